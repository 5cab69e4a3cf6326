// TEST INFRASTRUCTURE ONLY (oracle/_ref build): tf2_ros::Buffer / TransformListener stand-ins.  GroundGrid owns a
// private Buffer + listener (GroundGrid.h:78-79) and asks it for "base_link" <- "map" (GroundGrid.cpp:102); here every
// Buffer answers from one process-wide table that the harness (oracle/ref_harness.cpp) fills before each update().
#pragma once
#include <geometry_msgs/msgs.h>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>

namespace tf2 {
struct TransformException : std::runtime_error { using std::runtime_error::runtime_error; };
struct LookupException : TransformException { using TransformException::TransformException; };
struct ConnectivityException : TransformException { using TransformException::TransformException; };
struct ExtrapolationException : TransformException { using TransformException::TransformException; };
}  // namespace tf2

namespace tf2_ros {
inline std::map<std::pair<std::string, std::string>, geometry_msgs::TransformStamped>& ggr_tf_table() {
    static std::map<std::pair<std::string, std::string>, geometry_msgs::TransformStamped> t;
    return t;
}
class Buffer {
  public:
    geometry_msgs::TransformStamped lookupTransform(const std::string& target, const std::string& source, const ros::Time&,
                                                    const ros::Duration& = ros::Duration(0.0)) const {
        auto it = ggr_tf_table().find({target, source});
        if (it == ggr_tf_table().end()) throw tf2::LookupException("\"" + target + "\" <- \"" + source + "\": no transform");
        return it->second;
    }
};
class TransformListener {
  public:
    explicit TransformListener(Buffer&) {}
};
}  // namespace tf2_ros
