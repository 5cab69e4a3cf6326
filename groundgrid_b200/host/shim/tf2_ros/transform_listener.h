// tf2 / tf2_ros stand-ins.  Buffer keeps the latest transform per (target, source) pair; a
// TransformListener registers its Buffer with a process-wide "topic" so that
// tf2_ros::shim_broadcast() plays the role of the /tf topic.  doTransform restates
// tf2::Transform * Vector3 (Matrix3x3::setRotation + row dot products, fp64).
#pragma once
#include <geometry_msgs/msgs.h>
#include <map>
#include <mutex>
#include <stdexcept>
#include <utility>
#include <vector>

namespace tf2 {
struct TransformException : std::runtime_error { using std::runtime_error::runtime_error; };
struct LookupException : TransformException { using TransformException::TransformException; };
struct ExtrapolationException : TransformException { using TransformException::TransformException; };

// Row-major 3x4 [R|t] of a TransformStamped (tf2::Matrix3x3::setRotation).
inline void toMatrix(const geometry_msgs::TransformStamped& t, double m[12]) {
    const geometry_msgs::Quaternion& q = t.transform.rotation;
    const double d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    const double s = 2.0 / d;
    const double xs = q.x * s, ys = q.y * s, zs = q.z * s;
    const double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    const double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    const double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    m[0] = 1.0 - (yy + zz); m[1] = xy - wz;         m[2] = xz + wy;          m[3] = t.transform.translation.x;
    m[4] = xy + wz;         m[5] = 1.0 - (xx + zz); m[6] = yz - wx;          m[7] = t.transform.translation.y;
    m[8] = xz - wy;         m[9] = yz + wx;         m[10] = 1.0 - (xx + yy); m[11] = t.transform.translation.z;
}
inline void doTransform(const geometry_msgs::PointStamped& in, geometry_msgs::PointStamped& out, const geometry_msgs::TransformStamped& t) {
    double m[12];
    toMatrix(t, m);
    const double x = in.point.x, y = in.point.y, z = in.point.z;
    geometry_msgs::PointStamped r;
    r.point.x = (m[0] * x + m[1] * y + m[2] * z) + m[3];
    r.point.y = (m[4] * x + m[5] * y + m[6] * z) + m[7];
    r.point.z = (m[8] * x + m[9] * y + m[10] * z) + m[11];
    r.header.stamp = t.header.stamp;
    r.header.frame_id = t.header.frame_id;
    out = r;
}
}  // namespace tf2

namespace tf2_ros {
class Buffer {
  public:
    bool setTransform(const geometry_msgs::TransformStamped& t, const std::string& = "shim", bool = false) {
        std::lock_guard<std::mutex> g(mu_);
        store_[{t.header.frame_id, t.child_frame_id}] = t;
        return true;
    }
    // target_frame <- source_frame; exact (target, source) pairs only (no chaining in the shim)
    geometry_msgs::TransformStamped lookupTransform(const std::string& target, const std::string& source, const ros::Time&,
                                                    const ros::Duration& = ros::Duration(0.0)) const {
        std::lock_guard<std::mutex> g(mu_);
        auto it = store_.find({target, source});
        if (it == store_.end()) throw tf2::LookupException("\"" + target + "\" <- \"" + source + "\": no transform");
        return it->second;
    }
    bool canTransform(const std::string& target, const std::string& source, const ros::Time&, const ros::Duration& = ros::Duration(0.0)) const {
        std::lock_guard<std::mutex> g(mu_);
        return store_.count({target, source}) != 0;
    }
  private:
    mutable std::mutex mu_;
    std::map<std::pair<std::string, std::string>, geometry_msgs::TransformStamped> store_;
};
inline std::vector<Buffer*>& shim_listeners() {
    static std::vector<Buffer*> v;
    return v;
}
class TransformListener {
  public:
    explicit TransformListener(Buffer& b) : b_(&b) { shim_listeners().push_back(b_); }
    ~TransformListener() {
        auto& v = shim_listeners();
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == b_) { v.erase(v.begin() + i); break; }
    }
  private:
    Buffer* b_;
};
// the "/tf topic": header.frame_id = target (parent), child_frame_id = source
inline void shim_broadcast(const geometry_msgs::TransformStamped& t) {
    for (Buffer* b : shim_listeners()) b->setTransform(t);
}
}  // namespace tf2_ros
