#pragma once
#include <geometry_msgs/msgs.h>
namespace nav_msgs {
struct Odometry {
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
};
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
typedef std::shared_ptr<Odometry> OdometryPtr;
}  // namespace nav_msgs
