// TEST INFRASTRUCTURE ONLY (oracle/_ref build): nav_msgs::Odometry fields read by GroundGrid.cpp:58,65-66,72,91-97,102,117.
#pragma once
#include <geometry_msgs/msgs.h>
#include <memory>
namespace nav_msgs {
struct Odometry {
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
};
typedef std::shared_ptr<Odometry> OdometryPtr;
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
}  // namespace nav_msgs
