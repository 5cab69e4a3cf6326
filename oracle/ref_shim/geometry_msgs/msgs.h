// TEST INFRASTRUCTURE ONLY (oracle/_ref build): plain structs with the field names of the ROS messages the
// reference touches (GroundGrid.cpp:53,65-66,90-93,116-131,139-140; GroundSegmentation.cpp:405-411).
#pragma once
#include <std_msgs/Header.h>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {0}; };
struct PoseWithCovarianceStamped { std_msgs::Header header; PoseWithCovariance pose; };
struct PointStamped { std_msgs::Header header; Point point; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
}  // namespace geometry_msgs
