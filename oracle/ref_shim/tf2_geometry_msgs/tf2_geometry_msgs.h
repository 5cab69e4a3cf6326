// TEST INFRASTRUCTURE ONLY (oracle/_ref build): tf2::doTransform(PointStamped) as tf2_geometry_msgs 0.7.x (ROS Noetic)
// defines it -- tf2::Transform t; fromMsg(transform.transform, t); v_out = t * v_in -- restated from the published
// tf2 sources (tf2 is not installed here):
//   tf2::Matrix3x3::setRotation(q):  d = q.length2(); s = 2/d; xs = x*s ... ; rows as below
//   tf2::Transform::operator*(v):    (row0.dot(v) + origin.x, row1.dot(v) + origin.y, row2.dot(v) + origin.z),
//   tf2::Vector3::dot:               x*v.x + y*v.y + z*v.z   (left to right, fp64)
// Call sites: GroundGrid.cpp:129, GroundSegmentation.cpp:408.
#pragma once
#include <geometry_msgs/msgs.h>
#include <tf2_ros/transform_listener.h>

namespace tf2 {
// row-major 3x4 [R|t] of a TransformStamped
inline void ggr_to_matrix(const geometry_msgs::TransformStamped& t, double m[12]) {
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
inline void doTransform(const geometry_msgs::PointStamped& t_in, geometry_msgs::PointStamped& t_out, const geometry_msgs::TransformStamped& transform) {
    double m[12];
    ggr_to_matrix(transform, m);
    const double x = t_in.point.x, y = t_in.point.y, z = t_in.point.z;
    geometry_msgs::Point p;
    p.x = (m[0] * x + m[1] * y + m[2] * z) + m[3];
    p.y = (m[4] * x + m[5] * y + m[6] * z) + m[7];
    p.z = (m[8] * x + m[9] * y + m[10] * z) + m[11];
    t_out.point = p;
    t_out.header.stamp = transform.header.stamp;
    t_out.header.frame_id = transform.header.frame_id;
}
}  // namespace tf2
