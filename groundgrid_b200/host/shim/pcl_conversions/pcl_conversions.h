// pcl::fromROSMsg / toROSMsg for velodyne_pointcloud::PointXYZIR (field-offset driven, like PCL's
// generic converter): x, y, z, intensity float32 and ring uint16 at arbitrary offsets / point_step
// (the KITTI player publishes 18-byte points, scripts/kitti_data_publisher.py:139-150).
#pragma once
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
#include <cstring>
#include <stdexcept>
namespace pcl {
template <typename PointT>
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointT>& cloud) {
    int ox = -1, oy = -1, oz = -1, oi = -1, orr = -1;
    for (const auto& f : msg.fields) {
        if (f.name == "x") ox = (int)f.offset;
        else if (f.name == "y") oy = (int)f.offset;
        else if (f.name == "z") oz = (int)f.offset;
        else if (f.name == "intensity") oi = (int)f.offset;
        else if (f.name == "ring") orr = (int)f.offset;
    }
    if (ox < 0 || oy < 0 || oz < 0) throw std::runtime_error("fromROSMsg: x/y/z fields missing");
    const size_t n = (size_t)msg.width * msg.height;
    cloud.points.resize(n);
    cloud.width = msg.width;
    cloud.height = msg.height;
    cloud.header.frame_id = msg.header.frame_id;
    cloud.header.seq = msg.header.seq;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t* p = msg.data.data() + i * msg.point_step;
        PointT q;
        std::memset(&q, 0, sizeof(q));
        std::memcpy(&q.x, p + ox, 4);
        std::memcpy(&q.y, p + oy, 4);
        std::memcpy(&q.z, p + oz, 4);
        if (oi >= 0) std::memcpy(&q.intensity, p + oi, 4);
        if (orr >= 0) std::memcpy(&q.ring, p + orr, 2);
        cloud.points[i] = q;
    }
}
template <typename PointT>
void toROSMsg(const PointCloud<PointT>& cloud, sensor_msgs::PointCloud2& msg) {
    msg.height = 1;
    msg.width = (uint32_t)cloud.points.size();
    msg.point_step = sizeof(PointT);
    msg.row_step = msg.point_step * msg.width;
    msg.fields.clear();
    const char* names[5] = {"x", "y", "z", "intensity", "ring"};
    const uint32_t offs[5] = {0, 4, 8, 16, 20};
    for (int k = 0; k < 5; ++k) {
        sensor_msgs::PointField f;
        f.name = names[k];
        f.offset = offs[k];
        f.datatype = k < 4 ? sensor_msgs::PointField::FLOAT32 : sensor_msgs::PointField::UINT16;
        msg.fields.push_back(f);
    }
    msg.data.resize((size_t)msg.row_step);
    if (!cloud.points.empty()) std::memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
}
}  // namespace pcl
