// velodyne_pointcloud::PointXYZIR with the reference's memory layout
// (include/velodyne_pointcloud/point_types.h:27-33 there: PCL_ADD_POINT4D, float intensity,
// uint16 ring, EIGEN_ALIGN16 -> sizeof == 32).  Identical to gg_point of the C-ABI.
#pragma once
#include <cstdint>
namespace velodyne_pointcloud {
struct alignas(16) PointXYZIR {
    float x, y, z, data_pad;
    float intensity;
    std::uint16_t ring;
    std::uint16_t ring_pad;
    float tail_pad[2];
};
static_assert(sizeof(PointXYZIR) == 32, "PointXYZIR must match the reference's 32-byte record");
}  // namespace velodyne_pointcloud
