// TEST INFRASTRUCTURE ONLY (oracle/_ref build): pcl::PointCloud<T> as the reference uses it -- `points` vector, Ptr.
#pragma once
#include <Eigen/Core>
#include <memory>
#include <std_msgs/Header.h>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud {
  public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std_msgs::Header header;
    std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
};
}  // namespace pcl
