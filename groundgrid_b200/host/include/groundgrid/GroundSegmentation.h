// groundgrid::GroundSegmentation -- same public surface as the reference class
// (include/groundgrid/GroundSegmentation.h:48-62 there).  filter_cloud runs the whole per-scan
// path on the GPU through gg_filter_cloud; the per-phase helper methods of the reference
// (insert_cloud, detect_ground_patches, detect_ground_patch, spiral_ground_interpolation,
// interpolate_cell) are thread entry points of its CPU implementation and have no stand-alone
// meaning here: they are declared for source compatibility and throw std::logic_error.
#pragma once
#include <sensor_msgs/PointCloud2.h>
#include <geometry_msgs/TransformStamped.h>

#include <pcl_ros/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include "velodyne_pointcloud/point_types.h"

#include <grid_map_ros/grid_map_ros.hpp>
#include <grid_map_msgs/GridMap.h>
#include <grid_map_cv/GridMapCvConverter.hpp>

#include <groundgrid/GroundGridConfig.h>

namespace groundgrid {
class GroundSegmentation {
  public:
    typedef velodyne_pointcloud::PointXYZIR PCLPoint;

    GroundSegmentation() {}
    void init(ros::NodeHandle& nodeHandle, const size_t dimension, const float& resolution);
    pcl::PointCloud<PCLPoint>::Ptr filter_cloud(const pcl::PointCloud<PCLPoint>::Ptr cloud, const PCLPoint& cloudOrigin,
                                                const geometry_msgs::TransformStamped& mapToBase, grid_map::GridMap& map);
    void insert_cloud(const pcl::PointCloud<PCLPoint>::Ptr cloud, const size_t start, const size_t end, const PCLPoint& cloudOrigin,
                      std::vector<std::pair<size_t, grid_map::Index>>& point_index, std::vector<std::pair<size_t, grid_map::Index>>& ignored,
                      std::vector<size_t>& outliers, grid_map::GridMap& map);
    void setConfig(const groundgrid::GroundGridConfig& config);
    void detect_ground_patches(grid_map::GridMap& map, unsigned short section) const;
    template <int S> void detect_ground_patch(grid_map::GridMap& map, size_t i, size_t j) const;
    void spiral_ground_interpolation(grid_map::GridMap& map, const geometry_msgs::TransformStamped& toBase) const;
    void interpolate_cell(grid_map::GridMap& map, const size_t x, const size_t y) const;

    /** Not in the reference: per-input-point labels (0 absent / 49 / 99) of the last filter_cloud call. */
    const std::vector<uint8_t>& lastLabels() const { return labels_; }

  protected:
    groundgrid::GroundGridConfig mConfig;
    size_t mDimension = 0;
    float mResolutionInit = 0.f;
    std::vector<uint8_t> labels_;
};
}  // namespace groundgrid
