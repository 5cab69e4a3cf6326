// groundgrid::GroundSegmentation -- same public surface as the reference class
// (include/groundgrid/GroundSegmentation.h:48-62 there).  filter_cloud runs the whole per-scan
// path on the GPU through gg_filter_cloud.  The per-phase methods of the reference work too, at the
// granularity the device path has:
//   insert_cloud(cloud, 0, cloud->size(), ..)  resets the per-scan layers (what filter_cloud does at :61-75) and
//       rasterises the WHOLE cloud; a sub-range [start, end) other than the whole cloud throws std::invalid_argument
//       (the reference's ranges are its thread chunks accumulating into shared matrices; the device path has no
//       partial-cloud accumulation).  The three index lists are filled like the reference's.
//   detect_ground_patches(map, section)  section 0 runs the detection for the whole map, sections 1-3 are no-ops
//       (the four sections are disjoint and order-free: their union is what one launch computes).
//   detect_ground_patch<S>, spiral_ground_interpolation, interpolate_cell  as in the reference.
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

    /** Not in the reference: filter_cloud fed with the raw sensor_msgs/PointCloud2 payload.  The unpack
     *  (pcl::fromROSMsg, GroundGridNodelet.cpp:119-120) and the per-point transform into the map frame (:148-184;
     *  T_map_from_frame = row-major 3x4 of lookupTransform("map", frame_id), null when the cloud is in "map" already)
     *  run on the device (gg_upload_cloud_msg); returns the segmented cloud like filter_cloud. */
    pcl::PointCloud<PCLPoint>::Ptr filter_cloud_msg(const sensor_msgs::PointCloud2& msg, const double* T_map_from_frame, const PCLPoint& cloudOrigin,
                                                    const geometry_msgs::TransformStamped& mapToBase, grid_map::GridMap& map);

  protected:
    groundgrid::GroundGridConfig mConfig;
    size_t mDimension = 0;
    float mResolutionInit = 0.f;
    std::vector<uint8_t> labels_;

  private:
    void pushConfig(grid_map::GridMap& map) const;
};
}  // namespace groundgrid
