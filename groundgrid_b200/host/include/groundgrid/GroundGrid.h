// groundgrid::GroundGrid -- same public surface as the reference class
// (include/groundgrid/GroundGrid.h:50-71 there), backed by the groundgrid_b200 C-ABI: the
// grid map it hands out owns device-resident layers.
#pragma once
#include <grid_map_ros/grid_map_ros.hpp>
#include <grid_map_msgs/GridMap.h>

#include <geometry_msgs/PoseWithCovarianceStamped.h>
#include <nav_msgs/Odometry.h>

#include <tf2_ros/transform_listener.h>
#include <geometry_msgs/PointStamped.h>
#include <tf2_geometry_msgs/tf2_geometry_msgs.h>

#include <groundgrid/GroundGridConfig.h>

namespace groundgrid {

class GroundGrid {
  public:
    GroundGrid();
    virtual ~GroundGrid();

    /** Sets the current dynamic configuration. */
    void setConfig(groundgrid::GroundGridConfig& config);

    /** Creates the map around the first odometry pose (reference: src/GroundGrid.cpp:50-80). */
    void initGroundGrid(const nav_msgs::OdometryConstPtr& inOdom);
    /** Rolls the map to the new pose and seeds exposed cells (reference: src/GroundGrid.cpp:83-147). */
    std::shared_ptr<grid_map::GridMap> update(const nav_msgs::OdometryConstPtr& inOdom);

    const float mResolution = .33f;
    const float mDimension = 120.0f;

    /** Not in the reference: geometry / device of the map created by the next initGroundGrid
     *  (the reference hard-codes 120 m / 0.33 m, GroundGrid.h:70-71). */
    void setGeometryOverride(float dimension_m, float resolution, int device = 0, size_t max_points = 1u << 20);

  private:
    groundgrid::GroundGridConfig config_;
    tf2_ros::Buffer mTfBuffer;
    tf2_ros::TransformListener mTf2_listener;
    std::shared_ptr<grid_map::GridMap> mMap_ptr;
    geometry_msgs::PoseWithCovarianceStamped mLastPose;
    geometry_msgs::TransformStamped mBaseToMap;  // last good lookup ("static" in the reference, :100)
    float dim_override_ = 0.f, res_override_ = 0.f;
    int device_ = 0;
    size_t max_points_ = 1u << 20;
};
}  // namespace groundgrid
