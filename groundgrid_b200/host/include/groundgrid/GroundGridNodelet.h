// groundgrid::GroundGridNodelet -- the callback contract of the reference's nodelet
// (src/GroundGridNodelet.cpp:78-104 onInit, :107-112 odom_callback, :114-232 points_callback,
// :299-302 callbackReconfigure) without the ROS transport: subscriptions become public
// callbacks, publishers become std::function sinks (segmented cloud, grid map, per-layer colour images and the
// "terrain" image of publish_grid_map_layer, :219-228,234-291).  The ROS message transport itself is out of scope
// (SURVEY.md section 2).
#pragma once
#include <array>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include <nodelet/nodelet.h>
#include <ros/ros.h>
#include <nav_msgs/Odometry.h>
#include <pcl_conversions/pcl_conversions.h>
#include <velodyne_pointcloud/point_types.h>
#include <pcl_ros/point_cloud.h>
#include <tf2_ros/transform_listener.h>

#include <groundgrid/GroundGrid.h>
#include <groundgrid/GroundGridConfig.h>
#include <groundgrid/GroundGridFwd.h>
#include <groundgrid/GroundSegmentation.h>

namespace groundgrid {

class GroundGridNodelet : public nodelet::Nodelet {
  public:
    typedef velodyne_pointcloud::PointXYZIR PCLPoint;
    GroundGridNodelet() : mTfListener(mTfBuffer) {}
    virtual ~GroundGridNodelet() {}

    void onInit() override;
    /** "/localization/odometry/filtered_map" */
    virtual void odom_callback(const nav_msgs::OdometryConstPtr& inOdom);
    /** "/sensors/velodyne_points" */
    virtual void points_callback(const sensor_msgs::PointCloud2ConstPtr& cloud_msg);
    /** dynamic_reconfigure callback */
    void callbackReconfigure(groundgrid::GroundGridConfig& config, uint32_t level);

    /** publisher sinks: "groundgrid/segmented_cloud", "groundgrid/grid_map" */
    std::function<void(const sensor_msgs::PointCloud2&)> filtered_cloud_pub_;
    std::function<void(const std::shared_ptr<grid_map::GridMap>&, const ros::Time&)> grid_map_pub_;

    /** image sinks: "/groundgrid/grid_map_cv_<layer>" (8UC3) and "groundgrid/terrain" (32FC3), :219-228 */
    struct LayerImage {
        int rows = 0, cols = 0;
        std::string encoding;        // "8UC3" | "32FC3"
        std::vector<uint8_t> data;   // row-major, rows x cols x 3 channels
    };
    typedef std::function<void(const LayerImage&)> ImageSink;
    std::map<std::string, ImageSink> layer_pubs_;
    ImageSink terrain_im_pub_;
    /** the 256 BGR triples of cv::COLORMAP_TWILIGHT (OpenCV data; not shipped here) */
    void setColorMap(const std::array<std::array<uint8_t, 3>, 256>& bgr) { colormap_ = bgr; have_colormap_ = true; }
    void publish_grid_map_layer(const ImageSink& pub, const std::string& layer_name);

    /** geometry / device of the map (forwarded to GroundGrid::setGeometryOverride; call before onInit) */
    void setGeometryOverride(float dimension_m, float resolution, int device = 0, size_t max_points = 1u << 20) {
        dim_ = dimension_m; res_ = resolution; device_ = device; max_points_ = max_points;
    }
    GroundSegmentation& segmentation() { return ground_segmentation_; }
    std::shared_ptr<grid_map::GridMap> map() const { return map_ptr_; }

  private:
    GroundGridPtr groundgrid_;
    std::shared_ptr<grid_map::GridMap> map_ptr_;
    GroundSegmentation ground_segmentation_;
    tf2_ros::Buffer mTfBuffer;
    tf2_ros::TransformListener mTfListener;
    std::array<std::array<uint8_t, 3>, 256> colormap_{};
    bool have_colormap_ = false;
    float dim_ = 0.f, res_ = 0.f;
    int device_ = 0;
    size_t max_points_ = 1u << 20;
};
}  // namespace groundgrid
