// Nodelet callbacks over the GPU-backed GroundGrid / GroundSegmentation classes.
#include <groundgrid/GroundGridNodelet.h>

#include <chrono>

namespace groundgrid {

void GroundGridNodelet::onInit() {
    ros::NodeHandle nh = getNodeHandle();
    groundgrid_ = std::make_shared<GroundGrid>();
    if (dim_ > 0.f) groundgrid_->setGeometryOverride(dim_, res_, device_, max_points_);
    const float dim = dim_ > 0.f ? dim_ : groundgrid_->mDimension;
    const float res = res_ > 0.f ? res_ : groundgrid_->mResolution;
    ground_segmentation_.init(nh, (size_t)dim, res);
    GroundGridConfig defaults;
    callbackReconfigure(defaults, 0);  // the reconfigure server fires once with the defaults on setCallback
}

void GroundGridNodelet::odom_callback(const nav_msgs::OdometryConstPtr& inOdom) {
    auto start = std::chrono::steady_clock::now();
    map_ptr_ = groundgrid_->update(inOdom);
    auto end = std::chrono::steady_clock::now();
    ROS_DEBUG_STREAM("grid map update took " << std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() << "us");
}

void GroundGridNodelet::points_callback(const sensor_msgs::PointCloud2ConstPtr& cloud_msg) {
    auto start = std::chrono::steady_clock::now();
    pcl::PointCloud<PCLPoint>::Ptr cloud(new pcl::PointCloud<PCLPoint>);
    pcl::fromROSMsg(*cloud_msg, *cloud);
    geometry_msgs::TransformStamped mapToBaseTransform, cloudOriginTransform;

    // Map not initialized yet: no odometry message so far.
    if (!map_ptr_) return;

    try {
        mapToBaseTransform = mTfBuffer.lookupTransform("map", "base_link", cloud_msg->header.stamp, ros::Duration(0.0));
        cloudOriginTransform = mTfBuffer.lookupTransform("map", "velodyne", cloud_msg->header.stamp, ros::Duration(0.0));
    } catch (tf2::TransformException& ex) {
        ROS_WARN("Received point cloud but transforms are not available: %s", ex.what());
        return;  // the scan is dropped
    }

    geometry_msgs::PointStamped origin;
    origin.header = cloud_msg->header;
    origin.header.frame_id = "velodyne";
    tf2::doTransform(origin, origin, cloudOriginTransform);

    // Transform cloud into the map frame (per point, fp64, like tf2::doTransform)
    if (cloud_msg->header.frame_id != "map") {
        geometry_msgs::TransformStamped transformStamped;
        try {
            transformStamped = mTfBuffer.lookupTransform("map", cloud_msg->header.frame_id, cloud_msg->header.stamp, ros::Duration(0.0));
        } catch (tf2::TransformException& ex) {
            ROS_WARN("Failed to get map transform for point cloud transformation: %s", ex.what());
            return;
        }
        double m[12];
        tf2::toMatrix(transformStamped, m);
        pcl::PointCloud<PCLPoint>::Ptr transformed_cloud(new pcl::PointCloud<PCLPoint>);
        transformed_cloud->header = cloud->header;
        transformed_cloud->header.frame_id = "map";
        transformed_cloud->points.reserve(cloud->points.size());
        for (const auto& point : cloud->points) {
            const double x = point.x, y = point.y, z = point.z;
            PCLPoint p = point;
            p.x = (float)((m[0] * x + m[1] * y + m[2] * z) + m[3]);
            p.y = (float)((m[4] * x + m[5] * y + m[6] * z) + m[7]);
            p.z = (float)((m[8] * x + m[9] * y + m[10] * z) + m[11]);
            transformed_cloud->points.push_back(p);
        }
        cloud = transformed_cloud;
    }

    PCLPoint origin_pclPoint{};
    origin_pclPoint.x = (float)origin.point.x;
    origin_pclPoint.y = (float)origin.point.y;
    origin_pclPoint.z = (float)origin.point.z;
    sensor_msgs::PointCloud2 cloud_msg_out;
    pcl::toROSMsg(*(ground_segmentation_.filter_cloud(cloud, origin_pclPoint, mapToBaseTransform, *map_ptr_)), cloud_msg_out);
    cloud_msg_out.header = cloud_msg->header;
    cloud_msg_out.header.frame_id = "map";
    if (filtered_cloud_pub_) filtered_cloud_pub_(cloud_msg_out);
    if (grid_map_pub_) grid_map_pub_(map_ptr_, cloud_msg->header.stamp);
    auto end = std::chrono::steady_clock::now();
    ROS_DEBUG_STREAM("groundgrid took " << std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() << "us");
}

void GroundGridNodelet::callbackReconfigure(groundgrid::GroundGridConfig& config, uint32_t) {
    groundgrid_->setConfig(config);
    ground_segmentation_.setConfig(config);
}

}  // namespace groundgrid
