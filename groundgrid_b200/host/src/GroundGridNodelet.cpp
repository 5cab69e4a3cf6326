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

    // pcl::fromROSMsg (:119-120) and the per-point transform into the map frame (:148-184) run on the device:
    // the raw payload goes up as it is, together with the 3x4 matrix of lookupTransform("map", frame_id)
    double m[12];
    const double* T = nullptr;
    if (cloud_msg->header.frame_id != "map") {
        geometry_msgs::TransformStamped transformStamped;
        try {
            transformStamped = mTfBuffer.lookupTransform("map", cloud_msg->header.frame_id, cloud_msg->header.stamp, ros::Duration(0.0));
        } catch (tf2::TransformException& ex) {
            ROS_WARN("Failed to get map transform for point cloud transformation: %s", ex.what());
            return;
        }
        tf2::toMatrix(transformStamped, m);
        T = m;
    }

    PCLPoint origin_pclPoint{};
    origin_pclPoint.x = (float)origin.point.x;
    origin_pclPoint.y = (float)origin.point.y;
    origin_pclPoint.z = (float)origin.point.z;
    sensor_msgs::PointCloud2 cloud_msg_out;
    pcl::toROSMsg(*(ground_segmentation_.filter_cloud_msg(*cloud_msg, T, origin_pclPoint, mapToBaseTransform, *map_ptr_)), cloud_msg_out);
    cloud_msg_out.header = cloud_msg->header;
    cloud_msg_out.header.frame_id = "map";
    if (filtered_cloud_pub_) filtered_cloud_pub_(cloud_msg_out);
    if (grid_map_pub_) grid_map_pub_(map_ptr_, cloud_msg->header.stamp);
    // per-layer images (:219-228): only for the layers somebody subscribed to
    for (const auto& kv : layer_pubs_) publish_grid_map_layer(kv.second, kv.first);
    if (terrain_im_pub_) publish_grid_map_layer(terrain_im_pub_, "terrain");
    auto end = std::chrono::steady_clock::now();
    ROS_DEBUG_STREAM("groundgrid took " << std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() << "us");
}

// publish_grid_map_layer (:234-291).  Colour images: the device produces the 8UC1 image that
// grid_map::GridMapCvConverter::toImage hands to cv::applyColorMap; the 256-entry BGR table is OpenCV data
// (cv::COLORMAP_TWILIGHT) -- set it with setColorMap() (on a ROS machine: one cv::applyColorMap call on a 0..255 ramp);
// without a table the index image is replicated into the three channels.
void GroundGridNodelet::publish_grid_map_layer(const ImageSink& pub, const std::string& layer_name) {
    if (!pub || !map_ptr_) return;
    const int N = map_ptr_->getSize()(0);
    LayerImage img;
    img.rows = img.cols = N;
    if (layer_name != "terrain") {
        std::vector<uint8_t> idx((size_t)N * N);
        if (gg_layer_image_u8(map_ptr_->handle(), map_ptr_->slot(), layer_name.c_str(), idx.data(), nullptr, nullptr) != GG_OK) {
            ROS_WARN("layer image '%s': %s", layer_name.c_str(), gg_last_error());
            return;
        }
        img.encoding = "8UC3";
        img.data.resize(idx.size() * 3);
        for (size_t k = 0; k < idx.size(); ++k)
            for (int c = 0; c < 3; ++c) img.data[k * 3 + c] = have_colormap_ ? colormap_[idx[k]][c] : idx[k];
    } else {
        img.encoding = "32FC3";
        img.data.resize((size_t)N * N * 3 * sizeof(float));
        if (gg_terrain_image(map_ptr_->handle(), map_ptr_->slot(), reinterpret_cast<float*>(img.data.data())) != GG_OK) {
            ROS_WARN("terrain image: %s", gg_last_error());
            return;
        }
    }
    pub(img);
}

void GroundGridNodelet::callbackReconfigure(groundgrid::GroundGridConfig& config, uint32_t) {
    groundgrid_->setConfig(config);
    ground_segmentation_.setConfig(config);
}

}  // namespace groundgrid
