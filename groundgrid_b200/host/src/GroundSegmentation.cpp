// groundgrid::GroundSegmentation over the groundgrid_b200 C-ABI.
#include <groundgrid/GroundSegmentation.h>

#include <cmath>
#include <stdexcept>

using namespace groundgrid;

static_assert(sizeof(GroundSegmentation::PCLPoint) == sizeof(gg_point), "PointXYZIR and gg_point must share one layout");

void GroundSegmentation::init(ros::NodeHandle&, const size_t dimension, const float& resolution) {
    // The expectedPoints table of the reference (src/GroundSegmentation.cpp:37-48) is built inside
    // gg_create for the map's geometry; remember the arguments to cross-check them in filter_cloud.
    mDimension = dimension;
    mResolutionInit = resolution;
}

void GroundSegmentation::setConfig(const groundgrid::GroundGridConfig& config) { mConfig = config; }

pcl::PointCloud<GroundSegmentation::PCLPoint>::Ptr GroundSegmentation::filter_cloud(const pcl::PointCloud<PCLPoint>::Ptr cloud,
                                                                                      const PCLPoint& cloudOrigin,
                                                                                      const geometry_msgs::TransformStamped& mapToBase,
                                                                                      grid_map::GridMap& map) {
    gg_handle h = map.handle();
    if (!h) throw std::out_of_range("GridMap has no layers (initGroundGrid not called)");
    if (mDimension) {
        const size_t cells = (size_t)std::round(mDimension / mResolutionInit);
        if ((int)cells != map.getSize()(0)) throw std::runtime_error("GroundSegmentation::init geometry does not match the grid map");
    }
    gg_config c;
    c.point_count_cell_variance_threshold = mConfig.point_count_cell_variance_threshold;
    c.max_ring = mConfig.max_ring;
    c.groundpatch_detection_minimum_threshold = mConfig.groundpatch_detection_minimum_threshold;
    c.distance_factor = mConfig.distance_factor;
    c.minimum_distance_factor = mConfig.minimum_distance_factor;
    c.miminum_point_height_threshold = mConfig.miminum_point_height_threshold;
    c.minimum_point_height_obstacle_threshold = mConfig.minimum_point_height_obstacle_threshold;
    c.outlier_tolerance = mConfig.outlier_tolerance;
    c.ground_patch_detection_minimum_point_count_threshold = mConfig.ground_patch_detection_minimum_point_count_threshold;
    c.patch_size_change_distance = mConfig.patch_size_change_distance;
    c.occupied_cells_decrease_factor = mConfig.occupied_cells_decrease_factor;
    c.occupied_cells_point_count_factor = mConfig.occupied_cells_point_count_factor;
    c.min_outlier_detection_ground_confidence = mConfig.min_outlier_detection_ground_confidence;
    c.thread_count = mConfig.thread_count;
    if (gg_set_config(h, &c) != GG_OK) throw std::runtime_error(std::string("groundgrid_b200: ") + gg_last_error());

    // the six per-scan layers the reference adds to the map (:61-67,75)
    for (const char* l : {"groundCandidates", "planeDist", "m2", "meanVariance", "pointsRaw", "variance"}) map.add(l, 0.0);

    pcl::PointCloud<PCLPoint>::Ptr filtered_cloud(new pcl::PointCloud<PCLPoint>);
    const size_t n = cloud->points.size();
    filtered_cloud->points.resize(n);
    labels_.assign(n, 0);
    const float origin[3] = {cloudOrigin.x, cloudOrigin.y, cloudOrigin.z};
    // z of mapToBase * (0,0,0) (:405-411): the translation
    const double base_z = mapToBase.transform.translation.z;
    size_t n_out = 0;
    const int rc = gg_filter_cloud(h, map.slot(), reinterpret_cast<const gg_point*>(cloud->points.data()), n, origin, base_z, labels_.data(),
                                   nullptr, reinterpret_cast<gg_point*>(filtered_cloud->points.data()), &n_out);
    if (rc != GG_OK) throw std::runtime_error(std::string("groundgrid_b200: ") + gg_last_error());
    filtered_cloud->points.resize(n_out);
    filtered_cloud->width = (uint32_t)n_out;
    return filtered_cloud;
}

static void not_standalone(const char* what) {
    throw std::logic_error(std::string("groundgrid_b200: GroundSegmentation::") + what +
                           " is a thread entry point of the reference's CPU implementation; the GPU path runs it inside filter_cloud");
}
void GroundSegmentation::insert_cloud(const pcl::PointCloud<PCLPoint>::Ptr, const size_t, const size_t, const PCLPoint&,
                                      std::vector<std::pair<size_t, grid_map::Index>>&, std::vector<std::pair<size_t, grid_map::Index>>&,
                                      std::vector<size_t>&, grid_map::GridMap&) {
    not_standalone("insert_cloud");
}
void GroundSegmentation::detect_ground_patches(grid_map::GridMap&, unsigned short) const { not_standalone("detect_ground_patches"); }
template <int S> void GroundSegmentation::detect_ground_patch(grid_map::GridMap&, size_t, size_t) const { not_standalone("detect_ground_patch"); }
template void GroundSegmentation::detect_ground_patch<3>(grid_map::GridMap&, size_t, size_t) const;
template void GroundSegmentation::detect_ground_patch<5>(grid_map::GridMap&, size_t, size_t) const;
void GroundSegmentation::spiral_ground_interpolation(grid_map::GridMap&, const geometry_msgs::TransformStamped&) const {
    not_standalone("spiral_ground_interpolation");
}
void GroundSegmentation::interpolate_cell(grid_map::GridMap&, const size_t, const size_t) const { not_standalone("interpolate_cell"); }
