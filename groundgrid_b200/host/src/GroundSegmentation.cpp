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

namespace {
void check(int rc) {
    if (rc != GG_OK) throw std::runtime_error(std::string("groundgrid_b200: ") + gg_last_error());
}
gg_handle handle_of(grid_map::GridMap& map) {
    gg_handle h = map.handle();
    if (!h) throw std::out_of_range("GridMap has no layers (initGroundGrid not called)");
    return h;
}
}  // namespace

// The reference's setConfig is a struct copy (:468-471) and every phase reads mConfig when it runs.  Here the
// configuration lives on the device handle; it is pushed when it differs from what the handle has (a comparison, no
// device work in the steady state).
void GroundSegmentation::pushConfig(grid_map::GridMap& map) const {
    gg_handle h = handle_of(map);
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
    gg_config cur;
    check(gg_get_config(h, &cur));
    const bool same = cur.point_count_cell_variance_threshold == c.point_count_cell_variance_threshold && cur.max_ring == c.max_ring &&
                      cur.groundpatch_detection_minimum_threshold == c.groundpatch_detection_minimum_threshold &&
                      cur.distance_factor == c.distance_factor && cur.minimum_distance_factor == c.minimum_distance_factor &&
                      cur.miminum_point_height_threshold == c.miminum_point_height_threshold &&
                      cur.minimum_point_height_obstacle_threshold == c.minimum_point_height_obstacle_threshold &&
                      cur.outlier_tolerance == c.outlier_tolerance &&
                      cur.ground_patch_detection_minimum_point_count_threshold == c.ground_patch_detection_minimum_point_count_threshold &&
                      cur.patch_size_change_distance == c.patch_size_change_distance &&
                      cur.occupied_cells_decrease_factor == c.occupied_cells_decrease_factor &&
                      cur.occupied_cells_point_count_factor == c.occupied_cells_point_count_factor &&
                      cur.min_outlier_detection_ground_confidence == c.min_outlier_detection_ground_confidence && cur.thread_count == c.thread_count;
    if (!same) check(gg_set_config(h, &c));
    if (mDimension) {
        const size_t cells = (size_t)std::round(mDimension / mResolutionInit);
        if ((int)cells != map.getSize()(0)) throw std::runtime_error("GroundSegmentation::init geometry does not match the grid map");
    }
    // the six per-scan layers the reference adds to the map (:61-67,75)
    for (const char* l : {"groundCandidates", "planeDist", "m2", "meanVariance", "pointsRaw", "variance"}) map.add(l, 0.0);
}

pcl::PointCloud<GroundSegmentation::PCLPoint>::Ptr GroundSegmentation::filter_cloud(const pcl::PointCloud<PCLPoint>::Ptr cloud,
                                                                                      const PCLPoint& cloudOrigin,
                                                                                      const geometry_msgs::TransformStamped& mapToBase,
                                                                                      grid_map::GridMap& map) {
    gg_handle h = handle_of(map);
    pushConfig(map);
    pcl::PointCloud<PCLPoint>::Ptr filtered_cloud(new pcl::PointCloud<PCLPoint>);
    const size_t n = cloud->points.size();
    filtered_cloud->points.resize(n);
    labels_.assign(n, 0);
    const float origin[3] = {cloudOrigin.x, cloudOrigin.y, cloudOrigin.z};
    // z of mapToBase * (0,0,0) (:405-411): the translation
    const double base_z = mapToBase.transform.translation.z;
    size_t n_out = 0;
    check(gg_filter_cloud(h, map.slot(), reinterpret_cast<const gg_point*>(cloud->points.data()), n, origin, base_z, labels_.data(), nullptr,
                          reinterpret_cast<gg_point*>(filtered_cloud->points.data()), &n_out));
    filtered_cloud->points.resize(n_out);
    filtered_cloud->width = (uint32_t)n_out;
    return filtered_cloud;
}

pcl::PointCloud<GroundSegmentation::PCLPoint>::Ptr GroundSegmentation::filter_cloud_msg(const sensor_msgs::PointCloud2& msg, const double* T_map_from_frame,
                                                                                          const PCLPoint& cloudOrigin,
                                                                                          const geometry_msgs::TransformStamped& mapToBase,
                                                                                          grid_map::GridMap& map) {
    gg_handle h = handle_of(map);
    pushConfig(map);
    int off[5] = {-1, -1, -1, -1, -1};
    for (const auto& f : msg.fields) {
        if (f.name == "x") off[0] = (int)f.offset;
        else if (f.name == "y") off[1] = (int)f.offset;
        else if (f.name == "z") off[2] = (int)f.offset;
        else if (f.name == "intensity") off[3] = (int)f.offset;
        else if (f.name == "ring") off[4] = (int)f.offset;
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0) throw std::runtime_error("PointCloud2: x/y/z fields missing");
    const size_t n = (size_t)msg.width * msg.height;
    check(gg_upload_cloud_msg(h, map.slot(), msg.data.data(), n, (int)msg.point_step, off, T_map_from_frame));
    gg_scan_desc d{};
    d.slot = map.slot();
    d.n_points = n;
    d.origin[0] = cloudOrigin.x;
    d.origin[1] = cloudOrigin.y;
    d.origin[2] = cloudOrigin.z;
    d.base_z = mapToBase.transform.translation.z;
    check(gg_run_scans(h, 1, &d, 0));
    labels_.assign(n, 0);
    check(gg_download_labels(h, map.slot(), labels_.data(), n));
    pcl::PointCloud<PCLPoint>::Ptr filtered_cloud(new pcl::PointCloud<PCLPoint>);
    filtered_cloud->points.resize(n);
    size_t n_out = 0;
    check(gg_get_output(h, map.slot(), nullptr, reinterpret_cast<gg_point*>(filtered_cloud->points.data()), &n_out));  // synchronises
    filtered_cloud->points.resize(n_out);
    filtered_cloud->width = (uint32_t)n_out;
    return filtered_cloud;
}

// insert_cloud (:200-311) over the whole cloud, preceded by the per-scan layer reset of filter_cloud (:61-75)
void GroundSegmentation::insert_cloud(const pcl::PointCloud<PCLPoint>::Ptr cloud, const size_t start, const size_t end, const PCLPoint& cloudOrigin,
                                      std::vector<std::pair<size_t, grid_map::Index>>& point_index,
                                      std::vector<std::pair<size_t, grid_map::Index>>& ignored, std::vector<size_t>& outliers, grid_map::GridMap& map) {
    gg_handle h = handle_of(map);
    const size_t n = cloud->points.size();
    if (start != 0 || end != n)
        throw std::invalid_argument("groundgrid_b200: insert_cloud rasterises the whole cloud (start = 0, end = cloud size); see GroundSegmentation.h");
    pushConfig(map);
    check(gg_upload_points(h, map.slot(), reinterpret_cast<const gg_point*>(cloud->points.data()), n));
    gg_scan_desc d{};
    d.slot = map.slot();
    d.n_points = n;
    d.origin[0] = cloudOrigin.x;
    d.origin[1] = cloudOrigin.y;
    d.origin[2] = cloudOrigin.z;
    check(gg_run_scans(h, 1, &d, 1));
    std::vector<uint32_t> codes(n);
    check(gg_get_point_classes(h, map.slot(), codes.data(), n));
    const int N = map.getSize()(0);
    point_index.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t cls = codes[i] >> 24;
        const int cell = (int)(codes[i] & 0xffffffu);
        const grid_map::Index gi(cell % N, cell / N);
        if (cls == 1 || cls == 2) point_index.push_back(std::make_pair(i, gi));
        else if (cls == 3 || cls == 4) ignored.push_back(std::make_pair(i, gi));
        else if (cls == 5) outliers.push_back(i);
    }
}
void GroundSegmentation::detect_ground_patches(grid_map::GridMap& map, unsigned short section) const {
    if (section != 0) return;  // the launch of section 0 covers the union of the four sections
    check(gg_detect_ground_patches(handle_of(map), map.slot()));
}
template <int S> void GroundSegmentation::detect_ground_patch(grid_map::GridMap& map, size_t i, size_t j) const {
    check(gg_detect_ground_patch(handle_of(map), map.slot(), S, (int)i, (int)j));
}
template void GroundSegmentation::detect_ground_patch<3>(grid_map::GridMap&, size_t, size_t) const;
template void GroundSegmentation::detect_ground_patch<5>(grid_map::GridMap&, size_t, size_t) const;
void GroundSegmentation::spiral_ground_interpolation(grid_map::GridMap& map, const geometry_msgs::TransformStamped& toBase) const {
    check(gg_spiral_ground_interpolation(handle_of(map), map.slot(), toBase.transform.translation.z));
}
void GroundSegmentation::interpolate_cell(grid_map::GridMap& map, const size_t x, const size_t y) const {
    check(gg_interpolate_cell(handle_of(map), map.slot(), (int)x, (int)y));
}
