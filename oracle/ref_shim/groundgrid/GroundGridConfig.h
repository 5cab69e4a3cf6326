// TEST INFRASTRUCTURE ONLY (oracle/_ref build): what dynamic_reconfigure generates from the reference's
// cfg/GroundGrid.cfg:8-21 at catkin build time (CMakeLists.txt:47-49) -- a struct with one member per gen.add(),
// int_t -> int, double_t -> double, same names.  Defaults are the cfg file's.
#pragma once
namespace groundgrid {
class GroundGridConfig {
  public:
    int point_count_cell_variance_threshold = 10;
    int max_ring = 1024;
    double groundpatch_detection_minimum_threshold = 0.01;
    double distance_factor = 0.0001;
    double minimum_distance_factor = 0.0005;
    double miminum_point_height_threshold = 0.3;
    double minimum_point_height_obstacle_threshold = 0.1;
    double outlier_tolerance = 0.1;
    double ground_patch_detection_minimum_point_count_threshold = 0.25;
    double patch_size_change_distance = 20;
    double occupied_cells_decrease_factor = 5.0;
    double occupied_cells_point_count_factor = 20;
    double min_outlier_detection_ground_confidence = 1.25;
    int thread_count = 8;
};
}  // namespace groundgrid
