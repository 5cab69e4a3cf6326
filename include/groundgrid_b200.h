/*
 * groundgrid_b200 -- C-ABI of the B200-native GroundGrid per-scan hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The host
 * classes in groundgrid_b200/host (groundgrid::GroundGrid / groundgrid::GroundSegmentation,
 * same surface as the reference headers) are thin wrappers over these calls; a maintainer
 * of the reference binds them exactly the same way.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * dcmlr/groundgrid repository).  All functions return 0 on success or a negative GG_E_* code;
 * no exception crosses this boundary.  A handle is not thread-safe (same contract as the
 * reference: callbacks are serialised on one ROS queue, src/GroundGridNodelet.cpp:79).
 *
 * There is NO CPU fallback: every compute entry point fails with GG_E_CUDA when no sm_100
 * device is usable.
 *
 * Data layout at the boundary:
 *   - points: 32-byte PointXYZIR records (include/velodyne_pointcloud/point_types.h:27-33)
 *   - layers: N x N float, column-major like Eigen::MatrixXf: element (i, j) at i + j*N,
 *     row index i grows toward -x, column index j toward -y (grid_map convention).
 *   - one handle owns `n_slots` independent maps ("slots"); slot s is what one
 *     GroundGrid + GroundSegmentation pair owns in the reference.  n_slots == 1 is the
 *     plain drop-in; n_slots > 1 is the batched throughput mode (independent scans).
 */
#ifndef GROUNDGRID_B200_H
#define GROUNDGRID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_OK 0
#define GG_E_ARG (-1)      /* bad argument (null pointer, slot out of range, too many points ...) */
#define GG_E_CUDA (-2)     /* CUDA runtime error or no usable device; see gg_last_error() */
#define GG_E_STATE (-3)    /* map not initialised (reference: points_callback returns early, GroundGridNodelet.cpp:124-125) */
#define GG_E_LAYER (-4)    /* unknown layer name (reference: std::out_of_range from grid_map) */

#define GG_LABEL_ABSENT 0      /* point is not part of the output cloud (outside map / NaN / border cell) */
#define GG_LABEL_GROUND 49     /* src/GroundSegmentation.cpp:180,188 */
#define GG_LABEL_NONGROUND 99  /* src/GroundSegmentation.cpp:175 */

/* gg_create flags */
#define GG_FLAG_FULL_LAYERS 1u /* also maintain the layers the algorithm never reads back
                                  (groundCandidates, planeDist, m2, meanVariance, pointsRaw,
                                  maxGroundHeight; src/GroundSegmentation.cpp:61-67,234,296,303,307) */

/* include/velodyne_pointcloud/point_types.h:27-33 -- velodyne_pointcloud::PointXYZIR, sizeof == 32 */
typedef struct gg_point {
    float x, y, z, _pad0;
    float intensity;
    uint16_t ring;
    uint16_t _pad1;
    float _pad2[2];
} gg_point;

/* cfg/GroundGrid.cfg:8-21 -- groundgrid::GroundGridConfig (same order, same defaults) */
typedef struct gg_config {
    int point_count_cell_variance_threshold;                     /* 10 */
    int max_ring;                                                /* 1024 */
    double groundpatch_detection_minimum_threshold;              /* 0.01 (declared, never read) */
    double distance_factor;                                      /* 0.0001 */
    double minimum_distance_factor;                              /* 0.0005 */
    double miminum_point_height_threshold;                       /* 0.3 */
    double minimum_point_height_obstacle_threshold;              /* 0.1 */
    double outlier_tolerance;                                    /* 0.1 */
    double ground_patch_detection_minimum_point_count_threshold; /* 0.25 */
    double patch_size_change_distance;                           /* 20 */
    double occupied_cells_decrease_factor;                       /* 5 */
    double occupied_cells_point_count_factor;                    /* 20 */
    double min_outlier_detection_ground_confidence;              /* 1.25 */
    int thread_count;                                            /* 8 (accepted, unused: no host threads) */
} gg_config;

typedef struct gg_handle_s* gg_handle;

/* Per-scan inputs of one slot (arguments of GroundSegmentation::filter_cloud,
 * include/groundgrid/GroundSegmentation.h:55):
 *   origin  = cloudOrigin (x, y, z of the sensor in the map frame, GroundGridNodelet.cpp:139-146,190-194)
 *   base_z  = z of mapToBase * (0,0,0) (GroundSegmentation.cpp:405-411)                        */
typedef struct gg_scan_desc {
    int slot;
    int _reserved;
    size_t n_points;
    float origin[3];
    float _pad;
    double base_z;
} gg_scan_desc;

/* Fills *cfg with the defaults of cfg/GroundGrid.cfg:8-21. */
void gg_default_config(gg_config* cfg);

/* Text of the last error on this thread (CUDA error string or argument complaint). */
const char* gg_last_error(void);

/* Replaces GroundSegmentation::init (src/GroundSegmentation.cpp:37-48: builds the expectedPoints
 * table for round(dimension/resolution)^2 cells) plus the device-side allocation of `n_slots`
 * maps.  `max_points` is the per-scan capacity.  `stream` is a cudaStream_t (NULL = the handle
 * creates its own non-blocking streams).  `device` is the CUDA ordinal. */
int gg_create(double dimension_m, float resolution, int device, int n_slots, size_t max_points,
              unsigned flags, void* stream, gg_handle* out);
int gg_destroy(gg_handle h);

int gg_cells_per_side(gg_handle h);  /* N (364 for the reference default 120 m / 0.33 m) */
int gg_num_slots(gg_handle h);

/* Replaces GroundGrid::setConfig + GroundSegmentation::setConfig
 * (src/GroundGrid.cpp:48, src/GroundSegmentation.cpp:468-471; GroundGridNodelet.cpp:299-302). */
int gg_set_config(gg_handle h, const gg_config* cfg);
int gg_get_config(gg_handle h, gg_config* cfg);

/* Replaces GroundGrid::initGroundGrid (src/GroundGrid.cpp:50-80): map centred on (x, y),
 * ground = z, groundpatch = 1e-7, points = 0, min = 100, max = -100. */
int gg_init_map(gg_handle h, int slot, double x, double y, double z);

/* Replaces GroundGrid::update for an already initialised map (src/GroundGrid.cpp:83-147):
 * rolls the map to the odometry position (x, y) by whole cells, seeds exposed cells with
 * ground = -(T * (cx, cy, 0)).z and groundpatch = 0 where T = row-major 3x4 [R|t] of
 * lookupTransform("base_link", "map").  *moved (may be NULL) = 1 if a cell shift happened. */
int gg_update_pose(gg_handle h, int slot, double x, double y, const double T_base_from_map[12], int* moved);

/* Batched gg_update_pose: `count` distinct slots (a repeated slot is GG_E_ARG), xy = 2 doubles per slot, T = 12 doubles per
 * slot, moved (may be NULL) = 1 int per slot.  One roll launch covers all slots. */
int gg_update_pose_batch(gg_handle h, int count, const int* slots, const double* xy, const double* T, int* moved);

/* Current map centre position (grid_map::GridMap::getPosition()). */
int gg_get_map_position(gg_handle h, int slot, double xy[2]);

/* Replaces GroundSegmentation::filter_cloud (src/GroundSegmentation.cpp:50-197) for one slot with
 * HOST buffers: copies the cloud to the device, runs rasterise -> patch detection -> spiral
 * interpolation -> labelling, copies results back and returns when they are in host memory.
 *   labels_out : n bytes, per INPUT point: 0 absent / 49 ground / 99 non-ground   (may be NULL)
 *   index_out  : n_out uint32, input index of each output point in the reference's output
 *                order: kept, then ignored, then outliers (:112-117,150,185)        (may be NULL)
 *   cloud_out  : n_out records, the reference's returned cloud (intensity = 49/99)  (may be NULL)
 *   n_out      : number of points in the output cloud                              (may be NULL) */
int gg_filter_cloud(gg_handle h, int slot, const gg_point* points, size_t n, const float origin[3],
                    double base_z, uint8_t* labels_out, uint32_t* index_out, gg_point* cloud_out,
                    size_t* n_out);

/* Batched form of gg_filter_cloud: `count` independent scans (distinct slots), host buffers.
 * points[k] / labels_out[k] are per-scan host pointers (pinned memory makes the copies
 * asynchronous).  Copies and kernels of different scans overlap on internal streams.
 * PCIe is the bottleneck of this path.  Host worker threads (GG_HOST_THREADS; default: the usable
 * CPUs of the process minus one) repack clouds into pinned staging memory as x | y | z | ring
 * (14 of the 32 bytes of a PointXYZIR record are used by the algorithm) so that only those bytes
 * cross the bus; whenever the packers fall behind the bus (fewer than two packed clouds queued for
 * copying), the calling thread sends scans from the back of the batch, which no packer has reached
 * yet, as plain 32-byte records.  GG_HOST_PACK=1 packs every scan, GG_HOST_PACK=0 none.  Results
 * are identical in all modes. */
int gg_filter_cloud_batch(gg_handle h, int count, const gg_scan_desc* scans, const gg_point* const* points,
                          uint8_t* const* labels_out);

/* The same call in two halves, for callers that keep the bus busy across batches: _begin returns as
 * soon as every cloud has been handed to the copy engines and every kernel is enqueued; the labels of
 * that batch are complete after _wait(ticket) (or gg_synchronize).  At most two batches are in
 * flight: _begin first waits for the batch before the previous one.  points[k] and labels_out[k] of a
 * batch must stay untouched until its _wait; gg_update_pose_batch for the next scans may be called
 * right after _begin (stream order keeps roll -> scan -> roll -> scan per slot).  Other calls that
 * touch the same slots need a gg_synchronize first. */
int gg_filter_cloud_batch_begin(gg_handle h, int count, const gg_scan_desc* scans, const gg_point* const* points,
                                uint8_t* const* labels_out, int* ticket);
int gg_filter_cloud_batch_wait(gg_handle h, int ticket);

/* Device-resident pipeline pieces (what gg_filter_cloud_batch is made of; used by bench.py to time
 * the kernels with the inputs already in HBM, and by the tests to check single phases):
 *   gg_upload_points : async H2D of one scan into its slot
 *   gg_run_scans     : enqueue the kernels for `count` scans whose points are resident.
 *                      stop_after: 0 whole path, 1 after rasterisation, 2 after patch detection,
 *                      3 after spiral interpolation
 *   gg_download_labels : async D2H of the per-input-point labels of one slot
 *   gg_synchronize   : wait for everything enqueued on the handle                              */
int gg_upload_points(gg_handle h, int slot, const gg_point* points, size_t n);
int gg_run_scans(gg_handle h, int count, const gg_scan_desc* scans, int stop_after);
int gg_download_labels(gg_handle h, int slot, uint8_t* labels_out, size_t n);
int gg_synchronize(gg_handle h);

/* The per-phase methods of the reference class, for callers that drive the phases themselves
 * (include/groundgrid/GroundSegmentation.h:56-62; all enqueue on the slot's stream):
 *   gg_run_scans(.., stop_after = 1)  = the layer reset of filter_cloud (:61-75) + insert_cloud over the whole cloud
 *                                       (:200-311); gg_get_point_classes then returns, per input point,
 *                                       class << 24 | cell with class 0 absent, 1 kept, 2 kept (border cell),
 *                                       3 ignored, 4 ignored (border cell), 5 outlier -- the contents of the
 *                                       point_index / ignored / outliers lists of :112-117
 *   gg_detect_ground_patches          = detect_ground_patches for all four sections (:314-340; sections are disjoint
 *                                       and every cell only writes itself, so their union is order-free)
 *   gg_detect_ground_patch            = detect_ground_patch<patch_size>(map, i, j) for one cell (:343-395)
 *   gg_spiral_ground_interpolation    = spiral_ground_interpolation (:398-441), base_z = z of toBase * (0,0,0)
 *   gg_interpolate_cell               = interpolate_cell(map, x, y) (:445-465)                                   */
int gg_get_point_classes(gg_handle h, int slot, uint32_t* codes, size_t n);
int gg_detect_ground_patches(gg_handle h, int slot);
int gg_detect_ground_patch(gg_handle h, int slot, int patch_size, int i, int j);
int gg_spiral_ground_interpolation(gg_handle h, int slot, double base_z);
int gg_interpolate_cell(gg_handle h, int slot, int x, int y);

/* Like gg_run_scans, but scan k reads its cloud from the caller-owned DEVICE buffer dev_points[k]
 * (n_points 32-byte records, 16-byte aligned) instead of the slot's upload buffer.  The buffer
 * must stay valid until the work is complete (and until gg_get_output, if that is used). */
int gg_run_scans_device(gg_handle h, int count, const gg_scan_desc* scans, const gg_point* const* dev_points, int stop_after);

/* ---- steps next to the path -------------------------------------------
 * gg_upload_cloud_msg replaces pcl::fromROSMsg + the per-point tf2::doTransform loop of
 * GroundGridNodelet::points_callback (src/GroundGridNodelet.cpp:119-120,148-184): the raw
 * sensor_msgs/PointCloud2 payload (`point_step` bytes per point, fields x, y, z, intensity float32 and
 * ring uint16 at `field_offsets`, -1 = absent; e.g. {0,4,8,12,16} for the KITTI player's 18-byte points,
 * scripts/kitti_data_publisher.py:139-150) is copied to the device, unpacked into PointXYZIR records
 * and -- unless T is NULL (frame_id == "map") -- transformed with the row-major 3x4 T = lookupTransform
 * ("map", frame_id) in fp64.  The records land in the slot's cloud buffer: follow with gg_run_scans.
 *
 * gg_terrain_image replaces the "terrain" branch of publish_grid_map_layer (:247-270): N*N*3 floats,
 * pixel (i, j) = (ground, 3x3 pointsRaw sum >= 27 ? 1 : 0, pointsRaw); needs GG_FLAG_FULL_LAYERS.
 *
 * gg_eval_accumulate / gg_eval_read replace the tallies of scripts/eval_groundpoint_classifier.py:95-118:
 * counts[id][0] / counts[id][1] = points of ground-truth label `id` (taken from `ring`) predicted ground /
 * non-ground, accumulated over the scans it was called for (1024 ids). */
int gg_upload_cloud_msg(gg_handle h, int slot, const void* data, size_t n_points, int point_step, const int field_offsets[5],
                        const double T_map_from_frame[12]);
int gg_terrain_image(gg_handle h, int slot, float* dst);
/* The other branch of publish_grid_map_layer (src/GroundGridNodelet.cpp:238-245): the single-channel 8-bit image that
 * grid_map::GridMapCvConverter::toImage<unsigned char, 1>(map, layer, CV_8UC1, img) produces and cv::applyColorMap then
 * colours -- lower / upper = min / max over the finite cells, pixel (i, j) = (uchar)((v - lower) / (upper - lower) * 255),
 * non-finite cells 0.  dst: N*N bytes, row-major (i, j) like the cv::Mat; lower / upper (may be NULL) receive the range.
 * The colour table itself (cv::COLORMAP_TWILIGHT, 256 BGR triples) is OpenCV data: the caller applies it. */
int gg_layer_image_u8(gg_handle h, int slot, const char* name, uint8_t* dst, float* lower, float* upper);
int gg_eval_accumulate(gg_handle h, int slot);
int gg_eval_read(gg_handle h, uint64_t* counts, int reset);

/* Per-kernel CUDA-event timing on the launching stream (bench.py roofline).  While enabled every
 * kernel launch is bracketed by an event pair; gg_profile_read synchronises and returns the
 * accumulated milliseconds and launch counts per kernel id (arrays of gg_profile_kernel_count()). */
int gg_profile_enable(gg_handle h, int on);
int gg_profile_read(gg_handle h, double* ms_per_kernel, uint32_t* launches_per_kernel, int reset);
int gg_profile_kernel_count(void);
const char* gg_profile_kernel_name(int id);

/* Output cloud of the last scan of a slot in the reference's order (see gg_filter_cloud). */
int gg_get_output(gg_handle h, int slot, uint32_t* index_out, gg_point* cloud_out, size_t* n_out);

/* Layer access (grid_map::GridMap::operator[] / get(), e.g. GroundSegmentation.cpp:76-78):
 * names "points", "ground", "groundpatch", "minGroundHeight", "maxGroundHeight", "variance"
 * always; "groundCandidates", "planeDist", "m2", "meanVariance", "pointsRaw" with
 * GG_FLAG_FULL_LAYERS.  dst/src: N*N floats, column-major.  "expectedPoints" reads the
 * table of GroundSegmentation::init. */
int gg_get_layer(gg_handle h, int slot, const char* name, float* dst);
int gg_set_layer(gg_handle h, int slot, const char* name, const float* src);

/* Raw device pointer of a layer (for NCCL broadcast of the rolling terrain prior: "ground" and
 * "groundpatch" of one slot are contiguous, 2*N*N floats starting at "ground"). */
int gg_layer_device_ptr(gg_handle h, int slot, const char* name, void** dptr);
int gg_set_map_position(gg_handle h, int slot, double x, double y);

/* Streams.  Slots are bound to the handle's streams in contiguous groups (GG_STREAMS env,
 * default 4, capped by n_slots; 1 when the caller supplied a stream) and everything that
 * touches a slot is enqueued on its stream.  gg_stream() is the primary stream;
 * gg_fork_streams() makes all streams wait for work already enqueued on it and
 * gg_join_streams() makes it wait for all others, so an event pair recorded on gg_stream()
 * around fork ... join brackets the work of every stream. */
void* gg_stream(gg_handle h);
int gg_num_streams(gg_handle h);
int gg_host_pack_threads(gg_handle h);
/* How the last gg_filter_cloud_batch call moved its clouds: info[0] scans repacked on the host,
 * info[1] scans sent as 32-byte records, info[2] / info[3] the H2D bytes of either kind,
 * info[4] host microseconds until the last cloud was enqueued, info[5] until the call returned
 * (synchronous call only), info[6] / info[7] microseconds the packer threads spent packing / waiting
 * for a staging slot (summed over threads), info[8] microseconds the calling thread had nothing to
 * enqueue. */
int gg_last_batch_transfer(gg_handle h, size_t info[9]);
int gg_fork_streams(gg_handle h);
int gg_join_streams(gg_handle h);

/* Counters for bench.py: number of kernel launches issued by this handle so far. */
uint64_t gg_kernel_launches(gg_handle h);

/* Number of levels / visits of the wavefront schedule of the spiral interpolation (diagnostics). */
int gg_spiral_schedule_info(gg_handle h, int* levels, int* visits, int* max_per_level);

#ifdef __cplusplus
}
#endif
#endif /* GROUNDGRID_B200_H */
