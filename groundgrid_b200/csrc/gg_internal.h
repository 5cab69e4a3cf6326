// Internal definitions shared by the kernels (gg_kernels.cu) and the C-ABI (gg_capi.cu).
// Not installed; the public boundary is include/groundgrid_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "groundgrid_b200.h"

namespace gg {

// ---- device layer slots (per map slot, each N*N floats, column-major) -------------------
// "ground" and "groundpatch" come first and are contiguous: they are the rolling terrain
// prior, the only state that survives a scan and what NCCL broadcasts.
enum Layer : int {
    L_GROUND = 0,       // "ground"            terrain height G
    L_GROUNDPATCH = 1,  // "groundpatch"       confidence C
    L_OBSTACLES = 2,    // "points" after labelling: non-ground points per cell (GroundSegmentation.cpp:147,176)
    L_VARIANCE = 3,     // "variance"          m2 / (count + FLT_MIN)            (:323)
    L_MINH = 4,         // "minGroundHeight"
    L_COUNT = 5,        // "points" during rasterisation: kept points per cell   (:309)
    L_MAXH = 6,         // "maxGroundHeight"   (dead layer)
    L_GCAND = 7,        // "groundCandidates"  (dead layer)
    L_PLANEDIST = 8,    // "planeDist"         (dead layer)
    L_M2 = 9,           // "m2"
    L_MEAN = 10,        // "meanVariance"      Welford mean
    L_RAW = 11,         // "pointsRaw"         (dead layer)
    L_NUM = 12,
    L_NUM_LIVE = 6,     // layers [0, 6) exist always; [6, 12) only with GG_FLAG_FULL_LAYERS
};

// per-point class codes (upper byte of PointAux code)
enum PointClass : unsigned {
    PC_ABSENT = 0,          // outside the map / NaN: never part of the output cloud
    PC_KEPT = 1,            // rasterised, labelled by the height test
    PC_KEPT_BORDER = 2,     // rasterised, but cell index >= N-3: vanishes from the output (:167-168)
    PC_IGNORED = 3,         // ring > max_ring or closer than sqrt(12) m: labelled, not rasterised (:237-240)
    PC_IGNORED_BORDER = 4,  // ignored and in a border cell: vanishes
    PC_OUTLIER = 5,         // below-ground outlier: always labelled ground (:270,185-189)
};

// Constants derived once per set_config on the host, with the reference's own expressions
// (so the fp64 values are what the reference's compiler computes).
struct Const {
    int N, N2;
    int max_ring;
    int full_layers;
    float pc_var_thresh_f;  // (float) point_count_cell_variance_threshold (float >= int compare, :374)
    float res_f;            // (float) map.getResolution()
    double res;             // map.getResolution()  == (double) res_f
    double rres;            // 1 / res (only to skip IEEE divisions whose integer part is not in doubt)
    double len, half;       // N * res, 0.5 * len
    double res_sq;          // res * res  (std::pow(resolution, 2.0), :332,356,463)
    double min_outlier_conf, outlier_tol;
    double gp_thresh;       // ground_patch_detection_minimum_point_count_threshold
    double df_sq, mdf_sq, mdf10_sq, psc_sq;
    double occ_factor, occ_factor2, dec_factor;
    double lab_fac, lab_thres, lab_obs;
    int decay_floor_ok;     // decay of a confidence <= 0.001f gives 0.001f again (checked on the host for dec_factor)
};

// Per-scan, per-slot parameters (changes every scan / pose update).
struct SlotParams {
    double px, py;        // map centre position
    double t20, t21, t22, t23;  // row 2 of T_base_from_map (seed of exposed cells)
    float ox, oy, oz;     // cloud origin
    float base_z_f;       // (float) base_z
    int n_points;
    int shift_i, shift_j; // pending roll (index shift): new(r,c) = old(r + shift_i, c + shift_j)
    int slot;
    const gg_point* src;  // device pointer of this scan's cloud (the slot's own buffer or a caller-owned one)
    // packed cloud (host-side packing of the batched host-buffer path): x | y | z as float[n_pad]
    // followed by ring as uint16[n_pad], n_pad = n_points rounded up to 8; null -> `src` is used
    const float* packed;
};

// Device side of the skewed-layout spiral (gg_host.h:SkewTables); null sk -> not used.
constexpr int SKEW_XCH_ASYNC = 4;    // depth of the exchange ring of the barrier-free spiral variants (the barrier variant needs 2); the
                                     // synchronisation table is built for exactly this depth (gg_capi.cu -> build_skew_sync)
struct SkewView {
    float2* sk;             // [n_slots][slots] (G, C) in (side, level, ring) order
    float* sd;              // [n_slots][slots] decayed confidence the visit will store, -1: confidence unchanged
    size_t slots;
    const int* cell_home;   // [N2][4]
    // Lane threads are time-shared: ring k and ring k + M of one side are never active at the same
    // level (ring k spans levels ~[3k, 5k]), so thread (side, m) walks rings m, m + M, m + 2M ... one
    // after the other ("phases").  [phase][side * M + m] tables; an empty phase has begin == end.
    const int* ph_begin;    // level range [begin, end) of the regular run
    const int* ph_end;
    const int* ph_cell0;    // cell (x + y * N) of the first regular visit
    int M, phases;
    // second layout of the same tables with the smallest possible M (small CTAs: several scans per SM); the
    // launcher swaps it in for batches, the first one serves single scans (lowest latency)
    const int* thr_ph_begin;
    const int* thr_ph_end;
    const int* thr_ph_cell0;
    int thr_M, thr_phases;
    // point-to-point synchronisation tables of the two layouts (gg_host.cpp:build_skew_sync), nullptr: CTA barrier per level
    //   [agent][level][32] progress agent b must have reached before `agent` starts the level
    const uint16_t* req;
    const uint16_t* thr_req;
    int sync_sleep;         // ns a waiting warp sleeps between two looks at the progress counters (0: spin)
    // irregular visits, one fixed-size block per level (irr_chunks x uint4):
    //   words [ (v * 9 + q) * 2 + {0, 1} ] = slot of neighbour q of visit v, producer lane if it was
    //                                        written one level ago (else 0xffffffff)
    //   words [ irr_max * 18 + v * 4 + {0..3} ] = own slot (0xffffffff: no visit), mirror slot, lane, cell
    const uint4* irr_blocks;
    int irr_max, irr_chunks;
    int KP, rows, row0, lanes, levels;
    int pattern[36];
};

// Pointers / strides of the handle's device arena, passed to kernels by value.
struct View {
    Const k;
    float* layers;        // [n_slots][n_layers][N2]
    int n_layers;
    const float* expected;  // [N2] expectedPoints table (GroundSegmentation.cpp:40-46)
    const float4* detect_tab;  // [N2] per-cell constants of the patch detection (gg_kernels.cu:k_build_detect_table)
    gg_point* points;     // [n_slots][pcap]
    unsigned char* packed;  // [n_slots][14 * pcap] packed clouds (allocated on first use)
    uint2* zw;            // [n_slots][pcap] per input point: (z bits, position in the cell's segment | run-head flag << 31)
    uint32_t* runj;       // [n_slots][pcap] run heads only: arrival number of the run in its cell | (run length - 1) << 26
    float* zsorted;       // [n_slots][pcap] z of kept points grouped by cell: the cell's segment is a sequence of runs
    uint2* rundir;        // [n_slots][pcap] run directory, cell by cell at cellstart: (run id = point index >> 5, first position | (length - 1) << 26)
    float* dist;          // [n_slots][pcap] hypotf(x - ox, y - oy)
    uint32_t* code;       // [n_slots][pcap] class << 24 | cell
    uint8_t* labels;      // [n_slots][pcap]
    unsigned long long* cnt64;  // [n_slots][N2] runs of the cell << 32 | kept points of the cell (one atomic per run)
    int* raw_i;           // [n_slots][N2] inside points per cell (full layers)
    int* cellstart;       // [n_slots][N2] exclusive scan of the per-cell point counts
    int4* worklist;       // [n_slots][N2] non-empty cells grouped by count class, heaviest first: (cell, points, runs, segment start)
    int* wl_count;        // [n_slots][2] entries of the worklist, kept points of the scan
    int* cell_agg;        // [n_slots][cell_tiles][65] per tile of 4096 cells: cells per count class, kept points
    int cell_tiles;       // ceil(N2 / 4096)
    uint32_t* out_index;  // [n_slots][pcap]
    int* out_counts;      // [n_slots][3 * out_blocks + 1]  (+1: n_out)
    gg_point* out_cloud;  // [n_slots][pcap] (allocated lazily)
    float* roll_scratch;  // [n_slots][2][N2]
    size_t pcap;
    int out_blocks;       // ceil(pcap / OUT_TILE)
    // spiral wavefront schedule (shared by all slots)
    const int* level_start;   // [levels + 1]
    const uint32_t* visits;   // [n_visits]  x | y << 16
    int levels;
    // pipelined spiral (k_spiral_pipe): 16-byte records, see gg_host.cpp:build_spiral_records
    const uint4* spiral_recs; // null -> plain k_spiral
    int spiral_dist;          // prefetch distance the records were built for (1..3)
    int spiral_threads;       // 512 or 1024 (>= max visits per level)
    SkewView skew;            // skew.sk != null -> k_skew + k_spiral_skew + k_unskew replace k_spiral_pipe

    __host__ __device__ float* layer(int slot, int l) const { return layers + ((size_t)slot * n_layers + l) * k.N2; }
};

constexpr int RASTER_TILE = 2048;
constexpr int RASTER_THREADS = 256;
constexpr int OUT_TILE = 1024;

// kernels of the pipeline, as reported by the profiling hooks (gg_profile_read)
enum KernelId : int {
    K_RASTERIZE = 0, K_CELL_TILES, K_CELL_PLACE, K_SCATTER, K_CELL_STATS, K_DETECT, K_SPIRAL, K_LABEL, K_ROLL_GATHER, K_ROLL_COMMIT, K_OUT_COUNT, K_OUT_SCAN,
    K_OUT_WRITE, K_UNPACK, K_TERRAIN, K_EVAL, K_NUM
};

// Optional per-kernel CUDA-event timing (bench.py's roofline needs the dominant kernel's own
// duration measured live on the launching stream).  Null = no events.
struct Profiler {
    virtual void begin(int kernel_id, cudaStream_t st) = 0;
    virtual void end(int kernel_id, cudaStream_t st) = 0;
    virtual ~Profiler() {}
};

// ---- launchers (gg_kernels.cu); every function enqueues on `st` and returns the number of
// kernel launches it issued (for gg_kernel_launches()). -----------------------------------
int launch_init_map(const View& v, int slot, float z, cudaStream_t st);
int launch_build_detect_table(const View& v, float4* tab, cudaStream_t st);
int launch_roll(const View& v, const SlotParams* batch, int count, cudaStream_t st, Profiler* prof);
// layer_map: TMA descriptor of the handle's layer arena as a 3-D tensor (i, j, slot * n_layers + layer), box
// 40 x 12 x 1 (k_detect_tma); null -> the patch detection stages its tile with plain loads (N % 4 != 0)
// after_detect (may be null): recorded on st right before the spiral kernel
int launch_scan_pipeline(const View& v, const SlotParams* batch, int count, int max_points, int stop_after, cudaStream_t st,
                         Profiler* prof, const CUtensorMap* layer_map, cudaEvent_t after_detect);
// single phases / single cells (the reference's public per-phase methods)
int launch_detect_only(const View& v, const SlotParams* batch, int count, cudaStream_t st, Profiler* prof, const CUtensorMap* layer_map);
int launch_spiral_only(const View& v, const SlotParams* batch, int count, cudaStream_t st, Profiler* prof);
int launch_interpolate_cell(const View& v, int slot, int x, int y, cudaStream_t st);
int launch_detect_cell(const View& v, int slot, int S, int i, int j, cudaStream_t st);
int launch_output(const View& v, const SlotParams* batch, int count, int max_points, bool want_cloud, cudaStream_t st,
                  Profiler* prof);
// "next" rows of SURVEY.md section 8(f)
struct UnpackDesc {   // f1: PointCloud2 payload -> PointXYZIR records in the map frame
    const unsigned char* raw;  // device copy of msg.data
    gg_point* dst;
    int n, point_step;
    int off[5];                // byte offsets of x, y, z, intensity, ring (-1: field absent)
    int transform;             // 0: frame_id == "map", copy only
    double T[12];              // row-major 3x4 [R|t] of lookupTransform("map", frame_id)
};
int launch_unpack(const UnpackDesc& d, cudaStream_t st, Profiler* prof);
int launch_terrain_image(const View& v, int slot, float* dst, cudaStream_t st, Profiler* prof);
// mm: 2 floats (ordered-int keys of min / max), preset by the caller to the keys of +inf / -inf; dst: N * N bytes, row-major (i, j)
int launch_layer_image_u8(const View& v, const float* layer, float* mm, unsigned char* dst, cudaStream_t st);
int launch_eval(const View& v, const SlotParams* batch, unsigned long long* counts, cudaStream_t st, Profiler* prof);
constexpr int EVAL_LABELS = 1024;  // ring values (SemanticKITTI label ids <= 259) x {ground, non-ground}

}  // namespace gg
