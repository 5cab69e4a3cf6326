// Host-only helpers (gg_host.cpp); see there for the reference citations.
#pragma once
#include <cstdint>
#include <vector>

#include "gg_internal.h"

namespace gg {
int cells_per_side(double dimension_m, float resolution);
void build_expected_points(int n, std::vector<float>& table);
void derive_constants(const gg_config& c, double dimension_m, float resolution, unsigned flags, Const& k);
void move_map(double res, double& px, double& py, double nx, double ny, int& shift_i, int& shift_j);
void build_spiral_schedule(int n, std::vector<int>& level_start, std::vector<uint32_t>& visits);
// cached: plain stores (destination meant to stay in the last-level cache until the DMA engine reads it)
// instead of streaming stores
void pack_cloud_range(const gg_point* src, size_t n, unsigned char* dst, size_t i0, size_t i1, bool cached = false);
int usable_cpus();

// Tables of the skewed-layout spiral kernel (k_spiral_skew), see gg_host.cpp:build_spiral_skew.
struct SkewTables {
    bool ok = false;
    int n = 0, K = 0, KP = 0, rows = 0, levels = 0, row0 = 8;
    int lanes = 0;                       // 4 * KP
    size_t slots = 0;                    // 4 * rows * KP
    int pattern[36] = {0};               // [side][q]: slot offset of neighbour q relative to the own slot (regular visits)
    int prev_q[4] = {1, 3, 7, 5};        // neighbour index of a lane's previous cell
    std::vector<int> lane_begin, lane_end;   // [lanes] level range [begin, end) of the lane's regular run
    std::vector<int> lane_cell0;             // [lanes] cell (x + y * n) of the first regular visit; the lane then steps by +n, +1, -n, -1 (side 0..3)
    std::vector<int> cell_home;          // [n * n * 4] slots holding a copy of the cell (-1 padded; [0] first visit, [1] second visit)
    std::vector<int> irr_level_start;    // [levels + 1] CSR over levels
    std::vector<uint32_t> irr_recs;      // 16 words per irregular visit: own, nb[9], recents01, recents23, mirror, lane, cell, pad
    int max_irr_per_level = 0;
    size_t n_regular = 0, n_irregular = 0;
    // every visit in SEQUENTIAL order with the memory it touches in k_spiral_skew (input of build_skew_sync)
    struct Access {
        int level, lane;          // lane = side * KP + ring + 1 (also the visit's exchange-buffer entry)
        int own, mirror, cell;    // slots written (mirror: -1 or the second home of a ring corner), cell of the normal layers
        int nb[9];                // slots read
        int recent_lane[9];       // >= 0: neighbour q was written one level ago and arrives through the exchange buffer entry of this lane
        bool regular;             // executed by its lane thread (else by the irregular warps)
    };
    std::vector<Access> acc;
};
// Point-to-point synchronisation of k_spiral_skew<.., ASYNC>: the CTA-wide barrier per level is replaced by progress
// counters, one per AGENT (a warp of lane threads, or the two irregular warps together).  req[(agent * levels + l) * 32 + b]
// = progress agent b must have reached (= number of levels it has completed) before `agent` may start level l.
// Derived from the exact read / write sets of every visit (RAW, WAR and WAW on the skewed copy, the normal layers and
// the exchange ring of depth xch_depth), with the kernel's timing: the slots of a visit at level l are loaded during
// level l - 1, neighbours written at level l - 1 are read from the exchange ring during level l, stores happen at level l.
// M = lane threads per side (thread layout).  Returns false if some dependence cannot be expressed (then the barrier
// kernel is used).
bool build_skew_sync(const SkewTables& t, int M, int xch_depth, std::vector<uint16_t>& req, int& n_agents);
void build_spiral_skew(int n, const std::vector<int>& level_start, const std::vector<uint32_t>& visits, SkewTables& t);
bool build_spiral_records(int n, double res_sq, const std::vector<int>& level_start, const std::vector<uint32_t>& visits,
                          int dist, std::vector<uint32_t>& recs, int& max_recent);
}  // namespace gg

extern "C" {
// CPU-testable exports (no device needed).
int gg_host_cells_per_side(double dimension_m, float resolution);
int gg_host_expected_points(double dimension_m, float resolution, float* dst);
int gg_host_spiral_schedule(int n, int* level_start, int level_cap, uint32_t* visits, int visit_cap, int* n_levels, int* n_visits);
int gg_host_spiral_records(int n, float resolution, int dist, uint32_t* recs, int rec_cap_words, int* max_recent);
int gg_host_move_map(double res, double* pos_xy, double nx, double ny, int* shift_ij);
int gg_host_pack_cloud(const gg_point* src, size_t n, unsigned char* dst);
int gg_host_pack_cloud_cached(const gg_point* src, size_t n, unsigned char* dst);
int gg_host_packer_selftest(int threads, int n_jobs, size_t n_points, int ring_slots, int rounds, int lag);
int gg_host_spiral_skew_sync(int n, int M, int xch_depth, uint16_t* req, int req_cap, int* n_agents);
int gg_host_spiral_skew(int n, int* header, int* pattern, int* lane_begin, int* lane_end, int* cell_home, int* irr_level_start, uint32_t* irr_recs, int irr_cap_words);
}
