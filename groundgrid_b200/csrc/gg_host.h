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
void pack_cloud_range(const gg_point* src, size_t n, unsigned char* dst, size_t i0, size_t i1);
int usable_cpus();
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
}
