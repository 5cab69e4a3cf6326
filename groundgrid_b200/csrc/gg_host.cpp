// Host-side (no CUDA) pieces of the library: expectedPoints table, config-derived constants,
// the map-move arithmetic and the wavefront schedule of the spiral interpolation.  They are
// exported with a gg_host_ prefix so the CPU test-suite can exercise them without a GPU.
#include <immintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <thread>
#include <cstdint>
#include <cstring>
#include <vector>

#include <sched.h>

#include "gg_host.h"

namespace gg {

// GroundSegmentation::init (src/GroundSegmentation.cpp:37-48).  Built on the host with the
// platform libm (hypot / atanf), exactly like the reference, then uploaded once.
int cells_per_side(double dimension_m, float resolution) {
    return (int)std::round((float)dimension_m / resolution);
}

void build_expected_points(int n, std::vector<float>& table) {
    const float vertical_point_ang_dist = 0.00174532925 * 2;  // GroundSegmentation.h:69
    table.assign((size_t)n * n, 0.0f);
    const size_t cells = (size_t)n;
    for (size_t i = 0; i < cells; ++i)
        for (size_t j = 0; j < cells; ++j) {
            const float dist = std::hypot(i - cells / 2.0, j - cells / 2.0);
            table[i + j * cells] = std::atan(1 / dist) / vertical_point_ang_dist;
        }
}

// grid_map::GridMap::setGeometry + the constants the kernels need; squares are written as
// x * x, which is what the reference's compiler emits for std::pow(x, 2.0).
void derive_constants(const gg_config& c, double dimension_m, float resolution, unsigned flags, Const& k) {
    const double res = (double)resolution;
    const int n = (int)std::round((double)(float)dimension_m / res);
    k.N = n;
    k.N2 = n * n;
    k.max_ring = c.max_ring;
    k.full_layers = (flags & GG_FLAG_FULL_LAYERS) ? 1 : 0;
    k.pc_var_thresh_f = (float)c.point_count_cell_variance_threshold;
    k.res_f = (float)res;
    k.res = res;
    k.len = (double)n * res;
    k.half = 0.5 * k.len;
    k.res_sq = (double)k.res_f * (double)k.res_f;
    k.min_outlier_conf = c.min_outlier_detection_ground_confidence;
    k.outlier_tol = c.outlier_tolerance;
    k.gp_thresh = c.ground_patch_detection_minimum_point_count_threshold;
    k.df_sq = c.distance_factor * c.distance_factor;
    k.mdf_sq = c.minimum_distance_factor * c.minimum_distance_factor;
    const double m10 = c.minimum_distance_factor * 10;
    k.mdf10_sq = m10 * m10;
    k.psc_sq = c.patch_size_change_distance * c.patch_size_change_distance;
    k.occ_factor = c.occupied_cells_point_count_factor;
    k.occ_factor2 = c.occupied_cells_point_count_factor * 2.0f;
    k.dec_factor = c.occupied_cells_decrease_factor;
    k.lab_fac = c.minimum_distance_factor * 5;
    k.lab_thres = c.miminum_point_height_threshold;
    k.lab_obs = c.minimum_point_height_obstacle_threshold;
}

// grid_map::GridMap::move (getIndexShiftFromPositionShift / getPositionShiftFromIndexShift):
// whole-cell shift, rounded half away from zero; the map position advances by the aligned
// shift.  Returns the buffer index shift: new(r, c) = old(r + shift_i, c + shift_j).
void move_map(double res, double& px, double& py, double nx, double ny, int& shift_i, int& shift_j) {
    const double tx = (nx - px) / res, ty = (ny - py) / res;
    const int cx = (int)(tx + 0.5 * (tx > 0 ? 1 : -1));
    const int cy = (int)(ty + 0.5 * (ty > 0 ? 1 : -1));
    shift_i = -cx;
    shift_j = -cy;
    px += (double)cx * res;
    py += (double)cy * res;
}

// Wavefront schedule of GroundSegmentation::spiral_ground_interpolation (:398-441).
// Every visit reads the 3x3 neighbourhood of G and C and writes its own cell (:453-464).
// level(v) = 1 + max(level of the last writer of any cell it reads, level of any earlier
// reader of the cell it writes); visits of one level are mutually independent, so running
// the levels in order reproduces the sequential sweep exactly.
void build_spiral_schedule(int n, std::vector<int>& level_start, std::vector<uint32_t>& visits) {
    const int c = n / 2 - 1;
    std::vector<int> last_write((size_t)n * n, -1), last_read((size_t)n * n, -1);
    std::vector<uint32_t> seq;
    std::vector<int> lvl;
    auto visit = [&](int x, int y) {
        int l = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) l = std::max(l, last_write[(x + dx) + (size_t)(y + dy) * n] + 1);
        l = std::max(l, last_read[x + (size_t)y * n] + 1);
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int& r = last_read[(x + dx) + (size_t)(y + dy) * n];
                r = std::max(r, l);
            }
        last_write[x + (size_t)y * n] = l;
        seq.push_back((uint32_t)x | ((uint32_t)y << 16));
        lvl.push_back(l);
    };
    for (int p = c - 1; p >= 1; --p) {
        const int side = (c - p) * 2;
        const int q = p + side;
        for (int pos = p; pos < q; ++pos) visit(p, pos);       // side 0: x fixed
        for (int pos = p; pos < q; ++pos) visit(pos, p);       // side 1: y fixed
        for (int pos = q; pos >= p; --pos) visit(q, pos);      // far sides, descending
        for (int pos = q; pos >= p; --pos) visit(pos, q);
    }
    int levels = 0;
    for (int l : lvl) levels = std::max(levels, l + 1);
    level_start.assign((size_t)levels + 1, 0);
    for (int l : lvl) ++level_start[(size_t)l + 1];
    for (int l = 0; l < levels; ++l) level_start[(size_t)l + 1] += level_start[l];
    visits.resize(seq.size());
    std::vector<int> cursor(level_start.begin(), level_start.end() - 1);
    for (size_t v = 0; v < seq.size(); ++v) visits[(size_t)cursor[lvl[v]]++] = seq[v];
}

// Schedule records for the pipelined spiral kernel (k_spiral_pipe): per visit
//   w0 = x | y << 16
//   w1, w2 = up to four "recent" entries (16 bit each): bits 14..15 = how many levels earlier
//            the neighbour was written (1..3, 0 = unused entry), bits 10..13 neighbour index q
//            (block(q % 3, q / 3) of the 3x3 neighbourhood), bits 0..9 the slot (index inside
//            its level) of the visit that wrote it
//   w3 = bit 0: the cell lies beyond minDistSquared (confidence decays, :463-464)
//        bit 1: second visit of this cell (ring corners are visited twice, :421-438)
// The kernel prefetches a visit's 3x3 neighbourhood `dist` levels ahead; a neighbour written
// less than `dist` levels before the visit cannot come from that prefetch and is delivered
// through shared memory instead.  Returns false if a visit needs more than four such recents
// or a level has more than 1024 visits (the caller then uses the plain k_spiral).
bool build_spiral_records(int n, double res_sq, const std::vector<int>& level_start, const std::vector<uint32_t>& visits,
                          int dist, std::vector<uint32_t>& recs, int& max_recent) {
    const int c = n / 2 - 1;
    const size_t nv = visits.size();
    if (dist < 1 || dist > 3) return false;
    // level and slot of every visit, looked up by (level-sorted) position
    std::vector<int> level_of(nv), slot_of(nv);
    for (size_t l = 0; l + 1 < level_start.size(); ++l) {
        if (level_start[l + 1] - level_start[l] > 1024) return false;
        for (int t = level_start[l]; t < level_start[l + 1]; ++t) {
            level_of[t] = (int)l;
            slot_of[t] = t - level_start[l];
        }
    }
    // Replay the sequential order to know, for each visit, the last writer of every neighbour.
    // Sequential order != level-sorted order, so map (cell, occurrence) -> sorted position.
    std::vector<std::vector<int>> pos_of_cell((size_t)n * n);
    for (size_t t = 0; t < nv; ++t) {
        const int x = visits[t] & 0xffff, y = visits[t] >> 16;
        pos_of_cell[x + (size_t)y * n].push_back((int)t);  // level order == visit order for one cell (WAW)
    }
    std::vector<int> seen((size_t)n * n, 0);      // how many visits of the cell happened so far
    std::vector<int> last_writer((size_t)n * n, -1);  // sorted position of the cell's last writer
    recs.assign(nv * 4, 0u);
    max_recent = 0;
    bool ok = true;
    auto visit = [&](int x, int y) {
        const size_t cell = x + (size_t)y * n;
        const int t = pos_of_cell[cell][seen[cell]];
        const int lvl = level_of[t];
        uint32_t r[4] = {0, 0, 0, 0};
        int nr = 0;
        for (int q = 0; q < 9; ++q) {
            const size_t nb = (size_t)(x - 1 + q % 3) + (size_t)(y - 1 + q / 3) * n;
            const int w = last_writer[nb];
            if (w >= 0 && level_of[w] < lvl && level_of[w] >= lvl - dist) {
                // bits 14..15: how many levels earlier the neighbour was written (1 .. dist <= 3); 0 = unused entry
                if (nr < 4) r[nr] = ((uint32_t)(lvl - level_of[w]) << 14) | ((uint32_t)q << 10) | (uint32_t)slot_of[w];
                ++nr;
            } else if (w >= 0 && level_of[w] >= lvl) {
                ok = false;  // would contradict the levelisation
            }
        }
        if (nr > 4) ok = false;
        max_recent = std::max(max_recent, nr);
        const float fx = (float)x - (float)c, fy = (float)y - (float)c;
        const bool far = ((double)fx * (double)fx + (double)fy * (double)fy) * res_sq > 12.0;
        recs[4 * (size_t)t + 0] = (uint32_t)x | ((uint32_t)y << 16);
        recs[4 * (size_t)t + 1] = r[0] | (r[1] << 16);
        recs[4 * (size_t)t + 2] = r[2] | (r[3] << 16);
        recs[4 * (size_t)t + 3] = (far ? 1u : 0u) | (seen[cell] ? 2u : 0u);  // bit 1: second visit of a ring corner
        last_writer[cell] = t;
        ++seen[cell];
    };
    for (int p = c - 1; p >= 1; --p) {
        const int side = (c - p) * 2;
        const int q = p + side;
        for (int pos = p; pos < q; ++pos) visit(p, pos);
        for (int pos = p; pos < q; ++pos) visit(pos, p);
        for (int pos = q; pos >= p; --pos) visit(q, pos);
        for (int pos = q; pos >= p; --pos) visit(pos, q);
    }
    return ok;
}

// ---- host-side cloud packing (gg_filter_cloud_batch) -------------------------------------
// PointXYZIR records (32 B, 14 useful) -> x | y | z (float[n_pad]) | ring (u16[n_pad]) with
// n_pad = n rounded up to 8, written with streaming stores into pinned staging memory.
// [i0, i1) is a chunk: i0 a multiple of 8, i1 == n for the last chunk (which also zero-fills
// the padding).  AVX2 variant selected at run time.
static void pack_tail(const gg_point* src, size_t n, size_t n_pad, float* x, float* y, float* z, uint16_t* r, size_t i, size_t i1) {
    for (; i < i1 && i < n; ++i) {
        x[i] = src[i].x;
        y[i] = src[i].y;
        z[i] = src[i].z;
        r[i] = src[i].ring;
    }
    if (i1 >= n)
        for (size_t t = n; t < n_pad; ++t) {
            x[t] = y[t] = z[t] = 0.f;
            r[t] = 0;
        }
}

static void pack_sse2(const gg_point* src, size_t n, unsigned char* dst, size_t i0, size_t i1) {
    const size_t n_pad = (n + 7) & ~(size_t)7;
    float* x = reinterpret_cast<float*>(dst);
    float* y = x + n_pad;
    float* z = y + n_pad;
    uint16_t* r = reinterpret_cast<uint16_t*>(z + n_pad);
    size_t i = i0;
    const size_t vec_end = i0 + ((std::min(i1, n) - i0) & ~(size_t)7);
    for (; i < vec_end; i += 8) {
        alignas(16) uint16_t rr[8];
        for (int h = 0; h < 2; ++h) {
            const gg_point* p = src + i + 4 * h;
            __m128 a = _mm_loadu_ps(&p[0].x), b = _mm_loadu_ps(&p[1].x), c = _mm_loadu_ps(&p[2].x), d = _mm_loadu_ps(&p[3].x);
            _MM_TRANSPOSE4_PS(a, b, c, d);  // a = x0..x3, b = y0..y3, c = z0..z3
            _mm_stream_ps(x + i + 4 * h, a);
            _mm_stream_ps(y + i + 4 * h, b);
            _mm_stream_ps(z + i + 4 * h, c);
            for (int q = 0; q < 4; ++q) rr[4 * h + q] = p[q].ring;
        }
        _mm_stream_si128(reinterpret_cast<__m128i*>(r + i), _mm_load_si128(reinterpret_cast<const __m128i*>(rr)));
    }
    pack_tail(src, n, n_pad, x, y, z, r, i, i1);
    _mm_sfence();
}

__attribute__((target("avx2"))) static void pack_avx2(const gg_point* src, size_t n, unsigned char* dst, size_t i0, size_t i1) {
    const size_t n_pad = (n + 7) & ~(size_t)7;
    float* x = reinterpret_cast<float*>(dst);
    float* y = x + n_pad;
    float* z = y + n_pad;
    uint16_t* r = reinterpret_cast<uint16_t*>(z + n_pad);
    size_t i = i0;
    const size_t vec_end = i0 + ((std::min(i1, n) - i0) & ~(size_t)7);
    for (; i < vec_end; i += 8) {
        const gg_point* p = src + i;
        _mm_prefetch(reinterpret_cast<const char*>(p + 32), _MM_HINT_NTA);
        _mm_prefetch(reinterpret_cast<const char*>(p + 34), _MM_HINT_NTA);
        _mm_prefetch(reinterpret_cast<const char*>(p + 36), _MM_HINT_NTA);
        _mm_prefetch(reinterpret_cast<const char*>(p + 38), _MM_HINT_NTA);
        // lane 0: records 0..3, lane 1: records 4..7; transpose 4x4 inside each 128-bit lane
        const __m256 a = _mm256_loadu2_m128(&p[4].x, &p[0].x), b = _mm256_loadu2_m128(&p[5].x, &p[1].x);
        const __m256 c = _mm256_loadu2_m128(&p[6].x, &p[2].x), d = _mm256_loadu2_m128(&p[7].x, &p[3].x);
        const __m256 t0 = _mm256_unpacklo_ps(a, b), t1 = _mm256_unpackhi_ps(a, b);
        const __m256 t2 = _mm256_unpacklo_ps(c, d), t3 = _mm256_unpackhi_ps(c, d);
        _mm256_stream_ps(x + i, _mm256_shuffle_ps(t0, t2, 0x44));  // x0..x3 | x4..x7
        _mm256_stream_ps(y + i, _mm256_shuffle_ps(t0, t2, 0xEE));
        _mm256_stream_ps(z + i, _mm256_shuffle_ps(t1, t3, 0x44));
        alignas(16) uint16_t rr[8];
        for (int q = 0; q < 8; ++q) rr[q] = p[q].ring;
        _mm_stream_si128(reinterpret_cast<__m128i*>(r + i), _mm_load_si128(reinterpret_cast<const __m128i*>(rr)));
    }
    pack_tail(src, n, n_pad, x, y, z, r, i, i1);
    _mm_sfence();
}

void pack_cloud_range(const gg_point* src, size_t n, unsigned char* dst, size_t i0, size_t i1) {
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0)
        pack_avx2(src, n, dst, i0, i1);
    else
        pack_sse2(src, n, dst, i0, i1);
}

// CPUs this process may actually use: affinity mask capped by the cgroup CPU quota
int usable_cpus() {
    int n = (int)std::thread::hardware_concurrency();
#ifdef __linux__
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(n, CPU_COUNT(&set));
    if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char quota[32];
        long period = 0;
        if (std::fscanf(f, "%31s %ld", quota, &period) == 2 && period > 0 && quota[0] != 'm') n = std::min(n, (int)std::max(1L, std::atol(quota) / period));
        std::fclose(f);
    } else if (FILE* f1 = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {
        long q = -1, per = 100000;
        if (std::fscanf(f1, "%ld", &q) != 1) q = -1;
        std::fclose(f1);
        if (FILE* f2 = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) {
            if (std::fscanf(f2, "%ld", &per) != 1) per = 100000;
            std::fclose(f2);
        }
        if (q > 0 && per > 0) n = std::min(n, (int)std::max(1L, q / per));
    }
#endif
    return std::max(1, n);
}

}  // namespace gg

extern "C" {

int gg_host_cells_per_side(double dimension_m, float resolution) { return gg::cells_per_side(dimension_m, resolution); }

int gg_host_expected_points(double dimension_m, float resolution, float* dst) {
    std::vector<float> t;
    const int n = gg::cells_per_side(dimension_m, resolution);
    gg::build_expected_points(n, t);
    std::memcpy(dst, t.data(), t.size() * sizeof(float));
    return n;
}

int gg_host_spiral_schedule(int n, int* level_start, int level_cap, uint32_t* visits, int visit_cap, int* n_levels, int* n_visits) {
    std::vector<int> ls;
    std::vector<uint32_t> vs;
    gg::build_spiral_schedule(n, ls, vs);
    *n_levels = (int)ls.size() - 1;
    *n_visits = (int)vs.size();
    if (level_start && level_cap >= (int)ls.size()) std::memcpy(level_start, ls.data(), ls.size() * sizeof(int));
    if (visits && visit_cap >= (int)vs.size()) std::memcpy(visits, vs.data(), vs.size() * sizeof(uint32_t));
    return 0;
}

// records of the pipelined spiral kernel: 4 uint32 per visit (see build_spiral_records)
int gg_host_spiral_records(int n, float resolution, int dist, uint32_t* recs, int rec_cap_words, int* max_recent) {
    std::vector<int> ls;
    std::vector<uint32_t> vs, rc;
    gg::build_spiral_schedule(n, ls, vs);
    int mr = 0;
    const bool ok = gg::build_spiral_records(n, (double)resolution * (double)resolution, ls, vs, dist, rc, mr);
    if (max_recent) *max_recent = mr;
    if (recs && rec_cap_words >= (int)rc.size()) std::memcpy(recs, rc.data(), rc.size() * sizeof(uint32_t));
    return ok ? 1 : 0;
}

int gg_host_move_map(double res, double* pos_xy, double nx, double ny, int* shift_ij) {
    gg::move_map(res, pos_xy[0], pos_xy[1], nx, ny, shift_ij[0], shift_ij[1]);
    return (shift_ij[0] != 0 || shift_ij[1] != 0) ? 1 : 0;
}
}
