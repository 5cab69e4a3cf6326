// groundgrid_b200 -- hand-written sm_100a kernels of the GroundGrid per-scan hot path.
//
// Phases (reference: src/GroundSegmentation.cpp, see DESIGN.md for the data layout):
//   k_rasterize     point -> cell, ignore test, outlier ray-march; every warp claims, per cell it
//                   touches, one contiguous run of the cell's segment (one atomic per run)     (:200-280)
//   k_cell_tiles/place  exclusive scan of the per-cell counts -> segment starts; worklist of non-empty cells by count class
//   k_scatter       z of every kept point -> its slot of the cell's segment, run descriptors
//   k_cell_stats    per-cell sequential count / min / Welford mean+M2 -> variance, walking the
//                   runs of a cell in input order (the Welford recurrence of :298-305 is order
//                   dependent)                                                                 (:282-309,323)
//   k_detect        3x3 / 5x5 ground-patch stencil updating G, C              (:314-395)
//   k_spiral        level-scheduled wavefront of the in-place spiral sweep    (:398-465)
//   k_label         per-point ground / non-ground decision                    (:146-196)
//   k_out_*         output cloud order: kept, ignored, outliers               (:112-117,150,185)
//   k_roll_*        GroundGrid::update: whole-cell roll + seeding             (GroundGrid.cpp:83-147)
//
// Arithmetic rules: no FMA contraction (compiled with --fmad=false AND
// written with __f*_rn intrinsics where a product feeds a sum), IEEE division / sqrt,
// fp64 wherever the reference's C++ promotes to double, Eigen 3.3.7's binary-split
// reduction order for every fixed-size block sum.
#include <cuda.h>   // CUtensorMap (the descriptor is encoded on the host, gg_capi.cu)

#include <cfloat>
#include <cstddef>
#include <cstdio>

#include "gg_internal.h"

namespace gg {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int trunc_index(double v) {
    // grid_map: indexVector.cast<int>() (truncation toward zero); clamped so that NaN / huge
    // values map far outside instead of being undefined.
    if (!(v == v)) return 1000000000;
    if (v > 1.0e9) return 1000000000;
    if (v < -1.0e9) return -1000000000;
    return __double2int_rz(v);
}

// trunc(a / res) exactly as the IEEE division would give it, without paying for the division in the
// common case: q = a * (1 / res) differs from the correctly rounded quotient by a few ulp
// (|q| < 1e6 here, so < 1e-9 absolute); whenever q is farther than 1e-6 from every integer, no
// integer lies between the two values and trunc(q) == trunc(a / res).  Otherwise (a point within
// a micro-cell of a cell boundary, or a huge / non-finite value) the real division decides.
__device__ __forceinline__ int trunc_quotient(double a, double res, double rres) {
    const double q = __dmul_rn(a, rres);
    const double f = __dsub_rn(q, rint(q));
    if (fabs(q) < 1.0e6 && fabs(f) > 1.0e-6) return __double2int_rz(q);
    return trunc_index(__ddiv_rn(a, res));
}

// grid_map_core getIndexFromPosition with start index (0,0): -(int)((p - len/2 - pos) / res)
__device__ __forceinline__ void grid_index(const Const& k, double px, double py, double x, double y, int& ix, int& iy) {
    ix = -trunc_quotient(__dsub_rn(__dsub_rn(x, k.half), px), k.res, k.rres);
    iy = -trunc_quotient(__dsub_rn(__dsub_rn(y, k.half), py), k.res, k.rres);
}

// grid_map_core checkIfPositionWithinMap: t = -(p - pos - len/2); 0 <= t < len
__device__ __forceinline__ bool grid_inside(const Const& k, double px, double py, double x, double y) {
    const double tx = -__dsub_rn(__dsub_rn(x, px), k.half);
    const double ty = -__dsub_rn(__dsub_rn(y, py), k.half);
    return tx >= 0.0 && ty >= 0.0 && tx < k.len && ty < k.len;
}

// Eigen 3.3.7 redux_novec_unroller: binary split over the block's coefficients (column-major).
template <int Start, int Len>
struct TreeSum {
    template <typename F>
    __device__ __forceinline__ static float run(const F& e) {
        return __fadd_rn(TreeSum<Start, Len / 2>::run(e), TreeSum<Start + Len / 2, Len - Len / 2>::run(e));
    }
};
template <int Start>
struct TreeSum<Start, 1> {
    template <typename F>
    __device__ __forceinline__ static float run(const F& e) {
        return e(Start);
    }
};

__device__ __forceinline__ float tree9(const float* e) {
    return __fadd_rn(__fadd_rn(__fadd_rn(e[0], e[1]), __fadd_rn(e[2], e[3])),
                     __fadd_rn(__fadd_rn(e[4], e[5]), __fadd_rn(e[6], __fadd_rn(e[7], e[8]))));
}

__device__ __forceinline__ int warp_inclusive_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Exclusive scan of n ints (in -> out, may alias) by one block of 1024 threads: tiles of 4096
// elements, 4 consecutive ints per thread (coalesced), next tile prefetched while the current
// one is scanned.  Returns the grand total to every thread.
template <int STRIDE = 1>
__device__ int block_exclusive_scan_1024(const int* in, int* out, int n) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    int nx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) nx[q] = (tid * 4 + q < n) ? in[(size_t)(tid * 4 + q) * STRIDE] : 0;
    __syncthreads();
    for (int base = 0; base < n; base += 4096) {
        int cur[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = nx[q];
        const int nb = base + 4096 + tid * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) nx[q] = (nb + q < n) ? in[(size_t)(nb + q) * STRIDE] : 0;
        const int sum = cur[0] + cur[1] + cur[2] + cur[3];
        const int incl = warp_inclusive_scan(sum);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            const int wi = warp_inclusive_scan(w);
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const int carry = s_carry;
        int run = carry + s_warp[warp] + incl - sum;
        const int idx = base + tid * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (idx + q < n) out[idx + q] = run;
            run += cur[q];
        }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
    return s_carry;
}

// ------------------------------------------------------------------------------------------
// fills / map init / roll
// ------------------------------------------------------------------------------------------
// GroundGrid::update (GroundGrid.cpp:96-133,143): new(r, c) = old(r + shift_i, c + shift_j);
// exposed cells: ground = -(T * (cx, cy, 0)).z in fp64, groundpatch = 0.
// Four consecutive cells per thread: the eight (shifted, hence unaligned but contiguous) loads are issued together and,
// when N*N is a multiple of 4 (every layer then starts 16-byte aligned), the results leave as 16-byte stores.
constexpr int ROLL_ILP = 4;

__global__ void __launch_bounds__(256) k_roll_gather(View v, const SlotParams* __restrict__ batch) {
    const SlotParams& sp = batch[blockIdx.y];
    if (sp.shift_i == 0 && sp.shift_j == 0) return;
    const Const& k = v.k;
    const int cell0 = (blockIdx.x * 256 + threadIdx.x) * ROLL_ILP;
    if (cell0 >= k.N2) return;
    const int N = k.N;
    const float* G = v.layer(sp.slot, L_GROUND);
    const float* C = v.layer(sp.slot, L_GROUNDPATCH);
    float* sg = v.roll_scratch + (size_t)sp.slot * 2 * k.N2;
    float* sc = sg + k.N2;
    float g[ROLL_ILP], c[ROLL_ILP];
    bool seed[ROLL_ILP];
#pragma unroll
    for (int u = 0; u < ROLL_ILP; ++u) {
        const int cell = cell0 + u;
        const int r = cell % N, cc = cell / N;
        const int orr = r + sp.shift_i, occ = cc + sp.shift_j;
        const bool live = cell < k.N2;
        seed[u] = live && !(orr >= 0 && orr < N && occ >= 0 && occ < N);
        const bool ld = live && !seed[u];
        g[u] = ld ? G[orr + occ * N] : 0.0f;
        c[u] = ld ? C[orr + occ * N] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < ROLL_ILP; ++u)
        if (seed[u]) {
            const int cell = cell0 + u;
            const int r = cell % N, cc = cell / N;
            // grid_map getPositionFromIndex: pos + (len/2 - res/2) + res * (-(double)index)
            const double off = __dsub_rn(k.half, __dmul_rn(0.5, k.res));
            const double x = __dadd_rn(__dadd_rn(sp.px, off), __dmul_rn(k.res, (double)(-r)));
            const double y = __dadd_rn(__dadd_rn(sp.py, off), __dmul_rn(k.res, (double)(-cc)));
            // tf2::Transform * Vector3(x, y, 0): row2.dot(v) + origin.z
            const double tz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sp.t20, x), __dmul_rn(sp.t21, y)), __dmul_rn(sp.t22, 0.0)), sp.t23);
            g[u] = (float)(-tz);
            c[u] = 0.0f;
        }
    if ((k.N2 & 3) == 0) {
        *reinterpret_cast<float4*>(sg + cell0) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(sc + cell0) = make_float4(c[0], c[1], c[2], c[3]);
    } else {
#pragma unroll
        for (int u = 0; u < ROLL_ILP; ++u)
            if (cell0 + u < k.N2) {
                sg[cell0 + u] = g[u];
                sc[cell0 + u] = c[u];
            }
    }
}

__global__ void __launch_bounds__(256) k_roll_commit(View v, const SlotParams* __restrict__ batch) {
    const SlotParams& sp = batch[blockIdx.y];
    if (sp.shift_i == 0 && sp.shift_j == 0) return;
    const int N2 = v.k.N2;
    const int cell0 = (blockIdx.x * 256 + threadIdx.x) * ROLL_ILP;
    if (cell0 >= N2) return;
    const float* sg = v.roll_scratch + (size_t)sp.slot * 2 * N2;
    float* G = v.layer(sp.slot, L_GROUND);
    float* C = v.layer(sp.slot, L_GROUNDPATCH);
    if ((N2 & 3) == 0) {
        const float4 a = *reinterpret_cast<const float4*>(sg + cell0);
        const float4 b = *reinterpret_cast<const float4*>(sg + N2 + cell0);
        *reinterpret_cast<float4*>(G + cell0) = a;
        *reinterpret_cast<float4*>(C + cell0) = b;
    } else {
#pragma unroll
        for (int u = 0; u < ROLL_ILP; ++u)
            if (cell0 + u < N2) {
                G[cell0 + u] = sg[cell0 + u];
                C[cell0 + u] = sg[N2 + cell0 + u];
            }
    }
}

// ------------------------------------------------------------------------------------------
// phase 1a: per-point rasterisation front end (insert_cloud up to the accumulate step)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rasterize_point(const View& v, const SlotParams& sp, int i, float& zout) {
    const Const& k = v.k;
    const int N = k.N;
    const size_t base = (size_t)sp.slot * v.pcap;

    float x, y, z;
    int ring;
    if (sp.packed) {
        // packed SoA cloud: four fully coalesced streams, 14 bytes per point
        const int n_pad = (sp.n_points + 7) & ~7;
        x = __ldg(sp.packed + i);
        y = __ldg(sp.packed + n_pad + i);
        z = __ldg(sp.packed + 2 * n_pad + i);
        ring = (int)__ldg(reinterpret_cast<const unsigned short*>(sp.packed + 3 * n_pad) + i);
    } else {
        // coalesced 2 x 16-byte loads of the 32-byte PointXYZIR record
        const uint4* rec = reinterpret_cast<const uint4*>(sp.src + i);
        const uint4 a = __ldg(rec);
        const uint4 b = __ldg(rec + 1);
        x = __uint_as_float(a.x), y = __uint_as_float(a.y), z = __uint_as_float(a.z);
        ring = (int)(b.y & 0xffffu);
    }
    const float ox = sp.ox, oy = sp.oy, oz = sp.oz;

    const float dxo = __fsub_rn(x, ox), dyo = __fsub_rn(y, oy);
    // std::pow(dx, 2.0) + std::pow(dy, 2.0) in double (:223); the same sum feeds hypotf (:170)
    const double sq = __dadd_rn(__dmul_rn((double)dxo, (double)dxo), __dmul_rn((double)dyo, (double)dyo));
    const float sqdist = (float)sq;
    v.dist[base + i] = (float)__dsqrt_rn(sq);  // glibc hypotf: (float) sqrt((double)x*x + (double)y*y)

    uint32_t key = (uint32_t)k.N2;  // sentinel: not rasterised
    uint32_t code = PC_ABSENT << 24;

    const double px = sp.px, py = sp.py;
    int g0, g1;
    grid_index(k, px, py, (double)x, (double)y, g0, g1);
    if (grid_inside(k, px, py, (double)x, (double)y) && g0 >= 0 && g1 >= 0 && g0 < N && g1 < N) {
        const int cell = g0 + g1 * N;
        const bool border = (N <= g0 + 3) || (N <= g1 + 3);  // :167
        if (k.full_layers) atomicAdd(v.raw_i + (size_t)sp.slot * k.N2 + cell, 1);  // pointsRaw, :234
        if (ring > k.max_ring || sqdist < 12.0f) {  // :237
            code = ((border ? PC_IGNORED_BORDER : PC_IGNORED) << 24) | (uint32_t)cell;
        } else {
            const float* G = v.layer(sp.slot, L_GROUND);
            const float* C = v.layer(sp.slot, L_GROUNDPATCH);
            bool outlier = false;
            const float oldg = G[cell];
            if ((double)z < __dsub_rn((double)oldg, 0.2)) {  // :244
                float vx = dxo, vy = dyo, vz = __fsub_rn(z, oz);
                const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
                vx = __fdiv_rn(vx, len);
                vy = __fdiv_rn(vy, len);
                vz = __fdiv_rn(vz, len);
                const double len2 = __dmul_rn((double)len, (double)len);
                if (vz < -0.01f) {
                    // step cap: the reference loop is unbounded for non-finite lengths
                    for (int step = 3; step < (1 << 20); ++step) {
                        const float fs = (float)step;
                        const float sx = __fmul_rn(fs, vx), sy = __fmul_rn(fs, vy), sz = __fmul_rn(fs, vz);
                        const double lhs = __dadd_rn(__dadd_rn(__dmul_rn((double)sx, (double)sx), __dmul_rn((double)sy, (double)sy)),
                                                     __dmul_rn((double)sz, (double)sz));
                        if (!(lhs < len2)) break;
                        int ix, iy;
                        grid_index(k, px, py, (double)__fadd_rn(sx, ox), (double)__fadd_rn(sy, oy), ix, iy);
                        if (ix <= 0 || iy <= 0 || ix >= N - 1 || iy >= N - 1) continue;
                        const int r0 = max(ix - 1, 2), c0 = max(iy - 1, 2);  // :268 (clamped, not centred)
                        float e[9];
#pragma unroll
                        for (int q = 0; q < 9; ++q) e[q] = C[(r0 + q % 3) + (c0 + q / 3) * N];
                        const float bs = tree9(e);
                        if ((double)bs > k.min_outlier_conf && C[ix + iy * N] > 0.01f &&
                            (double)G[ix + iy * N] >= __dadd_rn((double)__fadd_rn(sz, oz), k.outlier_tol)) {
                            outlier = true;
                            break;
                        }
                    }
                }
            }
            if (outlier) {
                code = (PC_OUTLIER << 24) | (uint32_t)cell;
            } else {
                key = (uint32_t)cell;
                code = ((border ? PC_KEPT_BORDER : PC_KEPT) << 24) | (uint32_t)cell;
            }
        }
    }
    v.code[base + i] = code;
    zout = z;
    return key;
}

// Words that travel from k_rasterize to k_scatter:
//   zw[i]   = (z bits, position inside the cell's segment | run-head flag in bit 31)      every point
//   runj[i] = arrival number of the run among the cell's runs | (run length - 1) << 26     run heads only
constexpr uint32_t RUN_LOW26 = (1u << 26) - 1u;

// RASTER_TILE consecutive points per block, RASTER_THREADS threads, thread order == point order inside a round, so
// the 32 lanes of a warp always hold 32 CONSECUTIVE points (point index >> 5 is warp-uniform: the "run id").
// Kept points of a warp that fall into the same cell form one run: the lowest lane claims `len` consecutive
// positions of the cell's segment and the run's directory entry with ONE 64-bit atomicAdd (low word: points of the
// cell, high word: runs of the cell); lane order inside the run is input order.  Runs of one cell never interleave
// (different warps own disjoint index ranges), so input order inside a cell = runs sorted by run id -- which
// k_cell_stats restores from the (few) directory entries without any sort pass over the points.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(RASTER_THREADS, MIN_BLOCKS) k_rasterize(View v, const SlotParams* __restrict__ batch) {
    const SlotParams& sp = batch[blockIdx.y];
    const int n = sp.n_points;
    const int tile0 = blockIdx.x * RASTER_TILE;
    if (tile0 >= n) return;
    const size_t base = (size_t)sp.slot * v.pcap;
    unsigned long long* cnt = v.cnt64 + (size_t)sp.slot * v.k.N2;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int r = 0; r < RASTER_TILE / RASTER_THREADS; ++r) {
        const int i = tile0 + r * RASTER_THREADS + threadIdx.x;
        if (__all_sync(0xffffffffu, i >= n)) break;
        uint32_t key = (uint32_t)v.k.N2;
        float z = 0.0f;
        if (i < n) key = rasterize_point(v, sp, i, z);
        const bool kept = key != (uint32_t)v.k.N2;
        // lanes without a kept point get a private pseudo key (never equal to a cell id)
        const uint32_t peers = __match_any_sync(0xffffffffu, kept ? key : (0x80000000u | (uint32_t)lane));
        const int leader = __ffs(peers) - 1;
        const int len = __popc(peers);
        const bool head = kept && lane == leader;
        unsigned long long old = 0ull;
        if (head) old = atomicAdd(cnt + key, (1ull << 32) | (unsigned long long)len);
        const int start = __shfl_sync(0xffffffffu, (int)(uint32_t)old, leader);
        if (i < n) {
            const uint32_t w = kept ? ((uint32_t)(start + __popc(peers & lt_mask)) | (head ? 0x80000000u : 0u)) : 0u;
            v.zw[base + i] = make_uint2(__float_as_uint(z), w);
            if (head) v.runj[base + i] = ((uint32_t)(old >> 32) & RUN_LOW26) | ((uint32_t)(len - 1) << 26);
        }
    }
}

// ------------------------------------------------------------------------------------------
// phase 1b: segment starts (exclusive scan of the per-cell point counts), then every kept z goes to its
// position and every run head files the run in the cell's directory
// ------------------------------------------------------------------------------------------
// Count classes of the cell worklist: exact up to 32 points, then steps of 1/4 octave (cells of one class differ by
// less than 20 % in their walk length); heaviest class first.
constexpr int WL_CLASSES = 64;
__device__ __forceinline__ int worklist_class(int c) {
    if (c <= 32) return c;                                  // 1 .. 32 exact (0 never enters the list)
    const int e = 31 - __clz(c);                            // floor(log2 c) >= 5
    const int q = (c >> (e - 2)) & 3;                       // two bits below the leading one
    const int k = 33 + (e - 5) * 4 + q;
    return k < WL_CLASSES ? k : WL_CLASSES - 1;
}

// Segment starts and worklist in two short, fully parallel launches over tiles of CELL_TILE cells (a single block
// per scan would be a long serial chain):
//   k_cell_tiles   per tile: kept points and cells per count class            -> cell_agg[scan][tile][0 .. WL_CLASSES]
//   k_cell_place   per tile: prefix over the earlier tiles' aggregates, local exclusive scan -> cellstart; every
//                  non-empty cell goes to the scan's worklist, grouped by class (heaviest class first) so that the 32
//                  cells a warp of k_cell_stats walks have (almost) the same length.  The order inside a class is
//                  arbitrary -- every cell's result is independent of it.
constexpr int CELL_TILE = 4096;      // CT_THREADS threads x CT_PER cells
constexpr int CT_THREADS = 256, CT_PER = 16;
constexpr int AGG_STRIDE = WL_CLASSES + 1;   // [0, WL_CLASSES): cells per class, [WL_CLASSES]: kept points

// 16 consecutive cells per thread: eight 16-byte loads of the (runs << 32 | points) counters; c = points, r (optional) = runs
__device__ __forceinline__ void load_tile_counts(const unsigned long long* cnt64, int N2, int cell0, int c[CT_PER], int* r = nullptr) {
    if (cell0 + CT_PER <= N2 && (((size_t)cnt64 & 15) == 0) && (cell0 & 1) == 0) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(cnt64 + cell0);
#pragma unroll
        for (int q = 0; q < CT_PER / 2; ++q) {
            const ulonglong2 v = p[q];
            c[2 * q] = (int)(uint32_t)v.x;
            c[2 * q + 1] = (int)(uint32_t)v.y;
            if (r) {
                r[2 * q] = (int)(v.x >> 32);
                r[2 * q + 1] = (int)(v.y >> 32);
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < CT_PER; ++q) {
            const unsigned long long v = (cell0 + q < N2) ? cnt64[cell0 + q] : 0ull;
            c[q] = (int)(uint32_t)v;
            if (r) r[q] = (int)(v >> 32);
        }
    }
}

__global__ void __launch_bounds__(CT_THREADS) k_cell_tiles(View v, const SlotParams* __restrict__ batch) {
    __shared__ int s_cls[AGG_STRIDE];
    const SlotParams& sp = batch[blockIdx.y];
    const int N2 = v.k.N2, tid = threadIdx.x, lane = tid & 31;
    if (tid < AGG_STRIDE) s_cls[tid] = 0;
    __syncthreads();
    int c[CT_PER];
    load_tile_counts(v.cnt64 + (size_t)sp.slot * N2, N2, blockIdx.x * CELL_TILE + tid * CT_PER, c);
    int sum = 0;
#pragma unroll
    for (int q = 0; q < CT_PER; ++q) {
        sum += c[q];
        if (c[q] > 0) atomicAdd(&s_cls[worklist_class(c[q])], 1);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0 && sum) atomicAdd(&s_cls[WL_CLASSES], sum);
    __syncthreads();
    if (tid < AGG_STRIDE) v.cell_agg[((size_t)sp.slot * v.cell_tiles + blockIdx.x) * AGG_STRIDE + tid] = s_cls[tid];
}

__global__ void __launch_bounds__(CT_THREADS) k_cell_place(View v, const SlotParams* __restrict__ batch) {
    __shared__ int s_cur[AGG_STRIDE];   // worklist cursor of every class for this tile; [WL_CLASSES]: first segment start of the tile
    __shared__ int s_warp[CT_THREADS / 32];
    const SlotParams& sp = batch[blockIdx.y];
    const int N2 = v.k.N2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = v.cell_tiles, tile = blockIdx.x;
    const int* agg = v.cell_agg + (size_t)sp.slot * tiles * AGG_STRIDE;
    const size_t off = (size_t)sp.slot * N2;
    const int cell0 = tile * CELL_TILE + tid * CT_PER;
    int c[CT_PER], rn[CT_PER];
    load_tile_counts(v.cnt64 + off, N2, cell0, c, rn);
    // one warp: per class (2 per lane) the total over all tiles and the part of the earlier tiles
    if (warp == 0) {
        int tot[2] = {0, 0}, pre[2] = {0, 0};
        const int k0 = WL_CLASSES - 1 - 2 * lane, k1 = k0 - 1;   // lane 0: classes 63, 62 ... lane 31: classes 1, 0
        for (int p = 0; p < tiles; ++p) {
            const int a0 = agg[p * AGG_STRIDE + k0], a1 = agg[p * AGG_STRIDE + k1];
            tot[0] += a0;
            tot[1] += a1;
            if (p < tile) {
                pre[0] += a0;
                pre[1] += a1;
            }
        }
        const int incl = warp_inclusive_scan(tot[0] + tot[1]);   // classes in descending order: heaviest first
        s_cur[k0] = incl - tot[0] - tot[1] + pre[0];
        s_cur[k1] = incl - tot[1] + pre[1];
        int pts = 0, pts_all = 0;
        for (int p = lane; p < tiles; p += 32) {
            const int a = agg[p * AGG_STRIDE + WL_CLASSES];
            pts_all += a;
            if (p < tile) pts += a;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            pts += __shfl_xor_sync(0xffffffffu, pts, d);
            pts_all += __shfl_xor_sync(0xffffffffu, pts_all, d);
        }
        if (lane == 0) s_cur[WL_CLASSES] = pts;
        if (tile == 0 && lane == 31) {
            v.wl_count[2 * sp.slot] = incl;          // non-empty cells of the scan
            v.wl_count[2 * sp.slot + 1] = pts_all;   // kept points of the scan (= end of the last segment)
        }
    }
    int sum = 0;
#pragma unroll
    for (int q = 0; q < CT_PER; ++q) sum += c[q];
    const int incl = warp_inclusive_scan(sum);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int run = s_cur[WL_CLASSES] + incl - sum;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    // a worklist entry carries everything k_cell_stats needs to start on the cell: (cell, points, runs, segment start)
    int4* wl = v.worklist + off;
    int cs[CT_PER];
#pragma unroll
    for (int q = 0; q < CT_PER; ++q) {
        cs[q] = run;
        if (c[q] > 0) wl[atomicAdd(&s_cur[worklist_class(c[q])], 1)] = make_int4(cell0 + q, c[q], rn[q], run);
        run += c[q];
    }
    if (cell0 + CT_PER <= N2 && (N2 & 3) == 0) {
        int4* dst = reinterpret_cast<int4*>(v.cellstart + off + cell0);
#pragma unroll
        for (int q = 0; q < CT_PER / 4; ++q) dst[q] = make_int4(cs[4 * q], cs[4 * q + 1], cs[4 * q + 2], cs[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < CT_PER; ++q)
            if (cell0 + q < N2) v.cellstart[off + cell0 + q] = cs[q];
    }
}

constexpr int SCATTER_ILP = 4;

// Directory entry of a run, at rundir[cellstart + arrival number] (a cell has at most as many runs as points):
//   x = run id (point index >> 5), y = first position inside the segment | (length - 1) << 26
__global__ void __launch_bounds__(256) k_scatter(View v, const SlotParams* __restrict__ batch) {
    const SlotParams& sp = batch[blockIdx.y];
    const int n = sp.n_points;
    const int i0 = blockIdx.x * (256 * SCATTER_ILP) + threadIdx.x;
    if (i0 >= n) return;
    const size_t base = (size_t)sp.slot * v.pcap;
    const int* cellstart = v.cellstart + (size_t)sp.slot * v.k.N2;
    float* zs = v.zsorted + base;
    uint2* dir = v.rundir + base;
    uint32_t code[SCATTER_ILP];
    uint2 zw[SCATTER_ILP];
    int cs[SCATTER_ILP];
    uint32_t jl[SCATTER_ILP];
#pragma unroll
    for (int u = 0; u < SCATTER_ILP; ++u) {
        const int i = i0 + u * 256;
        code[u] = i < n ? v.code[base + i] : (PC_ABSENT << 24);
        zw[u] = i < n ? v.zw[base + i] : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < SCATTER_ILP; ++u) {
        const uint32_t cls = code[u] >> 24;
        const bool kept = cls == PC_KEPT || cls == PC_KEPT_BORDER;
        cs[u] = kept ? cellstart[code[u] & 0xffffffu] : -1;
        jl[u] = (kept && (zw[u].y & 0x80000000u)) ? v.runj[base + i0 + u * 256] : 0u;
    }
#pragma unroll
    for (int u = 0; u < SCATTER_ILP; ++u) {
        if (cs[u] < 0) continue;
        const uint32_t p = zw[u].y & 0x7fffffffu;
        zs[cs[u] + (int)p] = __uint_as_float(zw[u].x);
        if (zw[u].y & 0x80000000u) dir[cs[u] + (int)(jl[u] & RUN_LOW26)] = make_uint2((uint32_t)((i0 + u * 256) >> 5), p | (jl[u] & ~RUN_LOW26));
    }
}

// ------------------------------------------------------------------------------------------
// phase 1c: per-cell sequential statistics (the accumulate step of insert_cloud, :282-309,
// in input order) + variance (:323).  One thread per cell.
// ------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 256;

template <bool FULL>
__global__ void __launch_bounds__(CS_THREADS, FULL ? 2 : 4) k_cell_stats(View v, const SlotParams* __restrict__ batch) {
    const SlotParams& sp = batch[blockIdx.y];
    const Const& k = v.k;
    const size_t coff = (size_t)sp.slot * k.N2;
    const int gt = blockIdx.x * CS_THREADS + threadIdx.x;
    const int kept_total = v.wl_count[2 * sp.slot + 1];
    // (1) four consecutive cells per thread, coalesced: the per-scan resets every cell needs, and the results of an
    //     EMPTY cell (count 0, m2 / (0 + FLT_MIN) = 0, min = FLT_MAX, :61-75,323).  Empty <=> zero-length segment
    //     (the counters themselves are being consumed by other threads of this launch).
    for (int t = gt * 4; t < k.N2; t += gridDim.x * CS_THREADS * 4) {
        int cs[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) cs[q] = (t + q < k.N2) ? v.cellstart[coff + t + q] : kept_total;
        if ((k.N2 & 3) == 0) {
            *reinterpret_cast<float4*>(v.layer(sp.slot, L_OBSTACLES) + t) = make_float4(0.f, 0.f, 0.f, 0.f);  // map["points"].setConstant(0.0), :147
            if (FULL) {
                const int4 r = *reinterpret_cast<const int4*>(v.raw_i + coff + t);
                *reinterpret_cast<float4*>(v.layer(sp.slot, L_RAW) + t) = make_float4((float)r.x, (float)r.y, (float)r.z, (float)r.w);
                *reinterpret_cast<int4*>(v.raw_i + coff + t) = make_int4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int cell = t + q;
            if (cell >= k.N2) break;
            if ((k.N2 & 3) != 0) {
                v.layer(sp.slot, L_OBSTACLES)[cell] = 0.0f;
                if (FULL) {
                    v.layer(sp.slot, L_RAW)[cell] = (float)v.raw_i[coff + cell];
                    v.raw_i[coff + cell] = 0;
                }
            }
            if (cs[q + 1] == cs[q]) {
                v.layer(sp.slot, L_COUNT)[cell] = 0.0f;
                v.layer(sp.slot, L_VARIANCE)[cell] = 0.0f;
                v.layer(sp.slot, L_MINH)[cell] = FLT_MAX;
                if (FULL) {
                    v.layer(sp.slot, L_M2)[cell] = 0.0f;
                    v.layer(sp.slot, L_MEAN)[cell] = 0.0f;
                    v.layer(sp.slot, L_GCAND)[cell] = 0.0f;
                    v.layer(sp.slot, L_PLANEDIST)[cell] = 0.0f;
                    v.layer(sp.slot, L_MAXH)[cell] = FLT_MIN;
                }
            }
        }
    }
    // (2) the worklist: non-empty cells grouped by count class (k_cell_place), so the 32 cells of a warp need (almost)
    //     the same number of sequential steps; grid-stride (a warp's entries stay 32 consecutive ones)
    const int wl_n = v.wl_count[2 * sp.slot];
    const float oz = sp.oz;
    for (int t = gt; t < wl_n; t += gridDim.x * CS_THREADS) {
    const int4 entry = v.worklist[coff + t];   // one load instead of a chain of three dependent ones
    const int cell = entry.x, cnt = entry.y, runs = entry.z, cstart = entry.w;
    const float* zs = v.zsorted + (size_t)sp.slot * v.pcap + cstart;
    uint2* dir = v.rundir + (size_t)sp.slot * v.pcap + cstart;
    // the kernel is bound by the latency of dependent global loads: pull the first lines of the segment and of the
    // run directory towards L1 right away
    asm volatile("prefetch.global.L1 [%0];" ::"l"(zs));
    if (cnt > 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(zs + 32));
    if (runs > 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(dir));

    float n = 0.0f, mean = 0.0f, m2 = 0.0f;
    float mn = FLT_MAX;
    float mx = FLT_MIN, gc = 0.0f, pdm = 0.0f;  // dead layers (FULL only)
    // one accumulate step of insert_cloud (:282-309)
    auto step = [&](float z) {
        const float pd = __fsub_rn(z, oz);  // planeDist, :295
        if (FULL) gc = (float)__ddiv_rn((double)__fadd_rn(z, __fmul_rn(n, gc)), __dadd_rn((double)n, 1.0));  // :296
        if (mean == 0.0f) mean = pd;  // :298-299
        if (!(pd != pd)) {            // :300
            const float delta = __fsub_rn(pd, mean);
            mean = __fadd_rn(mean, __fdiv_rn(delta, __fadd_rn(n, 1.0f)));
            if (FULL) pdm = (float)__ddiv_rn((double)__fadd_rn(pd, __fmul_rn(n, pdm)), __dadd_rn((double)n, 1.0));
            m2 = __fadd_rn(m2, __fmul_rn(delta, __fsub_rn(pd, mean)));
        }
        if (FULL) mx = (mx < z) ? z : mx;                   // std::max(maxHeight, z)
        const float zl = __fsub_rn(z, 0.0001f);
        mn = (zl < mn) ? zl : mn;                           // std::min(minHeight, z - 0.0001f)
        n = __fadd_rn(n, 1.0f);
    };

    // The segment is a sequence of runs (k_rasterize), each internally in input order; input order of the cell = runs
    // by ascending id.  The cell's directory (one entry per run, contiguous) is sorted by id first: up to 8 runs in
    // registers (odd-even transposition), more -- cells crossed by many rings: walls, vehicles -- by an in-place Shell
    // sort (gaps 57 / 23 / 10 / 4 / 1).
    if (runs > 1 && runs <= 8) {
        uint2 e[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) e[q] = (q < runs) ? dir[q] : make_uint2(0xffffffffu, 0u);
        bool moved = false;
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
#pragma unroll
            for (int q = pass & 1; q + 1 < 8; q += 2) {
                if (e[q + 1].x < e[q].x) {
                    const uint2 tmp = e[q];
                    e[q] = e[q + 1];
                    e[q + 1] = tmp;
                    moved = true;
                }
            }
        }
        if (moved) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < runs) dir[q] = e[q];
        }
    } else if (runs > 8) {
        const int gaps[5] = {57, 23, 10, 4, 1};
#pragma unroll
        for (int gi = 0; gi < 5; ++gi) {
            const int gap = gaps[gi];
            if (gap >= runs) continue;
            for (int a = gap; a < runs; ++a) {
                const uint2 e = dir[a];
                int b = a;
                while (b >= gap) {
                    const uint2 f = dir[b - gap];
                    if (f.x <= e.x) break;
                    dir[b] = f;
                    b -= gap;
                }
                if (b != a) dir[b] = e;
            }
        }
    }
    // Streaming walk: chunks of up to 8 consecutive heights of the current run; the loads of the next chunk (and the
    // directory entry of the next run) are in flight while the current chunk goes through the sequential recurrence.
    {
        int rj = 0, pb = 0, pe = cnt;   // runs == 1: the whole segment
        if (runs > 1) {
            const uint2 e0 = dir[0];
            pb = (int)(e0.y & RUN_LOW26);
            pe = pb + (int)(e0.y >> 26) + 1;
        }
        float zb[8], zn[8];
        int mb = 0, mnx = 0;
        auto fetch = [&](float* dst, int& m) {
            m = min(8, pe - pb);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = (q < m) ? zs[pb + q] : 0.0f;
            pb += m;
            if (pb == pe && ++rj < runs) {
                const uint2 en = dir[rj];
                pb = (int)(en.y & RUN_LOW26);
                pe = pb + (int)(en.y >> 26) + 1;
                if (((pb + 8) & ~31) != (pb & ~31)) asm volatile("prefetch.global.L1 [%0];" ::"l"(zs + pb + 8));
            }
        };
        fetch(zb, mb);
        while (mb > 0) {
            fetch(zn, mnx);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < mb) step(zb[q]);
#pragma unroll
            for (int q = 0; q < 8; ++q) zb[q] = zn[q];
            mb = mnx;
        }
    }
    v.cnt64[coff + cell] = 0ull;                   // consumed: zero again for the next scan
    v.layer(sp.slot, L_COUNT)[cell] = n;
    v.layer(sp.slot, L_VARIANCE)[cell] = __fdiv_rn(m2, __fadd_rn(n, FLT_MIN));
    v.layer(sp.slot, L_MINH)[cell] = mn;
    if (FULL) {
        v.layer(sp.slot, L_M2)[cell] = m2;
        v.layer(sp.slot, L_MEAN)[cell] = mean;
        v.layer(sp.slot, L_GCAND)[cell] = gc;
        v.layer(sp.slot, L_PLANEDIST)[cell] = pdm;
        v.layer(sp.slot, L_MAXH)[cell] = mx;
    }
    }
}

// ------------------------------------------------------------------------------------------
// phase 2: ground patch detection stencil (detect_ground_patches / detect_ground_patch<S>)
// ------------------------------------------------------------------------------------------
constexpr int DT_X = 32, DT_Y = 8, DT_H = 2;
constexpr int DT_W = DT_X + 2 * DT_H;   // 36
constexpr int DT_R = DT_Y + 2 * DT_H;   // 12

// confidence decay of interpolate_cell (:464): std::max(c - c / decrease_factor, 0.001) in fp64
__device__ __forceinline__ float decay_confidence(const Const& k, float occ) {
    // Confidences sit at the 0.001 floor in most of the map; from there (and below) the result is the floor
    // again (host-checked for the configured factor, Const::decay_floor_ok), which skips the fp64 division.
    if (k.decay_floor_ok && occ <= 0.001f) return 0.001f;
    const double o = (double)occ;
    const double dec = __dsub_rn(o, __ddiv_rn(o, k.dec_factor));
    return (float)((dec < 0.001) ? 0.001 : dec);
}

// Fused into k_detect: every cell is copied to its home slot(s) as soon as its final G, C are known.
__device__ __forceinline__ void skew_store_cell(const View& v, const SlotParams& sp, int cell, int x, int y, float g, float c, bool far) {
    const Const& k = v.k;
    const int4 home = __ldg(reinterpret_cast<const int4*>(v.skew.cell_home) + cell);
    if (home.x < 0) return;
    const int cidx = k.N / 2 - 1;
    if (x == cidx && y == cidx) {  // spiral_ground_interpolation :405,411 (the normal layers get it in k_spiral_skew)
        g = sp.base_z_f;
        c = 1.0f;
    }
    // :463: beyond minDistSquared (`far`, from the per-cell table) the visit stores the decayed confidence (:464)
    const float d1 = far ? decay_confidence(k, c) : -1.0f;
    float2* SK = v.skew.sk + (size_t)sp.slot * v.skew.slots;
    float* SD = v.skew.sd + (size_t)sp.slot * v.skew.slots;
    const float2 gc = make_float2(g, c);
    SK[home.x] = gc;
    SD[home.x] = d1;
    if (home.y >= 0) {
        SK[home.y] = gc;
        SD[home.y] = far ? decay_confidence(k, d1) : -1.0f;  // second visit of a ring corner
    }
    if (home.z >= 0) SK[home.z] = gc;
    if (home.w >= 0) SK[home.w] = gc;
}

// Per-cell quantities of detect_ground_patches that depend only on the grid and the configuration, computed
// once (gg_create / gg_set_config) with the very operations of the per-scan code they replace:
//   x: max(3, floor(threshold * S * expectedPoints))  (:364; integer valued, exact in float, capped at 2^24)
//   y: variance threshold (:369)      z: expectedPoints      w: flags
constexpr int DTF_S5 = 1, DTF_FAR = 2, DTF_INNER = 4;

__global__ void k_build_detect_table(View v, float4* __restrict__ tab) {
    const Const& k = v.k;
    const int N = k.N;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= k.N2) return;
    const int i = cell % N, j = cell / N;
    const double di = __dsub_rn((double)i, (double)N / 2.0), dj = __dsub_rn((double)j, (double)N / 2.0);
    const float sqdist = (float)__dmul_rn(__dadd_rn(__dmul_rn(di, di), __dmul_rn(dj, dj)), k.res_sq);  // :332,356
    const float e = v.expected[cell];
    int flags = 0;
    if (!((double)sqdist <= k.psc_sq)) flags |= DTF_S5;
    // union of the four sections, :325-328: first index in [2, 2 * (N / 2) - 2) (the upper half starts at N / 2 and has
    // N / 2 - 2 entries: one row short of N - 2 when N is odd), second index in [2, N - 2)
    if (i >= 2 && i < 2 * (N / 2) - 2 && j >= 2 && j < N - 2) flags |= DTF_INNER;
    const int cidx = N / 2 - 1;
    const float fx = __fsub_rn((float)i, (float)cidx), fy = __fsub_rn((float)j, (float)cidx);
    if (__dmul_rn(__dadd_rn(__dmul_rn((double)fx, (double)fx), __dmul_rn((double)fy, (double)fy)), k.res_sq) > 12.0) flags |= DTF_FAR;  // :463
    const double S = (flags & DTF_S5) ? 5.0 : 3.0;
    double need = floor(__dmul_rn(__dmul_rn(k.gp_thresh, S), (double)e));
    need = (need < 3.0) ? 3.0 : need;                 // NaN -> NaN: the comparison below stays false, as in the reference
    if (need > 16777216.0) need = 16777216.0;         // patch sums are counts far below 2^24
    const double a = __dmul_rn((double)sqdist, k.df_sq);
    const double m = (a < k.mdf_sq) ? k.mdf_sq : a;
    const float vt = (float)((k.mdf10_sq < m) ? k.mdf10_sq : m);
    tab[cell] = make_float4((float)need, vt, e, __int_as_float(flags));
}

// returns true when (g, c) changed
// sPV / sPM hold the products count * variance and count * minHeight of every tile entry (the very fmul the
// reference's cwiseProduct performs, done once per entry instead of once per window position)
template <int S>
__device__ __forceinline__ bool detect_patch(const Const& k, const float (*sP)[DT_W], const float (*sPV)[DT_W], const float (*sPM)[DT_W],
                                             const float (*sM)[DT_W], int li, int lj, float variance, float need, float vt, float e, float& g,
                                             float& c) {
    constexpr int H = S / 2;
    const int r0 = li - H, c0 = lj - H;  // block origin in the shared tile (row index = i, col = j)
    const float psum = TreeSum<0, S * S>::run([&](int q) { return sP[c0 + q / S][r0 + q % S]; });
    const float oc = c, og = g;

    // early skipping of (almost) empty areas, :364 (both sides integer valued: the float compare is the double one)
    if (psum < need) return false;

    float localmin = sM[c0][r0];
#pragma unroll
    for (int q = 1; q < S * S; ++q) {
        const float t = sM[c0 + q / S][r0 + q % S];
        localmin = (t < localmin) ? t : localmin;
    }
    const float maxVar = (sP[lj][li] >= k.pc_var_thresh_f)
                             ? variance
                             : __fdiv_rn(TreeSum<0, S * S>::run([&](int q) { return sPV[c0 + q / S][r0 + q % S]; }), psum);
    const float groundlevel = __fdiv_rn(TreeSum<0, S * S>::run([&](int q) { return sPM[c0 + q / S][r0 + q % S]; }), psum);
    const float gd = __fmul_rn(__fsub_rn(groundlevel, og), __fmul_rn(2.0f, oc));
    const float groundDiff = (gd < 1.0f) ? 1.0f : gd;  // std::max(gd, 1.0f)

    // do not update known high confidence estimations upward, :379
    if ((double)oc > 0.5 && (double)groundlevel >= __dadd_rn((double)og, k.outlier_tol)) return false;

    if ((double)vt > __dmul_rn((double)maxVar, (double)maxVar) && maxVar > 0.0f &&
        (double)psum > __dmul_rn((double)__fmul_rn(__fmul_rn(groundDiff, e), (float)S), k.gp_thresh)) {
        const double ncd = __ddiv_rn((double)psum, k.occ_factor);
        const float nc = (float)((1.0 < ncd) ? 1.0 : ncd);  // std::min(ncd, 1.0)
        const float num = __fadd_rn(__fmul_rn(groundlevel, nc), __fmul_rn(__fmul_rn(oc, og), 2.0f));
        const float den = __fadd_rn(nc, __fmul_rn(oc, 2.0f));
        g = __fdiv_rn(num, den);
        const double cd = __ddiv_rn(__dadd_rn(__ddiv_rn((double)psum, k.occ_factor2), (double)oc), 2.0);
        c = (float)((1.0 < cd) ? 1.0 : cd);
        return true;
    }
    if (localmin < og) {
        g = localmin;
        const float t = __fadd_rn(oc, 0.1f);
        c = (0.5f < t) ? 0.5f : t;  // std::min(oc + 0.1f, 0.5f)
        return true;
    }
    return false;
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(DT_X* DT_Y, MIN_BLOCKS) k_detect_ldg(View v, const SlotParams* __restrict__ batch) {
    __shared__ float sP[DT_R][DT_W], sPV[DT_R][DT_W], sPM[DT_R][DT_W], sM[DT_R][DT_W];
    const SlotParams& sp = batch[blockIdx.z];
    const Const& k = v.k;
    const int N = k.N;
    const int i0 = blockIdx.x * DT_X, j0 = blockIdx.y * DT_Y;
    // one 64-bit base per scan, 32-bit offsets below it (all layers of a slot span far less than 2^31 floats)
    float* const L0 = v.layer(sp.slot, 0);
    const int N2 = k.N2;
    const float* P = L0 + L_COUNT * N2;
    const float* V = L0 + L_VARIANCE * N2;
    const float* M = L0 + L_MINH * N2;
    // tile + halo (36 x 12): a thread fetches column tx of rows ty and ty + 8 (ty < 4) and, for tx < 4, the halo
    // columns 32 + tx of the same rows; four fixed positions o, o + 32, o + 8 N, o + 8 N + 32, every global load issued
    // before the first shared store.  Outside the map: count 0, variance 0, min FLT_MAX (products 0), as before.
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gi = i0 - DT_H + tx, gj = j0 - DT_H + ty;
    const bool col0 = gi >= 0 && gi < N, col1 = tx < 2 * DT_H && gi + DT_X < N;
    const bool row0 = gj >= 0 && gj < N, row1 = ty < 2 * DT_H && gj + DT_Y < N;
    const int o = gi + gj * N;
    const bool in[4] = {col0 && row0, col1 && row0, col0 && row1, col1 && row1};
    const int off[4] = {o, o + DT_X, o + DT_Y * N, o + DT_Y * N + DT_X};
    float tp[4], tv[4], tm[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        tp[u] = in[u] ? P[off[u]] : 0.0f;
        tv[u] = in[u] ? V[off[u]] : 0.0f;
        tm[u] = in[u] ? M[off[u]] : FLT_MAX;
    }
    const int i = i0 + tx, j = j0 + ty;
    const bool live = i < N && j < N;
    const int cell = i + j * N;
    float g = 0.0f, c = 0.0f, variance = 0.0f;
    float4 tb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (live) {
        g = L0[L_GROUND * N2 + cell];
        c = L0[L_GROUNDPATCH * N2 + cell];
        variance = V[cell];
        tb = __ldg(v.detect_tab + cell);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const bool mine = ((u & 1) == 0 || tx < 2 * DT_H) && ((u & 2) == 0 || ty < 2 * DT_H);
        if (mine) {
            const int li = (u & 1) ? DT_X + tx : tx, lj = (u & 2) ? ty + DT_Y : ty;
            sP[lj][li] = tp[u];
            sPV[lj][li] = __fmul_rn(tp[u], tv[u]);
            sPM[lj][li] = __fmul_rn(tp[u], tm[u]);
            sM[lj][li] = tm[u];
        }
    }
    __syncthreads();
    if (!live) return;
    const int flags = __float_as_int(tb.w);
    if (flags & DTF_INNER) {
        const int li = threadIdx.x + DT_H, lj = threadIdx.y + DT_H;
        const bool changed = (flags & DTF_S5) ? detect_patch<5>(k, sP, sPV, sPM, sM, li, lj, variance, tb.x, tb.y, tb.z, g, c)
                                              : detect_patch<3>(k, sP, sPV, sPM, sM, li, lj, variance, tb.x, tb.y, tb.z, g, c);
        if (changed) {
            L0[L_GROUND * N2 + cell] = g;
            L0[L_GROUNDPATCH * N2 + cell] = c;
        }
    }
    if (v.skew.sk) {
        skew_store_cell(v, sp, cell, i, j, g, c, (flags & DTF_FAR) != 0);
    } else if (v.spiral_recs) {
        // Decayed confidence for the spiral sweep, taken off its sequential critical path: the
        // confidence of a cell only changes at its own visit(s), so decay(C) after patch
        // detection is exactly what the (first) visit will store; ring corners (i == j) are
        // visited twice and need the second decay as well.
        const float d1 = decay_confidence(k, c);
        float* D1 = v.roll_scratch + (size_t)sp.slot * 2 * k.N2;
        D1[cell] = d1;
        if (i == j) D1[k.N2 + cell] = decay_confidence(k, d1);
    }
}

// ---- TMA-staged variant ------------------------------------------------------------------
// The tile (+halo) of the three per-scan layers arrives by three cp.async.bulk.tensor loads (one elected thread, one
// mbarrier); cells outside the map are zero-filled by the TMA unit, which is harmless because only cells [2, N-2)
// compute and their windows stay inside the map (:325-337).  The innermost start coordinate of a TMA box must be
// 16-byte aligned (measured on B200: a start at i0 - 2 raises "illegal instruction", tools/tma_probe4.cu), so the
// tile keeps a halo of 4 cells in i: box = 40 x 12 x 1 at (i0 - 4, j0 - 2).  All layers of all slots form ONE 3-D tensor
// (i, j, plane = slot * n_layers + layer), so one descriptor serves every scan.  While the tile is in flight the
// threads fetch their own cell's G, C and table entry.
//
// Window sums.  The three block reductions of :359,374-375 are Eigen's binary-split tree over the column-major
// coefficients e[k] = block(k % S, k / S).  Most inner nodes of that tree are runs of 2 or 3
// consecutive rows of ONE column: pair(r, c) = q(r, c) + q(r+1, c) and triple(r, c) = q(r, c) + pair(r+1, c) -- the
// very additions of the tree, so they can be computed once per tile position and shared by every window that contains
// them (a 5x5 tree then needs 14 shared loads and 13 additions instead of 25 and 24).  The point-count sum is a sum of
// small integers (exact in fp32 in any order) and the block minimum is order-free: both are separable (column
// partials of 3 / 5 rows).  Tiles without a single candidate cell (psum >= need nowhere) skip the product stage.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

constexpr int DT_HX = 4;                 // halo in i of the TMA tile (alignment of the box start)
constexpr int DT_WT = DT_X + 2 * DT_HX;  // 40
constexpr int DT_WC = DT_WT - 4;         // 36: columns that can be the first row of a 3 / 5-row window
constexpr int DT_WB = DT_WT - 2;         // 38: columns whose product is read by some window

struct __align__(128) DetectRaw {   // one pipeline stage: the three TMA destinations
    float P[DT_R][DT_WT];   // kept points per cell
    float V[DT_R][DT_WT];   // variance
    float M[DT_R][DT_WT];   // min height
};
struct __align__(128) DetectTile {
    DetectRaw raw[2];       // double buffer: the tile after next is in flight while this one is computed
    float C3[DT_R][DT_WT], C5[DT_R][DT_WT];       // sum of P over rows r .. r+2 / r .. r+4 of the column (exact integers)
    float N3[DT_R][DT_WT], N5[DT_R][DT_WT];       // min of M over the same rows
    float QV[DT_R][DT_WT], PV2[DT_R][DT_WT], TV[DT_R][DT_WT];   // q = P * V; pair; triple
    float QM[DT_R][DT_WT], PM2[DT_R][DT_WT], TM[DT_R][DT_WT];   // q = P * M; pair; triple
};
static_assert(offsetof(DetectRaw, V) % 128 == 0 && offsetof(DetectRaw, M) % 128 == 0 && sizeof(DetectRaw) % 128 == 0, "TMA destinations are 128-byte aligned");   // 12 * 40 * 4 = 1920 = 15 * 128

// Eigen tree of a 5x5 / 3x3 window from the shared partials; (r0, c0) = window origin (row = i, col = j)
__device__ __forceinline__ float tree25(const float (*Q)[DT_WT], const float (*P2)[DT_WT], const float (*T)[DT_WT], int r0, int c0) {
#define QQ(r, c) Q[c0 + (c)][r0 + (r)]
#define PP(r, c) P2[c0 + (c)][r0 + (r)]
#define TT(r, c) T[c0 + (c)][r0 + (r)]
    const float s0_6 = __fadd_rn(TT(0, 0), __fadd_rn(QQ(3, 0), __fadd_rn(QQ(4, 0), QQ(0, 1))));          // e0..e2 | e3 + (e4 + e5)
    const float s6_6 = __fadd_rn(TT(1, 1), __fadd_rn(QQ(4, 1), PP(0, 2)));                                  // e6..e8 | e9 + (e10 + e11)
    const float s12_6 = __fadd_rn(TT(2, 2), TT(0, 3));                                                      // e12..e14 | e15..e17
    const float s18_7 = __fadd_rn(__fadd_rn(QQ(3, 3), __fadd_rn(QQ(4, 3), QQ(0, 4))), __fadd_rn(PP(1, 4), PP(3, 4)));  // e18 + (e19 + e20) | (e21 + e22) + (e23 + e24)
    return __fadd_rn(__fadd_rn(s0_6, s6_6), __fadd_rn(s12_6, s18_7));
#undef TT
}
__device__ __forceinline__ float tree9s(const float (*Q)[DT_WT], const float (*P2)[DT_WT], int r0, int c0) {
    const float s0_4 = __fadd_rn(PP(0, 0), __fadd_rn(QQ(2, 0), QQ(0, 1)));                  // (e0 + e1) + (e2 + e3)
    const float s4_5 = __fadd_rn(PP(1, 1), __fadd_rn(QQ(0, 2), PP(1, 2)));                  // (e4 + e5) + (e6 + (e7 + e8))
    return __fadd_rn(s0_4, s4_5);
#undef QQ
#undef PP
}

// the decision part of detect_ground_patch<S> (:364-394) on the window quantities; returns true when (g, c) changed
template <int S>
__device__ __forceinline__ bool detect_decide(const Const& k, float psum, float localmin, float sumPV, float sumPM, float centerP, float variance, float vt,
                                              float e, float& g, float& c) {
    const float oc = c, og = g;
    const float maxVar = (centerP >= k.pc_var_thresh_f) ? variance : __fdiv_rn(sumPV, psum);
    const float groundlevel = __fdiv_rn(sumPM, psum);
    const float gd = __fmul_rn(__fsub_rn(groundlevel, og), __fmul_rn(2.0f, oc));
    const float groundDiff = (gd < 1.0f) ? 1.0f : gd;  // std::max(gd, 1.0f)
    // do not update known high confidence estimations upward, :379
    if ((double)oc > 0.5 && (double)groundlevel >= __dadd_rn((double)og, k.outlier_tol)) return false;
    if ((double)vt > __dmul_rn((double)maxVar, (double)maxVar) && maxVar > 0.0f &&
        (double)psum > __dmul_rn((double)__fmul_rn(__fmul_rn(groundDiff, e), (float)S), k.gp_thresh)) {
        // std::min(psum / factor, 1.0): a quotient of at least one needs no division (psum >= factor > 0 <=> quotient >= 1)
        float nc = 1.0f;
        if (!(k.occ_factor > 0.0 && (double)psum >= k.occ_factor)) {
            const double ncd = __ddiv_rn((double)psum, k.occ_factor);
            nc = (float)((1.0 < ncd) ? 1.0 : ncd);
        }
        const float num = __fadd_rn(__fmul_rn(groundlevel, nc), __fmul_rn(__fmul_rn(oc, og), 2.0f));
        const float den = __fadd_rn(nc, __fmul_rn(oc, 2.0f));
        g = __fdiv_rn(num, den);
        // std::min((psum / (factor * 2.0f) + oc) / 2.0, 1.0): halving is exact (x * 0.5)
        const double cd = __dmul_rn(__dadd_rn(__ddiv_rn((double)psum, k.occ_factor2), (double)oc), 0.5);
        c = (float)((1.0 < cd) ? 1.0 : cd);
        return true;
    }
    if (localmin < og) {
        g = localmin;
        const float t = __fadd_rn(oc, 0.1f);
        c = (0.5f < t) ? 0.5f : t;  // std::min(oc + 0.1f, 0.5f)
        return true;
    }
    return false;
}

// A CTA walks DT_TILES consecutive tiles along j; the TMA loads of tile k + 1 are issued before tile k is computed
// (two stages of raw tiles, one mbarrier each, phase parity flips every second use).
constexpr int DT_TILES = 4;

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(DT_X* DT_Y, MIN_BLOCKS) k_detect_tma(View v, const SlotParams* __restrict__ batch, const __grid_constant__ CUtensorMap tmap) {
    __shared__ DetectTile s;
    __shared__ __align__(8) uint64_t s_bar[2];
    const SlotParams& sp = batch[blockIdx.z];
    const Const& k = v.k;
    const int N = k.N, N2 = k.N2;
    const int i0 = blockIdx.x * DT_X;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * DT_X + tx;
    const int tile0 = blockIdx.y * DT_TILES;
    const int n_tiles = min(DT_TILES, (N + DT_Y - 1) / DT_Y - tile0);
    const int plane0 = sp.slot * v.n_layers;
    auto issue = [&](int kk) {   // one thread: loads of tile kk into stage kk & 1
        constexpr uint32_t kBytes = 3u * DT_R * DT_WT * sizeof(float);
        DetectRaw& r = s.raw[kk & 1];
        uint64_t* bar = &s_bar[kk & 1];
        const int j0 = (tile0 + kk) * DT_Y;
        mbar_expect_tx(bar, kBytes);
        tma_load_3d(&r.P[0][0], &tmap, bar, i0 - DT_HX, j0 - DT_H, plane0 + L_COUNT);
        tma_load_3d(&r.V[0][0], &tmap, bar, i0 - DT_HX, j0 - DT_H, plane0 + L_VARIANCE);
        tma_load_3d(&r.M[0][0], &tmap, bar, i0 - DT_HX, j0 - DT_H, plane0 + L_MINH);
    };
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
    }
    __syncthreads();
    if (tid == 0) issue(0);
    float* const L0 = v.layer(sp.slot, 0);
    const int i = i0 + tx;
    const int li = tx + DT_HX, lj = ty + DT_H;
    for (int kk = 0; kk < n_tiles; ++kk) {
        // every thread is past the previous tile (barrier at the end of the loop body): its stage may be refilled
        if (tid == 0 && kk + 1 < n_tiles) issue(kk + 1);
        const DetectRaw& raw = s.raw[kk & 1];
        // own cell (overlaps the tile transfer)
        const int j = (tile0 + kk) * DT_Y + ty;
        const bool live = i < N && j < N;
        const int cell = i + j * N;
        float g = 0.0f, c = 0.0f;
        float4 tb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (live) {
            g = L0[L_GROUND * N2 + cell];
            c = L0[L_GROUNDPATCH * N2 + cell];
            tb = __ldg(v.detect_tab + cell);
        }
        const int flags = __float_as_int(tb.w);
        mbar_wait(&s_bar[kk & 1], (uint32_t)((kk >> 1) & 1));

        // stage A: column partials of the point counts (exact) -- positions p = tid, tid + 256 of the 12 x 40 tile
        // (only columns 0 .. 35 are ever the first row of a window: col + 4 <= 39 stays inside the tile, no bounds tests)
        for (int p = tid; p < DT_R * DT_WC; p += DT_X * DT_Y) {
            const int row = p / DT_WC, col = p % DT_WC;   // row = j index of the tile, col = i index
            float a[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) a[q] = raw.P[row][col + q];
            const float c3 = __fadd_rn(__fadd_rn(a[0], a[1]), a[2]);
            s.C3[row][col] = c3;
            s.C5[row][col] = __fadd_rn(__fadd_rn(c3, a[3]), a[4]);
        }
        __syncthreads();
        const bool inner = live && (flags & DTF_INNER);
        const bool s5 = (flags & DTF_S5) != 0;
        float psum = 0.0f;
        if (inner) {
            if (s5) {
                const int r0 = li - 2, c0 = lj - 2;
                psum = __fadd_rn(__fadd_rn(__fadd_rn(s.C5[c0][r0], s.C5[c0 + 1][r0]), __fadd_rn(s.C5[c0 + 2][r0], s.C5[c0 + 3][r0])), s.C5[c0 + 4][r0]);
            } else {
                const int r0 = li - 1, c0 = lj - 1;
                psum = __fadd_rn(__fadd_rn(s.C3[c0][r0], s.C3[c0 + 1][r0]), s.C3[c0 + 2][r0]);
            }
        }
        // early skipping of (almost) empty areas, :364 (both sides integer valued: the float compare is the double one)
        const bool cand = inner && !(psum < tb.x);
        const bool any = __syncthreads_or(cand);
        bool changed = false;
        if (any) {
            // stage B: products, pairs, triples and column minima
            // (products are needed up to column 37 -- the last row of the right-most window --, pairs up to 36, the rest up
            // to 35; reads of columns 40 / 41 for those two extra columns stay inside this struct and feed unused entries)
            for (int p = tid; p < DT_R * DT_WB; p += DT_X * DT_Y) {
                const int row = p / DT_WB, col = p % DT_WB;
                float qv[3], qm[3], m[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) m[q] = raw.M[row][col + q];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float pp = raw.P[row][col + q];
                    const float vv = raw.V[row][col + q];
                    qv[q] = __fmul_rn(pp, vv);
                    qm[q] = __fmul_rn(pp, m[q]);
                }
                s.QV[row][col] = qv[0];
                s.QM[row][col] = qm[0];
                const float pv12 = __fadd_rn(qv[1], qv[2]), pm12 = __fadd_rn(qm[1], qm[2]);
                s.PV2[row][col] = __fadd_rn(qv[0], qv[1]);
                s.PM2[row][col] = __fadd_rn(qm[0], qm[1]);
                s.TV[row][col] = __fadd_rn(qv[0], pv12);
                s.TM[row][col] = __fadd_rn(qm[0], pm12);
                float n3 = m[0];
                n3 = (m[1] < n3) ? m[1] : n3;
                n3 = (m[2] < n3) ? m[2] : n3;
                s.N3[row][col] = n3;
                float n5 = (m[3] < n3) ? m[3] : n3;
                n5 = (m[4] < n5) ? m[4] : n5;
                s.N5[row][col] = n5;
            }
            __syncthreads();
            if (cand) {
                const float variance = raw.V[lj][li], centerP = raw.P[lj][li];
                if (s5) {
                    const int r0 = li - 2, c0 = lj - 2;
                    float localmin = s.N5[c0][r0];
#pragma unroll
                    for (int q = 1; q < 5; ++q) {
                        const float t = s.N5[c0 + q][r0];
                        localmin = (t < localmin) ? t : localmin;
                    }
                    changed = detect_decide<5>(k, psum, localmin, tree25(s.QV, s.PV2, s.TV, r0, c0), tree25(s.QM, s.PM2, s.TM, r0, c0), centerP, variance,
                                               tb.y, tb.z, g, c);
                } else {
                    const int r0 = li - 1, c0 = lj - 1;
                    float localmin = s.N3[c0][r0];
#pragma unroll
                    for (int q = 1; q < 3; ++q) {
                        const float t = s.N3[c0 + q][r0];
                        localmin = (t < localmin) ? t : localmin;
                    }
                    changed = detect_decide<3>(k, psum, localmin, tree9s(s.QV, s.PV2, r0, c0), tree9s(s.QM, s.PM2, r0, c0), centerP, variance, tb.y, tb.z, g, c);
                }
            }
        }
        if (live) {
            if (changed) {
                L0[L_GROUND * N2 + cell] = g;
                L0[L_GROUNDPATCH * N2 + cell] = c;
            }
            if (v.skew.sk) {
                skew_store_cell(v, sp, cell, i, j, g, c, (flags & DTF_FAR) != 0);
            } else if (v.spiral_recs) {
                const float d1 = decay_confidence(k, c);
                float* D1 = v.roll_scratch + (size_t)sp.slot * 2 * k.N2;
                D1[cell] = d1;
                if (i == j) D1[k.N2 + cell] = decay_confidence(k, d1);
            }
        }
        __syncthreads();   // the derived arrays and this stage are free again
    }
}

int launch_build_detect_table(const View& v, float4* tab, cudaStream_t st) {
    k_build_detect_table<<<(v.k.N2 + 255) / 256, 256, 0, st>>>(v, tab);
    return 1;
}

// ------------------------------------------------------------------------------------------
// phase 3: spiral interpolation as a level-scheduled wavefront (one CTA per scan).
// The schedule (host-built from the exact RAW/WAR/WAW dependency DAG of the sequential
// sweep, GroundSegmentation.cpp:413-440) guarantees that visits of one level touch disjoint
// data, so executing levels in order reproduces the sequential result bit for bit.
// ------------------------------------------------------------------------------------------
constexpr int SPIRAL_THREADS = 512;

__global__ void __launch_bounds__(SPIRAL_THREADS) k_spiral(View v, const SlotParams* __restrict__ batch) {
    const SlotParams& sp = batch[blockIdx.x];
    const Const& k = v.k;
    const int N = k.N;
    float* G = v.layer(sp.slot, L_GROUND);
    float* C = v.layer(sp.slot, L_GROUNDPATCH);
    const int cidx = N / 2 - 1;
    if (threadIdx.x == 0) {
        C[cidx + cidx * N] = 1.0f;          // :405
        G[cidx + cidx * N] = sp.base_z_f;   // :411
    }
    __syncthreads();
    const float fc = (float)cidx;
    for (int lvl = 0; lvl < v.levels; ++lvl) {
        const int b = v.level_start[lvl], e = v.level_start[lvl + 1];
        for (int t = b + threadIdx.x; t < e; t += SPIRAL_THREADS) {
            const uint32_t vis = v.visits[t];
            const int x = (int)(vis & 0xffffu), y = (int)(vis >> 16);
            float cc[9], pr[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int g = (x - 1 + q % 3) + (y - 1 + q / 3) * N;
                cc[q] = C[g];
                pr[q] = __fmul_rn(cc[q], G[g]);
            }
            const float h = G[x + y * N];
            const float occ = cc[4];
            const float s = __fadd_rn(tree9(cc), FLT_MIN);  // :457
            const float avg = __fdiv_rn(tree9(pr), s);      // :458
            G[x + y * N] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, occ), avg), __fmul_rn(occ, h));  // :460
            const float fx = __fsub_rn((float)x, fc), fy = __fsub_rn((float)y, fc);
            const double d2 = __dmul_rn(__dadd_rn(__dmul_rn((double)fx, (double)fx), __dmul_rn((double)fy, (double)fy)), k.res_sq);
            if (d2 > 12.0) {  // :463
                const double o = (double)occ;
                const double dec = __dsub_rn(o, __ddiv_rn(o, k.dec_factor));
                C[x + y * N] = (float)((dec < 0.001) ? 0.001 : dec);  // std::max(dec, 0.001)
            }
        }
        __syncthreads();
    }
}

// Pipelined wavefront.  Same schedule, same arithmetic, but the per-level critical path no
// longer contains a global-memory round trip:
//   * schedule records are fetched DIST+1 levels ahead, the 3x3 neighbourhoods of G and C (and
//     the pre-decayed confidence) DIST levels ahead, so their L2 latency overlaps earlier levels;
//   * values written fewer than DIST levels before a visit cannot come from that prefetch; the
//     record names them (neighbour index + producer slot) and they travel through a small
//     shared-memory ring written by every visit;
//   * the fp64 confidence decay is read from the table k_detect prepared.
// Per level: barrier -> <= 3 shared loads -> 9 mul, tree sum, div, 2 mul + add -> stores.
template <int THREADS, int DIST>
__global__ void __launch_bounds__(THREADS) k_spiral_pipe(View v, const SlotParams* __restrict__ batch) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    int* s_ls = reinterpret_cast<int*>(s_raw);
    const int L = v.levels;
    float2* xch = reinterpret_cast<float2*>(s_raw + (size_t)((L + 4) & ~3) * sizeof(int));  // [DIST + 1][THREADS]
    const SlotParams& sp = batch[blockIdx.x];
    const Const& k = v.k;
    const int N = k.N;
    const int tid = threadIdx.x;
    float* G = v.layer(sp.slot, L_GROUND);
    float* C = v.layer(sp.slot, L_GROUNDPATCH);
    const float* D1 = v.roll_scratch + (size_t)sp.slot * 2 * k.N2;
    const float* D2 = D1 + k.N2;
    const uint4* __restrict__ recs = v.spiral_recs;
    for (int t = tid; t <= L; t += THREADS) s_ls[t] = v.level_start[t];
    const int cidx = N / 2 - 1;
    if (tid == 0) {
        C[cidx + cidx * N] = 1.0f;          // :405
        G[cidx + cidx * N] = sp.base_z_f;   // :411
    }
    __syncthreads();

    // pipeline state: rec[s] / act[s] describe this thread's visit at level lvl + s
    uint4 rec[DIST + 1];
    bool act[DIST + 1];
    float cc[DIST][9], gg[DIST][9], dd[DIST];
#pragma unroll
    for (int s = 0; s <= DIST; ++s) {
        act[s] = (s < L) && (tid < s_ls[s + 1] - s_ls[s]);
        rec[s] = act[s] ? recs[s_ls[s] + tid] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int s = 0; s < DIST; ++s) {
        dd[s] = 0.0f;
#pragma unroll
        for (int q = 0; q < 9; ++q) cc[s][q] = gg[s][q] = 0.0f;
        if (act[s]) {
            const int x = (int)(rec[s].x & 0xffffu), y = (int)(rec[s].x >> 16);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int g = (x - 1 + q % 3) + (y - 1 + q / 3) * N;
                cc[s][q] = C[g];
                gg[s][q] = G[g];
            }
            dd[s] = (rec[s].w & 2u) ? D2[x + y * N] : D1[x + y * N];
        }
    }

    int buf = 0;  // lvl % (DIST + 1)
    for (int lvl = 0; lvl < L; ++lvl) {
        // (1) prefetch: neighbourhood of the visit at lvl + DIST, record of lvl + DIST + 1
        float ncc[9], ngg[9], ndd = 0.0f;
#pragma unroll
        for (int q = 0; q < 9; ++q) ncc[q] = ngg[q] = 0.0f;
        if (act[DIST]) {
            const int x = (int)(rec[DIST].x & 0xffffu), y = (int)(rec[DIST].x >> 16);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int g = (x - 1 + q % 3) + (y - 1 + q / 3) * N;
                ncc[q] = C[g];
                ngg[q] = G[g];
            }
            ndd = (rec[DIST].w & 2u) ? D2[x + y * N] : D1[x + y * N];
        }
        uint4 nrec = make_uint4(0, 0, 0, 0);
        bool nact = false;
        {
            const int l2 = lvl + DIST + 1;
            if (l2 < L) {
                nact = tid < s_ls[l2 + 1] - s_ls[l2];
                if (nact) nrec = recs[s_ls[l2] + tid];
            }
        }
        // (2) this level's visit
        if (act[0]) {
            const uint32_t ents[4] = {rec[0].y & 0xffffu, rec[0].y >> 16, rec[0].z & 0xffffu, rec[0].z >> 16};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t e = ents[r];
                const int back = (int)(e >> 14);  // written `back` levels ago (0: unused entry)
                if (back) {
                    int b = buf - back;
                    if (b < 0) b += DIST + 1;
                    const float2 val = xch[b * THREADS + (int)(e & 1023u)];
                    const int q = (int)((e >> 10) & 15u);
#pragma unroll
                    for (int qq = 0; qq < 9; ++qq)
                        if (q == qq) {
                            gg[0][qq] = val.x;
                            cc[0][qq] = val.y;
                        }
                }
            }
            const int x = (int)(rec[0].x & 0xffffu), y = (int)(rec[0].x >> 16);
            float pr[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) pr[q] = __fmul_rn(cc[0][q], gg[0][q]);
            const float occ = cc[0][4], h = gg[0][4];
            const float ssum = __fadd_rn(tree9(cc[0]), FLT_MIN);  // :457
            const float avg = __fdiv_rn(tree9(pr), ssum);         // :458
            const float newg = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, occ), avg), __fmul_rn(occ, h));  // :460
            const bool far = (rec[0].w & 1u) != 0u;               // :463 (geometry only, evaluated on the host)
            const float newc = far ? dd[0] : occ;                 // :464 (table from k_detect)
            xch[buf * THREADS + tid] = make_float2(newg, newc);
            G[x + y * N] = newg;
            if (far) C[x + y * N] = newc;
        }
        __syncthreads();
        // (3) advance the pipeline by one level
#pragma unroll
        for (int s = 0; s + 1 < DIST; ++s) {
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                cc[s][q] = cc[s + 1][q];
                gg[s][q] = gg[s + 1][q];
            }
            dd[s] = dd[s + 1];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            cc[DIST - 1][q] = ncc[q];
            gg[DIST - 1][q] = ngg[q];
        }
        dd[DIST - 1] = ndd;
#pragma unroll
        for (int s = 0; s < DIST; ++s) {
            rec[s] = rec[s + 1];
            act[s] = act[s + 1];
        }
        rec[DIST] = nrec;
        act[DIST] = nact;
        buf = (buf + 1 == DIST + 1) ? 0 : buf + 1;
    }
}

// ------------------------------------------------------------------------------------------
// Skewed-layout spiral.  k_detect copies G, C (and the confidence each visit will store) into
// (side, level, ring) order (skew_store_cell), k_spiral_skew runs the wavefront there with
// coalesced accesses and also stores every result to the normal layers.  See
// gg_host.cpp:build_spiral_skew for the layout.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 spiral_visit(const float2* nb, float d) {
    float cc[9], pr[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        cc[q] = nb[q].y;
        pr[q] = __fmul_rn(nb[q].y, nb[q].x);
    }
    const float occ = cc[4], h = nb[4].x;
    const float ssum = __fadd_rn(tree9(cc), FLT_MIN);  // :457
    const float avg = __fdiv_rn(tree9(pr), ssum);      // :458
    const float newg = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, occ), avg), __fmul_rn(occ, h));  // :460
    return make_float2(newg, d >= 0.0f ? d : occ);     // :463-464 via the table of k_skew
}

// The level time of this kernel is set by the longest per-warp INSTRUCTION sequence between two
// barriers (one SM, a handful of active warps: a warp issues its dependent instructions a few
// cycles apart), not by memory.  So both roles keep their per-level code short: lane threads are
// compiled per side (compile-time neighbour index of the lane's previous cell, running pointers),
// and the irregular visits are spread over 64 threads (one thread per (visit, neighbour) gathers,
// one thread per visit computes) instead of unrolled in one thread.
constexpr int SKEW_IRR_THREADS = 64;
constexpr int SKEW_RING = 16;      // levels of irregular-visit blocks kept in shared memory
constexpr int SKEW_STAGE_LEAD = 10;  // a block is staged this many levels before its visits run

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- point-to-point synchronisation (ASYNC variants) ---------------------------------------
// Instead of one CTA barrier per level every AGENT (a warp of lane threads; the two irregular warps together) publishes
// the number of levels it has completed in shared memory and, before a level, waits until the agents it depends on have
// reached the progress the host table asks for (gg_host.cpp:build_skew_sync: exact RAW / WAR / WAW sets of the level's
// visits).  Neighbouring rings trail each other by three levels, so most requirements leave a level of slack: a warp
// that is late (a cache miss, a lost issue slot) no longer stalls the whole CTA.  Lane b of a warp watches agent b.
__device__ __forceinline__ void skew_publish(int* s_prog, int agent, int completed, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(s_prog + agent)), "r"(completed) : "memory");
}
__device__ __forceinline__ void skew_wait(const int* s_prog, int need, int lane, int sleep_ns) {
    const uint32_t addr = smem_u32(s_prog + lane);
    for (;;) {
        int p;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(p) : "r"(addr) : "memory");
        if (__all_sync(0xffffffffu, p >= need)) break;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
}

template <int SIDE, bool ASYNC>
__device__ __forceinline__ void skew_lane_thread(const View& v, const SlotParams& sp, float2* s_xch, int* s_prog) {
    const SkewView& w = v.skew;
    constexpr int XM = (ASYNC ? SKEW_XCH_ASYNC : 2) - 1;   // exchange ring: entry of level l is [l & XM]
    const int lane_id = threadIdx.x & 31;
    const int agent = threadIdx.x >> 5;
    const uint16_t* __restrict__ req = ASYNC ? w.req + (size_t)agent * w.levels * 32 + lane_id : nullptr;
    int need_nxt = ASYNC ? (int)__ldg(req) : 0;   // requirement of the next level this warp executes, loaded one level ahead
    constexpr int PQ = SIDE == 0 ? 1 : (SIDE == 1 ? 3 : (SIDE == 2 ? 7 : 5));  // neighbour index of the lane's previous cell
    constexpr int PF_FAR = 8, PF_NEAR = 2;
    const int L = w.levels, KP = w.KP, lanes = w.lanes, M = w.M;
    const int m = threadIdx.x - SIDE * M;  // this thread walks rings m - 1, m - 1 + M, m - 1 + 2M, ... of side SIDE
    float2* __restrict__ SK = w.sk + (size_t)sp.slot * w.slots;
    const float* __restrict__ SD = w.sd + (size_t)sp.slot * w.slots;
    float* __restrict__ Gn = v.layer(sp.slot, L_GROUND);
    float* __restrict__ Cn = v.layer(sp.slot, L_GROUNDPATCH);
    const int cstep = SIDE == 0 ? v.k.N : (SIDE == 1 ? 1 : (SIDE == 2 ? -v.k.N : -1));  // the lane walks +y, +x, -y, -x
    int off[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) off[q] = w.pattern[SIDE * 9 + q];
    // The only slot a regular visit touches for the first time (i.e. that still sits in HBM / L2) is
    // the newest row of its outer ring, plus its SD entry; everything else was read by this SM a
    // few levels ago.  Those two are prefetched PF_FAR levels ahead into L2 and PF_NEAR levels ahead
    // into L1, so the one-level-ahead loads below are cache hits.
    int off_new = off[0];
#pragma unroll
    for (int q = 1; q < 9; ++q) off_new = max(off_new, off[q]);

    // current phase: level range [lb, le) of the regular run, slot of level l = base0 + l * KP,
    // cell of level l = cell0 + l * cstep, exchange-buffer index xid
    int ph = -1, lb = 0, le = 0, base0 = 0, cell0 = 0, xid = 0;
    int w_first = 0, w_last = 0;  // levels in which any lane of this warp has work (or prefetches)
    auto next_phase = [&]() {
        ++ph;
        const size_t e = ((size_t)ph * 4 + SIDE) * M + m;
        const int col = ph * M + m;
        lb = w.ph_begin[e];
        le = w.ph_end[e];
        base0 = (SIDE * w.rows + w.row0) * KP + col;
        cell0 = w.ph_cell0[e] - lb * cstep;
        xid = SIDE * KP + col;
    };
    auto warp_window = [&]() {
        w_first = lb < le ? lb - PF_FAR : 0x7fffffff;
        w_last = lb < le ? le : -1;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            w_first = min(w_first, __shfl_xor_sync(0xffffffffu, w_first, d));
            w_last = max(w_last, __shfl_xor_sync(0xffffffffu, w_last, d));
        }
        w_first = max(w_first, 0) & ~1;  // iterations cover two levels
    };
    next_phase();
    warp_window();

    float2 A[9], B[9];
    float dA = -1.0f, dB = -1.0f;
#pragma unroll
    for (int q = 0; q < 9; ++q) A[q] = B[q] = make_float2(0.f, 0.f);
    if (lb == 0 && le > 0) {
#pragma unroll
        for (int q = 0; q < 9; ++q) A[q] = SK[base0 + off[q]];
        dA = SD[base0];
    }
    // two levels per iteration, alternating register sets (a copy would wait for the loads)
#define GG_LANE_LEVEL(l_, CUR, CURD, NXT, NXTD)                                                      \
    {                                                                                               \
        const int l__ = (l_);                                                                       \
        if (ASYNC) {                                                                                \
            const int need__ = need_nxt;                                                            \
            if (l__ + 1 < L) need_nxt = (int)__ldg(req + (size_t)(l__ + 1) * 32);                   \
            skew_wait(s_prog, need__, lane_id, w.sync_sleep);                                                     \
        }                                                                                           \
        if (l__ + PF_FAR >= lb && l__ + PF_FAR < le) {                                              \
            prefetch_l2(SK + (base0 + (l__ + PF_FAR) * KP + off_new));                              \
            prefetch_l2(SD + (base0 + (l__ + PF_FAR) * KP));                                        \
        }                                                                                           \
        if (l__ + PF_NEAR >= lb && l__ + PF_NEAR < le) {                                            \
            prefetch_l1(SK + (base0 + (l__ + PF_NEAR) * KP + off_new));                             \
            prefetch_l1(SD + (base0 + (l__ + PF_NEAR) * KP));                                       \
        }                                                                                           \
        if (l__ + 1 >= lb && l__ + 1 < le) {                                                        \
            const float2* p__ = SK + (base0 + (l__ + 1) * KP);                                      \
            _Pragma("unroll") for (int q = 0; q < 9; ++q) NXT[q] = p__[off[q]];                    \
            NXTD = SD[base0 + (l__ + 1) * KP];                                                      \
        }                                                                                           \
        if (l__ >= lb && l__ < le) {                                                                \
            CUR[PQ] = s_xch[((l__ - 1) & XM) * lanes + xid];                                        \
            const float2 r__ = spiral_visit(CUR, CURD);                                             \
            s_xch[(l__ & XM) * lanes + xid] = r__;                                                  \
            SK[base0 + l__ * KP] = r__;                                                             \
            Gn[cell0 + l__ * cstep] = r__.x;                                                        \
            if (CURD >= 0.0f) Cn[cell0 + l__ * cstep] = r__.y;                                      \
        }                                                                                           \
        if (ASYNC)                                                                                  \
            skew_publish(s_prog, agent, l__ + 1, lane_id);                                          \
        else                                                                                        \
            __syncthreads();                                                                        \
    }
    for (int l = 0; l < L; l += 2) {
        // done with this ring: move on to ring + M (it starts well after this one ended)
        if (w.phases > 1) {
            bool moved = false;
            while (ph + 1 < w.phases && l >= le) {
                next_phase();
                moved = true;
            }
            if (__any_sync(0xffffffffu, moved)) warp_window();
        }
        // a warp (32 consecutive rings of one side) has work only in a window of levels; outside of it
        // the per-level cost must be the barrier alone (the level time is set by instruction issue)
        if (l < w_first || l >= w_last) {
            if (ASYNC) {   // nothing to do and nothing to wait for: the warp has "completed" both levels
                skew_publish(s_prog, agent, min(l + 2, L), lane_id);
                if (l + 2 < L) need_nxt = (int)__ldg(req + (size_t)(l + 2) * 32);
            } else {
                __syncthreads();
                if (l + 1 < L) __syncthreads();
            }
            continue;
        }
        GG_LANE_LEVEL(l, A, dA, B, dB)
        if (l + 1 < L) GG_LANE_LEVEL(l + 1, B, dB, A, dA)
    }
#undef GG_LANE_LEVEL
}

__device__ __forceinline__ void skew_named_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(SKEW_IRR_THREADS) : "memory"); }

template <bool ASYNC>
__device__ __forceinline__ void skew_irregular_thread(const View& v, const SlotParams& sp, float2* s_xch, uint4* s_ring, float2* s_nb, float* s_dd, int* s_prog) {
    const SkewView& w = v.skew;
    const int ti = threadIdx.x - 4 * w.M;  // 0 .. 63
    constexpr int XM = (ASYNC ? SKEW_XCH_ASYNC : 2) - 1;
    const int lane_id = threadIdx.x & 31;
    const int agent = 4 * w.M / 32;        // both warps form one agent
    const uint16_t* __restrict__ req = ASYNC ? w.req + (size_t)agent * w.levels * 32 + lane_id : nullptr;
    int need_nxt = ASYNC ? (int)__ldg(req) : 0;
    const int L = w.levels, lanes = w.lanes;
    const int chunks = w.irr_chunks, irr_max = w.irr_max;
    float2* __restrict__ SK = w.sk + (size_t)sp.slot * w.slots;
    const float* __restrict__ SD = w.sd + (size_t)sp.slot * w.slots;
    const uint4* __restrict__ blocks = w.irr_blocks;
    float* __restrict__ Gn = v.layer(sp.slot, L_GROUND);
    float* __restrict__ Cn = v.layer(sp.slot, L_GROUNDPATCH);
    if (ti == 0) {  // spiral_ground_interpolation :405,411 on the normal layers (the skewed copy has it from k_detect)
        const int cidx = v.k.N / 2 - 1;
        Cn[cidx + cidx * v.k.N] = 1.0f;
        Gn[cidx + cidx * v.k.N] = sp.base_z_f;
    }
    const uint32_t* ring_w = reinterpret_cast<const uint32_t*>(s_ring);
    const int ring_words = chunks * 4;
    const bool stager = ti < chunks;              // copies one 16-byte chunk of the level blocks per level
    const bool gatherer = ti < irr_max * 9;       // (visit ti / 9, neighbour ti % 9)
    const bool visitor = ti < irr_max;            // computes visit ti

    // prologue: blocks of levels 0 .. LEAD-1 -> ring; gather of level 0; block of level LEAD in flight
    if (stager) {
        for (int b = 0; b < SKEW_STAGE_LEAD; ++b) s_ring[b * chunks + ti] = blocks[(size_t)min(b, L + 3) * chunks + ti];  // the table has 4 padding levels
    }
    uint4 stage = stager ? blocks[(size_t)min(SKEW_STAGE_LEAD, L + 3) * chunks + ti] : make_uint4(0, 0, 0, 0);
    skew_named_barrier();
    float2 val = make_float2(0.f, 0.f);
    float dval = -1.0f;
    uint32_t info = 0xffffffffu;
    bool have = false;
    if (gatherer) {
        const uint32_t slot = ring_w[0 * ring_words + ti * 2], rl = ring_w[0 * ring_words + ti * 2 + 1];
        have = slot != 0xffffffffu;
        if (have) {
            val = SK[slot];
            info = rl;
            if (ti % 9 == 4) dval = SD[slot];
        }
    }
    for (int l = 0; l < L; ++l) {
        if (ASYNC) {
            const int need = need_nxt;
            if (l + 1 < L) need_nxt = (int)__ldg(req + (size_t)(l + 1) * 32);
            skew_wait(s_prog, need, lane_id, w.sync_sleep);
        }
        // (1) block of level l + LEAD (loaded during the previous level) -> ring; start loading the next one
        if (stager) {
            s_ring[((l + SKEW_STAGE_LEAD) % SKEW_RING) * chunks + ti] = stage;
            stage = blocks[(size_t)min(l + SKEW_STAGE_LEAD + 1, L + 3) * chunks + ti];
        }
        // (2) hand the gathered neighbour of THIS level to the visitor, gather the one of level l + 1,
        //     pull the one of level l + LEAD - 2 towards L2 / L1
        if (gatherer) {
            if (have) {
                if (info != 0xffffffffu) val = s_xch[((l - 1) & XM) * lanes + (int)info];  // written at level l - 1
                s_nb[ti] = val;
                if (ti % 9 == 4) s_dd[ti / 9] = dval;
            }
            const uint32_t far_slot = ring_w[((l + SKEW_STAGE_LEAD - 2) % SKEW_RING) * ring_words + ti * 2];
            if (far_slot != 0xffffffffu) {
                prefetch_l2(SK + far_slot);
                if (ti % 9 == 4) prefetch_l2(SD + far_slot);
            }
            const uint32_t near_slot = ring_w[((l + 3) % SKEW_RING) * ring_words + ti * 2];
            if (near_slot != 0xffffffffu) {
                prefetch_l1(SK + near_slot);
                if (ti % 9 == 4) prefetch_l1(SD + near_slot);
            }
            const uint32_t* e = ring_w + ((l + 1) % SKEW_RING) * ring_words + ti * 2;
            const uint32_t slot = e[0];
            have = (l + 1 < L) && slot != 0xffffffffu;
            if (have) {
                val = SK[slot];
                info = e[1];
                if (ti % 9 == 4) dval = SD[slot];
            }
        }
        skew_named_barrier();
        // (3) the visits of this level
        if (visitor) {
            const uint32_t* hd = ring_w + (l % SKEW_RING) * ring_words + irr_max * 18 + ti * 4;
            const uint32_t own = hd[0];
            if (own != 0xffffffffu) {
                float2 nb[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) nb[q] = s_nb[ti * 9 + q];
                const float2 r = spiral_visit(nb, s_dd[ti]);
                const int mirror = (int)hd[1];
                s_xch[(l & XM) * lanes + (int)hd[2]] = r;
                SK[own] = r;
                if (mirror >= 0) SK[mirror] = r;
                Gn[hd[3]] = r.x;
                if (s_dd[ti] >= 0.0f) Cn[hd[3]] = r.y;
            }
        }
        if (ASYNC) {
            skew_named_barrier();   // both warps are done with the level (and with s_nb / s_dd)
            if (ti == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(s_prog + agent)), "r"(l + 1) : "memory");
        } else {
            __syncthreads();
        }
    }
}

template <int MAXT, int MIN_CTAS = 1, bool ASYNC = false>
__global__ void __launch_bounds__(MAXT, MIN_CTAS) k_spiral_skew(View v, const SlotParams* __restrict__ batch) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const SkewView& w = v.skew;
    // [ring: SKEW_RING levels x irr_chunks uint4][xch: depth x lanes float2][nb: irr_max*9 float2][dd: irr_max float][progress: 32 int]
    uint4* s_ring = reinterpret_cast<uint4*>(s_raw);
    float2* s_xch = reinterpret_cast<float2*>(s_ring + SKEW_RING * w.irr_chunks);
    float2* s_nb = s_xch + (ASYNC ? SKEW_XCH_ASYNC : 2) * w.lanes;
    float* s_dd = reinterpret_cast<float*>(s_nb + w.irr_max * 9);
    int* s_prog = reinterpret_cast<int*>(s_dd + w.irr_max);
    const SlotParams& sp = batch[blockIdx.x];
    const int tid = threadIdx.x;
    if (ASYNC) {
        // agents that do not exist count as finished (their table entries are 0 anyway)
        if (tid < 32) s_prog[tid] = tid <= 4 * w.M / 32 ? 0 : 0x7fffffff;
        __syncthreads();
    }
    if (tid < 4 * w.M) {
        const int side = tid / w.M;  // warp-uniform: M is a multiple of 32
        if (side == 0)
            skew_lane_thread<0, ASYNC>(v, sp, s_xch, s_prog);
        else if (side == 1)
            skew_lane_thread<1, ASYNC>(v, sp, s_xch, s_prog);
        else if (side == 2)
            skew_lane_thread<2, ASYNC>(v, sp, s_xch, s_prog);
        else
            skew_lane_thread<3, ASYNC>(v, sp, s_xch, s_prog);
    } else {
        skew_irregular_thread<ASYNC>(v, sp, s_xch, s_ring, s_nb, s_dd, s_prog);
    }
}

// ------------------------------------------------------------------------------------------
// phase 4: labelling (:146-196)
// ------------------------------------------------------------------------------------------
// LABEL_ILP points per thread (strided by the block, so every access stays coalesced): the streaming loads of all
// of them are issued first, then the two gathers each, then the arithmetic -- the kernel is bound by memory latency.
constexpr int LABEL_ILP = 4;

__global__ void __launch_bounds__(256) k_label(View v, const SlotParams* __restrict__ batch) {
    const SlotParams& sp = batch[blockIdx.y];
    const Const& k = v.k;
    const size_t base = (size_t)sp.slot * v.pcap;
    const int i0 = blockIdx.x * (256 * LABEL_ILP) + threadIdx.x;
    const int n = sp.n_points;
    if (i0 >= n) return;
    const float* L0 = v.layer(sp.slot, 0);
    const float* G = L0 + L_GROUND * k.N2;
    const float* V = L0 + L_VARIANCE * k.N2;
    float* OBS = v.layer(sp.slot, L_OBSTACLES);

    uint32_t code[LABEL_ILP];
    float dist[LABEL_ILP], z[LABEL_ILP], gh[LABEL_ILP], var[LABEL_ILP];
#pragma unroll
    for (int u = 0; u < LABEL_ILP; ++u) {
        const int i = i0 + u * 256;
        const bool in = i < n;
        code[u] = in ? v.code[base + i] : (PC_ABSENT << 24);
        dist[u] = in ? v.dist[base + i] : 0.0f;
        z[u] = in ? __uint_as_float(v.zw[base + i].x) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < LABEL_ILP; ++u) {
        const uint32_t cls = code[u] >> 24;
        const int cell = (int)(code[u] & 0xffffffu);
        const bool use = cls == PC_KEPT || cls == PC_IGNORED;
        gh[u] = use ? G[cell] : 0.0f;
        var[u] = use ? V[cell] : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < LABEL_ILP; ++u) {
        const int i = i0 + u * 256;
        if (i >= n) break;
        const uint32_t cls = code[u] >> 24;
        const int cell = (int)(code[u] & 0xffffffu);
        uint8_t label = GG_LABEL_ABSENT;
        if (cls == PC_OUTLIER) {
            label = GG_LABEL_GROUND;
        } else if (cls == PC_KEPT || cls == PC_IGNORED) {
            const double groundheight = (double)gh[u];
            // std::max(std::min((f * dist) / variance * thres, thres), obs_thres) with C++ min/max semantics
            const double a = __dmul_rn(__ddiv_rn(__dmul_rn(k.lab_fac, (double)dist[u]), (double)var[u]), k.lab_thres);
            double t = (k.lab_thres < a) ? k.lab_thres : a;
            t = (t < k.lab_obs) ? k.lab_obs : t;
            if (__dadd_rn(t, groundheight) < (double)z[u]) {
                label = GG_LABEL_NONGROUND;
                atomicAdd(OBS + cell, 1.0f);  // small exact integers: order free
            } else {
                label = GG_LABEL_GROUND;
            }
        }
        v.labels[base + i] = label;
    }
}

// ------------------------------------------------------------------------------------------
// output cloud order: kept (input order), ignored (input order), outliers (input order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int out_class(uint32_t code) {
    const uint32_t cls = code >> 24;
    return cls == PC_KEPT ? 0 : (cls == PC_IGNORED ? 1 : (cls == PC_OUTLIER ? 2 : -1));
}

__global__ void __launch_bounds__(OUT_TILE) k_out_count(View v, const SlotParams* __restrict__ batch, int nblk) {
    const SlotParams& sp = batch[blockIdx.y];
    const int i = blockIdx.x * OUT_TILE + threadIdx.x;
    const int c = (i < sp.n_points) ? out_class(v.code[(size_t)sp.slot * v.pcap + i]) : -1;
    int* counts = v.out_counts + (size_t)sp.slot * (3 * v.out_blocks + 1);
    const int n0 = __syncthreads_count(c == 0);
    const int n1 = __syncthreads_count(c == 1);
    const int n2 = __syncthreads_count(c == 2);
    if (threadIdx.x == 0) {
        counts[0 * nblk + blockIdx.x] = n0;
        counts[1 * nblk + blockIdx.x] = n1;
        counts[2 * nblk + blockIdx.x] = n2;
    }
}

__global__ void __launch_bounds__(1024) k_out_scan(View v, const SlotParams* __restrict__ batch, int nblk) {
    const SlotParams& sp = batch[blockIdx.x];
    int* counts = v.out_counts + (size_t)sp.slot * (3 * v.out_blocks + 1);
    const int total = block_exclusive_scan_1024(counts, counts, 3 * nblk);
    if (threadIdx.x == 0) counts[3 * v.out_blocks] = total;
}

template <bool CLOUD>
__global__ void __launch_bounds__(OUT_TILE) k_out_write(View v, const SlotParams* __restrict__ batch, int nblk) {
    __shared__ int s_w[3][32];
    const SlotParams& sp = batch[blockIdx.y];
    const size_t base = (size_t)sp.slot * v.pcap;
    const int i = blockIdx.x * OUT_TILE + threadIdx.x;
    const int c = (i < sp.n_points) ? out_class(v.code[base + i]) : -1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int rank_in_warp = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const uint32_t bal = __ballot_sync(0xffffffffu, c == q);
        if (c == q) rank_in_warp = __popc(bal & lt_mask);
        if (lane == 0) s_w[q][warp] = __popc(bal);
    }
    __syncthreads();
    if (warp < 3) {  // warp q scans the 32 warp totals of class q
        const int w = s_w[warp][lane];
        const int wi = warp_inclusive_scan(w);
        s_w[warp][lane] = wi - w;
    }
    __syncthreads();
    if (c < 0) return;
    const int* counts = v.out_counts + (size_t)sp.slot * (3 * v.out_blocks + 1);
    const int pos = counts[c * nblk + blockIdx.x] + s_w[c][warp] + rank_in_warp;
    v.out_index[base + pos] = (uint32_t)i;
    if (CLOUD) {
        const uint4* rec = reinterpret_cast<const uint4*>(sp.src + i);
        uint4 a = rec[0], b = rec[1];
        b.x = __float_as_uint((float)v.labels[base + i]);  // intensity = 49 / 99 (:175,180,188)
        uint4* dst = reinterpret_cast<uint4*>(v.out_cloud + base + pos);
        dst[0] = a;
        dst[1] = b;
    }
}

// ------------------------------------------------------------------------------------------
// "next" rows
// ------------------------------------------------------------------------------------------
// f1: pcl::fromROSMsg (field-offset driven unpack, GroundGridNodelet.cpp:119-120) + the per-point
// tf2::doTransform into the map frame in fp64, stored as float (:166-181).
__global__ void __launch_bounds__(256) k_unpack_transform(UnpackDesc d) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= d.n) return;
    const unsigned char* p = d.raw + (size_t)i * d.point_step;
    auto rd32 = [&](int off) -> uint32_t {
        if (off < 0) return 0u;
        return (uint32_t)p[off] | ((uint32_t)p[off + 1] << 8) | ((uint32_t)p[off + 2] << 16) | ((uint32_t)p[off + 3] << 24);
    };
    float x = __uint_as_float(rd32(d.off[0])), y = __uint_as_float(rd32(d.off[1])), z = __uint_as_float(rd32(d.off[2]));
    const uint32_t inten = rd32(d.off[3]);
    const uint32_t ring = d.off[4] < 0 ? 0u : ((uint32_t)p[d.off[4]] | ((uint32_t)p[d.off[4] + 1] << 8));
    if (d.transform) {
        const double dx = (double)x, dy = (double)y, dz = (double)z;
        // tf2::Transform * Vector3: row.dot(v) (left to right) + origin
        x = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(d.T[0], dx), __dmul_rn(d.T[1], dy)), __dmul_rn(d.T[2], dz)), d.T[3]);
        y = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(d.T[4], dx), __dmul_rn(d.T[5], dy)), __dmul_rn(d.T[6], dz)), d.T[7]);
        z = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(d.T[8], dx), __dmul_rn(d.T[9], dy)), __dmul_rn(d.T[10], dz)), d.T[11]);
    }
    uint4* dst = reinterpret_cast<uint4*>(d.dst + i);
    dst[0] = make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(z), 0u);
    dst[1] = make_uint4(inten, ring, 0u, 0u);
}

// f3: the "terrain" image of GroundGridNodelet::publish_grid_map_layer (:247-270): CV_32FC3, pixel
// (i, j) = (ground height, 3x3 sum of pointsRaw >= 27 ? 1 : 0, pointsRaw).  The reference reads the
// 3x3 block out of bounds on the border cells; here border cells get visited = 0.
__global__ void __launch_bounds__(256) k_terrain_image(View v, int slot, float* __restrict__ dst) {
    const int cell = blockIdx.x * 256 + threadIdx.x;
    if (cell >= v.k.N2) return;
    const int N = v.k.N;
    const int i = cell % N, j = cell / N;
    const float* raw = v.layer(slot, L_RAW);
    float visited = 0.0f;
    if (i >= 1 && j >= 1 && i < N - 1 && j < N - 1) {
        float e[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) e[q] = raw[(i - 1 + q % 3) + (j - 1 + q / 3) * N];
        visited = tree9(e) >= 27.0f ? 1.0f : 0.0f;
    }
    float* px = dst + ((size_t)i * N + j) * 3;  // cv::Mat row = index(0), col = index(1)
    px[0] = v.layer(slot, L_GROUND)[cell];
    px[1] = visited;
    px[2] = raw[cell];
}

// f3, second half: the 8-bit image grid_map::GridMapCvConverter::toImage<unsigned char, 1>(map, layer, CV_8UC1, img)
// hands to cv::applyColorMap (GroundGridNodelet.cpp:238-245): lower / upper = min / max over the finite cells of the
// layer, pixel (i, j) = (unsigned char)(((value - lower) / (upper - lower)) * 255.f), non-finite cells stay 0.
__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= FLT_MAX; }

__global__ void __launch_bounds__(256) k_layer_minmax(const float* __restrict__ layer, int n, float* __restrict__ mm) {
    // mm[0] = min, mm[1] = max, as ordered-int atomics (initialised by the host to +inf / -inf patterns)
    float lo = INFINITY, hi = -INFINITY;
    for (int c = blockIdx.x * 256 + threadIdx.x; c < n; c += gridDim.x * 256) {
        const float x = layer[c];
        if (finite_f(x)) {
            lo = fminf(lo, x);
            hi = fmaxf(hi, x);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        // monotone float -> int mapping so that integer atomicMin / atomicMax order like the floats
        auto key = [](float f) { const int b = __float_as_int(f); return b >= 0 ? b : (b ^ 0x7fffffff); };
        atomicMin(reinterpret_cast<int*>(mm), key(lo));
        atomicMax(reinterpret_cast<int*>(mm) + 1, key(hi));
    }
}

__global__ void __launch_bounds__(256) k_layer_image_u8(const float* __restrict__ layer, int N, const float* __restrict__ mm, unsigned char* __restrict__ dst) {
    const int cell = blockIdx.x * 256 + threadIdx.x;
    if (cell >= N * N) return;
    auto unkey = [](int kx) { return __int_as_float(kx >= 0 ? kx : (kx ^ 0x7fffffff)); };
    const float lower = unkey(reinterpret_cast<const int*>(mm)[0]), upper = unkey(reinterpret_cast<const int*>(mm)[1]);
    const float x = layer[cell];
    unsigned char px = 0;
    if (finite_f(x)) {
        const float t = __fmul_rn(__fdiv_rn(__fsub_rn(x, lower), __fsub_rn(upper, lower)), 255.0f);
        px = (t == t) ? (unsigned char)(int)t : 0;   // (Type_) cast: truncation; a constant layer (0 / 0) has no defined image
    }
    const int i = cell % N, j = cell / N;
    dst[(size_t)i * N + j] = px;   // cv::Mat row = index(0), col = index(1)
}

int launch_layer_image_u8(const View& v, const float* layer, float* mm, unsigned char* dst, cudaStream_t st) {
    k_layer_minmax<<<148, 256, 0, st>>>(layer, v.k.N2, mm);
    k_layer_image_u8<<<(v.k.N2 + 255) / 256, 256, 0, st>>>(layer, v.k.N, mm, dst);
    return 2;
}

// f4: the tallies of scripts/eval_groundpoint_classifier.py:95-118 for one segmented cloud: per
// ground-truth label id (carried in `ring`, scripts/kitti_data_publisher.py:122-132) the number of
// points predicted non-ground (intensity 99) and ground (49).
__global__ void __launch_bounds__(256) k_eval_counts(View v, const SlotParams* __restrict__ batch, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_cnt[EVAL_LABELS * 2];
    for (int t = threadIdx.x; t < EVAL_LABELS * 2; t += 256) s_cnt[t] = 0u;
    __syncthreads();
    const SlotParams& sp = batch[0];
    const size_t base = (size_t)sp.slot * v.pcap;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < sp.n_points; i += gridDim.x * 256) {
        const unsigned label = v.labels[base + i];
        if (label == GG_LABEL_ABSENT) continue;
        unsigned ring;
        if (sp.packed)
            ring = reinterpret_cast<const unsigned short*>(sp.packed + 3 * ((sp.n_points + 7) & ~7))[i];
        else
            ring = reinterpret_cast<const uint4*>(sp.src + i)[1].y & 0xffffu;
        if (ring < EVAL_LABELS) atomicAdd(&s_cnt[ring * 2 + (label == GG_LABEL_NONGROUND ? 1 : 0)], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < EVAL_LABELS * 2; t += 256)
        if (s_cnt[t]) atomicAdd(&counts[t], (unsigned long long)s_cnt[t]);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// GroundGrid::initGroundGrid (src/GroundGrid.cpp:71-75): ground = z, groundpatch = 1e-7, points = 0,
// min = 100, max = -100; the per-scan layers start at 0.
__global__ void __launch_bounds__(256) k_init_map(View v, int slot, float z) {
    const int cell = blockIdx.x * 256 + threadIdx.x;
    if (cell >= v.k.N2) return;
    v.layer(slot, L_GROUND)[cell] = z;
    v.layer(slot, L_GROUNDPATCH)[cell] = (float)0.0000001;
    v.layer(slot, L_OBSTACLES)[cell] = 0.0f;
    v.layer(slot, L_COUNT)[cell] = 0.0f;
    v.layer(slot, L_VARIANCE)[cell] = 0.0f;
    v.layer(slot, L_MINH)[cell] = 100.0f;
    if (v.k.full_layers) {
        v.layer(slot, L_MAXH)[cell] = -100.0f;
        v.layer(slot, L_GCAND)[cell] = 0.0f;
        v.layer(slot, L_PLANEDIST)[cell] = 0.0f;
        v.layer(slot, L_M2)[cell] = 0.0f;
        v.layer(slot, L_MEAN)[cell] = 0.0f;
        v.layer(slot, L_RAW)[cell] = 0.0f;
    }
}

int launch_init_map(const View& v, int slot, float z, cudaStream_t st) {
    k_init_map<<<cdiv(v.k.N2, 256), 256, 0, st>>>(v, slot, z);
    return 1;
}

struct Mark {
    Profiler* p;
    int id;
    cudaStream_t st;
    Mark(Profiler* p_, int id_, cudaStream_t st_) : p(p_), id(id_), st(st_) {
        if (p) p->begin(id, st);
    }
    ~Mark() {
        if (p) p->end(id, st);
    }
};
#define GG_LAUNCH(id, ...)        \
    do {                          \
        Mark mark__(prof, id, st);\
        __VA_ARGS__;              \
    } while (0)

int launch_roll(const View& v, const SlotParams* batch, int count, cudaStream_t st, Profiler* prof) {
    dim3 grid(cdiv(v.k.N2, 256 * ROLL_ILP), count);
    GG_LAUNCH(K_ROLL_GATHER, k_roll_gather<<<grid, 256, 0, st>>>(v, batch));
    GG_LAUNCH(K_ROLL_COMMIT, k_roll_commit<<<grid, 256, 0, st>>>(v, batch));
    return 2;
}

int launch_scan_pipeline(const View& v, const SlotParams* batch, int count, int max_points, int stop_after, cudaStream_t st,
                         Profiler* prof, const CUtensorMap* layer_map, cudaEvent_t after_detect) {
    int launches = 0;
    const int nb = max(1, cdiv(max_points, RASTER_TILE));

    static const int raster_occ = getenv("GG_RASTER_OCC") ? atoi(getenv("GG_RASTER_OCC")) : 5;
    if (raster_occ >= 5)
        GG_LAUNCH(K_RASTERIZE, k_rasterize<5><<<dim3(nb, count), RASTER_THREADS, 0, st>>>(v, batch));
    else
        GG_LAUNCH(K_RASTERIZE, k_rasterize<4><<<dim3(nb, count), RASTER_THREADS, 0, st>>>(v, batch));
    GG_LAUNCH(K_CELL_TILES, k_cell_tiles<<<dim3(v.cell_tiles, count), CT_THREADS, 0, st>>>(v, batch));
    GG_LAUNCH(K_CELL_PLACE, k_cell_place<<<dim3(v.cell_tiles, count), CT_THREADS, 0, st>>>(v, batch));
    GG_LAUNCH(K_SCATTER, k_scatter<<<dim3(max(1, cdiv(max_points, 256 * SCATTER_ILP)), count), 256, 0, st>>>(v, batch));
    launches += 4;

    if (v.k.full_layers)
        GG_LAUNCH(K_CELL_STATS, k_cell_stats<true><<<dim3(cdiv(v.k.N2, CS_THREADS * 4), count), CS_THREADS, 0, st>>>(v, batch));
    else
        GG_LAUNCH(K_CELL_STATS, k_cell_stats<false><<<dim3(cdiv(v.k.N2, CS_THREADS * 4), count), CS_THREADS, 0, st>>>(v, batch));
    ++launches;
    if (stop_after == 1) return launches;

    static const int detect_occ = getenv("GG_DETECT_OCC") ? atoi(getenv("GG_DETECT_OCC")) : 5;
    const dim3 dgrid(cdiv(v.k.N, DT_X), cdiv(v.k.N, DT_Y), count);
    const dim3 tgrid(cdiv(v.k.N, DT_X), cdiv(cdiv(v.k.N, DT_Y), DT_TILES), count);
    if (layer_map) {   // TMA-staged tile (needs a 16-byte row pitch: N % 4 == 0)
        if (detect_occ >= 5)
            GG_LAUNCH(K_DETECT, k_detect_tma<5><<<tgrid, dim3(DT_X, DT_Y), 0, st>>>(v, batch, *layer_map));
        else
            GG_LAUNCH(K_DETECT, k_detect_tma<4><<<tgrid, dim3(DT_X, DT_Y), 0, st>>>(v, batch, *layer_map));
    } else if (detect_occ >= 5)
        GG_LAUNCH(K_DETECT, k_detect_ldg<5><<<dgrid, dim3(DT_X, DT_Y), 0, st>>>(v, batch));
    else
        GG_LAUNCH(K_DETECT, k_detect_ldg<4><<<dgrid, dim3(DT_X, DT_Y), 0, st>>>(v, batch));
    ++launches;
    if (after_detect) cudaEventRecord(after_detect, st);   // the spiral (few warps per SM, latency bound) starts here: see gg_capi.cu
    if (stop_after == 2) return launches;

    if (v.skew.sk) {
        // batches run the small-CTA layout (two or three scans share an SM and leave room for the other streams'
        // kernels); a few scans alone get one thread per lane.  GG_SPIRAL_SHARE_MIN: smallest batch that shares.
        static const int share_min = getenv("GG_SPIRAL_SHARE_MIN") ? atoi(getenv("GG_SPIRAL_SHARE_MIN")) : 9;
        View vs = v;
        if (count >= share_min) {
            vs.skew.M = v.skew.thr_M;
            vs.skew.phases = v.skew.thr_phases;
            vs.skew.ph_begin = v.skew.thr_ph_begin;
            vs.skew.ph_end = v.skew.thr_ph_end;
            vs.skew.ph_cell0 = v.skew.thr_ph_cell0;
            vs.skew.req = v.skew.thr_req;
        }
        const bool async = vs.skew.req != nullptr;   // point-to-point synchronisation table of this thread layout
        const int threads = 4 * vs.skew.M + SKEW_IRR_THREADS;
        const size_t shm = (size_t)SKEW_RING * v.skew.irr_chunks * sizeof(uint4) + (size_t)(async ? SKEW_XCH_ASYNC : 2) * v.skew.lanes * sizeof(float2) +
                           (size_t)v.skew.irr_max * 9 * sizeof(float2) + (size_t)v.skew.irr_max * sizeof(float) + 32 * sizeof(int) + 16;
#define GG_SKEW_LAUNCH(T, C)                                                                                                     \
    {                                                                                                                            \
        if (shm > 48 * 1024) {   /* large maps: opt in to more dynamic shared memory */                                         \
            cudaFuncSetAttribute(k_spiral_skew<T, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);             \
            cudaFuncSetAttribute(k_spiral_skew<T, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);              \
        }                                                                                                                        \
        if (async)                                                                                                               \
            GG_LAUNCH(K_SPIRAL, (k_spiral_skew<T, C, true><<<count, threads, shm, st>>>(vs, batch)));                            \
        else                                                                                                                     \
            GG_LAUNCH(K_SPIRAL, (k_spiral_skew<T, C, false><<<count, threads, shm, st>>>(vs, batch)));                           \
    }
        if (threads <= 320)       // time-shared lane threads (GG_SPIRAL_M): several scans share an SM
            GG_SKEW_LAUNCH(320, 3)
        else if (threads <= 448)
            GG_SKEW_LAUNCH(448, 2)
        else if (threads <= 768)
            GG_SKEW_LAUNCH(768, 1)
        else
            GG_SKEW_LAUNCH(1024, 1)
#undef GG_SKEW_LAUNCH
    } else if (v.spiral_recs) {
        const size_t shm = (size_t)((v.levels + 4) & ~3) * sizeof(int) + (size_t)(v.spiral_dist + 1) * v.spiral_threads * sizeof(float2);
#define GG_SPIRAL_CASE(T, D)                                                                      \
    if (v.spiral_threads == T && v.spiral_dist == D) {                                            \
        if (shm > 48 * 1024) cudaFuncSetAttribute(k_spiral_pipe<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm); \
        GG_LAUNCH(K_SPIRAL, k_spiral_pipe<T, D><<<count, T, shm, st>>>(v, batch));                \
    }
        GG_SPIRAL_CASE(512, 1)
        GG_SPIRAL_CASE(512, 2)
        GG_SPIRAL_CASE(512, 3)
        GG_SPIRAL_CASE(1024, 1)
        GG_SPIRAL_CASE(1024, 2)
        GG_SPIRAL_CASE(1024, 3)
#undef GG_SPIRAL_CASE
    } else {
        GG_LAUNCH(K_SPIRAL, k_spiral<<<count, SPIRAL_THREADS, 0, st>>>(v, batch));
    }
    ++launches;
    if (stop_after == 3) return launches;

    GG_LAUNCH(K_LABEL, k_label<<<dim3(max(1, cdiv(max_points, 256 * LABEL_ILP)), count), 256, 0, st>>>(v, batch));
    ++launches;
    return launches;
}

// ------------------------------------------------------------------------------------------
// Single phases / single cells on their own: the public per-phase methods of the reference class
// (GroundSegmentation.h:56-62) -- detect_ground_patches, detect_ground_patch<S>, spiral_ground_interpolation,
// interpolate_cell -- for callers that drive the phases themselves.  Same arithmetic as the pipeline kernels.
// ------------------------------------------------------------------------------------------
// GroundSegmentation::interpolate_cell (:445-465) for one cell, in place
__global__ void k_interpolate_cell(View v, int slot, int x, int y) {
    const Const& k = v.k;
    const int N = k.N;
    float* G = v.layer(slot, L_GROUND);
    float* C = v.layer(slot, L_GROUNDPATCH);
    float cc[9], pr[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int g = (x - 1 + q % 3) + (y - 1 + q / 3) * N;
        cc[q] = C[g];
        pr[q] = __fmul_rn(cc[q], G[g]);
    }
    const float h = G[x + y * N];
    const float occ = cc[4];
    const float s = __fadd_rn(tree9(cc), FLT_MIN);  // :457
    const float avg = __fdiv_rn(tree9(pr), s);      // :458
    G[x + y * N] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, occ), avg), __fmul_rn(occ, h));  // :460
    const float fc = (float)(N / 2 - 1);
    const float fx = __fsub_rn((float)x, fc), fy = __fsub_rn((float)y, fc);
    const double d2 = __dmul_rn(__dadd_rn(__dmul_rn((double)fx, (double)fx), __dmul_rn((double)fy, (double)fy)), k.res_sq);
    if (d2 > 12.0) {  // :463
        const double o = (double)occ;
        const double dec = __dsub_rn(o, __ddiv_rn(o, k.dec_factor));
        C[x + y * N] = (float)((dec < 0.001) ? 0.001 : dec);  // std::max(dec, 0.001)
    }
}

// GroundSegmentation::detect_ground_patch<S> (:343-395) for one cell, reading the layers directly
template <int S>
__global__ void k_detect_cell(View v, int slot, int i, int j) {
    const Const& k = v.k;
    const int N = k.N, H = S / 2;
    const float* P = v.layer(slot, L_COUNT);
    const float* V = v.layer(slot, L_VARIANCE);
    const float* M = v.layer(slot, L_MINH);
    float* G = v.layer(slot, L_GROUND);
    float* C = v.layer(slot, L_GROUNDPATCH);
    const int cell = i + j * N;
    auto at = [&](const float* L, int q) { return L[(i - H + q % S) + (j - H + q / S) * N]; };
    const double di = __dsub_rn((double)i, (double)N / 2.0), dj = __dsub_rn((double)j, (double)N / 2.0);
    const float sqdist = (float)__dmul_rn(__dadd_rn(__dmul_rn(di, di), __dmul_rn(dj, dj)), k.res_sq);  // :356
    const float e = v.expected[cell];
    const float psum = TreeSum<0, S * S>::run([&](int q) { return at(P, q); });
    double need = floor(__dmul_rn(__dmul_rn(k.gp_thresh, (double)S), (double)e));
    need = (need < 3.0) ? 3.0 : need;
    if ((double)psum < need) return;  // :364
    const double a = __dmul_rn((double)sqdist, k.df_sq);
    const double m = (a < k.mdf_sq) ? k.mdf_sq : a;
    const float vt = (float)((k.mdf10_sq < m) ? k.mdf10_sq : m);  // :369
    float localmin = at(M, 0);
    for (int q = 1; q < S * S; ++q) {
        const float t = at(M, q);
        localmin = (t < localmin) ? t : localmin;
    }
    const float sumPV = TreeSum<0, S * S>::run([&](int q) { return __fmul_rn(at(P, q), at(V, q)); });
    const float sumPM = TreeSum<0, S * S>::run([&](int q) { return __fmul_rn(at(P, q), at(M, q)); });
    float g = G[cell], c = C[cell];
    if (detect_decide<S>(k, psum, localmin, sumPV, sumPM, P[cell], V[cell], vt, e, g, c)) {
        G[cell] = g;
        C[cell] = c;
    }
}

int launch_detect_only(const View& v, const SlotParams* batch, int count, cudaStream_t st, Profiler* prof, const CUtensorMap* layer_map) {
    const dim3 dgrid(cdiv(v.k.N, DT_X), cdiv(v.k.N, DT_Y), count);
    if (layer_map)
        GG_LAUNCH(K_DETECT, k_detect_tma<4><<<dim3(cdiv(v.k.N, DT_X), cdiv(cdiv(v.k.N, DT_Y), DT_TILES), count), dim3(DT_X, DT_Y), 0, st>>>(v, batch, *layer_map));
    else
        GG_LAUNCH(K_DETECT, k_detect_ldg<4><<<dgrid, dim3(DT_X, DT_Y), 0, st>>>(v, batch));
    return 1;
}

// the plain level-scheduled wavefront on the normal layers (no skewed copy needed)
int launch_spiral_only(const View& v, const SlotParams* batch, int count, cudaStream_t st, Profiler* prof) {
    GG_LAUNCH(K_SPIRAL, k_spiral<<<count, SPIRAL_THREADS, 0, st>>>(v, batch));
    return 1;
}

int launch_interpolate_cell(const View& v, int slot, int x, int y, cudaStream_t st) {
    k_interpolate_cell<<<1, 1, 0, st>>>(v, slot, x, y);
    return 1;
}

int launch_detect_cell(const View& v, int slot, int S, int i, int j, cudaStream_t st) {
    if (S == 3)
        k_detect_cell<3><<<1, 1, 0, st>>>(v, slot, i, j);
    else
        k_detect_cell<5><<<1, 1, 0, st>>>(v, slot, i, j);
    return 1;
}

int launch_output(const View& v, const SlotParams* batch, int count, int max_points, bool want_cloud, cudaStream_t st,
                  Profiler* prof) {
    const int nblk = max(1, cdiv(max_points, OUT_TILE));
    GG_LAUNCH(K_OUT_COUNT, k_out_count<<<dim3(nblk, count), OUT_TILE, 0, st>>>(v, batch, nblk));
    GG_LAUNCH(K_OUT_SCAN, k_out_scan<<<count, 1024, 0, st>>>(v, batch, nblk));
    if (want_cloud)
        GG_LAUNCH(K_OUT_WRITE, k_out_write<true><<<dim3(nblk, count), OUT_TILE, 0, st>>>(v, batch, nblk));
    else
        GG_LAUNCH(K_OUT_WRITE, k_out_write<false><<<dim3(nblk, count), OUT_TILE, 0, st>>>(v, batch, nblk));
    return 3;
}

int launch_unpack(const UnpackDesc& d, cudaStream_t st, Profiler* prof) {
    GG_LAUNCH(K_UNPACK, k_unpack_transform<<<max(1, cdiv(d.n, 256)), 256, 0, st>>>(d));
    return 1;
}

int launch_terrain_image(const View& v, int slot, float* dst, cudaStream_t st, Profiler* prof) {
    GG_LAUNCH(K_TERRAIN, k_terrain_image<<<cdiv(v.k.N2, 256), 256, 0, st>>>(v, slot, dst));
    return 1;
}

int launch_eval(const View& v, const SlotParams* batch, unsigned long long* counts, cudaStream_t st, Profiler* prof) {
    GG_LAUNCH(K_EVAL, k_eval_counts<<<148, 256, 0, st>>>(v, batch, counts));
    return 1;
}

}  // namespace gg
