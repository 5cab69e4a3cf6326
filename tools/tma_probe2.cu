// Second probe: where may the tensor map live?  variant 0: __grid_constant__ param + libcu++ wrappers, 1: global memory copy + raw PTX,
// 2: __constant__ symbol + raw PTX, 3: __grid_constant__ + raw PTX with the shared::cta destination form
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;
constexpr int W = 36, R = 12;
__constant__ CUtensorMap c_map;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe_cde(const __grid_constant__ CUtensorMap tmap, float* out, int c0, int c1) {
    __shared__ alignas(128) float tile[R][W];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &tmap, c0, c1, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int k = threadIdx.x; k < R * W; k += blockDim.x) out[k] = (&tile[0][0])[k];
}

template <int MODE>
__global__ void probe_raw(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, float* out, int c0, int c1) {
    __shared__ __align__(128) float tile[R * W];
    __shared__ __align__(8) uint64_t bar;
    const CUtensorMap* mp = MODE == 1 ? gmap : (MODE == 2 ? &c_map : &tmap);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (MODE == 3)
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)), "l"(mp),
                         "r"(smem_u32(&bar)), "r"(c0), "r"(c1)
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)),
                         "l"(mp), "r"(smem_u32(&bar)), "r"(c0), "r"(c1)
                         : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(R * W * 4)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int k = threadIdx.x; k < R * W; k += blockDim.x) out[k] = tile[k];
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int N = 100;
    std::vector<float> h((size_t)N * N);
    for (size_t k = 0; k < h.size(); ++k) h[k] = (float)k;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, W * R * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    const cuuint32_t box[2] = {W, R};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d: encode -> %d (query %d)\n", variant, (int)r, (int)q);
    CUtensorMap* gmap;
    cudaMalloc(&gmap, sizeof(map));
    cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_map, &map, sizeof(map));
    const int c0 = 30, c1 = -2;
    if (variant == 0) probe_cde<<<1, 128>>>(map, out, c0, c1);
    if (variant == 1) probe_raw<1><<<1, 128>>>(map, gmap, out, c0, c1);
    if (variant == 2) probe_raw<2><<<1, 128>>>(map, gmap, out, c0, c1);
    if (variant == 3) probe_raw<3><<<1, 128>>>(map, gmap, out, c0, c1);
    if (variant == 4) probe_raw<0><<<1, 128>>>(map, gmap, out, c0, c1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: kernel -> %s\n", variant, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(W * R);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < R; ++rr)
            for (int cc = 0; cc < W; ++cc) {
                const int i = c0 + cc, j = c1 + rr;
                const float want = (i < 0 || i >= N || j < 0 || j >= N) ? 0.f : (float)((size_t)j * N + i);
                if (o[rr * W + cc] != want) ++bad;
            }
        printf("variant %d: %d wrong values\n", variant, bad);
    }
    return 0;
}
