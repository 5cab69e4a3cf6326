// Stand-alone probe of the TMA tile load used by k_detect_tma (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu && tools/tma_probe <variant>
// variant bits: 1 = 2-D map (one plane), 2 = box width 32 instead of 36, 4 = no L2 promotion
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W, int R, bool THREE_D>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int c0, int c1, int c2) {
    __shared__ __align__(128) float tile[R * W];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(R * W * 4)) : "memory");
        if (THREE_D)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(tile)),
                         "l"(&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2)
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)),
                         "l"(&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1)
                         : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0)
        : "memory");
    for (int k = threadIdx.x; k < R * W; k += blockDim.x) out[k] = tile[k];
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int N = 100, planes = 12;
    std::vector<float> h((size_t)N * N * planes);
    for (size_t k = 0; k < h.size(); ++k) h[k] = (float)k;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 36 * 12 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    const bool two_d = variant & 1;
    const int W = (variant & 2) ? 32 : 36;
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
    const cuuint32_t box[3] = {(cuuint32_t)W, 12u, 1u};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, two_d ? 2 : 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, (variant & 4) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d: encode -> %d\n", variant, (int)r);
    const int c0 = 30, c1 = -2, c2 = 3;
    if (W == 36) {
        if (two_d) probe<36, 12, false><<<1, 128>>>(map, out, c0, c1, c2);
        else probe<36, 12, true><<<1, 128>>>(map, out, c0, c1, c2);
    } else {
        if (two_d) probe<32, 12, false><<<1, 128>>>(map, out, c0, c1, c2);
        else probe<32, 12, true><<<1, 128>>>(map, out, c0, c1, c2);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: kernel -> %s\n", variant, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(W * 12);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < 12; ++rr)
            for (int cc = 0; cc < W; ++cc) {
                const int i = c0 + cc, j = c1 + rr;
                const float want = (i < 0 || i >= N || j < 0 || j >= N) ? 0.f : (float)((size_t)(two_d ? 0 : c2) * N * N + (size_t)j * N + i);
                if (o[rr * W + cc] != want) ++bad;
            }
        printf("variant %d: %d wrong values\n", variant, bad);
    }
    return 0;
}
