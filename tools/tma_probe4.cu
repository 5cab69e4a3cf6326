// Fourth probe: matrix of (N, box width, swizzle, start coordinates) for the 2-D tile load.  usage: tma_probe4 N W swz c0 c1
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int c0, int c1, int W, int R) {
    extern __shared__ __align__(1024) float tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(R * W * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)), "l"(&tmap),
                     "r"(smem_u32(&bar)), "r"(c0), "r"(c1)
                     : "memory");
    }
    asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int k = threadIdx.x; k < R * W; k += blockDim.x) out[k] = tile[k];
}
int main(int argc, char** argv) {
    const int N = atoi(argv[1]), W = atoi(argv[2]), swz = atoi(argv[3]), c0 = atoi(argv[4]), c1 = atoi(argv[5]);
    const int R = 12;
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
    const cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)R};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    probe<<<1, 128, W * R * 4>>>(map, out, c0, c1, W, R);
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1;
    if (e == cudaSuccess && !swz) {
        std::vector<float> o(W * R);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        bad = 0;
        for (int rr = 0; rr < R; ++rr)
            for (int cc = 0; cc < W; ++cc) {
                const int i = c0 + cc, j = c1 + rr;
                const float want = (i < 0 || i >= N || j < 0 || j >= N) ? 0.f : (float)((size_t)j * N + i);
                if (o[rr * W + cc] != want) ++bad;
            }
    }
    printf("N %d W %d swz %d c (%d,%d): encode %d kernel '%s' wrong %d\n", N, W, swz, c0, c1, (int)r, cudaGetErrorString(e), bad);
    return 0;
}
