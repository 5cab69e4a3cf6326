// Third probe: non-tensor bulk copy (UBLKCP) and a dump of the encoded descriptor
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe_bulk(const float* src, float* out) {
    __shared__ __align__(128) float tile[1024];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4096u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tile)), "l"(src), "r"(4096u),
                     "r"(smem_u32(&bar))
                     : "memory");
    }
    asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) out[k] = tile[k];
}
int main() {
    std::vector<float> h(1024);
    for (int k = 0; k < 1024; ++k) h[k] = (float)k;
    float *d, *out;
    cudaMalloc(&d, 4096);
    cudaMalloc(&out, 4096);
    cudaMemcpy(d, h.data(), 4096, cudaMemcpyHostToDevice);
    probe_bulk<<<1, 128>>>(d, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("bulk copy: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(1024);
        cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int k = 0; k < 1024; ++k) bad += o[k] != h[k];
        printf("bulk copy: %d wrong\n", bad);
    }
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    memset(&map, 0xab, sizeof(map));
    const cuuint64_t dims[2] = {100, 100};
    const cuuint64_t strides[1] = {400};
    const cuuint32_t box[2] = {32, 8};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d, gaddr %p, fn %p\n", (int)r, (void*)d, fn);
    const unsigned long long* w = (const unsigned long long*)&map;
    for (int k = 0; k < 16; ++k) printf("%016llx%c", w[k], (k & 3) == 3 ? '\n' : ' ');
    int drv = 0, rt = 0;
    cudaDriverGetVersion(&drv);
    cudaRuntimeGetVersion(&rt);
    printf("driver %d runtime %d\n", drv, rt);
    return 0;
}
