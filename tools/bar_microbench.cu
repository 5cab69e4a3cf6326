// Micro-benchmark: what does one wavefront level cost?  (one CTA, 512 threads, 1000 levels)
//   variant 0: L1-hit load chain + math + barrier                      (no stores)
//   variant 1: + st.global of the result before the barrier
//   variant 2: + st.shared instead of st.global
//   variant 3: st.global, then the NEXT level loads the value another thread just stored
//   variant 4: like 3 but the value travels through shared memory, global store still issued
//   variant 5: like 4 but no global store at all
//   variant 6: like 3 with __ldcg (L2) loads
//   variant 7: only 128 of 512 threads store
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) k(float* g, int variant, int levels, long long* out) {
    __shared__ float sh[1024];
    const int tid = threadIdx.x;
    float acc = (float)tid;
    sh[tid] = acc; sh[tid + 512] = acc;
    __syncthreads();
    const long long t0 = clock64();
    for (int l = 0; l < levels; ++l) {
        const int src = (tid + 1) & 511;          // neighbour's slot
        float v;
        if (variant == 3 || variant == 7) v = g[src + (l & 1) * 512];          // written last level by thread src
        else if (variant == 6) v = __ldcg(g + src + (l & 1) * 512);
        else if (variant == 4 || variant == 5) v = sh[src + (l & 1) * 512];
        else v = g[4096 + ((tid * 7 + l) & 1023)];                                // stable data (L1 hits)
        acc = __fdiv_rn(__fadd_rn(__fmul_rn(acc, 0.999f), v), 1.0001f);
        const int dst = tid + ((l + 1) & 1) * 512;
        if (variant == 1 || variant == 3 || variant == 4 || variant == 6) g[dst] = acc;
        if (variant == 7 && tid < 128) g[dst] = acc;
        if (variant == 2 || variant == 4 || variant == 5) sh[dst] = acc;
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    g[8192 + tid] = acc;
}
int main() {
    float* g; long long* out; cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20); cudaMalloc(&out, 8);
    for (int variant = 0; variant < 8; ++variant) {
        for (int rep = 0; rep < 2; ++rep) k<<<1, 512>>>(g, variant, 1000, out);
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("variant %d: %.1f cycles/level\n", variant, c / 1000.0);
    }
    // the same with 64 CTAs resident (one per SM) to see L2 contention effects
    for (int variant : {1, 3}) {
        k<<<64, 512>>>(g, variant, 1000, out); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("variant %d x64 CTAs (racy addresses, timing only): %.1f cycles/level\n", variant, c / 1000.0);
    }
    return 0;
}
