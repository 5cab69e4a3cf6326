// Micro-benchmark 2: per-level cost of a barrier-separated dependent step, by block size and
// arithmetic.  Exchange through shared memory only.
//   op 0: fadd only            op 1: fmul+fadd+__fdiv_rn       op 2: 9 mul + tree + div + 3 (the spiral chain)
//   op 3: op 2 + fp64 decay (ddiv)                             op 4: no arithmetic (ld, st, bar)
//   sync 0: __syncthreads   sync 1: no barrier at all (racy; lower bound)   sync 2: __syncwarp only (threads==32)
#include <cstdio>
#include <cfloat>
#include <cuda_runtime.h>
template <int OP, int SYNC>
__global__ void k(float* g, int levels, long long* out) {
    __shared__ float sh[2048];
    const int tid = threadIdx.x, nt = blockDim.x;
    float acc = 1.0f + tid * 1e-3f;
    sh[tid] = acc; sh[tid + 1024] = acc;
    __syncthreads();
    const long long t0 = clock64();
    for (int l = 0; l < levels; ++l) {
        const int src = (tid + 1 == nt ? 0 : tid + 1) + (l & 1) * 1024;
        float v = sh[src];
        if (OP == 0) acc = __fadd_rn(acc, v);
        if (OP == 1) acc = __fdiv_rn(__fadd_rn(__fmul_rn(acc, 0.999f), v), 1.0001f);
        if (OP == 2 || OP == 3) {
            float c[9], p[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) { c[q] = 0.1f + q * 0.01f; p[q] = __fmul_rn(c[q], q == 3 ? v : acc); }
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(c[0], c[1]), __fadd_rn(c[2], c[3])), __fadd_rn(__fadd_rn(c[4], c[5]), __fadd_rn(c[6], __fadd_rn(c[7], c[8])))), FLT_MIN);
            const float t = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), __fadd_rn(p[2], p[3])), __fadd_rn(__fadd_rn(p[4], p[5]), __fadd_rn(p[6], __fadd_rn(p[7], p[8]))));
            const float avg = __fdiv_rn(t, s);
            acc = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, c[4]), avg), __fmul_rn(c[4], acc));
            if (OP == 3) { const double o = (double)c[4]; const double d = __dsub_rn(o, __ddiv_rn(o, 5.0 + acc)); g[100000 + tid] = (float)(d < 0.001 ? 0.001 : d); }
        }
        if (OP == 4) acc = v;
        sh[tid + ((l + 1) & 1) * 1024] = acc;
        if (SYNC == 0) __syncthreads();
        if (SYNC == 2) __syncwarp();
    }
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    g[tid] = acc;
}
template <int OP, int SYNC>
void run(float* g, long long* out, int threads, const char* name) {
    for (int rep = 0; rep < 2; ++rep) k<OP, SYNC><<<1, threads>>>(g, 2000, out);
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s threads=%4d : %7.1f cycles/level\n", name, threads, c / 2000.0);
}
int main() {
    float* g; long long* out; cudaMalloc(&g, 1 << 22); cudaMemset(g, 0, 1 << 22); cudaMalloc(&out, 8);
    for (int t : {32, 64, 128, 256, 512, 1024}) run<4, 0>(g, out, t, "ld+st+bar (no math)");
    for (int t : {32, 64, 128, 256, 512, 1024}) run<0, 0>(g, out, t, "fadd");
    for (int t : {32, 128, 256, 512}) run<1, 0>(g, out, t, "fmul+fadd+fdiv");
    for (int t : {32, 128, 256, 512}) run<2, 0>(g, out, t, "spiral chain (9 mul, tree, div, 3)");
    for (int t : {32, 128, 256, 512}) run<3, 0>(g, out, t, "spiral chain + fp64 decay");
    for (int t : {32, 256}) run<2, 1>(g, out, t, "spiral chain, NO barrier (lower bound)");
    run<2, 2>(g, out, 32, "spiral chain, syncwarp only");
    run<4, 2>(g, out, 32, "ld+st, syncwarp only");
    run<4, 1>(g, out, 32, "ld+st, no sync");
    return 0;
}
