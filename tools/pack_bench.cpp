// Host-side microbenchmark of the cloud packer (gg_host_pack_cloud_range of libgroundgrid_b200_host.so):
// T threads repack disjoint clouds, reports points/s and bytes read/s.  Build:
//   g++ -O2 -std=c++17 tools/pack_bench.cpp -o /tmp/pack_bench -ldl -lpthread
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

typedef int (*pack_fn)(const void*, size_t, void*);

int main(int argc, char** argv) {
    const char* lib = argc > 1 ? argv[1] : "groundgrid_b200/libgroundgrid_b200_host.so";
    const int max_threads = argc > 2 ? atoi(argv[2]) : 8;
    void* h = dlopen(lib, RTLD_NOW);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 1; }
    // mode 0: streaming stores, one destination per cloud; 1: plain stores into a small reused ring of
    // destinations (stays in the last-level cache); 2: streaming stores into the ring
    const int mode = argc > 3 ? atoi(argv[3]) : 0;
    const size_t ring = argc > 4 ? atoi(argv[4]) : 16;
    pack_fn pack = (pack_fn)dlsym(h, mode == 1 ? "gg_host_pack_cloud_cached" : "gg_host_pack_cloud");
    if (!pack) { fprintf(stderr, "no symbol\n"); return 1; }
    const size_t n = 120000, clouds = 64;
    std::vector<void*> src(clouds), dst(clouds);
    for (size_t c = 0; c < clouds; ++c) {
        src[c] = aligned_alloc(4096, n * 32);
        dst[c] = aligned_alloc(4096, n * 14 + 4096);
        float* f = (float*)src[c];
        for (size_t i = 0; i < n * 8; ++i) f[i] = (float)(i * 0.001 + c);
        memset(dst[c], 0, n * 14 + 4096);
    }
    for (int t = max_threads >= 8 ? 8 : 1; t <= max_threads; t *= 2) {
        double best = 1e30;
        for (int rep = 0; rep < 5; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int k = 0; k < t; ++k)
                th.emplace_back([&, k] { for (size_t c = k; c < clouds; c += t) pack(src[c], n, dst[mode ? c % ring : c]); });
            for (auto& x : th) x.join();
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        printf("mode %d threads %2d: %.1f Mpts/s  read %.2f GB/s  (%.1f Mpts/s/thread)\n", mode, t, clouds * n / best / 1e6, clouds * n * 32 / best / 1e9, clouds * n / best / 1e6 / t);
    }
    return 0;
}
