// C-ABI of groundgrid_b200 (include/groundgrid_b200.h): handle management, device arena,
// parameter staging, stream pipeline.  All compute happens in gg_kernels.cu; there is no CPU
// fallback -- without a usable sm_100 device every compute call returns GG_E_CUDA.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "gg_host.h"
#include "gg_internal.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define GG_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) return fail(GG_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

constexpr int kStreams = 8;   // upper bound; GG_STREAMS (default 8) picks how many are used
constexpr int kRing = 256;

// CUDA-event pairs around every kernel launch while profiling is enabled.
struct EventProfiler : gg::Profiler {
    struct Rec {
        int id;
        cudaEvent_t a, b;
    };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    size_t dropped = 0;
    static constexpr size_t kMaxRecs = 16384;
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int id, cudaStream_t st) override {
        if (recs.size() >= kMaxRecs) {
            ++dropped;
            cur = nullptr;
            return;
        }
        recs.push_back({id, get(), get()});
        cur = &recs.back();
        cudaEventRecord(cur->a, st);
    }
    void end(int, cudaStream_t st) override {
        if (cur) cudaEventRecord(cur->b, st);
        cur = nullptr;
    }
    // after the streams were synchronised
    void collect(double* ms, uint32_t* count) {
        for (Rec& r : recs) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
                ms[r.id] += t;
                count[r.id] += 1;
            }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
    ~EventProfiler() override {
        for (Rec& r : recs) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    }
    Rec* cur = nullptr;
};

// Host-side packing of PointXYZIR clouds for the batched host-buffer path: the 32-byte records
// carry 14 useful bytes (x, y, z, ring); PCIe is the bottleneck of that path, so worker threads
// repack every cloud into pinned staging memory as x | y | z (float[n_pad]) | ring (u16[n_pad]), n_pad = n rounded up to 8,
// with streaming stores, and only 14 bytes per point cross the bus.
struct PackJob {
    const gg_point* src = nullptr;
    size_t n = 0;
    std::atomic<int> remaining{0};
};

class HostPacker {
  public:
    static constexpr size_t kChunk = 16384;  // points per work item (multiple of 8)
    explicit HostPacker(int threads) {
        for (int t = 0; t < threads; ++t) workers_.emplace_back([this] { loop(); });
    }
    ~HostPacker() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    int threads() const { return (int)workers_.size(); }

    // Staging ring: packed job number `seq` (counted over the lifetime of the handle) is written to slot
    // seq % slots.  The ring is small enough to stay in the last-level cache, so the packers' (plain) stores
    // and the DMA reads that follow mostly stay out of DRAM; a slot is reused once `copied` -- the number of
    // packed jobs whose H2D copy has completed, advanced by the feeding thread -- has passed its last user.
    void set_ring(unsigned char* base, size_t stride, int slots) {
        ring_ = base;
        stride_ = stride;
        slots_ = slots;
    }
    unsigned char* slot_of(uint64_t seq) const { return ring_ + (size_t)(seq % (uint64_t)slots_) * stride_; }
    std::atomic<uint64_t> copied{0};
    std::atomic<uint64_t> pack_ns{0}, slot_wait_ns{0};  // summed over the worker threads (diagnostics)

    // jobs must stay alive until every job is either packed or claimed raw (or cancel() has returned); chunks go
    // out in job order; job j is packed job number seq_base + j
    void start(std::vector<PackJob>* jobs, uint64_t seq_base) {
        while (busy_.load(std::memory_order_acquire) != 0) std::this_thread::yield();  // stragglers of the previous run
        std::vector<std::pair<int, int>> chunks;
        std::vector<size_t> first_chunk;
        for (size_t j = 0; j < jobs->size(); ++j) {
            PackJob& job = (*jobs)[j];
            const int nch = (int)std::max<size_t>(1, (job.n + kChunk - 1) / kChunk);
            job.remaining.store(nch, std::memory_order_relaxed);
            first_chunk.push_back(chunks.size());
            for (int c = 0; c < nch; ++c) chunks.push_back({(int)j, c});
        }
        {
            // everything a claim reads changes under the claim lock, so a worker that wakes up late sees either the
            // finished previous run (nothing to claim) or this one completely
            std::lock_guard<std::mutex> g(claim_mu_);
            chunks_.swap(chunks);
            first_chunk_.swap(first_chunk);
            jobs_ = jobs;
            seq_base_ = seq_base;
            next_ = 0;
            limit_ = chunks_.size();
            raw_from_ = (int)jobs->size();
            cancel_.store(false, std::memory_order_relaxed);
        }
        {
            std::lock_guard<std::mutex> g(mu_);
            ++epoch_;
        }
        cv_.notify_all();
    }
    // Error path of the caller: nothing more is claimed, chunks in flight are waited for; afterwards the job list
    // may be destroyed.
    void cancel() {
        {
            std::lock_guard<std::mutex> g(claim_mu_);
            limit_ = next_;
            cancel_.store(true, std::memory_order_release);
        }
        while (inflight_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    }
    // Take the last job nobody has started packing yet out of the packers' hands (it will be sent as
    // plain 32-byte records).  Returns its index or -1.
    int claim_raw_from_back() {
        std::lock_guard<std::mutex> g(claim_mu_);
        if (raw_from_ <= 0) return -1;
        const int j = raw_from_ - 1;
        if (first_chunk_[j] < next_) return -1;  // a packer is already on it
        raw_from_ = j;
        limit_ = first_chunk_[j];
        return j;
    }
    bool packed(const PackJob& job) const { return job.remaining.load(std::memory_order_acquire) <= 0; }
    bool help() { return work_one(false); }  // the calling thread packs one chunk if one can be started right away

    static void pack_range(const gg_point* src, size_t n, unsigned char* dst, size_t i0, size_t i1, bool cached = false) {
        gg::pack_cloud_range(src, n, dst, i0, i1, cached);
    }

  private:
    bool work_one(bool may_wait) {
        std::vector<PackJob>* jobs;
        size_t c;
        uint64_t seq;
        {
            std::lock_guard<std::mutex> g(claim_mu_);
            jobs = jobs_;
            if (!jobs || next_ >= limit_) return false;
            seq = seq_base_ + (uint64_t)chunks_[next_].first;
            if (!may_wait && seq >= copied.load(std::memory_order_acquire) + (uint64_t)slots_) return false;
            c = next_++;
            inflight_.fetch_add(1, std::memory_order_acq_rel);
        }
        const std::pair<int, int> chunk = chunks_[c];  // chunks_ only changes in start(), which waits for busy_ == 0
        const auto t0 = std::chrono::steady_clock::now();
        bool go = true;
        while (seq >= copied.load(std::memory_order_acquire) + (uint64_t)slots_) {  // the slot's previous cloud is still on its way
            if (stop_ || cancel_.load(std::memory_order_acquire)) {
                go = false;
                break;
            }
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        const auto t1 = std::chrono::steady_clock::now();
        if (go) {
            PackJob& job = (*jobs)[chunk.first];
            const size_t i0 = (size_t)chunk.second * kChunk;
            pack_range(job.src, job.n, slot_of(seq), i0, std::min(job.n, i0 + kChunk), true);
            job.remaining.fetch_sub(1, std::memory_order_release);
        }
        const auto t2 = std::chrono::steady_clock::now();
        slot_wait_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count(), std::memory_order_relaxed);
        pack_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t2 - t1).count(), std::memory_order_relaxed);
        inflight_.fetch_sub(1, std::memory_order_acq_rel);
        return go;
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                busy_.fetch_add(1, std::memory_order_acq_rel);
            }
            while (work_one(true)) {
            }
            busy_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, claim_mu_;
    std::condition_variable cv_;
    std::vector<PackJob>* jobs_ = nullptr;
    std::vector<std::pair<int, int>> chunks_;
    std::vector<size_t> first_chunk_;
    unsigned char* ring_ = nullptr;
    size_t stride_ = 0;
    int slots_ = 1;
    uint64_t seq_base_ = 0;
    size_t next_ = 0, limit_ = 0;
    int raw_from_ = 0;
    std::atomic<int> busy_{0}, inflight_{0};
    std::atomic<bool> cancel_{false};
    uint64_t epoch_ = 0;
    std::atomic<bool> stop_{false};
};

struct SlotState {
    bool have_map = false;
    double px = 0.0, py = 0.0;
    size_t n_points = 0;       // points of the resident / last scan
    int last_stop = 0;         // stop_after of the last run (decides what "points" names)
    bool ran = false;
    bool output_valid = false;
    const gg_point* src = nullptr;  // caller-owned device cloud of the last scan (null: the slot's own buffer)
    const float* packed_input = nullptr;  // last scan came through the packed host path (no 32-byte records on the device)
};

}  // namespace

struct gg_handle_s {
    int device = 0;
    int n_slots = 0;
    size_t pcap = 0;
    unsigned flags = 0;
    double dimension_m = 0.0;
    float resolution = 0.f;
    gg_config cfg{};
    gg::View view{};
    std::vector<SlotState> slots;
    cudaStream_t streams[kStreams] = {};
    bool own_streams = true;
    int n_streams = kStreams;
    // parameter staging ring: pinned host copy + device copy per entry
    gg::SlotParams* h_ring = nullptr;
    gg::SlotParams* d_ring = nullptr;
    cudaEvent_t ring_ev[kRing] = {};
    bool ring_used[kRing] = {};
    int ring_pos = 0;
    uint64_t launches = 0;
    std::vector<void*> dev_allocs;
    std::vector<unsigned char> seen_scratch;  // duplicate-slot check of the batch calls
    cudaEvent_t stagger_ev[kStreams] = {};    // recorded by an even stream before its spiral, awaited by its odd partner
    bool stagger_armed[kStreams] = {};
    int stagger = 0;                          // GG_STAGGER (measured on B200: no gain, the bulk kernels leave the spiral no room; off)
    int sched_levels = 0, sched_visits = 0, sched_max = 0;
    bool out_cloud_ready = false;
    // f1: device copy of a PointCloud2 payload, one buffer per stream group (the copy and the unpack kernel of a slot
    // run on the slot's stream; stream order then keeps two slots of one group from overwriting each other's payload)
    unsigned char* d_raw[kStreams] = {};
    size_t d_raw_cap[kStreams] = {};
    float* d_image = nullptr;        // f3: terrain image staging (N * N * 3)
    unsigned char* d_image_u8 = nullptr;  // f3: 8-bit layer image staging (N * N)
    float* d_minmax = nullptr;       //     and its min / max keys
    unsigned long long* d_eval = nullptr;  // f4: [EVAL_LABELS][2] tallies
    HostPacker* packer = nullptr;    // created on the first packed batch call
    // gg_filter_cloud_batch[_begin] alternates between two sets of input / label buffers ("parity"), so the
    // clouds of batch t+1 can be packed and copied while the kernels of batch t still read theirs
    unsigned char* h_stage = nullptr;  // pinned staging ring of the packers, [pack_slots][14 * pcap] (HostPacker::set_ring)
    int pack_slots = 32;               // GG_PACK_RING
    std::vector<cudaEvent_t> slot_ev;  // [pack_slots] H2D of the slot's current cloud
    uint64_t pack_issued = 0;          // packed jobs whose copy has been enqueued (HostPacker::copied counts the completed ones)
    unsigned char* in_packed[2] = {};  // device, same shape
    gg_point* in_raw[2] = {};          // device 32-byte records; [0] is the slots' own buffer (view.points)
    uint8_t* labels_buf[2] = {};       // [0] is the buffer the handle was created with
    int batch_parity = 0;
    bool batch_outstanding[2] = {};
    cudaEvent_t batch_done[2] = {};    // on copy_out: kernels and label copies of the batch have finished
    int host_pack = 1;               // GG_HOST_PACK=0 sends the 32-byte records as they are
    int launch_unit = 32;            // GG_LAUNCH_UNIT: scans per kernel launch set in gg_filter_cloud_batch
    int host_pack_mix = 1;           // GG_HOST_PACK=1 pins "pack everything"; default: pack and send raw side by side
    cudaStream_t copy_in[10] = {}, copy_out = nullptr;  // [0, n_copy_in): packed clouds, then 2 for raw clouds;  // H2D / D2H of gg_filter_cloud_batch, never behind kernels
    int n_copy_in = 4;                                   // GG_COPY_STREAMS
    int main_help = 1;                                   // GG_MAIN_HELP
    float4* detect_tab = nullptr;  // per-cell constants of the patch detection, rebuilt by gg_set_config
    CUtensorMap layer_map{};        // TMA descriptor of the layer arena (k_detect_tma); valid iff have_layer_map
    bool have_layer_map = false;
    bool inputs_busy = false;  // asynchronous work that reads or writes the slots' input buffers may be in flight
    std::vector<cudaEvent_t> batch_ev;                    // unit hand-over events of gg_filter_cloud_batch
    cudaEvent_t raw_ev[8] = {};      // throttle of the raw copies issued by the mixing loop
    int raw_gate = 2;                // GG_RAW_GATE (legacy, only with GG_BUS_TARGET_MB=0): no raw copies while this many packed clouds wait for the bus
    size_t bus_target = 8u << 20;    // GG_BUS_TARGET_MB: raw clouds are added while fewer bytes than this are in flight on the bus
    int raw_depth = 4;               // GG_RAW_DEPTH: raw copies in flight before the loop stops claiming more
    size_t last_raw = 0, last_packed = 0;  // scans sent raw / packed by the last batch call
    size_t last_raw_bytes = 0, last_packed_bytes = 0;
    size_t last_feed_us = 0, last_total_us = 0;
    size_t last_pack_us = 0, last_slot_wait_us = 0, last_idle_us = 0;  // worker-thread sums; feeder time with nothing to enqueue  // host time until the last cloud was enqueued / until everything was done
    EventProfiler* prof = nullptr;   // non-null while profiling is enabled
    double prof_ms[gg::K_NUM] = {};
    uint32_t prof_count[gg::K_NUM] = {};
};

namespace {

template <typename T>
int dev_alloc(gg_handle h, T** p, size_t count) {
    void* q = nullptr;
    GG_CUDA(cudaMalloc(&q, count * sizeof(T) + 256));
    h->dev_allocs.push_back(q);
    *p = static_cast<T*>(q);
    return GG_OK;
}

int check_slot(gg_handle h, int slot) {
    if (!h) return fail(GG_E_ARG, "null handle");
    if (slot < 0 || slot >= h->n_slots) return fail(GG_E_ARG, "slot %d out of range [0, %d)", slot, h->n_slots);
    return GG_OK;
}

int layer_index(gg_handle h, int slot, const char* name, int* idx) {
    struct Entry {
        const char* name;
        int idx;
        bool full_only;
    };
    static const Entry table[] = {
        {"ground", gg::L_GROUND, false},        {"groundpatch", gg::L_GROUNDPATCH, false},
        {"variance", gg::L_VARIANCE, false},    {"minGroundHeight", gg::L_MINH, false},
        {"maxGroundHeight", gg::L_MAXH, true},  {"groundCandidates", gg::L_GCAND, true},
        {"planeDist", gg::L_PLANEDIST, true},   {"m2", gg::L_M2, true},
        {"meanVariance", gg::L_MEAN, true},     {"pointsRaw", gg::L_RAW, true},
        {"count", gg::L_COUNT, false},          {"obstacles", gg::L_OBSTACLES, false},
    };
    if (!name) return fail(GG_E_ARG, "null layer name");
    if (std::strcmp(name, "points") == 0) {
        // the reference reuses "points": kept-point count while rasterising (:309), non-ground
        // point count after the label loop (:147,176)
        const SlotState& s = h->slots[slot];
        *idx = (s.ran && s.last_stop != 0) ? gg::L_COUNT : gg::L_OBSTACLES;
        return GG_OK;
    }
    for (const Entry& e : table)
        if (std::strcmp(name, e.name) == 0) {
            if (e.full_only && !(h->flags & GG_FLAG_FULL_LAYERS)) return fail(GG_E_LAYER, "layer '%s' needs GG_FLAG_FULL_LAYERS", name);
            *idx = e.idx;
            return GG_OK;
        }
    return fail(GG_E_LAYER, "unknown layer '%s'", name);
}

// Reserve the next staging entry (waits only if the ring wrapped onto an in-flight entry).
int ring_acquire(gg_handle h, gg::SlotParams** host, gg::SlotParams** dev, int* pos) {
    const int p = h->ring_pos;
    h->ring_pos = (p + 1) % kRing;
    if (h->ring_used[p]) GG_CUDA(cudaEventSynchronize(h->ring_ev[p]));
    *host = h->h_ring + (size_t)p * h->n_slots;
    *dev = h->d_ring + (size_t)p * h->n_slots;
    *pos = p;
    return GG_OK;
}

int ring_commit(gg_handle h, int pos, int count, cudaStream_t st) {
    gg::SlotParams* host = h->h_ring + (size_t)pos * h->n_slots;
    gg::SlotParams* dev = h->d_ring + (size_t)pos * h->n_slots;
    GG_CUDA(cudaMemcpyAsync(dev, host, (size_t)count * sizeof(gg::SlotParams), cudaMemcpyHostToDevice, st));
    return GG_OK;
}

// The entry may be reused once the kernels that read it have finished (they may run on any of
// the handle's streams, so the event is recorded AFTER the launches, not after the copy).
int ring_release(gg_handle h, int pos, cudaStream_t st) {
    GG_CUDA(cudaEventRecord(h->ring_ev[pos], st));
    h->ring_used[pos] = true;
    return GG_OK;
}

// Slots are bound to streams in contiguous groups; everything that touches a slot is enqueued
// on its stream, so scans of different groups overlap (the latency-bound spiral of one group
// runs under the bandwidth-bound kernels of another) without any cross-stream dependency.
int stream_index(gg_handle h, int slot) { return (int)((long long)slot * h->n_streams / h->n_slots); }
cudaStream_t stream_of(gg_handle h, int slot) { return h->streams[stream_index(h, slot)]; }

void fill_params(gg_handle h, const gg_scan_desc& d, gg::SlotParams& p, const gg_point* src, const float* packed = nullptr) {
    const SlotState& s = h->slots[d.slot];
    std::memset(&p, 0, sizeof(p));
    p.px = s.px;
    p.py = s.py;
    p.ox = d.origin[0];
    p.oy = d.origin[1];
    p.oz = d.origin[2];
    p.base_z_f = (float)d.base_z;  // ggl(c, c) = ps.point.z (double -> float), GroundSegmentation.cpp:411
    p.n_points = (int)d.n_points;
    p.slot = d.slot;
    p.src = src ? src : h->view.points + (size_t)d.slot * h->pcap;
    p.packed = packed;
}

// enqueue the kernels of `count` scans, each group of slots on its own stream
int run_scans_grouped(gg_handle h, int count, const gg_scan_desc* scans, int stop_after, const gg_point* const* dev_points = nullptr,
                      const float* const* packed_ptrs = nullptr, uint8_t* labels_base = nullptr) {
    if (count <= 0) return GG_OK;
    if (count > h->n_slots) return fail(GG_E_ARG, "count %d exceeds the number of slots %d", count, h->n_slots);
    int rc;
    std::vector<unsigned char>& seen = h->seen_scratch;
    seen.assign((size_t)h->n_slots, 0);
    for (int i = 0; i < count; ++i) {
        const gg_scan_desc& d = scans[i];
        if ((rc = check_slot(h, d.slot))) return rc;
        if (seen[d.slot]++) return fail(GG_E_ARG, "slot %d appears twice in one batch (scans of a batch run concurrently)", d.slot);
        if (!h->slots[d.slot].have_map) return fail(GG_E_STATE, "slot %d: map not initialised", d.slot);
        if (d.n_points > h->pcap) return fail(GG_E_ARG, "slot %d: %zu points exceed capacity %zu", d.slot, d.n_points, h->pcap);
    }
    gg::View view = h->view;
    if (labels_base) view.labels = labels_base;
    for (int g = 0; g < h->n_streams; ++g) {
        gg::SlotParams *hp = nullptr, *dp = nullptr;
        int pos = 0, m = 0, max_points = 0;
        for (int i = 0; i < count; ++i) {
            const gg_scan_desc& d = scans[i];
            if (stream_index(h, d.slot) != g) continue;
            if (m == 0 && (rc = ring_acquire(h, &hp, &dp, &pos))) return rc;
            const float* packed = packed_ptrs ? packed_ptrs[i] : nullptr;
            fill_params(h, d, hp[m], dev_points ? dev_points[i] : nullptr, packed);
            ++m;
            max_points = std::max(max_points, (int)d.n_points);
            SlotState& s = h->slots[d.slot];
            s.n_points = d.n_points;
            s.last_stop = stop_after;
            s.ran = true;
            s.output_valid = false;
            s.src = dev_points ? dev_points[i] : nullptr;
            s.packed_input = packed;
        }
        if (m == 0) continue;
        cudaStream_t st = h->streams[g];
        if ((rc = ring_commit(h, pos, m, st))) return rc;
        // Staggered stream pairs: the spiral is latency bound (a CTA per scan, a few warps per SM), every other kernel
        // fills the machine.  Stream 2k + 1 starts its scans when stream 2k has reached its spiral, so in steady state the
        // spirals of one half of the batch run underneath the bulk kernels of the other half instead of all at once.
        cudaEvent_t after_detect = nullptr;
        if (h->stagger && h->n_streams > 1 && stop_after == 0) {
            if ((g & 1) == 0 && g + 1 < h->n_streams) {
                after_detect = h->stagger_ev[g];
                h->stagger_armed[g] = true;
            } else if ((g & 1) == 1 && h->stagger_armed[g - 1]) {
                GG_CUDA(cudaStreamWaitEvent(st, h->stagger_ev[g - 1], 0));
                h->stagger_armed[g - 1] = false;
            }
        }
        h->launches += gg::launch_scan_pipeline(view, dp, m, max_points, stop_after, st, h->prof, h->have_layer_map ? &h->layer_map : nullptr, after_detect);
        GG_CUDA(cudaGetLastError());
        if ((rc = ring_release(h, pos, st))) return rc;
    }
    return GG_OK;
}

// TMA descriptor of the layer arena seen as a 3-D fp32 tensor (i fastest, j, plane = slot * n_layers + layer), box
// 40 x 12 x 1 = the halo tile of k_detect_tma (the box starts at i0 - 4: TMA wants a 16-byte aligned innermost start).  cuTensorMapEncodeTiled is a driver-API call; it is resolved through the
// runtime (cudaGetDriverEntryPoint) so that the library needs no link-time libcuda.  Row pitch must be a multiple of
// 16 bytes: maps with N % 4 != 0 keep the plain-load kernel.  GG_DETECT_TMA=0 forces that kernel too.
bool encode_layer_map(gg_handle h) {
    const gg::View& v = h->view;
    if (v.k.N % 4 != 0) return false;
    if (const char* e = getenv("GG_DETECT_TMA"))
        if (atoi(e) == 0) return false;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)v.k.N, (cuuint64_t)v.k.N, (cuuint64_t)h->n_slots * (cuuint64_t)v.n_layers};
    const cuuint64_t strides[2] = {(cuuint64_t)v.k.N * sizeof(float), (cuuint64_t)v.k.N2 * sizeof(float)};
    const cuuint32_t box[3] = {40u, 12u, 1u};   // DT_WT x DT_R of k_detect_tma
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(&h->layer_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v.layers, dims, strides, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int ensure_out_cloud(gg_handle h) {
    if (h->out_cloud_ready) return GG_OK;
    int rc = dev_alloc(h, &h->view.out_cloud, (size_t)h->n_slots * h->pcap);
    if (rc) return rc;
    h->out_cloud_ready = true;
    return GG_OK;
}

int run_output_on(gg_handle h, int slot, bool want_cloud, cudaStream_t st) {
    SlotState& s = h->slots[slot];
    if (!s.ran || s.last_stop != 0) return fail(GG_E_STATE, "slot %d: no completed scan", slot);
    if (want_cloud && s.packed_input)
        return fail(GG_E_STATE, "slot %d: the output cloud needs the 32-byte records on the device (use gg_filter_cloud or GG_HOST_PACK=0)", slot);
    int rc;
    if (want_cloud && (rc = ensure_out_cloud(h))) return rc;
    gg::SlotParams *hp = nullptr, *dp = nullptr;
    int pos = 0;
    if ((rc = ring_acquire(h, &hp, &dp, &pos))) return rc;
    std::memset(&hp[0], 0, sizeof(gg::SlotParams));
    hp[0].slot = slot;
    hp[0].n_points = (int)s.n_points;
    hp[0].src = s.src ? s.src : h->view.points + (size_t)slot * h->pcap;
    if ((rc = ring_commit(h, pos, 1, st))) return rc;
    h->launches += gg::launch_output(h->view, dp, 1, (int)s.n_points, want_cloud, st, h->prof);
    GG_CUDA(cudaGetLastError());
    if ((rc = ring_release(h, pos, st))) return rc;
    s.output_valid = true;
    return GG_OK;
}

}  // namespace

extern "C" {

void gg_default_config(gg_config* c) {
    if (!c) return;
    c->point_count_cell_variance_threshold = 10;
    c->max_ring = 1024;
    c->groundpatch_detection_minimum_threshold = 0.01;
    c->distance_factor = 0.0001;
    c->minimum_distance_factor = 0.0005;
    c->miminum_point_height_threshold = 0.3;
    c->minimum_point_height_obstacle_threshold = 0.1;
    c->outlier_tolerance = 0.1;
    c->ground_patch_detection_minimum_point_count_threshold = 0.25;
    c->patch_size_change_distance = 20.0;
    c->occupied_cells_decrease_factor = 5.0;
    c->occupied_cells_point_count_factor = 20.0;
    c->min_outlier_detection_ground_confidence = 1.25;
    c->thread_count = 8;
}

const char* gg_last_error(void) { return g_last_error.c_str(); }

int gg_create(double dimension_m, float resolution, int device, int n_slots, size_t max_points, unsigned flags, void* stream,
              gg_handle* out) {
    if (!out) return fail(GG_E_ARG, "null out pointer");
    *out = nullptr;
    if (n_slots <= 0 || max_points == 0 || !(resolution > 0.f) || !(dimension_m > 0.0)) return fail(GG_E_ARG, "bad geometry / sizes");
    if (max_points > (1u << 26) - 256) return fail(GG_E_ARG, "max_points above 2^26 per cloud is not supported");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(GG_E_CUDA, "no CUDA device available (groundgrid_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(GG_E_ARG, "device %d out of range", device);
    cudaDeviceProp prop;
    GG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(GG_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    GG_CUDA(cudaSetDevice(device));

    const int n = gg::cells_per_side(dimension_m, resolution);
    if (n < 8 || n > 4096) return fail(GG_E_ARG, "unsupported map size %d cells", n);

    gg_handle h = new gg_handle_s();
    h->device = device;
    h->n_slots = n_slots;
    h->pcap = (max_points + 255) / 256 * 256;
    h->flags = flags;
    h->dimension_m = dimension_m;
    h->resolution = resolution;
    gg_default_config(&h->cfg);
    h->slots.resize(n_slots);
    gg::View& v = h->view;
    gg::derive_constants(h->cfg, dimension_m, resolution, flags, v.k);
    if (v.k.N != n) {
        const int n_geo = v.k.N;
        delete h;
        return fail(GG_E_ARG, "cell count mismatch between init (%d) and setGeometry (%d)", n, n_geo);
    }
    const size_t N2 = (size_t)v.k.N2;
    v.n_layers = (flags & GG_FLAG_FULL_LAYERS) ? gg::L_NUM : gg::L_NUM_LIVE;
    v.pcap = h->pcap;
    v.out_blocks = (int)((h->pcap + gg::OUT_TILE - 1) / gg::OUT_TILE);

#define GG_TRY(expr)            \
    do {                        \
        int rc__ = (expr);      \
        if (rc__) {             \
            gg_destroy(h);      \
            return rc__;        \
        }                       \
    } while (0)
#define GG_CUDA_TRY(call)                                                                             \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            gg_destroy(h);                                                                            \
            return fail(GG_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                             \
    } while (0)

    const size_t S = (size_t)n_slots, P = h->pcap;
    GG_TRY(dev_alloc(h, &v.layers, S * v.n_layers * N2));
    h->have_layer_map = encode_layer_map(h);
    float* expected = nullptr;
    GG_TRY(dev_alloc(h, &expected, N2));
    v.expected = expected;
    GG_TRY(dev_alloc(h, &h->detect_tab, N2));
    v.detect_tab = h->detect_tab;
    GG_TRY(dev_alloc(h, &v.points, S * P));
    GG_TRY(dev_alloc(h, &v.zw, S * P));
    GG_TRY(dev_alloc(h, &v.zsorted, S * P));
    GG_TRY(dev_alloc(h, &v.runj, S * P));
    GG_TRY(dev_alloc(h, &v.rundir, S * P));
    GG_TRY(dev_alloc(h, &v.dist, S * P));
    GG_TRY(dev_alloc(h, &v.code, S * P));
    GG_TRY(dev_alloc(h, &v.labels, S * P));
    GG_TRY(dev_alloc(h, &v.cnt64, S * N2));
    GG_TRY(dev_alloc(h, &v.raw_i, (flags & GG_FLAG_FULL_LAYERS) ? S * N2 : 1));
    GG_TRY(dev_alloc(h, &v.cellstart, S * N2));
    GG_TRY(dev_alloc(h, &v.worklist, S * N2));
    GG_TRY(dev_alloc(h, &v.wl_count, 2 * S));
    v.cell_tiles = (int)((N2 + 4095) / 4096);
    GG_TRY(dev_alloc(h, &v.cell_agg, S * v.cell_tiles * 65));
    GG_TRY(dev_alloc(h, &v.out_index, S * P));
    GG_TRY(dev_alloc(h, &v.out_counts, S * (3 * (size_t)v.out_blocks + 1)));
    GG_TRY(dev_alloc(h, &v.roll_scratch, S * 2 * N2));
    v.out_cloud = nullptr;
    GG_CUDA_TRY(cudaMemset(v.layers, 0, S * v.n_layers * N2 * sizeof(float)));
    // per-cell counters are zero between scans (k_cell_stats resets what it consumed)
    GG_CUDA_TRY(cudaMemset(v.cnt64, 0, S * N2 * sizeof(unsigned long long)));
    if (flags & GG_FLAG_FULL_LAYERS) GG_CUDA_TRY(cudaMemset(v.raw_i, 0, S * N2 * sizeof(int)));

    // expectedPoints table (host libm, like the reference) and the spiral wavefront schedule
    {
        std::vector<float> table;
        gg::build_expected_points(n, table);
        GG_CUDA_TRY(cudaMemcpy(expected, table.data(), N2 * sizeof(float), cudaMemcpyHostToDevice));
        std::vector<int> ls;
        std::vector<uint32_t> vs;
        gg::build_spiral_schedule(n, ls, vs);
        int* d_ls = nullptr;
        uint32_t* d_vs = nullptr;
        GG_TRY(dev_alloc(h, &d_ls, ls.size()));
        GG_TRY(dev_alloc(h, &d_vs, vs.size() + 1));
        GG_CUDA_TRY(cudaMemcpy(d_ls, ls.data(), ls.size() * sizeof(int), cudaMemcpyHostToDevice));
        GG_CUDA_TRY(cudaMemcpy(d_vs, vs.data(), vs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        v.level_start = d_ls;
        v.visits = d_vs;
        v.levels = (int)ls.size() - 1;
        h->sched_levels = v.levels;
        h->sched_visits = (int)vs.size();
        for (size_t l = 0; l + 1 < ls.size(); ++l) h->sched_max = std::max(h->sched_max, ls[l + 1] - ls[l]);
        // records of the pipelined spiral kernel (GG_SPIRAL_DIST=0 selects the plain wavefront kernel)
        v.spiral_recs = nullptr;
        v.spiral_dist = 0;
        v.spiral_threads = h->sched_max <= 512 ? 512 : 1024;
        int dist = 2;
        if (const char* e = getenv("GG_SPIRAL_DIST")) dist = atoi(e);
        std::vector<uint32_t> rc;
        int max_recent = 0;
        if (dist >= 1 && dist <= 3 && h->sched_max <= 1024 && gg::build_spiral_records(n, v.k.res_sq, ls, vs, dist, rc, max_recent)) {
            uint32_t* d_rc = nullptr;
            GG_TRY(dev_alloc(h, &d_rc, rc.size() + 4));
            GG_CUDA_TRY(cudaMemcpy(d_rc, rc.data(), rc.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            v.spiral_recs = reinterpret_cast<const uint4*>(d_rc);
            v.spiral_dist = dist;
        }
        // skewed-layout spiral (GG_SPIRAL_SKEW=0 keeps the pipelined kernel)
        std::memset(&v.skew, 0, sizeof(v.skew));
        int want_skew = 1;
        if (const char* e = getenv("GG_SPIRAL_SKEW")) want_skew = atoi(e);
        gg::SkewTables sk;
        if (want_skew) gg::build_spiral_skew(n, ls, vs, sk);
        // time-sharing of lane threads: the smallest M (multiple of 32) such that ring k + M of a side
        // starts (prefetch window included) only after ring k has finished
        // (only when one thread per lane does not fit a CTA: measured, a dedicated thread per lane is faster)
        // Two thread layouts of the same schedule.  "Latency": one thread per lane whenever that fits a CTA (measured:
        // fastest for a single scan).  "Throughput": the smallest M, so that the CTA is small and two or three scans
        // share an SM (measured: +4 % on batches).  GG_SPIRAL_M raises the throughput layout's M (0: same as latency).
        int M = 32, phases = 1, M_thr = 32, phases_thr = 1;
        if (sk.ok) {
            const int kGap = 2 * 8 + 4;  // 2 * PF_FAR of k_spiral_skew + slack
            auto smallest_fitting = [&](int m0) {
                int m = m0;
                for (; m < sk.KP; m += 32) {
                    bool fits = true;
                    for (int sd = 0; sd < 4 && fits; ++sd)
                        for (int c0 = 0; c0 + m < sk.KP && fits; ++c0) {
                            const int a = sd * sk.KP + c0, b2 = a + m;
                            if (sk.lane_begin[a] < sk.lane_end[a] && sk.lane_begin[b2] < sk.lane_end[b2] &&
                                sk.lane_end[a] + kGap > sk.lane_begin[b2])
                                fits = false;
                        }
                    if (fits) break;
                }
                return std::min(m, sk.KP);
            };
            M = (4 * sk.KP + 64 <= 1024) ? sk.KP : smallest_fitting(32);
            phases = (sk.KP + M - 1) / M;
            int m0 = 32;
            if (const char* e = getenv("GG_SPIRAL_M")) m0 = atoi(e) / 32 * 32;
            M_thr = m0 <= 0 ? M : smallest_fitting(std::max(32, m0));
            phases_thr = (sk.KP + M_thr - 1) / M_thr;
            // one CTA: lane threads + the two irregular warps (one thread per (visit, neighbour))
            if (4 * M + 64 > 1024 || 4 * M_thr + 64 > 1024 || sk.max_irr_per_level * 9 > 64) sk.ok = false;
        }
        if (sk.ok) {
            // re-layout of the irregular records: one dense block per level (see SkewView)
            const int irr_max = std::max(1, sk.max_irr_per_level);
            const int irr_words = ((irr_max * 22 + 3) / 4) * 4;
            std::vector<uint32_t> blocks((size_t)(sk.levels + 4) * irr_words, 0xffffffffu);
            for (int l = 0; l < sk.levels; ++l)
                for (int r = sk.irr_level_start[l]; r < sk.irr_level_start[l + 1]; ++r) {
                    const uint32_t* w = &sk.irr_recs[(size_t)r * 16];
                    const int vv = r - sk.irr_level_start[l];
                    uint32_t* blk = &blocks[(size_t)l * irr_words];
                    for (int q = 0; q < 9; ++q) {
                        blk[(vv * 9 + q) * 2] = w[1 + q];
                        blk[(vv * 9 + q) * 2 + 1] = 0xffffffffu;
                    }
                    const uint32_t ents[4] = {w[10] & 0xffffu, w[10] >> 16, w[11] & 0xffffu, w[11] >> 16};
                    for (uint32_t e : ents)
                        if (e != 0xffffu) blk[(vv * 9 + (e >> 12)) * 2 + 1] = e & 4095u;
                    uint32_t* hd = blk + irr_max * 18 + vv * 4;
                    hd[0] = w[0];
                    hd[1] = w[12];
                    hd[2] = w[13];
                    hd[3] = w[14];
                }
            // [phase][side * M + m] tables of a layout
            auto upload_phases = [&](int m_, int phases_, const int** d_b, const int** d_e, const int** d_c) -> int {
                std::vector<int> ph_b((size_t)phases_ * 4 * m_, 0), ph_e((size_t)phases_ * 4 * m_, 0), ph_c((size_t)phases_ * 4 * m_, 0);
                for (int ph = 0; ph < phases_; ++ph)
                    for (int sd = 0; sd < 4; ++sd)
                        for (int m = 0; m < m_; ++m) {
                            const int col = ph * m_ + m;
                            if (col >= sk.KP) continue;
                            const size_t dst = ((size_t)ph * 4 + sd) * m_ + m;
                            ph_b[dst] = sk.lane_begin[sd * sk.KP + col];
                            ph_e[dst] = sk.lane_end[sd * sk.KP + col];
                            ph_c[dst] = sk.lane_cell0[sd * sk.KP + col];
                        }
                int *d_lb = nullptr, *d_le = nullptr, *d_lc = nullptr;
                int rc2;
                if ((rc2 = dev_alloc(h, &d_lb, ph_b.size())) || (rc2 = dev_alloc(h, &d_le, ph_e.size())) || (rc2 = dev_alloc(h, &d_lc, ph_c.size())))
                    return rc2;
                if (cudaMemcpy(d_lb, ph_b.data(), ph_b.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(d_le, ph_e.data(), ph_e.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(d_lc, ph_c.data(), ph_c.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
                    return fail(GG_E_CUDA, "upload of the spiral phase tables failed");
                *d_b = d_lb;
                *d_e = d_le;
                *d_c = d_lc;
                return GG_OK;
            };
            GG_TRY(upload_phases(M, phases, &v.skew.ph_begin, &v.skew.ph_end, &v.skew.ph_cell0));
            GG_TRY(upload_phases(M_thr, phases_thr, &v.skew.thr_ph_begin, &v.skew.thr_ph_end, &v.skew.thr_ph_cell0));
            int* d_home = nullptr;
            uint32_t* d_irr = nullptr;
            GG_TRY(dev_alloc(h, &d_home, sk.cell_home.size()));
            GG_TRY(dev_alloc(h, &d_irr, blocks.size() + 16));
            GG_CUDA_TRY(cudaMemcpy(d_home, sk.cell_home.data(), sk.cell_home.size() * sizeof(int), cudaMemcpyHostToDevice));
            GG_CUDA_TRY(cudaMemcpy(d_irr, blocks.data(), blocks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            float2* d_sk = nullptr;
            float* d_sd = nullptr;
            GG_TRY(dev_alloc(h, &d_sk, S * sk.slots));
            GG_TRY(dev_alloc(h, &d_sd, S * sk.slots));
            GG_CUDA_TRY(cudaMemset(d_sk, 0, S * sk.slots * sizeof(float2)));
            GG_CUDA_TRY(cudaMemset(d_sd, 0, S * sk.slots * sizeof(float)));
            v.skew.sk = d_sk;
            v.skew.sd = d_sd;
            v.skew.slots = sk.slots;
            v.skew.cell_home = d_home;
            v.skew.M = M;
            v.skew.phases = phases;
            v.skew.thr_M = M_thr;
            v.skew.thr_phases = phases_thr;
            // barrier-free synchronisation of the spiral kernel: opt-in (GG_SPIRAL_ASYNC=1).  Measured on B200 it is SLOWER than
            // the CTA barrier per level (1 780 vs 1 484 us per 444 scans, 0.745 vs 0.688 ms for one scan): the release fence of
            // every publish also waits for the warp's loads of the next level, which the barrier lets stay in flight.
            int want_async = 0;
            if (const char* e = getenv("GG_SPIRAL_ASYNC")) want_async = atoi(e);
            auto upload_req = [&](int m_, const uint16_t** d_req) -> int {
                *d_req = nullptr;
                std::vector<uint16_t> rq;
                int n_agents = 0;
                if (!want_async || !gg::build_skew_sync(sk, m_, gg::SKEW_XCH_ASYNC, rq, n_agents)) return GG_OK;
                uint16_t* d = nullptr;
                int rc2;
                if ((rc2 = dev_alloc(h, &d, rq.size()))) return rc2;
                if (cudaMemcpy(d, rq.data(), rq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess)
                    return fail(GG_E_CUDA, "upload of the spiral synchronisation table failed");
                *d_req = d;
                return GG_OK;
            };
            v.skew.sync_sleep = 20;
            if (const char* e = getenv("GG_SPIRAL_SLEEP")) v.skew.sync_sleep = std::max(0, atoi(e));
            GG_TRY(upload_req(M, &v.skew.req));
            if (M_thr == M)
                v.skew.thr_req = v.skew.req;
            else
                GG_TRY(upload_req(M_thr, &v.skew.thr_req));
            v.skew.irr_blocks = reinterpret_cast<const uint4*>(d_irr);
            v.skew.irr_max = irr_max;
            v.skew.irr_chunks = irr_words / 4;
            v.skew.KP = sk.KP;
            v.skew.rows = sk.rows;
            v.skew.row0 = sk.row0;
            v.skew.lanes = sk.lanes;
            v.skew.levels = sk.levels;
            std::memcpy(v.skew.pattern, sk.pattern, sizeof(sk.pattern));
        }
    }

    if (const char* e = getenv("GG_LAUNCH_UNIT")) h->launch_unit = std::max(1, atoi(e));
    if (const char* e = getenv("GG_COPY_STREAMS")) h->n_copy_in = std::min(8, std::max(1, atoi(e)));
    if (const char* e = getenv("GG_MAIN_HELP")) h->main_help = atoi(e);
    if (const char* e = getenv("GG_RAW_GATE")) h->raw_gate = std::max(1, atoi(e));
    if (const char* e = getenv("GG_BUS_TARGET_MB")) h->bus_target = (size_t)std::max(0, atoi(e)) << 20;
    if (const char* e = getenv("GG_RAW_DEPTH")) h->raw_depth = std::min(8, std::max(1, atoi(e)));
    if (const char* e = getenv("GG_HOST_PACK")) {
        h->host_pack = atoi(e) ? 1 : 0;
        h->host_pack_mix = 0;
    }
    v.packed = nullptr;
    if (stream) {
        h->own_streams = false;
        h->n_streams = 1;
        h->streams[0] = static_cast<cudaStream_t>(stream);
    } else {
        int want = 8;
        if (const char* e = getenv("GG_STREAMS")) want = atoi(e);
        if (const char* e = getenv("GG_STAGGER")) h->stagger = atoi(e);
        h->n_streams = std::max(1, std::min(std::min(want, kStreams), n_slots));
        for (int i = 0; i < h->n_streams; ++i) GG_CUDA_TRY(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
    }
    GG_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h->h_ring), sizeof(gg::SlotParams) * kRing * S, cudaHostAllocDefault));
    GG_TRY(dev_alloc(h, &h->d_ring, (size_t)kRing * S));
    for (int i = 0; i < kRing; ++i) GG_CUDA_TRY(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < kStreams; ++i) GG_CUDA_TRY(cudaEventCreateWithFlags(&h->stagger_ev[i], cudaEventDisableTiming));
    h->launches += gg::launch_build_detect_table(v, h->detect_tab, h->streams[0]);
    GG_CUDA_TRY(cudaGetLastError());
    GG_CUDA_TRY(cudaStreamSynchronize(h->streams[0]));
#undef GG_TRY
#undef GG_CUDA_TRY
    *out = h;
    return GG_OK;
}

int gg_destroy(gg_handle h) {
    if (!h) return GG_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    delete h->prof;
    delete h->packer;
    for (int g = 0; g < kStreams; ++g)
        if (h->d_raw[g]) cudaFree(h->d_raw[g]);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    for (cudaEvent_t e : h->slot_ev) cudaEventDestroy(e);
    for (int i = 0; i < kStreams; ++i)
        if (h->stagger_ev[i]) cudaEventDestroy(h->stagger_ev[i]);
    for (int e = 0; e < 2; ++e)
        if (h->batch_done[e]) cudaEventDestroy(h->batch_done[e]);
    for (int e = 0; e < 8; ++e)
        if (h->raw_ev[e]) cudaEventDestroy(h->raw_ev[e]);
    for (cudaEvent_t e : h->batch_ev) cudaEventDestroy(e);
    for (int e = 0; e < 10; ++e)
        if (h->copy_in[e]) cudaStreamDestroy(h->copy_in[e]);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    for (void* p : h->dev_allocs) cudaFree(p);
    if (h->h_ring) cudaFreeHost(h->h_ring);
    for (int i = 0; i < kRing; ++i)
        if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    if (h->own_streams)
        for (int i = 0; i < h->n_streams; ++i)
            if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    delete h;
    return GG_OK;
}

int gg_cells_per_side(gg_handle h) { return h ? h->view.k.N : GG_E_ARG; }
int gg_num_slots(gg_handle h) { return h ? h->n_slots : GG_E_ARG; }
void* gg_stream(gg_handle h) { return h ? h->streams[0] : nullptr; }
uint64_t gg_kernel_launches(gg_handle h) { return h ? h->launches : 0; }

int gg_spiral_schedule_info(gg_handle h, int* levels, int* visits, int* max_per_level) {
    if (!h) return fail(GG_E_ARG, "null handle");
    if (levels) *levels = h->sched_levels;
    if (visits) *visits = h->sched_visits;
    if (max_per_level) *max_per_level = h->sched_max;
    return GG_OK;
}

int gg_set_config(gg_handle h, const gg_config* cfg) {
    if (!h || !cfg) return fail(GG_E_ARG, "null argument");
    GG_CUDA(cudaSetDevice(h->device));
    int rc = gg_synchronize(h);  // kernels in flight keep the tables of the old configuration
    if (rc) return rc;
    h->cfg = *cfg;
    gg::derive_constants(h->cfg, h->dimension_m, h->resolution, h->flags, h->view.k);  // by-value kernel argument: applies to the next launch
    h->launches += gg::launch_build_detect_table(h->view, h->detect_tab, h->streams[0]);
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaStreamSynchronize(h->streams[0]));
    return GG_OK;
}

int gg_get_config(gg_handle h, gg_config* cfg) {
    if (!h || !cfg) return fail(GG_E_ARG, "null argument");
    *cfg = h->cfg;
    return GG_OK;
}

int gg_init_map(gg_handle h, int slot, double x, double y, double z) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    GG_CUDA(cudaSetDevice(h->device));
    SlotState& s = h->slots[slot];
    s = SlotState();
    s.px = x;
    s.py = y;
    s.have_map = true;
    h->launches += gg::launch_init_map(h->view, slot, (float)z, stream_of(h, slot));
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}

int gg_update_pose_batch(gg_handle h, int count, const int* slots, const double* xy, const double* T, int* moved) {
    if (!h || !slots || !xy || !T) return fail(GG_E_ARG, "null argument");
    if (count <= 0) return GG_OK;
    if (count > h->n_slots) return fail(GG_E_ARG, "count exceeds slots");
    GG_CUDA(cudaSetDevice(h->device));
    int rc;
    std::vector<unsigned char>& seen = h->seen_scratch;
    seen.assign((size_t)h->n_slots, 0);
    for (int i = 0; i < count; ++i) {
        if ((rc = check_slot(h, slots[i]))) return rc;
        if (seen[slots[i]]++) return fail(GG_E_ARG, "slot %d appears twice in one batch", slots[i]);
        if (!h->slots[slots[i]].have_map) return fail(GG_E_STATE, "slot %d: map not initialised", slots[i]);
    }
    for (int g = 0; g < h->n_streams; ++g) {
        gg::SlotParams *hp = nullptr, *dp = nullptr;
        int pos = 0, m = 0, any = 0;
        for (int i = 0; i < count; ++i) {
            if (stream_index(h, slots[i]) != g) continue;
            if (m == 0 && (rc = ring_acquire(h, &hp, &dp, &pos))) return rc;
            SlotState& s = h->slots[slots[i]];
            gg::SlotParams& p = hp[m++];
            std::memset(&p, 0, sizeof(p));
            gg::move_map(h->view.k.res, s.px, s.py, xy[2 * i], xy[2 * i + 1], p.shift_i, p.shift_j);
            p.px = s.px;
            p.py = s.py;
            const double* t = T + 12 * (size_t)i;
            p.t20 = t[8];
            p.t21 = t[9];
            p.t22 = t[10];
            p.t23 = t[11];
            p.slot = slots[i];
            const int mv = (p.shift_i != 0 || p.shift_j != 0) ? 1 : 0;
            if (moved) moved[i] = mv;
            any |= mv;
        }
        if (m == 0) continue;
        cudaStream_t st = h->streams[g];
        if (any) {  // else: "We havent moved so we have nothing to do", GroundGrid.cpp:136-137
            if ((rc = ring_commit(h, pos, m, st))) return rc;
            h->launches += gg::launch_roll(h->view, dp, m, st, h->prof);
            GG_CUDA(cudaGetLastError());
        }
        if ((rc = ring_release(h, pos, st))) return rc;
    }
    return GG_OK;
}

int gg_update_pose(gg_handle h, int slot, double x, double y, const double T[12], int* moved) {
    const double xy[2] = {x, y};
    return gg_update_pose_batch(h, 1, &slot, xy, T, moved);
}

int gg_get_map_position(gg_handle h, int slot, double xy[2]) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!xy) return fail(GG_E_ARG, "null argument");
    xy[0] = h->slots[slot].px;
    xy[1] = h->slots[slot].py;
    return GG_OK;
}

int gg_set_map_position(gg_handle h, int slot, double x, double y) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    h->slots[slot].px = x;
    h->slots[slot].py = y;
    return GG_OK;
}

int gg_upload_points(gg_handle h, int slot, const gg_point* points, size_t n) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (n > h->pcap) return fail(GG_E_ARG, "%zu points exceed capacity %zu", n, h->pcap);
    if (n && !points) return fail(GG_E_ARG, "null points");
    h->inputs_busy = true;
    GG_CUDA(cudaSetDevice(h->device));
    if (n) GG_CUDA(cudaMemcpyAsync(h->view.points + (size_t)slot * h->pcap, points, n * sizeof(gg_point), cudaMemcpyHostToDevice, stream_of(h, slot)));
    h->slots[slot].n_points = n;
    return GG_OK;
}

int gg_run_scans(gg_handle h, int count, const gg_scan_desc* scans, int stop_after) {
    if (!h || !scans) return fail(GG_E_ARG, "null argument");
    if (stop_after < 0 || stop_after > 3) return fail(GG_E_ARG, "stop_after must be 0..3");
    h->inputs_busy = true;
    GG_CUDA(cudaSetDevice(h->device));
    return run_scans_grouped(h, count, scans, stop_after);
}

int gg_run_scans_device(gg_handle h, int count, const gg_scan_desc* scans, const gg_point* const* dev_points, int stop_after) {
    if (!h || !scans || !dev_points) return fail(GG_E_ARG, "null argument");
    if (stop_after < 0 || stop_after > 3) return fail(GG_E_ARG, "stop_after must be 0..3");
    h->inputs_busy = true;
    for (int i = 0; i < count; ++i)
        if (!dev_points[i] && scans[i].n_points) return fail(GG_E_ARG, "scan %d: null device cloud", i);
    GG_CUDA(cudaSetDevice(h->device));
    return run_scans_grouped(h, count, scans, stop_after, dev_points);
}

// ---- "next" rows of SURVEY.md section 8(f) --------------------------------------------------
int gg_upload_cloud_msg(gg_handle h, int slot, const void* data, size_t n_points, int point_step, const int field_offsets[5],
                        const double T_map_from_frame[12]) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (n_points > h->pcap) return fail(GG_E_ARG, "%zu points exceed capacity %zu", n_points, h->pcap);
    if ((n_points && !data) || !field_offsets || point_step < 12) return fail(GG_E_ARG, "bad PointCloud2 layout");
    for (int f = 0; f < 5; ++f) {
        const int width = f == 4 ? 2 : 4;
        if ((f < 3 && field_offsets[f] < 0) || field_offsets[f] + width > point_step) return fail(GG_E_ARG, "field %d does not fit point_step", f);
    }
    GG_CUDA(cudaSetDevice(h->device));
    h->inputs_busy = true;
    cudaStream_t st = stream_of(h, slot);
    const size_t bytes = n_points * (size_t)point_step;
    const int sg = stream_index(h, slot);
    if (bytes > h->d_raw_cap[sg]) {
        GG_CUDA(cudaStreamSynchronize(st));   // the only users of this buffer are on `st`
        if (h->d_raw[sg]) GG_CUDA(cudaFree(h->d_raw[sg]));
        h->d_raw[sg] = nullptr;
        h->d_raw_cap[sg] = 0;
        GG_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_raw[sg]), bytes + 256));
        h->d_raw_cap[sg] = bytes;
    }
    if (bytes) GG_CUDA(cudaMemcpyAsync(h->d_raw[sg], data, bytes, cudaMemcpyHostToDevice, st));
    gg::UnpackDesc d;
    std::memset(&d, 0, sizeof(d));
    d.raw = h->d_raw[sg];
    d.dst = h->view.points + (size_t)slot * h->pcap;
    d.n = (int)n_points;
    d.point_step = point_step;
    for (int f = 0; f < 5; ++f) d.off[f] = field_offsets[f];
    d.transform = T_map_from_frame ? 1 : 0;
    if (T_map_from_frame) std::memcpy(d.T, T_map_from_frame, sizeof(d.T));
    h->launches += gg::launch_unpack(d, st, h->prof);
    GG_CUDA(cudaGetLastError());
    h->slots[slot].n_points = n_points;
    return GG_OK;
}

int gg_terrain_image(gg_handle h, int slot, float* dst) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!dst) return fail(GG_E_ARG, "null dst");
    if (!(h->flags & GG_FLAG_FULL_LAYERS)) return fail(GG_E_LAYER, "the terrain image needs 'pointsRaw' (GG_FLAG_FULL_LAYERS)");
    GG_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->view.k.N2 * 3;
    if (!h->d_image && (rc = dev_alloc(h, &h->d_image, n))) return rc;
    cudaStream_t st = stream_of(h, slot);
    h->launches += gg::launch_terrain_image(h->view, slot, h->d_image, st, h->prof);
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaMemcpyAsync(dst, h->d_image, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    return GG_OK;
}

// f3: the single-channel 8-bit image cv::applyColorMap receives (GroundGridNodelet.cpp:238-245)
int gg_layer_image_u8(gg_handle h, int slot, const char* name, uint8_t* dst, float* lower, float* upper) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!name || !dst) return fail(GG_E_ARG, "null argument");
    int l = 0;
    if ((rc = layer_index(h, slot, name, &l))) return rc;
    GG_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->view.k.N2;
    if (!h->d_image_u8) {
        if ((rc = dev_alloc(h, &h->d_image_u8, n))) return rc;
        if ((rc = dev_alloc(h, &h->d_minmax, 2))) return rc;
    }
    cudaStream_t st = stream_of(h, slot);
    const int init[2] = {0x7f800000, (int)(0xff800000u ^ 0x7fffffffu)};   // ordered keys of +inf / -inf
    GG_CUDA(cudaMemcpyAsync(h->d_minmax, init, sizeof(init), cudaMemcpyHostToDevice, st));
    h->launches += gg::launch_layer_image_u8(h->view, h->view.layer(slot, l), h->d_minmax, h->d_image_u8, st);
    GG_CUDA(cudaGetLastError());
    int keys[2];
    GG_CUDA(cudaMemcpyAsync(dst, h->d_image_u8, n, cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaMemcpyAsync(keys, h->d_minmax, sizeof(keys), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    auto unkey = [](int k) { const int b = k >= 0 ? k : (k ^ 0x7fffffff); float f; std::memcpy(&f, &b, 4); return f; };
    if (lower) *lower = unkey(keys[0]);
    if (upper) *upper = unkey(keys[1]);
    return GG_OK;
}

// ---- single phases (GroundSegmentation.h:56-62 of the reference) ------------------------------
namespace {
int one_slot_params(gg_handle h, int slot, double base_z, gg::SlotParams** dp_out, int* pos_out, cudaStream_t st) {
    gg::SlotParams *hp = nullptr, *dp = nullptr;
    int rc, pos = 0;
    if ((rc = ring_acquire(h, &hp, &dp, &pos))) return rc;
    gg_scan_desc d;
    std::memset(&d, 0, sizeof(d));
    d.slot = slot;
    d.n_points = h->slots[slot].n_points;
    d.base_z = base_z;
    fill_params(h, d, hp[0], h->slots[slot].src);
    if ((rc = ring_commit(h, pos, 1, st))) return rc;
    *dp_out = dp;
    *pos_out = pos;
    return GG_OK;
}
}  // namespace

int gg_detect_ground_patches(gg_handle h, int slot) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!h->slots[slot].have_map) return fail(GG_E_STATE, "slot %d: map not initialised", slot);
    GG_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream_of(h, slot);
    gg::SlotParams* dp = nullptr;
    int pos = 0;
    if ((rc = one_slot_params(h, slot, 0.0, &dp, &pos, st))) return rc;
    h->launches += gg::launch_detect_only(h->view, dp, 1, st, h->prof, h->have_layer_map ? &h->layer_map : nullptr);
    GG_CUDA(cudaGetLastError());
    return ring_release(h, pos, st);
}

int gg_spiral_ground_interpolation(gg_handle h, int slot, double base_z) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!h->slots[slot].have_map) return fail(GG_E_STATE, "slot %d: map not initialised", slot);
    GG_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream_of(h, slot);
    gg::SlotParams* dp = nullptr;
    int pos = 0;
    if ((rc = one_slot_params(h, slot, base_z, &dp, &pos, st))) return rc;
    h->launches += gg::launch_spiral_only(h->view, dp, 1, st, h->prof);
    GG_CUDA(cudaGetLastError());
    return ring_release(h, pos, st);
}

int gg_interpolate_cell(gg_handle h, int slot, int x, int y) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    const int N = h->view.k.N;
    if (x < 1 || y < 1 || x >= N - 1 || y >= N - 1) return fail(GG_E_ARG, "cell (%d, %d) has no 3x3 neighbourhood", x, y);
    GG_CUDA(cudaSetDevice(h->device));
    h->launches += gg::launch_interpolate_cell(h->view, slot, x, y, stream_of(h, slot));
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}

int gg_detect_ground_patch(gg_handle h, int slot, int patch_size, int i, int j) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (patch_size != 3 && patch_size != 5) return fail(GG_E_ARG, "patch size must be 3 or 5");
    const int N = h->view.k.N, H = patch_size / 2;
    if (i < H || j < H || i >= N - H || j >= N - H) return fail(GG_E_ARG, "cell (%d, %d) has no %dx%d neighbourhood", i, j, patch_size, patch_size);
    GG_CUDA(cudaSetDevice(h->device));
    h->launches += gg::launch_detect_cell(h->view, slot, patch_size, i, j, stream_of(h, slot));
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}

// class << 24 | cell of every input point of the slot's last rasterisation (the index lists insert_cloud fills)
int gg_get_point_classes(gg_handle h, int slot, uint32_t* codes, size_t n) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (n > h->slots[slot].n_points || (n && !codes)) return fail(GG_E_ARG, "bad class buffer");
    if (!h->slots[slot].ran) return fail(GG_E_STATE, "slot %d: no rasterised scan", slot);
    GG_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream_of(h, slot);
    if (n) GG_CUDA(cudaMemcpyAsync(codes, h->view.code + (size_t)slot * h->pcap, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    return GG_OK;
}

int gg_eval_accumulate(gg_handle h, int slot) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    SlotState& s = h->slots[slot];
    if (!s.ran || s.last_stop != 0) return fail(GG_E_STATE, "slot %d: no completed scan", slot);
    GG_CUDA(cudaSetDevice(h->device));
    if (!h->d_eval) {
        if ((rc = dev_alloc(h, &h->d_eval, (size_t)gg::EVAL_LABELS * 2))) return rc;
        GG_CUDA(cudaMemset(h->d_eval, 0, sizeof(unsigned long long) * gg::EVAL_LABELS * 2));
    }
    cudaStream_t st = stream_of(h, slot);
    gg::SlotParams *hp = nullptr, *dp = nullptr;
    int pos = 0;
    if ((rc = ring_acquire(h, &hp, &dp, &pos))) return rc;
    std::memset(&hp[0], 0, sizeof(gg::SlotParams));
    hp[0].slot = slot;
    hp[0].n_points = (int)s.n_points;
    hp[0].src = s.src ? s.src : h->view.points + (size_t)slot * h->pcap;
    hp[0].packed = s.packed_input;
    if ((rc = ring_commit(h, pos, 1, st))) return rc;
    h->launches += gg::launch_eval(h->view, dp, h->d_eval, st, h->prof);
    GG_CUDA(cudaGetLastError());
    return ring_release(h, pos, st);
}

int gg_eval_read(gg_handle h, uint64_t* counts, int reset) {
    if (!h || !counts) return fail(GG_E_ARG, "null argument");
    int rc = gg_synchronize(h);
    if (rc) return rc;
    const size_t bytes = sizeof(unsigned long long) * gg::EVAL_LABELS * 2;
    if (!h->d_eval) {
        std::memset(counts, 0, bytes);
        return GG_OK;
    }
    GG_CUDA(cudaMemcpy(counts, h->d_eval, bytes, cudaMemcpyDeviceToHost));
    if (reset) GG_CUDA(cudaMemset(h->d_eval, 0, bytes));
    return GG_OK;
}

int gg_profile_enable(gg_handle h, int on) {
    if (!h) return fail(GG_E_ARG, "null handle");
    GG_CUDA(cudaSetDevice(h->device));
    if (on && !h->prof) h->prof = new EventProfiler();
    if (!on && h->prof) {
        int rc = gg_synchronize(h);
        if (rc) return rc;
        h->prof->collect(h->prof_ms, h->prof_count);
        delete h->prof;
        h->prof = nullptr;
    }
    return GG_OK;
}

int gg_profile_read(gg_handle h, double* ms_per_kernel, uint32_t* launches_per_kernel, int reset) {
    if (!h) return fail(GG_E_ARG, "null handle");
    int rc = gg_synchronize(h);
    if (rc) return rc;
    if (h->prof) h->prof->collect(h->prof_ms, h->prof_count);
    for (int k = 0; k < gg::K_NUM; ++k) {
        if (ms_per_kernel) ms_per_kernel[k] = h->prof_ms[k];
        if (launches_per_kernel) launches_per_kernel[k] = h->prof_count[k];
        if (reset) {
            h->prof_ms[k] = 0.0;
            h->prof_count[k] = 0;
        }
    }
    return GG_OK;
}

int gg_profile_kernel_count(void) { return gg::K_NUM; }

const char* gg_profile_kernel_name(int id) {
    static const char* names[gg::K_NUM] = {"k_rasterize",   "k_cell_tiles",    "k_cell_place",    "k_scatter",
                                           "k_cell_stats",  "k_detect",        "k_spiral",           "k_label",         "k_roll_gather",
                                           "k_roll_commit", "k_out_count",     "k_out_scan",         "k_out_write",     "k_unpack_transform",
                                           "k_terrain_image", "k_eval_counts"};
    return (id >= 0 && id < gg::K_NUM) ? names[id] : "";
}

int gg_download_labels(gg_handle h, int slot, uint8_t* labels_out, size_t n) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (n > h->pcap || (n && !labels_out)) return fail(GG_E_ARG, "bad label buffer");
    GG_CUDA(cudaSetDevice(h->device));
    if (n) GG_CUDA(cudaMemcpyAsync(labels_out, h->view.labels + (size_t)slot * h->pcap, n, cudaMemcpyDeviceToHost, stream_of(h, slot)));
    return GG_OK;
}

int gg_synchronize(gg_handle h) {
    if (!h) return fail(GG_E_ARG, "null handle");
    GG_CUDA(cudaSetDevice(h->device));
    for (int i = 0; i < h->n_streams; ++i) GG_CUDA(cudaStreamSynchronize(h->streams[i]));
    if (h->copy_out) GG_CUDA(cudaStreamSynchronize(h->copy_out));
    h->batch_outstanding[0] = h->batch_outstanding[1] = false;
    h->inputs_busy = false;
    return GG_OK;
}

int gg_get_output(gg_handle h, int slot, uint32_t* index_out, gg_point* cloud_out, size_t* n_out) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    GG_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream_of(h, slot);
    if ((rc = run_output_on(h, slot, cloud_out != nullptr, st))) return rc;
    int total = 0;
    const gg::View& v = h->view;
    GG_CUDA(cudaMemcpyAsync(&total, v.out_counts + (size_t)slot * (3 * v.out_blocks + 1) + 3 * v.out_blocks, sizeof(int), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    if (n_out) *n_out = (size_t)total;
    if (index_out && total) GG_CUDA(cudaMemcpyAsync(index_out, v.out_index + (size_t)slot * h->pcap, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (cloud_out && total) GG_CUDA(cudaMemcpyAsync(cloud_out, v.out_cloud + (size_t)slot * h->pcap, (size_t)total * sizeof(gg_point), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    return GG_OK;
}

int gg_filter_cloud(gg_handle h, int slot, const gg_point* points, size_t n, const float origin[3], double base_z,
                    uint8_t* labels_out, uint32_t* index_out, gg_point* cloud_out, size_t* n_out) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!origin) return fail(GG_E_ARG, "null origin");
    if (!h->slots[slot].have_map) return fail(GG_E_STATE, "slot %d: map not initialised", slot);
    if ((rc = gg_upload_points(h, slot, points, n))) return rc;
    gg_scan_desc d;
    std::memset(&d, 0, sizeof(d));
    d.slot = slot;
    d.n_points = n;
    d.origin[0] = origin[0];
    d.origin[1] = origin[1];
    d.origin[2] = origin[2];
    d.base_z = base_z;
    if ((rc = run_scans_grouped(h, 1, &d, 0))) return rc;
    if (labels_out && (rc = gg_download_labels(h, slot, labels_out, n))) return rc;
    if (index_out || cloud_out || n_out) return gg_get_output(h, slot, index_out, cloud_out, n_out);
    GG_CUDA(cudaStreamSynchronize(stream_of(h, slot)));
    return GG_OK;
}

namespace {

int batch_wait(gg_handle h, int parity) {
    if (h->batch_outstanding[parity]) {
        GG_CUDA(cudaEventSynchronize(h->batch_done[parity]));
        h->batch_outstanding[parity] = false;
    }
    return GG_OK;
}

// streams, events and the second buffer set of the batch path
int batch_prepare(gg_handle h) {
    int rc;
    if (!h->copy_out) {
        for (int e = 0; e < 8; ++e) GG_CUDA(cudaEventCreateWithFlags(&h->raw_ev[e], cudaEventDisableTiming));
        for (int e = 0; e < 2; ++e) GG_CUDA(cudaEventCreateWithFlags(&h->batch_done[e], cudaEventDisableTiming));
        for (int e = 0; e < h->n_copy_in + 2; ++e) GG_CUDA(cudaStreamCreateWithFlags(&h->copy_in[e], cudaStreamNonBlocking));
        GG_CUDA(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
        h->in_raw[0] = h->view.points;
        h->labels_buf[0] = h->view.labels;
    }
    if (h->host_pack && !h->h_stage) {
        if ((rc = dev_alloc(h, &h->in_raw[1], (size_t)h->n_slots * h->pcap))) return rc;
        if ((rc = dev_alloc(h, &h->labels_buf[1], (size_t)h->n_slots * h->pcap))) return rc;
        for (int e = 0; e < 2; ++e)
            if ((rc = dev_alloc(h, &h->in_packed[e], (size_t)h->n_slots * 14 * h->pcap))) return rc;
        if (const char* e = getenv("GG_PACK_RING")) h->pack_slots = std::min(256, std::max(2, atoi(e)));
        GG_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->h_stage), (size_t)h->pack_slots * 14 * h->pcap, cudaHostAllocDefault));
        h->slot_ev.resize(h->pack_slots);
        for (cudaEvent_t& e : h->slot_ev) GG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->packer->set_ring(h->h_stage, 14 * h->pcap, h->pack_slots);
    }
    return GG_OK;
}

}  // namespace

int gg_filter_cloud_batch_begin(gg_handle h, int count, const gg_scan_desc* scans, const gg_point* const* points, uint8_t* const* labels_out,
                                int* ticket) {
    if (!h || !scans || !points) return fail(GG_E_ARG, "null argument");
    if (count < 0) return fail(GG_E_ARG, "negative count");
    if (count > h->n_slots) return fail(GG_E_ARG, "count exceeds slots");
    GG_CUDA(cudaSetDevice(h->device));
    int rc;
    std::vector<unsigned char> seen((size_t)h->n_slots, 0);   // (run_scans_grouped re-checks its own sub-batches with the handle's scratch)
    for (int i = 0; i < count; ++i) {
        const gg_scan_desc& d = scans[i];
        if ((rc = check_slot(h, d.slot))) return rc;
        if (seen[d.slot]++) return fail(GG_E_ARG, "slot %d appears twice in one batch (scans of a batch run concurrently)", d.slot);
        if (d.n_points > h->pcap) return fail(GG_E_ARG, "slot %d: too many points", d.slot);
        if (d.n_points && !points[i]) return fail(GG_E_ARG, "scan %d: null cloud", i);
        if (!h->slots[d.slot].have_map) return fail(GG_E_STATE, "slot %d: map not initialised", d.slot);
    }
    if (h->host_pack && !h->packer) {
        // GG_HOST_THREADS forces a thread count; by default all usable CPUs (affinity, cgroup quota,
        // shared between the local ranks) but one
        int threads = 0;
        if (const char* e = getenv("GG_HOST_THREADS")) threads = atoi(e);
        if (threads <= 0) {
            int local = 1;
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(e));
            threads = std::min(48, gg::usable_cpus() / local - 1);
        }
        if (threads < 1)
            h->host_pack = 0;
        else
            h->packer = new HostPacker(threads);
    }
    if ((rc = batch_prepare(h))) return rc;
    while ((int)h->batch_ev.size() < h->n_streams) {  // [0, n_streams): one event per compute stream
        cudaEvent_t ev;
        GG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->batch_ev.push_back(ev);
    }
    const int par = h->host_pack ? h->batch_parity : 0;
    // the batch that used this buffer set two calls ago must be complete (it normally is, long since)
    if ((rc = batch_wait(h, par))) return rc;
    h->last_raw = h->last_packed = h->last_raw_bytes = h->last_packed_bytes = 0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto us_since = [&] { return (size_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_begin).count(); };
    if (h->host_pack) {
        // Two ways to get a cloud across PCIe: repacked by host worker threads into x | y | z | ring
        // (14 useful bytes of every 32-byte record, costs CPU time) or as it is (costs bus time).  Both
        // resources are used at once: the packers walk the scans from the front; whenever fewer than
        // raw_depth raw copies of this loop are pending, the calling thread takes the LAST scan nobody has
        // started packing and sends it raw.  Whatever the CPU quota of the host, neither the packers
        // nor the bus sit idle.
        //
        // All clouds go through a few copy streams and all labels come back through another one, so a
        // transfer never queues behind kernels.  Events hand the scans over: copies -> kernels (the
        // stream of the slot) -> label read-back.  Kernels are enqueued for `launch_unit` delivered scans
        // of a stream group at a time, so that only the kernels of the last few scans run after the
        // transfers have ended.
        unsigned char* const dpk = h->in_packed[par];
        gg_point* const draw = h->in_raw[par];
        uint8_t* const dlab = h->labels_buf[par];
        h->view.labels = dlab;  // what gg_download_labels / gg_get_output read after this batch
        const int KP = h->n_copy_in, KC = KP + 2;  // packed clouds rotate over the first KP copy streams, raw ones over the last 2
        std::vector<int> order;
        for (int g = 0; g < h->n_streams; ++g)
            for (int i = 0; i < count; ++i)
                if (stream_index(h, scans[i].slot) == g) order.push_back(i);
        std::vector<PackJob> jobs(count);
        for (int k = 0; k < count; ++k) {
            const gg_scan_desc& d = scans[order[k]];
            jobs[k].src = points[order[k]];
            jobs[k].n = d.n_points;
        }
        if (h->inputs_busy) {  // earlier asynchronous calls may still use the slots' own input buffers
            for (int g = 0; g < h->n_streams; ++g) {
                GG_CUDA(cudaEventRecord(h->batch_ev[g], h->streams[g]));
                for (int c = 0; c < KC; ++c) GG_CUDA(cudaStreamWaitEvent(h->copy_in[c], h->batch_ev[g], 0));
            }
        }
        size_t ev_used = h->n_streams;
        auto next_event = [&](cudaEvent_t* ev) -> int {
            if (ev_used == h->batch_ev.size()) {
                cudaEvent_t e;
                GG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->batch_ev.push_back(e);
            }
            *ev = h->batch_ev[ev_used++];
            return GG_OK;
        };
        std::vector<std::vector<int>> pending(h->n_streams);  // delivered, kernels not yet enqueued (positions in `order`)
        std::vector<gg_scan_desc> ud;
        std::vector<const gg_point*> usrc;
        std::vector<const float*> upk;
        std::vector<unsigned char> sent_packed(count, 0);
        auto flush = [&](int g) -> int {
            std::vector<int>& pg = pending[g];
            if (pg.empty()) return GG_OK;
            ud.clear();
            usrc.clear();
            upk.clear();
            for (int k : pg) {
                const gg_scan_desc& d = scans[order[k]];
                ud.push_back(d);
                usrc.push_back(draw + (size_t)d.slot * h->pcap);
                upk.push_back(sent_packed[k] ? reinterpret_cast<const float*>(dpk + (size_t)d.slot * 14 * h->pcap) : nullptr);
            }
            cudaStream_t st = h->streams[g];
            cudaEvent_t ev;
            int rc2;
            for (int c = 0; c < KC; ++c) {  // the clouds are spread over all copy streams
                if ((rc2 = next_event(&ev))) return rc2;
                GG_CUDA(cudaEventRecord(ev, h->copy_in[c]));
                GG_CUDA(cudaStreamWaitEvent(st, ev, 0));
            }
            if ((rc2 = run_scans_grouped(h, (int)ud.size(), ud.data(), 0, usrc.data(), upk.data(), dlab))) return rc2;
            if (labels_out) {  // the labels travel under the remaining H2D traffic
                if ((rc2 = next_event(&ev))) return rc2;
                GG_CUDA(cudaEventRecord(ev, st));
                GG_CUDA(cudaStreamWaitEvent(h->copy_out, ev, 0));
                for (int k : pg)
                    if (labels_out[order[k]] && scans[order[k]].n_points)
                        GG_CUDA(cudaMemcpyAsync(labels_out[order[k]], dlab + (size_t)scans[order[k]].slot * h->pcap, scans[order[k]].n_points,
                                                cudaMemcpyDeviceToHost, h->copy_out));
            }
            pg.clear();
            return GG_OK;
        };
        auto delivered = [&](int k) -> int {
            const int g = stream_index(h, scans[order[k]].slot);
            pending[g].push_back(k);
            return (int)pending[g].size() >= h->launch_unit ? flush(g) : GG_OK;
        };
        const uint64_t seq_base = h->pack_issued;
        auto poll_copies = [&] {  // staging slots whose cloud has reached the device go back to the packers, in order
            uint64_t done = h->packer->copied.load(std::memory_order_relaxed);
            while (done < h->pack_issued && cudaEventQuery(h->slot_ev[done % (uint64_t)h->pack_slots]) == cudaSuccess) ++done;
            h->packer->copied.store(done, std::memory_order_release);
        };
        h->packer->pack_ns = 0;
        h->packer->slot_wait_ns = 0;
        uint64_t idle_ns = 0;
        h->packer->start(&jobs, seq_base);
        struct CancelOnExit {  // an error return below must not leave packers working on `jobs`
            HostPacker* p;
            bool armed = true;
            ~CancelOnExit() {
                if (armed) p->cancel();
            }
        } pack_guard{h->packer};
        int front = 0, back = count, raw_issued = 0, raw_done = 0, n_copies = 0;
        size_t last_packed_bytes = 14 * (size_t)h->pcap, last_raw_bytes = sizeof(gg_point) * (size_t)h->pcap;   // size of the latest copy of either kind
        while (front < back) {
            poll_copies();
            if (h->packer->packed(jobs[front])) {
                const gg_scan_desc& d = scans[order[front]];
                const size_t n_pad = (d.n_points + 7) & ~(size_t)7;
                cudaStream_t cs = h->copy_in[n_copies++ % KP];
                if (d.n_points)
                    GG_CUDA(cudaMemcpyAsync(dpk + (size_t)d.slot * 14 * h->pcap, h->packer->slot_of(h->pack_issued), 14 * n_pad, cudaMemcpyHostToDevice, cs));
                GG_CUDA(cudaEventRecord(h->slot_ev[h->pack_issued % (uint64_t)h->pack_slots], cs));
                ++h->pack_issued;
                sent_packed[front] = 1;
                ++h->last_packed;
                h->last_packed_bytes += 14 * n_pad;
                last_packed_bytes = 14 * n_pad;
                if ((rc = delivered(front))) return rc;
                ++front;
                continue;
            }
            // A raw cloud costs 32 B/point of bus time, a packed one 14, so packed clouds go first whenever one is ready.
            // What the packers cannot fill is topped up with raw clouds: the measure is the number of BYTES in flight on
            // the bus (copies issued and not yet completed, both kinds).  Below the target the DMA engines would run dry
            // before the next packed cloud arrives, so a raw one is added; above it the bus is the limit and a raw cloud
            // would only delay cheaper packed ones.  The split therefore follows the measured rates of the packers and of
            // the bus on this host (CPU quota, ranks sharing the socket) instead of a fixed gate.
            while (raw_done < raw_issued && cudaEventQuery(h->raw_ev[raw_done % h->raw_depth]) == cudaSuccess) ++raw_done;
            const uint64_t packed_in_flight = h->pack_issued - h->packer->copied.load(std::memory_order_relaxed);
            const size_t in_flight_bytes = (size_t)packed_in_flight * last_packed_bytes + (size_t)(raw_issued - raw_done) * last_raw_bytes;
            const bool below_target = h->bus_target ? in_flight_bytes < h->bus_target : packed_in_flight < (uint64_t)h->raw_gate;
            const bool bus_free = h->host_pack_mix && below_target && (raw_issued - raw_done) < h->raw_depth;
            int j = bus_free ? h->packer->claim_raw_from_back() : -1;
            if (j >= 0) {
                const gg_scan_desc& d = scans[order[j]];
                if (d.n_points)
                    GG_CUDA(cudaMemcpyAsync(draw + (size_t)d.slot * h->pcap, jobs[j].src, d.n_points * sizeof(gg_point), cudaMemcpyHostToDevice,
                                            h->copy_in[KP + (raw_issued & 1)]));
                GG_CUDA(cudaEventRecord(h->raw_ev[raw_issued % h->raw_depth], h->copy_in[KP + (raw_issued & 1)]));
                ++raw_issued;
                ++h->last_raw;
                h->last_raw_bytes += d.n_points * sizeof(gg_point);
                last_raw_bytes = d.n_points * sizeof(gg_point);
                back = j;
                if ((rc = delivered(j))) return rc;
                continue;
            }
            // nothing to enqueue: with the raw queue full (or no mixing) this thread packs a chunk as well
            const auto i0 = std::chrono::steady_clock::now();
            if ((h->host_pack_mix && (bus_free || !h->main_help)) || !h->packer->help()) std::this_thread::yield();
            idle_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - i0).count();
        }
        h->last_pack_us = h->packer->pack_ns.load() / 1000;
        h->last_slot_wait_us = h->packer->slot_wait_ns.load() / 1000;
        h->last_idle_us = idle_ns / 1000;
        pack_guard.armed = false;  // every job is packed or was sent raw
        for (int g = 0; g < h->n_streams; ++g)
            if ((rc = flush(g))) return rc;
        h->batch_parity ^= 1;
    } else {
        // H2D of every 32-byte cloud on its slot's stream, then the kernels of each stream group: everything is
        // stream-ordered, one buffer set is enough
        for (int i = 0; i < count; ++i) {
            const gg_scan_desc& d = scans[i];
            if (d.n_points)
                GG_CUDA(cudaMemcpyAsync(h->view.points + (size_t)d.slot * h->pcap, points[i], d.n_points * sizeof(gg_point), cudaMemcpyHostToDevice,
                                        stream_of(h, d.slot)));
        }
        h->last_raw = (size_t)count;
        for (int i = 0; i < count; ++i) h->last_raw_bytes += scans[i].n_points * sizeof(gg_point);
        if ((rc = run_scans_grouped(h, count, scans, 0))) return rc;
        if (labels_out)
            for (int i = 0; i < count; ++i)
                if (labels_out[i] && scans[i].n_points)
                    GG_CUDA(cudaMemcpyAsync(labels_out[i], h->view.labels + (size_t)scans[i].slot * h->pcap, scans[i].n_points, cudaMemcpyDeviceToHost,
                                            stream_of(h, scans[i].slot)));
    }
    // batch_done: everything enqueued so far on the compute streams and on the label stream
    for (int g = 0; g < h->n_streams; ++g) {
        GG_CUDA(cudaEventRecord(h->batch_ev[g], h->streams[g]));
        GG_CUDA(cudaStreamWaitEvent(h->copy_out, h->batch_ev[g], 0));
    }
    GG_CUDA(cudaEventRecord(h->batch_done[par], h->copy_out));
    h->batch_outstanding[par] = true;
    if (ticket) *ticket = par;
    h->last_feed_us = us_since();
    return GG_OK;
}

int gg_filter_cloud_batch_wait(gg_handle h, int ticket) {
    if (!h) return fail(GG_E_ARG, "null handle");
    if (ticket < 0 || ticket > 1) return fail(GG_E_ARG, "bad ticket %d", ticket);
    GG_CUDA(cudaSetDevice(h->device));
    return batch_wait(h, ticket);
}

int gg_filter_cloud_batch(gg_handle h, int count, const gg_scan_desc* scans, const gg_point* const* points, uint8_t* const* labels_out) {
    const auto t_begin = std::chrono::steady_clock::now();
    int rc = gg_filter_cloud_batch_begin(h, count, scans, points, labels_out, nullptr);
    if (rc) return rc;
    rc = gg_synchronize(h);
    h->last_total_us = (size_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_begin).count();
    return rc;
}

// Timing helpers: make every stream wait for the primary one / the primary one for all others,
// so that a CUDA-event pair on gg_stream() brackets work spread over the handle's streams.
int gg_fork_streams(gg_handle h) {
    if (!h) return fail(GG_E_ARG, "null handle");
    if (h->n_streams <= 1) return GG_OK;
    GG_CUDA(cudaSetDevice(h->device));
    cudaEvent_t ev;
    GG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GG_CUDA(cudaEventRecord(ev, h->streams[0]));
    for (int g = 1; g < h->n_streams; ++g) GG_CUDA(cudaStreamWaitEvent(h->streams[g], ev, 0));
    GG_CUDA(cudaEventDestroy(ev));
    return GG_OK;
}

int gg_join_streams(gg_handle h) {
    if (!h) return fail(GG_E_ARG, "null handle");
    if (h->n_streams <= 1) return GG_OK;
    GG_CUDA(cudaSetDevice(h->device));
    for (int g = 1; g < h->n_streams; ++g) {
        cudaEvent_t ev;
        GG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        GG_CUDA(cudaEventRecord(ev, h->streams[g]));
        GG_CUDA(cudaStreamWaitEvent(h->streams[0], ev, 0));
        GG_CUDA(cudaEventDestroy(ev));
    }
    return GG_OK;
}

int gg_num_streams(gg_handle h) { return h ? h->n_streams : GG_E_ARG; }

// host-only: the repacking of gg_filter_cloud_batch for one cloud (dst: 14 * ((n + 7) & ~7) bytes,
// 32-byte aligned), chunked like the worker threads do it -- exported for the CPU tests
static int pack_whole_cloud(const gg_point* src, size_t n, unsigned char* dst, bool cached) {
    if ((!src && n) || !dst) return GG_E_ARG;
    size_t i0 = 0;
    do {
        const size_t i1 = std::min(n, i0 + HostPacker::kChunk);
        HostPacker::pack_range(src, n, dst, i0, i1, cached);
        i0 = i1;
    } while (i0 < n);
    return GG_OK;
}
int gg_host_pack_cloud(const gg_point* src, size_t n, unsigned char* dst) { return pack_whole_cloud(src, n, dst, false); }
// same with plain (cache-allocating) stores
int gg_host_pack_cloud_cached(const gg_point* src, size_t n, unsigned char* dst) { return pack_whole_cloud(src, n, dst, true); }

// host-only self-test of the packer pool (no CUDA): `rounds` batches of n_jobs clouds go through a ring of
// `ring_slots` staging slots; this thread plays the feeder of gg_filter_cloud_batch_begin -- it checks every packed
// cloud against the single-threaded packing, "completes" its copy `lag` jobs later (so packers have to wait for
// slots), takes some jobs away from the back like the raw path does, and cancels the last batch half way.
// Returns 0, or a negative code telling which check failed.
int gg_host_packer_selftest(int threads, int n_jobs, size_t n_points, int ring_slots, int rounds, int lag) {
    if (threads < 1 || n_jobs < 1 || ring_slots < 2 || rounds < 1 || lag < 0 || lag >= ring_slots) return GG_E_ARG;
    const size_t n_pad = (n_points + 7) & ~(size_t)7, stride = 14 * n_pad + 64;
    std::vector<std::vector<gg_point>> src(n_jobs);
    std::vector<std::vector<unsigned char>> want(n_jobs);
    uint32_t lcg = 12345u;
    for (int j = 0; j < n_jobs; ++j) {
        const size_t n = n_points - (size_t)(j % 5);  // ragged sizes
        src[j].resize(n);
        unsigned char* raw = reinterpret_cast<unsigned char*>(src[j].data());
        for (size_t b = 0; b < n * sizeof(gg_point); ++b) {
            lcg = lcg * 1664525u + 1013904223u;
            raw[b] = (unsigned char)(lcg >> 24);
        }
        want[j].assign(stride + 32, 0);
        unsigned char* w = want[j].data() + ((32 - (reinterpret_cast<uintptr_t>(want[j].data()) & 31)) & 31);
        pack_whole_cloud(src[j].data(), n, w, false);
    }
    auto want_ptr = [&](int j) { return want[j].data() + ((32 - (reinterpret_cast<uintptr_t>(want[j].data()) & 31)) & 31); };
    std::vector<unsigned char> ring_mem((size_t)ring_slots * stride + 64);
    unsigned char* ring = ring_mem.data() + ((64 - (reinterpret_cast<uintptr_t>(ring_mem.data()) & 63)) & 63);
    HostPacker packer(threads);
    packer.set_ring(ring, stride, ring_slots);
    uint64_t issued = 0;
    for (int round = 0; round < rounds; ++round) {
        std::vector<PackJob> jobs(n_jobs);
        for (int j = 0; j < n_jobs; ++j) {
            jobs[j].src = src[j].data();
            jobs[j].n = src[j].size();
        }
        const bool cancel_round = round == rounds - 1;
        const uint64_t base = issued;
        packer.start(&jobs, base);
        int front = 0, back = n_jobs, spins = 0;
        while (front < back) {
            if (issued >= (uint64_t)lag) packer.copied.store(issued - (uint64_t)lag, std::memory_order_release);
            if (cancel_round && front >= n_jobs / 2) {
                packer.cancel();
                break;
            }
            if (packer.packed(jobs[front])) {
                const size_t np = (jobs[front].n + 7) & ~(size_t)7;
                if (std::memcmp(packer.slot_of(base + front), want_ptr(front), 14 * np) != 0) return -100 - front;
                ++issued;
                ++front;
                spins = 0;
                continue;
            }
            if ((front + round) % 3 == 0) {
                const int j = packer.claim_raw_from_back();
                if (j >= 0) {
                    if (j != back - 1) return -50;
                    back = j;
                    continue;
                }
            }
            if (!packer.help()) std::this_thread::yield();
            if (++spins > 200000000) return -60;  // stuck
        }
        packer.copied.store(issued, std::memory_order_release);  // all copies of this batch "done"
    }
    return GG_OK;
}


// number of host threads that repack clouds in gg_filter_cloud_batch (0: packing disabled or not used yet)
int gg_host_pack_threads(gg_handle h) { return (h && h->host_pack && h->packer) ? h->packer->threads() + 1 : 0; }

int gg_last_batch_transfer(gg_handle h, size_t info[9]) {
    if (!h || !info) return fail(GG_E_ARG, "null argument");
    info[0] = h->last_packed;
    info[1] = h->last_raw;
    info[2] = h->last_packed_bytes;
    info[3] = h->last_raw_bytes;
    info[4] = h->last_feed_us;
    info[5] = h->last_total_us;
    info[6] = h->last_pack_us;
    info[7] = h->last_slot_wait_us;
    info[8] = h->last_idle_us;
    return GG_OK;
}

int gg_get_layer(gg_handle h, int slot, const char* name, float* dst) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!dst) return fail(GG_E_ARG, "null dst");
    GG_CUDA(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->view.k.N2 * sizeof(float);
    if (name && std::strcmp(name, "expectedPoints") == 0) {
        GG_CUDA(cudaMemcpy(dst, h->view.expected, bytes, cudaMemcpyDeviceToHost));
        return GG_OK;
    }
    int idx;
    if ((rc = layer_index(h, slot, name, &idx))) return rc;
    if ((rc = gg_synchronize(h))) return rc;
    GG_CUDA(cudaMemcpy(dst, h->view.layer(slot, idx), bytes, cudaMemcpyDeviceToHost));
    return GG_OK;
}

int gg_set_layer(gg_handle h, int slot, const char* name, const float* src) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!src) return fail(GG_E_ARG, "null src");
    GG_CUDA(cudaSetDevice(h->device));
    int idx;
    if ((rc = layer_index(h, slot, name, &idx))) return rc;
    if ((rc = gg_synchronize(h))) return rc;
    GG_CUDA(cudaMemcpy(h->view.layer(slot, idx), src, (size_t)h->view.k.N2 * sizeof(float), cudaMemcpyHostToDevice));
    return GG_OK;
}

int gg_layer_device_ptr(gg_handle h, int slot, const char* name, void** dptr) {
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!dptr) return fail(GG_E_ARG, "null dptr");
    int idx;
    if ((rc = layer_index(h, slot, name, &idx))) return rc;
    *dptr = h->view.layer(slot, idx);
    return GG_OK;
}

}  // extern "C"
