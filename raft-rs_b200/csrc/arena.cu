// arena.cu -- host side of the engine: the HBM arena, staging rings, streams and
// the extern "C" surface declared in include/raftgpu.h.
//
// The arena owns every byte of device and pinned memory.  There is no CPU
// compute path: without a CUDA device arena creation fails with
// RAFTGPU_ERR_NO_DEVICE and nothing else can be called.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "kernels.cuh"

using namespace raftgpu;

namespace {

constexpr int kNumSets = 3;        // staging sets: one filling, up to two steps in flight
constexpr uint32_t kChunk = 2048;  // records per staging chunk (48 KiB)

// A record moved to a later wave: its packed form (1..4 PackedRec) plus where its result goes.
struct OverflowRec {
    PackedRec pk[4];
    int n_pk;
    uint32_t ring;
    uint64_t seq;  // position of the (main) record in the ring's enqueue order
};

// One caller thread's staging ring: a list of chunks of the set's shared pinned
// buffer.  Only its owner thread touches it between steps.
struct Ring {
    std::vector<uint32_t> chunks;        // chunk indices, in fill order
    uint32_t fill = kChunk;              // records used in the last chunk (kChunk = need a new one)
    uint64_t seq = 0;                    // records enqueued on this ring (wave 0 + overflow)
    std::vector<uint64_t> overflow_seq;  // seq numbers that went to a later wave (ascending)
};

struct StagingSet {
    // host (pinned)
    PackedRec *h_recs = nullptr;      // [n_chunks][kChunk] packed wave-0 records, shared by all rings
    PackedRec *h_overflow = nullptr;  // later waves, laid out at submit time
    uint32_t *h_adv_bitmap = nullptr;
    uint64_t *h_committed = nullptr;
    uint8_t *h_results = nullptr;
    uint32_t *h_step_adv = nullptr;
    // host (pageable)
    std::vector<Ring> rings;
    std::atomic<uint32_t> next_chunk{0};
    uint8_t *touched = nullptr;  // [cap] one bit per peer slot: cell has a record in wave 0
    std::mutex overflow_mu;
    std::unordered_map<uint64_t, uint32_t> overflow_depth;      // cell -> waves used beyond 0
    std::vector<std::vector<OverflowRec>> overflow_waves;       // wave w+1 records
    std::vector<std::pair<uint32_t, uint64_t>> overflow_order;  // (ring, seq | UINT64_MAX for EXT) per packed rec
    // device
    PackedRec *d_recs = nullptr;
    uint32_t *d_adv_bitmap = nullptr;
    uint64_t *d_commit_out = nullptr;
    uint8_t *d_results = nullptr;
    uint32_t *d_tile_off = nullptr;  // [cap / kFTile + 2]: tile index of a tileable compact stream
    uint32_t *d_step_adv = nullptr;  // [0] advanced groups of the step, [1] duplicate records (zero-copy)
    uint32_t *d_touched = nullptr;   // [cap/4] zero-copy steps: one bit per (group, slot)
    // sync
    cudaEvent_t ev_h2d = nullptr, ev_compute = nullptr, ev_done = nullptr;
    bool in_flight = false;
    bool dirty = false;  // touched[] has bits set
    uint32_t flags = 0;
    uint64_t wave0_slots = 0;  // staged wave-0 slots (chunks used * kChunk)
    raftgpu_step_result result{};
};

// The library's own staging workers for raftgpu_enqueue_bulk: persistent threads, pinned to
// the GPU-local CPUs, one staging ring each.
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    std::function<void(int)> job;

    void run(const std::function<void(int)> &fn) {
        std::unique_lock<std::mutex> lk(mu);
        job = fn;
        pending = static_cast<int>(threads.size());
        generation++;
        cv_start.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    void worker(int idx, cpu_set_t cpus, bool pin, int first_cpu) {
        if (pin) {
            // one CPU per worker, taken in order from the GPU-local list (its first half are
            // distinct physical cores on these hosts; SMT siblings come after), starting at an
            // offset derived from the device index so that the arenas of several GPUs on the same
            // socket (one process per GPU) do not pile onto the same cores
            cpu_set_t one;
            CPU_ZERO(&one);
            int seen = 0, chosen = -1;
            const int want = (first_cpu + idx) % std::max(1, CPU_COUNT(&cpus));
            for (int c = 0; c < CPU_SETSIZE; c++)
                if (CPU_ISSET(c, &cpus) && seen++ == want) {
                    chosen = c;
                    break;
                }
            if (chosen >= 0) {
                CPU_SET(chosen, &one);
                sched_setaffinity(0, sizeof(one), &one);
            } else {
                sched_setaffinity(0, sizeof(cpus), &cpus);
            }
        }
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int)> fn;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_start.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                fn = job;
            }
            fn(idx);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
            cv_start.notify_all();
        }
        for (auto &t : threads) t.join();
    }
};

}  // namespace

struct raftgpu_arena {
    int device = 0;
    uint32_t cap = 0;
    uint32_t n_rings = 0;
    uint32_t n_chunks = 0;  // chunks in each set's shared staging buffer
    uint64_t overflow_records = 0;
    uint32_t voter_hint = 0;                 // superset of every group's voter slots (recompute_kernel)
    int grid_recompute = 0, grid_recompute5 = 0, grid_apply = 0;  // persistent grid sizes (blocks)
    uint32_t n_simple5 = 0;                  // groups whose meta is the plain 5-voter configuration
    bool force_general = false;              // RAFTGPU_FORCE_GENERAL=1: never take the simple5 kernels
    bool prefetch = false;                   // RAFTGPU_PREFETCH=1: kernels with the L2 prefetch stage
    bool use_tma = false;                    // RAFTGPU_TMA=1 selects the TMA-fed recompute kernel
    uint32_t tma_smem = 0;                   // dynamic shared memory for the TMA stage ring
    uint32_t tile_smem = 0;                  // dynamic shared memory for the fused tile kernel
    int tma_stages_cap = 0;                  // RAFTGPU_TMA_STAGES (tuning knob)
    Columns cols{};
    unsigned long long *d_counters = nullptr;
    void *d_scratch = nullptr;  // 256 B for single-group queries
    void *h_scratch = nullptr;  // pinned mirror
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    StagingSet sets[kNumSets];
    int fill = 0;        // set currently being filled by enqueue
    int last_done = -1;  // set whose results raftgpu_step_results exposes
    int inflight[2] = {-1, -1};  // sets submitted by step_begin and not yet waited for, oldest first
    int n_inflight = 0;
    // control plane: single-group calls share the scratch buffers and the meta mirror
    std::mutex ctl_mu;
    // group allocation
    std::mutex alloc_mu;
    std::vector<uint32_t> free_list;
    std::vector<uint8_t> allocated;
    std::vector<uint32_t> h_meta;  // host mirror of the meta column
    uint32_t hi = 0, n_alloc = 0;
    // info
    int sm_count = 0;
    uint64_t l2_bytes = 0, device_bytes = 0, pinned_bytes = 0;
    std::string last_error;
    std::vector<void *> user_allocs;
    std::vector<void *> host_allocs;  // raftgpu_host_alloc
    HostPool *pool = nullptr;  // created on first raftgpu_enqueue_bulk
    cpu_set_t local_cpus;      // GPU-local CPUs (empty when unknown)
    bool have_local_cpus = false;
};

namespace {

thread_local std::string g_create_error;

// Pinned staging must live on the NUMA node the GPU's PCIe root hangs off: DMA reads of
// remote-socket memory cross the inter-socket link and lose about half of the H2D bandwidth
// (measured on the 2-socket B200 hosts: 27 GB/s remote vs ~55 GB/s local).  cudaHostAlloc
// places pages by first touch, so the allocating thread is moved onto the GPU-local CPUs
// (sysfs local_cpulist of the device) for the duration of the allocations.
struct LocalCpuGuard {
    cpu_set_t saved;
    cpu_set_t local;  // the GPU-local CPUs we are allowed to run on
    bool active = false;
    explicit LocalCpuGuard(int device) {
        char bus[32] = {0};
        if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return;
        for (char *p = bus; *p; p++) *p = static_cast<char>(tolower(*p));
        std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
        FILE *f = fopen(path.c_str(), "r");
        if (!f) return;
        char line[512] = {0};
        const bool ok = fgets(line, sizeof(line), f) != nullptr;
        fclose(f);
        if (!ok) return;
        cpu_set_t want;
        CPU_ZERO(&want);
        int n_set = 0;
        for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int lo = 0, hi = 0;
            const int k = sscanf(tok, "%d-%d", &lo, &hi);
            if (k == 1) hi = lo;
            if (k < 1) continue;
            for (int c = lo; c <= hi && c < CPU_SETSIZE; c++) {
                CPU_SET(c, &want);
                n_set++;
            }
        }
        if (!n_set) return;
        if (sched_getaffinity(0, sizeof(saved), &saved) != 0) return;
        cpu_set_t both;
        CPU_AND(&both, &want, &saved);  // stay inside the cpuset we are allowed to use
        if (CPU_COUNT(&both) == 0) return;
        local = both;
        if (sched_setaffinity(0, sizeof(both), &both) == 0) active = true;
    }
    ~LocalCpuGuard() {
        if (active) sched_setaffinity(0, sizeof(saved), &saved);
    }
};

int32_t fail(raftgpu_arena *a, int32_t code, const std::string &msg) {
    if (a) a->last_error = msg;
    return code;
}

#define CK(a, call)                                                                          \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            return fail((a), RAFTGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
        }                                                                                    \
    } while (0)

#define CKL(a)                                                                                    \
    do {                                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess)                                                                    \
            return fail((a), RAFTGPU_ERR_CUDA, std::string("launch: ") + cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
int32_t dev_alloc(raftgpu_arena *a, T **p, size_t count, bool zero = true) {
    const size_t bytes = count * sizeof(T);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), bytes ? bytes : 1);
    if (e != cudaSuccess)
        return fail(a, e == cudaErrorMemoryAllocation ? RAFTGPU_ERR_NOMEM : RAFTGPU_ERR_CUDA,
                    std::string("cudaMalloc: ") + cudaGetErrorString(e));
    a->device_bytes += bytes;
    if (zero) CK(a, cudaMemset(*p, 0, bytes ? bytes : 1));
    return RAFTGPU_OK;
}

template <typename T>
int32_t pin_alloc(raftgpu_arena *a, T **p, size_t count) {
    const size_t bytes = count * sizeof(T);
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(p), bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess)
        return fail(a, RAFTGPU_ERR_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    a->pinned_bytes += bytes;
    memset(*p, 0, bytes);  // first touch decides the NUMA node
    return RAFTGPU_OK;
}

inline bool group_ok(const raftgpu_arena *a, uint32_t g) { return g < a->cap && a->allocated[g]; }

inline uint32_t present_mask(uint32_t meta) {
    return RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta) | RAFTGPU_META_LEARN(meta);
}
inline uint32_t voter_mask(uint32_t meta) { return RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta); }

inline cudaStream_t pick_stream(raftgpu_arena *a, void *stream) {
    return stream ? static_cast<cudaStream_t>(stream) : a->s_compute;
}

inline uint32_t div_up(uint64_t a, uint32_t b) { return static_cast<uint32_t>((a + b - 1) / b); }

// True when every group of [first, first+n) is the plain 5-voter configuration in slots 0..4
// (no joint half, no group commit): the host's meta mirror is authoritative, so the kernels
// specialised for that case need no per-group check.
inline bool is_simple5(uint32_t meta) { return (meta & (0xffffu | RAFTGPU_META_GROUP_COMMIT)) == 0x1fu; }
inline bool range_simple5(const raftgpu_arena *a, uint32_t first, uint32_t n) {
    return a->n_simple5 == a->hi && static_cast<uint64_t>(first) + n <= a->hi;
}
inline void set_meta(raftgpu_arena *a, uint32_t g, uint32_t meta) {
    a->n_simple5 += static_cast<uint32_t>(is_simple5(meta)) - static_cast<uint32_t>(is_simple5(a->h_meta[g]));
    a->h_meta[g] = meta;
    a->voter_hint |= voter_mask(meta);
}

int32_t launch_recompute(raftgpu_arena *a, cudaStream_t st, uint32_t first, uint32_t n, uint32_t hint,
                         uint32_t *d_adv, uint64_t *d_commit, uint64_t *d_mci, uint8_t *d_gc,
                         uint32_t *d_step_adv) {
    if (n == 0) return RAFTGPU_OK;
    const bool simple5 = range_simple5(a, first, n) && !a->force_general;
    if (simple5) hint = 0x1fu;
    if (a->use_tma && n >= 16u * kTile && hint != 0) {
        // TMA-fed pipeline: one persistent CTA per SM, as many stages as fit in shared memory
        const uint32_t stage_bytes =
            (static_cast<uint32_t>(__builtin_popcount(hint & 0xffu)) + 3u) * kTile * 8u + kTile * 4u;
        int stages = std::min<int>(kMaxStages, static_cast<int>(a->tma_smem / stage_bytes));
        if (a->tma_stages_cap > 0) stages = std::min(stages, a->tma_stages_cap);
        if (stages >= 2) {
            const uint32_t base = first - (first % kTile);
            const uint32_t tiles = div_up(static_cast<uint64_t>(first - base) + n, kTile);
            const uint32_t blocks = std::min<uint32_t>(tiles, static_cast<uint32_t>(a->sm_count));
            const size_t smem = static_cast<size_t>(stages) * stage_bytes;
            if (simple5)
                recompute_tma_kernel<true><<<blocks, kTmaThreads, smem, st>>>(
                    a->cols, first, n, hint, stages, d_adv, d_commit, d_mci, d_gc, d_step_adv, a->d_counters);
            else
                recompute_tma_kernel<false><<<blocks, kTmaThreads, smem, st>>>(
                    a->cols, first, n, hint, stages, d_adv, d_commit, d_mci, d_gc, d_step_adv, a->d_counters);
            CKL(a);
            return RAFTGPU_OK;
        }
    }
    const uint32_t base = first & ~31u;
    const uint64_t threads = static_cast<uint64_t>(first - base) + n;
    if (simple5 && a->prefetch) {
        const uint32_t blocks = std::min<uint32_t>(div_up(threads, 256), static_cast<uint32_t>(a->grid_recompute5));
        recompute_kernel<true, true><<<blocks, 256, 0, st>>>(a->cols, first, n, hint, d_adv, d_commit, d_mci, d_gc,
                                                            d_step_adv, a->d_counters);
    } else if (simple5) {
        const uint32_t blocks = std::min<uint32_t>(div_up(threads, 256), static_cast<uint32_t>(a->grid_recompute5));
        recompute_kernel<true><<<blocks, 256, 0, st>>>(a->cols, first, n, hint, d_adv, d_commit, d_mci, d_gc,
                                                      d_step_adv, a->d_counters);
    } else {
        const uint32_t blocks = std::min<uint32_t>(div_up(threads, 256), static_cast<uint32_t>(a->grid_recompute));
        recompute_kernel<false><<<blocks, 256, 0, st>>>(a->cols, first, n, hint, d_adv, d_commit, d_mci, d_gc,
                                                       d_step_adv, a->d_counters);
    }
    CKL(a);
    return RAFTGPU_OK;
}

int32_t launch_apply(raftgpu_arena *a, cudaStream_t st, const void *d_recs, uint64_t n, uint8_t *d_results,
                     bool packed) {
    if (n == 0) return RAFTGPU_OK;
    const uint32_t blocks = std::min<uint32_t>(div_up(n, 256), static_cast<uint32_t>(a->grid_apply));
    if (packed && a->prefetch)
        apply_kernel<true, false, true><<<blocks, 256, 0, st>>>(a->cols, d_recs, n, d_results, a->d_counters);
    else if (packed)
        apply_kernel<true><<<blocks, 256, 0, st>>>(a->cols, d_recs, n, d_results, a->d_counters);
    else if (a->prefetch)
        apply_kernel<false, false, true><<<blocks, 256, 0, st>>>(a->cols, d_recs, n, d_results, a->d_counters);
    else
        apply_kernel<false><<<blocks, 256, 0, st>>>(a->cols, d_recs, n, d_results, a->d_counters);
    CKL(a);
    return RAFTGPU_OK;
}

void free_set(StagingSet &s) {
    cudaFreeHost(s.h_recs);
    cudaFreeHost(s.h_overflow);
    cudaFreeHost(s.h_adv_bitmap);
    cudaFreeHost(s.h_committed);
    cudaFreeHost(s.h_results);
    cudaFreeHost(s.h_step_adv);
    free(s.touched);
    cudaFree(s.d_recs);
    cudaFree(s.d_adv_bitmap);
    cudaFree(s.d_commit_out);
    cudaFree(s.d_results);
    cudaFree(s.d_tile_off);
    cudaFree(s.d_step_adv);
    cudaFree(s.d_touched);
    if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
    if (s.ev_compute) cudaEventDestroy(s.ev_compute);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
}

void destroy(raftgpu_arena *a) {
    if (!a) return;
    cudaSetDevice(a->device);
    cudaDeviceSynchronize();
    Columns &c = a->cols;
    cudaFree(c.matched);
    cudaFree(c.next_idx);
    cudaFree(c.peer_committed);
    cudaFree(c.pending_snapshot);
    cudaFree(c.pending_req_snapshot);
    cudaFree(c.commit_group_id);
    cudaFree(c.pflags);
    cudaFree(c.votes);
    cudaFree(c.meta);
    cudaFree(c.committed);
    cudaFree(c.term_start);
    cudaFree(c.last_index);
    cudaFree(a->d_counters);
    cudaFree(a->d_scratch);
    cudaFreeHost(a->h_scratch);
    for (void *p : a->user_allocs) cudaFree(p);
    for (void *p : a->host_allocs) cudaFreeHost(p);
    delete a->pool;
    for (auto &s : a->sets) free_set(s);
    if (a->s_compute) cudaStreamDestroy(a->s_compute);
    if (a->s_h2d) cudaStreamDestroy(a->s_h2d);
    if (a->s_d2h) cudaStreamDestroy(a->s_d2h);
    delete a;
}

int32_t create(int32_t device, uint32_t max_groups, uint32_t slots, uint32_t n_rings,
               uint32_t ring_records, raftgpu_arena **out) {
    if (!out || max_groups == 0 || slots != RAFTGPU_SLOTS) return RAFTGPU_ERR_INVALID;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return RAFTGPU_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n_dev) return RAFTGPU_ERR_INVALID;
    raftgpu_arena *a = new (std::nothrow) raftgpu_arena();
    if (!a) return RAFTGPU_ERR_NOMEM;
    a->device = device;
    // pad the column stride so every column (and every slot row of it) is 128-byte aligned
    a->cap = (max_groups + 127u) & ~127u;
    a->n_rings = n_rings ? n_rings : 16;
    // staging capacity per step, shared by all rings: ring_records * n_rings when given, else
    // ~5 records per group (a full round of a 5-peer deployment is ~3.5); plus one partially
    // filled chunk per ring
    const uint64_t want = ring_records ? static_cast<uint64_t>(ring_records) * a->n_rings
                                       : std::max<uint64_t>(4096, 5ull * a->cap);
    a->n_chunks = div_up(want, kChunk) + a->n_rings;
    a->overflow_records = std::max<uint64_t>(4096, a->cap / 4);
    int32_t rc = RAFTGPU_OK;
    auto bail = [&](int32_t code) {
        g_create_error = a->last_error;
        destroy(a);
        return code;
    };
#define TRY(x)                                 \
    do {                                       \
        rc = (x);                              \
        if (rc != RAFTGPU_OK) return bail(rc); \
    } while (0)
#define TRYC(call)                                                              \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            a->last_error = std::string(#call) + ": " + cudaGetErrorString(e_); \
            return bail(RAFTGPU_ERR_CUDA);                                      \
        }                                                                       \
    } while (0)
    TRYC(cudaSetDevice(device));
    cudaDeviceProp prop{};
    TRYC(cudaGetDeviceProperties(&prop, device));
    a->sm_count = prop.multiProcessorCount;
    a->l2_bytes = static_cast<uint64_t>(prop.l2CacheSize);
    int occ = 0;
    TRYC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, recompute_kernel<false>, 256, 0));
    a->grid_recompute = std::max(1, occ) * a->sm_count;
    TRYC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, recompute_kernel<true>, 256, 0));
    a->grid_recompute5 = std::max(1, occ) * a->sm_count;
    TRYC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, apply_kernel<true>, 256, 0));
    a->grid_apply = std::max(1, occ) * a->sm_count;
    a->tma_smem = 200u * 1024u;
    TRYC(cudaFuncSetAttribute(recompute_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(a->tma_smem)));
    TRYC(cudaFuncSetAttribute(recompute_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(a->tma_smem)));
    a->tile_smem = 227u * 1024u - 512u;  // 227 KB per CTA minus the kernels' static shared memory
#define RAFTGPU_TILE_ATTR(CT, NG)                                                                             \
    TRYC(cudaFuncSetAttribute(step_tile_kernel<false, CT, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                              static_cast<int>(a->tile_smem)));                                              \
    TRYC(cudaFuncSetAttribute(step_tile_kernel<true, CT, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                              static_cast<int>(a->tile_smem)))
#if RAFTGPU_TILE_GROUPS == 192
    RAFTGPU_TILE_ATTR(192, 3);
    RAFTGPU_TILE_ATTR(192, 4);
    RAFTGPU_TILE_ATTR(192, 5);
#elif RAFTGPU_TILE_GROUPS == 128
    RAFTGPU_TILE_ATTR(128, 3);
    RAFTGPU_TILE_ATTR(128, 4);
    RAFTGPU_TILE_ATTR(128, 6);
    RAFTGPU_TILE_ATTR(256, 2);
    RAFTGPU_TILE_ATTR(256, 3);
#else
    RAFTGPU_TILE_ATTR(256, 1);
    RAFTGPU_TILE_ATTR(256, 2);
    RAFTGPU_TILE_ATTR(256, 3);
    RAFTGPU_TILE_ATTR(512, 1);
#endif
#undef RAFTGPU_TILE_ATTR
#define RAFTGPU_CTILE_ATTR(NG)                                                                                   \
    TRYC(cudaFuncSetAttribute(step_tile_compact_kernel<false, NG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                              static_cast<int>(a->tile_smem)));                                                      \
    TRYC(cudaFuncSetAttribute(step_tile_compact_kernel<true, NG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                              static_cast<int>(a->tile_smem)));                                                      \
    TRYC(cudaFuncSetAttribute(step_tile_compact_kernel<false, NG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                              static_cast<int>(a->tile_smem)));                                                      \
    TRYC(cudaFuncSetAttribute(step_tile_compact_kernel<true, NG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                              static_cast<int>(a->tile_smem)))
    RAFTGPU_CTILE_ATTR(1);
    RAFTGPU_CTILE_ATTR(2);
    RAFTGPU_CTILE_ATTR(3);
#undef RAFTGPU_CTILE_ATTR
    {
        const char *e = getenv("RAFTGPU_TMA");
        a->use_tma = e && e[0] == '1';
        const char *f = getenv("RAFTGPU_FORCE_GENERAL");
        a->force_general = f && f[0] == '1';
        const char *pf = getenv("RAFTGPU_PREFETCH");
        a->prefetch = pf && pf[0] == '1';
        const char *t = getenv("RAFTGPU_TMA_STAGES");
        if (t) a->tma_stages_cap = atoi(t);
    }
    TRYC(cudaStreamCreateWithFlags(&a->s_compute, cudaStreamNonBlocking));
    TRYC(cudaStreamCreateWithFlags(&a->s_h2d, cudaStreamNonBlocking));
    TRYC(cudaStreamCreateWithFlags(&a->s_d2h, cudaStreamNonBlocking));

    const size_t cells = static_cast<size_t>(kSlots) * a->cap;
    Columns &c = a->cols;
    c.cap = a->cap;
    TRY(dev_alloc(a, &c.matched, cells));
    TRY(dev_alloc(a, &c.next_idx, cells));
    TRY(dev_alloc(a, &c.peer_committed, cells));
    TRY(dev_alloc(a, &c.pending_snapshot, cells));
    TRY(dev_alloc(a, &c.pending_req_snapshot, cells));
    TRY(dev_alloc(a, &c.commit_group_id, cells));
    TRY(dev_alloc(a, &c.pflags, cells));
    TRY(dev_alloc(a, &c.votes, cells));
    TRY(dev_alloc(a, &c.meta, a->cap));
    TRY(dev_alloc(a, &c.committed, a->cap));
    TRY(dev_alloc(a, &c.term_start, a->cap));
    TRY(dev_alloc(a, &c.last_index, a->cap));
    TRYC(cudaMemset(c.term_start, 0xff, sizeof(uint64_t) * a->cap));  // RAFTGPU_NO_TERM_START
    TRY(dev_alloc(a, &a->d_counters, kCntCount + 8));  // + 8 diagnostic slots (fused kernel phase cycles)
    TRY(dev_alloc(a, reinterpret_cast<uint8_t **>(&a->d_scratch), 256));
    TRY(pin_alloc(a, reinterpret_cast<uint8_t **>(&a->h_scratch), 256));

    const uint64_t wave0_total = static_cast<uint64_t>(a->n_chunks) * kChunk;
    const uint64_t rec_total = wave0_total + a->overflow_records;
    LocalCpuGuard numa_guard(device);  // pinned pages on the GPU-local NUMA node
    CPU_ZERO(&a->local_cpus);
    if (numa_guard.active) {
        a->local_cpus = numa_guard.local;
        a->have_local_cpus = true;
    }
    for (auto &s : a->sets) {
        TRY(pin_alloc(a, &s.h_recs, wave0_total));
        TRY(pin_alloc(a, &s.h_overflow, a->overflow_records));
        TRY(pin_alloc(a, &s.h_adv_bitmap, a->cap / 32));
        TRY(pin_alloc(a, &s.h_committed, a->cap));
        TRY(pin_alloc(a, &s.h_results, 4 * rec_total));  // one byte per slot; a compact stream has 4-byte units
        TRY(pin_alloc(a, &s.h_step_adv, 4));
        s.rings.assign(a->n_rings, Ring());
        s.touched = static_cast<uint8_t *>(calloc(a->cap, 1));
        if (!s.touched) return bail(RAFTGPU_ERR_NOMEM);
        TRY(dev_alloc(a, &s.d_recs, rec_total, false));
        TRY(dev_alloc(a, &s.d_adv_bitmap, a->cap / 32));
        TRY(dev_alloc(a, &s.d_commit_out, a->cap));
        TRY(dev_alloc(a, &s.d_results, 4 * rec_total));
        TRY(dev_alloc(a, &s.d_tile_off, 3 * (a->cap / kFTile + 2) + 2));
        TRY(dev_alloc(a, &s.d_step_adv, 4));
        TRY(dev_alloc(a, &s.d_touched, a->cap / 4));
        TRYC(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        TRYC(cudaEventCreateWithFlags(&s.ev_compute, cudaEventDisableTiming));
        TRYC(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    }
    a->allocated.assign(a->cap, 0);
    a->h_meta.assign(a->cap, 0);
    TRYC(cudaDeviceSynchronize());
#undef TRY
#undef TRYC
    *out = a;
    return RAFTGPU_OK;
}

// Control-plane kernel(s) + sync on the compute stream.
template <typename F>
int32_t sync_op(raftgpu_arena *a, F &&launch) {
    CK(a, cudaSetDevice(a->device));
    launch(a->s_compute);
    CKL(a);
    CK(a, cudaStreamSynchronize(a->s_compute));
    return RAFTGPU_OK;
}

struct ColumnDesc {
    void *base;
    size_t elem;
    bool per_peer;
};

bool column_desc(raftgpu_arena *a, int32_t col, ColumnDesc *d) {
    Columns &c = a->cols;
    switch (col) {
    case RAFTGPU_COL_MATCHED: *d = {c.matched, 8, true}; return true;
    case RAFTGPU_COL_NEXT_IDX: *d = {c.next_idx, 8, true}; return true;
    case RAFTGPU_COL_PEER_COMMITTED: *d = {c.peer_committed, 8, true}; return true;
    case RAFTGPU_COL_PENDING_SNAPSHOT: *d = {c.pending_snapshot, 8, true}; return true;
    case RAFTGPU_COL_PENDING_REQ_SNAPSHOT: *d = {c.pending_req_snapshot, 8, true}; return true;
    case RAFTGPU_COL_COMMIT_GROUP_ID: *d = {c.commit_group_id, 8, true}; return true;
    case RAFTGPU_COL_PFLAGS: *d = {c.pflags, 1, true}; return true;
    case RAFTGPU_COL_VOTES: *d = {c.votes, 1, true}; return true;
    case RAFTGPU_COL_META: *d = {c.meta, 4, false}; return true;
    case RAFTGPU_COL_COMMITTED: *d = {c.committed, 8, false}; return true;
    case RAFTGPU_COL_TERM_START: *d = {c.term_start, 8, false}; return true;
    case RAFTGPU_COL_LAST_INDEX: *d = {c.last_index, 8, false}; return true;
    default: return false;
    }
}

// Move a duplicate-cell record (already packed) to a later wave.
void push_overflow(StagingSet &s, uint64_t cell, const PackedRec *pk, int n_pk, uint32_t ring, uint64_t seq) {
    std::lock_guard<std::mutex> lk(s.overflow_mu);
    const uint32_t depth = s.overflow_depth[cell]++;  // 0 -> wave 1
    if (s.overflow_waves.size() <= depth) s.overflow_waves.resize(depth + 1);
    OverflowRec o{};
    for (int k = 0; k < n_pk; k++) o.pk[k] = pk[k];
    o.n_pk = n_pk;
    o.ring = ring;
    o.seq = seq;
    s.overflow_waves[depth].push_back(o);
}

// make a finished set reusable for filling
int32_t reclaim_set(raftgpu_arena *a, StagingSet &s) {
    if (s.in_flight) return RAFTGPU_ERR_BUSY;
    if (s.dirty) {
        memset(s.touched, 0, a->cap);
        s.dirty = false;
    }
    for (auto &r : s.rings) {
        r.chunks.clear();
        r.fill = kChunk;
        r.seq = 0;
        r.overflow_seq.clear();
    }
    s.next_chunk.store(0);
    s.overflow_depth.clear();
    s.overflow_waves.clear();
    return RAFTGPU_OK;
}

}  // namespace

// ===========================================================================
extern "C" {

const char *raftgpu_strerror(int32_t status) {
    switch (status) {
    case RAFTGPU_OK: return "ok";
    case RAFTGPU_ERR_INVALID: return "invalid argument";
    case RAFTGPU_ERR_CUDA: return "CUDA error";
    case RAFTGPU_ERR_NOMEM: return "out of memory / arena full";
    case RAFTGPU_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case RAFTGPU_ERR_RANGE: return "group or peer slot out of range";
    case RAFTGPU_ERR_FULL: return "staging ring full";
    case RAFTGPU_ERR_PEER_NOT_FOUND: return "peer not found (StepPeerNotFound)";
    case RAFTGPU_ERR_COMMIT_RANGE: return "to_commit is out of range [last_index]";
    case RAFTGPU_ERR_BUSY: return "step in flight";
    default: return "unknown status";
    }
}

uint32_t raftgpu_abi_version(void) { return RAFTGPU_ABI_VERSION; }

int32_t raftgpu_arena_create(int32_t device, uint32_t max_groups, uint32_t slots_per_group,
                             uint32_t n_rings, uint32_t ring_records, raftgpu_arena **out) {
    return create(device, max_groups, slots_per_group, n_rings, ring_records, out);
}

int32_t raftgpu_arena_destroy(raftgpu_arena *arena) {
    if (!arena) return RAFTGPU_ERR_INVALID;
    destroy(arena);
    return RAFTGPU_OK;
}

int32_t raftgpu_arena_info(const raftgpu_arena *a, raftgpu_info *out) {
    if (!a || !out) return RAFTGPU_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->abi_version = RAFTGPU_ABI_VERSION;
    out->device = a->device;
    out->cap = a->cap;
    out->slots = RAFTGPU_SLOTS;
    out->n_alloc = a->n_alloc;
    out->hi = a->hi;
    out->sm_count = static_cast<uint32_t>(a->sm_count);
    out->l2_bytes = a->l2_bytes;
    out->device_bytes = a->device_bytes;
    out->pinned_bytes = a->pinned_bytes;
    return RAFTGPU_OK;
}

const char *raftgpu_last_error(const raftgpu_arena *a) {
    return a ? a->last_error.c_str() : g_create_error.c_str();
}

// ---- group lifecycle -------------------------------------------------------

int32_t raftgpu_group_alloc(raftgpu_arena *a, uint32_t *out_group) {
    if (!a || !out_group) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> lk(a->alloc_mu);
    uint32_t g;
    if (!a->free_list.empty()) {
        g = a->free_list.back();
        a->free_list.pop_back();
    } else if (a->hi < a->cap) {
        g = a->hi;
    } else {
        return fail(a, RAFTGPU_ERR_NOMEM, "arena full");
    }
    a->allocated[g] = 1;
    a->n_alloc++;
    if (g >= a->hi) a->hi = g + 1;
    *out_group = g;
    return RAFTGPU_OK;
}

int32_t raftgpu_group_alloc_range(raftgpu_arena *a, uint32_t n, uint32_t *out_first) {
    if (!a || !out_first || n == 0) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> lk(a->alloc_mu);
    if (static_cast<uint64_t>(a->hi) + n > a->cap) return fail(a, RAFTGPU_ERR_NOMEM, "arena full");
    const uint32_t first = a->hi;
    memset(&a->allocated[first], 1, n);
    a->hi += n;
    a->n_alloc += n;
    *out_first = first;
    return RAFTGPU_OK;
}

int32_t raftgpu_group_free(raftgpu_arena *a, uint32_t g) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    // an empty configuration with no term range is inert in every kernel
    int32_t rc = sync_op(a, [&](cudaStream_t st) {
        conf_kernel<<<1, 1, 0, st>>>(a->cols, g, 0u, 0u, present_mask(a->h_meta[g]), 0);
        group_op_kernel<<<1, 1, 0, st>>>(a->cols, g, 0, RAFTGPU_NO_TERM_START, 0, nullptr);
    });
    if (rc != RAFTGPU_OK) return rc;
    CK(a, cudaMemsetAsync(a->cols.committed + g, 0, 8, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    std::lock_guard<std::mutex> lk(a->alloc_mu);
    set_meta(a, g, 0);
    a->allocated[g] = 0;
    a->n_alloc--;
    a->free_list.push_back(g);
    return RAFTGPU_OK;
}

int32_t raftgpu_group_set_conf(raftgpu_arena *a, uint32_t g, uint32_t incoming_mask,
                               uint32_t outgoing_mask, uint32_t learner_mask, int32_t self_slot,
                               uint64_t next_idx) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    if ((incoming_mask | outgoing_mask | learner_mask) > 0xffu || self_slot >= RAFTGPU_SLOTS)
        return RAFTGPU_ERR_INVALID;
    const uint32_t old_meta = a->h_meta[g];
    uint32_t meta = incoming_mask | (outgoing_mask << 8) | (learner_mask << 16) |
                    (old_meta & RAFTGPU_META_GROUP_COMMIT);
    if (self_slot >= 0) meta |= (static_cast<uint32_t>(self_slot) << 24) | RAFTGPU_META_HAS_SELF;
    const uint32_t was = present_mask(old_meta), now = present_mask(meta);
    int32_t rc = sync_op(a, [&](cudaStream_t st) {
        conf_kernel<<<1, 1, 0, st>>>(a->cols, g, meta, now & ~was, was & ~now, next_idx);
    });
    if (rc == RAFTGPU_OK) set_meta(a, g, meta);
    return rc;
}

int32_t raftgpu_group_reset(raftgpu_arena *a, uint32_t g, uint64_t term_start, uint64_t last_index,
                            uint64_t committed, uint64_t persisted) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    return sync_op(a, [&](cudaStream_t st) {
        reset_kernel<<<1, 1, 0, st>>>(a->cols, g, term_start, last_index, committed, persisted);
    });
}

int32_t raftgpu_group_become_leader(raftgpu_arena *a, uint32_t g) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    return sync_op(a, [&](cudaStream_t st) { become_leader_kernel<<<1, 1, 0, st>>>(a->cols, g); });
}

int32_t raftgpu_group_set_log_bounds(raftgpu_arena *a, uint32_t g, uint64_t term_start,
                                     uint64_t last_index) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    return sync_op(a, [&](cudaStream_t st) {
        group_op_kernel<<<1, 1, 0, st>>>(a->cols, g, 0, term_start, last_index, nullptr);
    });
}

int32_t raftgpu_group_commit_to(raftgpu_arena *a, uint32_t g, uint64_t to_commit) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    uint32_t *d = static_cast<uint32_t *>(a->d_scratch);
    uint32_t *h = static_cast<uint32_t *>(a->h_scratch);
    CK(a, cudaSetDevice(a->device));
    group_op_kernel<<<1, 1, 0, a->s_compute>>>(a->cols, g, 1, to_commit, 0, d);
    CKL(a);
    CK(a, cudaMemcpyAsync(h, d, 4, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    if (*h) return fail(a, RAFTGPU_ERR_COMMIT_RANGE, "to_commit is out of range [last_index]");
    return RAFTGPU_OK;
}

int32_t raftgpu_group_get(raftgpu_arena *a, uint32_t g, raftgpu_group_state *out) {
    if (!a || !out) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    auto *d = static_cast<raftgpu_group_state *>(a->d_scratch);
    CK(a, cudaSetDevice(a->device));
    group_get_kernel<<<1, 1, 0, a->s_compute>>>(a->cols, g, d);
    CKL(a);
    CK(a, cudaMemcpyAsync(a->h_scratch, d, sizeof(*out), cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    memcpy(out, a->h_scratch, sizeof(*out));
    return RAFTGPU_OK;
}

int32_t raftgpu_progress_get(raftgpu_arena *a, uint32_t g, uint32_t peer_slot, raftgpu_progress *out) {
    if (!a || !out) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g) || peer_slot >= RAFTGPU_SLOTS) return RAFTGPU_ERR_RANGE;
    auto *d = static_cast<raftgpu_progress *>(a->d_scratch);
    CK(a, cudaSetDevice(a->device));
    progress_get_kernel<<<1, 1, 0, a->s_compute>>>(a->cols, g, peer_slot, d);
    CKL(a);
    CK(a, cudaMemcpyAsync(a->h_scratch, d, sizeof(*out), cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    memcpy(out, a->h_scratch, sizeof(*out));
    return out->present ? RAFTGPU_OK : RAFTGPU_ERR_PEER_NOT_FOUND;
}

int32_t raftgpu_progress_set(raftgpu_arena *a, uint32_t g, uint32_t peer_slot,
                             const raftgpu_progress *in) {
    if (!a || !in) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g) || peer_slot >= RAFTGPU_SLOTS) return RAFTGPU_ERR_RANGE;
    if (!((present_mask(a->h_meta[g]) >> peer_slot) & 1u)) return RAFTGPU_ERR_PEER_NOT_FOUND;
    if (in->state > RAFTGPU_STATE_SNAPSHOT) return RAFTGPU_ERR_INVALID;
    raftgpu_progress p = *in;
    return sync_op(a, [&](cudaStream_t st) {
        progress_set_kernel<<<1, 1, 0, st>>>(a->cols, g, peer_slot, p);
    });
}

static int32_t scratch_i32_op(raftgpu_arena *a, int32_t *out, const std::function<void(cudaStream_t, int32_t *)> &launch) {
    int32_t *d = static_cast<int32_t *>(a->d_scratch) + 24;  // offset 96
    uint8_t *hs = static_cast<uint8_t *>(a->h_scratch);
    CK(a, cudaSetDevice(a->device));
    launch(a->s_compute, d);
    CKL(a);
    CK(a, cudaMemcpyAsync(hs + 96, d, 4, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    if (out) *out = *reinterpret_cast<int32_t *>(hs + 96);
    return RAFTGPU_OK;
}

int32_t raftgpu_progress_op(raftgpu_arena *a, uint32_t g, uint32_t peer_slot, int32_t op, uint64_t a0,
                            uint64_t a1, uint64_t a2, int32_t *out_result) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g) || peer_slot >= RAFTGPU_SLOTS) return RAFTGPU_ERR_RANGE;
    if (!((present_mask(a->h_meta[g]) >> peer_slot) & 1u)) return RAFTGPU_ERR_PEER_NOT_FOUND;
    if (op < 0 || op > RAFTGPU_POP_RESET) return RAFTGPU_ERR_INVALID;
    return scratch_i32_op(a, out_result, [&](cudaStream_t st, int32_t *d) {
        progress_op_kernel<<<1, 1, 0, st>>>(a->cols, g, peer_slot, op, a0, a1, a2, d);
    });
}

int32_t raftgpu_has_quorum(raftgpu_arena *a, uint32_t g, uint32_t slot_mask, int32_t *out) {
    if (!a || !out) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    return scratch_i32_op(a, out, [&](cudaStream_t st, int32_t *d) {
        quorum_kernel<<<1, 1, 0, st>>>(a->cols, g, 0, slot_mask & 0xffu, d);
    });
}

int32_t raftgpu_quorum_recently_active(raftgpu_arena *a, uint32_t g, uint32_t perspective_of_slot,
                                       int32_t *out) {
    if (!a || !out) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g) || perspective_of_slot >= RAFTGPU_SLOTS) return RAFTGPU_ERR_RANGE;
    return scratch_i32_op(a, out, [&](cudaStream_t st, int32_t *d) {
        quorum_kernel<<<1, 1, 0, st>>>(a->cols, g, 1, perspective_of_slot, d);
    });
}

int32_t raftgpu_group_maybe_commit_to(raftgpu_arena *a, uint32_t g, uint64_t max_index, int32_t *out_advanced) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    return scratch_i32_op(a, out_advanced, [&](cudaStream_t st, int32_t *d) {
        group_op_kernel<<<1, 1, 0, st>>>(a->cols, g, 6, max_index, 0, reinterpret_cast<uint32_t *>(d));
    });
}

int32_t raftgpu_set_group_commit(raftgpu_arena *a, uint32_t g, int32_t enable) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    int32_t rc = sync_op(a, [&](cudaStream_t st) {
        group_op_kernel<<<1, 1, 0, st>>>(a->cols, g, 2, RAFTGPU_META_GROUP_COMMIT, enable ? 1 : 0,
                                        nullptr);
    });
    if (rc == RAFTGPU_OK)
        set_meta(a, g, enable ? (a->h_meta[g] | RAFTGPU_META_GROUP_COMMIT)
                              : (a->h_meta[g] & ~RAFTGPU_META_GROUP_COMMIT));
    return rc;
}

int32_t raftgpu_assign_commit_group(raftgpu_arena *a, uint32_t g, uint32_t peer_slot,
                                    uint64_t commit_group_id) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g) || peer_slot >= RAFTGPU_SLOTS) return RAFTGPU_ERR_RANGE;
    // raft.rs:534-540: unknown peers are skipped silently
    if (!((present_mask(a->h_meta[g]) >> peer_slot) & 1u)) return RAFTGPU_OK;
    return sync_op(a, [&](cudaStream_t st) {
        group_op_kernel<<<1, 1, 0, st>>>(a->cols, g, 3, peer_slot, commit_group_id, nullptr);
    });
}

int32_t raftgpu_column_write(raftgpu_arena *a, int32_t column, uint32_t peer_slot, uint32_t first_group,
                             uint32_t n, const void *host_src) {
    ColumnDesc d;
    if (!a || !host_src || !column_desc(a, column, &d)) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (static_cast<uint64_t>(first_group) + n > a->cap || (d.per_peer && peer_slot >= RAFTGPU_SLOTS))
        return RAFTGPU_ERR_RANGE;
    CK(a, cudaSetDevice(a->device));
    const size_t off = ((d.per_peer ? static_cast<size_t>(peer_slot) * a->cap : 0) + first_group) * d.elem;
    CK(a, cudaMemcpyAsync(static_cast<uint8_t *>(d.base) + off, host_src, n * d.elem,
                          cudaMemcpyHostToDevice, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    if (column == RAFTGPU_COL_META) {
        const uint32_t *m = static_cast<const uint32_t *>(host_src);
        for (uint32_t i = 0; i < n; i++) set_meta(a, first_group + i, m[i]);
    }
    return RAFTGPU_OK;
}

int32_t raftgpu_column_read(raftgpu_arena *a, int32_t column, uint32_t peer_slot, uint32_t first_group,
                            uint32_t n, void *host_dst) {
    ColumnDesc d;
    if (!a || !host_dst || !column_desc(a, column, &d)) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (static_cast<uint64_t>(first_group) + n > a->cap || (d.per_peer && peer_slot >= RAFTGPU_SLOTS))
        return RAFTGPU_ERR_RANGE;
    CK(a, cudaSetDevice(a->device));
    const size_t off = ((d.per_peer ? static_cast<size_t>(peer_slot) * a->cap : 0) + first_group) * d.elem;
    CK(a, cudaMemcpyAsync(host_dst, static_cast<uint8_t *>(d.base) + off, n * d.elem,
                          cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    return RAFTGPU_OK;
}

// ---- hot path ---------------------------------------------------------------

int32_t raftgpu_maximal_committed_index(raftgpu_arena *a, uint32_t g, uint64_t *out_index,
                                        int32_t *out_use_group_commit) {
    if (!a || !out_index) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    CK(a, cudaSetDevice(a->device));
    // tracker.rs:294-298 has no side effect: the query kernel only evaluates the quorum.
    uint64_t *d_mci = static_cast<uint64_t *>(a->d_scratch);
    uint8_t *d_gc = static_cast<uint8_t *>(a->d_scratch) + 8;
    mci_kernel<<<1, 32, 0, a->s_compute>>>(a->cols, g, d_mci, d_gc);
    CKL(a);
    CK(a, cudaMemcpyAsync(a->h_scratch, a->d_scratch, 16, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    *out_index = *static_cast<uint64_t *>(a->h_scratch);
    if (out_use_group_commit) *out_use_group_commit = static_cast<uint8_t *>(a->h_scratch)[8];
    return RAFTGPU_OK;
}

int32_t raftgpu_maybe_commit(raftgpu_arena *a, uint32_t g, int32_t *out_advanced, uint64_t *out_committed) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    CK(a, cudaSetDevice(a->device));
    uint32_t *d_word = static_cast<uint32_t *>(a->d_scratch) + 8;  // offset 32
    CK(a, cudaMemsetAsync(d_word, 0, 4, a->s_compute));
    // the bitmap pointer is indexed by g >> 5 from group 0: bias it so that word lands on d_word
    int32_t rc = launch_recompute(a, a->s_compute, g, 1, voter_mask(a->h_meta[g]), d_word - (g >> 5),
                                  nullptr, nullptr, nullptr, nullptr);
    if (rc != RAFTGPU_OK) return rc;
    uint8_t *hs = static_cast<uint8_t *>(a->h_scratch);
    CK(a, cudaMemcpyAsync(hs + 32, d_word, 4, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaMemcpyAsync(hs + 40, a->cols.committed + g, 8, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    const uint32_t word = *reinterpret_cast<uint32_t *>(hs + 32);
    if (out_advanced) *out_advanced = (word >> (g & 31)) & 1u;
    if (out_committed) *out_committed = *reinterpret_cast<uint64_t *>(hs + 40);
    return RAFTGPU_OK;
}

int32_t raftgpu_recompute(raftgpu_arena *a, void *stream, uint32_t first, uint32_t n,
                          uint32_t *d_adv_bitmap, uint64_t *d_commit_out, uint64_t *d_mci_out,
                          uint8_t *d_gc_out) {
    if (!a) return RAFTGPU_ERR_INVALID;
    if (static_cast<uint64_t>(first) + n > a->cap) return RAFTGPU_ERR_RANGE;
    CK(a, cudaSetDevice(a->device));
    return launch_recompute(a, pick_stream(a, stream), first, n, a->voter_hint, d_adv_bitmap,
                            d_commit_out, d_mci_out, d_gc_out, nullptr);
}

int32_t raftgpu_apply_device(raftgpu_arena *a, void *stream, const raftgpu_append_resp *d_records,
                             uint64_t n, uint8_t *d_results) {
    if (!a || (!d_records && n)) return RAFTGPU_ERR_INVALID;
    CK(a, cudaSetDevice(a->device));
    return launch_apply(a, pick_stream(a, stream), d_records, n, d_results, /*packed=*/false);
}

int32_t raftgpu_step_sorted_device(raftgpu_arena *a, void *stream, const void *d_packed_records, uint64_t n_packed,
                                   const uint32_t *d_tile_off, uint8_t *d_results, uint32_t *d_adv_bitmap,
                                   uint64_t *d_commit_out) {
    if (!a || !d_tile_off || (!d_packed_records && n_packed)) return RAFTGPU_ERR_INVALID;
    CK(a, cudaSetDevice(a->device));
    const uint32_t hi = a->hi;
    if (hi == 0) return RAFTGPU_OK;
    const bool simple5 = range_simple5(a, 0, hi) && !a->force_general;
    const uint32_t hint = simple5 ? 0x1fu : (a->voter_hint & 0xffu);
    const uint32_t H = static_cast<uint32_t>(__builtin_popcount(hint));
    // variant knobs (tuning): RAFTGPU_TILE_VARIANT = <threads per consumer group><groups>, e.g. 2562, 5122, 2563
    static const int variant = getenv("RAFTGPU_TILE_VARIANT") ? atoi(getenv("RAFTGPU_TILE_VARIANT")) : 2563;
    static const int cap_env = getenv("RAFTGPU_TILE_RECCAP") ? atoi(getenv("RAFTGPU_TILE_RECCAP")) : 0;
    const uint32_t rec_cap = static_cast<uint32_t>(cap_env) & ~3u;
    const uint32_t stage_bytes = tile_stage_bytes(H, rec_cap);
    static const int stages_env = getenv("RAFTGPU_TILE_STAGES") ? atoi(getenv("RAFTGPU_TILE_STAGES")) : kFMaxStages;
    int stages = std::min<int>(std::min<int>(kFMaxStages, stages_env), static_cast<int>(a->tile_smem / stage_bytes));
    if (stages < 2 || H == 0) return fail(a, RAFTGPU_ERR_INVALID, "configuration too wide for the fused tile kernel");
    TileArgs t{};
    t.recs = static_cast<const PackedRec *>(d_packed_records);
    t.tile_off = d_tile_off;
    t.n_groups = hi;
    t.hint = hint;
    t.n_stages = stages;
    t.rec_cap = rec_cap;
    t.results = d_results;
    t.adv_bitmap = d_adv_bitmap;
    t.commit_out = d_commit_out;
    t.step_advanced = nullptr;
    t.counters = a->d_counters;
    static const bool tile_debug = getenv("RAFTGPU_TILE_DEBUG") != nullptr;
    t.dbg = tile_debug ? a->d_counters + kCntCount : nullptr;  // 8 spare u64 behind the counters
    const uint32_t n_tiles = div_up(hi, kFTile);
    const uint32_t blocks = std::min<uint32_t>(n_tiles, static_cast<uint32_t>(a->sm_count));
    const size_t smem = static_cast<size_t>(stages) * stage_bytes;
    cudaStream_t st = pick_stream(a, stream);
#define RAFTGPU_LAUNCH_TILE(CT, NG)                                                        \
    do {                                                                                   \
        if (simple5)                                                                       \
            step_tile_kernel<true, CT, NG><<<blocks, CT * NG + 64, smem, st>>>(a->cols, t); \
        else                                                                               \
            step_tile_kernel<false, CT, NG><<<blocks, CT * NG + 64, smem, st>>>(a->cols, t); \
    } while (0)
#if RAFTGPU_TILE_GROUPS == 192
    switch (variant) {
    case 1923: RAFTGPU_LAUNCH_TILE(192, 3); break;
    case 1925: RAFTGPU_LAUNCH_TILE(192, 5); break;
    default: RAFTGPU_LAUNCH_TILE(192, 4); break;
    }
#elif RAFTGPU_TILE_GROUPS == 128
    switch (variant) {
    case 1283: RAFTGPU_LAUNCH_TILE(128, 3); break;
    case 1286: RAFTGPU_LAUNCH_TILE(128, 6); break;
    case 2562: RAFTGPU_LAUNCH_TILE(256, 2); break;
    case 2563: RAFTGPU_LAUNCH_TILE(256, 3); break;
    default: RAFTGPU_LAUNCH_TILE(128, 4); break;
    }
#else
    switch (variant) {
    case 2561: RAFTGPU_LAUNCH_TILE(256, 1); break;
    case 2563: RAFTGPU_LAUNCH_TILE(256, 3); break;
    case 5121: RAFTGPU_LAUNCH_TILE(512, 1); break;
    default: RAFTGPU_LAUNCH_TILE(256, 2); break;
    }
#endif
#undef RAFTGPU_LAUNCH_TILE
    CKL(a);
    return RAFTGPU_OK;
}

namespace {

__global__ void fill_u32_kernel(uint32_t *p, uint32_t n, uint32_t v) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

CompactSrc compact_src(const void *d_blob, const raftgpu_compact_hdr &h) {
    const uint8_t *b = static_cast<const uint8_t *>(d_blob);
    CompactSrc src;
    src.units = reinterpret_cast<const uint32_t *>(b + h.off_units);
    src.g_base = reinterpret_cast<const uint32_t *>(b + h.off_blocks);
    src.side = reinterpret_cast<const raftgpu_append_resp *>(b + h.off_side);
    src.n_units = h.n_units;
    src.n_side = h.n_side;
    return src;
}

bool compact_hdr_ok(const raftgpu_compact_hdr &h, uint64_t blob_bytes) {
    const uint64_t need_blocks = (static_cast<uint64_t>(h.n_units) + RAFTGPU_COMPACT_BLOCK - 1) / RAFTGPU_COMPACT_BLOCK;
    return h.magic == RAFTGPU_COMPACT_MAGIC && h.total_bytes <= blob_bytes && h.n_blocks >= need_blocks &&
           !(h.off_blocks & 3u) && !(h.off_units & 15u) && !(h.off_side & 15u) &&
           h.off_blocks + 4ull * h.n_blocks <= h.total_bytes &&
           h.off_units + 4ull * ((static_cast<uint64_t>(h.n_units) + 3u) & ~3ull) <= h.total_bytes &&
           h.off_side + 24ull * h.n_side <= h.total_bytes;
}

// the tile table is one device buffer: u32 tile_off[n_tiles + 1], padding to 8 bytes, uint2 tile_gb[n_tiles]
inline uint2 *tile_gb_of(uint32_t *d_tile_off, uint32_t n_tiles) {
    return reinterpret_cast<uint2 *>(d_tile_off + ((n_tiles + 2u) & ~1u));
}
inline const uint2 *tile_gb_of_c(const uint32_t *d_tile_off, uint32_t n_tiles) {
    return reinterpret_cast<const uint2 *>(d_tile_off + ((n_tiles + 2u) & ~1u));
}

// tile_off[0..n_tiles] of a tileable stream over [0, hi): everything = n_units, then the headers write
int32_t launch_compact_tile_index(raftgpu_arena *a, cudaStream_t st, const CompactSrc &src, uint32_t *d_tile_off,
                                  uint32_t *d_bad) {
    const uint32_t n_tiles = div_up(a->hi, kFTile);
    fill_u32_kernel<<<div_up(n_tiles + 1, 256), 256, 0, st>>>(d_tile_off, n_tiles + 1, src.n_units);
    CKL(a);
    if (src.n_units) {
        const uint32_t blocks = std::min<uint32_t>(div_up(src.n_units, 256), 8u * static_cast<uint32_t>(a->sm_count));
        compact_tile_index_kernel<<<blocks, 256, 0, st>>>(src, a->hi, d_tile_off, d_bad);
        CKL(a);
    }
    // behind the offsets (8-byte aligned): the g_base pair of every tile
    compact_tile_gb_kernel<<<div_up(n_tiles, 256), 256, 0, st>>>(src, n_tiles, d_tile_off, tile_gb_of(d_tile_off, n_tiles));
    CKL(a);
    return RAFTGPU_OK;
}

// Can the fused compact kernel run this arena's configuration?  (stages that fit in shared memory)
bool ctile_plan(const raftgpu_arena *a, bool simple5, uint32_t *hint_out, uint32_t *unit_cap_out, int *stages_out,
                int *ng_out) {
    static const int ng_env = getenv("RAFTGPU_CTILE_GROUPS") ? atoi(getenv("RAFTGPU_CTILE_GROUPS")) : 3;
    static const int cap_env = getenv("RAFTGPU_CTILE_UNITCAP") ? atoi(getenv("RAFTGPU_CTILE_UNITCAP")) : 1536;
    const uint32_t hint = simple5 ? 0x1fu : (a->voter_hint & 0xffu);
    const uint32_t H = static_cast<uint32_t>(__builtin_popcount(hint));
    const int ng = std::max(1, std::min(3, ng_env));
    uint32_t unit_cap = static_cast<uint32_t>(std::max(64, cap_env)) & ~3u;
    if (H == 0) return false;
    int stages = 0;
    for (int s = kFMaxStages; s >= 2; s--)
        if (ctile_smem_bytes(H, unit_cap, s, ng) <= a->tile_smem) {
            stages = s;
            break;
        }
    if (stages < 2) return false;
    *hint_out = hint;
    *unit_cap_out = unit_cap;
    *stages_out = stages;
    *ng_out = ng;
    return true;
}

int32_t launch_tile_compact(raftgpu_arena *a, cudaStream_t st, const CompactSrc &src, const uint32_t *d_tile_off,
                            uint8_t *d_results, uint32_t *d_adv_bitmap, uint64_t *d_commit_out, uint32_t *d_step_adv,
                            bool ordered, uint32_t *d_dup_count) {
    const uint32_t hi = a->hi;
    if (hi == 0) return RAFTGPU_OK;
    const bool simple5 = range_simple5(a, 0, hi) && !a->force_general;
    uint32_t hint = 0, unit_cap = 0;
    int stages = 0, ng = 0;
    if (!ctile_plan(a, simple5, &hint, &unit_cap, &stages, &ng))
        return fail(a, RAFTGPU_ERR_INVALID, "configuration too wide for the fused tile kernel");
    const uint32_t H = static_cast<uint32_t>(__builtin_popcount(hint));
    CTileArgs t{};
    t.src = src;
    t.tile_off = d_tile_off;
    t.tile_gb = tile_gb_of_c(d_tile_off, div_up(hi, kFTile));
    t.n_groups = hi;
    t.hint = hint;
    t.n_stages = stages;
    t.unit_cap = unit_cap;
    t.results = d_results;
    t.adv_bitmap = d_adv_bitmap;
    t.commit_out = d_commit_out;
    t.step_advanced = d_step_adv;
    t.counters = a->d_counters;
    static const bool tile_debug = getenv("RAFTGPU_TILE_DEBUG") != nullptr;
    t.dbg = tile_debug ? a->d_counters + kCntCount : nullptr;
    t.dup_count = ordered ? nullptr : d_dup_count;
    const uint32_t n_tiles = div_up(hi, kFTile);
    const uint32_t blocks = std::min<uint32_t>(n_tiles, static_cast<uint32_t>(a->sm_count));
    const size_t smem = ctile_smem_bytes(H, unit_cap, stages, ng);
#define RAFTGPU_LAUNCH_CTILE(NG)                                                                                \
    do {                                                                                                        \
        if (simple5 && ordered)                                                                                 \
            step_tile_compact_kernel<true, NG, true><<<blocks, kFTile * NG + 64, smem, st>>>(a->cols, t);       \
        else if (simple5)                                                                                       \
            step_tile_compact_kernel<true, NG, false><<<blocks, kFTile * NG + 64, smem, st>>>(a->cols, t);      \
        else if (ordered)                                                                                       \
            step_tile_compact_kernel<false, NG, true><<<blocks, kFTile * NG + 64, smem, st>>>(a->cols, t);      \
        else                                                                                                    \
            step_tile_compact_kernel<false, NG, false><<<blocks, kFTile * NG + 64, smem, st>>>(a->cols, t);     \
    } while (0)
    switch (ng) {
    case 1: RAFTGPU_LAUNCH_CTILE(1); break;
    case 3: RAFTGPU_LAUNCH_CTILE(3); break;
    default: RAFTGPU_LAUNCH_CTILE(2); break;
    }
#undef RAFTGPU_LAUNCH_CTILE
    CKL(a);
    return RAFTGPU_OK;
}

}  // namespace

int32_t raftgpu_compact_tile_index_device(raftgpu_arena *a, void *stream, const void *d_blob,
                                          const raftgpu_compact_hdr *hdr, uint32_t *d_tile_off, uint32_t *d_bad) {
    if (!a || !d_blob || !hdr || !d_tile_off || !d_bad) return RAFTGPU_ERR_INVALID;
    if (!compact_hdr_ok(*hdr, hdr->total_bytes)) return fail(a, RAFTGPU_ERR_INVALID, "malformed compact batch header");
    CK(a, cudaSetDevice(a->device));
    return launch_compact_tile_index(a, pick_stream(a, stream), compact_src(d_blob, *hdr), d_tile_off, d_bad);
}

int32_t raftgpu_step_compact_device(raftgpu_arena *a, void *stream, const void *d_blob, const raftgpu_compact_hdr *hdr,
                                    const uint32_t *d_tile_off, uint8_t *d_results, uint32_t *d_adv_bitmap,
                                    uint64_t *d_commit_out, uint32_t *d_dup_count, uint32_t flags) {
    if (!a || !d_blob || !hdr || !d_tile_off) return RAFTGPU_ERR_INVALID;
    if (!compact_hdr_ok(*hdr, hdr->total_bytes)) return fail(a, RAFTGPU_ERR_INVALID, "malformed compact batch header");
    if (!(hdr->flags & RAFTGPU_COMPACT_TILEABLE))
        return fail(a, RAFTGPU_ERR_INVALID, "the stream is not tileable (groups not ascending, or a run without a header)");
    CK(a, cudaSetDevice(a->device));
    const bool ordered = (flags & RAFTGPU_COMPACT_STEP_ORDERED) || !(hdr->flags & RAFTGPU_COMPACT_ONE_WAVE);
    return launch_tile_compact(a, pick_stream(a, stream), compact_src(d_blob, *hdr), d_tile_off, d_results, d_adv_bitmap,
                               d_commit_out, nullptr, ordered, d_dup_count);
}

int32_t raftgpu_tile_index(const raftgpu_packed_rec *packed, uint64_t n_packed, uint32_t n_groups, uint32_t *out,
                           uint64_t out_capacity) {
    if ((!packed && n_packed) || !out) return RAFTGPU_ERR_INVALID;
    const uint32_t n_tiles = div_up(n_groups, kFTile);
    if (out_capacity < static_cast<uint64_t>(n_tiles) + 1 || n_packed > UINT32_MAX) return RAFTGPU_ERR_INVALID;
    uint32_t t = 0;  // next tile whose start is still unknown
    uint32_t prev_group = 0;
    for (uint64_t i = 0; i < n_packed; i++) {
        if (packed[i].w0 & kPkExt) continue;  // payloads / padding ride behind their record
        const uint32_t g = static_cast<uint32_t>(packed[i].w0);
        if (g < prev_group) return RAFTGPU_ERR_INVALID;  // not in group order
        prev_group = g;
        const uint32_t tile = g / kFTile;
        while (t <= tile && t <= n_tiles) out[t++] = static_cast<uint32_t>(i);
    }
    while (t <= n_tiles) out[t++] = static_cast<uint32_t>(n_packed);
    return RAFTGPU_OK;
}

int32_t raftgpu_apply_device_packed(raftgpu_arena *a, void *stream, const void *d_packed_records, uint64_t n,
                                    uint8_t *d_results) {
    if (!a || (!d_packed_records && n)) return RAFTGPU_ERR_INVALID;
    CK(a, cudaSetDevice(a->device));
    return launch_apply(a, pick_stream(a, stream), d_packed_records, n, d_results, /*packed=*/true);
}

// Streaming store of one packed record into the pinned ring: the destination is written once
// and next read by the DMA engine, so bypass the cache (no read-for-ownership traffic).
static inline void store_rec(PackedRec *dst, const PackedRec &r) {
#if defined(__x86_64__)
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst),
                     _mm_set_epi64x(static_cast<long long>(r.w1), static_cast<long long>(r.w0)));
#else
    *dst = r;
#endif
}

// Public 24-byte record (+ the EXT record of a REJECT) -> 1..4 packed 16-byte records
// (layout: kernels.cuh PackedRec).
static inline int pack_record(const raftgpu_append_resp &r, const raftgpu_append_resp *ext, PackedRec out[4]) {
    uint64_t w0 = static_cast<uint64_t>(r.group) | (static_cast<uint64_t>(r.peer_slot & 7u) << 32);
    if (r.peer_slot >= RAFTGPU_SLOTS) w0 = 0xffffffffull;  // no such slot: an out-of-range group says so
    if (r.flags & RAFTGPU_REC_REJECT) w0 |= kPkReject;
    if (r.flags & RAFTGPU_REC_LOCAL) w0 |= kPkLocal;
    if (ext) w0 |= kPkHasExt;
    bool wide = false;
    uint64_t delta = 0;
    if (r.flags & RAFTGPU_REC_LOCAL) {
        if (r.commit == 0)
            delta = kPkNoCommit;
        else if (r.commit >= r.index && r.commit - r.index < kPkNoCommit)
            delta = r.commit - r.index;
        else
            wide = true;
    } else if (r.commit <= r.index && r.index - r.commit <= 0xFFFFFFull) {
        delta = r.index - r.commit;
    } else {
        wide = true;
    }
    if (wide) w0 |= kPkWide;
    int n = 0;
    out[n++] = PackedRec{w0 | (delta << 40), r.index};
    if (ext) {
        out[n++] = PackedRec{kPkExt | (1ull << 40), ext->index};                  // next_probe_index
        if (ext->commit != RAFTGPU_INVALID_INDEX) out[n++] = PackedRec{kPkExt | (2ull << 40), ext->commit};
    }
    if (wide) out[n++] = PackedRec{kPkExt | (3ull << 40), r.commit};
    return n;
}

// Stage `n` records on one ring.  `sorted` = the caller promised non-decreasing group order
// (verified here): cells of one group then never straddle two rings, so the per-cell
// bookkeeping needs no atomics.  Without it the touched bits are updated atomically because
// another ring's thread may hold records of the same group.
static int32_t enqueue_ring(raftgpu_arena *a, StagingSet &s, uint32_t ring, const raftgpu_append_resp *recs,
                            uint64_t n, bool sorted, bool atomic_touch) {
    Ring &rg = s.rings[ring];
    PackedRec *dst = rg.chunks.empty() ? nullptr : s.h_recs + static_cast<size_t>(rg.chunks.back()) * kChunk;
    uint32_t fill = rg.fill;
    uint64_t seq = rg.seq;
    int32_t rc = RAFTGPU_OK;
    uint32_t prev_group = 0;
    const PackedRec pad{kPkExt, 0};
    PackedRec pk[4];
    for (uint64_t i = 0; i < n; i++) {
        const raftgpu_append_resp &r = recs[i];
        if (r.flags & RAFTGPU_REC_EXT) continue;  // travels with its REJECT
        if (sorted) {
            if (r.group < prev_group) {
                rc = RAFTGPU_ERR_INVALID;
                break;
            }
            prev_group = r.group;
        }
        bool has_ext = false;
        int n_pk;
        if (__builtin_expect(r.flags == 0 && r.commit <= r.index && r.index - r.commit <= 0xFFFFFFull &&
                                 r.peer_slot < RAFTGPU_SLOTS, 1)) {
            // the common record: an accepted AppendResponse
            pk[0] = PackedRec{static_cast<uint64_t>(r.group) | (static_cast<uint64_t>(r.peer_slot) << 32) |
                                  ((r.index - r.commit) << 40),
                              r.index};
            n_pk = 1;
        } else {
            has_ext = (r.flags & RAFTGPU_REC_REJECT) && i + 1 < n && (recs[i + 1].flags & RAFTGPU_REC_EXT);
            n_pk = pack_record(r, has_ext ? &recs[i + 1] : nullptr, pk);
        }
        const int n_orig = has_ext ? 2 : 1;
        bool dup = false;
        if (r.group < a->cap && r.peer_slot < RAFTGPU_SLOTS) {
            const uint8_t bit = static_cast<uint8_t>(1u << r.peer_slot);
            uint8_t &t = s.touched[r.group];
            if (atomic_touch) {
                dup = (__atomic_fetch_or(&t, bit, __ATOMIC_RELAXED) & bit) != 0;
            } else {
                dup = (t & bit) != 0;
                t |= bit;
            }
        }
        if (dup) {
            push_overflow(s, (static_cast<uint64_t>(r.group) << 3) | r.peer_slot, pk, n_pk, ring, seq);
            for (int k = 0; k < n_orig; k++) rg.overflow_seq.push_back(seq + k);
            seq += n_orig;
            continue;
        }
        if (fill + n_pk > kChunk) {
            // pad the tail of the chunk with no-op records (EXT records are skipped by the kernel)
            for (; dst && fill < kChunk; fill++) store_rec(&dst[fill], pad);
            const uint32_t ch = s.next_chunk.fetch_add(1);
            if (ch >= a->n_chunks) {
                rc = RAFTGPU_ERR_FULL;
                fill = kChunk;
                break;
            }
            rg.chunks.push_back(ch);
            dst = s.h_recs + static_cast<size_t>(ch) * kChunk;
            fill = 0;
        }
        for (int k = 0; k < n_pk; k++) store_rec(&dst[fill++], pk[k]);
        seq += n_orig;
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
    rg.fill = fill;
    rg.seq = seq;
    return rc;
}

int32_t raftgpu_enqueue_append_resp(raftgpu_arena *a, uint32_t ring, const raftgpu_append_resp *recs,
                                    uint64_t n) {
    if (!a || (!recs && n)) return RAFTGPU_ERR_INVALID;
    if (ring >= a->n_rings) return RAFTGPU_ERR_RANGE;
    StagingSet &s = a->sets[a->fill];
    if (s.in_flight) return RAFTGPU_ERR_BUSY;
    s.dirty = true;
    const int32_t rc = enqueue_ring(a, s, ring, recs, n, false, false);
    if (rc == RAFTGPU_ERR_FULL) return fail(a, rc, "staging ring full");
    return rc;
}

static void ensure_pool(raftgpu_arena *a) {
    if (a->pool) return;
    int want = 16;
    if (const char *e = getenv("RAFTGPU_HOST_THREADS")) want = atoi(e);
    int avail = a->have_local_cpus ? CPU_COUNT(&a->local_cpus) : static_cast<int>(std::thread::hardware_concurrency());
    want = std::max(1, std::min({want, static_cast<int>(a->n_rings), std::max(1, avail)}));
    a->pool = new HostPool();
    for (int t = 0; t < want; t++)
        a->pool->threads.emplace_back(&HostPool::worker, a->pool, t, a->local_cpus, a->have_local_cpus, a->device * want);
}

int32_t raftgpu_enqueue_bulk(raftgpu_arena *a, const raftgpu_append_resp *recs, uint64_t n, uint32_t flags) {
    if (!a || (!recs && n)) return RAFTGPU_ERR_INVALID;
    StagingSet &s = a->sets[a->fill];
    if (s.in_flight) return RAFTGPU_ERR_BUSY;
    s.dirty = true;
    const bool sorted = (flags & RAFTGPU_BULK_SORTED) != 0;
    ensure_pool(a);
    const int T = static_cast<int>(a->pool->threads.size());
    if (n < 4096 || T == 1) {
        const int32_t rc = enqueue_ring(a, s, 0, recs, n, sorted, false);
        if (rc == RAFTGPU_ERR_FULL) return fail(a, rc, "staging ring full");
        if (rc == RAFTGPU_ERR_INVALID) return fail(a, rc, "RAFTGPU_BULK_SORTED but records are not in group order");
        return rc;
    }
    // contiguous slices; a cut never separates a REJECT from its EXT, nor (sorted) one group
    std::vector<uint64_t> cut(T + 1, n);
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
        uint64_t c = std::max(cut[t - 1], n * t / T);
        while (c < n && c > 0 &&
               ((recs[c].flags & RAFTGPU_REC_EXT) || (sorted && recs[c].group == recs[c - 1].group)))
            c++;
        cut[t] = c;
    }
    std::vector<int32_t> rcs(T, RAFTGPU_OK);
    static const bool trace = getenv("RAFTGPU_TRACE") != nullptr;
    std::vector<double> wt(T, 0.0);
    const auto t_all = std::chrono::steady_clock::now();
    a->pool->run([&](int t) {
        const auto t0 = std::chrono::steady_clock::now();
        if (cut[t + 1] > cut[t])
            rcs[t] = enqueue_ring(a, s, static_cast<uint32_t>(t), recs + cut[t], cut[t + 1] - cut[t], sorted, !sorted);
        wt[t] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    });
    if (trace) {
        const double total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_all).count();
        fprintf(stderr, "[raftgpu] enqueue_bulk n=%llu T=%d total %.3f ms, workers min %.3f max %.3f ms\n",
                static_cast<unsigned long long>(n), T, total, *std::min_element(wt.begin(), wt.end()),
                *std::max_element(wt.begin(), wt.end()));
    }
    bool order_ok = true;
    if (sorted)  // slices are internally ordered (checked by the workers); check the seams too
        for (int t = 1; t < T; t++)
            if (cut[t] < n && cut[t] > 0 && recs[cut[t]].group < recs[cut[t] - 1].group) order_ok = false;
    for (int t = 0; t < T; t++) {
        if (rcs[t] == RAFTGPU_ERR_FULL) return fail(a, RAFTGPU_ERR_FULL, "staging ring full");
        if (rcs[t] == RAFTGPU_ERR_INVALID) order_ok = false;
    }
    if (!order_ok) return fail(a, RAFTGPU_ERR_INVALID, "RAFTGPU_BULK_SORTED but records are not in group order");
    return RAFTGPU_OK;
}

// Submit one step.  ext != nullptr: zero-copy -- wave 0 is the caller's pinned packed buffer and the
// GPU verifies the one-record-per-cell promise; otherwise wave 0 is what the rings staged.
static int32_t step_submit(raftgpu_arena *a, uint32_t flags, const PackedRec *ext, uint64_t ext_n,
                           const raftgpu_compact_hdr *cb = nullptr, bool cb_resident = false) {
    if (!a) return RAFTGPU_ERR_INVALID;
    if (a->n_inflight >= 2) return fail(a, RAFTGPU_ERR_BUSY, "two steps already in flight: call raftgpu_step_wait");
    CK(a, cudaSetDevice(a->device));
    StagingSet &s = a->sets[a->fill];
    if (ext || cb) {
        if (!cb_resident && (s.next_chunk.load() != 0 || !s.overflow_waves.empty()))
            return fail(a, RAFTGPU_ERR_INVALID, "records were enqueued for this step: cannot mix with a zero-copy batch");
        if (ext_n > static_cast<uint64_t>(a->n_chunks) * kChunk)
            return fail(a, RAFTGPU_ERR_FULL, "zero-copy batch larger than the device staging buffer");
        if (cb && cb->total_bytes > (static_cast<uint64_t>(a->n_chunks) * kChunk + a->overflow_records) * sizeof(PackedRec))
            return fail(a, RAFTGPU_ERR_FULL, "compact batch larger than the device staging buffer");
    }
    // pad every ring's last chunk, then ONE H2D of the used prefix of the shared buffer
    uint64_t n_real = 0;
    for (auto &rg : s.rings) {
        n_real += rg.seq - rg.overflow_seq.size();
        if (rg.chunks.empty()) continue;
        PackedRec *dst = s.h_recs + static_cast<size_t>(rg.chunks.back()) * kChunk;
        for (; rg.fill < kChunk; rg.fill++) dst[rg.fill] = PackedRec{kPkExt, 0};
    }
    const uint32_t used_chunks = std::min(s.next_chunk.load(), a->n_chunks);
    const uint64_t wave0 = cb ? cb->n_units : ext ? ext_n : static_cast<uint64_t>(used_chunks) * kChunk;
    if (ext) n_real = ext_n;
    if (cb) n_real = cb->n_records;
    if (cb && !cb_resident)
        CK(a, cudaMemcpyAsync(s.d_recs, cb, cb->total_bytes, cudaMemcpyHostToDevice, a->s_h2d));
    else if (cb)
        ;  // raftgpu_step_begin_records has already queued the copies of its segments on s_h2d
    else if (wave0)
        CK(a, cudaMemcpyAsync(s.d_recs, ext ? ext : s.h_recs, wave0 * sizeof(PackedRec),
                              cudaMemcpyHostToDevice, a->s_h2d));
    std::vector<uint64_t> wave_sizes;
    uint64_t ov = 0;
    s.overflow_order.clear();
    uint64_t ov_orig = 0;
    for (auto &w : s.overflow_waves) {
        const uint64_t before = ov;
        for (const OverflowRec &o : w) {
            if (ov + o.n_pk > a->overflow_records) return fail(a, RAFTGPU_ERR_FULL, "overflow staging full");
            for (int k = 0; k < o.n_pk; k++) {
                s.h_overflow[ov++] = o.pk[k];
                s.overflow_order.emplace_back(o.ring, k == 0 ? o.seq : UINT64_MAX);
            }
            ov_orig += (o.pk[0].w0 & kPkHasExt) ? 2 : 1;
        }
        wave_sizes.push_back(ov - before);
    }
    if (ov)
        CK(a, cudaMemcpyAsync(s.d_recs + wave0, s.h_overflow, ov * sizeof(PackedRec),
                              cudaMemcpyHostToDevice, a->s_h2d));
    CK(a, cudaEventRecord(s.ev_h2d, a->s_h2d));

    // compute: apply per wave, then one recompute pass over [0, hi)
    CK(a, cudaStreamWaitEvent(a->s_compute, s.ev_h2d, 0));
    uint8_t *d_res = (flags & RAFTGPU_STEP_READ_RESULTS) ? s.d_results : nullptr;
    CK(a, cudaMemsetAsync(s.d_step_adv, 0, 8, a->s_compute));
    int32_t rc = RAFTGPU_OK;
    bool fused = false;  // the fused tile kernel did apply AND recompute
    if (cb) {
        const CompactSrc src = compact_src(s.d_recs, *cb);
        static const bool force_scatter = getenv("RAFTGPU_COMPACT_SCATTER") != nullptr;
        uint32_t ph = 0, pc = 0;
        int ps = 0, pn = 0;
        fused = (cb->flags & RAFTGPU_COMPACT_TILEABLE) && !force_scatter && a->hi > 0 &&
                ctile_plan(a, range_simple5(a, 0, a->hi) && !a->force_general, &ph, &pc, &ps, &pn);
        if (fused) {
            // a tileable stream: tile index (d_step_adv[1] counts violations of the promise), then ONE kernel
            if (d_res && wave0) CK(a, cudaMemsetAsync(d_res, 0, wave0, a->s_compute));  // header units get no result
            rc = launch_compact_tile_index(a, a->s_compute, src, s.d_tile_off, s.d_step_adv + 1);
            if (rc != RAFTGPU_OK) return rc;
            static const bool force_ordered = getenv("RAFTGPU_COMPACT_ORDERED") != nullptr;
            rc = launch_tile_compact(a, a->s_compute, src, s.d_tile_off, d_res, s.d_adv_bitmap,
                                     (flags & RAFTGPU_STEP_READ_COMMITTED) ? s.d_commit_out : nullptr, s.d_step_adv,
                                     force_ordered || !(cb->flags & RAFTGPU_COMPACT_ONE_WAVE), s.d_step_adv + 1);
            if (rc != RAFTGPU_OK) return rc;
        } else if (wave0) {
            CK(a, cudaMemsetAsync(s.d_touched, 0, a->cap, a->s_compute));
            const uint32_t blocks = std::min<uint32_t>(div_up(wave0, 256), static_cast<uint32_t>(a->grid_apply));
            apply_compact_kernel<true><<<blocks, 256, 0, a->s_compute>>>(a->cols, src, d_res, a->d_counters, s.d_touched,
                                                                        s.d_step_adv + 1);
            CKL(a);
        }
    } else if (ext) {
        CK(a, cudaMemsetAsync(s.d_touched, 0, a->cap, a->s_compute));
        if (wave0) {
            const uint32_t blocks = std::min<uint32_t>(div_up(wave0, 256), static_cast<uint32_t>(a->grid_apply));
            apply_kernel<true, true><<<blocks, 256, 0, a->s_compute>>>(a->cols, s.d_recs, wave0, d_res, a->d_counters,
                                                                     s.d_touched, s.d_step_adv + 1);
            CKL(a);
        }
    } else {
        rc = launch_apply(a, a->s_compute, s.d_recs, wave0, d_res, /*packed=*/true);
    }
    if (rc != RAFTGPU_OK) return rc;
    uint64_t woff = wave0;
    for (uint64_t wsz : wave_sizes) {
        rc = launch_apply(a, a->s_compute, s.d_recs + woff, wsz, d_res ? d_res + woff : nullptr, /*packed=*/true);
        if (rc != RAFTGPU_OK) return rc;
        woff += wsz;
    }
    const uint32_t hi = a->hi;
    if (!fused) {
        rc = launch_recompute(a, a->s_compute, 0, hi, a->voter_hint, s.d_adv_bitmap,
                              (flags & RAFTGPU_STEP_READ_COMMITTED) ? s.d_commit_out : nullptr, nullptr, nullptr,
                              s.d_step_adv);
        if (rc != RAFTGPU_OK) return rc;
    }
    CK(a, cudaEventRecord(s.ev_compute, a->s_compute));

    // D2H of the results
    CK(a, cudaStreamWaitEvent(a->s_d2h, s.ev_compute, 0));
    CK(a, cudaMemcpyAsync(s.h_step_adv, s.d_step_adv, 8, cudaMemcpyDeviceToHost, a->s_d2h));
    if (hi)
        CK(a, cudaMemcpyAsync(s.h_adv_bitmap, s.d_adv_bitmap, 4ull * ((hi + 31) / 32),
                              cudaMemcpyDeviceToHost, a->s_d2h));
    if ((flags & RAFTGPU_STEP_READ_COMMITTED) && hi)
        CK(a, cudaMemcpyAsync(s.h_committed, s.d_commit_out, 8ull * hi, cudaMemcpyDeviceToHost, a->s_d2h));
    if (d_res && woff)
        CK(a, cudaMemcpyAsync(s.h_results, s.d_results, woff, cudaMemcpyDeviceToHost, a->s_d2h));
    CK(a, cudaEventRecord(s.ev_done, a->s_d2h));

    s.in_flight = true;
    s.flags = flags;
    s.wave0_slots = wave0;
    s.result.n_records = n_real + ov_orig;
    s.result.h2d_bytes = cb ? cb->total_bytes : (wave0 + ov) * sizeof(PackedRec);
    s.result.d2h_bytes = 8 + (hi ? 4ull * ((hi + 31) / 32) : 0) +
                         (((flags & RAFTGPU_STEP_READ_COMMITTED) && hi) ? 8ull * hi : 0) + ((d_res && woff) ? woff : 0);
    s.result.n_waves = static_cast<uint32_t>((n_real ? 1 : 0) + wave_sizes.size());
    s.result.n_groups = hi;
    a->inflight[a->n_inflight++] = a->fill;
    // flip: a set that is neither in flight nor holding the results of the last completed step
    // becomes the fill target (with 3 sets and <= 2 in flight there is always one)
    int other = -1;
    for (int k = 0; k < kNumSets; k++) {
        if (a->sets[k].in_flight) continue;
        if (k == a->last_done && other >= 0) continue;  // prefer keeping the last results alive
        if (other < 0 || k != a->last_done) other = k;
    }
    if (other < 0) return fail(a, RAFTGPU_ERR_BUSY, "no free staging set");
    rc = reclaim_set(a, a->sets[other]);
    if (rc != RAFTGPU_OK) return rc;
    if (a->last_done == other) a->last_done = -1;  // its results are gone
    a->fill = other;
    return RAFTGPU_OK;
}

uint32_t raftgpu_tile_groups(void) { return RAFTGPU_TILE_GROUPS; }

int32_t raftgpu_step_begin(raftgpu_arena *a, uint32_t flags) { return step_submit(a, flags, nullptr, 0); }

int32_t raftgpu_step_begin_packed(raftgpu_arena *a, const raftgpu_packed_rec *pinned_records, uint64_t n_packed,
                                  uint32_t flags) {
    if (!a || (!pinned_records && n_packed)) return RAFTGPU_ERR_INVALID;
    static_assert(sizeof(raftgpu_packed_rec) == sizeof(PackedRec), "packed record layout");
    static const PackedRec kNone{kPkExt, 0};
    return step_submit(a, flags, n_packed ? reinterpret_cast<const PackedRec *>(pinned_records) : &kNone,
                       n_packed);
}

int32_t raftgpu_pack_records(const raftgpu_append_resp *records, uint64_t n, raftgpu_packed_rec *out,
                             uint64_t out_capacity, uint64_t *out_n) {
    if ((!records && n) || !out || !out_n) return RAFTGPU_ERR_INVALID;
    uint64_t k = 0;
    PackedRec pk[4];
    for (uint64_t i = 0; i < n; i++) {
        const raftgpu_append_resp &r = records[i];
        if (r.flags & RAFTGPU_REC_EXT) continue;
        const bool has_ext = (r.flags & RAFTGPU_REC_REJECT) && i + 1 < n && (records[i + 1].flags & RAFTGPU_REC_EXT);
        const int n_pk = pack_record(r, has_ext ? &records[i + 1] : nullptr, pk);
        if (k + n_pk > out_capacity) return RAFTGPU_ERR_FULL;
        for (int j = 0; j < n_pk; j++) out[k++] = raftgpu_packed_rec{pk[j].w0, pk[j].w1};
    }
    *out_n = k;
    return RAFTGPU_OK;
}

static inline uint64_t align16(uint64_t x) { return (x + 15u) & ~15ull; }

uint64_t raftgpu_compact_bound(uint64_t n) {
    // worst case per record: its own run (2 header units + 1) or an ESC unit plus 24 side bytes
    const uint64_t units = 3 * n + 8;
    return sizeof(raftgpu_compact_hdr) + align16(4 * (units / RAFTGPU_COMPACT_BLOCK + 2)) + align16(4 * units) + 24 * n + 64;
}

namespace {

// One packer pass over records[lo, hi) -- a slice that starts at a main record and never separates a
// REJECT from its EXT: units and g_base words are written in place, side-table records are
// collected in a vector.  Unit positions (and therefore g_base blocks) are relative to
// the start of `units`, which the caller places on a RAFTGPU_COMPACT_BLOCK boundary of the stream.
struct PackOut {
    uint32_t *units = nullptr;
    uint64_t unit_cap = 0;
    uint32_t *g_base = nullptr;
    uint64_t gbase_cap = 0;
    std::vector<raftgpu_append_resp> side;
    std::vector<uint32_t> esc_pos;  // unit positions of the ESC units (their side index is relative to `side`)
    bool want_esc_pos = false;
    uint64_t nu = 0, n_rec = 0, blocks_set = 0;
    bool tileable = true, one_wave = true, any = false;
    uint32_t first_group = 0, last_group = 0;
};

int32_t pack_core(const raftgpu_append_resp *records, uint64_t lo, uint64_t hi, PackOut &o, uint32_t *unit_of_record,
                  uint32_t unit_base) {
    const uint64_t n = hi, n_total = hi;
    uint32_t *units = o.units;
    uint32_t *g_base = o.g_base;
    const uint64_t unit_cap = o.unit_cap;
    uint64_t nu = o.nu;
    bool have_prev_group = false;
    uint32_t prev_group = 0, seen_slots = 0;
    auto esc = [&](uint64_t i) -> bool {  // record i (and its EXT) to the side table, one ESC unit
        if (nu >= unit_cap || o.side.size() >= kCuPad - 2) return false;
        units[nu] = kCuEsc | (static_cast<uint32_t>(o.side.size()) << 2);
        if (o.want_esc_pos) o.esc_pos.push_back(static_cast<uint32_t>(nu));
        if (unit_of_record) unit_of_record[i] = unit_base + static_cast<uint32_t>(nu);
        nu++;
        o.side.push_back(records[i]);
        if ((records[i].flags & RAFTGPU_REC_REJECT) && i + 1 < n_total && (records[i + 1].flags & RAFTGPU_REC_EXT))
            o.side.push_back(records[i + 1]);
        return true;
    };
    uint64_t i = lo;
    while (i < n) {
        if (records[i].flags & RAFTGPU_REC_EXT) {  // stray continuation: carries nothing by itself
            if (unit_of_record) unit_of_record[i] = UINT32_MAX;
            i++;
            continue;
        }
        // the run: consecutive records of one group, at most 8 units (a REJECT may take two: its hint
        // rides in a payload unit)
        const uint32_t g = records[i].group;
        uint64_t e = i, max_index = 0;
        uint32_t run_units = 0;
        // Fast path for what a round mostly consists of: a run of plain accepts / leader-local
        // records.  Same segmentation and same units as the general code below.
        {
            uint32_t odd = 0;
            while (e < n && run_units < 8u) {
                const raftgpu_append_resp &r = records[e];
                if (r.group != g) break;
                odd |= (r.flags & ~RAFTGPU_REC_LOCAL) | (r.peer_slot >> 3);
                max_index = std::max(max_index, r.index);
                run_units++;
                e++;
            }
            // the run must end where the general scan would end it: not in front of a REJECT that
            // would still fit, nor in front of an EXT
            if (e < n && ((records[e].flags & RAFTGPU_REC_EXT) || (records[e].group == g && run_units < 8u))) odd = 1;
            const uint64_t base = max_index > 0x3fffu ? max_index - 0x3fffu : 0;
            const uint64_t b = nu / RAFTGPU_COMPACT_BLOCK;
            if (!odd && base < (1ull << 48) && b < o.gbase_cap && nu + 2 + run_units <= unit_cap) {
                while (o.blocks_set <= b) g_base[o.blocks_set++] = g;
                const uint32_t gb = g_base[b];
                if (g >= gb && g - gb <= 0xfffu) {
                    if (have_prev_group && g < prev_group) o.tileable = false;
                    if (!have_prev_group || g != prev_group) seen_slots = 0;
                    if (!o.any) {
                        o.any = true;
                        o.first_group = g;
                    }
                    o.last_group = g;
                    prev_group = g;
                    have_prev_group = true;
                    units[nu++] = kCuHdrA | (static_cast<uint32_t>(base & 0x3fffffffu) << 2);
                    units[nu++] = kCuHdrB | ((g - gb) << 2) | (static_cast<uint32_t>(base >> 30) << 14);
                    bool ok = true;
                    for (uint32_t back = 0; back < run_units; back++) {
                        const raftgpu_append_resp &r = records[i + back];
                        const bool is_local = r.flags != 0;
                        if ((seen_slots >> r.peer_slot) & 1u) o.one_wave = false;
                        seen_slots |= 1u << r.peer_slot;
                        const uint64_t cdl = is_local ? (r.commit == 0 ? kCuNoCommit : r.commit - r.index) : r.index - r.commit;
                        const bool fits = r.index >= base && (is_local ? (r.commit == 0 || (r.commit >= r.index && cdl < kCuNoCommit))
                                                                         : (r.commit <= r.index && cdl <= 255u));
                        if (fits) {
                            units[nu] = kCuRec | (is_local ? kCuLocal : 0u) | (back << 3) | (static_cast<uint32_t>(r.peer_slot) << 6) |
                                        (static_cast<uint32_t>(r.index - base) << 10) | (static_cast<uint32_t>(cdl) << 24);
                            if (unit_of_record) unit_of_record[i + back] = unit_base + static_cast<uint32_t>(nu);
                            nu++;
                        } else if (!esc(i + back)) {
                            ok = false;
                            break;
                        }
                    }
                    if (!ok) return RAFTGPU_ERR_FULL;
                    o.n_rec += run_units;
                    i = e;
                    continue;
                }
            }
            e = i;  // general path
            max_index = 0;
            run_units = 0;
        }
        while (e < n) {
            const raftgpu_append_resp &r = records[e];
            if (r.flags & RAFTGPU_REC_EXT) {
                e++;
                continue;
            }
            const uint32_t need = (r.flags & RAFTGPU_REC_REJECT) ? 2u : 1u;
            if (r.group != g || run_units + need > 8u) break;
            run_units += need;
            if (r.peer_slot < RAFTGPU_SLOTS) max_index = std::max(max_index, r.index);
            e++;
        }
        while (e < n && (records[e].flags & RAFTGPU_REC_EXT)) e++;  // the EXT of the run's last record
        const uint64_t base = max_index > 0x3fffu ? max_index - 0x3fffu : 0;
        bool header = base < (1ull << 48);  // also for a run of ESC units only: the fused kernel finds records by their run
        if (header) {
            const uint64_t b = nu / RAFTGPU_COMPACT_BLOCK;
            if (b >= o.gbase_cap) return RAFTGPU_ERR_FULL;
            while (o.blocks_set <= b) g_base[o.blocks_set++] = g;  // first header of the block names its g_base
            const uint32_t gb = g_base[b];
            if (g < gb || g - gb > 0xfffu) header = false;
        }
        if (nu + 2 + run_units > unit_cap) return RAFTGPU_ERR_FULL;
        if (!header || (have_prev_group && g < prev_group)) o.tileable = false;
        if (!have_prev_group || g != prev_group) seen_slots = 0;
        if (!o.any) {
            o.any = true;
            o.first_group = g;
        }
        o.last_group = g;
        prev_group = g;
        have_prev_group = true;
        if (header) {
            const uint32_t gl = g - g_base[nu / RAFTGPU_COMPACT_BLOCK];
            units[nu++] = kCuHdrA | (static_cast<uint32_t>(base & 0x3fffffffu) << 2);
            units[nu++] = kCuHdrB | (gl << 2) | (static_cast<uint32_t>(base >> 30) << 14);
        }
        uint32_t back = 0;
        for (uint64_t k = i; k < e; k++) {
            const raftgpu_append_resp &r = records[k];
            if (r.flags & RAFTGPU_REC_EXT) {
                if (unit_of_record) unit_of_record[k] = UINT32_MAX;
                continue;
            }
            o.n_rec++;
            if (r.peer_slot < RAFTGPU_SLOTS) {
                if ((seen_slots >> r.peer_slot) & 1u) o.one_wave = false;
                seen_slots |= 1u << r.peer_slot;
            }
            const bool is_local = r.flags == RAFTGPU_REC_LOCAL, is_reject = r.flags == RAFTGPU_REC_REJECT;
            bool compact = header && (r.flags == 0 || is_local || is_reject) && r.peer_slot < RAFTGPU_SLOTS &&
                           r.index >= base && r.index - base <= 0x3fffu;
            uint32_t cd = 0, payload = 0;
            if (compact) {
                if (is_local) {
                    if (r.commit == 0)
                        cd = kCuNoCommit;
                    else if (r.commit >= r.index && r.commit - r.index < kCuNoCommit)
                        cd = static_cast<uint32_t>(r.commit - r.index);
                    else
                        compact = false;
                } else if (r.commit <= r.index && r.index - r.commit <= 255u) {
                    cd = static_cast<uint32_t>(r.index - r.commit);
                } else {
                    compact = false;
                }
            }
            if (compact && is_reject) {
                // the EXT's next_probe_index hint as a signed 29-bit delta from the index; a snapshot
                // request (rare) sends the record to the side table
                const bool has_ext = k + 1 < n_total && (records[k + 1].flags & RAFTGPU_REC_EXT);
                const uint64_t hint = has_ext ? records[k + 1].index : 0;
                const uint64_t snapshot = has_ext ? records[k + 1].commit : RAFTGPU_INVALID_INDEX;
                const int64_t d = static_cast<int64_t>(hint - r.index);
                if (snapshot != RAFTGPU_INVALID_INDEX || d < -(1ll << 28) || d >= (1ll << 28))
                    compact = false;
                else
                    payload = kCuEsc | ((kCuPayload | (static_cast<uint32_t>(d) & (kCuPayload - 1u))) << 2);
            }
            if (compact) {
                units[nu] = kCuRec | (is_local ? kCuLocal : 0u) | (is_reject ? kCuReject : 0u) | (back << 3) |
                            (static_cast<uint32_t>(r.peer_slot) << 6) | (static_cast<uint32_t>(r.index - base) << 10) | (cd << 24);
                if (unit_of_record) unit_of_record[k] = unit_base + static_cast<uint32_t>(nu);
                nu++;
                back++;
                if (is_reject) {
                    units[nu++] = payload;
                    back++;
                }
            } else {
                if (!esc(k)) return RAFTGPU_ERR_FULL;
                back++;
            }
        }
        i = e;
    }
    o.nu = nu;
    return RAFTGPU_OK;
}

}  // namespace

int32_t raftgpu_pack_compact(const raftgpu_append_resp *records, uint64_t n, void *out, uint64_t out_capacity,
                             uint64_t *out_bytes, uint32_t *unit_of_record) {
    if ((!records && n) || !out || !out_bytes) return RAFTGPU_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(out) & 15u) return RAFTGPU_ERR_INVALID;
    const uint64_t max_units = 3 * n + 8;
    const uint64_t off_blocks = sizeof(raftgpu_compact_hdr);
    const uint64_t off_units = off_blocks + align16(4 * (max_units / RAFTGPU_COMPACT_BLOCK + 2));
    if (off_units > out_capacity) return RAFTGPU_ERR_FULL;
    uint8_t *blob = static_cast<uint8_t *>(out);
    PackOut o;
    o.g_base = reinterpret_cast<uint32_t *>(blob + off_blocks);
    o.gbase_cap = max_units / RAFTGPU_COMPACT_BLOCK + 2;
    o.units = reinterpret_cast<uint32_t *>(blob + off_units);
    o.unit_cap = std::min<uint64_t>((out_capacity - off_units) / 4, 0xfffffff0ull);
    const int32_t rc = pack_core(records, 0, n, o, unit_of_record, 0);
    if (rc != RAFTGPU_OK) return rc;
    uint64_t nu = o.nu;
    uint32_t *units = o.units;
    const uint64_t n_blocks = (nu + RAFTGPU_COMPACT_BLOCK - 1) / RAFTGPU_COMPACT_BLOCK;
    while (o.blocks_set < n_blocks) o.g_base[o.blocks_set++] = 0;
    while (nu & 3u) {
        if (nu >= o.unit_cap) return RAFTGPU_ERR_FULL;
        units[nu++] = kCuEsc | (kCuPad << 2);  // the fused kernel fetches units in 16-byte pieces
    }
    const uint64_t off_side = off_units + align16(4 * nu);
    const uint64_t total = off_side + align16(o.side.size() * sizeof(raftgpu_append_resp));
    if (total > out_capacity) return RAFTGPU_ERR_FULL;
    if (!o.side.empty()) memcpy(blob + off_side, o.side.data(), o.side.size() * sizeof(raftgpu_append_resp));
    raftgpu_compact_hdr h{};
    h.magic = RAFTGPU_COMPACT_MAGIC;
    h.n_units = static_cast<uint32_t>(nu);
    h.n_blocks = static_cast<uint32_t>(n_blocks);
    h.n_side = static_cast<uint32_t>(o.side.size());
    h.n_records = o.n_rec;
    h.off_blocks = off_blocks;
    h.off_units = off_units;
    h.off_side = off_side;
    h.total_bytes = total;
    h.flags = (o.tileable ? RAFTGPU_COMPACT_TILEABLE : 0u) | (o.tileable && o.one_wave ? RAFTGPU_COMPACT_ONE_WAVE : 0u);
    memcpy(blob, &h, sizeof(h));
    *out_bytes = total;
    return RAFTGPU_OK;
}

// Records in pageable host memory -> compact stream -> step, in one call.  The library's staging
// threads each pack one slice of the batch (cut at group boundaries) straight into the staging set's
// pinned buffer; a slice's units start on a unit-block boundary of the stream, so the slices need
// nothing from each other (g_base words are per block).  The device copy is the standard blob:
// header + g_base table, the unit segments (each padded to a block), the side records.
static int32_t step_begin_records_compact(raftgpu_arena *a, const raftgpu_append_resp *recs, uint64_t n, uint32_t flags) {
    if (!a || (!recs && n)) return RAFTGPU_ERR_INVALID;
    if (a->n_inflight >= 2) return fail(a, RAFTGPU_ERR_BUSY, "two steps already in flight: call raftgpu_step_wait");
    StagingSet &s = a->sets[a->fill];
    if (s.in_flight) return RAFTGPU_ERR_BUSY;
    if (s.next_chunk.load() != 0 || !s.overflow_waves.empty())
        return fail(a, RAFTGPU_ERR_INVALID, "records were enqueued for this step: cannot mix with raftgpu_step_begin_records");
    CK(a, cudaSetDevice(a->device));
    ensure_pool(a);
    const int T = (n < 8192) ? 1 : static_cast<int>(a->pool->threads.size());
    // slices: a cut never separates a REJECT from its EXT, nor the records of one group
    std::vector<uint64_t> cut(T + 1, n);
    cut[0] = 0;
    while (cut[0] < n && (recs[cut[0]].flags & RAFTGPU_REC_EXT)) cut[0]++;  // stray continuations carry nothing
    for (int t = 1; t < T; t++) {
        uint64_t c = std::max(cut[t - 1], n * t / T);
        while (c < n && c > 0 && ((recs[c].flags & RAFTGPU_REC_EXT) || recs[c].group == recs[c - 1].group)) c++;
        cut[t] = c;
    }
    // the staging set's pinned buffer: [header + g_base table | T unit regions | side records]
    uint8_t *buf = reinterpret_cast<uint8_t *>(s.h_recs);
    const uint64_t cap_bytes = static_cast<uint64_t>(a->n_chunks) * kChunk * sizeof(PackedRec);
    const uint64_t block_bytes = 4ull * RAFTGPU_COMPACT_BLOCK;
    const uint64_t meta_bytes = (sizeof(raftgpu_compact_hdr) + cap_bytes / RAFTGPU_COMPACT_BLOCK + 4096) & ~4095ull;
    if (cap_bytes < meta_bytes + (T + 1) * block_bytes * 2) return RAFTGPU_ERR_FULL;
    const uint64_t side_bytes = ((cap_bytes - meta_bytes) / 4) & ~4095ull;
    const uint64_t region_bytes = ((cap_bytes - meta_bytes - side_bytes) / T) / block_bytes * block_bytes;
    std::vector<PackOut> po(T);
    std::vector<int32_t> rcs(T, RAFTGPU_OK);
    std::vector<std::vector<uint32_t>> gb(T);
    auto work = [&](int t) {
        PackOut &o = po[t];
        o.units = reinterpret_cast<uint32_t *>(buf + meta_bytes + static_cast<uint64_t>(t) * region_bytes);
        o.unit_cap = region_bytes / 4;
        gb[t].assign(region_bytes / block_bytes + 1, 0u);
        o.g_base = gb[t].data();
        o.gbase_cap = gb[t].size();
        o.want_esc_pos = true;
        if (cut[t + 1] > cut[t]) rcs[t] = pack_core(recs, cut[t], cut[t + 1], o, nullptr, 0);
        if (rcs[t] != RAFTGPU_OK) return;
        while (o.nu % RAFTGPU_COMPACT_BLOCK) o.units[o.nu++] = kCuEsc | (kCuPad << 2);  // next slice starts a block
    };
    if (T == 1)
        work(0);
    else
        a->pool->run(work);
    for (int t = 0; t < T; t++)
        if (rcs[t] != RAFTGPU_OK) return rcs[t];
    // stitch: offsets, flags, g_base table, side records (+ their ESC indexes)
    raftgpu_compact_hdr h{};
    h.magic = RAFTGPU_COMPACT_MAGIC;
    h.off_blocks = sizeof(raftgpu_compact_hdr);
    bool tileable = true, one_wave = true, have_prev = false;
    uint32_t prev_last = 0;
    uint64_t nu = 0, ns = 0;
    uint32_t *g_base = reinterpret_cast<uint32_t *>(buf + h.off_blocks);
    raftgpu_append_resp *side = reinterpret_cast<raftgpu_append_resp *>(buf + cap_bytes - side_bytes);
    std::vector<uint64_t> unit_off(T, 0);
    for (int t = 0; t < T; t++) {
        PackOut &o = po[t];
        unit_off[t] = nu;
        tileable = tileable && o.tileable && (!o.any || !have_prev || o.first_group > prev_last);
        one_wave = one_wave && o.one_wave;
        if (o.any) {
            have_prev = true;
            prev_last = o.last_group;
        }
        const uint64_t nb = o.nu / RAFTGPU_COMPACT_BLOCK;
        for (uint64_t b2 = 0; b2 < nb; b2++) g_base[nu / RAFTGPU_COMPACT_BLOCK + b2] = b2 < o.blocks_set ? gb[t][b2] : 0u;
        if (!o.side.empty()) {
            if ((ns + o.side.size()) * sizeof(raftgpu_append_resp) > side_bytes || ns + o.side.size() >= kCuPad - 2)
                return RAFTGPU_ERR_FULL;
            memcpy(side + ns, o.side.data(), o.side.size() * sizeof(raftgpu_append_resp));
            for (uint32_t pos : o.esc_pos) o.units[pos] += static_cast<uint32_t>(ns) << 2;
        }
        nu += o.nu;
        ns += o.side.size();
        h.n_records += o.n_rec;
    }
    if (nu > 0xfffffff0ull) return RAFTGPU_ERR_FULL;
    h.n_units = static_cast<uint32_t>(nu);
    h.n_blocks = static_cast<uint32_t>(nu / RAFTGPU_COMPACT_BLOCK);
    h.n_side = static_cast<uint32_t>(ns);
    h.off_units = meta_bytes;
    h.off_side = align16(h.off_units + 4 * nu);
    h.total_bytes = h.off_side + align16(ns * sizeof(raftgpu_append_resp));
    h.flags = (tileable ? RAFTGPU_COMPACT_TILEABLE : 0u) | (tileable && one_wave ? RAFTGPU_COMPACT_ONE_WAVE : 0u);
    const uint64_t dev_bytes = (static_cast<uint64_t>(a->n_chunks) * kChunk + a->overflow_records) * sizeof(PackedRec);
    if (h.total_bytes > dev_bytes) return RAFTGPU_ERR_FULL;
    memcpy(buf, &h, sizeof(h));
    // H2D: header + g_base, every slice's units to its place in the stream, the side records
    uint8_t *d = reinterpret_cast<uint8_t *>(s.d_recs);
    CK(a, cudaMemcpyAsync(d, buf, h.off_blocks + 4ull * h.n_blocks, cudaMemcpyHostToDevice, a->s_h2d));
    for (int t = 0; t < T; t++)
        if (po[t].nu)
            CK(a, cudaMemcpyAsync(d + h.off_units + 4 * unit_off[t], po[t].units, 4 * po[t].nu, cudaMemcpyHostToDevice, a->s_h2d));
    if (ns) CK(a, cudaMemcpyAsync(d + h.off_side, side, ns * sizeof(raftgpu_append_resp), cudaMemcpyHostToDevice, a->s_h2d));
    s.dirty = true;
    return step_submit(a, flags, nullptr, 0, &h, /*cb_resident=*/true);
}

int32_t raftgpu_step_begin_records(raftgpu_arena *a, const raftgpu_append_resp *recs, uint64_t n, uint32_t flags) {
    int32_t rc = step_begin_records_compact(a, recs, n, flags);
    if (rc != RAFTGPU_ERR_FULL) return rc;
    // A batch the compact form cannot hold in the staging buffer (groups in no order, every record
    // escaping to the side table): the general staging path -- packed 16-byte records, waves.
    rc = raftgpu_enqueue_bulk(a, recs, n, 0);
    if (rc != RAFTGPU_OK) return rc;
    return raftgpu_step_begin(a, flags);
}

int32_t raftgpu_step_begin_compact(raftgpu_arena *a, const void *pinned_blob, uint64_t blob_bytes, uint32_t flags) {
    if (!a || !pinned_blob || blob_bytes < sizeof(raftgpu_compact_hdr)) return RAFTGPU_ERR_INVALID;
    const raftgpu_compact_hdr *h = static_cast<const raftgpu_compact_hdr *>(pinned_blob);
    if (!compact_hdr_ok(*h, blob_bytes)) return fail(a, RAFTGPU_ERR_INVALID, "malformed compact batch header");
    return step_submit(a, flags, nullptr, 0, h);
}

int32_t raftgpu_host_alloc(raftgpu_arena *a, uint64_t bytes, void **out_pinned) {
    if (!a || !out_pinned) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    CK(a, cudaSetDevice(a->device));
    LocalCpuGuard numa_guard(a->device);  // first touch on the GPU-local NUMA node
    uint8_t *p = nullptr;
    int32_t rc = pin_alloc(a, &p, bytes);
    if (rc != RAFTGPU_OK) return rc;
    a->host_allocs.push_back(p);
    *out_pinned = p;
    return RAFTGPU_OK;
}

int32_t raftgpu_host_free(raftgpu_arena *a, void *pinned) {
    if (!a || !pinned) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    for (size_t i = 0; i < a->host_allocs.size(); i++)
        if (a->host_allocs[i] == pinned) {
            a->host_allocs.erase(a->host_allocs.begin() + i);
            CK(a, cudaFreeHost(pinned));
            return RAFTGPU_OK;
        }
    return RAFTGPU_ERR_INVALID;
}

int32_t raftgpu_step_wait(raftgpu_arena *a, raftgpu_step_result *out) {
    if (!a) return RAFTGPU_ERR_INVALID;
    if (a->n_inflight == 0) return fail(a, RAFTGPU_ERR_INVALID, "no step in flight");
    const int cur = a->inflight[0];
    StagingSet &s = a->sets[cur];
    CK(a, cudaEventSynchronize(s.ev_done));
    s.in_flight = false;
    s.result.n_advanced = s.h_step_adv[0];
    s.result.n_duplicates = s.h_step_adv[1];
    if (out) *out = s.result;
    a->last_done = cur;
    a->inflight[0] = a->inflight[1];
    a->inflight[1] = -1;
    a->n_inflight--;
    if (s.result.n_duplicates)
        return fail(a, RAFTGPU_ERR_INVALID, "zero-copy batch had more than one record for a (group, peer) cell");
    return RAFTGPU_OK;
}

int32_t raftgpu_step(raftgpu_arena *a, uint32_t flags, raftgpu_step_result *out) {
    int32_t rc = raftgpu_step_begin(a, flags);
    if (rc != RAFTGPU_OK) return rc;
    return raftgpu_step_wait(a, out);
}

int32_t raftgpu_step_results(raftgpu_arena *a, const uint32_t **adv_bitmap, const uint64_t **committed) {
    if (!a) return RAFTGPU_ERR_INVALID;
    if (a->last_done < 0) return fail(a, RAFTGPU_ERR_INVALID, "no completed step");
    StagingSet &s = a->sets[a->last_done];
    if (adv_bitmap) *adv_bitmap = s.h_adv_bitmap;
    if (committed) *committed = (s.flags & RAFTGPU_STEP_READ_COMMITTED) ? s.h_committed : nullptr;
    return RAFTGPU_OK;
}

int32_t raftgpu_step_slot_results(raftgpu_arena *a, const uint8_t **results, uint64_t *n_slots) {
    if (!a || !results || !n_slots) return RAFTGPU_ERR_INVALID;
    if (a->last_done < 0) return fail(a, RAFTGPU_ERR_INVALID, "no completed step");
    StagingSet &s = a->sets[a->last_done];
    if (!(s.flags & RAFTGPU_STEP_READ_RESULTS))
        return fail(a, RAFTGPU_ERR_INVALID, "step ran without RAFTGPU_STEP_READ_RESULTS");
    *results = s.h_results;
    *n_slots = s.wave0_slots;
    return RAFTGPU_OK;
}

int32_t raftgpu_step_record_results(raftgpu_arena *a, uint32_t ring, uint8_t *out, uint64_t cap,
                                    uint64_t *out_n) {
    if (!a || !out_n) return RAFTGPU_ERR_INVALID;
    if (ring >= a->n_rings) return RAFTGPU_ERR_RANGE;
    if (a->last_done < 0) return fail(a, RAFTGPU_ERR_INVALID, "no completed step");
    StagingSet &s = a->sets[a->last_done];
    if (!(s.flags & RAFTGPU_STEP_READ_RESULTS))
        return fail(a, RAFTGPU_ERR_INVALID, "step ran without RAFTGPU_STEP_READ_RESULTS");
    const Ring &rg = s.rings[ring];
    *out_n = rg.seq;
    if (!out) return RAFTGPU_OK;
    if (cap < rg.seq) return RAFTGPU_ERR_INVALID;
    memset(out, 0, rg.seq);
    // later-wave records of this ring, by seq
    for (size_t k = 0; k < s.overflow_order.size(); k++)
        if (s.overflow_order[k].first == ring && s.overflow_order[k].second != UINT64_MAX)
            out[s.overflow_order[k].second] = s.h_results[s.wave0_slots + k];
    // wave-0 records fill the remaining seq positions in chunk order: one result per main packed
    // record, which stands for 1 public record (2 with its EXT)
    size_t ov = 0;
    uint64_t seq = 0;
    for (size_t ci = 0; ci < rg.chunks.size() && seq < rg.seq; ci++) {
        const uint8_t *res = s.h_results + static_cast<size_t>(rg.chunks[ci]) * kChunk;
        const PackedRec *src = s.h_recs + static_cast<size_t>(rg.chunks[ci]) * kChunk;
        for (uint32_t k = 0; k < kChunk && seq < rg.seq; k++) {
            if (src[k].w0 & kPkExt) continue;  // EXT payloads and padding
            while (ov < rg.overflow_seq.size() && rg.overflow_seq[ov] == seq) {
                ov++;
                seq++;
            }
            if (seq >= rg.seq) break;
            out[seq] = res[k];
            seq += (src[k].w0 & kPkHasExt) ? 2 : 1;
        }
    }
    return RAFTGPU_OK;
}

// ---- votes ------------------------------------------------------------------

int32_t raftgpu_reset_votes(raftgpu_arena *a, uint32_t g) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    return sync_op(a, [&](cudaStream_t st) { group_op_kernel<<<1, 1, 0, st>>>(a->cols, g, 4, 0, 0, nullptr); });
}

int32_t raftgpu_record_vote(raftgpu_arena *a, uint32_t g, uint32_t peer_slot, int32_t vote) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g) || peer_slot >= RAFTGPU_SLOTS) return RAFTGPU_ERR_RANGE;
    return sync_op(a, [&](cudaStream_t st) {
        group_op_kernel<<<1, 1, 0, st>>>(a->cols, g, 5, peer_slot, vote ? 2 : 1, nullptr);
    });
}

int32_t raftgpu_send_list_device(raftgpu_arena *a, void *stream, uint32_t first, uint32_t n,
                                 const uint32_t *d_adv_bitmap, raftgpu_send_entry *d_out, uint64_t capacity,
                                 uint64_t *d_count) {
    static_assert(sizeof(raftgpu_send_entry) == 16, "send entry layout");
    if (!a || !d_count || (!d_out && capacity)) return RAFTGPU_ERR_INVALID;
    if (static_cast<uint64_t>(first) + n > a->cap) return RAFTGPU_ERR_RANGE;
    CK(a, cudaSetDevice(a->device));
    cudaStream_t st = pick_stream(a, stream);
    CK(a, cudaMemsetAsync(d_count, 0, 8, st));
    if (n == 0) return RAFTGPU_OK;
    const uint32_t blocks = std::min<uint32_t>(div_up(n, 256), 8u * static_cast<uint32_t>(a->sm_count));
    send_list_kernel<<<blocks, 256, 0, st>>>(a->cols, first, n, d_adv_bitmap, d_out, capacity,
                                             reinterpret_cast<unsigned long long *>(d_count));
    CKL(a);
    return RAFTGPU_OK;
}

int32_t raftgpu_step_send_list(raftgpu_arena *a, raftgpu_send_entry *out, uint64_t capacity, uint64_t *out_n) {
    if (!a || !out_n || (!out && capacity)) return RAFTGPU_ERR_INVALID;
    if (a->last_done < 0) return fail(a, RAFTGPU_ERR_INVALID, "no completed step");
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    StagingSet &s = a->sets[a->last_done];
    CK(a, cudaSetDevice(a->device));
    // the set's record staging is idle between its step_wait and its next submission: the entries go there
    const uint64_t room = (static_cast<uint64_t>(a->n_chunks) * kChunk + a->overflow_records) * sizeof(PackedRec) / sizeof(raftgpu_send_entry);
    const uint64_t cap_dev = std::min(capacity, room);
    uint64_t *d_count = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(a->d_scratch) + 128);
    int32_t rc = raftgpu_send_list_device(a, a->s_compute, 0, a->hi, s.d_adv_bitmap,
                                          reinterpret_cast<raftgpu_send_entry *>(s.d_recs), cap_dev, d_count);
    if (rc != RAFTGPU_OK) return rc;
    uint8_t *hs = static_cast<uint8_t *>(a->h_scratch);
    CK(a, cudaMemcpyAsync(hs + 128, d_count, 8, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    const uint64_t total = *reinterpret_cast<uint64_t *>(hs + 128);
    *out_n = total;
    if (total > cap_dev) return fail(a, RAFTGPU_ERR_FULL, "send list larger than the buffer");
    if (total) CK(a, cudaMemcpy(out, s.d_recs, total * sizeof(raftgpu_send_entry), cudaMemcpyDeviceToHost));
    return RAFTGPU_OK;
}

int32_t raftgpu_tally_votes(raftgpu_arena *a, void *stream, uint32_t first, uint32_t n, uint32_t *d_out) {
    if (!a || !d_out) return RAFTGPU_ERR_INVALID;
    if (static_cast<uint64_t>(first) + n > a->cap) return RAFTGPU_ERR_RANGE;
    if (n == 0) return RAFTGPU_OK;
    CK(a, cudaSetDevice(a->device));
    tally_kernel<<<div_up(n, 256), 256, 0, pick_stream(a, stream)>>>(a->cols, first, n, d_out, a->d_counters);
    CKL(a);
    return RAFTGPU_OK;
}

int32_t raftgpu_vote_result(raftgpu_arena *a, uint32_t g, int32_t *out_result, uint32_t *out_granted,
                            uint32_t *out_rejected) {
    if (!a) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    if (!group_ok(a, g)) return RAFTGPU_ERR_RANGE;
    CK(a, cudaSetDevice(a->device));
    uint32_t *d = static_cast<uint32_t *>(a->d_scratch) + 16;  // offset 64
    tally_kernel<<<1, 32, 0, a->s_compute>>>(a->cols, g, 1, d - g, a->d_counters);
    CKL(a);
    uint8_t *hs = static_cast<uint8_t *>(a->h_scratch);
    CK(a, cudaMemcpyAsync(hs + 64, d, 4, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    const uint32_t w = *reinterpret_cast<uint32_t *>(hs + 64);
    if (out_result) *out_result = w & 0xff;
    if (out_granted) *out_granted = (w >> 8) & 0xff;
    if (out_rejected) *out_rejected = (w >> 16) & 0xff;
    return RAFTGPU_OK;
}

// ---- plumbing ---------------------------------------------------------------

int32_t raftgpu_counters_read(raftgpu_arena *a, raftgpu_counters *out) {
    if (!a || !out) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    CK(a, cudaSetDevice(a->device));
    static_assert(sizeof(raftgpu_counters) == kCntCount * 8, "counter layout");
    CK(a, cudaMemcpyAsync(a->h_scratch, a->d_counters, sizeof(*out), cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    memcpy(out, a->h_scratch, sizeof(*out));
    return RAFTGPU_OK;
}

int32_t raftgpu_debug_read(raftgpu_arena *a, uint64_t *out8) {
    if (!a || !out8) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    CK(a, cudaSetDevice(a->device));
    CK(a, cudaMemcpyAsync(a->h_scratch, a->d_counters + kCntCount, 64, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    memcpy(out8, a->h_scratch, 64);
    return RAFTGPU_OK;
}

int32_t raftgpu_synchronize(raftgpu_arena *a) {
    if (!a) return RAFTGPU_ERR_INVALID;
    CK(a, cudaSetDevice(a->device));
    CK(a, cudaDeviceSynchronize());
    return RAFTGPU_OK;
}

int32_t raftgpu_device_alloc(raftgpu_arena *a, uint64_t bytes, void **out) {
    if (!a || !out) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    CK(a, cudaSetDevice(a->device));
    uint8_t *p = nullptr;
    int32_t rc = dev_alloc(a, &p, bytes, false);
    if (rc != RAFTGPU_OK) return rc;
    a->user_allocs.push_back(p);
    *out = p;
    return RAFTGPU_OK;
}

int32_t raftgpu_device_free(raftgpu_arena *a, void *p) {
    if (!a || !p) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    for (size_t i = 0; i < a->user_allocs.size(); i++) {
        if (a->user_allocs[i] == p) {
            a->user_allocs.erase(a->user_allocs.begin() + i);
            CK(a, cudaSetDevice(a->device));
            CK(a, cudaFree(p));
            return RAFTGPU_OK;
        }
    }
    return RAFTGPU_ERR_INVALID;
}

int32_t raftgpu_memcpy_h2d(raftgpu_arena *a, void *dst, const void *src, uint64_t bytes) {
    if (!a || !dst || !src) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    CK(a, cudaSetDevice(a->device));
    CK(a, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    return RAFTGPU_OK;
}

int32_t raftgpu_memcpy_d2h(raftgpu_arena *a, void *dst, const void *src, uint64_t bytes) {
    if (!a || !dst || !src) return RAFTGPU_ERR_INVALID;
    std::lock_guard<std::mutex> ctl_lock(a->ctl_mu);
    CK(a, cudaSetDevice(a->device));
    CK(a, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, a->s_compute));
    CK(a, cudaStreamSynchronize(a->s_compute));
    return RAFTGPU_OK;
}

}  // extern "C"
