// kernels.cuh -- sm_100a device code of the batched commit-index engine.
//
// Every kernel here is HBM-bound u64 index arithmetic over the SoA arena
// (DESIGN.md); there is no floating point and no tensor-core work on this path.
// Reference semantics (file:line relative to the raft-rs checkout) are cited at
// each step; bit-exactness against oracle/raft_oracle.c is the contract.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "raftgpu.h"

namespace raftgpu {

constexpr int kSlots = RAFTGPU_SLOTS;

// Device view of the arena: per-peer columns are [kSlots][cap], per-group [cap].
struct Columns {
    uint32_t cap;
    uint64_t *matched;
    uint64_t *next_idx;
    uint64_t *peer_committed;
    uint64_t *pending_snapshot;
    uint64_t *pending_req_snapshot;
    uint64_t *commit_group_id;
    uint8_t *pflags;
    uint8_t *votes;
    uint32_t *meta;
    uint64_t *committed;
    uint64_t *term_start;
    uint64_t *last_index;
};

enum Counter : int {
    kCntRecomputes = 0,
    kCntAdvanced,
    kCntRecords,
    kCntUpdates,
    kCntRejects,
    kCntDecrements,
    kCntNoProgress,
    kCntVotes,
    kCntCount
};

// Fire-and-forget L2 prefetch: costs no destination register, so it deepens the memory pipeline
// beyond what registers x occupancy allow (the kernels here are long-scoreboard bound).
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------
// MajorityConfig::committed_index without group commit (majority.rs:70-101):
// the q-th largest acked index of the voters in `mask`, q = n/2 + 1
// (util.rs:118-120); the empty config yields u64::MAX (majority.rs:71-75).

// compare-exchange, larger value first
__device__ __forceinline__ void cex(uint64_t &a, uint64_t &b) {
    const bool lt = a < b;
    const uint64_t hi = lt ? b : a, lo = lt ? a : b;
    a = hi;
    b = lo;
}

// General form.  Non-members are zeroed, which leaves the top-q ranks of the
// members intact (q <= n); a 19-comparator network sorts the 8 slots in
// descending order (the reference's stable sort_by, majority.rs:95 -- ties are
// equal values, so any order of them selects the same index) and the q-th
// element is picked.
__device__ __forceinline__ uint64_t quorum_index(const uint64_t (&v)[kSlots], uint32_t mask) {
    if (mask == 0) return UINT64_MAX;
    const uint32_t q = (static_cast<uint32_t>(__popc(mask)) >> 1) + 1;
    uint64_t w0 = (mask & 1u) ? v[0] : 0, w1 = (mask & 2u) ? v[1] : 0, w2 = (mask & 4u) ? v[2] : 0,
             w3 = (mask & 8u) ? v[3] : 0, w4 = (mask & 16u) ? v[4] : 0, w5 = (mask & 32u) ? v[5] : 0,
             w6 = (mask & 64u) ? v[6] : 0, w7 = (mask & 128u) ? v[7] : 0;
    // Batcher / optimal 19-comparator network for 8 inputs
    cex(w0, w1); cex(w2, w3); cex(w4, w5); cex(w6, w7);
    cex(w0, w2); cex(w1, w3); cex(w4, w6); cex(w5, w7);
    cex(w1, w2); cex(w5, w6); cex(w0, w4); cex(w3, w7);
    cex(w1, w5); cex(w2, w6);
    cex(w1, w4); cex(w3, w6);
    cex(w2, w4); cex(w3, w5);
    cex(w3, w4);
    // q in 1..5 for up to 8 voters
    uint64_t r = w0;
    r = q == 2 ? w1 : r;
    r = q == 3 ? w2 : r;
    r = q == 4 ? w3 : r;
    r = q == 5 ? w4 : r;
    return r;
}

// The common 5-voter case (q = 3): the median, by the classic 10 min/max form
// med5(a..e) = med3(e, max(min(a,b),min(c,d)), min(max(a,b),max(c,d))).
__device__ __forceinline__ uint64_t median5(uint64_t a, uint64_t b, uint64_t c, uint64_t d,
                                            uint64_t e) {
    const uint64_t lo = umax64(umin64(a, b), umin64(c, d));
    const uint64_t hi = umin64(umax64(a, b), umax64(c, d));
    return umax64(umin64(lo, hi), umin64(umax64(lo, hi), e));
}

// MajorityConfig::committed_index WITH group commit (majority.rs:70-124), the
// literal algorithm: gather, stable descending sort, then the scan of :102-123.
// Rare path (ProgressTracker::group_commit is off by default), kept out of line
// so its local arrays do not cost the common path registers.
__device__ __noinline__ void majority_group_commit(const uint64_t *v, const uint64_t *gid,
                                                   uint32_t mask, uint64_t *out_index,
                                                   bool *out_use_gc) {
    if (mask == 0) {  // :71-75
        *out_index = UINT64_MAX;
        *out_use_gc = true;
        return;
    }
    uint64_t idx[kSlots], grp[kSlots];
    int n = 0;
    for (int s = 0; s < kSlots; s++) {
        if ((mask >> s) & 1u) {  // :77-85 (voters without progress do not occur in a tracker)
            idx[n] = v[s];
            grp[n] = gid[s];
            n++;
        }
    }
    for (int i = 1; i < n; i++) {  // :95 stable sort, descending by index
        uint64_t xi = idx[i], xg = grp[i];
        int j = i;
        while (j > 0 && idx[j - 1] < xi) {
            idx[j] = idx[j - 1];
            grp[j] = grp[j - 1];
            j--;
        }
        idx[j] = xi;
        grp[j] = xg;
    }
    const int quorum = n / 2 + 1;  // :97
    const uint64_t quorum_commit_index = idx[quorum - 1];
    uint64_t checked_group_id = grp[quorum - 1];
    bool single_group = true;
    for (int i = 0; i < n; i++) {  // :105-118
        if (grp[i] == 0) {
            single_group = false;
            continue;
        }
        if (checked_group_id == 0) {
            checked_group_id = grp[i];
            continue;
        }
        if (checked_group_id == grp[i]) continue;
        *out_index = umin64(idx[i], quorum_commit_index);
        *out_use_gc = true;
        return;
    }
    *out_index = single_group ? quorum_commit_index : idx[n - 1];  // :119-123
    *out_use_gc = false;
}

// ProgressTracker::maximal_committed_index (tracker.rs:294-298) of group g:
// JointConfig::committed_index (joint.rs:47-51) over both majority halves, reading
// matched / commit_group_id through the ProgressMap AckedIndexer (tracker.rs:183-190).
__device__ __forceinline__ void group_mci(const Columns &c, uint32_t g, uint32_t meta, uint64_t &mci,
                                          bool &use_gc) {
    const uint32_t in = RAFTGPU_META_IN(meta), out = RAFTGPU_META_OUT(meta);
    const uint32_t voters = in | out;
    uint64_t v[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; s++)
        v[s] = ((voters >> s) & 1u) ? c.matched[static_cast<size_t>(s) * c.cap + g] : 0ull;
    if (!(meta & RAFTGPU_META_GROUP_COMMIT)) {
        const uint64_t i_idx = quorum_index(v, in);
        const uint64_t o_idx = quorum_index(v, out);  // empty outgoing => u64::MAX
        mci = umin64(i_idx, o_idx);                    // joint.rs:50
        // a non-empty half reports false (majority.rs:99-101), an empty one true (:71-75)
        use_gc = (in == 0) && (out == 0);
    } else {
        uint64_t gid[kSlots];
        for (int s = 0; s < kSlots; s++)
            gid[s] = ((voters >> s) & 1u) ? c.commit_group_id[static_cast<size_t>(s) * c.cap + g] : 0ull;
        uint64_t i_idx, o_idx;
        bool i_gc, o_gc;
        majority_group_commit(v, gid, in, &i_idx, &i_gc);
        majority_group_commit(v, gid, out, &o_idx, &o_gc);
        mci = umin64(i_idx, o_idx);
        use_gc = i_gc && o_gc;  // joint.rs:50
    }
}

// Side-effect-free single-group query (thread 0 of one warp).
__global__ void mci_kernel(Columns c, uint32_t g, uint64_t *out_mci, uint8_t *out_gc) {
    if (threadIdx.x != 0) return;
    uint64_t mci;
    bool use_gc;
    group_mci(c, g, c.meta[g], mci, use_gc);
    *out_mci = mci;
    *out_gc = use_gc ? 1 : 0;
}

// Block-level counter flush: per-thread tallies -> warp shuffle reduce -> shared
// -> ONE global atomic per counter per block.  (v1 issued one atomic per warp per
// counter; ~10^5 same-address atomics serialise in L2 and dominated both kernels.)
template <int kN>
__device__ __forceinline__ void block_flush_counts(const uint32_t (&local)[kN], const int (&which)[kN],
                                                   unsigned long long *counters,
                                                   uint32_t *extra_u32 /* nullable, gets local[1] */) {
    __shared__ uint32_t s_cnt[kN];
    if (threadIdx.x < kN) s_cnt[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kN; k++) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, local[k]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_cnt[k], w);
    }
    __syncthreads();
    if (threadIdx.x < kN && s_cnt[threadIdx.x]) {
        atomicAdd(&counters[which[threadIdx.x]], static_cast<unsigned long long>(s_cnt[threadIdx.x]));
        if (extra_u32 && threadIdx.x == 1) atomicAdd(extra_u32, s_cnt[1]);
    }
}

// ---------------------------------------------------------------------------
// The recompute pass: one Raft::maybe_commit (raft.rs:893-904) per group.
//   mci  = ProgressTracker::maximal_committed_index      tracker.rs:294-298
//        = min(incoming.committed_index, outgoing.committed_index)   joint.rs:47-51
//   if mci > committed && term(mci) == term              raft_log.rs:487-499
//        committed = mci; prs[self].update_committed      raft.rs:896-900
// term(mci) == term is the range test term_start <= mci <= last_index.
// Algorithmic bytes per group: 8K (matched) + 4 (meta) + 24 (committed,
// term_start, last_index) read, 8 written when advanced.
//
// Two feeds (LDG, TMA) x two specialisations.  kSimple5 = the host has verified
// from its mirror of the meta column that EVERY group in the range is the plain
// 5-voter configuration in slots 0..4 (no joint half, no group commit): the
// kernel then carries no mask logic and no general selection network, which
// roughly halves its instructions and registers.  The general form handles any
// configuration; `hint` is a superset guess of the voter slots in the range (the
// host keeps the union of all voter masks) so that the matched loads of the
// hinted slots are issued together with meta / committed / term_start /
// last_index -- ONE round trip to HBM instead of two.  Voter slots outside the
// hint are fetched after meta arrives: correct for any hint, fast for a tight one.

// maximal_committed_index of one group from its matched values v[].
template <bool kSimple5>
__device__ __forceinline__ void eval_mci(const Columns &c, uint32_t g, uint32_t meta, uint64_t (&v)[kSlots],
                                         uint32_t hint, uint64_t &mci, bool &use_gc) {
    if constexpr (kSimple5) {
        mci = median5(v[0], v[1], v[2], v[3], v[4]);  // 5 voters: q = 3 = the median
        use_gc = false;
    } else {
        const uint32_t in = RAFTGPU_META_IN(meta), out = RAFTGPU_META_OUT(meta);
        const uint32_t voters = in | out;
        const uint32_t missing = voters & ~hint;
        if (missing) {  // hint was too small for this group: second trip for the rest
#pragma unroll
            for (int s = 0; s < kSlots; s++)
                if ((missing >> s) & 1u) v[s] = c.matched[static_cast<size_t>(s) * c.cap + g];
        }
        if ((meta & (0xffffu | RAFTGPU_META_GROUP_COMMIT)) == 0x1fu) {
            mci = median5(v[0], v[1], v[2], v[3], v[4]);
            use_gc = false;
        } else if (!(meta & RAFTGPU_META_GROUP_COMMIT)) {
            const uint64_t i_idx = quorum_index(v, in);
            const uint64_t o_idx = quorum_index(v, out);  // empty outgoing => u64::MAX
            mci = umin64(i_idx, o_idx);                    // joint.rs:50
            use_gc = (in == 0) && (out == 0);              // majority.rs:71-75 vs :99-101
        } else {
            uint64_t gid[kSlots];
            for (int s = 0; s < kSlots; s++)
                gid[s] = ((voters >> s) & 1u) ? c.commit_group_id[static_cast<size_t>(s) * c.cap + g] : 0ull;
            uint64_t i_idx, o_idx;
            bool i_gc, o_gc;
            majority_group_commit(v, gid, in, &i_idx, &i_gc);
            majority_group_commit(v, gid, out, &o_idx, &o_gc);
            mci = umin64(i_idx, o_idx);
            use_gc = i_gc && o_gc;
        }
    }
}

// RaftLog::maybe_commit (raft_log.rs:487-499, range form) + raft.rs:896-900.
__device__ __forceinline__ bool commit_group(const Columns &c, uint32_t g, uint32_t meta, uint64_t mci,
                                             bool use_gc, uint64_t committed, uint64_t term_start,
                                             uint64_t last_index, uint64_t *commit_out, uint64_t *mci_out,
                                             uint8_t *gc_out) {
    if (mci_out) mci_out[g] = mci;
    if (gc_out) gc_out[g] = use_gc ? 1 : 0;
    const bool advanced = mci > committed && mci >= term_start && mci <= last_index;
    if (advanced) {
        c.committed[g] = mci;  // commit_to: mci <= last_index, never the fatal! branch
        if (commit_out) commit_out[g] = mci;
        if (meta & RAFTGPU_META_HAS_SELF) {  // raft.rs:896-900
            const size_t cell = static_cast<size_t>(RAFTGPU_META_SELF(meta)) * c.cap + g;
            if (mci > c.peer_committed[cell]) c.peer_committed[cell] = mci;
        }
    }
    return advanced;
}

// One word of the advanced bitmap per warp-tile.
__device__ __forceinline__ void publish_tile(uint32_t *adv_bitmap, uint64_t g64, uint32_t lane, bool active,
                                             bool advanced, uint32_t (&local)[2]) {
    const unsigned act = __ballot_sync(0xffffffffu, active);
    const unsigned adv = __ballot_sync(0xffffffffu, advanced);
    if (lane == 0 && act != 0 && adv_bitmap) {
        uint32_t *word = &adv_bitmap[g64 >> 5];
        if (act == 0xffffffffu) {
            *word = adv;
        } else {  // range starts / ends inside this word: leave the other bits alone
            atomicAnd(word, ~act);
            if (adv) atomicOr(word, adv);
        }
    }
    local[0] += active ? 1u : 0u;
    local[1] += advanced ? 1u : 0u;
}

// ---- LDG feed: persistent grid, each warp walks 32-group tiles with a grid stride.
template <bool kSimple5, bool kPrefetch = false>
__global__ void __launch_bounds__(256, kSimple5 ? 6 : 4)
recompute_kernel(Columns c, uint32_t first, uint32_t n, uint32_t hint_arg,
                 uint32_t *__restrict__ adv_bitmap, uint64_t *__restrict__ commit_out,
                 uint64_t *__restrict__ mci_out, uint8_t *__restrict__ gc_out,
                 uint32_t *__restrict__ step_advanced, unsigned long long *__restrict__ counters) {
    const uint32_t hint = kSimple5 ? 0x1fu : hint_arg;
    const uint32_t base = first & ~31u;
    const uint32_t n_tiles = static_cast<uint32_t>((static_cast<uint64_t>(first - base) + n + 31) >> 5);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t end = static_cast<uint64_t>(first) + n;
    uint32_t local[2] = {0, 0};  // recomputes, advanced

    for (uint32_t tile = warp; tile < n_tiles; tile += n_warps) {
        const uint64_t g64 = static_cast<uint64_t>(base) + (static_cast<uint64_t>(tile) << 5) + lane;
        const bool active = g64 >= first && g64 < end;
        const uint32_t g = static_cast<uint32_t>(g64);
        // pull this warp's NEXT tile into L2 while the current one is processed; each row of a tile
        // is 256 contiguous bytes = two 128-byte lines, so lanes 0..1 cover it (slots by lane / 2)
        if (kPrefetch && tile + n_warps < n_tiles) {
            const uint64_t gn = static_cast<uint64_t>(base) + (static_cast<uint64_t>(tile + n_warps) << 5);
            const uint32_t row = lane >> 1, half = (lane & 1u) * 16u;
            if (row < kSlots) {
                if ((hint >> row) & 1u) prefetch_l2(c.matched + static_cast<size_t>(row) * c.cap + gn + half);
            } else if (row == kSlots) {
                prefetch_l2(c.committed + gn + half);
            } else if (row == kSlots + 1) {
                prefetch_l2(c.term_start + gn + half);
            } else if (row == kSlots + 2) {
                prefetch_l2(c.last_index + gn + half);
            } else if (row == kSlots + 3 && half == 0) {
                prefetch_l2(c.meta + gn);
            }
        }
        bool advanced = false;
        if (active) {
            // one batch of independent loads
            const uint32_t meta = c.meta[g];
            uint64_t v[kSlots];
#pragma unroll
            for (int s = 0; s < kSlots; s++)
                v[s] = ((hint >> s) & 1u) ? c.matched[static_cast<size_t>(s) * c.cap + g] : 0ull;
            const uint64_t committed = c.committed[g];
            const uint64_t term_start = c.term_start[g];
            const uint64_t last_index = c.last_index[g];
            uint64_t mci;
            bool use_gc;
            eval_mci<kSimple5>(c, g, meta, v, hint, mci, use_gc);
            advanced = commit_group(c, g, meta, mci, use_gc, committed, term_start, last_index, commit_out,
                                    mci_out, gc_out);
        }
        publish_tile(adv_bitmap, g64, lane, active, advanced, local);
    }
    const int which[2] = {kCntRecomputes, kCntAdvanced};
    block_flush_counts<2>(local, which, counters, step_advanced);
}

// ---- TMA feed.
// The LDG feed is long-scoreboard bound: the bytes it keeps in flight are capped
// by registers x occupancy.  Here a producer warp streams whole column tiles into
// a ring of shared-memory stages with 1-D bulk copies (cp.async.bulk, SASS
// UBLKCP) that complete on an mbarrier, so up to ~200 KB per SM are in flight
// whatever the consumer warps are doing; the consumers only touch shared memory
// and write `committed` back with coalesced stores.
//
//   stage layout:  [rows][kTile] u64   rows = hinted matched slots (ascending),
//                                      then committed, term_start, last_index
//                  [kTile] u32         meta
//   full[stage]  : producer arms with expect_tx(bytes); the copies complete it
//   empty[stage] : one arrival per consumer warp releases the stage
constexpr int kTile = 512;               // groups per stage = consumer threads
constexpr int kTmaThreads = kTile + 32;  // + one producer warp
constexpr int kMaxStages = 12;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <bool kSimple5>
__global__ void __launch_bounds__(kTmaThreads, 1)
recompute_tma_kernel(Columns c, uint32_t first, uint32_t n, uint32_t hint_arg, int n_stages,
                     uint32_t *__restrict__ adv_bitmap, uint64_t *__restrict__ commit_out,
                     uint64_t *__restrict__ mci_out, uint8_t *__restrict__ gc_out,
                     uint32_t *__restrict__ step_advanced, unsigned long long *__restrict__ counters) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];

    const uint32_t hint = kSimple5 ? 0x1fu : hint_arg;
    const uint32_t n_hint = kSimple5 ? 5u : static_cast<uint32_t>(__popc(hint & 0xffu));
    const uint32_t rows = n_hint + 3;
    const uint32_t stage_bytes = rows * kTile * 8 + kTile * 4;
    const uint32_t base = first - (first % kTile);
    const uint64_t end = static_cast<uint64_t>(first) + n;
    const uint32_t n_tiles = static_cast<uint32_t>((end - base + kTile - 1) / kTile);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kTile / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t local[2] = {0, 0};  // recomputes, advanced
    if (warp == kTile / 32) {
        // ===== producer warp: lane 0 arms the stage, then one lane per row issues its copy =====
        // row -> source: rows [0, n_hint) = matched of the r-th hinted slot, then committed,
        // term_start, last_index, and row `rows` = meta (u32)
        const uint8_t *src_base = nullptr;
        uint32_t elem = 8;
        if (lane < n_hint) {
            uint32_t seen = 0;
            for (int s = 0; s < kSlots; s++) {
                if (!((hint >> s) & 1u)) continue;
                if (seen == lane)
                    src_base = reinterpret_cast<const uint8_t *>(c.matched + static_cast<size_t>(s) * c.cap);
                seen++;
            }
        } else if (lane == n_hint) {
            src_base = reinterpret_cast<const uint8_t *>(c.committed);
        } else if (lane == n_hint + 1) {
            src_base = reinterpret_cast<const uint8_t *>(c.term_start);
        } else if (lane == n_hint + 2) {
            src_base = reinterpret_cast<const uint8_t *>(c.last_index);
        } else if (lane == rows) {
            src_base = reinterpret_cast<const uint8_t *>(c.meta);
            elem = 4;
        }
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int st = it % n_stages;
            const uint32_t ph = (it / n_stages) & 1u;
            const uint64_t g0 = static_cast<uint64_t>(base) + static_cast<uint64_t>(tile) * kTile;
            uint32_t ng = static_cast<uint32_t>(end - g0 < kTile ? end - g0 : kTile);
            ng = (ng + 3u) & ~3u;  // 16-byte multiples for the u32 row; stays inside the padded stride
            if (lane == 0) {
                mbar_wait(&empty_bar[st], ph ^ 1u);  // fresh barrier: the parity-1 wait passes at once
                mbar_expect_tx(&full_bar[st], rows * ng * 8 + ng * 4);
            }
            __syncwarp();
            if (src_base) {
                uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
                tma_load_1d(sb + static_cast<size_t>(lane) * kTile * 8, src_base + g0 * elem, ng * elem,
                            &full_bar[st]);
            }
        }
    } else {
        // ===== consumers: one group per thread per tile =====
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int st = it % n_stages;
            const uint32_t ph = (it / n_stages) & 1u;
            const uint64_t g64 = static_cast<uint64_t>(base) + static_cast<uint64_t>(tile) * kTile + threadIdx.x;
            const bool active = g64 >= first && g64 < end;
            const uint32_t g = static_cast<uint32_t>(g64);
            const uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
            const uint64_t *row = reinterpret_cast<const uint64_t *>(sb) + threadIdx.x;
            mbar_wait(&full_bar[st], ph);
            // everything this thread needs from the stage, into registers
            const uint32_t meta =
                reinterpret_cast<const uint32_t *>(sb + static_cast<size_t>(rows) * kTile * 8)[threadIdx.x];
            uint64_t v[kSlots];
            uint32_t r = 0;
#pragma unroll
            for (int s = 0; s < kSlots; s++) {
                v[s] = 0;
                if ((hint >> s) & 1u) {
                    v[s] = row[static_cast<size_t>(r) * kTile];
                    r++;
                }
            }
            const uint64_t committed = row[static_cast<size_t>(n_hint) * kTile];
            const uint64_t term_start = row[static_cast<size_t>(n_hint + 1) * kTile];
            const uint64_t last_index = row[static_cast<size_t>(n_hint + 2) * kTile];
            // the stage can be refilled as soon as every lane of this warp has its values
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            bool advanced = false;
            if (active) {
                uint64_t mci;
                bool use_gc;
                eval_mci<kSimple5>(c, g, meta, v, hint, mci, use_gc);
                advanced = commit_group(c, g, meta, mci, use_gc, committed, term_start, last_index,
                                        commit_out, mci_out, gc_out);
            }
            publish_tile(adv_bitmap, g64, lane, active, advanced, local);
        }
    }
    const int which[2] = {kCntRecomputes, kCntAdvanced};
    block_flush_counts<2>(local, which, counters, step_advanced);
}

// ---------------------------------------------------------------------------
// Progress state helpers on a register copy of one cell.
struct Cell {
    uint64_t matched, next_idx;
    uint32_t flags;  // pflags byte
};

// progress.rs:75-80 reset_state: paused = false, pending_snapshot = 0, state, ins.reset()
__device__ __forceinline__ void reset_state(Cell &p, uint32_t state, uint64_t *pending_snapshot) {
    p.flags &= ~(RAFTGPU_PF_PAUSED | RAFTGPU_PF_INS_FULL | RAFTGPU_PF_STATE_MASK);
    p.flags |= state;
    // a plain store, not "if non-zero then clear": the cold column would otherwise cost a dependent
    // HBM read on every state transition (and, in the fused kernel, stall the tile's barrier)
    *pending_snapshot = 0;
}

// apply_kernel: the per-message prefix of Raft::handle_append_response
// (raft.rs:1663-1743) for one wave of records, one thread per record, persistent
// grid-stride loop.  Within a wave every (group, peer) cell is touched by at most
// one record, so threads never race on a cell and no atomics are needed on the
// columns.
//
// The loop is software-pipelined three deep, because the work is two dependent
// HBM round trips (the record names the cell; the cell decides the update):
//   iteration k issues   the record load of element k+2,
//                        the cell loads (meta, matched, next_idx, pflags,
//                        committed_index) of element k+1,
//   and computes / stores element k,
// so every load has a whole iteration to land.  (An extra L2-prefetch stage, kPrefetch, was
// measured and does not help: at ~12 MB in flight the kernel is limited by the DRAM
// efficiency of sector-granular scattered accesses, not by latency -- profiles/.)
// Algorithmic bytes per record: 24 (record) + RMW of matched, next_idx,
// committed_index (48) + flag byte and meta (~4) = 76.
struct RecRegs {
    uint64_t w0, index, commit;
};
struct CellRegs {
    uint32_t meta;
    uint32_t flags;
    uint64_t matched, next_idx, peer_committed;
};
// Where one cell's hot fields live: HBM (scatter kernel) or a shared-memory tile (fused kernel).
// The cold columns (pending_snapshot, pending_request_snapshot) are always addressed in HBM.
struct CellPtrs {
    uint64_t *matched, *next_idx, *peer_committed, *last_index;
    uint8_t *pflags;
};

// Packed staging record, 16 bytes: what raftgpu_enqueue_* writes into the pinned rings and the
// step path ships over PCIe (the public 24-byte raftgpu_append_resp stays the API; packing cuts
// the H2D bytes -- the end-to-end bottleneck -- and the apply kernel's record traffic by a third).
//   w0: [0,32) group  [32,35) slot  35 REJECT  36 LOCAL  37 EXT  38 WIDE  39 HAS_EXT  [40,64) delta
//   w1: m.index   (EXT: the payload)
// commit is carried as a 24-bit delta: index - commit for a message (a follower's commit never
// exceeds what it acknowledges), commit - index for a LOCAL record (0xFFFFFF = "no new
// last_index"); anything else sets WIDE and the exact value follows in an EXT record.
// EXT kinds (in the delta field): 1 = next_probe_index, 2 = request_snapshot, 3 = wide commit,
// 0 = padding.
struct PackedRec {
    uint64_t w0, w1;
};
constexpr uint64_t kPkReject = 1ull << 35, kPkLocal = 1ull << 36, kPkExt = 1ull << 37, kPkWide = 1ull << 38,
                   kPkHasExt = 1ull << 39;
constexpr uint32_t kPkNoCommit = 0xFFFFFFu;

template <bool kPacked>
__device__ __forceinline__ RecRegs load_rec(const void *recs_v, uint64_t i, uint64_t n) {
    RecRegs r;
    if (i >= n) {  // past the end: a no-op (EXT) record
        r.w0 = static_cast<uint64_t>(RAFTGPU_REC_EXT) << 40;
        r.index = 0;
        r.commit = 0;
        return r;
    }
    if constexpr (!kPacked) {
        const uint64_t *p = reinterpret_cast<const uint64_t *>(static_cast<const raftgpu_append_resp *>(recs_v) + i);
        r.w0 = p[0];
        r.index = p[1];
        r.commit = p[2];
    } else {
        const ulonglong2 q = reinterpret_cast<const ulonglong2 *>(recs_v)[i];  // one 128-bit load
        const uint64_t w0 = q.x;
        const uint32_t delta = static_cast<uint32_t>(w0 >> 40);
        const uint32_t flags = ((w0 & kPkReject) ? RAFTGPU_REC_REJECT : 0u) | ((w0 & kPkLocal) ? RAFTGPU_REC_LOCAL : 0u) |
                               ((w0 & kPkExt) ? RAFTGPU_REC_EXT : 0u);
        // the public layout: group | slot << 32 | flags << 40
        r.w0 = (w0 & 0xffffffffull) | (((w0 >> 32) & 7ull) << 32) | (static_cast<uint64_t>(flags) << 40);
        r.index = q.y;
        if (w0 & kPkLocal)
            r.commit = delta == kPkNoCommit ? 0 : q.y + delta;
        else
            r.commit = q.y - delta;
        if ((w0 & kPkWide) && !(w0 & kPkExt)) {  // rare: the exact commit follows in an EXT of kind 3
            for (uint64_t j = i + 1; j < n && j <= i + 3; j++) {
                const ulonglong2 e = reinterpret_cast<const ulonglong2 *>(recs_v)[j];
                if (!(e.x & kPkExt)) break;
                if ((e.x >> 40) == 3) r.commit = e.y;
            }
        }
    }
    return r;
}

// next_probe_index / request_snapshot of the REJECT at position i (raft.rs:1560-1661, 1709)
template <bool kPacked>
__device__ __forceinline__ void load_reject_ext(const void *recs_v, uint64_t i, uint64_t n, uint64_t &hint,
                                                uint64_t &request_snapshot) {
    hint = 0;
    request_snapshot = RAFTGPU_INVALID_INDEX;
    if constexpr (!kPacked) {
        if (i + 1 < n) {
            const uint64_t *e = reinterpret_cast<const uint64_t *>(static_cast<const raftgpu_append_resp *>(recs_v) + i + 1);
            if ((e[0] >> 40) & RAFTGPU_REC_EXT) {
                hint = e[1];
                request_snapshot = e[2];
            }
        }
    } else {
        for (uint64_t j = i + 1; j < n && j <= i + 3; j++) {
            const ulonglong2 e = reinterpret_cast<const ulonglong2 *>(recs_v)[j];
            if (!(e.x & kPkExt)) break;
            const uint32_t kind = static_cast<uint32_t>(e.x >> 40);
            if (kind == 1) hint = e.y;
            if (kind == 2) request_snapshot = e.y;
        }
    }
}

__device__ __forceinline__ CellPtrs global_cell_ptrs(const Columns &c, const RecRegs &r) {
    const uint32_t g = static_cast<uint32_t>(r.w0);
    const uint32_t slot = static_cast<uint32_t>(r.w0 >> 32) & 0xffu;
    const bool ok = g < c.cap && slot < kSlots;
    const size_t cell = ok ? static_cast<size_t>(slot) * c.cap + g : 0;
    return CellPtrs{c.matched + cell, c.next_idx + cell, c.peer_committed + cell, c.last_index + (ok ? g : 0),
                    c.pflags + cell};
}

// Stage 2 of the apply pipeline: pull the record's cell (and its group's meta word) into L2.
__device__ __forceinline__ void prefetch_cell(const Columns &c, const RecRegs &r) {
    const uint32_t g = static_cast<uint32_t>(r.w0);
    const uint32_t slot = static_cast<uint32_t>(r.w0 >> 32) & 0xffu;
    if (((r.w0 >> 40) & RAFTGPU_REC_EXT) || g >= c.cap || slot >= kSlots) return;
    const size_t cell = static_cast<size_t>(slot) * c.cap + g;
    prefetch_l2(c.matched + cell);
    prefetch_l2(c.next_idx + cell);
    prefetch_l2(c.peer_committed + cell);
    prefetch_l2(c.pflags + cell);
    prefetch_l2(c.meta + g);
}

__device__ __forceinline__ CellRegs load_cell(const Columns &c, const RecRegs &r) {
    const uint32_t g = static_cast<uint32_t>(r.w0);
    const uint32_t slot = static_cast<uint32_t>(r.w0 >> 32) & 0xffu;
    // cell 0 / group 0 is a harmless stand-in for EXT and out-of-range records
    // (nothing is written for them)
    const bool ok = !((r.w0 >> 40) & RAFTGPU_REC_EXT) && g < c.cap && slot < kSlots;
    const size_t cell = ok ? static_cast<size_t>(slot) * c.cap + g : 0;
    CellRegs d;
    d.meta = c.meta[ok ? g : 0];
    d.matched = c.matched[cell];
    d.next_idx = c.next_idx[cell];
    d.flags = c.pflags[cell];
    d.peer_committed = c.peer_committed[cell];
    return d;
}

// One record against its cell: raft.rs:1663-1743.  Returns the result byte.
// kFmt: 0 = records in the public 24-byte layout, 1 = packed 16-byte records (a REJECT's EXT is
// looked up behind position i in either), 2 = the caller has decoded them already (compact
// streams) and hands them in through the two integer arguments: n = next_probe_index hint,
// i = request_snapshot (recs unused) -- registers only, no stack traffic.
template <int kFmt>
__device__ __forceinline__ uint32_t apply_one(const Columns &c, const void *recs, uint64_t n, uint64_t i,
                                              const RecRegs &rec, const CellRegs &cd, const CellPtrs &ptr,
                                              uint32_t *local) {
    const uint64_t index = rec.index, commit = rec.commit;
    const uint32_t g = static_cast<uint32_t>(rec.w0);
    const uint32_t slot = static_cast<uint32_t>(rec.w0 >> 32) & 0xffu;
    const uint32_t rflags = static_cast<uint32_t>(rec.w0 >> 40) & 0xffu;
    if (rflags & RAFTGPU_REC_EXT) return 0;
    local[0]++;
    const bool in_range = g < c.cap && slot < kSlots;
    const uint32_t present =
        RAFTGPU_META_IN(cd.meta) | RAFTGPU_META_OUT(cd.meta) | RAFTGPU_META_LEARN(cd.meta);
    if (!in_range || !((present >> slot) & 1u)) {
        // raft.rs:1663-1673: no progress available for m.from
        local[4]++;
        return RAFTGPU_RES_NO_PROGRESS;
    }
    const size_t cell = static_cast<size_t>(slot) * c.cap + g;
    Cell pr;
    pr.matched = cd.matched;
    pr.next_idx = cd.next_idx;
    pr.flags = cd.flags;
    const uint32_t state = pr.flags & RAFTGPU_PF_STATE_MASK;
    uint32_t res = 0;

    if (rflags & RAFTGPU_REC_LOCAL) {
        // raft.rs:974-991 append_entry: last_index grew
        if (commit != 0) *ptr.last_index = commit;
        // raft.rs:1010-1014 on_persist_entries: prs[self].maybe_update(index)
        if (pr.matched < index) {  // progress.rs:138-150
            pr.matched = index;
            pr.flags &= ~RAFTGPU_PF_PAUSED;
            local[1]++;
            res = RAFTGPU_RES_OK;
        }
        if (pr.next_idx < index + 1) pr.next_idx = index + 1;
    } else {
        pr.flags |= RAFTGPU_PF_RECENT_ACTIVE;  // raft.rs:1674
        // raft.rs:1677 pr.update_committed(m.commit), progress.rs:153-157
        if (commit > cd.peer_committed) *ptr.peer_committed = commit;

        if (rflags & RAFTGPU_REC_REJECT) {
            local[2]++;
            uint64_t hint, request_snapshot;
            if constexpr (kFmt == 2) {
                hint = n;
                request_snapshot = i;
            } else {
                load_reject_ext<(kFmt == 1)>(recs, i, n, hint, request_snapshot);
            }
            // Progress::maybe_decr_to, progress.rs:168-206
            bool ok;
            if (state == RAFTGPU_STATE_REPLICATE) {
                if (index < pr.matched || (index == pr.matched && request_snapshot == RAFTGPU_INVALID_INDEX)) {
                    ok = false;  // :173-177 stale
                } else {
                    if (request_snapshot == RAFTGPU_INVALID_INDEX)
                        pr.next_idx = pr.matched + 1;  // :178-179
                    else
                        c.pending_req_snapshot[cell] = request_snapshot;  // :181
                    ok = true;
                }
            } else if ((pr.next_idx == 0 || pr.next_idx - 1 != index) &&
                       request_snapshot == RAFTGPU_INVALID_INDEX) {
                ok = false;  // :188-192 stale
            } else {
                if (request_snapshot == RAFTGPU_INVALID_INDEX) {  // :195-199
                    pr.next_idx = umin64(index, hint + 1);
                    if (pr.next_idx < 1) pr.next_idx = 1;
                } else if (c.pending_req_snapshot[cell] == RAFTGPU_INVALID_INDEX) {
                    c.pending_req_snapshot[cell] = request_snapshot;  // :200-203
                }
                pr.flags &= ~RAFTGPU_PF_PAUSED;  // :204 resume()
                ok = true;
            }
            if (ok) {
                local[3]++;
                res = RAFTGPU_RES_OK | RAFTGPU_RES_SEND;
                if (state == RAFTGPU_STATE_REPLICATE) {
                    // raft.rs:1716-1718 become_probe (progress.rs:95-107, not Snapshot)
                    reset_state(pr, RAFTGPU_STATE_PROBE, &c.pending_snapshot[cell]);
                    pr.next_idx = pr.matched + 1;
                }
            }
        } else {
            // raft.rs:1724 old_paused = pr.is_paused(), progress.rs:210-216
            const bool old_paused =
                state == RAFTGPU_STATE_PROBE
                    ? (pr.flags & RAFTGPU_PF_PAUSED) != 0
                    : (state == RAFTGPU_STATE_REPLICATE ? (pr.flags & RAFTGPU_PF_INS_FULL) != 0 : true);
            // raft.rs:1725 pr.maybe_update(m.index), progress.rs:138-150
            const bool need_update = pr.matched < index;
            if (need_update) {
                pr.matched = index;
                pr.flags &= ~RAFTGPU_PF_PAUSED;
            }
            if (pr.next_idx < index + 1) pr.next_idx = index + 1;
            if (need_update) {
                local[1]++;
                res = RAFTGPU_RES_OK | (old_paused ? RAFTGPU_RES_OLD_PAUSED : 0u);
                if (state == RAFTGPU_STATE_PROBE) {
                    // raft.rs:1730 become_replicate, progress.rs:110-114
                    reset_state(pr, RAFTGPU_STATE_REPLICATE, &c.pending_snapshot[cell]);
                    pr.next_idx = pr.matched + 1;
                } else if (state == RAFTGPU_STATE_SNAPSHOT) {
                    // raft.rs:1731-1741 maybe_snapshot_abort -> become_probe
                    const uint64_t pending = c.pending_snapshot[cell];
                    if (pr.matched >= pending) {  // progress.rs:131-134
                        reset_state(pr, RAFTGPU_STATE_PROBE, &c.pending_snapshot[cell]);
                        pr.next_idx = umax64(pr.matched + 1, pending + 1);  // :99-102
                    }
                }
                // Replicate: pr.ins.free_to(m.index) -- Inflights stays host-side
            }
        }
    }
    if (pr.matched != cd.matched) *ptr.matched = pr.matched;
    if (pr.next_idx != cd.next_idx) *ptr.next_idx = pr.next_idx;
    if (pr.flags != cd.flags) *ptr.pflags = static_cast<uint8_t>(pr.flags);
    return res;
}

// kCheckDup (zero-copy submissions, where no host code has seen the records): every record marks
// its cell in `touched` ([cap] bytes, one bit per peer slot, cleared by the caller beforehand)
// with an L2 atomic; a cell marked twice breaks the one-wave precondition -- the record is NOT
// applied and *dup_count is bumped so the step can fail loudly.
template <bool kPacked, bool kCheckDup = false, bool kPrefetch = false>
__global__ void __launch_bounds__(256, 4)
apply_kernel(Columns c, const void *__restrict__ recs, uint64_t n, uint8_t *__restrict__ results,
             unsigned long long *__restrict__ counters, uint32_t *__restrict__ touched = nullptr,
             uint32_t *__restrict__ dup_count = nullptr) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint32_t local[5] = {0, 0, 0, 0, 0};  // records, updates, rejects, decrements, no_progress
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    // prologue: fill the pipeline
    RecRegs rec_a = load_rec<kPacked>(recs, i, n);
    RecRegs rec_b = load_rec<kPacked>(recs, i + stride, n);
    CellRegs cell_a = load_cell(c, rec_a);
    for (; i < n; i += stride) {
        const RecRegs rec_c = load_rec<kPacked>(recs, i + 2 * stride, n);  // element k+2: record
        if (kPrefetch) prefetch_cell(c, rec_c);                    // (optional) its cell -> L2
        const CellRegs cell_b = load_cell(c, rec_b);               // element k+1: cell -> registers
        if constexpr (kCheckDup) {
            const uint32_t g = static_cast<uint32_t>(rec_a.w0), slot = static_cast<uint32_t>(rec_a.w0 >> 32) & 0xffu;
            if (!((rec_a.w0 >> 40) & RAFTGPU_REC_EXT) && g < c.cap && slot < kSlots) {
                const uint32_t bit = 1u << (8 * (g & 3u) + slot);
                if (atomicOr(&touched[g >> 2], bit) & bit) {  // second record for this cell in one wave
                    atomicAdd(dup_count, 1u);
                    if (results) results[i] = 0;
                    rec_a = rec_b;
                    cell_a = cell_b;
                    rec_b = rec_c;
                    continue;
                }
            }
        }
        const CellPtrs gp = global_cell_ptrs(c, rec_a);
        const uint32_t res = apply_one<(kPacked ? 1 : 0)>(c, recs, n, i, rec_a, cell_a, gp, local);  // element k
        if (results) results[i] = static_cast<uint8_t>(res);
        rec_a = rec_b;
        cell_a = cell_b;
        rec_b = rec_c;
    }
    const int which[5] = {kCntRecords, kCntUpdates, kCntRejects, kCntDecrements, kCntNoProgress};
    block_flush_counts<5>(local, which, counters, nullptr);
}

// ---------------------------------------------------------------------------
// The compact stream (raftgpu.h "compact stream"): 4-byte units, group runs with a header.
constexpr uint32_t kCuRec = 0, kCuHdrA = 1, kCuHdrB = 2, kCuEsc = 3;
constexpr uint32_t kCuLocal = 4u, kCuReject = 1u << 9, kCuNoCommit = 255u;
constexpr uint32_t kCuPayload = 1u << 29;   // in the ESC field: a REJECT's hint rides here, not a side index
constexpr uint32_t kCuPad = 0x1fffffffu;    // ESC field value of a padding unit (side indexes stay below it)
struct CompactSrc {
    const uint32_t *units;
    const uint32_t *g_base;              // one per block of RAFTGPU_COMPACT_BLOCK units
    const raftgpu_append_resp *side;     // ESC targets, public layout (a REJECT is followed by its EXT)
    uint32_t n_units, n_side;
};

// A REJECT's payload unit (the one behind it): hint = index + signed 29-bit delta.
__device__ __forceinline__ uint64_t compact_hint(uint64_t index, uint32_t payload_unit) {
    const int32_t d = static_cast<int32_t>(payload_unit << 1) >> 3;  // bits [2,31), sign-extended
    return index + static_cast<uint64_t>(static_cast<int64_t>(d));
}

// Unit i as a record in the public register layout.  Headers, payload units, padding and
// malformed units come back as EXT (a no-op).
__device__ __forceinline__ RecRegs load_compact(const CompactSrc &s, uint64_t i) {
    RecRegs r;
    r.w0 = static_cast<uint64_t>(RAFTGPU_REC_EXT) << 40;
    r.index = 0;
    r.commit = 0;
    if (i >= s.n_units) return r;
    const uint32_t u = s.units[i];
    const uint32_t kind = u & 3u;
    if (kind == kCuEsc) {
        const uint32_t k = u >> 2;
        if (k < s.n_side && k < kCuPad) {
            const uint64_t *p = reinterpret_cast<const uint64_t *>(s.side + k);
            r.w0 = p[0];
            r.index = p[1];
            r.commit = p[2];
        }
        return r;
    }
    if (kind != kCuRec) return r;
    const uint32_t back = (u >> 3) & 7u;
    if (i < back + 2u) return r;
    const uint64_t h = i - back - 2u;
    const uint32_t ha = s.units[h], hb = s.units[h + 1];
    if ((ha & 3u) != kCuHdrA || (hb & 3u) != kCuHdrB) return r;
    const uint32_t g = s.g_base[h / RAFTGPU_COMPACT_BLOCK] + ((hb >> 2) & 0xfffu);
    const uint64_t base = static_cast<uint64_t>(ha >> 2) | (static_cast<uint64_t>(hb >> 14) << 30);
    const uint32_t slot = (u >> 6) & 7u;
    const uint64_t index = base + ((u >> 10) & 0x3fffu);
    const uint32_t cd = u >> 24;
    const bool local = (u & kCuLocal) != 0;
    uint32_t flags = local ? RAFTGPU_REC_LOCAL : 0u;
    r.index = index;
    r.commit = local ? (cd == kCuNoCommit ? 0 : index + cd) : (index >= cd ? index - cd : 0);
    if (u & kCuReject) flags = RAFTGPU_REC_REJECT;
    r.w0 = static_cast<uint64_t>(g) | (static_cast<uint64_t>(slot) << 32) | (static_cast<uint64_t>(flags) << 40);
    return r;
}

// {next_probe_index hint, request_snapshot} of the REJECT that unit i decoded to (rare: looked up
// when the record is applied, not carried through the pipeline).
__device__ __forceinline__ void compact_reject_ext(const CompactSrc &s, uint64_t i, uint64_t index, uint64_t (&ext)[2]) {
    ext[0] = 0;
    ext[1] = RAFTGPU_INVALID_INDEX;
    const uint32_t u = s.units[i];
    if ((u & 3u) == kCuEsc) {
        load_reject_ext<false>(s.side, u >> 2, s.n_side, ext[0], ext[1]);
    } else {
        const uint32_t pl = i + 1 < s.n_units ? s.units[i + 1] : 0u;
        if ((pl & 3u) == kCuEsc && ((pl >> 2) & kCuPayload)) ext[0] = compact_hint(index, pl);
    }
}

// apply_kernel for a compact stream: the same three-deep pipeline, one thread per UNIT (header
// units idle).  A run's header sits in the cache lines its records' neighbours load, so the
// decode adds L1/L2 hits, not HBM trips.  REJECTs only arrive through ESC units, whose payload
// (and EXT continuation) is read from the side table in the public layout.
template <bool kCheckDup>
__global__ void __launch_bounds__(256, 4)
apply_compact_kernel(Columns c, CompactSrc src, uint8_t *__restrict__ results, unsigned long long *__restrict__ counters,
                     uint32_t *__restrict__ touched, uint32_t *__restrict__ dup_count) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t n = src.n_units;
    uint32_t local[5] = {0, 0, 0, 0, 0};  // records, updates, rejects, decrements, no_progress
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    RecRegs rec_a = load_compact(src, i);
    RecRegs rec_b = load_compact(src, i + stride);
    CellRegs cell_a = load_cell(c, rec_a);
    for (; i < n; i += stride) {
        const RecRegs rec_c = load_compact(src, i + 2 * stride);
        const CellRegs cell_b = load_cell(c, rec_b);
        bool skip = false;
        if constexpr (kCheckDup) {
            const uint32_t g = static_cast<uint32_t>(rec_a.w0), slot = static_cast<uint32_t>(rec_a.w0 >> 32) & 0xffu;
            if (!((rec_a.w0 >> 40) & RAFTGPU_REC_EXT) && g < c.cap && slot < kSlots) {
                const uint32_t bit = 1u << (8 * (g & 3u) + slot);
                if (atomicOr(&touched[g >> 2], bit) & bit) {  // second record for this cell in one wave
                    atomicAdd(dup_count, 1u);
                    skip = true;
                }
            }
        }
        uint32_t res = 0;
        if (!skip) {
            const CellPtrs gp = global_cell_ptrs(c, rec_a);
            uint64_t ext[2] = {0, RAFTGPU_INVALID_INDEX};
            if ((rec_a.w0 >> 40) & RAFTGPU_REC_REJECT) compact_reject_ext(src, i, rec_a.index, ext);
            res = apply_one<2>(c, nullptr, ext[0], ext[1], rec_a, cell_a, gp, local);
        }
        if (results) results[i] = static_cast<uint8_t>(res);
        rec_a = rec_b;
        cell_a = cell_b;
        rec_b = rec_c;
    }
    const int which[5] = {kCntRecords, kCntUpdates, kCntRejects, kCntDecrements, kCntNoProgress};
    block_flush_counts<5>(local, which, counters, nullptr);
}

// ---------------------------------------------------------------------------
// step_tile_kernel: apply + recompute FUSED, for batches whose records are in group order.
//
// The scatter apply kernel above moves ~12 MB in flight but tops out near 3.7 TB/s: its cell
// accesses are sector-granular (8 useful bytes per 32-byte sector request, several rows per
// record), which HBM serves at roughly half the efficiency of dense bursts.  When the batch is
// ordered by group, a tile of kFTile consecutive groups owns a CONTIGUOUS record range, so the
// whole step becomes dense traffic: per tile the producer warp bulk-loads (TMA, cp.async.bulk)
// the tile's rows of matched / next_idx / committed_index / pflags / meta / committed /
// term_start / last_index and its record range into one shared-memory stage; the consumers
//   A. apply every record of the tile to the shared-memory rows (one thread per record;
//      distinct cells per wave, so no conflicts),
//   B. recompute the commit index of the tile's groups from the same shared-memory rows
//      (matched is read from HBM once per step instead of twice),
//   C. bulk-store the rows back (cp.async.bulk.global.shared::cta).
// Peer slots outside `hint` (learners) and the cold columns are handled through HBM directly.
// tile_off[t] = index of the first packed record of tile t (raftgpu_tile_index builds it).
constexpr int kFTile = RAFTGPU_TILE_GROUPS;  // groups per tile
// consumer groups: tile i of a CTA is handled by group i % kNG, so the phases of kNG tiles overlap
// inside one CTA; each group has kCT threads (kCT >= kFTile).  Template parameters of the kernel.
constexpr int kFMaxStages = 8;
// Row strides inside a stage.  Records arrive in group order, so the ~3.5 records of one group sit
// in neighbouring lanes and touch the SAME column index of DIFFERENT rows: with a 256-element row
// stride they would all fall on the same shared-memory banks (4-way conflicts on every access).
// 258 u64 (= 2064 B, still 16-byte aligned for TMA) shifts consecutive rows by 4 banks.
constexpr uint32_t kFRow64 = (kFTile + 2) * 8;   // bytes per u64 row
constexpr uint32_t kFRow8 = kFTile + 16;         // bytes per u8 (pflags) row

__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
// 1-D TMA store: shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void tma_store_1d(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                 "r"(bytes)
                 : "memory");
}

struct TileArgs {
    const PackedRec *recs;     // packed records in group order
    const uint32_t *tile_off;  // [n_tiles + 1]
    uint32_t n_groups;         // groups [0, n_groups)
    uint32_t hint;
    int n_stages;
    uint32_t rec_cap;          // packed records staged per tile (multiple of 4); the rest is read from HBM
    uint8_t *results;          // nullable, one byte per packed record
    uint32_t *adv_bitmap;      // nullable
    uint64_t *commit_out;      // nullable
    uint32_t *step_advanced;   // nullable
    unsigned long long *counters;
    unsigned long long *dbg;   // nullable: [8] cycle totals per phase (diagnostics, RAFTGPU_TILE_DEBUG=1)
};

// bytes of one stage for H hinted slots (shared by host and device)
__host__ __device__ constexpr uint32_t tile_stage_bytes(uint32_t H, uint32_t rec_cap) {
    return 3u * H * kFRow64 + 3u * kFRow64 + kFTile * 4u + H * kFRow8 + rec_cap * 16u;
}

template <bool kSimple5, int kCT, int kNG>
__global__ void __launch_bounds__(kCT *kNG + 64, 1) step_tile_kernel(Columns c, TileArgs a) {
    static_assert(kCT >= kFTile && kCT % 32 == 0, "a consumer group covers a tile");
    const uint32_t kFRecCap = a.rec_cap;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kFMaxStages];   // loads landed            (load warp -> consumers)
    __shared__ __align__(8) uint64_t done_bar[kFMaxStages];   // rows final in smem      (consumers -> store warp)
    __shared__ __align__(8) uint64_t empty_bar[kFMaxStages];  // rows read by the stores (store warp -> load warp)

    const uint32_t hint = kSimple5 ? 0x1fu : (a.hint & 0xffu);
    const uint32_t H = kSimple5 ? 5u : static_cast<uint32_t>(__popc(hint));
    // stage layout (bytes)
    const uint32_t o_matched = 0, o_next = H * kFRow64, o_pc = 2u * H * kFRow64, o_committed = 3u * H * kFRow64,
                   o_ts = o_committed + kFRow64, o_li = o_ts + kFRow64, o_meta = o_li + kFRow64,
                   o_flags = o_meta + kFTile * 4u, o_recs = o_flags + H * kFRow8, stage_bytes = tile_stage_bytes(H, a.rec_cap);
    const uint32_t n_tiles = (a.n_groups + kFTile - 1) / kFTile;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&done_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t local[7] = {0, 0, 0, 0, 0, 0, 0};  // records, updates, rejects, decrements, no_progress | recomputes, advanced
    if (warp == kNG * kCT / 32 + 1) {
        // ===================== store warp: rows back to HBM, then the stage is free =====================
        const uint32_t n_out = 4u * H + 2u;
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int st = it % a.n_stages;
            const uint32_t ph = (it / a.n_stages) & 1u;
            const uint32_t g0 = tile * kFTile;
            const uint32_t ng = a.n_groups - g0 < kFTile ? a.n_groups - g0 : kFTile;
            const uint32_t ng16 = (ng + 15u) & ~15u;
            uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
            mbar_wait(&done_bar[st], ph);  // every lane waits: the barrier's completion orders the consumers' writes
            for (uint32_t j = lane; j < n_out; j += 32) {
                if (j < 4u * H) {
                    const uint32_t col = j / H, r = j % H;
                    uint32_t slot = 0, seen = 0;
                    for (uint32_t s2 = 0; s2 < kSlots; s2++)
                        if ((hint >> s2) & 1u) {
                            if (seen == r) slot = s2;
                            seen++;
                        }
                    const size_t cell = static_cast<size_t>(slot) * c.cap + g0;
                    if (col == 0) tma_store_1d(c.matched + cell, sb + o_matched + r * kFRow64, ng16 * 8u);
                    if (col == 1) tma_store_1d(c.next_idx + cell, sb + o_next + r * kFRow64, ng16 * 8u);
                    if (col == 2) tma_store_1d(c.peer_committed + cell, sb + o_pc + r * kFRow64, ng16 * 8u);
                    if (col == 3) tma_store_1d(c.pflags + cell, sb + o_flags + r * kFRow8, ng16);
                } else if (j == 4u * H) {
                    tma_store_1d(c.committed + g0, sb + o_committed, ng16 * 8u);
                } else {
                    tma_store_1d(c.last_index + g0, sb + o_li, ng16 * 8u);
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory has been read
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores have landed
    } else if (warp == kNG * kCT / 32) {
        // ===================== load warp =====================
        const uint32_t n_copies = 4u * H + 5u;
        uint32_t it = 0;
        // the tile index of the NEXT tile is fetched while this one is being issued: a dependent global
        // load at the top of every iteration would sit on the critical path of the ring
        uint32_t nx_beg = blockIdx.x < n_tiles ? a.tile_off[blockIdx.x] : 0u, nx_end = blockIdx.x < n_tiles ? a.tile_off[blockIdx.x + 1] : 0u;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int st = it % a.n_stages;
            const uint32_t ph = (it / a.n_stages) & 1u;
            const uint32_t g0 = tile * kFTile;
            const uint32_t ng = a.n_groups - g0 < kFTile ? a.n_groups - g0 : kFTile;
            const uint32_t ng16 = (ng + 15u) & ~15u;  // 16-byte multiples for every row; inside the padded stride
            const uint32_t rbeg = nx_beg, rend = nx_end;
            if (tile + gridDim.x < n_tiles) {
                nx_beg = a.tile_off[tile + gridDim.x];
                nx_end = a.tile_off[tile + gridDim.x + 1];
            }
            const uint32_t staged = rend - rbeg < kFRecCap ? rend - rbeg : kFRecCap;
            if (lane == 0) {
                mbar_wait(&empty_bar[st], ph ^ 1u);
                mbar_expect_tx(&full_bar[st], 3u * H * ng16 * 8u + H * ng16 + 3u * ng16 * 8u + ng16 * 4u + staged * 16u);
            }
            __syncwarp();
            uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
            for (uint32_t j = lane; j < n_copies; j += 32) {
                if (j < 4u * H) {
                    const uint32_t col = j / H, r = j % H;  // col: 0 matched, 1 next_idx, 2 committed_index, 3 pflags
                    uint32_t slot = 0, seen = 0;
                    for (uint32_t s2 = 0; s2 < kSlots; s2++)
                        if ((hint >> s2) & 1u) {
                            if (seen == r) slot = s2;
                            seen++;
                        }
                    const size_t cell = static_cast<size_t>(slot) * c.cap + g0;
                    if (col == 0) tma_load_1d(sb + o_matched + r * kFRow64, c.matched + cell, ng16 * 8u, &full_bar[st]);
                    if (col == 1) tma_load_1d(sb + o_next + r * kFRow64, c.next_idx + cell, ng16 * 8u, &full_bar[st]);
                    if (col == 2) tma_load_1d(sb + o_pc + r * kFRow64, c.peer_committed + cell, ng16 * 8u, &full_bar[st]);
                    if (col == 3) tma_load_1d(sb + o_flags + r * kFRow8, c.pflags + cell, ng16, &full_bar[st]);
                } else if (j == 4u * H) {
                    tma_load_1d(sb + o_committed, c.committed + g0, ng16 * 8u, &full_bar[st]);
                } else if (j == 4u * H + 1) {
                    tma_load_1d(sb + o_ts, c.term_start + g0, ng16 * 8u, &full_bar[st]);
                } else if (j == 4u * H + 2) {
                    tma_load_1d(sb + o_li, c.last_index + g0, ng16 * 8u, &full_bar[st]);
                } else if (j == 4u * H + 3) {
                    tma_load_1d(sb + o_meta, c.meta + g0, ng16 * 4u, &full_bar[st]);
                } else if (staged) {
                    tma_load_1d(sb + o_recs, a.recs + rbeg, staged * 16u, &full_bar[st]);
                }
            }
        }
    } else {
        // ===================== consumers: group cg takes every kFGroups-th tile of this CTA =====================
        const uint32_t cg = warp / (kCT / 32);
        const uint32_t tid = threadIdx.x - cg * kCT;
        const int bar_id = 1 + static_cast<int>(cg);
        constexpr uint32_t R64 = kFRow64 / 8;  // row stride in u64 elements
        uint32_t it = 0;
        const uint32_t first_tile = blockIdx.x + cg * gridDim.x, tile_step = kNG * gridDim.x;
        uint32_t nx_beg = first_tile < n_tiles ? a.tile_off[first_tile] : 0u, nx_end = first_tile < n_tiles ? a.tile_off[first_tile + 1] : 0u;
        it = cg;
        for (uint32_t tile = first_tile; tile < n_tiles; tile += tile_step, it += kNG) {

            const int st = it % a.n_stages;
            const uint32_t ph = (it / a.n_stages) & 1u;
            const uint32_t g0 = tile * kFTile;
            const uint32_t ng = a.n_groups - g0 < kFTile ? a.n_groups - g0 : kFTile;
            const uint32_t rbeg = nx_beg, rend = nx_end;
            if (tile + tile_step < n_tiles) {  // next tile of this consumer group: fetched during this one
                nx_beg = a.tile_off[tile + tile_step];
                nx_end = a.tile_off[tile + tile_step + 1];
            }
            const uint32_t cnt = rend - rbeg, staged = cnt < kFRecCap ? cnt : kFRecCap;
            uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
            uint64_t *s_matched = reinterpret_cast<uint64_t *>(sb + o_matched);
            uint64_t *s_next = reinterpret_cast<uint64_t *>(sb + o_next);
            uint64_t *s_pc = reinterpret_cast<uint64_t *>(sb + o_pc);
            uint64_t *s_committed = reinterpret_cast<uint64_t *>(sb + o_committed);
            uint64_t *s_ts = reinterpret_cast<uint64_t *>(sb + o_ts);
            uint64_t *s_li = reinterpret_cast<uint64_t *>(sb + o_li);
            uint32_t *s_meta = reinterpret_cast<uint32_t *>(sb + o_meta);
            uint8_t *s_flags = sb + o_flags;
            const PackedRec *s_recs = reinterpret_cast<const PackedRec *>(sb + o_recs);
            const ulonglong2 *g_recs = reinterpret_cast<const ulonglong2 *>(a.recs + rbeg);
            ulonglong2 q_next = make_ulonglong2(kPkExt, 0ull);
            if (staged == 0 && tid < cnt) q_next = g_recs[tid];  // direct records: in flight during the wait
            long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
            if (a.dbg && tid == 0) t0 = clock64();
            mbar_wait(&full_bar[st], ph);
            if (a.dbg && tid == 0) t1 = clock64();

            // ---- A: the tile's records against the shared-memory rows (raft.rs:1663-1743)
            // Records come from the stage when they were staged (rec_cap > 0), else straight from HBM /
            // L2 (rec_cap == 0: the stage holds rows only, which buys a fifth stage; the records of a
            // tile were prefetched into L2 while the previous tile was processed, and the loop fetches
            // record k + kCT while it works on record k).
            auto rec_at = [&](uint32_t j) -> ulonglong2 {
                if (j < staged) return reinterpret_cast<const ulonglong2 *>(s_recs)[j];
                if (j < cnt) return g_recs[j];
                return make_ulonglong2(kPkExt, 0ull);
            };
            if (staged != 0) q_next = rec_at(tid);
            for (uint32_t k = tid; k < cnt; k += kCT) {
                // Fast path: a record for a staged cell of a peer in Replicate or Probe state -- accept,
                // leader-local, or a rejection without a snapshot request -- straight on the packed
                // words and the shared-memory cell.  Statement for statement the branches of apply_one
                // (raft.rs:1674-1743, 1010-1014; progress.rs:95-114, 138-206); everything else (Snapshot
                // state, request_snapshot, WIDE commits, learners) takes the general path below.
                const ulonglong2 q = q_next;
                q_next = rec_at(k + kCT);
                const bool in_smem = true;
                if (in_smem) {
                    const uint64_t w0 = q.x;
                    if (w0 & kPkExt) {
                        if (a.results) a.results[rbeg + k] = 0;
                        continue;
                    }
                    const uint32_t slot = static_cast<uint32_t>(w0 >> 32) & 7u;
                    const uint32_t g = static_cast<uint32_t>(w0);
                    const uint32_t gl = g - g0;
                    if (!(w0 & kPkWide) && gl < ng && ((hint >> slot) & 1u)) {
                        const uint32_t r = kSimple5 ? slot : static_cast<uint32_t>(__popc(hint & ((1u << slot) - 1u)));
                        const uint32_t ci = r * R64 + gl;
                        const uint32_t f0 = s_flags[r * kFRow8 + gl];
                        const uint32_t state = f0 & RAFTGPU_PF_STATE_MASK;
                        const bool present = kSimple5 || (((RAFTGPU_META_IN(s_meta[gl]) | RAFTGPU_META_OUT(s_meta[gl]) |
                                                            RAFTGPU_META_LEARN(s_meta[gl])) >> slot) & 1u);
                        const bool simple = present && state != RAFTGPU_STATE_SNAPSHOT;
                        if (simple && !(w0 & kPkReject)) {
                            // accept / leader-local: maybe_update (progress.rs:138-150), shared by the accept path
                            // (raft.rs:1674-1677, 1724-1730) and the leader-local path (raft.rs:974-991, 1010-1014);
                            // only an accept looks at is_paused() and may move a probing peer to Replicate.
                            // Written as straight-line selects: this is ~98 % of all records.
                            const uint64_t index = q.y;
                            const uint32_t delta = static_cast<uint32_t>(w0 >> 40);
                            const bool is_local = (w0 & kPkLocal) != 0;
                            const uint64_t m = s_matched[ci], nx = s_next[ci], pcv = s_pc[ci];
                            local[0]++;
                            if (is_local && delta != kPkNoCommit) s_li[gl] = index + delta;   // raft.rs:974-991
                            const uint64_t commit = index - delta;
                            if (!is_local && commit > pcv) s_pc[ci] = commit;                 // raft.rs:1677
                            const bool probe = state == RAFTGPU_STATE_PROBE;
                            const bool need = m < index;
                            const bool old_paused = !is_local && (f0 & (probe ? RAFTGPU_PF_PAUSED : RAFTGPU_PF_INS_FULL)) != 0;
                            const bool trans = need && !is_local && probe;                    // raft.rs:1730 become_replicate
                            uint32_t f = is_local ? f0 : (f0 | RAFTGPU_PF_RECENT_ACTIVE);     // raft.rs:1674
                            if (need) f &= ~RAFTGPU_PF_PAUSED;
                            if (trans)
                                f = (f & ~(RAFTGPU_PF_PAUSED | RAFTGPU_PF_INS_FULL | RAFTGPU_PF_STATE_MASK)) | RAFTGPU_STATE_REPLICATE;
                            uint64_t nnx = nx < index + 1 ? index + 1 : nx;
                            if (trans) {
                                nnx = index + 1;                                              // next_idx = matched + 1
                                c.pending_snapshot[static_cast<size_t>(slot) * c.cap + g] = 0;
                            }
                            local[1] += need ? 1u : 0u;
                            if (need) s_matched[ci] = index;
                            if (nnx != nx) s_next[ci] = nnx;
                            if (f != f0) s_flags[r * kFRow8 + gl] = static_cast<uint8_t>(f);
                            if (a.results)
                                a.results[rbeg + k] = static_cast<uint8_t>(need ? (RAFTGPU_RES_OK | (old_paused ? RAFTGPU_RES_OLD_PAUSED : 0u)) : 0u);
                            continue;
                        }
                        if (simple) {  // a rejection: look at its EXT payloads: [kind 1 hint] [kind 2 snapshot request]
                            uint64_t hint_idx = 0;
                            bool snapshot_req = false;
                            const ulonglong2 e1 = rec_at(k + 1);   // (past the end: a padding EXT of kind 0)
                            const ulonglong2 e2 = rec_at(k + 2);
                            const bool x1 = (e1.x & kPkExt) != 0, x2 = x1 && (e2.x & kPkExt) != 0;
                            if (x1 && (e1.x >> 40) == 1) hint_idx = e1.y;
                            if ((x1 && (e1.x >> 40) == 2) || (x2 && (e2.x >> 40) == 2)) snapshot_req = true;
                            if (!snapshot_req) {
                                // maybe_decr_to without a snapshot request (progress.rs:168-206)
                                const uint64_t index = q.y;
                                const uint32_t delta = static_cast<uint32_t>(w0 >> 40);
                                uint64_t m = s_matched[ci], nx = s_next[ci];
                                const uint64_t nx0 = nx;
                                uint32_t f = f0 | RAFTGPU_PF_RECENT_ACTIVE, res = 0;         // raft.rs:1674
                                local[0]++;
                                local[2]++;
                                const uint64_t commit = index - delta;
                                if (commit > s_pc[ci]) s_pc[ci] = commit;                   // raft.rs:1677
                                bool ok;
                                if (state == RAFTGPU_STATE_REPLICATE) {
                                    ok = index > m;                                          // :173-177 stale otherwise
                                    if (ok) nx = m + 1;                                      // :178-179
                                } else if (nx == 0 || nx - 1 != index) {
                                    ok = false;                                              // :188-192 stale
                                } else {
                                    nx = umin64(index, hint_idx + 1);                        // :195-199
                                    if (nx < 1) nx = 1;
                                    f &= ~RAFTGPU_PF_PAUSED;                                 // :204
                                    ok = true;
                                }
                                if (ok) {
                                    local[3]++;
                                    res = RAFTGPU_RES_OK | RAFTGPU_RES_SEND;
                                    if (state == RAFTGPU_STATE_REPLICATE) {                  // raft.rs:1716-1718 become_probe
                                        f = (f & ~(RAFTGPU_PF_PAUSED | RAFTGPU_PF_INS_FULL | RAFTGPU_PF_STATE_MASK)) | RAFTGPU_STATE_PROBE;
                                        c.pending_snapshot[static_cast<size_t>(slot) * c.cap + g] = 0;
                                        nx = m + 1;
                                    }
                                }
                                if (nx != nx0) s_next[ci] = nx;
                                if (f != f0) s_flags[r * kFRow8 + gl] = static_cast<uint8_t>(f);
                                if (a.results) a.results[rbeg + k] = static_cast<uint8_t>(res);
                                continue;
                            }
                        }
                    }
                }
                // General path: from the stage when the whole tile is staged, else from HBM.
                const bool from_smem = staged == cnt;
                const void *base = from_smem ? static_cast<const void *>(s_recs) : static_cast<const void *>(a.recs + rbeg);
                const uint64_t nn = from_smem ? staged : cnt;
                const RecRegs rec = load_rec<true>(base, k, nn);
                uint32_t res = 0;
                if (!((rec.w0 >> 40) & RAFTGPU_REC_EXT)) {
                    const uint32_t g = static_cast<uint32_t>(rec.w0), slot = static_cast<uint32_t>(rec.w0 >> 32) & 0xffu;
                    const uint32_t gl = g - g0;
                    if (gl >= ng) {  // not this tile's group: the batch is not in group order / bad index
                        local[0]++;
                        local[4]++;
                        res = RAFTGPU_RES_NO_PROGRESS;
                    } else if (slot < kSlots && ((hint >> slot) & 1u)) {
                        const uint32_t r = __popc(hint & ((1u << slot) - 1u));
                        CellRegs cd;
                        cd.meta = s_meta[gl];
                        cd.matched = s_matched[r * R64 + gl];
                        cd.next_idx = s_next[r * R64 + gl];
                        cd.flags = s_flags[r * kFRow8 + gl];
                        cd.peer_committed = s_pc[r * R64 + gl];
                        const CellPtrs sp{&s_matched[r * R64 + gl], &s_next[r * R64 + gl], &s_pc[r * R64 + gl], &s_li[gl],
                                          &s_flags[r * kFRow8 + gl]};
                        res = apply_one<1>(c, base, nn, k, rec, cd, sp, local);
                    } else {  // a peer slot outside the hint (a learner): its cell lives in HBM
                        CellRegs cd = load_cell(c, rec);
                        cd.meta = s_meta[gl];
                        CellPtrs gp = global_cell_ptrs(c, rec);
                        gp.last_index = &s_li[gl];
                        res = apply_one<1>(c, base, nn, k, rec, cd, gp, local);
                    }
                }
                if (a.results) a.results[rbeg + k] = static_cast<uint8_t>(res);
            }
            named_bar_sync(bar_id, kCT);
            if (a.dbg && tid == 0) t2 = clock64();

            // ---- B: Raft::maybe_commit for the tile's groups (raft.rs:893-904)
            if (tid < kFTile) {  // warp-uniform: kFTile is a multiple of 32
                const uint32_t gl = tid;
                const bool active = gl < ng;
                const uint32_t g = g0 + gl;
                bool advanced = false;
                if (active) {
                    const uint32_t meta = s_meta[gl];
                    uint64_t v[kSlots];
                    uint32_t r = 0;
#pragma unroll
                    for (int s2 = 0; s2 < kSlots; s2++) {
                        v[s2] = 0;
                        if ((hint >> s2) & 1u) {
                            v[s2] = s_matched[r * R64 + gl];
                            r++;
                        }
                    }
                    uint64_t mci;
                    bool use_gc;
                    eval_mci<kSimple5>(c, g, meta, v, hint, mci, use_gc);
                    advanced = mci > s_committed[gl] && mci >= s_ts[gl] && mci <= s_li[gl];  // raft_log.rs:488
                    if (advanced) {
                        s_committed[gl] = mci;
                        if (a.commit_out) a.commit_out[g] = mci;
                        if (meta & RAFTGPU_META_HAS_SELF) {  // raft.rs:896-900
                            const uint32_t self = RAFTGPU_META_SELF(meta);
                            uint64_t *pc = ((hint >> self) & 1u)
                                               ? &s_pc[__popc(hint & ((1u << self) - 1u)) * R64 + gl]
                                               : &c.peer_committed[static_cast<size_t>(self) * c.cap + g];
                            if (mci > *pc) *pc = mci;
                        }
                    }
                }
                uint32_t lc[2] = {0, 0};
                publish_tile(a.adv_bitmap, static_cast<uint64_t>(g0) + gl, lane, active, advanced, lc);
                local[5] += lc[0];
                local[6] += lc[1];
            }

            // ---- C: hand the stage to the store warp (generic writes -> async proxy: fence, then signal)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            named_bar_sync(bar_id, kCT);
            if (a.dbg && tid == 0) t3 = clock64();
            if (tid == 0) mbar_arrive(&done_bar[st]);
            if (kFRecCap == 0 && tile + tile_step < n_tiles) {  // direct records: next tile's range -> L2 (one line = 8 records)
                const uint32_t nn = nx_end - nx_beg;
                for (uint32_t k = tid * 8u; k < nn; k += kCT * 8u) prefetch_l2(a.recs + nx_beg + k);
            }
            if (a.dbg && tid == 0) {
                t4 = clock64();
                atomicAdd(&a.dbg[0], static_cast<unsigned long long>(t1 - t0));  // waiting for the TMA loads
                atomicAdd(&a.dbg[1], static_cast<unsigned long long>(t2 - t1));  // A: records
                atomicAdd(&a.dbg[2], static_cast<unsigned long long>(t3 - t2));  // B: recompute (+ fence, barrier)
                atomicAdd(&a.dbg[3], static_cast<unsigned long long>(t4 - t3));  // C: stores + drain
                atomicAdd(&a.dbg[4], 1ull);                                       // tiles
            }
        }
    }
    const int which[7] = {kCntRecords, kCntUpdates, kCntRejects, kCntDecrements, kCntNoProgress, kCntRecomputes, kCntAdvanced};
    block_flush_counts<7>(local, which, a.counters, nullptr);
    if (a.step_advanced) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, local[6]);
        if (lane == 0 && w) atomicAdd(a.step_advanced, w);
    }
}

// step_tile_compact_kernel: the fused step for a TILEABLE compact stream (groups ascending, every
// run headed -- raftgpu_compact_hdr.flags & RAFTGPU_COMPACT_TILEABLE).  Same producer / store
// warps and stage ring as step_tile_kernel; the consumers work per GROUP instead of per record:
//   0. a unit-parallel pre-pass finds, for every group of the tile, the first unit of its first
//      run (shared-memory atomicMin on the run headers),
//   1. thread gl then walks the runs of group g0+gl SEQUENTIALLY -- header decoded once, each
//      record a handful of instructions against the shared-memory cells -- so the records of one
//      cell apply in stream order (no one-wave restriction) and
//   2. goes straight on to the group's Raft::maybe_commit: its cells were written by this very
//      thread, so no barrier separates apply and recompute.
// Against the per-record form this is ~2.5x fewer warp instructions per tile (no per-record
// header/cell address arithmetic, no three-way divergence on the record kind) and a third of the
// record bytes.  tile_off[t] = unit position of the first run header of tile t.
struct CTileArgs {
    CompactSrc src;
    const uint32_t *tile_off;  // [n_tiles + 1], unit positions
    const uint2 *tile_gb;      // [n_tiles]: g_base of the unit block tile t starts in, and of the next block
    uint32_t n_groups;
    uint32_t hint;
    int n_stages;
    uint32_t unit_cap;         // units staged in shared memory per tile (multiple of 4); the rest is read from HBM
    uint8_t *results;          // nullable, one byte per unit
    uint32_t *adv_bitmap;      // nullable
    uint64_t *commit_out;      // nullable
    uint32_t *step_advanced;   // nullable
    unsigned long long *counters;
    unsigned long long *dbg;   // nullable
    uint32_t *dup_count;       // nullable (!kOrdered): bumped for a second record on one cell, which is not applied
};

__host__ __device__ constexpr uint32_t ctile_stage_bytes(uint32_t H, uint32_t unit_cap) {
    return 3u * H * kFRow64 + 3u * kFRow64 + kFTile * 4u + H * kFRow8 + unit_cap * 4u;
}
// dynamic shared memory: the stages, then one u32 run index per group per consumer group
__host__ __device__ constexpr uint32_t ctile_smem_bytes(uint32_t H, uint32_t unit_cap, int stages, int n_groups_c) {
    return static_cast<uint32_t>(stages) * ctile_stage_bytes(H, unit_cap) + static_cast<uint32_t>(n_groups_c) * kFTile * 4u;
}

struct TileRows {
    uint64_t *matched, *next, *pc, *committed, *ts, *li;
    uint32_t *meta;
    uint8_t *flags;
};

// A record that is not on the fast path (Snapshot state, a REJECT / hostile value from the side
// table, a peer slot outside the hint): the literal apply_one against the shared-memory cell, or
// against HBM for a slot the tile does not stage.
__device__ __forceinline__ uint32_t tile_general_apply(const Columns &c, const RecRegs &rec, uint64_t ext_hint,
                                                    uint64_t ext_snapshot, uint32_t gl, uint32_t hint,
                                                    const TileRows &t, uint32_t *local) {
    constexpr uint32_t R64 = kFRow64 / 8;
    const uint32_t slot = static_cast<uint32_t>(rec.w0 >> 32) & 0xffu;
    if (slot < kSlots && ((hint >> slot) & 1u)) {
        const uint32_t r = __popc(hint & ((1u << slot) - 1u));
        CellRegs cd;
        cd.meta = t.meta[gl];
        cd.matched = t.matched[r * R64 + gl];
        cd.next_idx = t.next[r * R64 + gl];
        cd.flags = t.flags[r * kFRow8 + gl];
        cd.peer_committed = t.pc[r * R64 + gl];
        const CellPtrs sp{&t.matched[r * R64 + gl], &t.next[r * R64 + gl], &t.pc[r * R64 + gl], &t.li[gl],
                          &t.flags[r * kFRow8 + gl]};
        return apply_one<2>(c, nullptr, ext_hint, ext_snapshot, rec, cd, sp, local);
    }
    CellRegs cd = load_cell(c, rec);
    cd.meta = t.meta[gl];
    CellPtrs gp = global_cell_ptrs(c, rec);
    gp.last_index = &t.li[gl];
    return apply_one<2>(c, nullptr, ext_hint, ext_snapshot, rec, cd, gp, local);
}

// The common records -- accept, leader-local, rejection without a snapshot request -- for a peer
// in Replicate or Probe state whose cell is staged in shared memory: statement for statement the
// branches of apply_one (raft.rs:1674-1677, 1709-1730, 974-991, 1010-1014; progress.rs:95-114,
// 138-157, 168-206).  Returns false (nothing touched) when the record needs the general path.
template <bool kSimple5>
__device__ __forceinline__ bool tile_fast_apply(const Columns &c, const TileRows &t, uint32_t gl, uint32_t g, uint32_t slot,
                                                uint64_t index, uint64_t commit, bool is_local, bool is_reject,
                                                uint64_t ext_hint, uint32_t hint, uint32_t present_mask, uint32_t *local,
                                                uint32_t &res) {
    constexpr uint32_t R64 = kFRow64 / 8;
    if (!((hint >> slot) & 1u)) return false;
    const uint32_t r = kSimple5 ? slot : static_cast<uint32_t>(__popc(hint & ((1u << slot) - 1u)));
    const uint32_t ci = r * R64 + gl;
    const uint32_t f0 = t.flags[r * kFRow8 + gl];
    const uint32_t state = f0 & RAFTGPU_PF_STATE_MASK;
    if (!((present_mask >> slot) & 1u) || state == RAFTGPU_STATE_SNAPSHOT) return false;
    uint64_t m = t.matched[ci], nx = t.next[ci];
    const uint64_t m0 = m, nx0 = nx;
    uint32_t f = f0;
    local[0]++;
    if (is_local) {
        if (commit != 0) t.li[gl] = commit;
    } else {
        f |= RAFTGPU_PF_RECENT_ACTIVE;
        if (commit > t.pc[ci]) t.pc[ci] = commit;
    }
    if (is_reject) {  // maybe_decr_to without a snapshot request
        local[2]++;
        bool ok;
        if (state == RAFTGPU_STATE_REPLICATE) {
            ok = index > m;                              // progress.rs:173-177 stale otherwise
            if (ok) nx = m + 1;                          // :178-179
        } else if (nx == 0 || nx - 1 != index) {
            ok = false;                                  // :188-192 stale
        } else {
            nx = umin64(index, ext_hint + 1);            // :195-199
            if (nx < 1) nx = 1;
            f &= ~RAFTGPU_PF_PAUSED;                     // :204
            ok = true;
        }
        if (ok) {
            local[3]++;
            res = RAFTGPU_RES_OK | RAFTGPU_RES_SEND;
            if (state == RAFTGPU_STATE_REPLICATE) {      // raft.rs:1716-1718 become_probe
                f = (f & ~(RAFTGPU_PF_PAUSED | RAFTGPU_PF_INS_FULL | RAFTGPU_PF_STATE_MASK)) | RAFTGPU_STATE_PROBE;
                c.pending_snapshot[static_cast<size_t>(slot) * c.cap + g] = 0;
                nx = m + 1;
            }
        }
    } else {
        const bool old_paused = !is_local && (state == RAFTGPU_STATE_PROBE ? (f & RAFTGPU_PF_PAUSED) != 0
                                                                           : (f & RAFTGPU_PF_INS_FULL) != 0);
        const bool need = m < index;
        if (need) {
            m = index;
            f &= ~RAFTGPU_PF_PAUSED;
            local[1]++;
            res = RAFTGPU_RES_OK | (old_paused ? RAFTGPU_RES_OLD_PAUSED : 0u);
        }
        if (nx < index + 1) nx = index + 1;
        if (need && !is_local && state == RAFTGPU_STATE_PROBE) {
            f = (f & ~(RAFTGPU_PF_PAUSED | RAFTGPU_PF_INS_FULL | RAFTGPU_PF_STATE_MASK)) | RAFTGPU_STATE_REPLICATE;
            c.pending_snapshot[static_cast<size_t>(slot) * c.cap + g] = 0;
            nx = m + 1;
        }
    }
    if (m != m0) t.matched[ci] = m;
    if (nx != nx0) t.next[ci] = nx;
    if (f != f0) t.flags[r * kFRow8 + gl] = static_cast<uint8_t>(f);
    return true;
}

// kOrdered: per-group walk (records of a cell apply in stream order; any number per cell).
// !kOrdered: one thread per unit (faster; at most one record per (group, peer) cell per batch --
// checked on the fly in a shared-memory bitmap when a.dup_count is given).
template <bool kSimple5, int kNG, bool kOrdered>
__global__ void __launch_bounds__(kFTile *kNG + 64, 1) step_tile_compact_kernel(Columns c, CTileArgs a) {
    constexpr int kCT = kFTile;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kFMaxStages];
    __shared__ __align__(8) uint64_t done_bar[kFMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kFMaxStages];

    const uint32_t hint = kSimple5 ? 0x1fu : (a.hint & 0xffu);
    const uint32_t H = kSimple5 ? 5u : static_cast<uint32_t>(__popc(hint));
    const uint32_t o_matched = 0, o_next = H * kFRow64, o_pc = 2u * H * kFRow64, o_committed = 3u * H * kFRow64,
                   o_ts = o_committed + kFRow64, o_li = o_ts + kFRow64, o_meta = o_li + kFRow64,
                   o_flags = o_meta + kFTile * 4u, o_units = o_flags + H * kFRow8,
                   stage_bytes = ctile_stage_bytes(H, a.unit_cap);
    const uint32_t n_tiles = (a.n_groups + kFTile - 1) / kFTile;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *s_run_all = reinterpret_cast<uint32_t *>(smem + static_cast<size_t>(a.n_stages) * stage_bytes);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&done_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < kNG * kCT) s_run_all[threadIdx.x] = kOrdered ? 0xffffffffu : 0u;
    __syncthreads();

    uint32_t local[7] = {0, 0, 0, 0, 0, 0, 0};  // records, updates, rejects, decrements, no_progress | recomputes, advanced
    if (warp == kNG * kCT / 32 + 1) {
        // ===================== store warp =====================
        const uint32_t n_out = 4u * H + 2u;
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int st = it % a.n_stages;
            const uint32_t ph = (it / a.n_stages) & 1u;
            const uint32_t g0 = tile * kFTile;
            const uint32_t ng = a.n_groups - g0 < kFTile ? a.n_groups - g0 : kFTile;
            const uint32_t ng16 = (ng + 15u) & ~15u;
            uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
            mbar_wait(&done_bar[st], ph);
            for (uint32_t j = lane; j < n_out; j += 32) {
                if (j < 4u * H) {
                    const uint32_t col = j / H, r = j % H;
                    uint32_t slot = 0, seen = 0;
                    for (uint32_t s2 = 0; s2 < kSlots; s2++)
                        if ((hint >> s2) & 1u) {
                            if (seen == r) slot = s2;
                            seen++;
                        }
                    const size_t cell = static_cast<size_t>(slot) * c.cap + g0;
                    if (col == 0) tma_store_1d(c.matched + cell, sb + o_matched + r * kFRow64, ng16 * 8u);
                    if (col == 1) tma_store_1d(c.next_idx + cell, sb + o_next + r * kFRow64, ng16 * 8u);
                    if (col == 2) tma_store_1d(c.peer_committed + cell, sb + o_pc + r * kFRow64, ng16 * 8u);
                    if (col == 3) tma_store_1d(c.pflags + cell, sb + o_flags + r * kFRow8, ng16);
                } else if (j == 4u * H) {
                    tma_store_1d(c.committed + g0, sb + o_committed, ng16 * 8u);
                } else {
                    tma_store_1d(c.last_index + g0, sb + o_li, ng16 * 8u);
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (warp == kNG * kCT / 32) {
        // ===================== load warp =====================
        const uint32_t n_copies = 4u * H + 5u;
        uint32_t it = 0;
        uint32_t nx_u0 = blockIdx.x < n_tiles ? a.tile_off[blockIdx.x] : 0u, nx_u1 = blockIdx.x < n_tiles ? a.tile_off[blockIdx.x + 1] : 0u;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int st = it % a.n_stages;
            const uint32_t ph = (it / a.n_stages) & 1u;
            const uint32_t g0 = tile * kFTile;
            const uint32_t ng = a.n_groups - g0 < kFTile ? a.n_groups - g0 : kFTile;
            const uint32_t ng16 = (ng + 15u) & ~15u;
            // the tile's units, from the 16-byte boundary below its first one
            const uint32_t u0 = nx_u0, u1 = nx_u1;
            if (tile + gridDim.x < n_tiles) {  // next tile's index: fetched while this one is being issued
                nx_u0 = a.tile_off[tile + gridDim.x];
                nx_u1 = a.tile_off[tile + gridDim.x + 1];
            }
            const uint32_t ua = u0 & ~3u;
            const uint32_t cnt4 = (u1 - ua + 3u) & ~3u;
            const uint32_t staged = u1 > u0 ? (cnt4 < a.unit_cap ? cnt4 : a.unit_cap) : 0u;
            if (lane == 0) {
                mbar_wait(&empty_bar[st], ph ^ 1u);
                mbar_expect_tx(&full_bar[st], 3u * H * ng16 * 8u + H * ng16 + 3u * ng16 * 8u + ng16 * 4u + staged * 4u);
            }
            __syncwarp();
            uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
            for (uint32_t j = lane; j < n_copies; j += 32) {
                if (j < 4u * H) {
                    const uint32_t col = j / H, r = j % H;
                    uint32_t slot = 0, seen = 0;
                    for (uint32_t s2 = 0; s2 < kSlots; s2++)
                        if ((hint >> s2) & 1u) {
                            if (seen == r) slot = s2;
                            seen++;
                        }
                    const size_t cell = static_cast<size_t>(slot) * c.cap + g0;
                    if (col == 0) tma_load_1d(sb + o_matched + r * kFRow64, c.matched + cell, ng16 * 8u, &full_bar[st]);
                    if (col == 1) tma_load_1d(sb + o_next + r * kFRow64, c.next_idx + cell, ng16 * 8u, &full_bar[st]);
                    if (col == 2) tma_load_1d(sb + o_pc + r * kFRow64, c.peer_committed + cell, ng16 * 8u, &full_bar[st]);
                    if (col == 3) tma_load_1d(sb + o_flags + r * kFRow8, c.pflags + cell, ng16, &full_bar[st]);
                } else if (j == 4u * H) {
                    tma_load_1d(sb + o_committed, c.committed + g0, ng16 * 8u, &full_bar[st]);
                } else if (j == 4u * H + 1) {
                    tma_load_1d(sb + o_ts, c.term_start + g0, ng16 * 8u, &full_bar[st]);
                } else if (j == 4u * H + 2) {
                    tma_load_1d(sb + o_li, c.last_index + g0, ng16 * 8u, &full_bar[st]);
                } else if (j == 4u * H + 3) {
                    tma_load_1d(sb + o_meta, c.meta + g0, ng16 * 4u, &full_bar[st]);
                } else if (staged) {
                    tma_load_1d(sb + o_units, a.src.units + ua, staged * 4u, &full_bar[st]);
                }
            }
        }
    } else {
        // ===================== consumers: thread gl of group cg owns group g0 + gl of its tiles =====================
        const uint32_t cg = warp / (kCT / 32);
        const uint32_t tid = threadIdx.x - cg * kCT;
        const int bar_id = 1 + static_cast<int>(cg);
        constexpr uint32_t R64 = kFRow64 / 8;
        uint32_t *s_run = s_run_all + cg * kFTile;
        uint32_t it = 0;
        // tile index and g_base words of this consumer group's NEXT tile are fetched during the current
        // one (tile_gb[t] = the g_base words of the two unit blocks tile t starts in)
        const uint32_t first_tile = blockIdx.x + cg * gridDim.x, tile_step = kNG * gridDim.x;
        uint32_t nx_u0 = 0, nx_u1 = 0;
        uint2 nx_gb = make_uint2(0u, 0u);
        if (first_tile < n_tiles) {
            nx_u0 = a.tile_off[first_tile];
            nx_u1 = a.tile_off[first_tile + 1];
            nx_gb = a.tile_gb[first_tile];
        }
        it = cg;
        for (uint32_t tile = first_tile; tile < n_tiles; tile += tile_step, it += kNG) {

            const int st = it % a.n_stages;
            const uint32_t ph = (it / a.n_stages) & 1u;
            const uint32_t g0 = tile * kFTile;
            const uint32_t ng = a.n_groups - g0 < kFTile ? a.n_groups - g0 : kFTile;
            const uint32_t u0 = nx_u0, u1 = nx_u1;
            const uint32_t gb0 = nx_gb.x, gb1 = nx_gb.y;
            if (tile + tile_step < n_tiles) {
                nx_u0 = a.tile_off[tile + tile_step];
                nx_u1 = a.tile_off[tile + tile_step + 1];
                nx_gb = a.tile_gb[tile + tile_step];
            }
            const uint32_t ua = u0 & ~3u;
            const uint32_t lead = u0 - ua, cnt = u1 - ua;  // local unit positions [lead, cnt)
            const uint32_t cnt4 = (cnt + 3u) & ~3u;
            const uint32_t staged = u1 > u0 ? (cnt4 < a.unit_cap ? cnt4 : a.unit_cap) : 0u;
            uint8_t *sb = smem + static_cast<size_t>(st) * stage_bytes;
            TileRows t;
            t.matched = reinterpret_cast<uint64_t *>(sb + o_matched);
            t.next = reinterpret_cast<uint64_t *>(sb + o_next);
            t.pc = reinterpret_cast<uint64_t *>(sb + o_pc);
            t.committed = reinterpret_cast<uint64_t *>(sb + o_committed);
            t.ts = reinterpret_cast<uint64_t *>(sb + o_ts);
            t.li = reinterpret_cast<uint64_t *>(sb + o_li);
            t.meta = reinterpret_cast<uint32_t *>(sb + o_meta);
            t.flags = sb + o_flags;
            const uint32_t *s_units = reinterpret_cast<const uint32_t *>(sb + o_units);
            const uint32_t *g_units = a.src.units + ua;
            const uint32_t blk0 = ua / RAFTGPU_COMPACT_BLOCK;
            long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
            if (a.dbg && tid == 0) t0 = clock64();
            mbar_wait(&full_bar[st], ph);
            if (a.dbg && tid == 0) t1 = clock64();

            // ---- 0: first run header of every group of the tile.  The g_base words of the (at most
            // two, unless the tile is crowded) unit blocks the tile spans were fetched before the wait.
            auto group_of = [&](uint32_t p, uint32_t hb) -> uint32_t {
                const uint32_t blk = (ua + p) / RAFTGPU_COMPACT_BLOCK;
                const uint32_t gb = blk == blk0 ? gb0 : (blk == blk0 + 1u ? gb1 : a.src.g_base[blk]);
                return gb + ((hb >> 2) & 0xfffu);
            };
            const uint32_t gl = tid;
            const bool active = gl < ng;
            const uint32_t g = g0 + gl;
            if constexpr (kOrdered) {
                for (uint32_t p = lead + tid; p < cnt; p += kCT) {
                    const uint32_t u = p < staged ? s_units[p] : g_units[p];
                    const uint32_t kind = u & 3u;
                    if (kind == kCuEsc) {  // a side-table record: start pulling it towards L1 for the walk below
                        const uint32_t k = u >> 2;
                        if (k < kCuPad && k < a.src.n_side) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.src.side + k));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const uint8_t *>(a.src.side + k) + 40));
                        }
                        continue;
                    }
                    if (kind != kCuHdrA || p + 1u >= cnt) continue;
                    const uint32_t hb = p + 1u < staged ? s_units[p + 1u] : g_units[p + 1u];
                    if ((hb & 3u) != kCuHdrB) continue;
                    const uint32_t gt = group_of(p, hb) - g0;
                    if (gt < ng) atomicMin(&s_run[gt], p);
                }
                named_bar_sync(bar_id, kCT);
                if (a.dbg && tid == 0) t2 = clock64();

                // ---- 1: the runs of group g, record by record (raft.rs:1663-1743)
                uint32_t p = s_run[gl];
                s_run[gl] = 0xffffffffu;  // ready for this consumer group's next tile
                if (active && p != 0xffffffffu) {
                    const uint32_t meta = t.meta[gl];
                    const uint32_t present_mask = kSimple5 ? 0x1fu
                                                           : (RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta) | RAFTGPU_META_LEARN(meta));
                    uint64_t base = 0;
                    while (p < cnt) {
                        const uint32_t u = p < staged ? s_units[p] : g_units[p];
                        const uint32_t kind = u & 3u;
                        if (kind == kCuHdrA) {  // a run header: ours (the first, or a continuation) or the next group's
                            if (p + 1u >= cnt) break;
                            const uint32_t hb = p + 1u < staged ? s_units[p + 1u] : g_units[p + 1u];
                            if ((hb & 3u) != kCuHdrB) break;
                            if (group_of(p, hb) != g) break;
                            base = static_cast<uint64_t>(u >> 2) | (static_cast<uint64_t>(hb >> 14) << 30);
                            p += 2u;
                            continue;
                        }
                        if (kind == kCuHdrB) break;  // malformed
                        uint32_t res = 0;
                        RecRegs rec;
                        uint64_t ext_hint = 0, ext_snapshot = RAFTGPU_INVALID_INDEX;
                        bool general = false;
                        uint32_t step = 1;
                        if (kind == kCuRec) {
                            const uint32_t slot = (u >> 6) & 7u;
                            const uint64_t index = base + ((u >> 10) & 0x3fffu);
                            const uint32_t cd = u >> 24;
                            const bool is_local = (u & kCuLocal) != 0;
                            const bool is_reject = (u & kCuReject) != 0;
                            const uint64_t commit = is_local ? (cd == kCuNoCommit ? 0 : index + cd) : (index >= cd ? index - cd : 0);
                            if (is_reject && p + 1u < cnt) {  // its hint rides in the next unit
                                const uint32_t pl = p + 1u < staged ? s_units[p + 1u] : g_units[p + 1u];
                                if ((pl & 3u) == kCuEsc && ((pl >> 2) & kCuPayload)) {
                                    ext_hint = compact_hint(index, pl);
                                    step = 2;
                                }
                            }
                            const bool fast = tile_fast_apply<kSimple5>(c, t, gl, g, slot, index, commit, is_local, is_reject, ext_hint,
                                                                        hint, present_mask, local, res);
                            if (!fast) {
                                general = true;
                                rec.w0 = static_cast<uint64_t>(g) | (static_cast<uint64_t>(slot) << 32) |
                                         (static_cast<uint64_t>(is_reject ? RAFTGPU_REC_REJECT : (is_local ? RAFTGPU_REC_LOCAL : 0u)) << 40);
                                rec.index = index;
                                rec.commit = commit;
                            }
                        } else {  // ESC: the full record sits in the side table (a stray payload unit: nothing)
                            const uint32_t k = u >> 2;
                            if (k == kCuPad) break;  // padding only ever follows the last run of a slice
                            if (k < kCuPad && k < a.src.n_side) {
                                const uint64_t *sp = reinterpret_cast<const uint64_t *>(a.src.side + k);
                                rec.w0 = sp[0];
                                rec.index = sp[1];
                                rec.commit = sp[2];
                                if ((rec.w0 >> 40) & RAFTGPU_REC_EXT) {
                                    // a continuation by itself carries nothing
                                } else if (static_cast<uint32_t>(rec.w0) != g) {  // not this run's group: the stream lied
                                    local[0]++;
                                    local[4]++;
                                    res = RAFTGPU_RES_NO_PROGRESS;
                                } else {
                                    general = true;
                                    if ((rec.w0 >> 40) & RAFTGPU_REC_REJECT)
                                        load_reject_ext<false>(a.src.side, k, a.src.n_side, ext_hint, ext_snapshot);
                                }
                            }
                        }
                        if (general) res = tile_general_apply(c, rec, ext_hint, ext_snapshot, gl, hint, t, local);
                        if (a.results) a.results[ua + p] = static_cast<uint8_t>(res);
                        p += step;
                    }
                }
            } else {
                // ---- 1': one thread per unit; a record finds its run header `back` units behind it
                uint32_t *s_touch = s_run;  // 2048 bits: (group, peer slot) cells seen in this tile
                for (uint32_t k = lead + tid; k < cnt; k += kCT) {
                    const uint32_t u = k < staged ? s_units[k] : g_units[k];
                    const uint32_t kind = u & 3u;
                    if (kind == kCuHdrA || kind == kCuHdrB) continue;
                    uint32_t res = 0, gt = 0, slot = 0;
                    RecRegs rec;
                    uint64_t ext_hint = 0, ext_snapshot = RAFTGPU_INVALID_INDEX;
                    bool general = false, fast_ok = false, is_local = false, is_reject = false;
                    if (kind == kCuRec) {
                        const uint32_t back = (u >> 3) & 7u;
                        if (k < lead + back + 2u) continue;  // malformed
                        const uint32_t hp = k - back - 2u;
                        const uint32_t ha = hp < staged ? s_units[hp] : g_units[hp];
                        const uint32_t hb = hp + 1u < staged ? s_units[hp + 1u] : g_units[hp + 1u];
                        if ((ha & 3u) != kCuHdrA || (hb & 3u) != kCuHdrB) continue;
                        gt = group_of(hp, hb) - g0;
                        slot = (u >> 6) & 7u;
                        rec.index = (static_cast<uint64_t>(ha >> 2) | (static_cast<uint64_t>(hb >> 14) << 30)) + ((u >> 10) & 0x3fffu);
                        const uint32_t cd = u >> 24;
                        is_local = (u & kCuLocal) != 0;
                        is_reject = (u & kCuReject) != 0;
                        rec.commit = is_local ? (cd == kCuNoCommit ? 0 : rec.index + cd) : (rec.index >= cd ? rec.index - cd : 0);
                        if (is_reject && k + 1u < cnt) {  // its hint rides in the next unit
                            const uint32_t pl = k + 1u < staged ? s_units[k + 1u] : g_units[k + 1u];
                            if ((pl & 3u) == kCuEsc && ((pl >> 2) & kCuPayload)) ext_hint = compact_hint(rec.index, pl);
                        }
                        rec.w0 = static_cast<uint64_t>(g0 + gt) | (static_cast<uint64_t>(slot) << 32) |
                                 (static_cast<uint64_t>(is_reject ? RAFTGPU_REC_REJECT : (is_local ? RAFTGPU_REC_LOCAL : 0u)) << 40);
                        fast_ok = true;
                    } else {  // ESC: the full record sits in the side table (payload units / padding: nothing)
                        const uint32_t k2 = u >> 2;
                        if (k2 >= kCuPad || k2 >= a.src.n_side) continue;
                        const uint64_t *sp = reinterpret_cast<const uint64_t *>(a.src.side + k2);
                        rec.w0 = sp[0];
                        rec.index = sp[1];
                        rec.commit = sp[2];
                        if ((rec.w0 >> 40) & RAFTGPU_REC_EXT) continue;  // a continuation by itself carries nothing
                        gt = static_cast<uint32_t>(rec.w0) - g0;
                        slot = static_cast<uint32_t>(rec.w0 >> 32) & 0xffu;
                        if ((rec.w0 >> 40) & RAFTGPU_REC_REJECT)
                            load_reject_ext<false>(a.src.side, k2, a.src.n_side, ext_hint, ext_snapshot);
                    }
                    if (gt >= ng) {  // not this tile's group: the stream lied
                        local[0]++;
                        local[4]++;
                        res = RAFTGPU_RES_NO_PROGRESS;
                    } else {
                        bool dup = false;
                        if (a.dup_count && slot < kSlots) {
                            const uint32_t bit = 1u << (8u * (gt & 3u) + slot);
                            dup = (atomicOr(&s_touch[gt >> 2], bit) & bit) != 0;
                            if (dup) atomicAdd(a.dup_count, 1u);
                        }
                        if (!dup) {
                            const uint32_t present_mask =
                                kSimple5 ? 0x1fu
                                         : (RAFTGPU_META_IN(t.meta[gt]) | RAFTGPU_META_OUT(t.meta[gt]) | RAFTGPU_META_LEARN(t.meta[gt]));
                            general = !(fast_ok && tile_fast_apply<kSimple5>(c, t, gt, g0 + gt, slot, rec.index, rec.commit, is_local,
                                                                             is_reject, ext_hint, hint, present_mask, local, res));
                            if (general) res = tile_general_apply(c, rec, ext_hint, ext_snapshot, gt, hint, t, local);
                        }
                    }
                    if (a.results) a.results[ua + k] = static_cast<uint8_t>(res);
                }
                named_bar_sync(bar_id, kCT);
                if (a.dup_count && tid < kFTile / 4) s_touch[tid] = 0;  // clean for this consumer group's next tile
                if (a.dbg && tid == 0) t2 = clock64();
            }
            __syncwarp();
            if (a.dbg && tid == 0) t3 = clock64();

            // ---- 2: Raft::maybe_commit for this thread's group (raft.rs:893-904)
            {
                bool advanced = false;
                if (active) {
                    const uint32_t meta = t.meta[gl];
                    uint64_t v[kSlots];
                    uint32_t r = 0;
#pragma unroll
                    for (int s2 = 0; s2 < kSlots; s2++) {
                        v[s2] = 0;
                        if ((hint >> s2) & 1u) {
                            v[s2] = t.matched[r * R64 + gl];
                            r++;
                        }
                    }
                    uint64_t mci;
                    bool use_gc;
                    eval_mci<kSimple5>(c, g, meta, v, hint, mci, use_gc);
                    advanced = mci > t.committed[gl] && mci >= t.ts[gl] && mci <= t.li[gl];  // raft_log.rs:488
                    if (advanced) {
                        t.committed[gl] = mci;
                        if (a.commit_out) a.commit_out[g] = mci;
                        if (meta & RAFTGPU_META_HAS_SELF) {  // raft.rs:896-900
                            const uint32_t self = RAFTGPU_META_SELF(meta);
                            uint64_t *pc = ((hint >> self) & 1u)
                                               ? &t.pc[__popc(hint & ((1u << self) - 1u)) * R64 + gl]
                                               : &c.peer_committed[static_cast<size_t>(self) * c.cap + g];
                            if (mci > *pc) *pc = mci;
                        }
                    }
                }
                uint32_t lc[2] = {0, 0};
                publish_tile(a.adv_bitmap, static_cast<uint64_t>(g0) + gl, lane, active, advanced, lc);
                local[5] += lc[0];
                local[6] += lc[1];
            }

            // ---- 3: hand the stage to the store warp
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            named_bar_sync(bar_id, kCT);
            if (tid == 0) mbar_arrive(&done_bar[st]);
            if (a.dbg && tid == 0) {
                t4 = clock64();
                atomicAdd(&a.dbg[0], static_cast<unsigned long long>(t1 - t0));  // waiting for the TMA loads
                atomicAdd(&a.dbg[1], static_cast<unsigned long long>(t2 - t1));  // 0: run index
                atomicAdd(&a.dbg[2], static_cast<unsigned long long>(t3 - t2));  // 1: records
                atomicAdd(&a.dbg[3], static_cast<unsigned long long>(t4 - t3));  // 2 + 3: recompute, fence, barrier
                atomicAdd(&a.dbg[4], 1ull);
            }
        }
    }
    const int which[7] = {kCntRecords, kCntUpdates, kCntRejects, kCntDecrements, kCntNoProgress, kCntRecomputes, kCntAdvanced};
    block_flush_counts<7>(local, which, a.counters, nullptr);
    if (a.step_advanced) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, local[6]);
        if (lane == 0 && w) atomicAdd(a.step_advanced, w);
    }
}

// Tile index of a tileable compact stream, on the device (the zero-copy step has no host pass
// over the units): tile_off[t] = position of the first run header whose group is >= t * kFTile.
// One thread per unit; a header finds the previous run's group by looking back over at most one
// run (2 + 8 units).  *bad is bumped when groups do not ascend or an ESC unit has no run.
__global__ void __launch_bounds__(256)
compact_tile_index_kernel(CompactSrc src, uint32_t n_groups, uint32_t *__restrict__ tile_off, uint32_t *__restrict__ bad) {
    const uint32_t n_tiles = (n_groups + kFTile - 1) / kFTile;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < src.n_units; i += stride) {
        const uint32_t u = src.units[i];
        if ((u & 3u) == kCuEsc && (u >> 2) < kCuPad) {  // a side-table record must sit inside a run
            bool headed = false;
            for (uint32_t b = 1; b <= 8u && b <= i; b++) {
                const uint32_t k = src.units[i - b] & 3u;
                if (k == kCuHdrB) {
                    headed = true;
                    break;
                }
            }
            if (!headed) atomicAdd(bad, 1u);
            continue;
        }
        if ((u & 3u) != kCuHdrA || i + 1 >= src.n_units) continue;
        const uint32_t hb = src.units[i + 1];
        if ((hb & 3u) != kCuHdrB) continue;
        const uint32_t g = src.g_base[i / RAFTGPU_COMPACT_BLOCK] + ((hb >> 2) & 0xfffu);
        // previous run header, if any: at most one run (2 + 8 units) back, not counting the padding
        // that fills a staging slice up to its block boundary (raftgpu_step_begin_records)
        bool have_prev = false;
        uint32_t gp = 0;
        for (uint32_t b = 2, real = 0; b <= i && real <= 10u; b++) {
            const uint32_t ua = src.units[i - b];
            if (ua == (kCuEsc | (kCuPad << 2))) continue;
            real++;
            if ((ua & 3u) == kCuHdrA && (src.units[i - b + 1] & 3u) == kCuHdrB) {
                gp = src.g_base[(i - b) / RAFTGPU_COMPACT_BLOCK] + ((src.units[i - b + 1] >> 2) & 0xfffu);
                have_prev = true;
                break;
            }
        }
        if (!have_prev) {  // the first run of the stream: everything before it must be padding
            for (uint32_t b = 1; b <= i && b <= 12u; b++)
                if (src.units[i - b] != (kCuEsc | (kCuPad << 2))) {
                    atomicAdd(bad, 1u);
                    break;
                }
        }
        if (have_prev && gp > g) atomicAdd(bad, 1u);       // groups must ascend
        if (g >= n_groups) {
            atomicAdd(bad, 1u);
            continue;
        }
        const uint32_t t_hi = g / kFTile;
        const uint32_t t_lo = have_prev ? (gp / kFTile) + 1u : 0u;  // tiles (prev tile, this tile] start here
        if (!have_prev || gp / kFTile != t_hi)
            for (uint32_t t2 = t_lo; t2 <= t_hi; t2++) tile_off[t2] = i;
    }
    // tiles behind the last run are filled by the host-side launch (tile_off pre-set to n_units)
    (void)n_tiles;
}

// tile_gb[t] = the g_base words of the unit block tile t's (16-byte aligned) first unit lies in and of
// the block after it, so that the fused kernel decodes group ids without dependent global loads.
__global__ void __launch_bounds__(256)
compact_tile_gb_kernel(CompactSrc src, uint32_t n_tiles, const uint32_t *__restrict__ tile_off, uint2 *__restrict__ tile_gb) {
    const uint32_t n_blk = (src.n_units + RAFTGPU_COMPACT_BLOCK - 1) / RAFTGPU_COMPACT_BLOCK;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
        const uint32_t blk0 = (tile_off[t] & ~3u) / RAFTGPU_COMPACT_BLOCK;
        tile_gb[t] = make_uint2(blk0 < n_blk ? src.g_base[blk0] : 0u, blk0 + 1u < n_blk ? src.g_base[blk0 + 1u] : 0u);
    }
}

// ---------------------------------------------------------------------------
// send_list_kernel: bcast_append (raft.rs:857-865) behind Raft::maybe_commit (raft.rs:1745-1748) as a
// stream compaction.  One lane per group, a warp per 32 groups = one word of the advanced bitmap;
// a selected group contributes one entry per present peer other than itself that is not paused
// (progress.rs:210-216).  Lanes count their entries, a warp scan turns the counts into offsets, ONE
// global atomic per warp reserves the range, each lane writes its 16-byte entries.
// Algorithmic bytes: 4 per 32 groups (bitmap) + per advanced group 4 (meta) + K x 17 (pflags,
// next_idx, pending_request_snapshot of its peers) read, 16 written per entry.
__global__ void __launch_bounds__(256)
send_list_kernel(Columns c, uint32_t first, uint32_t n, const uint32_t *__restrict__ adv_bitmap,
                 raftgpu_send_entry *__restrict__ out, unsigned long long capacity, unsigned long long *__restrict__ count) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = first & ~31u;
    const uint32_t n_tiles = static_cast<uint32_t>((static_cast<uint64_t>(first - base) + n + 31) >> 5);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t tile = warp; tile < n_tiles; tile += n_warps) {
        const uint32_t g = base + tile * 32u + lane;
        const uint32_t word = adv_bitmap ? adv_bitmap[g >> 5] : 0xffffffffu;
        const bool sel = g >= first && g < first + n && ((word >> lane) & 1u);
        uint32_t send = 0;
        uint64_t nx[kSlots], prs[kSlots];  // loaded together with the flag bytes: one round trip, not one per entry
        if (sel) {
            const uint32_t meta = c.meta[g];
            uint32_t peers = RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta) | RAFTGPU_META_LEARN(meta);
            if (meta & RAFTGPU_META_HAS_SELF) peers &= ~(1u << RAFTGPU_META_SELF(meta));  // raft.rs:863 id != self_id
            uint32_t f[kSlots];
#pragma unroll
            for (int s = 0; s < kSlots; s++) {
                f[s] = RAFTGPU_STATE_SNAPSHOT;
                nx[s] = 0;
                prs[s] = 0;
                if ((peers >> s) & 1u) {
                    const size_t cell = static_cast<size_t>(s) * c.cap + g;
                    f[s] = c.pflags[cell];
                    nx[s] = c.next_idx[cell];
                    prs[s] = c.pending_req_snapshot[cell];
                }
            }
#pragma unroll
            for (int s = 0; s < kSlots; s++) {
                const uint32_t state = f[s] & RAFTGPU_PF_STATE_MASK;
                const bool paused = state == RAFTGPU_STATE_PROBE ? (f[s] & RAFTGPU_PF_PAUSED) != 0
                                                                 : (state == RAFTGPU_STATE_REPLICATE ? (f[s] & RAFTGPU_PF_INS_FULL) != 0 : true);
                if (((peers >> s) & 1u) && !paused) send |= 1u << s;
            }
        }
        const uint32_t cnt = __popc(send);
        uint32_t incl = cnt;  // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= static_cast<uint32_t>(d)) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long pos = 0;
        if (lane == 31) pos = atomicAdd(count, static_cast<unsigned long long>(total));
        pos = __shfl_sync(0xffffffffu, pos, 31) + (incl - cnt);
#pragma unroll
        for (int s = 0; s < kSlots; s++) {
            if (!((send >> s) & 1u)) continue;
            if (pos < capacity) {
                const uint64_t w0 = static_cast<uint64_t>(g) | (static_cast<uint64_t>(s) << 32) |
                                    (static_cast<uint64_t>(prs[s] != RAFTGPU_INVALID_INDEX ? RAFTGPU_SEND_SNAPSHOT : 0u) << 40);
                reinterpret_cast<ulonglong2 *>(out)[pos] = make_ulonglong2(w0, nx[s]);
            }
            pos++;
        }
    }
}

// ---------------------------------------------------------------------------
// tally_kernel: ProgressTracker::tally_votes (tracker.rs:313-340) per group:
// granted / rejected over voters, JointConfig::vote_result (joint.rs:56-67) over
// MajorityConfig::vote_result (majority.rs:130-154).
__device__ __forceinline__ uint32_t majority_vote(uint32_t mask, uint32_t yes, uint32_t no) {
    if (mask == 0) return RAFTGPU_VOTE_WON;  // majority.rs:131-136
    const uint32_t n = __popc(mask), q = (n >> 1) + 1;
    const uint32_t y = __popc(yes & mask), missing = n - y - __popc(no & mask);
    if (y >= q) return RAFTGPU_VOTE_WON;
    if (y + missing >= q) return RAFTGPU_VOTE_PENDING;
    return RAFTGPU_VOTE_LOST;
}

__global__ void __launch_bounds__(256)
tally_kernel(Columns c, uint32_t first, uint32_t n, uint32_t *__restrict__ out,
             unsigned long long *__restrict__ counters) {
    const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool active = t < n;
    if (active) {
        const uint32_t g = first + static_cast<uint32_t>(t);
        const uint32_t meta = c.meta[g];
        const uint32_t in = RAFTGPU_META_IN(meta), outm = RAFTGPU_META_OUT(meta);
        uint32_t yes = 0, no = 0;
#pragma unroll
        for (int s = 0; s < kSlots; s++) {
            const uint32_t v = c.votes[static_cast<size_t>(s) * c.cap + g];
            yes |= (v == 2u) << s;
            no |= (v == 1u) << s;
        }
        const uint32_t i = majority_vote(in, yes, no), o = majority_vote(outm, yes, no);
        uint32_t r;
        if (i == RAFTGPU_VOTE_WON && o == RAFTGPU_VOTE_WON)
            r = RAFTGPU_VOTE_WON;
        else if (i == RAFTGPU_VOTE_LOST || o == RAFTGPU_VOTE_LOST)
            r = RAFTGPU_VOTE_LOST;
        else
            r = RAFTGPU_VOTE_PENDING;
        const uint32_t voters = in | outm;  // tracker.rs:320-322
        out[g] = r | (__popc(yes & voters) << 8) | (__popc(no & voters) << 16);
    }
    const uint32_t local[1] = {active ? 1u : 0u};
    const int which[1] = {kCntVotes};
    block_flush_counts<1>(local, which, counters, nullptr);
}

// ---------------------------------------------------------------------------
// Control-plane helpers (single thread; launched <<<1,1>>>).

// ProgressTracker::apply_conf (tracker.rs:380-397)
__global__ void conf_kernel(Columns c, uint32_t g, uint32_t new_meta, uint32_t added,
                            uint32_t removed, uint64_t next_idx) {
    for (int s = 0; s < kSlots; s++) {
        const size_t cell = static_cast<size_t>(s) * c.cap + g;
        if (((added | removed) >> s) & 1u) {
            const bool add = (added >> s) & 1u;
            c.matched[cell] = 0;
            c.next_idx[cell] = add ? next_idx : 0;  // Progress::new(next_idx, ..), progress.rs:60-73
            c.peer_committed[cell] = 0;
            c.pending_snapshot[cell] = 0;
            c.pending_req_snapshot[cell] = 0;
            c.commit_group_id[cell] = 0;
            c.pflags[cell] = add ? RAFTGPU_PF_RECENT_ACTIVE : 0;  // tracker.rs:385-389
            c.votes[cell] = 0;
        }
    }
    c.meta[g] = new_meta;
}

// Raft::reset (raft.rs:942-971) for the tracker + log bookkeeping of one group.
__global__ void reset_kernel(Columns c, uint32_t g, uint64_t term_start, uint64_t last_index,
                             uint64_t committed, uint64_t persisted) {
    const uint32_t meta = c.meta[g];
    const uint32_t present = RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta) | RAFTGPU_META_LEARN(meta);
    for (int s = 0; s < kSlots; s++) {
        if (!((present >> s) & 1u)) continue;
        const size_t cell = static_cast<size_t>(s) * c.cap + g;
        // Progress::reset(last_index + 1), progress.rs:82-92
        c.matched[cell] = 0;
        c.next_idx[cell] = last_index + 1;
        c.pending_snapshot[cell] = 0;
        c.pending_req_snapshot[cell] = RAFTGPU_INVALID_INDEX;
        c.pflags[cell] = RAFTGPU_STATE_PROBE;
        c.votes[cell] = 0;  // prs.reset_votes(), raft.rs:953
        if ((meta & RAFTGPU_META_HAS_SELF) && RAFTGPU_META_SELF(meta) == static_cast<uint32_t>(s)) {
            c.matched[cell] = persisted;         // raft.rs:967
            c.peer_committed[cell] = committed;  // raft.rs:968
        }
    }
    c.committed[g] = committed;
    c.term_start[g] = term_start;
    c.last_index[g] = last_index;
}

// Raft::become_leader's tracker side (raft.rs:1176-1192): self.become_replicate(),
// then the empty entry of the new term is appended at last_index + 1.
__global__ void become_leader_kernel(Columns c, uint32_t g) {
    const uint32_t meta = c.meta[g];
    if (meta & RAFTGPU_META_HAS_SELF) {
        const size_t cell = static_cast<size_t>(RAFTGPU_META_SELF(meta)) * c.cap + g;
        c.pflags[cell] = static_cast<uint8_t>(
            (c.pflags[cell] & ~(RAFTGPU_PF_PAUSED | RAFTGPU_PF_INS_FULL | RAFTGPU_PF_STATE_MASK)) |
            RAFTGPU_STATE_REPLICATE);
        c.pending_snapshot[cell] = 0;
        c.next_idx[cell] = c.matched[cell] + 1;  // progress.rs:110-114
    }
    const uint64_t li = c.last_index[g] + 1;  // raft.rs:1192 append_entry(&mut [Entry::default()])
    c.last_index[g] = li;
    c.term_start[g] = li;
}

__global__ void progress_get_kernel(Columns c, uint32_t g, uint32_t s, raftgpu_progress *out) {
    const size_t cell = static_cast<size_t>(s) * c.cap + g;
    const uint32_t f = c.pflags[cell];
    const uint32_t meta = c.meta[g];
    raftgpu_progress p{};
    p.matched = c.matched[cell];
    p.next_idx = c.next_idx[cell];
    p.pending_snapshot = c.pending_snapshot[cell];
    p.pending_request_snapshot = c.pending_req_snapshot[cell];
    p.commit_group_id = c.commit_group_id[cell];
    p.committed_index = c.peer_committed[cell];
    p.state = f & RAFTGPU_PF_STATE_MASK;
    p.paused = (f & RAFTGPU_PF_PAUSED) != 0;
    p.recent_active = (f & RAFTGPU_PF_RECENT_ACTIVE) != 0;
    p.ins_full = (f & RAFTGPU_PF_INS_FULL) != 0;
    p.present = ((RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta) | RAFTGPU_META_LEARN(meta)) >> s) & 1u;
    *out = p;
}

__global__ void progress_set_kernel(Columns c, uint32_t g, uint32_t s, raftgpu_progress p) {
    const size_t cell = static_cast<size_t>(s) * c.cap + g;
    c.matched[cell] = p.matched;
    c.next_idx[cell] = p.next_idx;
    c.pending_snapshot[cell] = p.pending_snapshot;
    c.pending_req_snapshot[cell] = p.pending_request_snapshot;
    c.commit_group_id[cell] = p.commit_group_id;
    c.peer_committed[cell] = p.committed_index;
    c.pflags[cell] = static_cast<uint8_t>((p.state & RAFTGPU_PF_STATE_MASK) |
                                          (p.paused ? RAFTGPU_PF_PAUSED : 0) |
                                          (p.recent_active ? RAFTGPU_PF_RECENT_ACTIVE : 0) |
                                          (p.ins_full ? RAFTGPU_PF_INS_FULL : 0));
}

__device__ __forceinline__ uint32_t majority_vote(uint32_t mask, uint32_t yes, uint32_t no);

// Every method of Progress (src/tracker/progress.rs:75-243) on one cell, literally, for the
// host mirror's ProgressRef: op codes are RAFTGPU_POP_*.  *out gets the bool / status result.
__global__ void progress_op_kernel(Columns c, uint32_t g, uint32_t s, int op, uint64_t a0, uint64_t a1,
                                   uint64_t a2, int32_t *out) {
    const size_t cell = static_cast<size_t>(s) * c.cap + g;
    uint64_t matched = c.matched[cell], next = c.next_idx[cell];
    uint32_t f = c.pflags[cell];
    const uint32_t state = f & RAFTGPU_PF_STATE_MASK;
    int32_t ret = 0;
    auto reset_st = [&](uint32_t st) {  // progress.rs:75-80
        f = (f & ~(RAFTGPU_PF_PAUSED | RAFTGPU_PF_INS_FULL | RAFTGPU_PF_STATE_MASK)) | st;
        c.pending_snapshot[cell] = 0;
    };
    switch (op) {
    case RAFTGPU_POP_MAYBE_UPDATE:  // progress.rs:138-150
        if (matched < a0) {
            matched = a0;
            f &= ~RAFTGPU_PF_PAUSED;
            ret = 1;
        }
        if (next < a0 + 1) next = a0 + 1;
        break;
    case RAFTGPU_POP_MAYBE_DECR_TO: {  // progress.rs:168-206 (a0 rejected, a1 match_hint, a2 request_snapshot)
        if (state == RAFTGPU_STATE_REPLICATE) {
            if (a0 < matched || (a0 == matched && a2 == RAFTGPU_INVALID_INDEX)) break;
            if (a2 == RAFTGPU_INVALID_INDEX)
                next = matched + 1;
            else
                c.pending_req_snapshot[cell] = a2;
            ret = 1;
            break;
        }
        if ((next == 0 || next - 1 != a0) && a2 == RAFTGPU_INVALID_INDEX) break;
        if (a2 == RAFTGPU_INVALID_INDEX) {
            next = umin64(a0, a1 + 1);
            if (next < 1) next = 1;
        } else if (c.pending_req_snapshot[cell] == RAFTGPU_INVALID_INDEX) {
            c.pending_req_snapshot[cell] = a2;
        }
        f &= ~RAFTGPU_PF_PAUSED;
        ret = 1;
        break;
    }
    case RAFTGPU_POP_UPDATE_COMMITTED:  // progress.rs:153-157
        if (a0 > c.peer_committed[cell]) c.peer_committed[cell] = a0;
        break;
    case RAFTGPU_POP_OPTIMISTIC_UPDATE:  // progress.rs:160-163
        next = a0 + 1;
        break;
    case RAFTGPU_POP_BECOME_PROBE:  // progress.rs:95-107
        if (state == RAFTGPU_STATE_SNAPSHOT) {
            const uint64_t pending = c.pending_snapshot[cell];
            reset_st(RAFTGPU_STATE_PROBE);
            next = umax64(matched + 1, pending + 1);
        } else {
            reset_st(RAFTGPU_STATE_PROBE);
            next = matched + 1;
        }
        break;
    case RAFTGPU_POP_BECOME_REPLICATE:  // progress.rs:110-114
        reset_st(RAFTGPU_STATE_REPLICATE);
        next = matched + 1;
        break;
    case RAFTGPU_POP_BECOME_SNAPSHOT:  // progress.rs:117-121
        reset_st(RAFTGPU_STATE_SNAPSHOT);
        c.pending_snapshot[cell] = a0;
        break;
    case RAFTGPU_POP_SNAPSHOT_FAILURE:  // progress.rs:124-127
        c.pending_snapshot[cell] = 0;
        break;
    case RAFTGPU_POP_MAYBE_SNAPSHOT_ABORT:  // progress.rs:131-134
        ret = state == RAFTGPU_STATE_SNAPSHOT && matched >= c.pending_snapshot[cell];
        break;
    case RAFTGPU_POP_IS_PAUSED:  // progress.rs:210-216
        ret = state == RAFTGPU_STATE_PROBE ? (f & RAFTGPU_PF_PAUSED) != 0
              : state == RAFTGPU_STATE_REPLICATE ? (f & RAFTGPU_PF_INS_FULL) != 0 : 1;
        break;
    case RAFTGPU_POP_RESUME:  // progress.rs:219-222
        f &= ~RAFTGPU_PF_PAUSED;
        break;
    case RAFTGPU_POP_PAUSE:  // progress.rs:225-228
        f |= RAFTGPU_PF_PAUSED;
        break;
    case RAFTGPU_POP_UPDATE_STATE:  // progress.rs:231-243 (a0 = last); -1 where the reference panics
        if (state == RAFTGPU_STATE_REPLICATE)
            next = a0 + 1;  // optimistic_update; ins.add(last) is the host's
        else if (state == RAFTGPU_STATE_PROBE)
            f |= RAFTGPU_PF_PAUSED;
        else
            ret = -1;
        break;
    case RAFTGPU_POP_RESET:  // progress.rs:82-92 (a0 = next_idx)
        matched = 0;
        next = a0;
        f = RAFTGPU_STATE_PROBE;
        c.pending_snapshot[cell] = 0;
        c.pending_req_snapshot[cell] = RAFTGPU_INVALID_INDEX;
        break;
    default:
        ret = -2;
    }
    c.matched[cell] = matched;
    c.next_idx[cell] = next;
    c.pflags[cell] = static_cast<uint8_t>(f);
    *out = ret;
}

// ProgressTracker::has_quorum (tracker.rs:367-372): vote_result(|id| set.get(id).map(|_| true)) == Won,
// and quorum_recently_active (tracker.rs:346-361), which also clears recent_active.
__global__ void quorum_kernel(Columns c, uint32_t g, int op, uint32_t arg, int32_t *out) {
    const uint32_t meta = c.meta[g];
    const uint32_t in = RAFTGPU_META_IN(meta), outm = RAFTGPU_META_OUT(meta);
    uint32_t active = arg;
    if (op == 1) {  // quorum_recently_active(perspective_of = slot arg)
        const uint32_t present = in | outm | RAFTGPU_META_LEARN(meta);
        active = 0;
        for (int s = 0; s < kSlots; s++) {
            if (!((present >> s) & 1u)) continue;
            uint8_t *f = &c.pflags[static_cast<size_t>(s) * c.cap + g];
            if (static_cast<uint32_t>(s) == arg) {
                *f |= RAFTGPU_PF_RECENT_ACTIVE;  // tracker.rs:350-352
                active |= 1u << s;
            } else if (*f & RAFTGPU_PF_RECENT_ACTIVE) {
                active |= 1u << s;  // tracker.rs:353-358
                *f &= ~RAFTGPU_PF_RECENT_ACTIVE;
            }
        }
    }
    // members of the set vote yes, everyone else is missing (None)
    const uint32_t i = majority_vote(in, active, 0), o = majority_vote(outm, active, 0);
    *out = (i == RAFTGPU_VOTE_WON && o == RAFTGPU_VOTE_WON) ? 1 : 0;
}

__global__ void group_get_kernel(Columns c, uint32_t g, raftgpu_group_state *out) {
    raftgpu_group_state s{};
    s.meta = c.meta[g];
    s.committed = c.committed[g];
    s.term_start = c.term_start[g];
    s.last_index = c.last_index[g];
    *out = s;
}

// op 0: set_log_bounds; op 1: commit_to (status in *out: 0 ok, 1 out of range);
// op 2: meta bit set/clear (a = mask, b = enable); op 3: assign commit group (a = slot, b = id);
// op 4: reset votes; op 5: record vote (a = slot, b = vote+1), first vote wins (tracker.rs:308-310)
__global__ void group_op_kernel(Columns c, uint32_t g, int op, uint64_t a, uint64_t b,
                                uint32_t *out) {
    switch (op) {
    case 0:
        c.term_start[g] = a;
        c.last_index[g] = b;
        break;
    case 1:  // RaftLog::commit_to, raft_log.rs:286-300
        if (c.committed[g] >= a) {
            *out = 0;
        } else if (c.last_index[g] < a) {
            *out = 1;
        } else {
            c.committed[g] = a;
            *out = 0;
        }
        break;
    case 2:
        c.meta[g] = b ? (c.meta[g] | static_cast<uint32_t>(a)) : (c.meta[g] & ~static_cast<uint32_t>(a));
        break;
    case 3:
        c.commit_group_id[static_cast<size_t>(a) * c.cap + g] = b;
        break;
    case 4:
        for (int s = 0; s < kSlots; s++) c.votes[static_cast<size_t>(s) * c.cap + g] = 0;
        break;
    case 5: {
        uint8_t *v = &c.votes[static_cast<size_t>(a) * c.cap + g];
        if (*v == 0) *v = static_cast<uint8_t>(b);  // entry(id).or_insert(vote)
        break;
    }
    case 6:  // RaftLog::maybe_commit(max_index = a, term = the leader's), raft_log.rs:487-499, range form
        if (a > c.committed[g] && a >= c.term_start[g] && a <= c.last_index[g]) {
            c.committed[g] = a;
            *out = 1;
        } else {
            *out = 0;
        }
        break;
    }
}

}  // namespace raftgpu
