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
// recompute_kernel: one Raft::maybe_commit (raft.rs:893-904) per group.
//   mci  = ProgressTracker::maximal_committed_index      tracker.rs:294-298
//        = min(incoming.committed_index, outgoing.committed_index)   joint.rs:47-51
//   if mci > committed && term(mci) == term              raft_log.rs:487-499
//        committed = mci; prs[self].update_committed      raft.rs:896-900
// term(mci) == term is the range test term_start <= mci <= last_index.
//
// Persistent grid: each warp walks 32-group tiles (one word of the advanced
// bitmap) with a grid stride.  `hint` is a superset guess of the voter slots in
// this arena (the host keeps the union of all voter masks): the matched loads of
// the hinted slots are issued together with meta / committed / term_start /
// last_index, so a tile costs ONE round trip to HBM instead of two (meta first,
// then the slots it names).  Voter slots outside the hint are fetched after
// meta arrives -- correct for any hint, fast for a tight one.
// Algorithmic bytes per group: 8K (matched) + 4 (meta) + 24 (committed,
// term_start, last_index) read, 8 written when advanced.
__global__ void __launch_bounds__(256)
recompute_kernel(Columns c, uint32_t first, uint32_t n, uint32_t hint,
                 uint32_t *__restrict__ adv_bitmap, uint64_t *__restrict__ commit_out,
                 uint64_t *__restrict__ mci_out, uint8_t *__restrict__ gc_out,
                 uint32_t *__restrict__ step_advanced, unsigned long long *__restrict__ counters) {
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

            const uint32_t in = RAFTGPU_META_IN(meta), out = RAFTGPU_META_OUT(meta);
            const uint32_t voters = in | out;
            const uint32_t missing = voters & ~hint;
            if (missing) {  // hint was too small for this group: second trip for the rest
#pragma unroll
                for (int s = 0; s < kSlots; s++)
                    if ((missing >> s) & 1u) v[s] = c.matched[static_cast<size_t>(s) * c.cap + g];
            }
            uint64_t mci;
            bool use_gc;
            if ((meta & (0xffffu | RAFTGPU_META_GROUP_COMMIT)) == 0x1fu) {
                // 5 voters in slots 0..4, no joint config, no group commit: q = 3 = the median
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
                    gid[s] = ((voters >> s) & 1u)
                                 ? c.commit_group_id[static_cast<size_t>(s) * c.cap + g] : 0ull;
                uint64_t i_idx, o_idx;
                bool i_gc, o_gc;
                majority_group_commit(v, gid, in, &i_idx, &i_gc);
                majority_group_commit(v, gid, out, &o_idx, &o_gc);
                mci = umin64(i_idx, o_idx);
                use_gc = i_gc && o_gc;
            }
            if (mci_out) mci_out[g] = mci;
            if (gc_out) gc_out[g] = use_gc ? 1 : 0;

            // RaftLog::maybe_commit, raft_log.rs:488, in range form.
            advanced = mci > committed && mci >= term_start && mci <= last_index;
            if (advanced) {
                c.committed[g] = mci;  // commit_to: mci <= last_index, never the fatal! branch
                if (commit_out) commit_out[g] = mci;
                if (meta & RAFTGPU_META_HAS_SELF) {  // raft.rs:896-900
                    const size_t cell = static_cast<size_t>(RAFTGPU_META_SELF(meta)) * c.cap + g;
                    if (mci > c.peer_committed[cell]) c.peer_committed[cell] = mci;
                }
            }
        }
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
    if (*pending_snapshot != 0) *pending_snapshot = 0;
}

// apply_kernel: the per-message prefix of Raft::handle_append_response
// (raft.rs:1663-1743) for one wave of records, one thread per record, persistent
// grid-stride loop.  Within a wave every (group, peer) cell is touched by at most
// one record, so threads never race on a cell and no atomics are needed on the
// columns.  The record names its cell, so meta and the cell's matched / next_idx /
// pflags / committed_index are fetched in ONE batch of independent loads.
// Algorithmic bytes per record: 24 (record) + RMW of matched, next_idx,
// committed_index (48) + flag byte and meta (~4) = 76.
__global__ void __launch_bounds__(256)
apply_kernel(Columns c, const raftgpu_append_resp *__restrict__ recs, uint64_t n,
             uint8_t *__restrict__ results, unsigned long long *__restrict__ counters) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint32_t local[5] = {0, 0, 0, 0, 0};  // records, updates, rejects, decrements, no_progress
    // keep whole warps in the loop so the flush below runs converged
    const uint64_t n_pad = (n + 31) & ~31ull;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        if (i >= n) continue;
        const uint64_t *p = reinterpret_cast<const uint64_t *>(recs + i);
        const uint64_t w0 = p[0];
        const uint64_t index = p[1];
        const uint64_t commit = p[2];
        const uint32_t g = static_cast<uint32_t>(w0);
        const uint32_t slot = static_cast<uint32_t>(w0 >> 32) & 0xffu;
        const uint32_t rflags = static_cast<uint32_t>(w0 >> 40) & 0xffu;
        uint32_t res = 0;
        if (!(rflags & RAFTGPU_REC_EXT)) {
            local[0]++;
            const bool in_range = g < c.cap && slot < kSlots;
            const size_t cell = in_range ? static_cast<size_t>(slot) * c.cap + g : 0;
            // one batch of independent loads (cell 0 / group 0 is a harmless stand-in when
            // the record is out of range; nothing is written in that case)
            const uint32_t meta = c.meta[in_range ? g : 0];
            Cell pr;
            pr.matched = c.matched[cell];
            pr.next_idx = c.next_idx[cell];
            pr.flags = c.pflags[cell];
            const uint64_t peer_committed = c.peer_committed[cell];
            const uint32_t present =
                RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta) | RAFTGPU_META_LEARN(meta);
            if (!in_range || !((present >> slot) & 1u)) {
                // raft.rs:1663-1673: no progress available for m.from
                local[4]++;
                res = RAFTGPU_RES_NO_PROGRESS;
            } else {
                const uint64_t matched0 = pr.matched, next0 = pr.next_idx;
                const uint32_t flags0 = pr.flags;
                const uint32_t state = pr.flags & RAFTGPU_PF_STATE_MASK;

                if (rflags & RAFTGPU_REC_LOCAL) {
                    // raft.rs:974-991 append_entry: last_index grew
                    if (commit != 0) c.last_index[g] = commit;
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
                    if (commit > peer_committed) c.peer_committed[cell] = commit;

                    if (rflags & RAFTGPU_REC_REJECT) {
                        local[2]++;
                        uint64_t hint = 0, request_snapshot = RAFTGPU_INVALID_INDEX;
                        if (i + 1 < n) {
                            const uint64_t *e = reinterpret_cast<const uint64_t *>(recs + i + 1);
                            if ((e[0] >> 40) & RAFTGPU_REC_EXT) {
                                hint = e[1];
                                request_snapshot = e[2];
                            }
                        }
                        // Progress::maybe_decr_to, progress.rs:168-206
                        bool ok;
                        if (state == RAFTGPU_STATE_REPLICATE) {
                            if (index < pr.matched || (index == pr.matched &&
                                                       request_snapshot == RAFTGPU_INVALID_INDEX)) {
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
                                : (state == RAFTGPU_STATE_REPLICATE
                                       ? (pr.flags & RAFTGPU_PF_INS_FULL) != 0
                                       : true);
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
                if (pr.matched != matched0) c.matched[cell] = pr.matched;
                if (pr.next_idx != next0) c.next_idx[cell] = pr.next_idx;
                if (pr.flags != flags0) c.pflags[cell] = static_cast<uint8_t>(pr.flags);
            }
        }
        if (results) results[i] = static_cast<uint8_t>(res);
    }
    const int which[5] = {kCntRecords, kCntUpdates, kCntRejects, kCntDecrements, kCntNoProgress};
    block_flush_counts<5>(local, which, counters, nullptr);
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
    }
}

}  // namespace raftgpu
