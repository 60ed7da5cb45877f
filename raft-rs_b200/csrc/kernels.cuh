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
// Non-members are zeroed, which leaves the top-q ranks of the members intact
// (q <= n), so the selection runs over all 8 slots branch-free: the answer is
// the largest value that at least q slots are >= to.
__device__ __forceinline__ uint64_t quorum_index(const uint64_t (&v)[kSlots], uint32_t mask) {
    if (mask == 0) return UINT64_MAX;
    const uint32_t q = (static_cast<uint32_t>(__popc(mask)) >> 1) + 1;
    uint64_t w[kSlots];
#pragma unroll
    for (int i = 0; i < kSlots; i++) w[i] = ((mask >> i) & 1u) ? v[i] : 0ull;
    uint64_t best = 0;
#pragma unroll
    for (int i = 0; i < kSlots; i++) {
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < kSlots; j++) cnt += (w[j] >= w[i]) ? 1u : 0u;
        if (cnt >= q && w[i] > best) best = w[i];
    }
    return best;
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

// Warp-aggregated counter bump: one atomic per warp.
__device__ __forceinline__ void warp_count(unsigned long long *counters, int which, bool pred) {
    const unsigned ballot = __ballot_sync(0xffffffffu, pred);
    if ((threadIdx.x & 31) == 0 && ballot != 0)
        atomicAdd(&counters[which], static_cast<unsigned long long>(__popc(ballot)));
}

// ---------------------------------------------------------------------------
// recompute_kernel: one Raft::maybe_commit (raft.rs:893-904) per group.
//   mci  = ProgressTracker::maximal_committed_index      tracker.rs:294-298
//        = min(incoming.committed_index, outgoing.committed_index)   joint.rs:47-51
//   if mci > committed && term(mci) == term              raft_log.rs:487-499
//        committed = mci; prs[self].update_committed      raft.rs:896-900
// term(mci) == term is the range test term_start <= mci <= last_index.
//
// One thread per group; thread t of the grid handles group (first & ~31) + t so
// that a warp always covers exactly one 32-bit word of the advanced bitmap.
// Algorithmic bytes per group: 8K (matched) + 4 (meta) + 24 (committed,
// term_start, last_index) read, 8 written when advanced.
__global__ void __launch_bounds__(256)
recompute_kernel(Columns c, uint32_t first, uint32_t n, uint32_t *__restrict__ adv_bitmap,
                 uint64_t *__restrict__ commit_out, uint64_t *__restrict__ mci_out,
                 uint8_t *__restrict__ gc_out, uint32_t *__restrict__ step_advanced,
                 unsigned long long *__restrict__ counters) {
    const uint32_t base = first & ~31u;
    const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t g64 = base + t;
    const bool active = g64 >= first && g64 < static_cast<uint64_t>(first) + n;
    const uint32_t g = static_cast<uint32_t>(g64);

    bool advanced = false;
    if (active) {
        const uint32_t meta = c.meta[g];
        const uint64_t committed = c.committed[g];
        const uint64_t term_start = c.term_start[g];
        const uint64_t last_index = c.last_index[g];
        uint64_t mci;
        bool use_gc;
        group_mci(c, g, meta, mci, use_gc);
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
    if ((threadIdx.x & 31) == 0 && act != 0) {
        if (adv_bitmap) {
            uint32_t *word = &adv_bitmap[g64 >> 5];
            if (act == 0xffffffffu) {
                *word = adv;
            } else {  // range starts / ends inside this word: leave the other bits alone
                atomicAnd(word, ~act);
                if (adv) atomicOr(word, adv);
            }
        }
        atomicAdd(&counters[kCntRecomputes], static_cast<unsigned long long>(__popc(act)));
        if (adv) {
            atomicAdd(&counters[kCntAdvanced], static_cast<unsigned long long>(__popc(adv)));
            if (step_advanced) atomicAdd(step_advanced, static_cast<uint32_t>(__popc(adv)));
        }
    }
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
// (raft.rs:1663-1743) for one wave of records, one thread per record.  Within a
// wave every (group, peer) cell is touched by at most one record, so threads
// never race on a cell and no atomics are needed on the columns.
__global__ void __launch_bounds__(256)
apply_kernel(Columns c, const raftgpu_append_resp *__restrict__ recs, uint64_t n,
             uint8_t *__restrict__ results, unsigned long long *__restrict__ counters) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool is_record = false, updated = false, is_reject = false, decremented = false,
         no_progress = false;
    if (i < n) {
        const uint64_t *p = reinterpret_cast<const uint64_t *>(recs + i);
        const uint64_t w0 = p[0];
        const uint64_t index = p[1];
        const uint64_t commit = p[2];
        const uint32_t g = static_cast<uint32_t>(w0);
        const uint32_t slot = static_cast<uint32_t>(w0 >> 32) & 0xffu;
        const uint32_t rflags = static_cast<uint32_t>(w0 >> 40) & 0xffu;
        uint32_t res = 0;
        if (!(rflags & RAFTGPU_REC_EXT)) {
            is_record = true;
            uint32_t present = 0;
            if (g < c.cap && slot < kSlots) {
                const uint32_t meta = c.meta[g];
                present = RAFTGPU_META_IN(meta) | RAFTGPU_META_OUT(meta) | RAFTGPU_META_LEARN(meta);
            }
            if (!((present >> slot) & 1u)) {
                // raft.rs:1663-1673: no progress available for m.from
                no_progress = true;
                res = RAFTGPU_RES_NO_PROGRESS;
            } else {
                const size_t cell = static_cast<size_t>(slot) * c.cap + g;
                Cell pr;
                pr.matched = c.matched[cell];
                pr.next_idx = c.next_idx[cell];
                pr.flags = c.pflags[cell];
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
                        updated = true;
                        res = RAFTGPU_RES_OK;
                    }
                    if (pr.next_idx < index + 1) pr.next_idx = index + 1;
                } else {
                    pr.flags |= RAFTGPU_PF_RECENT_ACTIVE;  // raft.rs:1674
                    // raft.rs:1677 pr.update_committed(m.commit), progress.rs:153-157
                    if (commit > c.peer_committed[cell]) c.peer_committed[cell] = commit;

                    if (rflags & RAFTGPU_REC_REJECT) {
                        is_reject = true;
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
                            decremented = true;
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
                            updated = true;
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
    warp_count(counters, kCntRecords, is_record);
    warp_count(counters, kCntUpdates, updated);
    warp_count(counters, kCntRejects, is_reject);
    warp_count(counters, kCntDecrements, decremented);
    warp_count(counters, kCntNoProgress, no_progress);
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
    warp_count(counters, kCntVotes, active);
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
