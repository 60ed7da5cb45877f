/*
 * raft_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see raft_oracle.h).
 *
 * Restates, function by function, the reference's quorum / progress / commit
 * logic on plain arrays.  Each function names the reference lines it follows.
 */
#define _POSIX_C_SOURCE 200809L
#include "raft_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------ */
/* src/util.rs:118-120                                                       */
size_t ro_majority(size_t total) { return total / 2 + 1; }

/* src/quorum.rs:69-74 (HashMap::get(...).cloned()) */
int ro_acked_index(const ro_ack_indexer *l, uint64_t voter, ro_index *out) {
    for (size_t i = 0; i < l->n; i++) {
        if (l->ids[i] == voter) {
            *out = l->idx[i];
            return 1;
        }
    }
    return 0;
}

/* Stable descending sort by .index == `matched.sort_by(|a, b| b.index.cmp(&a.index))`
 * (majority.rs:95; Rust's slice::sort_by is stable).  Insertion sort is stable. */
static void sort_desc_stable(ro_index *a, size_t n) {
    for (size_t i = 1; i < n; i++) {
        ro_index x = a[i];
        size_t j = i;
        while (j > 0 && a[j - 1].index < x.index) {
            a[j] = a[j - 1];
            j--;
        }
        a[j] = x;
    }
}

/* src/quorum/majority.rs:70-124 */
void ro_majority_committed_index(const uint64_t *voters, size_t n_voters, int use_group_commit,
                                 const ro_ack_indexer *l, uint64_t *out_index, int *out_use_gc) {
    /* :71-75 empty config commits "everything" */
    if (n_voters == 0) {
        *out_index = UINT64_MAX;
        *out_use_gc = 1;
        return;
    }
    /* :77-93 gather acked indexes; a voter without an entry counts as
     * Index::default() = (0, 0) (`unwrap_or_default`).  The reference uses a
     * 7-slot stack array or a Vec; one heap buffer covers both here. */
    ro_index stack_arr[7];
    ro_index *matched = stack_arr;
    ro_index *heap_arr = NULL;
    if (n_voters > 7) {
        heap_arr = (ro_index *)malloc(n_voters * sizeof(ro_index));
        matched = heap_arr;
    }
    for (size_t i = 0; i < n_voters; i++) {
        ro_index ix = {0, 0};
        if (!ro_acked_index(l, voters[i], &ix)) {
            ix.index = 0;
            ix.group_id = 0;
        }
        matched[i] = ix;
    }
    /* :95 reverse (stable) sort */
    sort_desc_stable(matched, n_voters);
    /* :97-101 */
    size_t quorum = ro_majority(n_voters);
    ro_index quorum_index = matched[quorum - 1];
    if (!use_group_commit) {
        *out_index = quorum_index.index;
        *out_use_gc = 0;
        free(heap_arr);
        return;
    }
    /* :102-123 group commit */
    uint64_t quorum_commit_index = quorum_index.index;
    uint64_t checked_group_id = quorum_index.group_id;
    int single_group = 1;
    for (size_t i = 0; i < n_voters; i++) {
        const ro_index *m = &matched[i];
        if (m->group_id == 0) {
            single_group = 0;
            continue;
        }
        if (checked_group_id == 0) {
            checked_group_id = m->group_id;
            continue;
        }
        if (checked_group_id == m->group_id) {
            continue;
        }
        *out_index = m->index < quorum_commit_index ? m->index : quorum_commit_index;
        *out_use_gc = 1;
        free(heap_arr);
        return;
    }
    if (single_group) {
        *out_index = quorum_commit_index;
        *out_use_gc = 0;
    } else {
        *out_index = matched[n_voters - 1].index;
        *out_use_gc = 0;
    }
    free(heap_arr);
}

/* src/quorum/joint.rs:47-51 */
void ro_joint_committed_index(const uint64_t *incoming, size_t n_in, const uint64_t *outgoing,
                              size_t n_out, int use_group_commit, const ro_ack_indexer *l,
                              uint64_t *out_index, int *out_use_gc) {
    uint64_t i_idx, o_idx;
    int i_gc, o_gc;
    ro_majority_committed_index(incoming, n_in, use_group_commit, l, &i_idx, &i_gc);
    ro_majority_committed_index(outgoing, n_out, use_group_commit, l, &o_idx, &o_gc);
    *out_index = i_idx < o_idx ? i_idx : o_idx;
    *out_use_gc = i_gc && o_gc;
}

/* check(id) -> Option<bool>: 2 = Some(true), 1 = Some(false), 0 = None */
static int vote_check(const ro_vote_map *votes, uint64_t id) {
    for (size_t i = 0; i < votes->n; i++) {
        if (votes->ids[i] == id) return votes->vote[i] ? 2 : 1;
    }
    return 0;
}

/* src/quorum/majority.rs:130-154 */
int ro_majority_vote_result(const uint64_t *voters, size_t n_voters, const ro_vote_map *votes) {
    if (n_voters == 0) return RO_VOTE_WON; /* :131-136 */
    size_t yes = 0, missing = 0;
    for (size_t i = 0; i < n_voters; i++) { /* :138-145 */
        int c = vote_check(votes, voters[i]);
        if (c == 2)
            yes++;
        else if (c == 0)
            missing++;
    }
    size_t q = ro_majority(n_voters); /* :146-153 */
    if (yes >= q) return RO_VOTE_WON;
    if (yes + missing >= q) return RO_VOTE_PENDING;
    return RO_VOTE_LOST;
}

/* src/quorum/joint.rs:56-67 */
int ro_joint_vote_result(const uint64_t *incoming, size_t n_in, const uint64_t *outgoing,
                         size_t n_out, const ro_vote_map *votes) {
    int i = ro_majority_vote_result(incoming, n_in, votes);
    int o = ro_majority_vote_result(outgoing, n_out, votes);
    if (i == RO_VOTE_WON && o == RO_VOTE_WON) return RO_VOTE_WON;
    if (i == RO_VOTE_LOST || o == RO_VOTE_LOST) return RO_VOTE_LOST;
    return RO_VOTE_PENDING;
}

/* ------------------------------------------------------------------------ */
/* src/tracker/inflights.rs                                                  */

int ro_inflights_full(const ro_inflights *in) { return in->count == in->cap; } /* :54-56 */

/* :65-82 */
int ro_inflights_add(ro_inflights *in, uint64_t inflight) {
    if (ro_inflights_full(in)) return -1; /* panic!("cannot add into a full inflights") */
    uint32_t next = in->start + in->count;
    if (next >= in->cap) next -= in->cap;
    in->buffer[next] = inflight; /* (the Vec grows by push until it holds cap entries: the same ring) */
    in->count += 1;
    return 0;
}

/* :85-110 */
void ro_inflights_free_to(ro_inflights *in, uint64_t to) {
    if (in->count == 0 || to < in->buffer[in->start]) return; /* out of the left side of the window */
    uint32_t i = 0, idx = in->start;
    while (i < in->count) {
        if (to < in->buffer[idx]) break; /* found the first large inflight */
        idx += 1;
        if (idx >= in->cap) idx -= in->cap;
        i += 1;
    }
    in->count -= i;
    in->start = idx;
}

/* :113-116 */
void ro_inflights_free_first_one(ro_inflights *in) { ro_inflights_free_to(in, in->buffer[in->start]); }

/* :119-123 */
void ro_inflights_reset(ro_inflights *in) {
    in->count = 0;
    in->start = 0;
}

/* ------------------------------------------------------------------------ */
/* src/tracker/progress.rs                                                   */

/* :60-73 */
void ro_progress_new(ro_progress *p, uint64_t next_idx) {
    memset(p, 0, sizeof(*p));
    p->matched = 0;
    p->next_idx = next_idx;
    p->state = RO_STATE_PROBE; /* state.rs:31-35 default */
    p->paused = 0;
    p->pending_snapshot = 0;
    p->pending_request_snapshot = 0;
    p->recent_active = 0;
    p->ins_full = 0;
    p->commit_group_id = 0;
    p->committed_index = 0;
}

/* :75-80 */
static void reset_state(ro_progress *p, uint8_t state) {
    p->paused = 0;
    p->pending_snapshot = 0;
    p->state = state;
    p->ins_full = 0; /* ins.reset() */
    if (p->ins) ro_inflights_reset(p->ins);
}

/* :82-92 */
void ro_progress_reset(ro_progress *p, uint64_t next_idx) {
    p->matched = 0;
    p->next_idx = next_idx;
    p->state = RO_STATE_PROBE;
    p->paused = 0;
    p->pending_snapshot = 0;
    p->pending_request_snapshot = RO_INVALID_INDEX;
    p->recent_active = 0;
    p->ins_full = 0; /* ins.reset() */
    if (p->ins) ro_inflights_reset(p->ins);
}

/* :95-107 */
void ro_progress_become_probe(ro_progress *p) {
    if (p->state == RO_STATE_SNAPSHOT) {
        uint64_t pending_snapshot = p->pending_snapshot;
        reset_state(p, RO_STATE_PROBE);
        uint64_t a = p->matched + 1, b = pending_snapshot + 1;
        p->next_idx = a > b ? a : b;
    } else {
        reset_state(p, RO_STATE_PROBE);
        p->next_idx = p->matched + 1;
    }
}

/* :110-114 */
void ro_progress_become_replicate(ro_progress *p) {
    reset_state(p, RO_STATE_REPLICATE);
    p->next_idx = p->matched + 1;
}

/* :117-121 */
void ro_progress_become_snapshot(ro_progress *p, uint64_t snapshot_idx) {
    reset_state(p, RO_STATE_SNAPSHOT);
    p->pending_snapshot = snapshot_idx;
}

/* :124-127 */
void ro_progress_snapshot_failure(ro_progress *p) { p->pending_snapshot = 0; }

/* :131-134 */
int ro_progress_maybe_snapshot_abort(const ro_progress *p) {
    return p->state == RO_STATE_SNAPSHOT && p->matched >= p->pending_snapshot;
}

/* :138-150 (n + 1 wraps like Rust release builds) */
int ro_progress_maybe_update(ro_progress *p, uint64_t n) {
    int need_update = p->matched < n;
    if (need_update) {
        p->matched = n;
        ro_progress_resume(p);
    }
    if (p->next_idx < n + 1) p->next_idx = n + 1;
    return need_update;
}

/* :153-157 */
void ro_progress_update_committed(ro_progress *p, uint64_t committed_index) {
    if (committed_index > p->committed_index) p->committed_index = committed_index;
}

/* :160-163 */
void ro_progress_optimistic_update(ro_progress *p, uint64_t n) { p->next_idx = n + 1; }

/* :168-206 */
int ro_progress_maybe_decr_to(ro_progress *p, uint64_t rejected, uint64_t match_hint,
                              uint64_t request_snapshot) {
    if (p->state == RO_STATE_REPLICATE) {
        /* :173-177 stale rejection */
        if (rejected < p->matched ||
            (rejected == p->matched && request_snapshot == RO_INVALID_INDEX)) {
            return 0;
        }
        if (request_snapshot == RO_INVALID_INDEX) { /* :178-182 */
            p->next_idx = p->matched + 1;
        } else {
            p->pending_request_snapshot = request_snapshot;
        }
        return 1;
    }
    /* :188-192 stale unless it answers next_idx - 1 (or requests a snapshot) */
    if ((p->next_idx == 0 || p->next_idx - 1 != rejected) &&
        request_snapshot == RO_INVALID_INDEX) {
        return 0;
    }
    if (request_snapshot == RO_INVALID_INDEX) { /* :195-199 */
        uint64_t h = match_hint + 1;
        p->next_idx = rejected < h ? rejected : h;
        if (p->next_idx < 1) p->next_idx = 1;
    } else if (p->pending_request_snapshot == RO_INVALID_INDEX) { /* :200-203 */
        p->pending_request_snapshot = request_snapshot;
    }
    ro_progress_resume(p); /* :204 */
    return 1;
}

/* :210-216 */
int ro_progress_is_paused(const ro_progress *p) {
    switch (p->state) {
    case RO_STATE_PROBE:
        return p->paused;
    case RO_STATE_REPLICATE:
        return p->ins ? ro_inflights_full(p->ins) : p->ins_full; /* self.ins.full() */
    default:
        return 1; /* Snapshot */
    }
}

/* :219-222, :225-228 */
void ro_progress_resume(ro_progress *p) { p->paused = 0; }
void ro_progress_pause(ro_progress *p) { p->paused = 1; }

/* :231-243 */
int ro_progress_update_state(ro_progress *p, uint64_t last) {
    switch (p->state) {
    case RO_STATE_REPLICATE:
        if (p->ins && ro_inflights_full(p->ins)) return -1; /* ins.add on a full window panics (inflights.rs:66-68) */
        ro_progress_optimistic_update(p, last);
        if (p->ins) ro_inflights_add(p->ins, last); /* else: the ring is the caller's */
        return 0;
    case RO_STATE_PROBE:
        ro_progress_pause(p);
        return 0;
    default:
        return -1; /* panic!("updating progress state in unhandled state") */
    }
}

/* ------------------------------------------------------------------------ */
/* src/raft_log.rs                                                           */

uint64_t ro_log_last_index(const ro_raft_log *l) { return l->first_index + l->n - 1; }

/* :122-140 */
uint64_t ro_log_term(const ro_raft_log *l, uint64_t idx) {
    uint64_t dummy_idx = l->first_index - 1;
    if (idx < dummy_idx || idx > ro_log_last_index(l)) return 0;
    if (idx == dummy_idx) return l->dummy_term;
    return l->terms[idx - l->first_index];
}

/* :286-300 */
int ro_log_commit_to(ro_raft_log *l, uint64_t to_commit) {
    if (l->committed >= to_commit) return 0; /* never decrease commit */
    if (ro_log_last_index(l) < to_commit) return -1; /* fatal!: out of range */
    l->committed = to_commit;
    return 0;
}

/* :487-499 */
int ro_log_maybe_commit(ro_raft_log *l, uint64_t max_index, uint64_t term) {
    if (max_index > l->committed && ro_log_term(l, max_index) == term) {
        ro_log_commit_to(l, max_index);
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* arena view                                                                */

static inline size_t cell(const ro_arena_view *a, uint32_t slot, uint32_t g) {
    return (size_t)slot * a->cap + g;
}

/* The Progress of one cell.  `ring` backs p->ins when the arena models the Inflights window. */
static void load_progress_ins(const ro_arena_view *a, uint32_t slot, uint32_t g, ro_progress *p, ro_inflights *ring) {
    size_t c = cell(a, slot, g);
    p->ins = NULL;
    if (a->ins_cap && ring) {
        ring->start = a->ins_meta[c] & 0xffffu;
        ring->count = a->ins_meta[c] >> 16;
        ring->cap = a->ins_cap;
        ring->buffer = a->ins_buf + c * (size_t)a->ins_cap;
        p->ins = ring;
    }
    uint8_t f = a->pflags[c];
    p->matched = a->matched[c];
    p->next_idx = a->next_idx[c];
    p->pending_snapshot = a->pending_snapshot[c];
    p->pending_request_snapshot = a->pending_request_snapshot[c];
    p->commit_group_id = a->commit_group_id[c];
    p->committed_index = a->peer_committed[c];
    p->state = f & RO_PF_STATE_MASK;
    p->paused = (f & RO_PF_PAUSED) != 0;
    p->recent_active = (f & RO_PF_RECENT_ACTIVE) != 0;
    p->ins_full = (f & RO_PF_INS_FULL) != 0;
}

static void store_progress(ro_arena_view *a, uint32_t slot, uint32_t g, ro_progress *p) {
    size_t c = cell(a, slot, g);
    if (p->ins) { /* the flag column mirrors ins.full() */
        a->ins_meta[c] = p->ins->start | (p->ins->count << 16);
        p->ins_full = (uint8_t)ro_inflights_full(p->ins);
    }
    a->matched[c] = p->matched;
    a->next_idx[c] = p->next_idx;
    a->pending_snapshot[c] = p->pending_snapshot;
    a->pending_request_snapshot[c] = p->pending_request_snapshot;
    a->commit_group_id[c] = p->commit_group_id;
    a->peer_committed[c] = p->committed_index;
    a->pflags[c] = (uint8_t)((p->state & RO_PF_STATE_MASK) | (p->paused ? RO_PF_PAUSED : 0) |
                             (p->recent_active ? RO_PF_RECENT_ACTIVE : 0) |
                             (p->ins_full ? RO_PF_INS_FULL : 0));
}

/* tracker.rs:294-298 -> joint.rs:47-51 -> majority.rs:70-124, with the
 * ProgressMap AckedIndexer of tracker.rs:183-190.  Peer ids are the slot
 * numbers + 1 (any injective id assignment gives the same result). */
void ro_arena_mci(const ro_arena_view *a, uint32_t g, uint64_t *out_index, int *out_use_gc) {
    uint32_t meta = a->meta[g];
    uint32_t halves = (meta & RO_META_WIDE_LO) ? 2 : 1; /* a wide group: peers 8..15 are the cells of g + 1 */
    uint64_t in_ids[2 * RO_SLOTS], out_ids[2 * RO_SLOTS], map_ids[2 * RO_SLOTS];
    ro_index map_idx[2 * RO_SLOTS];
    size_t n_in = 0, n_out = 0, n_map = 0;
    for (uint32_t h = 0; h < halves; h++) {
        uint32_t m = a->meta[g + h];
        uint32_t in = RO_META_IN(m), out = RO_META_OUT(m), learn = RO_META_LEARN(m);
        uint32_t present = in | out | learn; /* every voter / learner has a Progress */
        for (uint32_t s = 0; s < RO_SLOTS; s++) {
            uint64_t id = 8 * h + s + 1;
            if (in & (1u << s)) in_ids[n_in++] = id;
            if (out & (1u << s)) out_ids[n_out++] = id;
            if (present & (1u << s)) {
                size_t c = cell(a, s, g + h);
                map_ids[n_map] = id;
                map_idx[n_map].index = a->matched[c];           /* tracker.rs:186 */
                map_idx[n_map].group_id = a->commit_group_id[c]; /* tracker.rs:187 */
                n_map++;
            }
        }
    }
    ro_ack_indexer l = {map_ids, map_idx, n_map};
    ro_joint_committed_index(in_ids, n_in, out_ids, n_out, (meta & RO_META_GROUP_COMMIT) != 0, &l,
                             out_index, out_use_gc);
}

/* the leader's own Progress cell: (slot, group slot) -- in whichever half of a wide group it lives */
static int self_cell(const ro_arena_view *a, uint32_t g, size_t *c) {
    uint32_t halves = (a->meta[g] & RO_META_WIDE_LO) ? 2 : 1;
    for (uint32_t h = 0; h < halves; h++) {
        uint32_t m = a->meta[g + h];
        if (m & RO_META_HAS_SELF) {
            *c = cell(a, RO_META_SELF(m), g + h);
            return 1;
        }
    }
    return 0;
}

/* both halves of a wide group carry the group's log bounds and commit index */
static void wide_merge(ro_arena_view *a, uint32_t g) {
    if (!(a->meta[g] & RO_META_WIDE_LO)) return;
    uint64_t li = a->last_index[g] > a->last_index[g + 1] ? a->last_index[g] : a->last_index[g + 1];
    a->last_index[g] = a->last_index[g + 1] = li;
}

/* raft.rs:893-904 with raft_log.rs:487-499 in range form. */
int ro_arena_maybe_commit(ro_arena_view *a, uint32_t g) {
    uint64_t mci;
    int gc;
    if (a->meta[g] & RO_META_WIDE_HI) return 0; /* not a group of its own */
    ro_arena_mci(a, g, &mci, &gc);
    wide_merge(a, g);
    /* raft_log.rs:488: max_index > committed && term(max_index) == term.
     * On a leader the entries of its own term are exactly
     * [term_start, last_index] and term() is 0 above last_index. */
    if (mci > a->committed[g] && mci >= a->term_start[g] && mci <= a->last_index[g]) {
        a->committed[g] = mci; /* commit_to: mci <= last_index so no panic */
        if (a->meta[g] & RO_META_WIDE_LO) a->committed[g + 1] = mci;
        size_t c;
        if (self_cell(a, g, &c)) { /* raft.rs:896-900 */
            if (mci > a->peer_committed[c]) a->peer_committed[c] = mci;
        }
        return 1;
    }
    return 0;
}

/* SURVEY 8(d) synthetic log: first_index = 1, dummy (0, 0). */
static uint64_t synth_term_of(const ro_arena_view *a, uint32_t g, uint64_t idx) {
    if (idx > a->last_index[g]) return 0; /* raft_log.rs:125-126 */
    /* idx < dummy_idx = 0 is impossible for u64 */
    if (idx == 0) return 0; /* dummy entry */
    if (idx >= a->term_start[g]) return a->term[g];
    return a->term[g] - 1;
}

int ro_arena_maybe_commit_literal(ro_arena_view *a, uint32_t g) {
    uint64_t mci;
    int gc;
    ro_arena_mci(a, g, &mci, &gc);
    if (a->term_start[g] == UINT64_MAX) return 0; /* not a leader: no quorum commit */
    if (mci > a->committed[g] && synth_term_of(a, g, mci) == a->term[g]) {
        if (a->last_index[g] < mci) abort(); /* commit_to would fatal!; unreachable */
        a->committed[g] = mci;
        uint32_t meta = a->meta[g];
        if (meta & RO_META_HAS_SELF) {
            size_t c = cell(a, RO_META_SELF(meta), g);
            if (mci > a->peer_committed[c]) a->peer_committed[c] = mci;
        }
        return 1;
    }
    return 0;
}

/* raft.rs:1663-1751 */
uint8_t ro_arena_handle_append_response(ro_arena_view *a, const ro_append_resp *rec,
                                        const ro_append_resp *ext, int per_message_commit,
                                        int *advanced) {
    uint32_t g = rec->group, slot = rec->peer_slot;
    uint32_t meta = a->meta[g];
    uint32_t present = RO_META_IN(meta) | RO_META_OUT(meta) | RO_META_LEARN(meta);
    if (advanced) *advanced = 0;
    /* :1663-1673 prs.get_mut(m.from) == None */
    if (slot >= RO_SLOTS || !(present & (1u << slot))) return RO_RES_NO_PROGRESS;

    ro_progress pr;
    ro_inflights ring;
    load_progress_ins(a, slot, g, &pr, &ring);
    uint8_t res = 0;

    if (rec->flags & RO_REC_LOCAL) {
        /* raft.rs:974-991 append_entry: the leader's log grew to `commit` */
        if (rec->commit != 0) a->last_index[g] = rec->commit;
        /* raft.rs:1010-1014 on_persist_entries on a leader:
         * pr.maybe_update(index) && self.maybe_commit() */
        if (ro_progress_maybe_update(&pr, rec->index)) {
            res |= RO_RES_OK;
            store_progress(a, slot, g, &pr);
            if (per_message_commit) {
                int adv = ro_arena_maybe_commit(a, (meta & RO_META_WIDE_HI) ? g - 1 : g);
                if (advanced) *advanced = adv;
            }
        } else {
            store_progress(a, slot, g, &pr);
        }
        return res;
    }

    pr.recent_active = 1;                           /* :1674 */
    ro_progress_update_committed(&pr, rec->commit); /* :1677 */

    if (rec->flags & RO_REC_REJECT) { /* :1679-1722 */
        uint64_t next_probe_index = ext ? ext->index : 0;
        uint64_t request_snapshot = ext ? ext->commit : RO_INVALID_INDEX;
        if (ro_progress_maybe_decr_to(&pr, rec->index, next_probe_index, request_snapshot)) {
            res |= RO_RES_OK | RO_RES_SEND;
            if (pr.state == RO_STATE_REPLICATE) ro_progress_become_probe(&pr); /* :1716-1718 */
            /* :1719 self.send_append(m.from): message building is the caller's */
        }
        store_progress(a, slot, g, &pr);
        return res;
    }

    int old_paused = ro_progress_is_paused(&pr); /* :1724 */
    if (!ro_progress_maybe_update(&pr, rec->index)) { /* :1725-1727 */
        store_progress(a, slot, g, &pr);
        return res;
    }
    res |= RO_RES_OK;
    if (old_paused) res |= RO_RES_OLD_PAUSED;
    switch (pr.state) { /* :1729-1743 */
    case RO_STATE_PROBE:
        ro_progress_become_replicate(&pr);
        break;
    case RO_STATE_SNAPSHOT:
        if (ro_progress_maybe_snapshot_abort(&pr)) ro_progress_become_probe(&pr);
        break;
    default:
        if (pr.ins) ro_inflights_free_to(pr.ins, rec->index); /* :1742; else the ring is the caller's */
        break;
    }
    store_progress(a, slot, g, &pr);
    if (per_message_commit) { /* :1745 */
        int adv = ro_arena_maybe_commit(a, (meta & RO_META_WIDE_HI) ? g - 1 : g);
        if (advanced) *advanced = adv;
    }
    return res;
}

void ro_arena_apply(ro_arena_view *a, const ro_append_resp *recs, size_t n, int mode,
                    uint8_t *results) {
    for (size_t i = 0; i < n; i++) {
        const ro_append_resp *r = &recs[i];
        if (r->flags & RO_REC_EXT) {
            if (results) results[i] = 0;
            continue;
        }
        const ro_append_resp *ext = NULL;
        if ((r->flags & RO_REC_REJECT) && i + 1 < n && (recs[i + 1].flags & RO_REC_EXT))
            ext = &recs[i + 1];
        uint8_t res = ro_arena_handle_append_response(a, r, ext, mode == 1, NULL);
        if (results) results[i] = res;
    }
}

/* raft.rs:1777-1819 */
uint8_t ro_arena_handle_heartbeat_response(ro_arena_view *a, const ro_append_resp *rec) {
    uint32_t g = rec->group, slot = rec->peer_slot;
    uint32_t meta = a->meta[g];
    uint32_t present = RO_META_IN(meta) | RO_META_OUT(meta) | RO_META_LEARN(meta);
    if (slot >= RO_SLOTS || !(present & (1u << slot))) return RO_RES_NO_PROGRESS; /* :1779-1789 */
    ro_progress pr;
    ro_inflights ring;
    load_progress_ins(a, slot, g, &pr, &ring);
    ro_progress_update_committed(&pr, rec->commit); /* :1791 */
    pr.recent_active = 1;                           /* :1792 */
    ro_progress_resume(&pr);                        /* :1793 */
    if (pr.ins) {                                   /* :1796-1798 */
        if (pr.state == RO_STATE_REPLICATE && ro_inflights_full(pr.ins)) ro_inflights_free_first_one(pr.ins);
    } else if (pr.state == RO_STATE_REPLICATE && pr.ins_full) {
        pr.ins_full = 0; /* a full window that loses its first entry is no longer full */
    }
    uint8_t res = RO_RES_OK;
    if (pr.matched < a->last_index[g] || pr.pending_request_snapshot != RO_INVALID_INDEX) res |= RO_RES_SEND; /* :1800-1803 */
    store_progress(a, slot, g, &pr);
    return res;
}

void ro_arena_apply_heartbeat(ro_arena_view *a, const ro_append_resp *recs, size_t n, uint8_t *results) {
    for (size_t i = 0; i < n; i++) {
        uint8_t res = (recs[i].flags & RO_REC_HEARTBEAT) ? ro_arena_handle_heartbeat_response(a, &recs[i]) : 0;
        if (results) results[i] = res;
    }
}

/* progress.rs:231-243 over a send list */
void ro_arena_update_state(ro_arena_view *a, const ro_send_entry *e, size_t n, uint8_t *results) {
    for (size_t i = 0; i < n; i++) {
        uint32_t g = e[i].group, slot = e[i].peer_slot;
        uint32_t meta = a->meta[g];
        uint32_t present = RO_META_IN(meta) | RO_META_OUT(meta) | RO_META_LEARN(meta);
        uint8_t res;
        if (slot >= RO_SLOTS || !(present & (1u << slot))) {
            res = RO_RES_NO_PROGRESS;
        } else {
            ro_progress pr;
            ro_inflights ring;
            load_progress_ins(a, slot, g, &pr, &ring);
            res = ro_progress_update_state(&pr, e[i].next_idx) == 0 ? 1 : 0xff;
            store_progress(a, slot, g, &pr);
        }
        if (results) results[i] = res;
    }
}

uint64_t ro_arena_recompute(ro_arena_view *a, uint32_t first, uint32_t n, uint32_t *adv_bitmap,
                            uint64_t *mci_out, uint8_t *gc_out) {
    uint64_t advanced = 0;
    for (uint32_t g = first; g < first + n; g++) {
        if ((mci_out || gc_out) && !(a->meta[g] & RO_META_WIDE_HI)) {
            uint64_t mci;
            int gc;
            ro_arena_mci(a, g, &mci, &gc);
            if (mci_out) mci_out[g] = mci;
            if (gc_out) gc_out[g] = (uint8_t)gc;
        }
        int adv = ro_arena_maybe_commit(a, g);
        if (adv_bitmap) {
            if (adv)
                __atomic_fetch_or(&adv_bitmap[g >> 5], 1u << (g & 31), __ATOMIC_RELAXED);
            else
                __atomic_fetch_and(&adv_bitmap[g >> 5], ~(1u << (g & 31)), __ATOMIC_RELAXED);
        }
        advanced += (uint64_t)adv;
    }
    return advanced;
}

/* tracker.rs:313-340 (tally_votes) with the joint vote_result. */
int ro_arena_vote_result(const ro_arena_view *a, const uint8_t *votes, uint32_t g,
                         uint32_t *granted, uint32_t *rejected) {
    if (a->meta[g] & RO_META_WIDE_HI) g--; /* both halves of a wide group report the group's result */
    uint32_t halves = (a->meta[g] & RO_META_WIDE_LO) ? 2 : 1;
    uint64_t in_ids[2 * RO_SLOTS], out_ids[2 * RO_SLOTS], v_ids[2 * RO_SLOTS];
    uint8_t v_vote[2 * RO_SLOTS];
    size_t n_in = 0, n_out = 0, n_v = 0;
    uint32_t gr = 0, rj = 0;
    for (uint32_t h = 0; h < halves; h++) {
        uint32_t meta = a->meta[g + h];
        uint32_t in = RO_META_IN(meta), out = RO_META_OUT(meta);
        for (uint32_t s = 0; s < RO_SLOTS; s++) {
            uint64_t id = 8 * h + s + 1;
            if (in & (1u << s)) in_ids[n_in++] = id;
            if (out & (1u << s)) out_ids[n_out++] = id;
            uint8_t v = votes[cell(a, s, g + h)];
            if (v) {
                v_ids[n_v] = id;
                v_vote[n_v] = (v == 2);
                n_v++;
                if ((in | out) & (1u << s)) { /* tracker.rs:320-322: only voters count */
                    if (v == 2)
                        gr++;
                    else
                        rj++;
                }
            }
        }
    }
    ro_vote_map vm = {v_ids, v_vote, n_v};
    if (granted) *granted = gr;
    if (rejected) *rejected = rj;
    return ro_joint_vote_result(in_ids, n_in, out_ids, n_out, &vm);
}

/* raft.rs:880-889 bcast_heartbeat_with_ctx + raft.rs:838-840 send_heartbeat */
void ro_arena_heartbeat_commits(const ro_arena_view *a, uint32_t first, uint32_t n, uint64_t *out) {
    for (uint32_t i = 0; i < n; i++) {
        uint32_t g = first + i;
        uint32_t meta = a->meta[g];
        uint32_t peers = RO_META_IN(meta) | RO_META_OUT(meta) | RO_META_LEARN(meta);
        for (uint32_t s = 0; s < RO_SLOTS; s++) {
            uint64_t v = UINT64_MAX;
            if ((peers & (1u << s)) && !((meta & RO_META_HAS_SELF) && RO_META_SELF(meta) == s)) { /* :887 */
                uint64_t m = a->matched[cell(a, s, g)];
                v = m < a->committed[g] ? m : a->committed[g];                                   /* :839 */
            }
            out[(size_t)s * n + i] = v;
        }
    }
}

/* progress.rs:210-216 */
static int progress_is_paused(uint8_t flags) {
    switch (flags & RO_PF_STATE_MASK) {
    case RO_STATE_PROBE: return (flags & RO_PF_PAUSED) != 0;
    case RO_STATE_REPLICATE: return (flags & RO_PF_INS_FULL) != 0; /* ins.full(), reported by the host */
    default: return 1;                                              /* Snapshot */
    }
}

/* raft.rs:857-865 bcast_append over the advanced groups; raft.rs:780-788 maybe_send_append's gate */
uint64_t ro_arena_send_list(const ro_arena_view *a, uint32_t first, uint32_t n, const uint32_t *adv_bitmap,
                            ro_send_entry *out, uint64_t cap) {
    uint64_t k = 0;
    for (uint32_t g = first; g < first + n; g++) {
        uint32_t meta = a->meta[g];
        uint32_t bit = (meta & RO_META_WIDE_HI) ? g - 1 : g; /* peers 8..15 of a wide group follow the group's bit */
        if (adv_bitmap && !(adv_bitmap[bit >> 5] & (1u << (bit & 31)))) continue;
        uint32_t peers = RO_META_IN(meta) | RO_META_OUT(meta) | RO_META_LEARN(meta);
        for (uint32_t s = 0; s < RO_SLOTS; s++) {
            if (!(peers & (1u << s))) continue;
            if ((meta & RO_META_HAS_SELF) && RO_META_SELF(meta) == s) continue; /* :863 id != self_id */
            size_t c = cell(a, s, g);
            if (progress_is_paused(a->pflags[c])) continue;                       /* :780 */
            if (k < cap) {
                out[k].group = g;
                out[k].peer_slot = (uint8_t)s;
                out[k].flags = a->pending_request_snapshot[c] != 0 ? 1 : 0;           /* :792 */
                out[k].reserved = 0;
                out[k].next_idx = a->next_idx[c];
            }
            k++;
        }
    }
    return k;
}

/* ------------------------------------------------------------------------ */
/* Tuned CPU path for the baseline arm: same results as the literal functions  */
/* above (tests/test_oracle_fast.py), but written the way a CPU implementation */
/* over flat arrays would be: no id lists, no struct copies.                   */

static inline uint64_t fast_quorum_index(const uint64_t *v, uint32_t mask) {
    if (mask == 0) return UINT64_MAX; /* majority.rs:71-75 */
    uint64_t w[RO_SLOTS];
    int n = 0;
    for (uint32_t s = 0; s < RO_SLOTS; s++)
        if (mask & (1u << s)) {
            uint64_t x = v[s];
            int j = n++;
            while (j > 0 && w[j - 1] < x) { /* descending insertion sort, majority.rs:95 */
                w[j] = w[j - 1];
                j--;
            }
            w[j] = x;
        }
    return w[n / 2]; /* matched[majority(n) - 1], majority.rs:97-98 */
}

static inline int fast_maybe_commit(ro_arena_view *a, uint32_t g) {
    uint32_t meta = a->meta[g];
    if (meta & RO_META_GROUP_COMMIT) return ro_arena_maybe_commit(a, g); /* rare: literal path */
    uint32_t in = RO_META_IN(meta), out = RO_META_OUT(meta), voters = in | out;
    uint64_t v[RO_SLOTS];
    for (uint32_t s = 0; s < RO_SLOTS; s++)
        v[s] = (voters & (1u << s)) ? a->matched[(size_t)s * a->cap + g] : 0;
    uint64_t i_idx = fast_quorum_index(v, in), o_idx = fast_quorum_index(v, out);
    uint64_t mci = i_idx < o_idx ? i_idx : o_idx; /* joint.rs:50 */
    if (mci > a->committed[g] && mci >= a->term_start[g] && mci <= a->last_index[g]) {
        a->committed[g] = mci;
        if (meta & RO_META_HAS_SELF) {
            size_t c = (size_t)RO_META_SELF(meta) * a->cap + g;
            if (mci > a->peer_committed[c]) a->peer_committed[c] = mci;
        }
        return 1;
    }
    return 0;
}

static inline void fast_reset_state(ro_arena_view *a, size_t c, uint8_t *f, uint8_t state) {
    *f = (uint8_t)((*f & ~(RO_PF_PAUSED | RO_PF_INS_FULL | RO_PF_STATE_MASK)) | state);
    if (a->pending_snapshot[c] != 0) a->pending_snapshot[c] = 0;
}

/* raft.rs:1663-1743 on the columns in place (batched mode: no per-message commit). */
static inline void fast_apply_one(ro_arena_view *a, const ro_append_resp *r, const ro_append_resp *ext) {
    uint32_t g = r->group, slot = r->peer_slot;
    if (g >= a->cap || slot >= RO_SLOTS) return;
    uint32_t meta = a->meta[g];
    if (!((RO_META_IN(meta) | RO_META_OUT(meta) | RO_META_LEARN(meta)) & (1u << slot))) return;
    size_t c = (size_t)slot * a->cap + g;
    uint64_t matched = a->matched[c], next = a->next_idx[c];
    uint8_t f = a->pflags[c];
    uint8_t state = f & RO_PF_STATE_MASK;
    uint64_t index = r->index;
    if (r->flags & RO_REC_LOCAL) {
        if (r->commit != 0) a->last_index[g] = r->commit;
        if (matched < index) {
            matched = index;
            f &= (uint8_t)~RO_PF_PAUSED;
        }
        if (next < index + 1) next = index + 1;
    } else {
        f |= RO_PF_RECENT_ACTIVE;
        if (r->commit > a->peer_committed[c]) a->peer_committed[c] = r->commit;
        if (r->flags & RO_REC_REJECT) {
            uint64_t hint = ext ? ext->index : 0, rs = ext ? ext->commit : RO_INVALID_INDEX;
            int ok;
            if (state == RO_STATE_REPLICATE) {
                if (index < matched || (index == matched && rs == RO_INVALID_INDEX)) {
                    ok = 0;
                } else {
                    if (rs == RO_INVALID_INDEX)
                        next = matched + 1;
                    else
                        a->pending_request_snapshot[c] = rs;
                    ok = 1;
                }
            } else if ((next == 0 || next - 1 != index) && rs == RO_INVALID_INDEX) {
                ok = 0;
            } else {
                if (rs == RO_INVALID_INDEX) {
                    uint64_t h = hint + 1;
                    next = index < h ? index : h;
                    if (next < 1) next = 1;
                } else if (a->pending_request_snapshot[c] == RO_INVALID_INDEX) {
                    a->pending_request_snapshot[c] = rs;
                }
                f &= (uint8_t)~RO_PF_PAUSED;
                ok = 1;
            }
            if (ok && state == RO_STATE_REPLICATE) {
                fast_reset_state(a, c, &f, RO_STATE_PROBE);
                next = matched + 1;
            }
        } else {
            int need = matched < index;
            if (need) {
                matched = index;
                f &= (uint8_t)~RO_PF_PAUSED;
            }
            if (next < index + 1) next = index + 1;
            if (need) {
                if (state == RO_STATE_PROBE) {
                    fast_reset_state(a, c, &f, RO_STATE_REPLICATE);
                    next = matched + 1;
                } else if (state == RO_STATE_SNAPSHOT) {
                    uint64_t pending = a->pending_snapshot[c];
                    if (matched >= pending) {
                        fast_reset_state(a, c, &f, RO_STATE_PROBE);
                        uint64_t x = matched + 1, y = pending + 1;
                        next = x > y ? x : y;
                    }
                }
            }
        }
    }
    a->matched[c] = matched;
    a->next_idx[c] = next;
    a->pflags[c] = f;
}

/* ------------------------------------------------------------------------ */
/* CPU baseline timing                                                       */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    ro_arena_view *a;
    uint32_t first, n;
    int iters;
    const ro_append_resp *recs;
    size_t rec_lo, rec_hi;
    uint64_t advanced;
} bench_job;

static void *bench_recompute_worker(void *arg) {
    bench_job *j = (bench_job *)arg;
    uint64_t adv = 0;
    for (int it = 0; it < j->iters; it++) adv += ro_arena_recompute(j->a, j->first, j->n, NULL, NULL, NULL);
    j->advanced = adv;
    return NULL;
}

static void *bench_step_worker(void *arg) {
    bench_job *j = (bench_job *)arg;
    if (j->rec_hi > j->rec_lo)
        ro_arena_apply(j->a, j->recs + j->rec_lo, j->rec_hi - j->rec_lo, 0, NULL);
    j->advanced = ro_arena_recompute(j->a, j->first, j->n, NULL, NULL, NULL);
    return NULL;
}

static size_t lower_bound_group(const ro_append_resp *recs, size_t n, uint32_t g) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (recs[mid].group < g)
            lo = mid + 1;
        else
            hi = mid;
    }
    /* never split a REJECT from its EXT record */
    while (lo > 0 && lo < n && (recs[lo].flags & RO_REC_EXT)) lo++;
    return lo;
}

/* A persistent worker pool: creating 128 threads per step would cost more than the step. */
typedef struct {
    int n_threads;
    pthread_t *th;
    bench_job *jobs;
    void *(*fn)(void *);
    pthread_barrier_t start, done;
    int stop;
} ro_pool;

typedef struct {
    ro_pool *pool;
    int idx;
} pool_arg;

static void *pool_worker(void *argp) {
    pool_arg *pa = (pool_arg *)argp;
    ro_pool *p = pa->pool;
    int idx = pa->idx;
    free(pa);
    for (;;) {
        pthread_barrier_wait(&p->start);
        if (p->stop) return NULL;
        p->fn(&p->jobs[idx]);
        pthread_barrier_wait(&p->done);
    }
}

static ro_pool *g_pool = NULL;

static ro_pool *get_pool(int n_threads) {
    if (g_pool && g_pool->n_threads == n_threads) return g_pool;
    if (g_pool) {
        g_pool->stop = 1;
        pthread_barrier_wait(&g_pool->start);
        for (int t = 1; t < g_pool->n_threads; t++) pthread_join(g_pool->th[t], NULL);
        pthread_barrier_destroy(&g_pool->start);
        pthread_barrier_destroy(&g_pool->done);
        free(g_pool->th);
        free(g_pool->jobs);
        free(g_pool);
        g_pool = NULL;
    }
    ro_pool *p = (ro_pool *)calloc(1, sizeof(ro_pool));
    p->n_threads = n_threads;
    p->th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    p->jobs = (bench_job *)calloc((size_t)n_threads, sizeof(bench_job));
    pthread_barrier_init(&p->start, NULL, (unsigned)n_threads);
    pthread_barrier_init(&p->done, NULL, (unsigned)n_threads);
    for (int t = 1; t < n_threads; t++) {
        pool_arg *pa = (pool_arg *)malloc(sizeof(pool_arg));
        pa->pool = p;
        pa->idx = t;
        pthread_create(&p->th[t], NULL, pool_worker, pa);
    }
    g_pool = p;
    return p;
}

static double run_jobs(ro_arena_view *a, int n_threads, int iters, const ro_append_resp *recs,
                       size_t n_recs, void *(*fn)(void *), uint64_t *advanced_total) {
    if (n_threads < 1) n_threads = 1;
    ro_pool *p = get_pool(n_threads);
    bench_job *jobs = p->jobs;
    uint32_t per = (a->n_groups + (uint32_t)n_threads - 1) / (uint32_t)n_threads;
    for (int t = 0; t < n_threads; t++) {
        uint32_t first = (uint32_t)t * per;
        if (first > a->n_groups) first = a->n_groups;
        uint32_t n = a->n_groups - first < per ? a->n_groups - first : per;
        memset(&jobs[t], 0, sizeof(bench_job));
        jobs[t].a = a;
        jobs[t].first = first;
        jobs[t].n = n;
        jobs[t].iters = iters;
        jobs[t].recs = recs;
        if (recs) {
            jobs[t].rec_lo = lower_bound_group(recs, n_recs, first);
            jobs[t].rec_hi = lower_bound_group(recs, n_recs, first + n);
        }
    }
    p->fn = fn;
    double t0 = now_s();
    pthread_barrier_wait(&p->start); /* the caller is worker 0 */
    fn(&jobs[0]);
    pthread_barrier_wait(&p->done);
    double t1 = now_s();
    uint64_t adv = 0;
    for (int t = 0; t < n_threads; t++) adv += jobs[t].advanced;
    if (advanced_total) *advanced_total = adv;
    return t1 - t0;
}

static void *bench_step_fast_worker(void *arg) {
    bench_job *j = (bench_job *)arg;
    const ro_append_resp *recs = j->recs;
    for (size_t i = j->rec_lo; i < j->rec_hi; i++) {
        const ro_append_resp *r = &recs[i];
        if (r->flags & RO_REC_EXT) continue;
        const ro_append_resp *ext = NULL;
        if ((r->flags & RO_REC_REJECT) && i + 1 < j->rec_hi && (recs[i + 1].flags & RO_REC_EXT)) ext = &recs[i + 1];
        fast_apply_one(j->a, r, ext);
    }
    uint64_t adv = 0;
    for (uint32_t g = j->first; g < j->first + j->n; g++) adv += (uint64_t)fast_maybe_commit(j->a, g);
    j->advanced = adv;
    return NULL;
}

double ro_bench_recompute(ro_arena_view *a, int n_threads, int iters, uint64_t *advanced_total) {
    return run_jobs(a, n_threads, iters, NULL, 0, bench_recompute_worker, advanced_total);
}

double ro_bench_step(ro_arena_view *a, const ro_append_resp *recs, size_t n_recs, int n_threads,
                     uint64_t *advanced_total) {
    return run_jobs(a, n_threads, 1, recs, n_recs, bench_step_worker, advanced_total);
}

double ro_bench_step_fast(ro_arena_view *a, const ro_append_resp *recs, size_t n_recs, int n_threads,
                          uint64_t *advanced_total) {
    return run_jobs(a, n_threads, 1, recs, n_recs, bench_step_fast_worker, advanced_total);
}
