"""Pins the C oracle against the reference's table-driven unit/integration tests.

Each table below is the data of a reference test (file:line given); the
harness around it restates what that test drives.  Together with
test_oracle_golden.py this is what "parity pinned" means for oracle/.
"""
import ctypes as C

import random

import numpy as np
import pytest

from oracle import oracle as O

L = O.lib()
PROBE, REPL, SNAP = O.STATE_PROBE, O.STATE_REPLICATE, O.STATE_SNAPSHOT


def new_progress(state, matched, next_idx, pending_snapshot=0):
    # src/tracker/progress.rs:250-262 (test helper new_progress)
    p = O.Progress()
    L.ro_progress_new(C.byref(p), next_idx)
    p.state, p.matched, p.pending_snapshot = state, matched, pending_snapshot
    return p


# ---- src/tracker/progress.rs:264-283 test_progress_is_paused
@pytest.mark.parametrize("state,paused,want", [
    (PROBE, False, False), (PROBE, True, True), (REPL, False, False),
    (REPL, True, False), (SNAP, False, True), (SNAP, True, True)])
def test_progress_is_paused(state, paused, want):
    p = new_progress(state, 0, 0)
    p.paused = paused
    assert bool(L.ro_progress_is_paused(C.byref(p))) == want


def test_progress_is_paused_replicate_follows_inflights_full():
    # progress.rs:213: Replicate => self.ins.full()
    p = new_progress(REPL, 0, 0)
    p.ins_full = 1
    assert L.ro_progress_is_paused(C.byref(p)) == 1


# ---- progress.rs:285-295 test_progress_resume
def test_progress_resume():
    p = O.Progress()
    L.ro_progress_new(C.byref(p), 2)
    p.paused = 1
    L.ro_progress_maybe_decr_to(C.byref(p), 1, 1, 0)
    assert not p.paused
    p.paused = 1
    L.ro_progress_maybe_update(C.byref(p), 2)
    assert not p.paused


# ---- progress.rs:297-330 test_progress_become_probe
@pytest.mark.parametrize("state,next_idx,pending,wnext", [
    (REPL, 5, 0, 2), (SNAP, 5, 10, 11), (SNAP, 5, 0, 2)])
def test_progress_become_probe(state, next_idx, pending, wnext):
    p = new_progress(state, 1, next_idx, pending)
    L.ro_progress_become_probe(C.byref(p))
    assert (p.state, p.matched, p.next_idx) == (PROBE, 1, wnext)


# ---- progress.rs:332-349
def test_progress_become_replicate_and_snapshot():
    p = new_progress(PROBE, 1, 5)
    L.ro_progress_become_replicate(C.byref(p))
    assert (p.state, p.matched, p.next_idx) == (REPL, 1, 2)
    p = new_progress(PROBE, 1, 5)
    L.ro_progress_become_snapshot(C.byref(p), 10)
    assert (p.state, p.matched, p.pending_snapshot) == (SNAP, 1, 10)


# ---- progress.rs:351-373 test_progress_update  (prev_m=3, prev_n=5)
@pytest.mark.parametrize("update,wm,wn,wok", [(2, 3, 5, False), (3, 3, 5, False),
                                               (4, 4, 5, True), (5, 5, 6, True)])
def test_progress_update(update, wm, wn, wok):
    p = O.Progress()
    L.ro_progress_new(C.byref(p), 5)
    p.matched = 3
    assert bool(L.ro_progress_maybe_update(C.byref(p), update)) == wok
    assert (p.matched, p.next_idx) == (wm, wn)


# ---- progress.rs:375-412 test_progress_maybe_decr
@pytest.mark.parametrize("state,m,n,rejected,last,w,wn", [
    (REPL, 5, 10, 5, 5, False, 10), (REPL, 5, 10, 4, 4, False, 10), (REPL, 5, 10, 9, 9, True, 6),
    (PROBE, 0, 0, 0, 0, False, 0), (PROBE, 0, 10, 5, 5, False, 10), (PROBE, 0, 10, 9, 9, True, 9),
    (PROBE, 0, 2, 1, 1, True, 1), (PROBE, 0, 1, 0, 0, True, 1), (PROBE, 0, 10, 9, 2, True, 3),
    (PROBE, 0, 10, 9, 0, True, 1)])
def test_progress_maybe_decr(state, m, n, rejected, last, w, wn):
    p = new_progress(state, m, n)
    assert bool(L.ro_progress_maybe_decr_to(C.byref(p), rejected, last, 0)) == w
    assert (p.matched, p.next_idx) == (m, wn)


def test_progress_maybe_decr_request_snapshot():
    # progress.rs:173-183, 188-203: request_snapshot != INVALID_INDEX branches
    p = new_progress(REPL, 5, 10)
    assert L.ro_progress_maybe_decr_to(C.byref(p), 5, 5, 7) == 1   # rejected == matched but a request
    assert (p.next_idx, p.pending_request_snapshot) == (10, 7)
    p = new_progress(PROBE, 0, 10)
    assert L.ro_progress_maybe_decr_to(C.byref(p), 3, 3, 8) == 1   # not next-1, but a request
    assert (p.next_idx, p.pending_request_snapshot) == (10, 8)
    assert L.ro_progress_maybe_decr_to(C.byref(p), 3, 3, 9) == 1   # already pending: keep the first
    assert p.pending_request_snapshot == 8


def test_progress_update_state():
    # progress.rs:231-243
    p = new_progress(REPL, 1, 2)
    assert L.ro_progress_update_state(C.byref(p), 9) == 0 and p.next_idx == 10
    p = new_progress(PROBE, 1, 2)
    assert L.ro_progress_update_state(C.byref(p), 9) == 0 and p.paused and p.next_idx == 2
    p = new_progress(SNAP, 1, 2)
    assert L.ro_progress_update_state(C.byref(p), 9) == -1          # panic! in the reference


# ---- src/raft_log.rs:1497-1522 test_commit_to  (entries (t1,i1),(t2,i2),(t3,i3), committed=2)
@pytest.mark.parametrize("commit,wcommit,wpanic", [(3, 3, False), (1, 2, False), (4, 0, True)])
def test_commit_to(commit, wcommit, wpanic):
    terms = (C.c_uint64 * 3)(1, 2, 3)
    log = O.RaftLog(1, 0, terms, 3, 2)
    rc = L.ro_log_commit_to(C.byref(log), commit)
    assert (rc == -1) == wpanic
    if not wpanic:
        assert log.committed == wcommit


def test_log_term_bounds():
    # raft_log.rs:122-127: 0 outside [first_index-1, last_index]
    terms = (C.c_uint64 * 3)(4, 4, 5)
    log = O.RaftLog(10, 3, terms, 3, 0)
    assert [L.ro_log_term(C.byref(log), i) for i in (8, 9, 10, 11, 12, 13)] == [0, 3, 4, 4, 5, 0]


def one_group(matches, gids=None, incoming=None, outgoing=0, group_commit=False, committed=0,
              term_start=1, last_index=None, term=1):
    n = len(matches)
    c = O.new_columns(4, 1)
    for s, m in enumerate(matches):
        c.matched[s, 0] = m
        c.next_idx[s, 0] = m + 1
        if gids:
            c.commit_group_id[s, 0] = gids[s]
    c.meta[0] = O.make_meta((1 << n) - 1 if incoming is None else incoming, outgoing, 0, 0,
                            group_commit)
    c.committed[0] = committed
    c.term_start[0] = term_start
    c.last_index[0] = max(matches) if last_index is None else last_index
    c.term[0] = term
    return c


# ---- harness/tests/integration_cases/test_raft.rs:1145-1240 test_commit
# (matches, log [(term, index)], sm_term, want committed)
TEST_COMMIT = [
    ([1], [(1, 1)], 1, 1), ([1], [(1, 1)], 2, 0), ([2], [(1, 1), (2, 2)], 2, 2),
    ([1], [(2, 1)], 2, 1),
    ([2, 1, 1], [(1, 1), (2, 2)], 1, 1), ([2, 1, 1], [(1, 1), (1, 2)], 2, 0),
    ([2, 1, 2], [(1, 1), (2, 2)], 2, 2), ([2, 1, 2], [(1, 1), (1, 2)], 2, 0),
    ([2, 1, 1, 1], [(1, 1), (2, 2)], 1, 1), ([2, 1, 1, 1], [(1, 1), (1, 2)], 2, 0),
    ([2, 1, 1, 2], [(1, 1), (2, 2)], 1, 1), ([2, 1, 1, 2], [(1, 1), (1, 2)], 2, 0),
    ([2, 1, 2, 2], [(1, 1), (2, 2)], 2, 2), ([2, 1, 2, 2], [(1, 1), (1, 2)], 2, 0),
]


@pytest.mark.parametrize("matches,logs,sm_term,want", TEST_COMMIT)
def test_commit_literal_log_and_range_form(matches, logs, sm_term, want):
    ids = list(range(1, len(matches) + 1))
    mci, _ = O.majority_committed_index(ids, {i: (m, 0) for i, m in zip(ids, matches)})
    # (1) literal: RaftLog::maybe_commit over the explicit term array (raft_log.rs:487-499)
    terms = (C.c_uint64 * len(logs))(*[t for t, _ in logs])
    log = O.RaftLog(1, 0, terms, len(logs), 0)
    L.ro_log_maybe_commit(C.byref(log), mci, sm_term)
    assert log.committed == want
    # (2) the arena's range form: entries of term sm_term are [term_start, last_index]
    own = [i for t, i in logs if t == sm_term]
    term_start = min(own) if own else O.U64_MAX
    c = one_group(matches, term_start=term_start, last_index=logs[-1][1], term=sm_term)
    O.arena_maybe_commit(c, 0)
    assert int(c.committed[0]) == want


# ---- test_raft.rs:5092-5163 test_group_commit: (matches, group_ids, group-commit want, quorum want)
TEST_GROUP_COMMIT = [
    ([1], [0], 1, 1), ([1], [1], 1, 1),
    ([2, 2, 1], [1, 2, 1], 2, 2), ([2, 2, 1], [1, 1, 2], 1, 2), ([2, 2, 1], [1, 0, 1], 1, 2),
    ([2, 2, 1], [0, 0, 0], 1, 2),
    ([4, 2, 1, 3], [0, 0, 0, 0], 1, 2), ([4, 2, 1, 3], [1, 0, 0, 0], 1, 2),
    ([4, 2, 1, 3], [0, 1, 0, 2], 2, 2), ([4, 2, 1, 3], [0, 2, 1, 0], 1, 2),
    ([4, 2, 1, 3], [1, 1, 1, 1], 2, 2), ([4, 2, 1, 3], [1, 1, 2, 1], 1, 2),
    ([4, 2, 1, 3], [1, 2, 1, 1], 2, 2), ([4, 2, 1, 3], [4, 3, 2, 1], 2, 2),
]


@pytest.mark.parametrize("matches,gids,g_w,q_w", TEST_GROUP_COMMIT)
def test_group_commit(matches, gids, g_w, q_w):
    # log = entries min..max all of term 1, leader term 1 (test_raft.rs:5118-5124)
    c = one_group(matches, gids, group_commit=True, term_start=min(matches),
                  last_index=max(matches))
    O.arena_maybe_commit(c, 0)                      # assign_commit_groups on the leader
    assert int(c.committed[0]) == g_w
    c.meta[0] &= ~np.uint32(O.META_GROUP_COMMIT)     # enable_group_commit(false) -> maybe_commit
    O.arena_maybe_commit(c, 0)
    assert int(c.committed[0]) == q_w


# ---- test_raft.rs:5166-5287 test_group_commit_consistent (leader rows whose answer comes from
# maximal_committed_index: raft.rs:566-575 `Some(use_group_commit && index == committed)`)
@pytest.mark.parametrize("matches,gids,committed,want", [
    ([8], [0], 8, False), ([8, 2, 6], [1, 1, 2], 6, True), ([8, 6, 6], [0, 0, 0], 6, False),
    ([8, 6, 6], [1, 1, 1], 6, False), ([8, 6, 6], [1, 1, 0], 6, False)])
def test_group_commit_consistent(matches, gids, committed, want):
    c = one_group(matches, gids, group_commit=True, committed=committed)
    idx, use_gc = O.arena_mci(c, 0)
    assert (use_gc and idx == committed) == want
    if want:  # test_raft.rs:5270-5277: with group commit disabled it is Some(false)
        c.meta[0] &= ~np.uint32(O.META_GROUP_COMMIT)
        idx, use_gc = O.arena_mci(c, 0)
        assert not (use_gc and idx == committed)


def rec(group, slot, index, commit=0, reject=False, hint=0, request_snapshot=0):
    if reject:
        r = np.zeros(2, dtype=O.APPEND_RESP_DTYPE)
        r[0] = (group, slot, O.REC_REJECT, 0, index, commit)
        r[1] = (group, slot, O.REC_EXT, 0, hint, request_snapshot)
        return r
    r = np.zeros(1, dtype=O.APPEND_RESP_DTYPE)
    r[0] = (group, slot, 0, 0, index, commit)
    return r


# ---- test_raft.rs:2611-2675 test_leader_append_response.  Leader (slot 0) at term 1 with log
# terms [0, 1, 1]; its own matched = persisted = 2; followers Probe, paused (after bcast_append),
# matched 0, next 3.  wnext below is the value BEFORE the send path's optimistic_update (the
# reference's wnext=4 in row 3 includes bcast_append -> update_state, which is the caller's).
@pytest.mark.parametrize("index,reject,wmatch,wnext,wres,wcommitted", [
    (3, True, 0, 3, 0, 0),
    (2, True, 0, 2, O.RES_OK | O.RES_SEND, 0),
    (2, False, 2, 3, O.RES_OK | O.RES_OLD_PAUSED, 2),
    (0, False, 0, 3, 0, 0)])
def test_leader_append_response(index, reject, wmatch, wnext, wres, wcommitted):
    c = O.new_columns(4, 1)
    c.meta[0] = O.make_meta(0b111, 0, 0, 0)
    c.matched[0, 0], c.next_idx[0, 0], c.pflags[0, 0] = 2, 4, REPL
    for s in (1, 2):
        c.matched[s, 0], c.next_idx[s, 0], c.pflags[s, 0] = 0, 3, PROBE | O.PF_PAUSED
    c.committed[0], c.term_start[0], c.last_index[0], c.term[0] = 0, 2, 3, 1
    res = O.arena_apply(c, rec(0, 1, index, reject=reject, hint=index), mode=1)
    assert int(c.matched[1, 0]) == wmatch and int(c.next_idx[1, 0]) == wnext
    assert int(res[0]) == wres
    assert int(c.committed[0]) == wcommitted
    assert c.pflags[1, 0] & O.PF_RECENT_ACTIVE           # raft.rs:1674
    if wcommitted:
        assert int(c.peer_committed[0, 0]) == wcommitted  # raft.rs:896-900
        assert c.pflags[1, 0] & O.PF_STATE_MASK == REPL   # raft.rs:1730 Probe -> Replicate


# ---- harness/tests/integration_cases/test_raft_paper.rs:499-534 test_leader_acknowledge_commit
@pytest.mark.parametrize("size,acceptors,wack", [
    (1, [], True), (3, [], False), (3, [2], True), (3, [2, 3], True), (5, [], False),
    (5, [2], False), (5, [2, 3], True), (5, [2, 3, 4], True), (5, [2, 3, 4, 5], True)])
def test_leader_acknowledge_commit(size, acceptors, wack):
    # after commit_noop_entry everyone matched li=1 (committed 1); leader proposes + persists 2
    c = one_group([1] * size, committed=1, term_start=1, last_index=2)
    c.matched[0, 0] = 2
    c.pflags[:size, 0] = REPL
    for a in acceptors:
        O.arena_apply(c, rec(0, a - 1, 2), mode=1)
    O.arena_maybe_commit(c, 0)  # size 1: on_persist_entries -> maybe_commit (raft.rs:1010-1014)
    assert (int(c.committed[0]) > 1) == wack


# ---- test_raft_paper.rs:1012-1052 test_leader_only_commits_log_from_current_term
@pytest.mark.parametrize("index,wcommit", [(1, 0), (2, 0), (3, 3)])
def test_leader_only_commits_log_from_current_term(index, wcommit):
    # log terms [1, 2] + leader (term 3) noop... the proposal lands at index 3 of term 3
    # (become_leader's noop + propose are indexes 3 and 4 in the reference; the test's
    # w=3 row acks index 3, the first entry of the leader's term).
    c = one_group([4, 0], committed=0, term_start=3, last_index=4, term=3)
    c.pflags[:2, 0] = REPL
    O.arena_apply(c, rec(0, 1, index), mode=1)
    assert int(c.committed[0]) == wcommit
    # literal synthetic log agrees with the range form
    c2 = one_group([4, 0], committed=0, term_start=3, last_index=4, term=3)
    c2.matched[1, 0] = index
    O.arena_maybe_commit(c2, 0, literal=True)
    assert int(c2.committed[0]) == wcommit


def test_update_committed_from_append_response():
    # raft.rs:1677 + progress.rs:153-157 (pinned end-to-end by test_raft.rs:116-299)
    c = one_group([5, 3, 3], committed=3, last_index=5)
    O.arena_apply(c, rec(0, 1, 3, commit=3), mode=0)
    assert int(c.peer_committed[1, 0]) == 3
    O.arena_apply(c, rec(0, 1, 3, commit=2), mode=0)
    assert int(c.peer_committed[1, 0]) == 3   # never decreases


def test_unknown_responder_is_skipped():
    # raft.rs:1663-1673
    c = one_group([5, 3, 3], last_index=5)
    res = O.arena_apply(c, rec(0, 6, 4), mode=0)
    assert int(res[0]) == O.RES_NO_PROGRESS and int(c.matched[6, 0]) == 0


def test_snapshot_state_accept_paths():
    # raft.rs:1731-1741 + progress.rs:95-107, 131-134
    c = one_group([10, 2, 2], last_index=10)
    c.pflags[1, 0] = SNAP
    c.pending_snapshot[1, 0] = 8
    O.arena_apply(c, rec(0, 1, 5), mode=0)            # matched 5 < pending 8: stays Snapshot
    assert c.pflags[1, 0] & 3 == SNAP and int(c.matched[1, 0]) == 5
    O.arena_apply(c, rec(0, 1, 9), mode=0)            # matched 9 >= 8: abort -> Probe
    assert c.pflags[1, 0] & 3 == PROBE
    assert int(c.next_idx[1, 0]) == 10 and int(c.pending_snapshot[1, 0]) == 0


def test_vote_result_arena_matches_quorum_functions():
    import random
    rng = random.Random(5)
    for _ in range(500):
        inc, out = rng.randrange(0, 256), rng.choice([0, rng.randrange(0, 256)])
        c = O.new_columns(1, 1)
        c.meta[0] = O.make_meta(inc, out, 0, None)
        votes = np.zeros((O.SLOTS, 1), dtype=np.uint8)
        vm = {}
        for s in range(8):
            v = rng.randrange(0, 3)
            votes[s, 0] = v
            if v:
                vm[s + 1] = (v == 2)
        a = [s + 1 for s in range(8) if inc >> s & 1]
        b = [s + 1 for s in range(8) if out >> s & 1]
        gr, rj, r = O.arena_vote_result(c, votes, 0)
        assert r == O.joint_vote_result(a, b, vm)
        assert gr == sum(1 for s in range(8) if (inc | out) >> s & 1 and votes[s, 0] == 2)
        assert rj == sum(1 for s in range(8) if (inc | out) >> s & 1 and votes[s, 0] == 1)


# ---- bcast_append behind maybe_commit (raft.rs:857-865, 1745-1748), every
# send gated by Progress::is_paused (raft.rs:780-788).  The arena restatement must agree, peer by
# peer, with the Progress-level is_paused that progress.rs:264-283 pins above.
def test_send_list_arena_matches_progress_is_paused():
    import random
    rng = random.Random(9)
    n = 300
    c = O.new_columns(n, n)
    expect = []
    bm = np.zeros((n + 31) // 32, dtype=np.uint32)
    for g in range(n):
        inc, out = rng.randrange(0, 256), rng.choice([0, rng.randrange(0, 256)])
        learners = rng.randrange(0, 256) & ~(inc | out)
        present = inc | out | learners
        self_slot = rng.choice([None] + [s for s in range(8) if present >> s & 1]) if present else None
        c.meta[g] = O.make_meta(inc, out, learners, self_slot)
        adv = rng.random() < 0.5
        if adv:
            bm[g >> 5] |= np.uint32(1 << (g & 31))
        for s in range(8):
            state = rng.choice([PROBE, REPL, SNAP])
            paused, full = rng.random() < 0.4, rng.random() < 0.4
            c.pflags[s, g] = state | (O.PF_PAUSED if paused else 0) | (O.PF_INS_FULL if full else 0) | \
                (O.PF_RECENT_ACTIVE if rng.random() < 0.5 else 0)
            c.next_idx[s, g] = rng.randrange(1, 1 << 40)
            c.pending_request_snapshot[s, g] = rng.choice([0, 0, 0, rng.randrange(1, 1 << 30)])
            p = new_progress(state, 0, 0)
            p.paused, p.ins_full = int(paused), int(full)
            if adv and (present >> s & 1) and s != self_slot and not L.ro_progress_is_paused(C.byref(p)):
                expect.append((g, s, 1 if c.pending_request_snapshot[s, g] else 0, int(c.next_idx[s, g])))
    got = O.arena_send_list(c, bm)
    assert list(zip(got["group"].tolist(), got["peer_slot"].tolist(), got["flags"].tolist(),
                    got["next_idx"].tolist())) == expect
    # no bitmap = a plain bcast_append over every group of the range
    every = O.arena_send_list(c, None, 10, 100)
    assert set(every["group"].tolist()) <= set(range(10, 110)) and len(every) > len(got) // 4


# ---- SURVEY 8(f) rank 3 (leader side): send_heartbeat attaches min(pr.matched, raft_log.committed)
# (raft.rs:838-840); test_raft.rs:2713-2721 (test_bcast_beat) expects exactly that per follower ("want_commit =
# min(committed, matched)") and nothing for the leader itself (raft.rs:887).
def test_heartbeat_commits_arena():
    c = O.new_columns(2, 2)
    c.meta[0] = O.make_meta(0b0111, 0, 0b1000, 0)        # voters 0,1,2 + learner 3, self = slot 0
    c.committed[0] = 1005
    c.matched[:, 0] = [1011, 5, 1006, 1000, 9, 9, 9, 9]
    c.meta[1] = O.make_meta(0b0011, 0b0110, 0, None)     # joint, no self slot known
    c.committed[1] = 7
    c.matched[:, 1] = [7, 9, 3, 0, 0, 0, 0, 0]
    out = O.arena_heartbeat_commits(c)
    none = (1 << 64) - 1
    assert out[:, 0].tolist() == [none, 5, 1005, 1000, none, none, none, none]
    assert out[:, 1].tolist() == [7, 7, 3, none, none, none, none, none]


# ---- SURVEY 8(f) rank 3 (response side): handle_heartbeat_response, raft.rs:1777-1804.
def hb(group, slot, commit=0):
    r = np.zeros(1, dtype=O.APPEND_RESP_DTYPE)
    r[0] = (group, slot, 0x04, 0, 0, commit)
    return r


def test_heartbeat_response_resumes_and_frees_one_inflight():
    # test_raft.rs:329-347 test_progress_resume_by_heartbeat_resp: a paused peer in Replicate state is resumed
    c = one_group([11, 5], last_index=11)
    c.pflags[1, 0] = REPL | O.PF_PAUSED
    res = O.arena_apply_heartbeat(c, hb(0, 1))
    assert not (c.pflags[1, 0] & O.PF_PAUSED) and (c.pflags[1, 0] & O.PF_RECENT_ACTIVE)
    assert res[0] == O.RES_OK | O.RES_SEND            # matched 5 < last_index 11: send_append (raft.rs:1800-1803)
    # test_raft_flow_control.rs:100-187 test_msg_app_flow_control_recv_heartbeat: a full window is not full
    # any more after ONE heartbeat response (free_first_one), and stays so for further ones
    c.pflags[1, 0] = REPL | O.PF_INS_FULL
    for _ in range(3):
        O.arena_apply_heartbeat(c, hb(0, 1))
        assert not (c.pflags[1, 0] & O.PF_INS_FULL)
    # a Probe peer's "full" bit is not looked at (raft.rs:1796: only in Replicate); paused clears
    # (test_raft.rs:2855-2880: the heartbeat response lets ONE more append go out to a paused probing peer)
    c.pflags[1, 0] = PROBE | O.PF_PAUSED | O.PF_INS_FULL
    res = O.arena_apply_heartbeat(c, hb(0, 1))
    assert c.pflags[1, 0] & O.PF_INS_FULL and not (c.pflags[1, 0] & O.PF_PAUSED) and res[0] & O.RES_SEND
    # update_committed (raft.rs:1791) never decreases; a caught-up peer with no snapshot request gets no append
    c.matched[1, 0], c.peer_committed[1, 0] = 11, 9
    assert O.arena_apply_heartbeat(c, hb(0, 1, commit=7))[0] == O.RES_OK and c.peer_committed[1, 0] == 9
    assert O.arena_apply_heartbeat(c, hb(0, 1, commit=10))[0] == O.RES_OK and c.peer_committed[1, 0] == 10
    c.pending_request_snapshot[1, 0] = 4
    assert O.arena_apply_heartbeat(c, hb(0, 1))[0] == O.RES_OK | O.RES_SEND
    # unknown responder (raft.rs:1779-1789); a record without the HEARTBEAT flag is not this path's
    assert O.arena_apply_heartbeat(c, hb(0, 5))[0] == O.RES_NO_PROGRESS
    plain = hb(0, 1)
    plain["flags"] = 0
    assert O.arena_apply_heartbeat(c, plain)[0] == 0


def test_update_state_over_a_send_list():
    # progress.rs:231-243 (pinned Progress-level by test_progress_update_state above): the arena form
    c = one_group([20, 4, 7, 9], last_index=20)
    c.pflags[1, 0], c.pflags[2, 0], c.pflags[3, 0] = REPL, PROBE, SNAP
    e = np.zeros(5, dtype=[("group", "<u4"), ("peer_slot", "u1"), ("flags", "u1"), ("reserved", "<u2"), ("next_idx", "<u8")])
    e["peer_slot"] = [1, 2, 3, 6, 1]
    e["next_idx"] = [15, 12, 9, 1, 18]       # `last` of every MsgAppend built
    res = O.arena_update_state(c, e)
    assert res.tolist() == [1, 1, 0xFF, O.RES_NO_PROGRESS, 1]
    assert c.next_idx[1, 0] == 19             # optimistic_update of the second send to peer 1: last + 1
    assert c.pflags[2, 0] & O.PF_PAUSED and c.next_idx[2, 0] == 8     # Probe: paused, next_idx untouched
    assert c.next_idx[3, 0] == 10 and not (c.pflags[3, 0] & O.PF_PAUSED)


# ---- src/tracker/inflights.rs:131-256: the reference's own table tests, restated for ro_inflights
def _ins(cap, start=0, buffer=()):
    buf = (C.c_uint64 * cap)(*buffer)
    return O.Inflights(start, 0, cap, buf), buf


def _state(ins, buf, n):
    return ins.start, ins.count, list(buf)[:n]


def test_inflight_add():
    ins, buf = _ins(10)
    for i in range(5):
        assert L.ro_inflights_add(C.byref(ins), i) == 0
    assert _state(ins, buf, 5) == (0, 5, [0, 1, 2, 3, 4])
    for i in range(5, 10):
        L.ro_inflights_add(C.byref(ins), i)
    assert _state(ins, buf, 10) == (0, 10, list(range(10)))
    assert L.ro_inflights_full(C.byref(ins)) and L.ro_inflights_add(C.byref(ins), 99) == -1    # :66-68 panic
    ins2, buf2 = _ins(10, start=5, buffer=[0] * 5)
    for i in range(5):
        L.ro_inflights_add(C.byref(ins2), i)
    assert _state(ins2, buf2, 10) == (5, 5, [0, 0, 0, 0, 0, 0, 1, 2, 3, 4])
    for i in range(5, 10):
        L.ro_inflights_add(C.byref(ins2), i)
    assert _state(ins2, buf2, 10) == (5, 10, [5, 6, 7, 8, 9, 0, 1, 2, 3, 4])


def test_inflight_free_to():
    ins, buf = _ins(10)
    for i in range(10):
        L.ro_inflights_add(C.byref(ins), i)
    L.ro_inflights_free_to(C.byref(ins), 4)
    assert _state(ins, buf, 10) == (5, 5, list(range(10)))
    L.ro_inflights_free_to(C.byref(ins), 8)
    assert _state(ins, buf, 10) == (9, 1, list(range(10)))
    for i in range(10, 15):
        L.ro_inflights_add(C.byref(ins), i)
    L.ro_inflights_free_to(C.byref(ins), 12)
    assert _state(ins, buf, 10) == (3, 2, [10, 11, 12, 13, 14, 5, 6, 7, 8, 9])
    L.ro_inflights_free_to(C.byref(ins), 14)
    assert _state(ins, buf, 10) == (5, 0, [10, 11, 12, 13, 14, 5, 6, 7, 8, 9])


def test_inflight_free_first_one():
    ins, buf = _ins(10)
    for i in range(10):
        L.ro_inflights_add(C.byref(ins), i)
    L.ro_inflights_free_first_one(C.byref(ins))
    assert _state(ins, buf, 10) == (1, 9, list(range(10)))


def test_arena_inflights_follow_the_messages():
    """The window through the arena functions: update_state adds (progress.rs:231-236), an accepted append response
    frees up to its index (raft.rs:1742), a heartbeat response frees the first entry of a full window (raft.rs:1796-1798),
    a state change resets (progress.rs:75-80); INS_FULL mirrors ins.full() -- test_raft_flow_control.rs:26-98
    (test_msg_app_flow_control_full / _move_forward) in miniature."""
    c = O.enable_inflights(one_group([20, 4], last_index=20), 3)
    c.pflags[1, 0] = REPL
    e = np.zeros(1, dtype=[("group", "<u4"), ("peer_slot", "u1"), ("flags", "u1"), ("reserved", "<u2"), ("next_idx", "<u8")])
    e["peer_slot"] = 1
    for last in (5, 6, 7):
        e["next_idx"] = last
        assert O.arena_update_state(c, e).tolist() == [1]
    assert c.ins_meta[1, 0] == (3 << 16) and c.ins_buf[1, 0].tolist() == [5, 6, 7] and c.pflags[1, 0] & O.PF_INS_FULL
    e["next_idx"] = 8
    assert O.arena_update_state(c, e).tolist() == [0xFF] and c.next_idx[1, 0] == 8      # add on a full window panics: nothing changes
    r = rec(0, 1, 6, commit=0)
    assert O.arena_apply(c, r, mode=0)[0] == O.RES_OK | O.RES_OLD_PAUSED                 # was paused: the window was full
    assert c.ins_meta[1, 0] == (2 | (1 << 16)) and not (c.pflags[1, 0] & O.PF_INS_FULL)  # freed 5 and 6: start 2, count 1
    for last in (8, 9):
        e["next_idx"] = last
        O.arena_update_state(c, e)
    assert c.pflags[1, 0] & O.PF_INS_FULL and c.ins_buf[1, 0].tolist() == [8, 9, 7]
    assert O.arena_apply_heartbeat(c, hb(0, 1))[0] & O.RES_OK
    assert c.ins_meta[1, 0] == (0 | (2 << 16)) and not (c.pflags[1, 0] & O.PF_INS_FULL) # free_first_one: 7 is gone
    rej = rec(0, 1, 9, reject=True, hint=5)
    O.arena_apply(c, rej, mode=0)                                                         # Replicate -> Probe: ins.reset()
    assert (c.pflags[1, 0] & 3) == PROBE and c.ins_meta[1, 0] == 0


def test_wide_groups_in_the_arena_view_equal_the_id_list_functions():
    """A wide group (two consecutive slots, RO_META_WIDE_LO / _HI) in the arena view against the oracle's own
    id-list functions (majority.rs / joint.rs restated for any number of voters): committed index, group commit,
    vote result, and Raft::maybe_commit writing both halves."""
    rng = random.Random(77)
    WIDE_LO, WIDE_HI = 0x20000000, 0x40000000
    for it in range(200):
        c = O.new_columns(4, 4)
        inc, out = rng.randrange(1, 1 << 16), rng.choice([0, rng.randrange(0, 1 << 16)])
        learn = rng.randrange(0, 1 << 16) & ~(inc | out) if rng.random() < 0.3 else 0
        gc = rng.random() < 0.4
        voters = [s for s in range(16) if (inc | out) >> s & 1]
        self_slot = rng.choice(voters)
        for h in (0, 1):
            c.meta[2 + h] = O.make_meta((inc >> 8 * h) & 0xff, (out >> 8 * h) & 0xff, (learn >> 8 * h) & 0xff,
                                        self_slot - 8 * h if self_slot // 8 == h else None, group_commit=gc) | (WIDE_HI if h else WIDE_LO)
        lookup, votes_by_id = {}, {}
        votes = np.zeros((8, 4), dtype=np.uint8)
        for s in range(16):
            m, gid = rng.randrange(0, 60), rng.randrange(0, 3)
            c.matched[s & 7, 2 + (s >> 3)], c.commit_group_id[s & 7, 2 + (s >> 3)] = m, gid
            if (inc | out | learn) >> s & 1:
                lookup[s + 1] = (m, gid)
            v = rng.randrange(0, 3)
            votes[s & 7, 2 + (s >> 3)] = v
            if v:
                votes_by_id[s + 1] = v == 2
        a = [s + 1 for s in range(16) if inc >> s & 1]
        b = [s + 1 for s in range(16) if out >> s & 1]
        full = {i: lookup.get(i, (0, 0)) for i in set(a) | set(b)}
        want = O.joint_committed_index(a, b, full, gc)
        assert tuple(O.arena_mci(c, 2)) == tuple(want), (it, a, b)
        gr, rj, res = O.arena_vote_result(c, votes, 2)
        assert res == O.joint_vote_result(a, b, votes_by_id)
        assert O.arena_vote_result(c, votes, 3) == (gr, rj, res)          # the high half reports the group's result
        assert gr == sum(1 for i, v in votes_by_id.items() if v and i in full)
        # maybe_commit: log bounds on one half only (a LOCAL record wrote there), commit index lands on both
        c.term_start[2] = c.term_start[3] = 1
        c.last_index[2 + (self_slot >> 3)] = 100
        adv, bm, _, _ = O.arena_recompute(c)
        mci = want[0]
        if 0 < mci <= 100:
            assert adv == 1 and bm[0] == 0b100
            assert c.committed[2] == c.committed[3] == mci and c.last_index[2] == c.last_index[3] == 100
            assert c.peer_committed[self_slot & 7, 2 + (self_slot >> 3)] == mci
        else:
            assert adv == 0 and bm[0] == 0
