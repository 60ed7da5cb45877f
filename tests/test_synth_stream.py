"""The synthetic AppendResponse workload (csrc/synth.cpp) and what the
oracle proves on it (BASELINE config 1: 1 000 groups x 5 peers, CPU only):

* the stream is deterministic and every round is ONE wave (<= 1 record per cell);
* acks never exceed the leader's last_index, so ONE recompute per batch equals the
  reference's per-message Raft::maybe_commit (raft.rs:1745) -- the equivalence the
  batched engine relies on;
* the range form of RaftLog::maybe_commit agrees with the literal term lookup.
"""
import numpy as np
import pytest

from helpers import B, O, assert_columns_equal

SEED1 = 0x5EED0001


def test_stream_is_deterministic():
    a, b = B.Synth(1000, SEED1), B.Synth(1000, SEED1)
    assert_columns_equal(a.initial, b.initial, 1000)
    for _ in range(3):
        assert np.array_equal(a.next_round(), b.next_round())
    c = B.Synth(1000, SEED1 + 1)
    assert not np.array_equal(c.initial.matched, a.initial.matched)


@pytest.mark.parametrize("joint", [False, True])
def test_round_is_one_wave_and_respects_last_index(joint):
    s = B.Synth(1000, SEED1, joint=joint)
    last = s.initial.last_index.copy()
    for _ in range(8):
        r = s.next_round()
        main = r[(r["flags"] & B.REC_EXT) == 0]
        cells = main["group"].astype(np.uint64) * 8 + main["peer_slot"]
        assert len(np.unique(cells)) == len(cells)             # one wave
        assert np.all(np.diff(r["group"].astype(np.int64)) >= 0)  # group order
        acc = main[main["flags"] == 0]
        assert np.all(acc["index"] <= last[acc["group"]])       # acks <= leader's last_index
        loc = main[main["flags"] == B.REC_LOCAL]
        assert len(loc) == 1000 and np.all(loc["peer_slot"] == 0)
        assert np.all(loc["commit"] >= last[loc["group"]]) and np.all(loc["index"] <= loc["commit"])
        last[loc["group"]] = loc["commit"]
        rej = r[(r["flags"] & B.REC_REJECT) != 0]
        nxt = r[np.nonzero((r["flags"] & B.REC_REJECT) != 0)[0] + 1]
        assert np.all(nxt["flags"] == B.REC_EXT) and np.array_equal(nxt["group"], rej["group"])
    mix = s.next_round()
    frac_reject = np.mean((mix["flags"] & B.REC_REJECT) != 0)
    assert 0.002 < frac_reject < 0.03


@pytest.mark.parametrize("joint", [False, True])
def test_batched_step_equals_per_message_commit(joint):
    s = B.Synth(1000, SEED1, joint=joint)
    seq, bat = O.copy_columns(s.initial), O.copy_columns(s.initial)
    for rnd in range(16):
        recs = s.next_round().copy()
        r1 = O.arena_apply(seq, recs, mode=1)        # literal: maybe_commit after every message
        r0 = O.arena_apply(bat, recs, mode=0)        # batched: apply the wave ...
        O.arena_recompute(bat)                       # ... then ONE recompute pass
        O.arena_recompute(seq)                       # (no-op unless a LOCAL-only advance is pending)
        assert np.array_equal(r0, r1)
        assert_columns_equal(bat, seq, 1000, f"round {rnd}")
    assert np.any(bat.committed != s.initial.committed)


@pytest.mark.parametrize("joint", [False, True])
def test_range_form_equals_literal_term_lookup(joint):
    s = B.Synth(1000, SEED1, joint=joint)
    a, b = O.copy_columns(s.initial), O.copy_columns(s.initial)
    n_adv = 0
    for _ in range(6):
        recs = s.next_round().copy()
        O.arena_apply(a, recs, mode=0)
        O.arena_apply(b, recs, mode=0)
        for g in range(1000):
            x = O.arena_maybe_commit(a, g, literal=False)
            y = O.arena_maybe_commit(b, g, literal=True)
            assert x == y
            n_adv += x
        assert np.array_equal(a.committed, b.committed)
    assert n_adv > 1000
    # the guard bites: some quorum indexes fall before the leader's term
    mci = np.array([O.arena_mci(s.initial, g)[0] for g in range(1000)], dtype=np.uint64)
    assert np.any(mci < s.initial.term_start[:1000]) and np.any(mci >= s.initial.term_start[:1000])


def test_joint_config_masks():
    s = B.Synth(8, SEED1, joint=True)
    m = int(s.initial.meta[0])
    assert m & 0xFF == 0x1F and (m >> 8) & 0xFF == 0x67   # {0..4} and {0,1,2,5,6}
    s = B.Synth(8, SEED1, k_peers=5)
    assert int(s.initial.meta[0]) & 0xFFFF == 0x1F
