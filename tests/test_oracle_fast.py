"""The tuned CPU path the baseline arm times (ro_bench_step_fast) gives exactly the results
of the literal oracle, on plain and joint streams, single- and multi-threaded."""
import numpy as np
import pytest

from helpers import B, O, assert_columns_equal


@pytest.mark.parametrize("joint", [False, True])
@pytest.mark.parametrize("threads", [1, 3])
def test_fast_cpu_step_equals_literal(joint, threads):
    n = 3000
    s = B.Synth(n, 0x5EED0001, joint=joint)
    lit, fast = O.copy_columns(s.initial), O.copy_columns(s.initial)
    # make some groups use group commit and some peers Snapshot so every branch is hit
    for c in (lit, fast):
        c.meta[::7] |= np.uint32(O.META_GROUP_COMMIT)
        c.commit_group_id[1:4, ::7] = np.array([[1], [2], [1]], dtype=np.uint64)
        c.pflags[2, ::11] = O.STATE_SNAPSHOT
        c.pending_snapshot[2, ::11] = c.matched[2, ::11] + 20
    for rnd in range(10):
        recs = s.next_round().copy()
        O.arena_apply(lit, recs, mode=0)
        adv_lit, _, _, _ = O.arena_recompute(lit)
        _, adv_fast = O.bench_step(fast, recs, threads, fast=True)
        assert adv_fast == adv_lit
        assert_columns_equal(fast, lit, n, f"round {rnd}")
