"""Inflights on the device (raftgpu_arena_enable_inflights).  The reference's own table tests
(src/tracker/inflights.rs:131-256) through raftgpu_progress_op, then the whole flow -- update_state adds, accepted
append responses free, heartbeat responses free the first entry of a full window, state changes reset, INS_FULL
mirrors ins.full() -- on random traffic against the oracle (whose ring is pinned by the same table tests)."""
import numpy as np
import pytest

from helpers import B, O, assert_columns_equal

pytestmark = pytest.mark.gpu


@pytest.fixture()
def one():
    a = B.Arena(256)
    a.enable_inflights(10)
    g = a.group_alloc()
    a.group_set_conf(g, 0b11, 0, 0, 0, 1)
    yield a, g
    a.close()


def state(a, g):
    st, cnt, buf = a.inflights_get(g, 1)
    return st, cnt, buf.tolist()


def test_inflight_add_on_gpu(one):          # inflights.rs:131-184
    a, g = one
    for i in range(5):
        assert a.progress_op(g, 1, B.POP_INS_ADD, i) == 0
    assert state(a, g) == (0, 5, [0, 1, 2, 3, 4, 0, 0, 0, 0, 0])
    for i in range(5, 10):
        a.progress_op(g, 1, B.POP_INS_ADD, i)
    assert state(a, g) == (0, 10, list(range(10)))
    assert a.progress_op(g, 1, B.POP_INS_FULL) == 1 and a.progress_get(g, 1).ins_full == 1
    assert a.progress_op(g, 1, B.POP_INS_ADD, 99) == -1            # :66-68 panic: refused, nothing changes
    assert state(a, g) == (0, 10, list(range(10)))
    # the reference's second window starts at 5 over a buffer of five zeros: add five zeros, free them
    a.progress_op(g, 1, B.POP_INS_RESET)
    for _ in range(5):
        a.progress_op(g, 1, B.POP_INS_ADD, 0)
    a.progress_op(g, 1, B.POP_INS_FREE_TO, 0)
    assert state(a, g)[:2] == (5, 0)
    for i in range(5):
        a.progress_op(g, 1, B.POP_INS_ADD, i)
    assert state(a, g) == (5, 5, [0, 0, 0, 0, 0, 0, 1, 2, 3, 4])
    for i in range(5, 10):
        a.progress_op(g, 1, B.POP_INS_ADD, i)
    assert state(a, g) == (5, 10, [5, 6, 7, 8, 9, 0, 1, 2, 3, 4])


def test_inflight_free_to_and_free_first_one_on_gpu(one):      # inflights.rs:186-256
    a, g = one
    for i in range(10):
        a.progress_op(g, 1, B.POP_INS_ADD, i)
    a.progress_op(g, 1, B.POP_INS_FREE_TO, 4)
    assert state(a, g) == (5, 5, list(range(10))) and a.progress_get(g, 1).ins_full == 0
    a.progress_op(g, 1, B.POP_INS_FREE_TO, 8)
    assert state(a, g) == (9, 1, list(range(10)))
    for i in range(10, 15):
        a.progress_op(g, 1, B.POP_INS_ADD, i)
    a.progress_op(g, 1, B.POP_INS_FREE_TO, 12)
    assert state(a, g) == (3, 2, [10, 11, 12, 13, 14, 5, 6, 7, 8, 9])
    a.progress_op(g, 1, B.POP_INS_FREE_TO, 14)
    assert state(a, g) == (5, 0, [10, 11, 12, 13, 14, 5, 6, 7, 8, 9])
    a.progress_op(g, 1, B.POP_INS_RESET)
    for i in range(10):
        a.progress_op(g, 1, B.POP_INS_ADD, i)
    a.progress_op(g, 1, B.POP_INS_FREE_FIRST_ONE)
    assert state(a, g) == (1, 9, list(range(10)))
    # an arena without windows says so
    b = B.Arena(128)
    gb = b.group_alloc()
    b.group_set_conf(gb, 0b11, 0, 0, 0, 1)
    assert b.progress_op(gb, 1, B.POP_INS_FULL) == -3
    b.close()


def test_flow_control_on_random_traffic_vs_oracle():
    """Ticks of: a synthetic AppendResponse round (enqueue + step: free_to, resets), the send list of the step ->
    update_state (ins.add; a full window refuses), heartbeat responses for a third of the peers (free_first_one).
    A window of 3 fills up quickly.  Every column, every window's (start, count), and sampled rings."""
    n, W = 20_000, 3
    synth = B.Synth(n, 0x1F5, k_peers=5)
    a = B.Arena(n)
    a.enable_inflights(W)
    assert a.group_alloc_range(n) == 0
    a.load_columns(synth.initial)
    ref = O.enable_inflights(O.copy_columns(synth.initial), W)
    rng = np.random.default_rng(12)
    d_cnt, cap = a.device_alloc(8), 8 * n
    d_out = a.device_alloc(16 * cap)
    n_full_seen = n_refused = 0
    for tick in range(6):
        recs = synth.next_round().copy()
        a.enqueue(recs)
        r = a.step(B.STEP_READ_COMMITTED | B.STEP_READ_RESULTS)
        want_res = O.arena_apply(ref, recs, mode=0)
        want_adv, want_bm, _, _ = O.arena_recompute(ref)
        assert np.array_equal(a.record_results(0), want_res) and r.n_advanced == want_adv
        # bcast_append over the advanced groups, then Progress::update_state for every MsgAppend built
        entries = a.step_send_list(cap)
        entries = entries[np.lexsort((entries["peer_slot"], entries["group"]))]
        assert np.array_equal(entries, O.arena_send_list(ref, want_bm))
        sent = entries.copy()
        sent["next_idx"] = ref.last_index[sent["group"]] + np.uint64(tick)      # `last` of the message
        got = a.update_state(sent)
        want = O.arena_update_state(ref, sent)
        assert np.array_equal(got, want), tick
        n_refused += int(np.count_nonzero(want == 0xFF))
        # heartbeat responses from a third of the followers
        g = rng.integers(0, n, n // 3, dtype=np.uint32)
        s = rng.integers(1, 5, n // 3).astype(np.uint8)
        _, first = np.unique(g.astype(np.uint64) * 8 + s, return_index=True)
        hb = np.zeros(len(first), dtype=B.APPEND_RESP_DTYPE)
        hb["group"], hb["peer_slot"], hb["flags"] = g[first], s[first], B.REC_HEARTBEAT
        hb["commit"] = ref.committed[hb["group"]]
        n_full_seen += int(np.count_nonzero(ref.pflags[:, :n] & O.PF_INS_FULL))
        assert np.array_equal(a.heartbeat_resp(hb), O.arena_apply_heartbeat(ref, hb)), tick
        assert_columns_equal(a.read_columns(n), ref, n, f"inflights tick {tick}")
        for slot in range(B.SLOTS):
            m = a.column_read(B.COL_INS_META, slot, 0, n, np.uint32)
            assert np.array_equal(m, ref.ins_meta[slot, :n]), (tick, slot)
        for gg in rng.integers(0, n, 40):
            for slot in (1, 2, 3, 4):
                st, cnt, buf = a.inflights_get(int(gg), slot)
                assert (st, cnt) == (int(ref.ins_meta[slot, gg]) & 0xFFFF, int(ref.ins_meta[slot, gg]) >> 16)
                live = [(st + k) % W for k in range(cnt)]
                assert buf[live].tolist() == ref.ins_buf[slot, gg][live].tolist()
    assert n_full_seen > 1000 and n_refused == 0     # windows fill up; the send list never names a full (= paused) peer
    # ... and a send to a full window anyway is what the reference panics on: refused, nothing changes
    full = np.argwhere((ref.pflags[1:5, :n] & O.PF_INS_FULL) != 0)
    assert len(full)
    e = np.zeros(1, dtype=B.SEND_ENTRY_DTYPE)
    e["peer_slot"], e["group"], e["next_idx"] = int(full[0][0]) + 1, int(full[0][1]), 1 << 41
    assert a.update_state(e).tolist() == [0xFF] == O.arena_update_state(ref, e).tolist()
    assert_columns_equal(a.read_columns(n), ref, n, "refused send")
    a.close()
