"""N > 1 path on CPU: two gloo ranks each own half of the groups (BASELINE config 5 in
miniature), route the global record stream by group, run their shard, and all-reduce the
counters.  The union of the shards must equal the unsharded run bit for bit (groups are
independent), and the aggregated counters must equal the unsharded counters.  The compute in
this CPU test is the oracle (the GPU arm runs the same host logic around libraftgpu.so)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

N_TOTAL, ROUNDS, SEED = 10_000, 6, 0x5EED0005


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("raft-rs_b200")
    B, S = pkg.binding, importlib.import_module("raft-rs_b200.shard")
    from oracle import oracle as O
    synth = B.Synth(N_TOTAL, SEED)                      # every rank sees the same global stream
    lo, hi = S.shard_bounds(N_TOTAL, world, rank)
    cols = S.slice_columns(synth.initial, lo, hi, O.new_columns)
    recomputes = advanced = records = 0
    for _ in range(ROUNDS):
        recs = S.route_records(synth.next_round(), N_TOTAL, world, rank)
        res = O.arena_apply(cols, recs, mode=0)
        adv, _, _, _ = O.arena_recompute(cols)
        recomputes += hi - lo
        advanced += adv
        records += int(np.count_nonzero((recs["flags"] & B.REC_EXT) == 0))
    sums, maxes = S.aggregate(dist, torch, {"recomputes": recomputes, "advanced": advanced, "records": records},
                              {"elapsed": 1.0 + rank})
    np.savez(os.path.join(tmpdir, f"shard{rank}.npz"), lo=lo, hi=hi, committed=cols.committed,
             matched=cols.matched, next_idx=cols.next_idx, pflags=cols.pflags,
             sums=np.array([sums["recomputes"], sums["advanced"], sums["records"]]), tmax=maxes["elapsed"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_unsharded(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import B, O
    synth = B.Synth(N_TOTAL, SEED)
    ref = O.copy_columns(synth.initial)
    want = np.zeros(3)
    for _ in range(ROUNDS):
        recs = synth.next_round().copy()
        O.arena_apply(ref, recs, mode=0)
        adv, _, _, _ = O.arena_recompute(ref)
        want += [N_TOTAL, adv, np.count_nonzero((recs["flags"] & B.REC_EXT) == 0)]
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"shard{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        n = hi - lo
        covered += n
        assert np.array_equal(z["committed"][:n], ref.committed[lo:hi])
        for name in ("matched", "next_idx", "pflags"):
            assert np.array_equal(z[name][:, :n], getattr(ref, name)[:, lo:hi]), name
        assert np.array_equal(z["sums"], want)        # every rank holds the all-reduced totals
        assert float(z["tmax"]) == float(world)        # MAX over ranks of (1 + rank)
    assert covered == N_TOTAL


def test_shard_bounds_cover_exactly():
    S = importlib.import_module("raft-rs_b200.shard")
    for n in (1, 7, 10_000_000, 1_000_003):
        for w in (1, 2, 4, 8):
            spans = [S.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            # every block starts on an even group: a wide group (two consecutive slots from an even one) is never
            # split between ranks and keeps its parity after the rebase to rank-local slots
            assert all(lo % 2 == 0 for lo, hi in spans if hi > lo)
    g = np.array([0, 1_249_999, 1_250_000, 9_999_999], dtype=np.uint32)
    assert S.owner_of(g, 10_000_000, 8).tolist() == [0, 0, 1, 7]
    pair = np.array([7 * 142_858 - 2, 7 * 142_858 - 1], dtype=np.uint32)      # an even/odd pair near a block seam
    assert len(set(S.owner_of(pair, 1_000_003, 7).tolist())) == 1
