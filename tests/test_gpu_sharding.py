"""N > 1 path with the GPU arena as compute: ONE global cfg5-style stream, routed by group over two ranks
(raft-rs_b200/shard.py, the host logic bench.py --gpus N relies on); each rank owns its slice of the groups in
its own arena (its own GPU when the box has two, else both on cuda:0), steps it through the end-to-end C-ABI
call (raftgpu_step_begin_records) and all-reduces the counters over gloo.  The union of the shards must equal
the UNSHARDED oracle bit for bit, and the aggregated counters the unsharded counters."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

N_TOTAL, ROUNDS, SEED = 200_000, 5, 0x5EED0005


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("raft-rs_b200")
    B, S = pkg.binding, importlib.import_module("raft-rs_b200.shard")
    synth = B.Synth(N_TOTAL, SEED)                      # every rank sees the same global stream
    lo, hi = S.shard_bounds(N_TOTAL, world, rank)
    n = hi - lo
    arena = B.Arena(n, device=rank % torch.cuda.device_count(), n_rings=4)
    assert arena.group_alloc_range(n) == 0
    arena.load_columns(S.slice_columns(synth.initial, lo, hi, B.new_columns))
    advanced = records = 0
    bitmaps = []
    for _ in range(ROUNDS):
        recs = S.route_records(synth.next_round(), N_TOTAL, world, rank)
        arena.step_begin_records(recs, B.STEP_READ_COMMITTED)
        r = arena.step_wait()
        advanced += r.n_advanced
        records += int(np.count_nonzero((recs["flags"] & B.REC_EXT) == 0))
        bitmaps.append(arena.step_results(n)[0].copy())
    cnt = arena.counters()
    assert cnt["recomputes"] == n * ROUNDS and cnt["advanced"] == advanced and cnt["records"] == records
    sums, _ = S.aggregate(dist, torch, {"recomputes": cnt["recomputes"], "advanced": advanced, "records": records},
                          {"elapsed": 1.0})
    got = arena.read_columns(n)
    np.savez(os.path.join(tmpdir, f"shard{rank}.npz"), lo=lo, hi=hi, committed=got.committed, matched=got.matched,
             next_idx=got.next_idx, peer_committed=got.peer_committed, pflags=got.pflags, last_index=got.last_index,
             bitmaps=np.stack(bitmaps), sums=np.array([sums["recomputes"], sums["advanced"], sums["records"]]))
    arena.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_ranks_one_global_stream_equals_the_unsharded_oracle(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import B, O, bitmap_to_bool
    synth = B.Synth(N_TOTAL, SEED)
    ref = O.copy_columns(synth.initial)
    want = np.zeros(3)
    want_adv = []
    for _ in range(ROUNDS):
        recs = synth.next_round().copy()
        O.arena_apply(ref, recs, mode=0)
        adv, bm, _, _ = O.arena_recompute(ref)
        want_adv.append(bitmap_to_bool(bm, N_TOTAL))
        want += [N_TOTAL, adv, np.count_nonzero((recs["flags"] & B.REC_EXT) == 0)]
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"shard{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        n = hi - lo
        covered += n
        for name in ("committed", "last_index"):
            assert np.array_equal(z[name][:n], getattr(ref, name)[lo:hi]), name
        for name in ("matched", "next_idx", "peer_committed", "pflags"):
            assert np.array_equal(z[name][:, :n], getattr(ref, name)[:, lo:hi]), name
        for k in range(ROUNDS):     # the advanced bitmap of every round, rebased to the shard
            assert np.array_equal(bitmap_to_bool(z["bitmaps"][k], n), want_adv[k][lo:hi]), (r, k)
        assert np.array_equal(z["sums"], want)
    assert covered == N_TOTAL
