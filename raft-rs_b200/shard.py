"""Multi-GPU sharding of raft groups.

Groups are independent (one ProgressTracker per Raft, raft.rs:267-274), so they shard with NO
data-path collective: rank r owns a contiguous block of group slots in its own arena, records are
routed by group on the host, and the only thing that crosses ranks is a handful of counters
(NCCL all-reduce on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def block_size(n_total: int, world: int) -> int:
    """Groups per rank: ceil(n / world), rounded up to an EVEN count -- a wide group (include/raftgpu.h
    raftgpu_group_alloc_wide) is two consecutive slots starting at an even one, so a block boundary must
    neither split the pair nor change the parity of the rank-local slot numbers."""
    per = (n_total + world - 1) // world
    return (per + 1) & ~1


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of groups owned by `rank`: stable group -> rank = g // block_size."""
    per = block_size(n_total, world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def owner_of(groups: np.ndarray, n_total: int, world: int) -> np.ndarray:
    return (groups // block_size(n_total, world)).astype(np.int64)


def route_records(recs: np.ndarray, n_total: int, world: int, rank: int) -> np.ndarray:
    """The records of `rank`'s groups, group field rebased to the rank-local slot, arrival order
    kept (an EXT record travels with its REJECT: it carries the same group)."""
    lo, hi = shard_bounds(n_total, world, rank)
    g = recs["group"]
    mine = recs[(g >= lo) & (g < hi)].copy()
    mine["group"] -= np.uint32(lo)
    return mine


def slice_columns(cols, lo: int, hi: int, new_columns):
    """Rank-local copy of host columns [lo, hi)."""
    n = hi - lo
    out = new_columns(n, n)
    for name, arr in vars(cols).items():
        if not isinstance(arr, np.ndarray):
            continue
        dst = getattr(out, name, None)
        if dst is None:
            continue
        dst[...] = arr[..., lo:hi]
    return out


def aggregate(dist, torch, sums: dict, maxes: dict, device=None):
    """All-reduce per-rank counters (SUM) and times (MAX); the only collective of the system."""
    ks, km = sorted(sums), sorted(maxes)
    t_sum = torch.tensor([float(sums[k]) for k in ks], dtype=torch.float64, device=device)
    t_max = torch.tensor([float(maxes[k]) for k in km], dtype=torch.float64, device=device)
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    return ({k: t_sum[i].item() for i, k in enumerate(ks)},
            {k: t_max[i].item() for i, k in enumerate(km)})
