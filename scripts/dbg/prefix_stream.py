"""Diagnostics: a batch that only covers the first groups of a large arena, through raftgpu_step_begin_records."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
B = importlib.import_module("raft-rs_b200").binding
n = int(os.environ.get("N", 1_000_000))
frac = float(os.environ.get("FRAC", 0.01))
flags = int(os.environ.get("FLAGS", 0))
s = B.Synth(n, 0x77, k_peers=5)
a = B.Arena(n)
a.group_alloc_range(n); a.load_columns(s.initial)
recs = s.next_round().copy()
k = int(np.searchsorted(recs["group"], int(frac * n)))
while recs["flags"][k] & B.REC_EXT: k += 1
part = np.ascontiguousarray(recs[:k])
print("records", len(part), "of", len(recs), "flags", flags, flush=True)
a.step_begin_records(part, B.STEP_READ_COMMITTED | flags)
r = a.step_wait()
print("ok: advanced", r.n_advanced, "records", r.n_records, "h2d", r.h2d_bytes, flush=True)
from oracle import oracle as O
ref = O.copy_columns(s.initial)
O.arena_apply(ref, part, mode=0)
adv, bm, _, _ = O.arena_recompute(ref)
got = a.read_columns(n)
print("oracle advanced", adv)
for name in ("matched", "next_idx", "peer_committed", "pflags", "committed", "last_index"):
    x, y = getattr(got, name), getattr(ref, name)[..., :n]
    bad = np.argwhere(x != y)
    if len(bad):
        gs = np.unique(bad[:, -1])
        print(name, "differs in", len(gs), "groups; first", gs[:8].tolist(), "last", gs[-3:].tolist(), "tiles", np.unique(gs // 256)[:10].tolist())
gb, _ = a.step_results(n)
d = np.nonzero(gb != bm[: len(gb)])[0]
print("bitmap words differing:", len(d), d[:10].tolist())
