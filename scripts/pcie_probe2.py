"""Does the step overlap with host work?  begin -> (sleep | enqueue next) -> wait."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
B = importlib.import_module("raft-rs_b200").binding
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
cpus = bench.gpu_local_cpus(torch, 0)
if cpus: os.sched_setaffinity(0, cpus)
n = 1_000_000
T = int(os.environ.get("T", 32))
s = B.Synth(n, 0x5EED0003)
a = B.Arena(n, n_rings=T)
a.group_alloc_range(n); a.load_columns(s.initial)
pool = ThreadPoolExecutor(T)
bufs = [np.empty(5 * n + 64, dtype=B.APPEND_RESP_DTYPE) for _ in range(2)]
def split(recs):
    cuts = [0]
    for t in range(1, T):
        c = len(recs) * t // T
        while c < len(recs) and recs[c]["flags"] & B.REC_EXT: c += 1
        cuts.append(c)
    cuts.append(len(recs))
    return [recs[cuts[t]:cuts[t + 1]] for t in range(T)]
def enq(parts): list(pool.map(lambda t: a.enqueue(parts[t], ring=t), range(T)))
p0 = split(s.next_round(bufs[0])); p1 = split(s.next_round(bufs[1]))
enq(p0); a.step(B.STEP_READ_COMMITTED)
for mode in ("sleep", "enqueue", "enqueue"):
    p0 = split(s.next_round(bufs[0])); p1 = split(s.next_round(bufs[1]))
    t0 = time.perf_counter(); enq(p0); t1 = time.perf_counter()
    a.step_begin(B.STEP_READ_COMMITTED); t2 = time.perf_counter()
    if mode == "sleep": time.sleep(0.004)
    else: enq(p1)
    t3 = time.perf_counter(); a.step_wait(); t4 = time.perf_counter()
    print(f"T={T} {mode}: enqueue {1e3*(t1-t0):.2f}  begin {1e3*(t2-t1):.2f}  middle {1e3*(t3-t2):.2f}  wait {1e3*(t4-t3):.2f} ms")
    if mode != "sleep": a.step(B.STEP_READ_COMMITTED)
