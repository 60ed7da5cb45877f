"""Micro-benchmark of send_list_kernel: 1M groups x 5 peers, the advanced bitmap of
a real step, 4 arenas rotated; CUDA-event timing.  Prints one line for DESIGN.md."""
import importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = importlib.import_module("raft-rs_b200").binding

n, A = 1_000_000, 4
arenas, bms, outs, cnts, totals = [], [], [], [], []
for a in range(A):
    s = B.Synth(n, 0xABC0 + a)
    ar = B.Arena(n)
    ar.group_alloc_range(n)
    ar.load_columns(s.initial)
    for _ in range(2):
        ar.enqueue_bulk(s.next_round().copy(), B.BULK_SORTED) if hasattr(ar, "enqueue_bulk") else ar.enqueue(s.next_round().copy())
        ar.step(0)
    bm, _ = ar.step_results(n)
    d_bm = ar.device_alloc(4 * len(bm))
    ar.h2d(d_bm, np.ascontiguousarray(bm))
    arenas.append(ar); bms.append(d_bm)
    outs.append(ar.device_alloc(16 * 8 * n)); cnts.append(ar.device_alloc(8))
    totals.append(int(np.unpackbits(bm.view(np.uint8)).sum()))
stream = torch.cuda.Stream()
for bitmap in (True, False):
    for w in range(8):
        a = w % A
        arenas[a].send_list_device(0, n, bms[a] if bitmap else None, outs[a], 8 * n, cnts[a], stream=stream.cuda_stream)
    torch.cuda.synchronize()
    K = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(K):
        a = k % A
        arenas[a].send_list_device(0, n, bms[a] if bitmap else None, outs[a], 8 * n, cnts[a], stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / K
    c = np.zeros(1, dtype=np.uint64)
    arenas[0].d2h(c, cnts[0])
    print(f"send_list_kernel, 1M groups x 5 peers, {'advanced bitmap (%d groups advanced)' % totals[0] if bitmap else 'no bitmap (bcast_append for every group)'}: "
          f"{us:.1f} us per pass (memset + kernel), {int(c[0])} entries -> {c[0] / us:.0f} entries/us, {n / us * 1e6:.3e} groups/s")
