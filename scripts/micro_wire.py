"""Micro-benchmark: the wire path's kernels (wire_scan_kernel + wire_apply_kernel) on a device-resident
batch of serialized eraftpb.Message frames (one synthetic round of 1M groups), CUDA events."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("raft-rs_b200")
B, W = pkg.binding, pkg.wire
n = int(os.environ.get("N", 1_000_000))
s = B.Synth(n, 0x5EED0003)
ar = B.Arena(n, n_rings=1, ring_records=4096)
ar.group_alloc_range(n); ar.load_columns(s.initial)
wb = W.WireBuffers(None, 5 * n + 64)
rounds = []
for _ in range(6):
    wb.encode(s.next_round())
    nb = (wb.n_bytes + 15) & ~15
    d_bytes = ar.device_alloc(nb + 16); d_off = ar.device_alloc(4 * (wb.n + 1)); d_st = ar.device_alloc(wb.n)
    ar.h2d(d_bytes, wb.bytes[:nb]); ar.h2d(d_off, wb.offsets[: wb.n + 1])
    rounds.append((d_bytes, wb.n_bytes, d_off, wb.n, d_st))
st = torch.cuda.Stream(); sh = st.cuda_stream
evs = []
with torch.cuda.stream(st):
    for r in rounds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ar.wire_apply_device(r[0], r[1], r[2], r[3], r[4], stream=sh)
        e1.record(st)
        evs.append((e0, e1))
    torch.cuda.synchronize()
us = [1e3 * a.elapsed_time(b) for a, b in evs]
stt = np.zeros(rounds[-1][3], dtype=np.uint8); ar.d2h(stt, rounds[-1][4])
print(f"wire decode+apply: {rounds[-1][3]} frames, {rounds[-1][1]} bytes: {min(us[1:]):.1f} us per batch (scan + apply), "
      f"{rounds[-1][1] / min(us[1:]) / 1e3:.0f} GB/s of wire bytes; statuses {np.bincount(stt >> 4).tolist()}")
