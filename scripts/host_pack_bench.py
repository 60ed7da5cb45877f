#!/usr/bin/env python
"""Host-only timing of the 24-byte-record -> compact-stream packer (no GPU needed).

    python scripts/host_pack_bench.py [n_groups] [rounds]

Prints ns per record for raftgpu_pack_compact (single thread) and, when the library exports it,
for the multi-threaded slice packer raftgpu_pack_compact_mt at several thread counts.  Used to
size the staging path of raftgpu_step_begin_records."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = importlib.import_module("raft-rs_b200").binding


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    s = B.Synth(n, 0x5EED0003, k_peers=5)
    out = np.empty(B.compact_bound(5 * n + 64), dtype=np.uint8)
    bufs = [s.next_round().copy() for _ in range(rounds)]
    for r in bufs[:1]:
        B.pack_compact(r, out)
    ts = []
    for r in bufs:
        t0 = time.perf_counter()
        nb, _ = B.pack_compact(r, out)
        ts.append((time.perf_counter() - t0) / len(r))
    print(f"pack_compact 1 thread: {1e9 * min(ts):.2f} ns/record (best of {rounds}), "
          f"{len(bufs[0])} records -> {nb} bytes ({nb / len(bufs[0]):.2f} B/record)")
    L = B.lib()
    if hasattr(L, "raftgpu_pack_compact_mt"):
        for T in (1, 2, 4, 8, 16, 32, 64):
            if T > 2 * (os.cpu_count() or 1):
                break
            ts = []
            for r in bufs:
                t0 = time.perf_counter()
                nb = B.pack_compact_mt(r, out, T)
                ts.append(time.perf_counter() - t0)
            print(f"pack_compact_mt T={T:3d}: {1e3 * min(ts):.3f} ms per round, "
                  f"{1e9 * min(ts) * T / len(bufs[0]):.2f} ns/record/thread")


if __name__ == "__main__":
    main()
