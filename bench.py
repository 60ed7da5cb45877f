#!/usr/bin/env python
"""bench.py -- commit-index recomputes/s on synthetic AppendResponse streams.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU baseline arm (the oracle port)

One "step" = one pass of the hot path over one batch: apply ONE round of synthetic
AppendResponses (about 3.5 records per group) to a 1M-group x 5-peer arena and
recompute the commit index of every group (Raft::maybe_commit).  `value` = groups recomputed
per second over all GPUs with the records already resident in HBM; `e2e` = the same through
raftgpu_enqueue_append_resp / raftgpu_step with HOST buffers (H2D + D2H inside the timed
region).  Multi-GPU: groups shard across ranks, no data-path collective; NCCL only reduces
the per-rank counters and times.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "commit_index_recomputes_per_s"
UNIT = "recomputes/s"
N_GROUPS = 1_000_000          # BASELINE configs[2]: the headline config, per GPU
K_PEERS = 5
SEED = 0x5EED0003
B_ALG_RECOMPUTE = 8 * K_PEERS + 34   # 74 B per recompute at K = 5
B_ALG_APPLY = 76                      # bytes per applied AppendResponse
N_ARENAS = 4                          # rotated so consecutive steps never share L2 contents


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the measured phases run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, f"/tmp/raftgpu_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[4:8]):
                    if v == "Active":
                        reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def gpu_local_cpus(torch, device: int):
    """CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs local_cpulist)."""
    try:
        pr = torch.cuda.get_device_properties(device)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        cpus = set()
        for tok in open(path).read().strip().split(","):
            lo, _, hi = tok.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus & os.sched_getaffinity(0)
    except Exception:
        return set()


def cpu_leg(n_groups, seed, rounds_wanted, threads, budget_s=20.0, joint=False):
    """The oracle (oracle/raft_oracle.c: apply + recompute, range-partitioned over `threads`
    pthreads) on a bounded sample of the same workload.  Only used as the CPU baseline."""
    B = importlib.import_module("raft-rs_b200").binding
    from oracle import oracle as O
    synth = B.Synth(n_groups, seed, k_peers=K_PEERS, joint=joint)
    cols = O.copy_columns(synth.initial)
    total, done, times = 0.0, 0, []
    for _ in range(rounds_wanted):
        recs = synth.next_round()
        secs, _ = O.bench_step(cols, recs, threads, fast=True)
        times.append(secs)
        total += secs
        done += 1
        if total > budget_s:
            break
    return n_groups * done / total, done, times


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  raft-rs is Rust and
    this image has no rustc/cargo, so the arm runs the pinned C port (oracle/) with every host
    thread, on the same config / metric as the GPU arm."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    # warmup rounds are part of the same stream; time exactly `steps` rounds after them
    B = importlib.import_module("raft-rs_b200").binding
    from oracle import oracle as O
    joint = args.workload == "cfg4"
    seed0 = 0x5EED0004 if joint else SEED
    synth = B.Synth(N_GROUPS, seed0, k_peers=K_PEERS, joint=joint)
    cols = O.copy_columns(synth.initial)
    for _ in range(args.warmup):
        O.bench_step(cols, synth.next_round(), threads, fast=True)
    total = 0.0
    for _ in range(args.steps):
        secs, _ = O.bench_step(cols, synth.next_round(), threads, fast=True)
        total += secs
    value = N_GROUPS * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": ("cfg4: 1M raft groups x 7 peer slots under joint consensus, one synthetic "
                                "AppendResponse round per step (apply + recompute)") if joint else
                               ("cfg3: 1M raft groups x 5 peers, one synthetic AppendResponse round "
                                "per step (apply + recompute)"), "groups": N_GROUPS,
                   "peers": 7 if joint else K_PEERS, "seed": hex(seed0)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} rounds of the {args.workload} stream, {threads} pthreads, "
                                   "oracle/raft_oracle.c tuned path (ro_bench_step_fast, == the literal port)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--groups", type=int, default=N_GROUPS, help=argparse.SUPPRESS)
    ap.add_argument("--compact-device", action="store_true",
                    help="device-resident leg on the compact stream + its fused kernel (raftgpu_step_compact_device) "
                         "instead of the 16-byte packed records + per-record fused kernel (measured: 86 vs 75 us)")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4"],
                    help="cfg3 (default, the headline): 1M groups x 5 peers; cfg4: 1M groups x 7 peer slots under "
                         "joint consensus (incoming {0..4}, outgoing {0,1,2,5,6}), 90 B per recompute")
    ap.add_argument("--e2e-threads", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed e2e steps (default: --steps)")
    ap.add_argument("--e2e-chunk", type=int, default=8, help="pipelined e2e steps per timed chunk")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scatter", action="store_true",
                    help="device-resident leg uses the general two-kernel path (scatter apply + recompute) "
                         "instead of the fused tile kernel for group-ordered batches")
    ap.add_argument("--public-records", action="store_true",
                    help="device-resident leg reads 24-byte public records instead of the packed 16-byte form")
    ap.add_argument("--profile", action="store_true",
                    help="device-resident loop only (for ncu): no clock warm loop, no e2e, no CPU leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    B = importlib.import_module("raft-rs_b200").binding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference "
                         "for the CPU baseline arm)")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    local_cpus = gpu_local_cpus(torch, local_rank)
    if local_cpus:
        # like `numactl --cpunodebind`: host staging threads and their buffers next to the GPU
        os.sched_setaffinity(0, local_cpus)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.groups
    K, W = args.steps, args.warmup
    joint = args.workload == "cfg4"
    k_union = 7 if joint else K_PEERS
    seed0 = 0x5EED0004 if joint else SEED
    b_alg_recompute = 8 * k_union + 34
    peak_gbs, peak_src = peaks()

    # ---- synthetic inputs: N_ARENAS independent 1M-group stores, K+W rounds in total ----------
    # Host memory is kept small and reused (one record buffer per generator): fresh host pages
    # are slow on these VMs; HBM holds all K+W rounds.
    per_arena = [(W + K + N_ARENAS - 1 - a) // N_ARENAS for a in range(N_ARENAS)]
    arenas, round_len, d_recs, d_offs = [], [], [], []
    fused = not (args.scatter or args.public_records)
    compact = fused and args.compact_device    # compact stream + its fused kernel for the device-resident leg
    pack_buf = np.empty((7 * n + 64, 2), dtype=np.uint64)
    blob_buf = np.empty(B.compact_bound(7 * n + 64), dtype=np.uint8) if compact else None
    d_bad = None
    for a in range(N_ARENAS):
        seed = seed0 + 0x100 * a + 0x10000 * rank
        s = B.Synth(n, seed, k_peers=K_PEERS, joint=joint)
        ar = B.Arena(n, device=local_rank, n_rings=1, ring_records=4096)
        assert ar.group_alloc_range(n) == 0
        ar.load_columns(s.initial)
        ptrs, lens, offs = [], [], []
        for _ in range(per_arena[a]):
            recs = s.next_round()
            if compact:                   # the compact stream (4-byte units) + its tile index, both in HBM
                nb, _ = B.pack_compact(recs, blob_buf)
                hdr = blob_buf[:64].copy()
                assert hdr.view(B.COMPACT_HDR_DTYPE)[0]["flags"] & B.COMPACT_TILEABLE
                p = ar.device_alloc(nb)
                ar.h2d(p, blob_buf[:nb])
                po = ar.device_alloc(4 * (3 * (ar.cap // B.tile_groups() + 2) + 2))
                if d_bad is None:
                    d_bad = ar.device_alloc(4)
                    ar.h2d(d_bad, np.zeros(1, dtype=np.uint32))
                ar.compact_tile_index_device(p, hdr, po, d_bad)
                offs.append((po, hdr))
                lens.append((nb, len(recs)))
            elif args.public_records:     # the 24-byte public record layout in HBM
                p = ar.device_alloc(recs.nbytes)
                ar.h2d(p, recs)
                lens.append((len(recs), len(recs)))
            else:                         # the packed 16-byte wire form (what the staging path ships)
                k = ar.pack_records(recs, pack_buf)
                p = ar.device_alloc(16 * k)
                ar.h2d(p, pack_buf[:k])
                lens.append((k, len(recs)))
                if fused:                 # the batch is in group order: its 256-group tile index
                    off = B.tile_index(pack_buf, k, n)
                    po = ar.device_alloc(off.nbytes)
                    ar.h2d(po, off)
                    offs.append(po)
            ptrs.append(p)
        arenas.append(ar)
        round_len.append(lens)
        d_recs.append(ptrs)
        d_offs.append(offs)
        del s
    schedule = [(i % N_ARENAS, i // N_ARENAS) for i in range(W + K)]  # (arena, round) per step

    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    def run_step(i, ev=None):
        a, r = schedule[i]
        if ev:
            ev[0].record(stream)
        if fused:   # ONE kernel: apply + recompute on shared-memory tiles (group-ordered batch)
            if compact:
                arenas[a].step_compact_device(d_recs[a][r], d_offs[a][r][1], d_offs[a][r][0], stream=sh)
            else:
                arenas[a].step_sorted_device(d_recs[a][r], round_len[a][r][0], d_offs[a][r], stream=sh)
            if ev:
                ev[1].record(stream)
                ev[2].record(stream)
            return
        if args.public_records:
            arenas[a].apply_device(d_recs[a][r], round_len[a][r][0], stream=sh)
        else:
            arenas[a].apply_device_packed(d_recs[a][r], round_len[a][r][0], stream=sh)
        if ev:
            ev[1].record(stream)
        arenas[a].recompute(0, n, stream=sh)
        if ev:
            ev[2].record(stream)

    sampler = ClockSampler(local_rank)
    sampler.start()
    with torch.cuda.stream(stream):
        # clocks up: the recompute pass is idempotent on unchanged progress
        t_end = time.perf_counter() + (0.0 if args.profile else 0.3)
        while time.perf_counter() < t_end:
            for a in arenas:
                a.recompute(0, n, stream=sh)
            stream.synchronize()
        for i in range(W):
            run_step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            run_step(W + i, evs[i])
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    ms_apply = sum(e[0].elapsed_time(e[1]) for e in evs)
    ms_recompute = sum(e[1].elapsed_time(e[2]) for e in evs)
    n_records = sum(round_len[a][r][1] for a, r in schedule[W:])

    # ---- e2e: host buffers -> enqueue -> step (H2D, kernels, D2H) on a fresh arena ------------
    # The caller's records sit in ordinary host memory; raftgpu_enqueue_append_resp stages them
    # into pinned rings (T caller threads, T rings), raftgpu_step_begin DMAs + launches,
    # raftgpu_step_wait returns once the results are back in host memory.  Steps are timed in
    # chunks of `chunk` pipelined steps (the next batch is staged while one is in flight); the
    # records of the following chunk are regenerated between chunks, untimed, into the same few
    # host buffers.
    # staging threads per rank: one per physical GPU-local core, shared with the other ranks whose
    # GPU hangs off the same socket (two sockets per host)
    ranks_per_node = max(1, (world + 1) // 2)
    e2e_threads = args.e2e_threads or max(4, min(32, (len(local_cpus) or 64) // (2 * ranks_per_node)))
    os.environ.setdefault("RAFTGPU_HOST_THREADS", str(e2e_threads))
    chunk = max(2, args.e2e_chunk)
    e2e_steps = 0 if args.profile else (args.e2e_steps or K)
    es = B.Synth(n, seed0 + 0x10000 * rank, k_peers=K_PEERS, joint=joint)
    ea = B.Arena(n, device=local_rank, n_rings=e2e_threads)
    assert ea.group_alloc_range(n) == 0
    ea.load_columns(es.initial)
    bufs = [np.empty((7 if joint else 5) * n + 64, dtype=B.APPEND_RESP_DTYPE) for _ in range(chunk)]

    def split(recs):
        return recs

    def staged_leg(records_api):
        """Timed: the caller's 24-byte records (ordinary host memory) -> library staging threads ->
        H2D -> kernels -> D2H.  records_api: raftgpu_step_begin_records (compact stream, one call);
        else the general path raftgpu_enqueue_bulk(SORTED) + raftgpu_step_begin (16-byte records)."""
        def submit(recs):
            if records_api:
                ea.step_begin_records(recs, flags)
                return time.perf_counter()
            ea.enqueue_bulk(recs, sorted_by_group=True)
            t_mid = time.perf_counter()
            ea.step_begin(flags)
            return t_mid

        secs, timed, caller_b, first_chunk = 0.0, 0, 0, True
        phase = [0.0, 0.0, 0.0]   # host seconds in staging / step_begin / step_wait
        dma = [0, 0]              # bytes actually DMAed (h2d, d2h), as reported by the library
        while timed < e2e_steps:
            m = min(chunk, e2e_steps - timed)
            parts = [es.next_round(bufs[j]) for j in range(m)]      # untimed generation
            if first_chunk:                                           # untimed warm-up steps
                for j in range(m):
                    submit(parts[j])
                    ea.step_wait()
                first_chunk = False
                continue
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            t1 = submit(parts[0])
            phase[0] += t1 - t0
            phase[1] += time.perf_counter() - t1
            for j in range(m):
                ta = time.perf_counter()
                if j + 1 < m:             # stage AND submit the next batch while this one is in flight:
                    tb = submit(parts[j + 1])   # its H2D overlaps this step's kernels + D2H
                else:
                    tb = ta
                tc = time.perf_counter()
                sr = ea.step_wait()
                dma[0] += sr.h2d_bytes
                dma[1] += sr.d2h_bytes
                td = time.perf_counter()
                phase[0] += tb - ta
                phase[1] += tc - tb
                phase[2] += td - tc
            secs += time.perf_counter() - t0
            timed += m
            caller_b += sum(pj.nbytes for pj in parts)
        return {"seconds": secs, "steps": timed, "h2d": dma[0] / max(1, timed), "d2h": dma[1] / max(1, timed),
                "caller_bytes": caller_b / max(1, timed), "phase": [1e3 * x / max(1, timed) for x in phase]}

    flags = B.STEP_READ_COMMITTED
    sr_ = staged_leg(False) if e2e_steps else {"seconds": 0.0, "steps": 0}   # e2e_staged: the general staging path
    sg = staged_leg(True) if e2e_steps else {"seconds": 0.0, "steps": 0}     # e2e_records_api: one-call compact staging
    e2e_s, e2e_timed = sr_["seconds"], sr_["steps"]
    # ---- e2e, zero-copy: the caller builds its batch (packed 16-byte records) directly in the
    # arena's NUMA-local pinned memory (untimed, like the generation above); timed is
    # raftgpu_step_begin_packed (H2D straight from that buffer, the GPU verifies the one-wave
    # promise) + raftgpu_step_wait, two steps in flight.
    def zero_copy_leg(compact):
        if compact:
            cap_b = B.compact_bound((7 if joint else 5) * n + 64)
            pk = [ea.host_alloc_bytes(cap_b) for _ in range(chunk)]
        else:
            pk = [ea.host_alloc_packed((7 if joint else 5) * n + 64) for _ in range(chunk)]
        zc_s, zc_timed, zc_dma = 0.0, 0, [0, 0]
        while zc_timed < e2e_steps:
            m = min(chunk, e2e_steps - zc_timed)
            if compact:   # untimed, like the generation of the records themselves
                ks = [B.pack_compact(es.next_round(bufs[j]), pk[j])[0] for j in range(m)]
                begin = ea.step_begin_compact
            else:
                ks = [ea.pack_records(es.next_round(bufs[j]), pk[j]) for j in range(m)]
                begin = ea.step_begin_packed
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            begin(pk[0], ks[0], flags)
            for j in range(m):
                if j + 1 < m:
                    begin(pk[j + 1], ks[j + 1], flags)
                sr = ea.step_wait()
                zc_dma[0] += sr.h2d_bytes
                zc_dma[1] += sr.d2h_bytes
            zc_s += time.perf_counter() - t0
            zc_timed += m
        for b_ in pk:
            ea.host_free(b_)
        return {"seconds": zc_s, "steps": zc_timed, "h2d": zc_dma[0] / zc_timed, "d2h": zc_dma[1] / zc_timed}

    zc, zp = {"value": None}, {"value": None}
    if e2e_steps:
        zp = zero_copy_leg(False)
        zc = zero_copy_leg(True)
    clocks = sampler.stop()

    # ---- aggregate over ranks (NCCL: counters and times only) -----------------------------------
    if os.environ.get("RAFTGPU_TILE_DEBUG") and rank == 0:
        d = sum(a.debug_read().astype(np.float64) for a in arenas)
        if d[4]:
            print("[tile debug] cycles per tile: wait_loads %.0f  records %.0f  recompute %.0f  stores %.0f  (tiles %d)"
                  % (d[0] / d[4], d[1] / d[4], d[2] / d[4], d[3] / d[4], d[4]), file=sys.stderr)
    cnt = [a.counters() for a in arenas]
    S = importlib.import_module("raft-rs_b200.shard")
    sums, maxes = S.aggregate(
        dist if world > 1 else None, torch,
        {"groups_device": n * K, "groups_e2e": n * e2e_timed,
         "recomputes": sum(c["recomputes"] for c in cnt), "advanced": sum(c["advanced"] for c in cnt),
         "records": sum(c["records"] for c in cnt)},
        {"ms_total": ms_total, "e2e_s": e2e_s, "sg_s": sg.get("seconds", 0.0), "zc_s": zc.get("seconds", 0.0),
         "zp_s": zp.get("seconds", 0.0)},
        device="cuda")
    ms_max, e2e_max = maxes["ms_total"], maxes["e2e_s"]
    value = sums["groups_device"] / (ms_max * 1e-3)
    e2e_value = sums["groups_e2e"] / e2e_max if e2e_max > 0 else None

    if rank == 0:
        kernels = []
        if fused:
            klist = (("step_tile_compact_kernel" if compact else "step_tile_kernel", ms_apply, n_records * B_ALG_APPLY + n * K * b_alg_recompute),)
        else:
            klist = (("apply_kernel", ms_apply, n_records * B_ALG_APPLY),
                     ("recompute_kernel", ms_recompute, n * K * b_alg_recompute))
        for name, ms, alg_bytes in klist:
            gbs = alg_bytes / (ms * 1e-3) / 1e9
            kernels.append({"kernel": name, "avg_us": 1e3 * ms / K, "share": ms / ms_total,
                            "alg_bytes_per_launch": alg_bytes / K, "achieved": gbs,
                            "frac": gbs / peak_gbs})
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            traffic = {}
        for k in kernels:
            k["traffic"] = traffic.get(k["kernel"])   # ncu dram bytes per launch (profiles/)
        dom = max(kernels, key=lambda k: k["avg_us"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": ("cfg4: 1M raft groups x 7 peer slots per GPU under joint consensus (two 5-voter "
                             "majorities), one synthetic AppendResponse round per step (apply + recompute)")
                if joint else ("cfg3: 1M raft groups x 5 peers per GPU, one synthetic AppendResponse "
                               "round per step (apply + recompute)"),
                "groups_per_gpu": n, "peers": k_union, "seed": hex(seed0),
                "records_per_step": n_records / K,
                "record_format": "24 B public" if args.public_records else
                                 ("compact stream, 4 B units (raftgpu_pack_compact)" if compact else
                                  "16 B packed (raftgpu_pack_records)"),
                "device_path": "fused tile kernel (raftgpu_step_sorted_device, group-ordered batch + tile index)"
                               if fused else "scatter apply + recompute (raftgpu_apply_device[_packed] + raftgpu_recompute)",
                "l2": f"inputs larger than L2: {N_ARENAS} arenas rotated, fresh records every step",
                "parallelism": f"groups sharded over {world} GPU(s), no data-path collective",
            },
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"],
                         "peak": peak_gbs, "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"],
                         "peak_source": peak_src},
            "kernels": kernels,
            "e2e_staged": None if not e2e_timed else {
                "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sr_["h2d"], "d2h_bytes_per_step": sr_["d2h"],
                "steps": e2e_timed, "ms_per_step": 1e3 * e2e_max / max(1, e2e_timed),
                "caller_record_bytes_per_step": sr_["caller_bytes"],
                "host_threads": e2e_threads, "pipelined_chunk": chunk, "host_cpus_bound": len(local_cpus) or None,
                "host_ms_per_step": {"staging": sr_["phase"][0], "step_begin": sr_["phase"][1], "step_wait": sr_["phase"][2]},
                "api": "raftgpu_enqueue_bulk(SORTED) + raftgpu_step_begin/_wait (READ_COMMITTED): 24-byte records in "
                       "ordinary host memory, copied + packed to 16-byte records by the library's staging threads "
                       "inside the timed region"},
            "e2e_records_api": None if not sg.get("steps") else {
                "value": world * n * sg["steps"] / maxes["sg_s"], "unit": UNIT,
                "ms_per_step": 1e3 * maxes["sg_s"] / sg["steps"], "steps": sg["steps"],
                "h2d_bytes_per_step": sg["h2d"], "d2h_bytes_per_step": sg["d2h"],
                "host_ms_per_step": {"staging": sg["phase"][0], "step_begin": sg["phase"][1], "step_wait": sg["phase"][2]},
                "api": "raftgpu_step_begin_records + raftgpu_step_wait: one call, the staging threads pack slices of "
                       "the batch into the compact stream (host-bound: ~12 ns per record per thread)"},
            "e2e": None if not zc.get("steps") else {
                "value": world * n * zc["steps"] / maxes["zc_s"], "unit": UNIT,
                "ms_per_step": 1e3 * maxes["zc_s"] / zc["steps"], "steps": zc["steps"],
                "h2d_bytes_per_step": zc["h2d"], "d2h_bytes_per_step": zc["d2h"],
                "api": "raftgpu_step_begin_compact + raftgpu_step_wait: the step's records sit in pinned host "
                       "memory (raftgpu_host_alloc) as the compact stream the caller built them in "
                       "(raftgpu_pack_compact, untimed like the generation of the records); timed: H2D of the "
                       "stream, apply (device-side decode + one-wave check) + recompute kernels, D2H of the "
                       "advanced bitmap and new commit indexes; two steps in flight.  e2e_packed16 is the same "
                       "with the 16-byte packed form (raftgpu_step_begin_packed); e2e_staged goes through "
                       "raftgpu_enqueue_bulk, i.e. with the library copying + packing 24-byte records from "
                       "pageable memory first"},
            "e2e_packed16": None if not zp.get("steps") else {
                "value": world * n * zp["steps"] / maxes["zp_s"], "unit": UNIT,
                "ms_per_step": 1e3 * maxes["zp_s"] / zp["steps"], "steps": zp["steps"],
                "h2d_bytes_per_step": zp["h2d"], "d2h_bytes_per_step": zp["d2h"]},
            "gpu_launches": (1 if fused else 2) * K,
            "clocks": clocks,
            "counters": {"recomputes": sums["recomputes"], "advanced": sums["advanced"], "records": sums["records"]},
        }
        if world == 1 and not args.no_cpu_baseline and not args.profile:
            os.sched_setaffinity(0, all_cpus)   # the CPU baseline gets every host core
            threads = len(all_cpus)
            v, done, _ = cpu_leg(n, seed0, 64, threads, budget_s=15.0, joint=joint)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{done} rounds of the same {args.workload} stream (apply + recompute), {threads} "
                          "pthreads, oracle/raft_oracle.c tuned path (ro_bench_step_fast, == the literal port)"}
        print(json.dumps(line), flush=True)
    for a in arenas:
        a.close()
    ea.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
