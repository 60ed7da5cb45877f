"""The compact stream (include/raftgpu.h "compact stream", raftgpu_pack_compact in csrc/arena.cu,
load_compact in kernels.cuh) is lossless: decoding the blob the way the kernel does gives back the
public 24-byte records -- field for field, in order -- for ANY input (hostile values go through
the ESC side table), and a synthetic multi-raft round shrinks to ~22 bytes per group.  CPU only
(host functions of the library, no device call)."""
import numpy as np
import pytest

from helpers import B

M64 = (1 << 64) - 1


def pack(recs):
    out = np.zeros(B.compact_bound(len(recs)), dtype=np.uint8)
    nb, units = B.pack_compact(recs, out, want_units=True)
    return out[:nb], units


def normalise(recs):
    """What the stream promises to carry: main records in order, each REJECT with its EXT
    (next_probe_index hint, request_snapshot) -- a REJECT without one means (0, none), which is how
    the kernels read it (load_reject_ext); stray EXT records carry nothing and are dropped."""
    out = []
    for i, r in enumerate(recs):
        if r["flags"] & B.REC_EXT:
            continue
        out.append((r["group"], r["peer_slot"], r["flags"], 0, r["index"], r["commit"]))
        if r["flags"] & B.REC_REJECT:
            if i + 1 < len(recs) and (recs[i + 1]["flags"] & B.REC_EXT):
                e = recs[i + 1]
                out.append((r["group"], r["peer_slot"], B.REC_EXT, 0, e["index"], e["commit"]))
            else:
                out.append((r["group"], r["peer_slot"], B.REC_EXT, 0, 0, 0))
    return np.array(out, dtype=B.APPEND_RESP_DTYPE)


def check_round_trip(recs):
    blob, units = pack(recs)
    got = normalise(B.unpack_compact(blob))
    want = normalise(recs)
    assert len(got) == len(want)
    ext = (want["flags"] & B.REC_EXT) != 0
    for name in ("group", "peer_slot", "flags", "index", "commit"):
        if name in ("group", "peer_slot"):     # an EXT's own group / slot fields mean nothing
            np.testing.assert_array_equal(got[name][~ext], want[name][~ext], err_msg=name)
        else:
            np.testing.assert_array_equal(got[name], want[name], err_msg=name)
    hdr = blob[:B.COMPACT_HDR_DTYPE.itemsize].view(B.COMPACT_HDR_DTYPE)[0]
    assert hdr["total_bytes"] == len(blob)
    assert hdr["n_records"] == int(np.count_nonzero(~(recs["flags"] & B.REC_EXT).astype(bool)))
    # unit_of_record: every main record has a unit, units are distinct and increasing
    main = ~(recs["flags"] & B.REC_EXT).astype(bool)
    u = units[main].astype(np.int64)
    assert np.all(np.diff(u) > 0) and (len(u) == 0 or u[-1] < hdr["n_units"])
    assert np.all(units[~main] == 0xFFFFFFFF)
    return hdr


def test_empty_batch():
    hdr = check_round_trip(np.zeros(0, dtype=B.APPEND_RESP_DTYPE))
    assert hdr["n_units"] == 0 and hdr["n_side"] == 0


def hostile_records(seed=5):
    rng = np.random.default_rng(seed)
    edge = [0, 1, 2, 254, 255, 256, 0x3ffe, 0x3fff, 0x4000, 0x7ffe, 0x7fff, 0x8000, (1 << 28) - 1, 1 << 28, (1 << 28) + 1, (1 << 30) - 1, 1 << 30, (1 << 48) - 1, 1 << 48,
            (1 << 48) + 0x7fff, (1 << 48) + 0x8000, 1 << 63, M64 - 1, M64]
    rows = []
    g = 0
    for _ in range(3000):
        g += int(rng.integers(0, 3)) if rng.random() < 0.95 else int(rng.integers(0, 1 << 16))
        g &= 0xFFFFFFFF
        base = edge[int(rng.integers(0, len(edge)))] if rng.random() < 0.3 else int(rng.integers(0, 1 << 50))
        for _ in range(int(rng.integers(1, 12))):   # runs longer than 8 split
            kind = rng.random()
            off = int(rng.integers(0, 40000)) if rng.random() < 0.2 else int(rng.integers(0, 64))
            index = max(0, base - off) if rng.random() < 0.8 else min(M64, base + off)
            slot = int(rng.integers(0, 8)) if rng.random() < 0.97 else int(rng.integers(8, 256))
            if kind < 0.7:
                cd = int(rng.integers(0, 4)) if rng.random() < 0.8 else int(rng.integers(250, 260))
                commit = max(0, index - cd) if rng.random() < 0.9 else min(M64, index + cd + 1)
                rows.append((g, slot, 0, 0, index, commit))
            elif kind < 0.85:
                cd = int(rng.integers(0, 64)) if rng.random() < 0.8 else int(rng.integers(250, 260))
                commit = 0 if rng.random() < 0.2 else min(M64, index + cd)
                if rng.random() < 0.05:
                    commit = max(1, index - 1)     # a "new last_index" below index: still lossless
                rows.append((g, slot, B.REC_LOCAL, 0, index, commit))
            else:
                cd = int(rng.integers(0, 6)) if rng.random() < 0.7 else int(rng.integers(0, 1 << 40))
                rows.append((g, slot, B.REC_REJECT, 0, index, max(0, index - cd)))
                if rng.random() < 0.8:
                    hd = int(rng.integers(-8, 8)) if rng.random() < 0.7 else \
                        int(rng.choice([-(1 << 28) - 1, -(1 << 28), (1 << 28) - 1, 1 << 28, 1 << 40]))
                    rows.append((g, slot, B.REC_EXT, 0, min(M64, max(0, index + hd)),
                                 0 if rng.random() < 0.7 else int(rng.integers(1, 1 << 40))))
    return np.array(rows, dtype=B.APPEND_RESP_DTYPE)


def test_round_trip_random_records():
    hdr = check_round_trip(hostile_records())
    assert hdr["n_side"] > 0


def test_vector_packer_is_byte_identical_to_the_scalar_definition(tmp_path):
    """pack_compact.cpp has an AVX-512 form (whole runs per step) and the scalar state machine that
    defines the format; RAFTGPU_PACK_SCALAR=1 forces the latter.  Same bytes for hostile records,
    for synthetic rounds, and when the batch is packed in slices (as the staging threads do)."""
    import os
    import subprocess
    import sys
    batches = [hostile_records(11), hostile_records(12), B.Synth(30000, 0xABCDE, k_peers=5).next_round().copy(),
               B.Synth(9000, 3, k_peers=5, joint=True).next_round().copy()]
    for k, recs in enumerate(batches):
        np.save(tmp_path / f"recs{k}.npy", recs)
        blob, units = pack(recs)
        np.save(tmp_path / f"blob{k}.npy", blob)
        np.save(tmp_path / f"units{k}.npy", units)
    code = f"""
import sys, numpy as np
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r})
from test_compact_format import pack
for k in range({len(batches)}):
    recs = np.load({str(tmp_path)!r} + f"/recs{{k}}.npy")
    blob, units = pack(recs)
    assert np.array_equal(blob, np.load({str(tmp_path)!r} + f"/blob{{k}}.npy")), k
    assert np.array_equal(units, np.load({str(tmp_path)!r} + f"/units{{k}}.npy")), k
print("same")
"""
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RAFTGPU_PACK_SCALAR="1"),
                       capture_output=True, text=True)
    assert r.returncode == 0 and "same" in r.stdout, r.stderr


def test_unsorted_and_stray_ext():
    rng = np.random.default_rng(6)
    n = 5000
    recs = np.zeros(n, dtype=B.APPEND_RESP_DTYPE)
    recs["group"] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)   # random order, huge ids
    recs["peer_slot"] = rng.integers(0, 8, n)
    recs["index"] = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    recs["commit"] = recs["index"] - np.minimum(recs["index"], rng.integers(0, 4, n).astype(np.uint64))
    recs["flags"][rng.random(n) < 0.05] = B.REC_EXT   # stray EXT records: dropped
    hdr = check_round_trip(recs)
    assert not (hdr["flags"] & B.COMPACT_TILEABLE)


def test_descending_groups_escape():
    # groups below their block's g_base cannot be named by a header: ESC, still lossless
    n = 3000
    recs = np.zeros(n, dtype=B.APPEND_RESP_DTYPE)
    recs["group"] = np.arange(n, 0, -1)
    recs["index"] = 1000 + np.arange(n)
    recs["commit"] = recs["index"]
    hdr = check_round_trip(recs)
    assert hdr["n_side"] >= n - 4      # one header per block of units still works
    assert not (hdr["flags"] & B.COMPACT_TILEABLE)


def test_synth_round_is_compact():
    n = 20000
    s = B.Synth(n, 0xC0FFEE, k_peers=5)
    for _ in range(3):
        recs = s.next_round().copy()
        hdr = check_round_trip(recs)
        main = int(hdr["n_records"])
        # 2 header units per group + 1 per record; only REJECTs (2%) escape
        # 2 header units per group, 1 per record, 1 more for the hint of a REJECT; only a REJECT
        # that asks for a snapshot (0.25 % of the records) goes to the side table
        rej = np.nonzero(recs["flags"] & B.REC_REJECT)[0]
        want_snapshot = int(np.count_nonzero(recs["commit"][rej + 1]))
        assert hdr["n_units"] == (2 * n + main + (len(rej) - want_snapshot) + 3) & ~3      # padded to 16 bytes
        assert hdr["flags"] & B.COMPACT_TILEABLE
        assert hdr["n_side"] == 2 * want_snapshot
        assert hdr["total_bytes"] < 0.42 * 16 * (main + len(rej))    # vs the 16-byte packed form


def test_joint_round_is_compact():
    s = B.Synth(5000, 7, k_peers=5, joint=True)
    hdr = check_round_trip(s.next_round().copy())
    assert hdr["flags"] & B.COMPACT_TILEABLE and hdr["n_side"] < 0.01 * hdr["n_records"]


def test_capacity_errors():
    recs = B.Synth(100, 1, k_peers=5).next_round().copy()
    out = np.zeros(128, dtype=np.uint8)
    with pytest.raises(RuntimeError):
        B.pack_compact(recs, out)


def test_packer_differential_fuzz(tmp_path):
    """scripts/micro/pack_fuzz.cpp: pack_range (the vector form on this CPU) against pack_range_scalar on
    random traffic -- clean rounds, many REJECTs, hostile values, unsorted groups and stray EXT records --
    packed as a sequence of arbitrary [lo, hi) ranges on one PackState.  Units, g_base words, side
    records, ESC positions and the tileable / one-wave verdicts must be identical."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pack_fuzz")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "raft-rs_b200", "csrc"),
                    os.path.join(root, "scripts", "micro", "pack_fuzz.cpp"), os.path.join(root, "raft-rs_b200", "csrc", "pack_compact.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe, "800"], capture_output=True, text=True)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout + r.stderr
