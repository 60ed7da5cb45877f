"""The packed 16-byte wire form (raftgpu_pack_records, csrc/arena.cu pack_record; layout in
kernels.cuh PackedRec) is lossless: decoding it the way the kernels do gives back every field of
the public 24-byte records, for arbitrary u64 values.  Also raftgpu_tile_index.  CPU only (host
functions of the library, no device call)."""
import ctypes as C

import numpy as np
import pytest

from helpers import B

PK_REJECT, PK_LOCAL, PK_EXT, PK_WIDE, PK_HAS_EXT = (1 << 35, 1 << 36, 1 << 37, 1 << 38, 1 << 39)
NO_COMMIT = 0xFFFFFF
M64 = (1 << 64) - 1


def pack(recs):
    out = np.zeros((4 * len(recs) + 4, 2), dtype=np.uint64)
    n = C.c_uint64()
    rc = B.lib().raftgpu_pack_records(recs.ctypes.data, len(recs), out.ctypes.data, out.shape[0], C.byref(n))
    assert rc == B.OK
    return out[: n.value]


def unpack(pk):
    """What load_rec<true> / load_reject_ext<true> compute, as (group, slot, flags, index, commit,
    hint, request_snapshot) per main record."""
    out, i, n = [], 0, len(pk)
    while i < n:
        w0, w1 = int(pk[i, 0]), int(pk[i, 1])
        assert not (w0 & PK_EXT), "an EXT payload without its record"
        delta = w0 >> 40
        flags = (1 if w0 & PK_REJECT else 0) | (2 if w0 & PK_LOCAL else 0)
        if w0 & PK_LOCAL:
            commit = 0 if delta == NO_COMMIT else (w1 + delta) & M64
        else:
            commit = (w1 - delta) & M64
        hint, rs, j = 0, 0, i + 1
        while j < n and j <= i + 3 and int(pk[j, 0]) & PK_EXT:
            kind = int(pk[j, 0]) >> 40
            if kind == 1:
                hint = int(pk[j, 1])
            elif kind == 2:
                rs = int(pk[j, 1])
            elif kind == 3 and (w0 & PK_WIDE):
                commit = int(pk[j, 1])
            j += 1
        out.append((w0 & 0xFFFFFFFF, (w0 >> 32) & 7, flags, w1, commit, hint, rs, bool(w0 & PK_HAS_EXT)))
        i = j
    return out


def test_round_trip_random_records():
    rng = np.random.default_rng(11)
    edge = [0, 1, 2, (1 << 24) - 2, (1 << 24) - 1, 1 << 24, (1 << 40) + 5, (1 << 63), M64 - 1, M64]
    recs, want = [], []
    for _ in range(4000):
        kind = rng.integers(0, 3)
        g, slot = int(rng.integers(0, 1 << 20)), int(rng.integers(0, 8))
        pick = lambda: edge[int(rng.integers(0, len(edge)))] if rng.random() < 0.4 else int(rng.integers(0, 1 << 50))
        index, commit = pick(), pick()
        if rng.random() < 0.5:                      # the common shape: commit a little below / above index
            commit = max(0, index - int(rng.integers(0, 8))) if kind != 2 else min(M64, index + int(rng.integers(0, 8)))
        if kind == 0:
            recs.append((g, slot, 0, 0, index, commit))
            want.append((g, slot, 0, index, commit, 0, 0, False))
        elif kind == 1:
            hint, rs = pick(), (0 if rng.random() < 0.7 else pick())
            recs.append((g, slot, B.REC_REJECT, 0, index, commit))
            recs.append((g, slot, B.REC_EXT, 0, hint, rs))
            want.append((g, slot, 1, index, commit, hint, rs, True))
        else:
            recs.append((g, slot, B.REC_LOCAL, 0, index, commit))
            want.append((g, slot, 2, index, commit, 0, 0, False))
    arr = np.zeros(len(recs), dtype=B.APPEND_RESP_DTYPE)
    for i, r in enumerate(recs):
        arr[i] = r
    got = unpack(pack(arr))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a == b, (a, b)


def test_common_records_take_one_packed_record():
    recs = np.zeros(3, dtype=B.APPEND_RESP_DTYPE)
    recs[0] = (7, 2, 0, 0, 1000, 997)                      # accept
    recs[1] = (7, 0, B.REC_LOCAL, 0, 1040, 1043)           # leader-local
    recs[2] = (8, 1, 0, 0, 5, 5)
    assert len(pack(recs)) == 3
    rej = np.zeros(2, dtype=B.APPEND_RESP_DTYPE)
    rej[0] = (9, 3, B.REC_REJECT, 0, 50, 40)
    rej[1] = (9, 3, B.REC_EXT, 0, 44, 0)
    assert len(pack(rej)) == 2                             # reject + hint; request_snapshot = 0 is implicit
    rej[1]["commit"] = 45
    assert len(pack(rej)) == 3


def test_out_of_range_slot_is_marked_unroutable():
    recs = np.zeros(1, dtype=B.APPEND_RESP_DTYPE)
    recs[0] = (3, 200, 0, 0, 10, 9)
    pk = pack(recs)
    assert int(pk[0, 0]) & 0xFFFFFFFF == 0xFFFFFFFF        # no such group: the kernels report NO_PROGRESS


def test_tile_index():
    rng = np.random.default_rng(3)
    n_groups = 2000
    groups = np.sort(rng.integers(0, n_groups, 5000)).astype(np.uint32)
    recs = np.zeros(len(groups), dtype=B.APPEND_RESP_DTYPE)
    recs["group"], recs["peer_slot"], recs["index"], recs["commit"] = groups, 1, 100, 99
    pk = pack(recs)
    off = B.tile_index(pk, len(pk), n_groups)
    n_tiles = (n_groups + 255) // 256
    assert len(off) == n_tiles + 1 and off[0] == 0 and off[-1] == len(pk)
    want = np.searchsorted(groups, np.arange(n_tiles + 1) * 256, side="left")
    assert np.array_equal(off, want)
    with pytest.raises(B.RaftGpuError):
        B.tile_index(np.ascontiguousarray(pk[::-1]), len(pk), n_groups)
    assert np.array_equal(B.tile_index(pk[:0], 0, n_groups), np.zeros(n_tiles + 1, dtype=np.uint32))
    # a record of a group the arena does not have (no tile would visit it): refused, not dropped
    recs2 = recs[-3:].copy()
    recs2["group"] = [n_groups - 1, n_groups, n_groups + 7]
    with pytest.raises(B.RaftGpuError) as e:
        B.tile_index(pack(recs2), 3, n_groups)
    assert e.value.status == B.ERR_RANGE
