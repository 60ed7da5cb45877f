"""Shared test glue: product binding + oracle side by side."""
import importlib

import numpy as np

from oracle import oracle as O

B = importlib.import_module("raft-rs_b200").binding

COMPARE_COLUMNS = ("matched", "next_idx", "peer_committed", "pending_snapshot",
                   "pending_request_snapshot", "commit_group_id", "pflags", "meta", "committed",
                   "term_start", "last_index")


def assert_columns_equal(got, want, n, what=""):
    for name in COMPARE_COLUMNS:
        a, b = getattr(got, name)[..., :n], getattr(want, name)[..., :n]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)[:5]
            raise AssertionError(f"{what}: column {name} differs at {bad.tolist()}: "
                                 f"got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}")


def bitmap_to_bool(bm, n):
    return np.unpackbits(np.ascontiguousarray(bm).view(np.uint8), bitorder="little")[:n].astype(bool)


def checksum(arr) -> int:
    """Order-sensitive 64-bit checksum (wrapping multiply-add) for full-size comparisons."""
    a = np.ascontiguousarray(arr).astype(np.uint64).ravel()
    w = (np.arange(a.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return int((a * w).sum(dtype=np.uint64))
