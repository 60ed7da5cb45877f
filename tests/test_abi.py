"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every
symbol include/raftgpu.h declares, and fails loudly (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest

from helpers import B


def test_library_exports_every_declared_symbol():
    L = B.lib()
    declared = B.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # and the binding knows a signature for each of them
    assert sorted(L._signatures) == declared


def test_abi_version_and_strerror():
    L = B.lib()
    assert L.raftgpu_abi_version() == 1
    assert B.strerror(0) == "ok"
    assert "no CPU fallback" in B.strerror(B.ERR_NO_DEVICE)
    assert "StepPeerNotFound" in B.strerror(B.ERR_PEER_NOT_FOUND)


def test_struct_layouts_match_header():
    assert C.sizeof(B.Progress) == 56
    assert C.sizeof(B.GroupState) == 32
    assert C.sizeof(B.Counters) == 64
    assert C.sizeof(B.StepResult) == 48
    assert B.APPEND_RESP_DTYPE.itemsize == 24


def test_null_arena_is_rejected_not_crashing():
    L = B.lib()
    assert L.raftgpu_arena_destroy(None) == B.ERR_INVALID
    assert L.raftgpu_group_alloc(None, None) == B.ERR_INVALID
    assert L.raftgpu_step(None, 0, None) == B.ERR_INVALID
    assert L.raftgpu_recompute(None, None, 0, 1, None, None, None, None) == B.ERR_INVALID
    assert L.raftgpu_arena_create(0, 0, 8, 0, 0, None) == B.ERR_INVALID


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked tests")
    with pytest.raises(B.RaftGpuError) as e:
        B.Arena(1024)
    assert e.value.status == B.ERR_NO_DEVICE
