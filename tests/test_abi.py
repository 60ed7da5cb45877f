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


def test_synth_library_exports_every_symbol_of_its_header():
    """include/raftgpu_synth.h (the bench's workload generator, its own library -- not part of the product .so)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "raftgpu_synth.h"), encoding="utf-8").read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(raftgpu_synth_[a-z0-9_]+)\s*\(", text)))
    assert len(declared) >= 3
    S = B.synth_lib()
    missing = [s for s in declared if not hasattr(S, s)]
    assert not missing, missing
    # ... and none of them leaks into the product library
    L = B.lib()
    for s in declared:
        with pytest.raises(AttributeError):
            getattr(L, s)


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
    # the documented hard limit: 8 peer slots per group (a distinct code, not a generic failure)
    out = C.c_void_p()
    assert L.raftgpu_arena_create(0, 1024, 16, 0, 0, C.byref(out)) == B.ERR_TOO_MANY_PEERS
    assert "peer slots" in B.strerror(B.ERR_TOO_MANY_PEERS)


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked tests")
    with pytest.raises(B.RaftGpuError) as e:
        B.Arena(1024)
    assert e.value.status == B.ERR_NO_DEVICE


def test_rust_stub_in_integration_md_is_generated_and_matches_the_header():
    """INTEGRATION.md section 2 is scripts/gen_rust_stub.py's output (a stale hand-written stub declared a 24-byte
    raftgpu_step_result against the header's 48 bytes in round 1), and every struct's C layout -- computed from the
    header text -- agrees with the ctypes mirror in binding.py."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    import gen_rust_stub as G
    doc = open(os.path.join(root, "INTEGRATION.md"), encoding="utf-8").read()
    block = doc[doc.index(G.BEGIN) + len(G.BEGIN):doc.index(G.END)]
    assert block.strip() == ("```rust\n" + G.generate() + "```").strip(), "run: python scripts/gen_rust_stub.py --update"
    lay = G.struct_layouts()
    mirror = {"raftgpu_progress": B.Progress, "raftgpu_group_state": B.GroupState, "raftgpu_info": B.Info,
              "raftgpu_counters": B.Counters, "raftgpu_step_result": B.StepResult, "raftgpu_wire_batch": B.WireBatch}
    for name, cls in mirror.items():
        size, offs = lay[name]
        assert C.sizeof(cls) == size, name
        # same field order and offsets (ctypes names may differ where the header's is a Python keyword / differs)
        assert [o for _, o, _ in offs] == [getattr(cls, f).offset for f, *_ in cls._fields_], name
    assert lay["raftgpu_append_resp"][0] == B.APPEND_RESP_DTYPE.itemsize == 24
    assert lay["raftgpu_send_entry"][0] == B.SEND_ENTRY_DTYPE.itemsize == 16
    assert lay["raftgpu_compact_hdr"][0] == B.COMPACT_HDR_DTYPE.itemsize == 64
    assert lay["raftgpu_step_result"][0] == 48
    # the Rust text itself: every struct carries its size assertion
    for name, (size, _) in lay.items():
        assert f"size_of::<{name}>() == {size}" in block, name


@pytest.mark.parametrize("header", ["raftgpu.h", "raftgpu_synth.h"])
def test_headers_are_plain_c(header, tmp_path):
    """The boundary is a C-ABI: include/*.h must compile as C99 (what a cgo / bindgen / ctypes user feeds it to) and
    as C++17, warnings as errors, with nothing but the standard headers."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text(f'#include "{header}"\nint main(void) {{ return 0; }}\n')
    inc = "-I" + os.path.join(root, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", inc, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)
