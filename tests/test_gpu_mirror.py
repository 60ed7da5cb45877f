"""The C++ mirror of the raft-rs tracker surface (raft-rs_b200/host/raftgpu.hpp): the
reference's own tests restated on it (host/test_mirror.cpp), every assertion through the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "raft-rs_b200", "host", "test_mirror")


@pytest.mark.gpu
def test_reference_tracker_tests_on_the_cpp_mirror():
    assert os.path.exists(BIN), "run __graft_entry__.build() first"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


def test_mirror_binary_is_built_and_needs_a_gpu():
    """Built by build(); without a device it refuses to run (no CPU fallback)."""
    import torch
    assert os.path.exists(BIN), "run __graft_entry__.build() first"
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no CUDA device" in r.stdout
