#!/bin/bash
# first GPU contact: environment probe, smoke, parity tests, short bench
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv
nproc; lscpu | grep -E "Model name|Socket|Thread|^CPU\(s\)"; free -g | head -2
python - <<'PY'
import numpy as np, time
t=time.perf_counter(); a=np.zeros(200_000_000, dtype=np.uint8); a[::4096]=1; print("host first-touch 200MB: %.3fs" % (time.perf_counter()-t))
PY
} > gpurun_out/env.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench1.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench1.txt
tail -5 gpurun_out/env.txt gpurun_out/smoke.txt gpurun_out/pytest_gpu.txt gpurun_out/bench1.txt
