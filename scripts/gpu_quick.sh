#!/bin/bash
# quick GPU iteration: parity tests (optional) + bench
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt
  tail -n 5 $OUT/pytest_gpu.txt
fi
timeout 900 python bench.py --steps 60 --warmup 6 --no-cpu-baseline --e2e-steps 16 > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value=%.3e ms/step=%.4f"%(d["value"],d["ms_per_step"]), [(k["kernel"],round(k["avg_us"],1),round(k["frac"],3)) for k in d["kernels"]])
    print("e2e (zero-copy)=%.3e  %.3f ms/step | staged=%.3e %.3f ms/step"%(d["e2e"]["value"],d["e2e"]["ms_per_step"],d["e2e_staged"]["value"],d["e2e_staged"]["ms_per_step"]), d["e2e_staged"].get("host_ms_per_step"))
except Exception as e:
    print("failed", e); print(open("$OUT/bench.err").read()[-2000:])
PY
