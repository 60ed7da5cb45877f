#!/bin/bash
# compact stream: the new parity test, then a bench run
OUT=gpurun_out/compact; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact or zero_copy" > $OUT/pytest.txt 2>&1; tail -n 15 $OUT/pytest.txt
timeout 900 python bench.py --steps 60 --warmup 6 --no-cpu-baseline --e2e-steps 16 > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value=%.3e ms/step=%.4f"%(d["value"],d["ms_per_step"]), [(k["kernel"],round(k["avg_us"],1),round(k["frac"],3)) for k in d["kernels"]])
    for k in ("e2e","e2e_packed16","e2e_staged"):
        e=d[k]; print(k,"%.3e  %.3f ms/step  h2d %.1f MB"%(e["value"],e["ms_per_step"],e["h2d_bytes_per_step"]/1e6))
except Exception as e:
    print("failed", e); print(open("$OUT/bench.err").read()[-2000:])
PY
