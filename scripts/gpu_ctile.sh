#!/bin/bash
# fused compact kernel: parity tests, then bench variants
OUT=gpurun_out/ctile; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact" > $OUT/pytest.txt 2>&1; tail -n 25 $OUT/pytest.txt
for v in 2 3; do
  RAFTGPU_CTILE_GROUPS=$v RAFTGPU_TILE_DEBUG=1 timeout 300 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --e2e-steps 4 > $OUT/b$v.json 2> $OUT/b$v.err
  echo "consumer groups $v"; grep "tile debug" $OUT/b$v.err; python - <<PY
import json
try:
    d=json.loads(open("$OUT/b$v.json").read().strip().splitlines()[-1])
    print("value=%.3e ms/step=%.4f"%(d["value"],d["ms_per_step"]), [(k["kernel"],round(k["avg_us"],1),round(k["frac"],3)) for k in d["kernels"]], d["counters"])
    for k in ("e2e","e2e_packed16","e2e_staged"):
        e=d[k]; print(k,"%.3e  %.3f ms/step  h2d %.1f MB"%(e["value"],e["ms_per_step"],e["h2d_bytes_per_step"]/1e6))
except Exception as e:
    print("failed", e); print(open("$OUT/b$v.err").read()[-1500:])
PY
done
