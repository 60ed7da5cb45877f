#!/bin/bash
OUT=gpurun_out/exp; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_begin_records" > $OUT/pytest.txt 2>&1; tail -n 5 $OUT/pytest.txt
run() {
  echo "== $*"; env "$@" RAFTGPU_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --e2e-steps 16 > $OUT/b.json 2> $OUT/b.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
    for k in ("e2e","e2e_staged","e2e_records_api"):
        e=d[k]; print(k,"%.3e  %.3f ms/step  h2d %.1f MB"%(e["value"],e["ms_per_step"],e["h2d_bytes_per_step"]/1e6), e.get("host_ms_per_step"))
except Exception as e:
    print("failed", e); print(open("$OUT/b.err").read()[-1500:])
PY
}
run RAFTGPU_HOST_THREADS=16
run RAFTGPU_HOST_THREADS=32
