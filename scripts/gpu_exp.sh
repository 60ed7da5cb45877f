#!/bin/bash
OUT=gpurun_out/exp; mkdir -p $OUT
run() {
  echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --e2e-steps 24 --packed16 > $OUT/b.json 2> $OUT/b.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
    for k in ("e2e","e2e_packed16"):
        e=d[k]; print(k,"%.3e  %.3f ms/step  h2d %.1f MB"%(e["value"],e["ms_per_step"],e["h2d_bytes_per_step"]/1e6))
except Exception as e:
    print("failed", e); print(open("$OUT/b.err").read()[-1500:])
PY
}
run RAFTGPU_COMPACT_SCATTER=1
run RAFTGPU_CTILE_GROUPS=3
run RAFTGPU_CTILE_GROUPS=2
run RAFTGPU_CTILE_GROUPS=3 RAFTGPU_COMPACT_ORDERED=1
