#!/bin/bash
OUT=gpurun_out/exp; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" > $OUT/pytest.txt 2>&1; tail -n 3 $OUT/pytest.txt
run() {
  echo "== $*"; env "$@" RAFTGPU_TILE_DEBUG=1 timeout 300 python bench.py --steps 40 --warmup 4 --no-cpu-baseline --e2e-steps 2 $EXTRA > $OUT/b.json 2> $OUT/b.err
  grep "tile debug" $OUT/b.err; python - <<PY
import json
try:
    d=json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
    print("value=%.3e ms/step=%.4f"%(d["value"],d["ms_per_step"]), [(k["kernel"],round(k["avg_us"],1),round(k["frac"],3)) for k in d["kernels"]])
except Exception as e:
    print("failed", e); print(open("$OUT/b.err").read()[-1500:])
PY
}
run A=1
run RAFTGPU_TILE_STAGES=5
run RAFTGPU_TILE_VARIANT=2562
EXTRA="--workload cfg4" run A=1
