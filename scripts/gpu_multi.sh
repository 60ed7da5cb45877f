#!/bin/bash
# multi-GPU bench under torchrun, as the driver launches it
N=${1:-2}; OUT=gpurun_out/multi; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 60 --warmup 6 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "rc=$?"; tail -c 1500 $OUT/bench_n$N.err | grep -v Warning | tail -5
python - <<PY
import json
d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value=%.3e ms/step=%.4f e2e=%.3e (%.3f ms, %d thr) zc=%.3e"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e_staged"]["value"],d["e2e_staged"]["ms_per_step"],d["e2e_staged"]["host_threads"],d["e2e"]["value"]), d["counters"], d["clocks"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --impl reference --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-400
