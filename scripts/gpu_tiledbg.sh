#!/bin/bash
OUT=gpurun_out/tiledbg; mkdir -p $OUT
for v in 2562 2563 2561; do
  RAFTGPU_TILE_VARIANT=$v RAFTGPU_TILE_DEBUG=1 timeout 300 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --e2e-steps 2 > $OUT/b$v.json 2> $OUT/b$v.err
  echo "variant $v"; grep "tile debug" $OUT/b$v.err; python -c "
import json;d=json.loads(open('$OUT/b$v.json').read().strip().splitlines()[-1]);print(d['kernels'][0]['avg_us'])"
done
