#!/bin/bash
OUT=gpurun_out/tile128; mkdir -p $OUT
for v in 1284:512 1283:512 1286:512 2562:512 1284:640; do
  var=${v%%:*}; cap=${v##*:}
  RAFTGPU_TILE_RECCAP=$cap RAFTGPU_TILE_VARIANT=$var RAFTGPU_TILE_DEBUG=1 timeout 300 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --e2e-steps 2 > $OUT/b$var.json 2> $OUT/b$var.err
  echo "variant $var cap $cap"; grep "tile debug" $OUT/b$var.err; python -c "
import json;d=json.loads(open('$OUT/b$var.json').read().strip().splitlines()[-1]);print(d['kernels'][0]['avg_us'], d['counters'])" || tail -5 $OUT/b$var.err
done
