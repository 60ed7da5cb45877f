#!/bin/bash
run() { echo -n "$1: "; env $2 python bench.py --profile --steps 30 --warmup 5 $3 2> /tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value=%.3e ms/step=%.4f'%(d['value'],d['ms_per_step']), [(k['kernel'],round(k['avg_us'],1),round(k['frac'],3)) for k in d['kernels']])"; grep "tile debug" /tmp/err.txt; }
run "256x3 cap1024 (4st)      " "RAFTGPU_TILE_VARIANT=2563 RAFTGPU_TILE_RECCAP=1024" ""
run "256x3 cap1024 (4st) debug" "RAFTGPU_TILE_VARIANT=2563 RAFTGPU_TILE_RECCAP=1024 RAFTGPU_TILE_DEBUG=1" ""
run "256x2 cap1536            " "RAFTGPU_TILE_VARIANT=2562" ""
run "256x2 cap1024 (4st)      " "RAFTGPU_TILE_VARIANT=2562 RAFTGPU_TILE_RECCAP=1024" ""
run "256x1 cap1536 debug      " "RAFTGPU_TILE_VARIANT=2561 RAFTGPU_TILE_DEBUG=1" ""
run "scatter                  " "X=1" "--scatter"
RAFTGPU_TILE_VARIANT=2563 RAFTGPU_TILE_RECCAP=1024 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
