#!/bin/bash
# e2e-focused run: staging thread sweep with the library trace on
for t in 8 16 32; do
  RAFTGPU_TRACE=1 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 16 --e2e-threads $t 2> gpurun_out/e2e_$t.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('T=$t e2e=%.3e ms/step=%.3f'%(e['value'],e['ms_per_step']), e['host_ms_per_step'], 'h2d=%.1fMB'%(e['h2d_bytes_per_step']/1e6))"
  grep "enqueue_bulk" gpurun_out/e2e_$t.err | tail -3
done
