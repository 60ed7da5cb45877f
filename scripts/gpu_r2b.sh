#!/bin/bash
# e2e staging diagnosis: step_begin_records traced at several staging-thread counts
OUT=gpurun_out/${1:-r2b}; mkdir -p $OUT
for d in /sys/bus/pci/devices/*; do [ -f $d/local_cpulist ] && grep -qi "0x030[02]00" $d/class 2>/dev/null && echo "$d $(cat $d/local_cpulist) numa=$(cat $d/numa_node)"; done > $OUT/topo.txt
cat /sys/devices/system/cpu/cpu0/topology/thread_siblings_list /sys/devices/system/cpu/cpu1/topology/thread_siblings_list >> $OUT/topo.txt
numactl -H >> $OUT/topo.txt 2>&1
cat $OUT/topo.txt | head -20
for T in 32; do
  RAFTGPU_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 16 --e2e-threads $T --no-sublegs --no-cpu-baseline > $OUT/bench_T$T.json 2> $OUT/bench_T$T.err
  echo "== T=$T"; grep step_begin_records $OUT/bench_T$T.err | tail -n 3
  python - <<PY
import json
d=json.loads(open("$OUT/bench_T$T.json").read().strip().splitlines()[-1])
print({a:(round(b,4) if isinstance(b,float) else b) for a,b in d["e2e"].items() if a!="api"})
PY
done
