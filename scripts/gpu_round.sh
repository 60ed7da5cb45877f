#!/bin/bash
# One GPU round: parity tests, smoke, bench, ncu launch list + full captures of both hot kernels.
# usage: scripts/gpu_round.sh <tag> [skip_tests]
TAG=${1:-r}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 300 python __graft_entry__.py smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" >> $OUT/smoke.txt
  timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt
fi
timeout 900 python bench.py --steps 60 --warmup 6 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
PROF="python bench.py --profile --steps 8 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recompute_kernel -s 6 -c 2 -f -o $OUT/prof_recompute $PROF > $OUT/ncu_recompute.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_kernel -s 6 -c 2 -f -o $OUT/prof_apply $PROF > $OUT/ncu_apply.log 2>&1
for f in smoke.txt pytest_gpu.txt bench.err; do [ -f $OUT/$f ] && { echo "== $f"; tail -n 4 $OUT/$f; }; done
echo "== bench"; cat $OUT/bench.json | head -c 3000
