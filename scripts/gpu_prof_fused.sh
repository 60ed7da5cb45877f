#!/bin/bash
OUT=gpurun_out/fused; mkdir -p $OUT
PROF="python bench.py --profile --steps 6 --warmup 3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_tile_kernel -s 4 -c 1 -f -o $OUT/prof_fused $PROF > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
