#!/bin/bash
# ncu evidence for round 1 (run under gpurun, one GPU). Outputs into gpurun_out/.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/launches_r1.stdout 2>&1
# (2) the hot kernels, full set, source-level
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'partition_sorted|tile_insert|filter_hits|subtract' -s 8 -c 4 \
    -o gpurun_out/prof_r1 -f $CMD > gpurun_out/prof_r1.stdout 2>&1
ls -la gpurun_out
