#!/bin/bash
# ncu evidence (run under gpurun, one GPU). Outputs into gpurun_out/.
#   tools/profile_r1.sh [tag] [kernel-regex]
TAG=${1:-r1}
KRE=${2:-partition_sorted|tile_insert|filter_hits|subtract}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/launches_$TAG.stdout 2>&1
# (2) the hot kernels, full set, source-level
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"$KRE" -s 8 -c 4 \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/prof_$TAG.stdout 2>&1
ls -la gpurun_out
