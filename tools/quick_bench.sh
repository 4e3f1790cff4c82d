#!/bin/bash
# GPU tests + a short resident-only bench; prints the per-kernel times. usage: tools/quick_bench.sh [tag]
TAG=${1:-q}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_$TAG.log > gpurun_out/bench_$TAG.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", round(d["value"],2), "Gbases/s  ms/step", round(d["ms_per_step"],3), " roofline frac", round(d["roofline"]["frac"],4))
print({k:round(v["ms_avg"],4) for k,v in d["kernels_ms"].items()})
PY
