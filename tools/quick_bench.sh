#!/bin/bash
# Development helper (GPU box): parity tests once, then the device-resident bench for the default library and
# every variant under rufus_b200/variants/, kernel times only.
tag=$1; shift
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=90 2>&1 | tail -2
for lib in default "$@"; do
  if [ $lib = default ]; then unset RG_LIB; else export RG_LIB=rufus_b200/variants/$lib.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_$lib.json 2> gpurun_out/${tag}_$lib.log || tail -3 gpurun_out/${tag}_$lib.log
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_$lib.json"))
print("$lib", round(d["value"],1), "Gb/s", round(d["ms_per_step"],3), "ms/step frac", round(d["roofline"]["frac"],3), {k:round(v["ms_avg"],3) for k,v in d["kernels_ms"].items() if v["ms_avg"]>0.03}, d["result"])
PY
done
