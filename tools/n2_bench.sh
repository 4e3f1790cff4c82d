#!/bin/bash
# Development helper (2-GPU box): N=2 bench under a few settings; prints value, step time and kernel times.
tag=$1; shift
run() {
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NG:-2} --steps 10 --warmup 3 --no-e2e "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "Gb/s", round(d["ms_per_step"],3), "ms/step parity", d["parity"]["ok"], {k:round(v["ms_avg"],3) for k,v in d["kernels_ms"].items() if v["ms_avg"]>0.03})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/${tag}_$name.log").read()[-1500:])
PY
}
RG_P2P_SERIAL=1 run serial
run ctas96
run ctas148 --scatter-ctas 148
run ctas64 --scatter-ctas 64
