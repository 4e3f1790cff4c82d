#!/bin/bash
# Development helper (N-GPU box): bench at NG GPUs, overlapped (default) and serial (isolated kernel timings).
tag=$1; NG=${NG:-8}
run() {
  name=$1; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 10 --warmup 3 "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$name.json").read().strip().splitlines()[-1])
    print("$name N=$NG", round(d["value"],1), "Gb/s", round(d["ms_per_step"],3), "ms/step parity", d["parity"]["ok"], "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("parity_ok"), {k:round(v["ms_avg"],3) for k,v in d["kernels_ms"].items() if v["ms_avg"]>0.03})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/${tag}_$name.log").read()[-2500:])
PY
}
run overlap
[ -n "$BATCHED" ] && RG_P2P_SCHEDULE=batched run batched --no-e2e; true
