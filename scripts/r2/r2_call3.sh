#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "affine or msm_edge or golden_proof" > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest3.log
for mode in 0 1 3; do
  OG_AFFINE=$mode timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/r2_aff_$mode.json 2> gpurun_out/r2_aff_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_aff_$mode.json").read().strip().splitlines()[-1])
    print("affine $mode:", round(d["value"],1), "proofs/s", round(d["ms_per_step"],1), "ms", d["config"]["parity"], "launches", d["gpu_launches"])
    print("   ", {k:round(v["ms"]/d["steps"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e:
    print("affine $mode: FAILED", e); print(open("gpurun_out/r2_aff_$mode.err").read()[-1500:])
PY
done
