#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest7.log
run() {
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/r2_ab_$tag.json 2> gpurun_out/r2_ab_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_$tag.json").read().strip().splitlines()[-1])
    print("$tag:", round(d["value"],1), "proofs/s", round(d["ms_per_step"],1), "ms", d["config"]["parity"]["bit_exact"], d["config"]["parity"]["verified"], "launches", d["gpu_launches"])
    print("   ", {k:round(v["ms"]/d["steps"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e:
    print("$tag: FAILED", e); print(open("gpurun_out/r2_ab_$tag.err").read()[-1500:])
PY
}
run red4 OG_RED_OCC=4
run red6 OG_RED_OCC=6
run red8 OG_RED_OCC=8
