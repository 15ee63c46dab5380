#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "affine" > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest4.log
run() {
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/r2_aff_$tag.json 2> gpurun_out/r2_aff_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_aff_$tag.json").read().strip().splitlines()[-1])
    print("$tag:", round(d["value"],1), "proofs/s", round(d["ms_per_step"],1), "ms", d["config"]["parity"]["bit_exact"], d["config"]["parity"]["verified"], "launches", d["gpu_launches"])
    print("   ", {k:round(v["ms"]/d["steps"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e:
    print("$tag: FAILED", e); print(open("gpurun_out/r2_aff_$tag.err").read()[-1500:])
PY
}
run v2_g1_occ5 OG_AFFINE=1 OG_AFF_OCC=5
run v2_g1_occ6 OG_AFFINE=1 OG_AFF_OCC=6
run v2_g1_occ8 OG_AFFINE=1 OG_AFF_OCC=8
run v2_g1g2 OG_AFFINE=3
