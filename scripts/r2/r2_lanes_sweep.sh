#!/bin/bash
# round 2: GPU parity tests, then proofs/s of the lane-pipelined prover for (lanes, chunk) combinations
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest.log
for cfg in "1 1024" "1 256" "2 512" "2 256" "2 128" "2 64"; do
  set -- $cfg
  OG_LANES=$1 OG_CHUNK=$2 timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2_lanes_$1_$2.json 2> gpurun_out/r2_lanes_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_lanes_$1_$2.json").read().strip().splitlines()[-1])
    print("lanes $1 chunk $2:", round(d["value"],1), "proofs/s", round(d["ms_per_step"],1), "ms", "e2e", round(d["e2e"]["value"],1), d["config"]["proof0_verifies"], d["config"]["e2e_bytes_equal_device_path"])
    print("   ", {k:v["ms"] for k,v in list(d["kernels"].items())[:8]})
except Exception as e:
    print("lanes $1 chunk $2: FAILED", e); print(open("gpurun_out/r2_lanes_$1_$2.err").read()[-1500:])
PY
done
