#!/bin/bash
set -u
mkdir -p gpurun_out
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
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest9.log
run reorder OG_AFFINE=0
timeout 900 python scripts/bench_kernels.py --reps 5 > gpurun_out/r2_kernels2.jsonl 2> gpurun_out/r2_kernels2.err; echo "kernels rc=$?"; cut -c1-420 gpurun_out/r2_kernels2.jsonl; tail -3 gpurun_out/r2_kernels2.err
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --sharded-log-n 0"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_bucket_acc_sm --csv --log-file gpurun_out/r2_traffic_bucket_acc.csv $B > gpurun_out/r2_traffic.log 2>&1; echo "traffic rc=$?"; cat gpurun_out/r2_traffic_bucket_acc.csv | tail -12
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --sharded-log-n 0 > gpurun_out/r2_launches.log 2>&1; echo "launches rc=$?"; wc -l gpurun_out/r2_launches.csv
