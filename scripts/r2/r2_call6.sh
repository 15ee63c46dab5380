#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest6.log
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
run base OG_AFFINE=0
run widesqr OWSHEN_B200_LIB=$PWD/owshen_b200/libowshen_b200_wide.so
run aff_coalesced OG_AFFINE=1
run aff_coalesced_wide OG_AFFINE=1 OWSHEN_B200_LIB=$PWD/owshen_b200/libowshen_b200_wide.so
run lanes2_noprio OG_LANES=2 OG_CHUNK=512 OG_LANE_PRIO=0
timeout 600 python scripts/bench_kernels.py --reps 5 > gpurun_out/r2_kernels.jsonl 2> gpurun_out/r2_kernels.err; echo "kernels rc=$?"; cut -c1-260 gpurun_out/r2_kernels.jsonl
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --sharded-log-n 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass2 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_ncu_ntt $B > gpurun_out/r2_ncu_ntt.log 2>&1; echo "ncu ntt rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_bucket_acc_sm$' --launch-count 1 -f -o gpurun_out/r2_ncu_xyzz_g2 $B > gpurun_out/r2_ncu_g2.log 2>&1; echo "ncu g2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_digits --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2_ncu_digits $B > gpurun_out/r2_ncu_digits.log 2>&1; echo "ncu digits rc=$?"
ls -la gpurun_out/*.ncu-rep
