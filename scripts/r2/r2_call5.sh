#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --sharded-log-n 0"
OG_AFFINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_aff_round --launch-skip 72 --launch-count 1 -f -o gpurun_out/r2_ncu_aff_round $B > gpurun_out/r2_ncu_aff.log 2>&1; echo "ncu aff rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc_sm1 --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2_ncu_xyzz_g1 $B > gpurun_out/r2_ncu_xyzz.log 2>&1; echo "ncu xyzz rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_reduce_level --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2_ncu_reduce_g1 $B > gpurun_out/r2_ncu_reduce.log 2>&1; echo "ncu reduce rc=$?"
ls -la gpurun_out/*.ncu-rep
