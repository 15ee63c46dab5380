#!/bin/bash
# ncu --set full of the dominant kernel on the final build (wide squarer in the G1 unit): the C' MSM launch (second G1 launch of a step)
mkdir -p gpurun_out/c47
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --sharded-log-n 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc_sm1 --launch-skip 1 --launch-count 1 -f -o gpurun_out/c47/r2_ncu_xyzz_g1_final $B > gpurun_out/c47/ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/c47/*.ncu-rep
