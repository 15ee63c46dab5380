#!/bin/bash
# ncu --set full of the kernels that had no summary yet: MiMC path chain (config 2), witness chain, one-shot MSM tail + Horner
mkdir -p gpurun_out/c37
K="python scripts/bench_kernels.py --reps 1 --warmup 1"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_merkle_paths --launch-count 1 -f -o gpurun_out/c37/r2_ncu_merkle_paths $K > gpurun_out/c37/mp.log 2>&1; echo "mp rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tail_sums --launch-count 1 -f -o gpurun_out/c37/r2_ncu_tail_sums $K > gpurun_out/c37/ts.log 2>&1; echo "ts rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_horner --launch-count 1 -f -o gpurun_out/c37/r2_ncu_horner $K > gpurun_out/c37/ho.log 2>&1; echo "ho rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_withdraw_witness --launch-count 1 -f -o gpurun_out/c37/r2_ncu_witness python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --sharded-log-n 0 > gpurun_out/c37/wi.log 2>&1; echo "wi rc=$?"
ls -la gpurun_out/c37/*.ncu-rep
