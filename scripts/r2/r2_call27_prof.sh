#!/bin/bash
mkdir -p gpurun_out/c27
for t in 1 0; do for c in "20 g1" "18 g2" "21 g1" "21 g2"; do echo "== OG_MSM_TAIL=$t $c"; OG_MSM_TAIL=$t python scripts/prof_msm.py $c; done; done > gpurun_out/c27/prof.log 2>&1
cat gpurun_out/c27/prof.log
