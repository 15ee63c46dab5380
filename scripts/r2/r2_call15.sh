#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest15.log
timeout 900 python scripts/bench_kernels.py --reps 5 > gpurun_out/r2_kernels3.jsonl 2> gpurun_out/r2_kernels3.err; echo "kernels rc=$?"; grep -E "config 3|config 2" gpurun_out/r2_kernels3.jsonl | cut -c1-700
