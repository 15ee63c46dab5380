#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest2.log
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
