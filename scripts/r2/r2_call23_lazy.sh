#!/bin/bash
# lazy-reduction MiMC chain: gpu tests, kernel lines, one bench
mkdir -p gpurun_out/c23
python -m pytest tests -m gpu -x -q > gpurun_out/c23/gputest.log 2>&1; tail -2 gpurun_out/c23/gputest.log
python scripts/bench_kernels.py > gpurun_out/c23/kernels.jsonl 2> gpurun_out/c23/kernels.err; grep -o '"kernel": "mimc[^}]*' gpurun_out/c23/kernels.jsonl | cut -c1-200
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c23/bench.json 2> gpurun_out/c23/bench.err; cut -c1-200 gpurun_out/c23/bench.json
