#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus or two_contexts or sharded" > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest10.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 2500 gpurun_out/r2_bench_2gpu.json; tail -3 gpurun_out/r2_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29618 scripts/bench_sharded_msm.py --log-n 24 --steps 2 --warmup 1 > gpurun_out/r2_sharded_2gpu.json 2> gpurun_out/r2_sharded_2gpu.err; echo "sharded rc=$?"
cat gpurun_out/r2_sharded_2gpu.json | cut -c1-1500; tail -3 gpurun_out/r2_sharded_2gpu.err
