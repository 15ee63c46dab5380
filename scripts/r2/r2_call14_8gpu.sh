#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 scripts/bench_sharded_msm.py --log-n 24 --steps 3 --warmup 2 > gpurun_out/r2_sharded_8gpu.json 2> gpurun_out/r2_sharded_8gpu.err; echo "sharded rc=$?"
cut -c1-1200 gpurun_out/r2_sharded_8gpu.json; tail -2 gpurun_out/r2_sharded_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_8gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["parity"], d["sharded_msm"])
PY
tail -2 gpurun_out/r2_bench_8gpu.err
