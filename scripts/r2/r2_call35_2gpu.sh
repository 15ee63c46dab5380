#!/bin/bash
# final build on 2 GPUs: the gpu tests that need two devices, the bench at N=2
mkdir -p gpurun_out/c35
nvidia-smi -L | wc -l
python -m pytest tests -m gpu -x -q > gpurun_out/c35/gputest.log 2>&1; tail -3 gpurun_out/c35/gputest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c35/bench_2gpu.json 2> gpurun_out/c35/bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c35/bench_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["parity"], {k:d["sharded_msm"][k] for k in ("ms","single_gpu_ms","matches_single_gpu","strong_scaling_vs_n1")})
PY
