#!/bin/bash
# one-shot MSM: reduction tail as independent tree sums (OG_MSM_TAIL=1, new default) vs the fan-8 levels (0)
mkdir -p gpurun_out/c26
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm or sharded" > gpurun_out/c26/gputest.log 2>&1; tail -2 gpurun_out/c26/gputest.log
OG_MSM_TAIL=1 python scripts/bench_kernels.py > gpurun_out/c26/kernels_tail1.jsonl 2> gpurun_out/c26/k1.err
OG_MSM_TAIL=0 python scripts/bench_kernels.py > gpurun_out/c26/kernels_tail0.jsonl 2> gpurun_out/c26/k0.err
for f in gpurun_out/c26/kernels_tail1.jsonl gpurun_out/c26/kernels_tail0.jsonl; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    if 'msm' in d.get('kernel',''): print(d['kernel'], round(d['ms'],3), {k:round(v,3) for k,v in d.get('kernels_ms',{}).items()} if 'kernels_ms' in d else '')
PY
done
