#!/bin/bash
mkdir -p gpurun_out/c28
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm or sharded" > gpurun_out/c28/gputest.log 2>&1; tail -2 gpurun_out/c28/gputest.log
for c in "20 g1" "18 g2" "21 g2"; do echo "== $c"; python scripts/prof_msm.py $c | head -9; done > gpurun_out/c28/prof.log 2>&1
cat gpurun_out/c28/prof.log
python scripts/bench_kernels.py > gpurun_out/c28/kernels.jsonl 2> gpurun_out/c28/k.err
python - <<'PY'
import json
for l in open('gpurun_out/c28/kernels.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'msm' in d.get('kernel',''): print(d['kernel'], round(d['ms'],3))
PY
