#!/bin/bash
mkdir -p gpurun_out/c30
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm or sharded" > gpurun_out/c30/gputest.log 2>&1; tail -3 gpurun_out/c30/gputest.log
for g in 1 0; do echo "== OG_GLV=$g"; OG_GLV=$g python scripts/prof_msm.py 20 g1 2>&1 | head -12; done > gpurun_out/c30/prof.log 2>&1
cat gpurun_out/c30/prof.log
for g in 1 0; do OG_GLV=$g python scripts/bench_kernels.py > gpurun_out/c30/kernels_glv$g.jsonl 2> gpurun_out/c30/k$g.err; python - gpurun_out/c30/kernels_glv$g.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    if 'msm_g1' in d.get('kernel',''): print(sys.argv[1][-12:], d['kernel'], round(d['ms'],3))
PY
done
