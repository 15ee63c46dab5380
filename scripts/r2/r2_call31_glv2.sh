#!/bin/bash
mkdir -p gpurun_out/c31
python -m pytest tests -m gpu -x -q > gpurun_out/c31/gputest.log 2>&1; tail -3 gpurun_out/c31/gputest.log
for g in 1 0; do echo "== OG_GLV=$g"; OG_GLV=$g python scripts/prof_msm.py 20 g1 2>&1 | head -12; done > gpurun_out/c31/prof.log 2>&1
cat gpurun_out/c31/prof.log
python scripts/bench_kernels.py > gpurun_out/c31/kernels.jsonl 2> gpurun_out/c31/k.err; python - gpurun_out/c31/kernels.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    if 'msm' in d.get('kernel',''): print(d['kernel'], round(d['ms'],3))
PY
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/c31/bench.json 2> gpurun_out/c31/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c31/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity'), {k:v for k,v in d['sharded_msm'].items() if k in ('ms','single_gpu_ms','matches_single_gpu')})
PY
