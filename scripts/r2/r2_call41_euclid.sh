#!/bin/bash
# binary extended Euclid instead of Fermat in Fp::inv(): gpu tests, smoke, kernel lines, bench, one-shot breakdown
mkdir -p gpurun_out/c41
python -m pytest tests -m gpu -x -q > gpurun_out/c41/gputest.log 2>&1; tail -2 gpurun_out/c41/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python scripts/prof_msm.py 20 g1 2>&1 | head -6
python scripts/prof_msm.py 18 g2 2>&1 | head -5
python scripts/bench_kernels.py > gpurun_out/c41/kernels.jsonl 2> gpurun_out/c41/kernels.err
python bench.py > gpurun_out/c41/bench.json 2> gpurun_out/c41/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c41/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['config'].get('parity'), d['clocks']['reasons'])
print({a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()})
for l in open('gpurun_out/c41/kernels.jsonl'):
    try: k=json.loads(l)
    except Exception: continue
    if 'ms' in k: print(k['kernel'][:60], round(k['ms'],3))
PY
