#!/bin/bash
# final build (wide squarer in the G1 unit): gpu tests, smoke, default bench (both arms), kernel lines
mkdir -p gpurun_out/final2
python -m pytest tests -m gpu -x -q > gpurun_out/final2/gputest.log 2>&1; tail -2 gpurun_out/final2/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final2/bench.json 2> gpurun_out/final2/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final2/bench_ref.json 2> gpurun_out/final2/bench_ref.err; echo "ref rc=$?"
python scripts/bench_kernels.py > gpurun_out/final2/kernels.jsonl 2> gpurun_out/final2/kernels.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final2/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['config'].get('parity'), d['clocks'], d['roofline']['frac'], d['imad']['frac'])
print({a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()})
for l in open('gpurun_out/final2/kernels.jsonl'):
    try: k=json.loads(l)
    except Exception: continue
    if 'ms' in k: print(k['kernel'][:60], round(k['ms'],3))
PY
