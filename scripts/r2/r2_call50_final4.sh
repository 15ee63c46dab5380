#!/bin/bash
mkdir -p gpurun_out/final4
python bench.py > gpurun_out/final4/bench.json 2> gpurun_out/final4/bench.err; echo "bench rc=$?"
python scripts/bench_kernels.py > gpurun_out/final4/kernels.jsonl 2> gpurun_out/final4/kernels.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final4/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['config'].get('parity'), d['clocks'], d['roofline']['frac'], d['imad']['frac'])
print({a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()})
PY
