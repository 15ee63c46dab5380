#!/bin/bash
mkdir -p gpurun_out/final3
python -m pytest tests -m gpu -q > gpurun_out/final3/gputest.log 2>&1; tail -2 gpurun_out/final3/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final3/bench.json 2> gpurun_out/final3/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final3/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['config'].get('parity'), d['clocks']['reasons'], d['gpu_launches'])
PY
