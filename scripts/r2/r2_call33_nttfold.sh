#!/bin/bash
mkdir -p gpurun_out/c33
python -m pytest tests -m gpu -x -q > gpurun_out/c33/gputest.log 2>&1; tail -3 gpurun_out/c33/gputest.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c33/bench.json 2> gpurun_out/c33/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c33/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity'), {a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()})
PY
