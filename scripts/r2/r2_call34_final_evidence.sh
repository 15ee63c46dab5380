#!/bin/bash
# final evidence of round 2 on the final build: default bench (both arms), kernel lines, ncu launch list of the bench command
mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo "ref rc=$?"
python scripts/bench_kernels.py > gpurun_out/final/kernels.jsonl 2> gpurun_out/final/kernels.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --sharded-log-n 0 > gpurun_out/final/launches.log 2>&1; echo "launches rc=$?"; wc -l gpurun_out/final/launches.csv
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['config'].get('parity'), d['clocks'], d['roofline']['frac'], d['imad']['frac'])
print({a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()})
print(open('gpurun_out/final/bench_ref.json').read()[:300])
PY
