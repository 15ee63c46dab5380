#!/bin/bash
mkdir -p gpurun_out/c32
python -m pytest tests -m gpu -x -q > gpurun_out/c32/gputest.log 2>&1; tail -3 gpurun_out/c32/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in "20 g1" "22 g1"; do echo "== $c"; python scripts/prof_msm.py $c 2>&1 | head -13; done > gpurun_out/c32/prof.log 2>&1
cat gpurun_out/c32/prof.log
python scripts/bench_kernels.py > gpurun_out/c32/kernels.jsonl 2> gpurun_out/c32/k.err; python - gpurun_out/c32/kernels.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    if 'msm' in d.get('kernel',''): print(d['kernel'], round(d['ms'],3))
PY
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/c32/sanitizer_racecheck.log 2>&1; tail -2 gpurun_out/c32/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/c32/sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/c32/sanitizer_memcheck.log
