#!/bin/bash
mkdir -p gpurun_out/c29
python -m pytest tests -m gpu -x -q > gpurun_out/c29/gputest.log 2>&1; tail -2 gpurun_out/c29/gputest.log
for c in "20 g1" "18 g2"; do echo "== $c"; python scripts/prof_msm.py $c 2>&1 | head -9; done > gpurun_out/c29/prof.log 2>&1
cat gpurun_out/c29/prof.log
python scripts/bench_kernels.py > gpurun_out/c29/kernels.jsonl 2> gpurun_out/c29/k.err
python - <<'PY'
import json
for l in open('gpurun_out/c29/kernels.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'msm' in d.get('kernel','') or 'mimc' in d.get('kernel',''): print(d['kernel'], round(d['ms'],3))
PY
for tool in memcheck racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/c29/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/c29/sanitizer_$tool.log
done
