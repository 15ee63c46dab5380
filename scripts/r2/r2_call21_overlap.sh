#!/bin/bash
# A/B: digit sorts of the 2nd/3rd MSM on the side stream (persistent grids) under the accumulation of the previous MSM
mkdir -p gpurun_out/c21
run() { name=$1; shift; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c21/$name.json 2> gpurun_out/c21/$name.err; echo "$name rc=$?"; }
run base OG_OVERLAP=0
run ov_c2 OG_OVERLAP=1 OG_SIDE_CTAS=2
run ov_c1 OG_OVERLAP=1 OG_SIDE_CTAS=1
run ov_c4 OG_OVERLAP=1 OG_SIDE_CTAS=4
run ov_c2_cnt OG_OVERLAP=1 OG_SIDE_CTAS=2 OG_SIDE_COUNT=1
run ov_c4_cnt OG_OVERLAP=1 OG_SIDE_CTAS=4 OG_SIDE_COUNT=1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "groth16 or prove or withdraw" > gpurun_out/c21/tests.log 2>&1; tail -2 gpurun_out/c21/tests.log
for f in gpurun_out/c21/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()}
    print(sys.argv[1].split('/')[-1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity',{}).get('bit_exact'), k)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
