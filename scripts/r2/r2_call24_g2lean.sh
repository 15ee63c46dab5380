#!/bin/bash
# A/B: G2 bucket accumulation with every Fq2 value in shared memory (k_bucket_acc_sm2) vs the by-value kernel
mkdir -p gpurun_out/c24
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c24/gputest.log 2>&1; tail -2 gpurun_out/c24/gputest.log
run() { name=$1; shift; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c24/$name.json 2> gpurun_out/c24/$name.err; echo "$name rc=$?"; }
run lean OG_G2_LEAN=1
run old OG_G2_LEAN=0
for f in gpurun_out/c24/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()}
    print(sys.argv[1].split('/')[-1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity',{}).get('bit_exact'), k)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
