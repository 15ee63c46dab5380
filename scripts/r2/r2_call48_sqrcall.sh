#!/bin/bash
mkdir -p gpurun_out/c48
run() { name=$1; lib=$2; OWSHEN_B200_LIB=$lib python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c48/$name.json 2> gpurun_out/c48/$name.err; echo "$name rc=$?"; }
run base owshen_b200/libowshen_b200.so
run sc owshen_b200/libowshen_b200_sc.so
run base2 owshen_b200/libowshen_b200.so
run sc2 owshen_b200/libowshen_b200_sc.so
for f in gpurun_out/c48/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()}
    print(sys.argv[1].split('/')[-1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity',{}).get('bit_exact'), {x:k[x] for x in ('k_bucket_acc_g1','k_bucket_acc_g2')})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
