#!/bin/bash
# re-sweep of the run-time knobs on the final build: level-0 reduction fan, window bits of the three MSMs
mkdir -p gpurun_out/c40
run() { name=$1; shift; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity --sharded-log-n 0 > gpurun_out/c40/$name.json 2> gpurun_out/c40/$name.err; echo "$name rc=$?"; }
run base X=1
run fan4 OG_RED_FAN0=4
run fan5 OG_RED_FAN0=5
run cA16 OG_C_A=16
run cA14 OG_C_A=14
run cB16 OG_C_B=16
run cB14 OG_C_B=14
run cC15 OG_C_C=15
for f in gpurun_out/c40/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()}
    print(sys.argv[1].split('/')[-1], round(d['value'],1), round(d['ms_per_step'],1), {x:k.get(x) for x in ('k_bucket_acc_g1','k_bucket_acc_g2','k_reduce_level_g1','k_reduce_level_g2','k_digits_scatter','k_digits_count_tiled')})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
