#!/bin/bash
# L2 prefetch of the next entry's table point in the bucket accumulation kernels (pf: G1 + G2, pf1: G1 only) vs shipped
mkdir -p gpurun_out/c46
run() { name=$1; lib=$2; OWSHEN_B200_LIB=$lib python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c46/$name.json 2> gpurun_out/c46/$name.err; echo "$name rc=$?"; }
run base owshen_b200/libowshen_b200.so
run pf owshen_b200/libowshen_b200_pf.so
run pf1 owshen_b200/libowshen_b200_pf1.so
for f in gpurun_out/c46/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()}
    print(sys.argv[1].split('/')[-1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity',{}).get('bit_exact'), {x:k[x] for x in ('k_bucket_acc_g1','k_bucket_acc_g2')})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
for lib in owshen_b200/libowshen_b200.so owshen_b200/libowshen_b200_pf.so; do OWSHEN_B200_LIB=$lib python scripts/bench_kernels.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'msm' in d.get('kernel',''): print('$lib'[-9:], d['kernel'], round(d['ms'],3))
"; done
