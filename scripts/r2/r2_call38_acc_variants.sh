#!/bin/bash
# G1 bucket accumulation: squarer (interleaved / wide) x resident CTAs per SM (8 / 6 / 5), whole-library variants via OWSHEN_B200_LIB
mkdir -p gpurun_out/c38
run() { name=$1; lib=$2; OWSHEN_B200_LIB=$lib python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c38/$name.json 2> gpurun_out/c38/$name.err; echo "$name rc=$?"; }
run i8 owshen_b200/libowshen_b200.so
for v in w8 w6 i6 w5; do run $v owshen_b200/libowshen_b200_$v.so; done
for f in gpurun_out/c38/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()}
    print(sys.argv[1].split('/')[-1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity',{}).get('bit_exact'), {x:k[x] for x in ('k_bucket_acc_g1','k_reduce_level_g1','k_assemble_g1')})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
