#!/bin/bash
# A/B 2: which accumulation the side-stream scatters run under (order of the three MSMs), CTAs per SM of the persistent scatter
mkdir -p gpurun_out/c22
run() { name=$1; shift; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/c22/$name.json 2> gpurun_out/c22/$name.err; echo "$name rc=$?"; }
run acb_c2 OG_ORDER=ACB OG_SIDE_CTAS=2
run acb_c4 OG_ORDER=ACB OG_SIDE_CTAS=4
run acb_c6 OG_ORDER=ACB OG_SIDE_CTAS=6
run cab_c2 OG_ORDER=CAB OG_SIDE_CTAS=2
run cab_c4 OG_ORDER=CAB OG_SIDE_CTAS=4
run cba_c2 OG_ORDER=CBA OG_SIDE_CTAS=2
for f in gpurun_out/c22/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={a:round(b['ms']/d['steps'],1) for a,b in d['kernels'].items()}
    print(sys.argv[1].split('/')[-1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('parity',{}).get('bit_exact'), k)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
