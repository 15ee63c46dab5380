#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest11.log
OG_NTT_TMA=0 timeout 600 python -m pytest tests -m gpu -x -q -k "ntt or groth16_prove_bit_exact or golden" > gpurun_out/r2_pytest11b.log 2>&1; echo "pytest(no tma) rc=$?"
tail -3 gpurun_out/r2_pytest11b.log
run() {
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sharded-log-n 0 > gpurun_out/r2_ab_$tag.json 2> gpurun_out/r2_ab_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_$tag.json").read().strip().splitlines()[-1])
    print("$tag:", round(d["value"],1), "proofs/s", round(d["ms_per_step"],1), "ms", d["config"]["parity"]["bit_exact"], d["config"]["parity"]["verified"], "launches", d["gpu_launches"])
    print("   ", {k:round(v["ms"]/d["steps"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e:
    print("$tag: FAILED", e); print(open("gpurun_out/r2_ab_$tag.err").read()[-1500:])
PY
}
run ntt_tma0 OG_NTT_TMA=0
run ntt_tma1 OG_NTT_TMA=1
run ntt_tma0b OG_NTT_TMA=0
run ntt_tma1b OG_NTT_TMA=1
for t in 0 1; do OG_NTT_TMA=$t timeout 300 python - <<'PY'
import os, json, torch, sys
sys.path.insert(0, ".")
import owshen_b200 as ob
from owshen_b200 import api
ctx = ob.Context(0); L = api.lib(); dev = torch.device("cuda", 0)
for log_n, batch in ((15, 3072), (20, 8), (24, 1), (12, 8192)):
    n = 1 << log_n
    data = torch.randint(0, 255, (32 * n * batch,), dtype=torch.uint8, device=dev); data.view(-1, 32)[:, 31] = 0
    torch.cuda.synchronize()
    for _ in range(3): L.og_ntt_dev(ctx._h, data.data_ptr(), log_n, batch, 0, 0)
    ctx.sync(); ctx.timer_start()
    for _ in range(5): L.og_ntt_dev(ctx._h, data.data_ptr(), log_n, batch, 0, 0)
    ms = ctx.timer_stop() / 5
    print(json.dumps({"OG_NTT_TMA": os.environ["OG_NTT_TMA"], "log_n": log_n, "batch": batch, "ms": ms}))
ctx.close()
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass2 --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_ncu_ntt_tma python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --sharded-log-n 0 > gpurun_out/r2_ncu_ntt_tma.log 2>&1; echo "ncu rc=$?"
