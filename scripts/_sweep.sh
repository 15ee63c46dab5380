run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$tag.json"))
    k=d["kernels"]
    print("$tag", round(d["value"],1), "proofs/s", d["config"].get("proof0_verifies"), {n:round(v["ms"]/d["steps"],2) for n,v in k.items() if "acc" in n or "reduce" in n})
except Exception as e:
    print("$tag failed", e, open("gpurun_out/sw_$tag.err").read()[-400:])
PY
}
run default OG_X=0
run red_sm OG_RED_SM=1
OG_RED_SM=1 timeout 600 python -m pytest tests -m gpu -x -q -k "msm or prove or groth or golden" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
