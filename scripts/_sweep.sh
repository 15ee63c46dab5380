for f in 3 4 14 15 16; do
  OG_ACC_OCC_G2=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/occ$f.json 2> gpurun_out/occ$f.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/occ$f.json"))
    k=d["kernels"]
    print("g2 occ=$f", round(d["value"],1), "proofs/s", d["config"].get("proof0_verifies"), {n:round(v["ms"]/d["steps"],2) for n,v in k.items() if "acc" in n})
except Exception as e:
    print("occ=$f failed", e, open("gpurun_out/occ$f.err").read()[-400:])
PY
done
OG_ACC_OCC_G2=15 timeout 600 python -m pytest tests -m gpu -x -q -k "msm or prove or groth or golden" 2>&1 | tail -2
