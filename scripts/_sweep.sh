for f in 3 4 2; do
  OG_ACC_OCC_G2=$f timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/occ$f.json 2> gpurun_out/occ$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/occ$f.json"))
k=d["kernels"]
print("g2 minb=$f", round(d["value"],1), "proofs/s", {n:round(v["ms"]/d["steps"],2) for n,v in k.items() if "acc" in n})
PY
done
OG_ACC_OCC_G2=4 timeout 600 python -m pytest tests -m gpu -x -q -k "msm or prove or groth or golden" 2>&1 | tail -2
