"""Per-kernel CUDA-event breakdown of a one-shot MSM (og_profile)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import owshen_b200 as ob
from owshen_b200 import api
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
curve = sys.argv[2] if len(sys.argv) > 2 else "g1"
n = 1 << log_n
rng = random.Random(1)
ctx = ob.Context(0)
L = api.lib()
raw = rng.randbytes(31 * n); sc = b"".join(raw[31*i:31*i+31] + b"\0" for i in range(n))
raw = rng.randbytes(31 * n); k = b"".join(raw[31*i:31*i+31] + b"\0" for i in range(n))
pts = ctx.g1_generator_mul(k) if curve == "g1" else ctx.g2_generator_mul(k)
dev = torch.device("cuda", 0)
d_pts = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
d_sc = torch.frombuffer(bytearray(sc), dtype=torch.uint8).to(dev)
d_out = torch.empty(128, dtype=torch.uint8, device=dev)
fn = L.og_msm_g1_dev if curve == "g1" else L.og_msm_g2_dev
torch.cuda.synchronize()
for _ in range(2):
    assert fn(ctx._h, d_pts.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr()) == 0
ctx.sync()
ctx.profile(True)
ctx.timer_start()
assert fn(ctx._h, d_pts.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr()) == 0
ms = ctx.timer_stop()
ctx.profile(False)
print(f"msm {curve} 2^{log_n}: {ms:.3f} ms")
for k_, (cnt, t) in sorted(ctx.profile_dump().items(), key=lambda kv: -kv[1][1]):
    print(f"  {k_:28s} x{cnt:3d} {t:9.3f} ms")
