"""A few batched NTT launches (the prover's shape) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import owshen_b200 as ob
from owshen_b200 import api
log_n, batch = 15, 768
ctx = ob.Context(0); L = api.lib()
data = torch.randint(0, 255, (32 * batch << log_n,), dtype=torch.uint8, device="cuda")
data.view(-1, 32)[:, 31] = 0
torch.cuda.synchronize()
for _ in range(2):
    assert L.og_ntt_dev(ctx._h, data.data_ptr(), log_n, batch, 0, 1) == 0
ctx.sync()
