#!/usr/bin/env python
"""BASELINE config 5: one G1+G2 MSM of 2^LOG_N points sharded by point range over the GPUs of one box.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/bench_sharded_msm.py --log-n 24 [--steps 3 --warmup 2]

Every rank synthesises ITS slice of the inputs on its own GPU (points = k_i * G via og_g*_generator_mul
with seeded k_i, 254-bit seeded scalars shared by the G1 and G2 MSM), runs a full Pippenger on the
slice (og_msm_g1_dev / og_msm_g2_dev), then the partial sums (64 B + 128 B per rank) are all-gathered
over NCCL and added on every rank (og_g1_sum / og_g2_sum).  ncclSum cannot add curve points, so the
"reduce" is all-gather + local group addition; the exchange is < 2 KB, once.
Timing: barrier + sync, CUDA events on the library stream per rank, max over ranks; one JSON line.
Check: all ranks agree, and for log_n <= --verify-max rank 0 recomputes the whole MSM alone.
"""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import owshen_b200 as ob  # noqa: E402
from owshen_b200 import api  # noqa: E402
from owshen_b200.sharded import shard_range  # noqa: E402


def fr_bytes_254(seed, lo, hi):
    """Scalars i in [lo, hi): 253-bit seeded values (always canonical), reproducible per index block."""
    out = bytearray()
    blk = 1 << 16
    for b0 in range(lo - lo % blk, hi, blk):
        rng = random.Random(seed * 1000003 + b0 // blk)
        raw = rng.randbytes(32 * blk)
        a, b = max(lo, b0) - b0, min(hi, b0 + blk) - b0
        chunk = bytearray(raw[32 * a:32 * b])
        chunk[31::32] = bytes(x & 0x1f for x in chunk[31::32])        # < 2^253 < r
        out += chunk
    return bytes(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--verify-max", type=int, default=22)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    ctx = ob.Context(local)
    L = api.lib()
    n = 1 << args.log_n
    lo, hi = shard_range(n, rank, world)
    m = hi - lo
    t0 = time.time()
    ks = fr_bytes_254(5, lo, hi)
    sc = fr_bytes_254(55, lo, hi)
    step = 1 << 20
    p1 = b"".join(ctx.g1_generator_mul(ks[32 * i:32 * min(m, i + step)]) for i in range(0, m, step))
    p2 = b"".join(ctx.g2_generator_mul(ks[32 * i:32 * min(m, i + step)]) for i in range(0, m, step))
    d_p1 = torch.frombuffer(bytearray(p1), dtype=torch.uint8).to(device)
    d_p2 = torch.frombuffer(bytearray(p2), dtype=torch.uint8).to(device)
    d_sc = torch.frombuffer(bytearray(sc), dtype=torch.uint8).to(device)
    part = torch.zeros(192, dtype=torch.uint8, device=device)
    gathered = [torch.zeros(192, dtype=torch.uint8, device=device) for _ in range(world)]
    gen_s = time.time() - t0

    def one():
        rc = L.og_msm_g1_dev(ctx._h, d_p1.data_ptr(), d_sc.data_ptr(), m, part.data_ptr())
        rc |= L.og_msm_g2_dev(ctx._h, d_p2.data_ptr(), d_sc.data_ptr(), m, part.data_ptr() + 64)
        assert rc == 0, L.og_last_error(ctx._h)
        ctx.sync()                                   # the collective runs on torch's stream
        dist.all_gather(gathered, part)
        allb = b"".join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
        g1 = ctx.g1_sum(b"".join(allb[192 * r:192 * r + 64] for r in range(world)))
        g2 = ctx.g2_sum(b"".join(allb[192 * r + 64:192 * r + 192] for r in range(world)))
        return g1, g2

    for _ in range(args.warmup):
        res = one()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = one()
    torch.cuda.synchronize(); dist.barrier()
    ms = 1e3 * (time.perf_counter() - t0) / args.steps
    tt = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    # agreement across ranks
    mine = torch.frombuffer(bytearray(res[0] + res[1]), dtype=torch.uint8).to(device)
    ref = mine.clone(); dist.broadcast(ref, 0)
    agree = torch.tensor([int(torch.equal(mine, ref))], device=device); dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    verified = None
    if rank == 0 and args.log_n <= args.verify_max and world > 1:
        ksa, sca = fr_bytes_254(5, 0, n), fr_bytes_254(55, 0, n)
        full1 = ctx.msm_g1(b"".join(ctx.g1_generator_mul(ksa[32 * i:32 * min(n, i + step)]) for i in range(0, n, step)), sca)
        full2 = ctx.msm_g2(b"".join(ctx.g2_generator_mul(ksa[32 * i:32 * min(n, i + step)]) for i in range(0, n, step)), sca)
        verified = (full1 == res[0] and full2 == res[1])
    if rank == 0:
        alg = (64 + 128 + 32) * n
        print(json.dumps({"metric": "sharded_msm_g1_g2_points_per_sec", "value": n / (ms * 1e-3), "unit": "points/s", "n_gpus": world,
                          "log_n": args.log_n, "ms_per_msm_pair": ms, "steps": args.steps, "warmup": args.warmup, "scaling": "strong",
                          "algorithmic_bytes": alg, "hbm_gbs_aggregate": alg / (ms * 1e-3) / 1e9, "exchange_bytes_per_rank": 192,
                          "ranks_agree": bool(agree.item()), "matches_single_gpu": verified, "input_generation_s": gen_s,
                          "timing": "host wall between barriers (includes the all-gather and the host hop of 192 B/rank), max over ranks",
                          "config": {"workload": f"2^{args.log_n}-point G1+G2 MSM, shared scalars, point-range sharded (BASELINE config 5)"}}))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
