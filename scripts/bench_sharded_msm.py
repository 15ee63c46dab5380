#!/usr/bin/env python
"""BASELINE config 5: one G1+G2 MSM of 2^LOG_N points sharded by point range over the GPUs of one box.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/bench_sharded_msm.py --log-n 24 [--steps 3 --warmup 2]

Runs the product function owshen_b200.sharded.msm_sharded_dev (the same leg bench.py reports as `sharded_msm` on a
2^22 instance): every rank synthesises ITS slice of the inputs on its own GPU (points = k_i * G with seeded k_i,
253-bit seeded scalars shared by the G1 and the G2 MSM), runs a full Pippenger on the slice, the partial sums
(64 B + 128 B per rank) are all-gathered over NCCL on the library's stream and added on every rank.  ncclSum cannot add
curve points, so the "reduce" is all-gather + local group addition; nothing returns to the host between the steps.
Timing: CUDA events on the library stream, max over ranks.  Rank 0 recomputes the whole MSM alone for the check.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import owshen_b200 as ob  # noqa: E402
from owshen_b200 import api  # noqa: E402
from bench import sharded_msm_leg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    ctx = ob.Context(local)
    res = sharded_msm_leg(torch, dist if world > 1 else None, ob, api, ctx, device, rank, world, args.log_n, args.steps, args.warmup)
    if rank == 0:
        res.update({"metric": "sharded_msm_g1_g2_points_per_sec", "value": res["points_per_s"], "unit": "points/s", "scaling": "strong"})
        print(json.dumps(res))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
