"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two ways the path shards:
  * proofs/sec: proofs are independent -> each rank proves its own slice of the batch, no collective
    (split_batch);
  * one large MSM (BASELINE config 5): point-range sharding -> every rank runs a full Pippenger on its
    slice, the partial sums (64 B in G1, 128 B in G2 per rank) are all-gathered and added locally
    (`ncclSum` cannot add curve points, so the "reduce" is all-gather + og_g1_sum / og_g2_sum).
The collective is the only torch call; every group operation runs in the CUDA library.
"""
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n items for `rank`; the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def split_batch(n: int, rank: int, world: int):
    return shard_range(n, rank, world)


def gather_partials(partial: bytes, device=None, group=None):
    """All-gather one fixed-size byte string per rank; returns the concatenation in rank order.
    Works on the gloo backend with CPU tensors and on NCCL with device tensors."""
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(partial), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return b"".join(bytes(x.cpu().numpy().tobytes()) for x in out)


def msm_sharded(ctx, points: bytes, scalars: bytes, curve: str = "g1", device=None, group=None) -> bytes:
    """MSM over points/scalars that every rank holds in full (synthetic benchmark layout) or that the
    caller already sliced (pass the local slice and it is used as is when world == 1)."""
    pb = 64 if curve == "g1" else 128
    n = len(scalars) // 32
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n, rank, world)
    msm = ctx.msm_g1 if curve == "g1" else ctx.msm_g2
    partial = msm(points[pb * lo:pb * hi], scalars[32 * lo:32 * hi])
    allp = gather_partials(partial, device, group)
    return (ctx.g1_sum if curve == "g1" else ctx.g2_sum)(allp)
