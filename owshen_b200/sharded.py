"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two ways the path shards:
  * proofs/sec: proofs are independent -> each rank proves its own slice of the batch, no collective
    (split_batch);
  * one large MSM (BASELINE config 5): point-range sharding -> every rank runs a full Pippenger on its
    slice, the partial sums (64 B in G1, 128 B in G2 per rank) are all-gathered and added locally
    (`ncclSum` cannot add curve points, so the "reduce" is all-gather + og_g1_sum / og_g2_sum).

`msm_sharded_dev` is the product path: inputs, partial sums, the gathered partials and the result are device
tensors, the all-gather is enqueued on the library's own stream (torch.cuda.ExternalStream over og_stream), so
there is no host hop and no host synchronisation between the MSM, the exchange and the final sum -- CUDA events
on that stream time the whole thing.  `msm_sharded` is the host-buffer convenience over it; with a context that
has no device entry points (the CPU tests' stand-in) it uses the host-pointer calls and a CPU all-gather instead.
The collective is the only torch call; every group operation runs in the CUDA library.
"""
import torch
import torch.distributed as dist

POINT_BYTES = {"g1": 64, "g2": 128}


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


def library_stream(ctx):
    """The library's CUDA stream as a torch stream, so torch collectives can be ordered against the kernels."""
    return torch.cuda.ExternalStream(ctx.cuda_stream, device=torch.device("cuda", ctx.device))


def msm_sharded_dev(ctx, d_points: torch.Tensor, d_scalars: torch.Tensor, curve: str = "g1", group=None,
                    out: torch.Tensor = None) -> torch.Tensor:
    """This rank's slice (device uint8 tensors: affine points and canonical scalars) -> the MSM over ALL ranks'
    slices as a device tensor of 64 (G1) / 128 (G2) bytes, identical on every rank.  Nothing is synchronised."""
    from . import api
    if curve not in POINT_BYTES:
        raise ValueError("curve must be 'g1' or 'g2'")
    pb = POINT_BYTES[curve]
    if d_scalars.numel() % 32 or d_points.numel() != pb * (d_scalars.numel() // 32):
        raise ValueError("msm_sharded_dev: need one point per 32-byte scalar")
    if not (d_points.is_cuda and d_scalars.is_cuda and d_points.dtype == torch.uint8 and d_scalars.dtype == torch.uint8):
        raise ValueError("msm_sharded_dev: inputs must be CUDA uint8 tensors")
    n = d_scalars.numel() // 32
    L = api.lib()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = d_points.device
    msm = L.og_msm_g1_dev if curve == "g1" else L.og_msm_g2_dev
    gsum = L.og_g1_sum_dev if curve == "g1" else L.og_g2_sum_dev
    with torch.cuda.stream(library_stream(ctx)):
        part = torch.empty(pb, dtype=torch.uint8, device=dev)
        api._check(msm(ctx._h, d_points.data_ptr(), d_scalars.data_ptr(), n, part.data_ptr()), ctx)
        if world == 1:
            gathered = part
        else:
            gathered = torch.empty(pb * world, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, part, group=group)      # 64 / 128 B per rank, on the library's stream
        res = out if out is not None else torch.empty(pb, dtype=torch.uint8, device=dev)
        api._check(gsum(ctx._h, gathered.data_ptr(), world, res.data_ptr()), ctx)
    return res


def msm_sharded(ctx, points: bytes, scalars: bytes, curve: str = "g1", device=None, group=None) -> bytes:
    """MSM over points/scalars that every rank holds in full (host bytes): each rank takes its shard_range, the
    partial sums are all-gathered and added.  With a CUDA context the slice is uploaded and msm_sharded_dev does
    the rest on the device; a context without `cuda_stream` (CPU stand-in) goes through the host-pointer calls."""
    if curve not in POINT_BYTES:
        raise ValueError("curve must be 'g1' or 'g2'")
    pb = POINT_BYTES[curve]
    if len(scalars) % 32 or len(points) != pb * (len(scalars) // 32):
        raise ValueError("msm_sharded: need one point per 32-byte scalar")
    n = len(scalars) // 32
    if dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(n, rank, world)
    if hasattr(ctx, "cuda_stream"):
        dev = torch.device("cuda", ctx.device)
        with torch.cuda.stream(library_stream(ctx)):
            d_p = torch.frombuffer(bytearray(points[pb * lo:pb * hi]) or bytearray(1), dtype=torch.uint8).to(dev)[:pb * (hi - lo)]
            d_s = torch.frombuffer(bytearray(scalars[32 * lo:32 * hi]) or bytearray(1), dtype=torch.uint8).to(dev)[:32 * (hi - lo)]
        res = msm_sharded_dev(ctx, d_p, d_s, curve, group)
        ctx.sync()
        return bytes(res.cpu().numpy().tobytes())
    msm = ctx.msm_g1 if curve == "g1" else ctx.msm_g2
    partial = msm(points[pb * lo:pb * hi], scalars[32 * lo:32 * hi])
    allp = gather_partials(partial, device, group) if world > 1 else partial
    return (ctx.g1_sum if curve == "g1" else ctx.g2_sum)(allp)
