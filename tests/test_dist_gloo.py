"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shard ranges, the all-gather of partial
sums and the local combine, through owshen_b200.sharded.msm_sharded itself.  The group arithmetic of each
rank is answered by the oracle through a stand-in context (there is no GPU here and the product has no CPU
path); tests/test_gpu_parity.py::test_msm_sharded_* run the same function with the CUDA library (device
tensors, NCCL all-gather on the library's stream)."""
import os
import random
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bn254 as bn
from oracle import cport
from owshen_b200.sharded import gather_partials, msm_sharded, shard_range, split_batch
from tests.helpers import rand_g1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class OracleCtx:
    """What msm_sharded needs from a context without device entry points: msm_g1 / g1_sum on host bytes."""
    def msm_g1(self, points, scalars):
        return cport.g1_msm(points, scalars)

    def g1_sum(self, points):
        total = bytes(64)
        for i in range(0, len(points), 64):
            total = cport.g1_add(total, points[i:i + 64])
        return total


def _worker(rank, world, port, pts, sc, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = msm_sharded(OracleCtx(), pts, sc, "g1")
    # the pieces msm_sharded is made of, checked on their own as well
    n = len(sc) // 32
    lo, hi = shard_range(n, rank, world)
    allp = gather_partials(cport.g1_msm(pts[64 * lo:64 * hi], sc[32 * lo:32 * hi]))
    assert OracleCtx().g1_sum(allp) == total
    q.put((rank, total))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 1024, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert split_batch(1024, 3, 8) == (384, 512)


def test_point_range_sharded_msm_world2():
    rng = random.Random(5)
    n = 301
    pts = rand_g1(rng, n)
    sc = cport.frs([rng.randrange(bn.R) for _ in range(n)])
    expect = cport.g1_msm(pts, sc)
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, pts, sc, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == expect and got[1] == expect
