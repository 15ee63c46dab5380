"""Host logic of owshen_b200.api.MerkleTree (sparse tree bookkeeping, KvStore persistence, path extraction) on the CPU:
the tree only asks its context for batched two-to-one hashes, so a stand-in context that answers them from the oracle
exercises everything but the kernel (the kernel's own parity is tests/test_gpu_parity.py)."""
import random

import pytest

from owshen_b200 import api
from owshen_b200.kvstore import RamKvStore
from oracle import bn254 as bn
from oracle import mimc7


class OracleHashCtx:
    """What MerkleTree needs from a Context: mimc7_hash2(left bytes, right bytes) -> bytes."""
    def __init__(self):
        self.calls = 0

    def mimc7_hash2(self, left: bytes, right: bytes) -> bytes:
        self.calls += 1
        assert len(left) == len(right) and len(left) % 32 == 0
        out = []
        for i in range(0, len(left), 32):
            a, b = int.from_bytes(left[i:i + 32], "little"), int.from_bytes(right[i:i + 32], "little")
            out.append(mimc7.hash2(a, b).to_bytes(32, "little"))
        return b"".join(out)


def fr(x):
    return x.to_bytes(32, "little")


def test_insert_root_path_match_the_spec_tree():
    rng = random.Random(5)
    ctx = OracleHashCtx()
    t = api.MerkleTree(ctx, 6)
    ref = mimc7.MerkleTree(6)
    assert t.root() == fr(ref.root())                       # empty tree
    leaves = [rng.randrange(bn.R) for _ in range(11)]
    idx = t.insert_batch([fr(x) for x in leaves[:7]])
    assert idx == list(range(7))
    before = ctx.calls
    assert t.insert(fr(leaves[7])) == 7
    assert ctx.calls - before == 6                          # one batched hash call per level
    t.insert_batch(leaves[8:])                              # integers are accepted too
    for x in leaves:
        ref.insert(x)
    assert t.root() == fr(ref.root()) and t.n_leaves == 11
    for i in (0, 3, 10):
        sib, bits = t.path(i)
        rs, rb = ref.path(i)
        assert sib == b"".join(fr(x) for x in rs) and bits == sum(b << k for k, b in enumerate(rb))
        node = leaves[i]                                     # the path hashes back to the root
        for lvl in range(6):
            s = int.from_bytes(sib[32 * lvl:32 * lvl + 32], "little")
            node = mimc7.hash2(s, node) if (bits >> lvl) & 1 else mimc7.hash2(node, s)
        assert fr(node) == t.root()
    sibs, bl = t.paths([10, 0])
    assert sibs == t.path(10)[0] + t.path(0)[0] and bl == [t.path(10)[1], t.path(0)[1]]
    with pytest.raises(IndexError):
        t.path(11)


def test_reopen_over_the_same_store():
    ctx = OracleHashCtx()
    store = RamKvStore()
    t = api.MerkleTree(ctx, 5, store=store)
    t.insert_batch([fr(i + 1) for i in range(9)])
    again = api.MerkleTree(ctx, 5, store=store)
    assert again.n_leaves == 9 and again.root() == t.root() and again.path(4) == t.path(4)
    again.insert(fr(99))
    t2 = api.MerkleTree(ctx, 5, store=store)
    assert t2.n_leaves == 10 and t2.root() == again.root()
    with pytest.raises(ValueError):
        api.MerkleTree(ctx, 4, store=store)
    other = api.MerkleTree(ctx, 5, store=store, prefix=b"other/")    # a second tree in the same store
    assert other.n_leaves == 0 and other.root() != t2.root()
