"""Host logic of owshen_b200.api.MerkleTree (sparse tree bookkeeping, KvStore persistence, undo records, path
extraction) on the CPU: the tree only asks its context for two-to-one hashes (the empty-subtree roots) and for the
nodes touched by an append, so a stand-in context that answers both from the oracle exercises everything but the
kernels (their own parity is tests/test_gpu_parity.py)."""
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

    def merkle_append(self, depth, start, leaves, left_boundary, zeros):
        """Restates og_mimc7_merkle_append (include/owshen_b200.h) with the oracle hash."""
        self.append_calls = getattr(self, "append_calls", 0) + 1
        cur = [int.from_bytes(leaves[i:i + 32], "little") for i in range(0, len(leaves), 32)]
        c0, out = start, []
        for l in range(depth):
            lb = int.from_bytes(left_boundary[32 * l:32 * l + 32], "little")
            z = int.from_bytes(zeros[32 * l:32 * l + 32], "little")
            def child(i):
                return lb if i < c0 else (cur[i - c0] if i < c0 + len(cur) else z)
            p0, p1 = c0 >> 1, (c0 + len(cur) - 1) >> 1
            nxt = [mimc7.hash2(child(2 * p), child(2 * p + 1)) for p in range(p0, p1 + 1)]
            out += nxt
            cur, c0 = nxt, p0
        return b"".join(x.to_bytes(32, "little") for x in out)


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
    before = ctx.append_calls
    assert t.insert(fr(leaves[7])) == 7
    assert ctx.append_calls - before == 1                   # one library call per insert, whatever the depth
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


def test_capacity_is_enforced():
    ctx = OracleHashCtx()
    t = api.MerkleTree(ctx, 3)
    t.insert_batch([fr(i + 1) for i in range(6)])
    root, n = t.root(), t.n_leaves
    with pytest.raises(OverflowError):
        t.insert_batch([fr(7), fr(8), fr(9)])               # 6 + 3 > 2^3: nothing may be written
    assert t.root() == root and t.n_leaves == n
    t.insert_batch([fr(7), fr(8)])                          # exactly full is fine
    assert t.n_leaves == 8
    sib, bits = t.path(7)
    node = 8
    for lvl in range(3):
        s = int.from_bytes(sib[32 * lvl:32 * lvl + 32], "little")
        node = mimc7.hash2(s, node) if (bits >> lvl) & 1 else mimc7.hash2(node, s)
    assert fr(node) == t.root()
    with pytest.raises(OverflowError):
        t.insert(fr(9))


def test_pop_batch_and_rollback_restore_earlier_roots():
    """Undo semantics of the reference's pop_block (src/blockchain/mod.rs:291-315): applying the stored delta restores
    every overwritten key, and the delta itself is deleted."""
    rng = random.Random(9)
    ctx = OracleHashCtx()
    store = RamKvStore()
    t = api.MerkleTree(ctx, 5, store=store)
    roots, sizes = [t.root()], [0]
    snapshot0 = dict(store.db)
    for n in (3, 1, 6, 2):
        t.insert_batch([fr(rng.randrange(bn.R)) for _ in range(n)])
        roots.append(t.root()); sizes.append(t.n_leaves)
    assert t.n_batches == 4
    assert t.pop_batch() == 2 and t.root() == roots[3] and t.n_leaves == sizes[3] and t.n_batches == 3
    # a reopened tree sees the popped state; re-inserting other leaves gives a different root, popping them restores it
    t2 = api.MerkleTree(ctx, 5, store=store)
    t2.insert_batch([fr(1), fr(2)])
    assert t2.root() != roots[4]
    t2.pop_batch()
    assert t2.root() == roots[3]
    # rollback to a batch boundary and into the middle of a batch
    t.rollback(sizes[2])
    assert t.root() == roots[2] and t.n_leaves == 4
    ref = mimc7.MerkleTree(5)
    leaves = [int.from_bytes(store.get_raw(t._key(0, i)), "little") for i in range(4)]
    for x in leaves[:2]:
        ref.insert(x)
    t.rollback(2)                                            # inside the first batch of 3
    assert t.n_leaves == 2 and t.root() == fr(ref.root())
    assert t.path(1)[0] == b"".join(fr(x) for x in ref.path(1)[0])
    t.rollback(0)
    assert t.root() == roots[0] and t.n_leaves == 0 and t.n_batches == 0
    live = {k: v for k, v in store.db.items() if k != b"mt/depth"}
    assert live == {k: v for k, v in snapshot0.items() if k != b"mt/depth"}      # nothing left behind, deltas included
    assert t.pop_batch() == 0
    with pytest.raises(ValueError):
        t.rollback(1)


def test_mirror_kvstore_matches_the_reference_shape():
    from owshen_b200.kvstore import MirrorKvStore
    base = RamKvStore()
    base.batch_put_raw([(b"a", b"1"), (b"b", b"2")])
    m = MirrorKvStore(base)
    m.batch_put_raw([(b"a", b"9"), (b"c", b"3"), (b"b", None)])
    assert m.get_raw(b"a") == b"9" and m.get_raw(b"b") is None and m.get_raw(b"c") == b"3" and base.get_raw(b"a") == b"1"
    assert m.rollback() == {b"a": b"1", b"b": b"2", b"c": None}           # mirror.rs:19-27
    base.batch_put_raw(m.buffer().items())
    assert base.db == {b"a": b"9", b"c": b"3"}


def test_random_insert_pop_rollback_sequences_against_the_spec_tree():
    """Model-based: a random walk of insert_batch / pop_batch / rollback / reopen against a list of leaves and the
    oracle's spec tree rebuilt from that list; roots, leaf counts and a sample path must agree after every operation,
    and the store must hold no stale undo records."""
    from hypothesis import given, settings, strategies as st

    ops = st.lists(st.one_of(st.tuples(st.just("ins"), st.integers(1, 9)), st.tuples(st.just("pop"), st.just(0)),
                             st.tuples(st.just("roll"), st.integers(0, 40)), st.tuples(st.just("reopen"), st.just(0))),
                   min_size=1, max_size=14)

    @settings(max_examples=25, deadline=None)
    @given(ops, st.integers(0, 2**32 - 1))
    def run(seq, seed):
        rng = random.Random(seed)
        ctx = OracleHashCtx()
        store = RamKvStore()
        depth = 5
        t = api.MerkleTree(ctx, depth, store=store)
        leaves, batches = [], []                      # model: the leaf list and the sizes of the live batches
        for op, arg in seq:
            if op == "ins":
                n = min(arg, (1 << depth) - len(leaves))
                if n == 0:
                    with pytest.raises(OverflowError):
                        t.insert(fr(1))
                    continue
                new = [rng.randrange(bn.R) for _ in range(n)]
                t.insert_batch(new); leaves += new; batches.append(n)
            elif op == "pop":
                removed = t.pop_batch()
                assert removed == (batches.pop() if batches else 0)
                if removed:
                    del leaves[-removed:]
            elif op == "roll":
                target = min(arg, len(leaves))
                t.rollback(target)
                del leaves[target:]
                kept, acc = [], 0                     # batches shrink from the end; a partially kept batch is re-inserted
                for b in batches:
                    if acc + b <= target:
                        kept.append(b); acc += b
                    else:
                        if target - acc:
                            kept.append(target - acc)
                        break
                batches = kept
            else:
                t = api.MerkleTree(ctx, depth, store=store)
            ref = mimc7.MerkleTree(depth)
            for x in leaves:
                ref.insert(x)
            assert t.n_leaves == len(leaves) and t.root() == fr(ref.root()) and t.n_batches == len(batches)
            if leaves:
                i = rng.randrange(len(leaves))
                assert t.path(i)[0] == b"".join(fr(x) for x in ref.path(i)[0])
            deltas = [k for k in store.db if k.startswith(b"mt/delta")]
            assert len(deltas) == len(batches)
    run()
