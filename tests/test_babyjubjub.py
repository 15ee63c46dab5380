"""The reference's own BabyJubJub tests (/root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/tests.rs)
re-run against the oracle's restatement (CPU), and the GPU batch verifier against the oracle (gpu)."""
import random

import pytest

from oracle import babyjubjub as bjj
from oracle import bn254 as bn

R = bn.R


def test_twisted_edwards_curve_ops():            # tests.rs:3-29
    a = bjj.add(bjj.add(bjj.double(bjj.BASE), bjj.BASE), bjj.BASE)
    b = bjj.double(bjj.double(bjj.BASE))
    assert a == b
    c = bjj.BASE
    for _ in range(3):
        c = bjj.add(c, bjj.BASE)
    assert b == c
    base_p = (bjj.BASE[0], bjj.BASE[1], 1)
    pnt1 = bjj.p_add(bjj.p_double(bjj.p_double(base_p)), base_p)
    pnt2 = bjj.add(bjj.double(bjj.double(bjj.BASE)), bjj.BASE)
    assert bjj.p_to_affine(pnt1) == pnt2


def test_jubjub_public_key_compression():        # tests.rs:31-37
    p1 = bjj.multiply(bjj.BASE, 123)
    assert bjj.decompress(bjj.compress(p1)) == p1


def test_jubjub_signature_verification():        # tests.rs:39-51
    sk, pk = 12345, bjj.to_pub(12345)
    sig = bjj.sign(sk, 2345, 123456)
    assert bjj.verify(pk, 123456, sig)
    assert not bjj.verify(pk, 123457, sig)


def test_curve_facts_the_reference_relies_on():
    assert bjj.is_on_curve(bjj.BASE) and bjj.is_on_curve(bjj.ZERO)
    assert bjj.ORDER % 8 == 0 and bjj.multiply(bjj.BASE, (bjj.ORDER // 8) % R) == bjj.ZERO   # BASE has prime order l
    assert bjj.multiply(bjj.BASE, 0) == bjj.ZERO
    assert bjj.multiply(bjj.BASE, 5) == bjj.add(bjj.multiply(bjj.BASE, 2), bjj.multiply(bjj.BASE, 3))
    with pytest.raises(bjj.CannotInvert):
        bjj.decompress((3, 0))                   # x = 3: (1 - a x^2)/(1 - d x^2) is a non-residue


def _cases(rng, n, hash_kind):
    pks, msgs, sigs, expect = [], [], [], []
    for i in range(n):
        sk = rng.randrange(1, R)
        msg = rng.randrange(R)
        while True:
            try:
                sig = bjj.sign(sk, rng.randrange(R), msg, hash_kind)
                break
            except ValueError:
                continue
        pk = bjj.to_pub(sk)
        kind = i % 4
        if kind == 1:
            msg = (msg + 1) % R                                  # wrong message
        elif kind == 2:
            sig = (sig[0], (sig[1] + 1) % R)                     # wrong s
        elif kind == 3 and i % 8 == 3:
            sig = ((sig[0][0], (sig[0][1] + 1) % R), sig[1])     # R off the curve
        pks.append(pk); msgs.append(msg); sigs.append(sig)
        expect.append(bjj.verify(pk, msg, sig, hash_kind))
    return pks, msgs, sigs, expect


def _pack(pks, msgs, sigs):
    fb = bn.fr_to_bytes
    pkx = b"".join(fb(p[0]) for p in pks)
    odd = bytes(p[1] for p in pks)
    m = b"".join(fb(x) for x in msgs)
    s = b"".join(fb(g[0][0]) + fb(g[0][1]) + fb(g[1]) for g in sigs)
    return pkx, odd, m, s


@pytest.mark.gpu
@pytest.mark.parametrize("hash_kind", [0, 1])
def test_gpu_batch_verify_matches_oracle(ctx, hash_kind):
    rng = random.Random(77 + hash_kind)
    pks, msgs, sigs, expect = _cases(rng, 48, hash_kind)
    assert any(expect) and not all(expect)
    got = ctx.bjj_verify_batch(*_pack(pks, msgs, sigs), hash_kind=hash_kind)
    assert list(got) == [1 if e else 0 for e in expect]
    # the reference's own vector (tests.rs:39-51) through the GPU path
    sig = bjj.sign(12345, 2345, 123456)
    pk = bjj.to_pub(12345)
    assert list(ctx.bjj_verify_batch(*_pack([pk, pk], [123456, 123457], [sig, sig]))) == [1, 0]
    # undecompressible public key -> status 2 (the reference returns Err)
    assert list(ctx.bjj_verify_batch(*_pack([(3, 0)], [1], [sig]))) == [2]


@pytest.mark.gpu
@pytest.mark.parametrize("hash_kind", [0, 1])
def test_gpu_batch_sign_matches_oracle(ctx, hash_kind):
    """og_bjj_sign_batch = PrivateKey::to_pub + PrivateKey::sign (mod.rs:206-237) for a batch of keys: public keys and
    signatures equal the oracle's restatement bit for bit, verify on the GPU, and the reference's own test vector
    (tests.rs:39-51: sk 12345, randomness 2345, message 123456) comes out the same."""
    rng = random.Random(990 + hash_kind)
    n = 40
    sks = [rng.randrange(R) for _ in range(n - 3)] + [12345, 0, 1]
    rnds = [rng.randrange(R) for _ in range(n - 3)] + [2345, 7, 0]
    msgs = [rng.randrange(R) for _ in range(n - 3)] + [123456, 9, 11]
    fb = bn.fr_to_bytes
    pkx, odd, sigs, st = ctx.bjj_sign_batch(b"".join(map(fb, sks)), b"".join(map(fb, rnds)), b"".join(map(fb, msgs)), hash_kind)
    assert list(st) == [1] * n
    exp_pk = [bjj.to_pub(k) for k in sks]
    exp_sig = [bjj.sign(k, r_, m, hash_kind) for k, r_, m in zip(sks, rnds, msgs)]
    epx, eodd, em, es = _pack(exp_pk, msgs, exp_sig)
    assert pkx == epx and odd == eodd and sigs == es
    assert list(ctx.bjj_verify_batch(pkx, odd, em, sigs, hash_kind=hash_kind)) == [1] * n
    with pytest.raises(ValueError):
        ctx.bjj_sign_batch(bytes(32), bytes(31), bytes(32))
