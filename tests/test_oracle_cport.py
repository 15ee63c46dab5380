"""The C port (oracle/cpu) against the pure-Python spec, bit for bit (CPU only)."""
import random

import pytest

from oracle import bn254 as bn
from oracle import cport, groth16 as g16, mimc7
from oracle import ntt as pntt
from oracle.withdraw_circuit import build_r1cs, witness

R, P = bn.R, bn.P


def test_field_ops():
    rng = random.Random(7)
    for name, mod, pack in (("fr", R, cport.frs), ("fq", P, cport.fqs)):
        xs = [rng.randrange(mod) for _ in range(200)] + [0, 1, mod - 1, mod - 1]
        ys = [rng.randrange(mod) for _ in range(200)] + [mod - 1, mod - 1, mod - 1, 1]
        assert cport.unfr(cport.field_binop(f"oc_{name}_mul", pack(xs), pack(ys))) == [a * b % mod for a, b in zip(xs, ys)]
        assert cport.unfr(cport.field_binop(f"oc_{name}_add", pack(xs), pack(ys))) == [(a + b) % mod for a, b in zip(xs, ys)]
        assert cport.unfr(cport.field_binop(f"oc_{name}_sub", pack(xs), pack(ys))) == [(a - b) % mod for a, b in zip(xs, ys)]
        assert cport.unfr(cport.field_inv(f"oc_{name}_inv", pack(xs[:5]))) == [pow(a, -1, mod) for a in xs[:5]]
    with pytest.raises(ValueError):
        cport.field_binop("oc_fr_mul", R.to_bytes(32, "little"), bytes(32))


def test_curve_ops_and_msm():
    rng = random.Random(8)
    g1, g2 = bn.g1_to_bytes(bn.G1_GEN), bn.g2_to_bytes(bn.G2_GEN)
    for k in [0, 1, 2, R - 1, rng.randrange(R)]:
        assert cport.g1_mul(g1, bn.fr_to_bytes(k)) == bn.g1_to_bytes(bn.g1_mul(bn.G1_GEN, k))
        assert cport.g2_mul(g2, bn.fr_to_bytes(k)) == bn.g2_to_bytes(bn.g2_mul(bn.G2_GEN, k))
    n = 40
    pts = [bn.g1_mul(bn.G1_GEN, rng.randrange(R)) for _ in range(n)]
    pts[3] = None; pts[5] = pts[4]
    sc = [rng.randrange(R) for _ in range(n)]
    sc[0] = 0; sc[1] = 1; sc[2] = R - 1; sc[5] = sc[4]
    assert cport.g1_msm(b"".join(map(bn.g1_to_bytes, pts)), cport.frs(sc)) == bn.g1_to_bytes(bn.g1_msm(pts, sc))
    pts2 = [bn.g2_mul(bn.G2_GEN, rng.randrange(R)) for _ in range(10)]
    assert cport.g2_msm(b"".join(map(bn.g2_to_bytes, pts2)), cport.frs(sc[:10])) == bn.g2_to_bytes(bn.g2_msm(pts2, sc[:10]))
    assert cport.g1_fixed_mul_batch(g1, cport.frs(sc[:8])) == b"".join(bn.g1_to_bytes(bn.g1_mul(bn.G1_GEN, k)) for k in sc[:8])
    assert cport.g2_fixed_mul_batch(g2, cport.frs(sc[:4])) == b"".join(bn.g2_to_bytes(bn.g2_mul(bn.G2_GEN, k)) for k in sc[:4])
    assert cport.g1_msm(b"", b"") == bytes(64)


def test_ntt():
    rng = random.Random(9)
    for ln in (0, 1, 2, 5, 8):
        v = [rng.randrange(R) for _ in range(1 << ln)]
        for inv in (False, True):
            for co in (False, True):
                assert cport.unfr(cport.ntt(cport.frs(v), inv, co)) == pntt.ntt(v, inv, co)


def test_mimc_merkle_witness():
    rng = random.Random(10)
    assert cport.mimc7_hash(1, 2) == mimc7.mimc7_hash(1, 2)
    assert cport.mimc7_multi_hash([1, 2, 3, 4]) == mimc7.multi_hash([1, 2, 3, 4])
    depth = 5
    leaves = [rng.randrange(R) for _ in range(3)]
    sibs = [[rng.randrange(R) for _ in range(depth)] for _ in range(3)]
    bits = [rng.randrange(1 << depth) for _ in range(3)]
    out = cport.unfr(cport.merkle_paths(cport.frs(leaves), cport.frs(sum(sibs, [])), bits, depth))
    for p in range(3):
        assert out[p * (depth + 1):(p + 1) * (depth + 1)] == mimc7.merkle_path_nodes(
            leaves[p], sibs[p], [(bits[p] >> l) & 1 for l in range(depth)])
    w = witness(11, 22, 33, sibs[0], [(bits[0] >> l) & 1 for l in range(depth)])
    wc = cport.unfr(cport.withdraw_witness(cport.frs([11]), cport.frs([22]), cport.frs([33]), cport.frs(sibs[0]), [bits[0]], depth))
    assert w == wc


def test_groth16_small_circuit_bit_exact():
    rng = random.Random(11)
    nr, depth = 3, 2
    cport.set_mimc_rounds(nr)
    try:
        cs = build_r1cs(depth, nr)
        tw = [rng.randrange(1, R) for _ in range(5)]
        pk, vk = g16.setup(cs, *tw)
        pkb, vkb = cport.setup_bytes(cs, *tw)
        assert pkb["a"] == b"".join(map(bn.g1_to_bytes, pk["a"])) and pkb["h"] == b"".join(map(bn.g1_to_bytes, pk["h"]))
        assert pkb["b1"] == b"".join(map(bn.g1_to_bytes, pk["b1"])) and pkb["b2"] == b"".join(map(bn.g2_to_bytes, pk["b2"]))
        assert pkb["l"] == b"".join(map(bn.g1_to_bytes, pk["l"])) and vkb["ic"] == b"".join(map(bn.g1_to_bytes, vk["ic"]))
        w = witness(11, 22, 33, [5, 6], [1, 0], nr)
        pr = cport.Prover(cs, pkb)
        assert cport.unfr(pr.h_evals(cport.frs(w))) == g16.h_evals(cs, w, pk["log_m"])
        r, s = rng.randrange(R), rng.randrange(R)
        pb = pr.prove(cport.frs(w), r, s)
        assert pb == g16.proof_to_bytes(g16.prove(cs, pk, w, r, s))
        assert pr.prove_batch(cport.frs(w) * 2, (bn.fr_to_bytes(r) + bn.fr_to_bytes(s)) * 2) == pb * 2
        assert g16.verify(vk, w[1:4], g16.proof_from_bytes(pb))
    finally:
        cport.set_mimc_rounds(91)
