"""External pins ("any external pin for BN254 group/pairing arithmetic"): published known answers that this
repository's code did not produce -- tests/golden/external_vectors.json states where each one comes from.  Both oracles
(pure Python, C port) and the product's host-side pairing must reproduce them; the CUDA library is checked against the
same file in tests/test_gpu_parity.py::test_external_known_answers."""
import json
import os
import struct

from oracle import babyjubjub as bj
from oracle import bn254 as bn
from oracle import cport
from oracle import pairing as pr
from owshen_b200 import api, formats

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "external_vectors.json")))
H = lambda s: int(s, 16)


def _pt(xy):
    return (H(xy[0]), H(xy[1]))


def jeff1_pairs():
    w = [H(x) for x in V["eip197"]["pairing_jeff1"]["words"]]
    out = []
    for k in range(len(w) // 6):
        x, y, x1, x0, y1, y0 = w[6 * k:6 * k + 6]
        out.append(((x, y), ((x0, x1), (y0, y1))))
    return out


def test_bn_parameter_generates_both_moduli():
    u = int(V["bn_parameter_u"])
    assert 36 * u**4 + 36 * u**3 + 24 * u**2 + 6 * u + 1 == bn.P
    assert 36 * u**4 + 36 * u**3 + 18 * u**2 + 6 * u + 1 == bn.R == api.FR_MODULUS
    assert 6 * u + 2 == int(V["ate_loop_count_6u_plus_2"])


def test_eip196_add_and_mul_both_oracles():
    e = V["eip196"]
    two = (H(e["g1_generator_doubled"]["x"]), H(e["g1_generator_doubled"]["y"]))
    assert bn.g1_add(bn.G1_GEN, bn.G1_GEN) == two == bn.g1_mul(bn.G1_GEN, 2)
    a, b, s = _pt(e["add_chfast1"]["a"]), _pt(e["add_chfast1"]["b"]), _pt(e["add_chfast1"]["sum"])
    assert bn.g1_on_curve(a) and bn.g1_on_curve(b) and bn.g1_add(a, b) == s
    assert cport.g1_add(bn.g1_to_bytes(a), bn.g1_to_bytes(b)) == bn.g1_to_bytes(s)
    p, k, q = _pt(e["mul_chfast1"]["p"]), H(e["mul_chfast1"]["k"]), _pt(e["mul_chfast1"]["product"])
    assert bn.g1_mul(p, k) == q
    assert cport.g1_msm(bn.g1_to_bytes(p), cport.frs([k])) == bn.g1_to_bytes(q)
    assert cport.g1_msm(bn.g1_to_bytes(bn.G1_GEN), cport.frs([2])) == bn.g1_to_bytes(two)


def test_eip197_pairing_check_python_oracle():
    pairs = jeff1_pairs()
    for g1, g2 in pairs:
        assert bn.g1_on_curve(g1) and bn.g2_on_curve(g2)
    assert pairs[1][1] == bn.G2_GEN                          # the vector's second G2 point is the standard generator
    assert pr.pairing_product_is_one(pairs) is V["eip197"]["pairing_jeff1"]["expected"]
    assert not pr.pairing_product_is_one([pairs[0], (bn.g1_mul(pairs[1][0], 2), pairs[1][1])])


def test_eip197_pairing_check_product_host_pairing():
    """The same published pairs through the product's own C++ pairing (og_groth16_verify, pairing.cu): a verifying key
    and "proof" assembled so that the Groth16 equation e(-A,B) e(alpha,beta) e(X,gamma) e(C,delta) = 1 IS the vector's
    product e(P1,Q1) e(P2,G2) = 1 times e(T,G2) e(-T,G2).  No GPU involved: verify() is a host function by design."""
    (p1, q1), (p2, q2) = jeff1_pairs()
    t = bn.g1_mul(bn.G1_GEN, 5)
    g1b, g2b = bn.g1_to_bytes, bn.g2_to_bytes
    vk = b"OGVK" + struct.pack("<II", 1, 0) + g1b(p2) + g2b(q2) + g2b(bn.G2_GEN) + g2b(bn.G2_GEN) + g1b(t)
    proof = g1b(bn.g1_neg(p1)) + g2b(q1) + g1b(bn.g1_neg(t))
    assert api.verify(vk, b"", proof) is True
    bad = g1b(bn.g1_neg(bn.g1_mul(p1, 3))) + g2b(q1) + g1b(bn.g1_neg(t))
    assert api.verify(vk, b"", bad) is False
    # and the EIP-197 word order of owshen_b200.formats is the vector's: imaginary part first, big-endian
    words = b"".join(bytes.fromhex(x) for x in V["eip197"]["pairing_jeff1"]["words"])
    enc = formats.proof_to_eip197(g1b(p1) + g2b(q1) + g1b(p2))
    assert enc[:192] == words[:192] and enc[192:256] == words[192:256]
    assert formats.proof_from_eip197(enc) == g1b(p1) + g2b(q1) + g1b(p2)


def test_babyjubjub_constants_match_circomlib_and_the_reference():
    c = V["babyjubjub_circomlib"]
    base8 = (int(c["base8"][0]), int(c["base8"][1]))
    gen = (int(c["generator"][0]), int(c["generator"][1]))
    assert bj.BASE == base8 and bj.A == c["a"] and bj.D == c["d"] and bj.ORDER == int(c["order"])
    assert bj.is_on_curve(gen) and bj.multiply(gen, 8) == base8          # Base8 = 8 * generator (circomlib)
    assert int(c["order"]) == 8 * int(c["suborder"])
    assert bj.multiply(base8, int(c["suborder"])) == bj.ZERO             # Base8 generates the prime-order subgroup
