"""The committed golden vectors (tests/golden/vectors.json, made by tests/golden/gen_golden.py from the
pure-Python spec) against the oracle's C port -- the CPU half of the pinning; the GPU half is in
tests/test_gpu_parity.py."""
import hashlib
import json
import os
import random

from oracle import bn254 as bn
from oracle import cport, mimc7
from oracle.withdraw_circuit import build_r1cs

R = bn.R
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


def test_field_and_curve_vectors():
    for v in GOLD["field"]:
        pack = cport.fqs if v["field"] == "fq" else cport.frs
        a, b = pack([int(x) for x in v["a"]]), pack([int(x) for x in v["b"]])
        for op in ("mul", "add", "sub"):
            assert cport.unfr(cport.field_binop(f"oc_{v['field']}_{op}", a, b)) == [int(x) for x in v[op]]
    g1, g2 = bn.g1_to_bytes(bn.G1_GEN), bn.g2_to_bytes(bn.G2_GEN)
    for v in GOLD["g1_mul"]:
        assert cport.g1_mul(g1, bn.fr_to_bytes(int(v["k"]))).hex() == v["out"]
    for v in GOLD["g2_mul"]:
        assert cport.g2_mul(g2, bn.fr_to_bytes(int(v["k"]))).hex() == v["out"]
    for v in GOLD["msm"]:
        f = cport.g1_msm if v["curve"] == "g1" else cport.g2_msm
        assert f(bytes.fromhex(v["points"]), bytes.fromhex(v["scalars"])).hex() == v["out"]


def test_ntt_vectors():
    for v in GOLD["ntt"]:
        r2 = random.Random(v["seed"])
        data = cport.frs([r2.randrange(R) for _ in range(1 << v["log_n"])])
        got = cport.ntt(data, v["inverse"], v["coset"])
        assert hashlib.sha256(got).hexdigest() == v["sha256"]
        if v["out"]:
            assert got.hex() == v["out"]


def test_mimc_vectors():
    g = GOLD["mimc7"]
    assert int(g["c1"]) == mimc7.CONSTANTS[1] and int(g["c90"]) == mimc7.CONSTANTS[90]
    assert cport.mimc7_hash(1, 2) == int(g["hash_1_2"])
    assert cport.mimc7_multi_hash([1, 2]) == int(g["multi_hash_1_2"])
    assert cport.mimc7_multi_hash([1, 2, 3, 4]) == int(g["multi_hash_1_2_3_4"])
    p = g["path"]
    got = cport.unfr(cport.merkle_paths(cport.frs([int(p["leaf"])]), cport.frs([int(x) for x in p["siblings"]]), [p["bits"]], 3))
    assert got == [int(x) for x in p["nodes"]]
    lv = [int(x) for x in g["tree4_leaves"]]
    l1 = [cport.mimc7_multi_hash(lv[0:2]), cport.mimc7_multi_hash(lv[2:4])]
    assert cport.mimc7_multi_hash(l1) == int(g["tree4_root"])


def test_groth16_golden_proof_reproduced_by_c_port():
    g = GOLD["groth16"]
    cs = build_r1cs(g["depth"])
    pkb, vkb = cport.setup_bytes(cs, *[int(x) for x in g["toxic"]])
    q = pkb["a"] + pkb["b1"] + pkb["b2"] + pkb["l"] + pkb["h"]
    assert hashlib.sha256(q).hexdigest() == g["pk_queries_sha256"]
    v = g["vk"]
    assert (vkb["alpha1"] + vkb["beta2"] + vkb["gamma2"] + vkb["delta2"] + vkb["ic"]).hex() == v["alpha1"] + v["beta2"] + v["gamma2"] + v["delta2"] + v["ic"]
    f = lambda k: bn.fr_to_bytes(int(g[k]))
    wit = cport.withdraw_witness(f("nullifier"), f("secret"), f("recipient"), f("sibling"), [g["bits"]], g["depth"])
    assert hashlib.sha256(wit).hexdigest() == g["witness_sha256"]
    assert cport.unfr(wit[32:128]) == [int(x) for x in g["public"]]
    assert cport.Prover(cs, pkb).prove(wit, int(g["r"]), int(g["s"])).hex() == g["proof"]
