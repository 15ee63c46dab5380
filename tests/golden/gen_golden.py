"""Generates tests/golden/vectors.json from the pure-Python spec (oracle/*.py) with fixed seeds.

The reference holds no golden vectors for this path (its tests are algebraic identities only,
/root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/tests.rs:3-51), so these fixtures come from
the repo's own obviously-correct spec; the two public known answers they embed (MiMC7 constants and
hash(1,2) from circomlib) are asserted in tests/test_oracle_spec.py.
Run from the repo root:  python -m tests.golden.gen_golden      (~3 minutes, pure Python Groth16)
"""
import hashlib
import json
import os
import random

from oracle import bn254 as bn
from oracle import groth16 as g16
from oracle import mimc7, ntt
from oracle.withdraw_circuit import build_r1cs, witness

R, P = bn.R, bn.P
HERE = os.path.dirname(os.path.abspath(__file__))


def hx(b):
    return b.hex()


def main():
    rng = random.Random(20260922)
    out = {}
    fe = []
    for name, mod in (("fq", P), ("fr", R)):
        xs = [0, 1, mod - 1, mod - 1, 2**253, rng.randrange(mod), rng.randrange(mod)]
        ys = [mod - 1, mod - 1, mod - 1, 1, 2**253, rng.randrange(mod), rng.randrange(mod)]
        fe.append(dict(field=name, a=[str(x) for x in xs], b=[str(y) for y in ys],
                       mul=[str(a * b % mod) for a, b in zip(xs, ys)], add=[str((a + b) % mod) for a, b in zip(xs, ys)],
                       sub=[str((a - b) % mod) for a, b in zip(xs, ys)]))
    out["field"] = fe
    ks = [0, 1, 2, R - 1, rng.randrange(R)]
    out["g1_mul"] = [dict(k=str(k), out=hx(bn.g1_to_bytes(bn.g1_mul(bn.G1_GEN, k)))) for k in ks]
    out["g2_mul"] = [dict(k=str(k), out=hx(bn.g2_to_bytes(bn.g2_mul(bn.G2_GEN, k)))) for k in ks]
    msm = []
    for n in (1, 2, 33):
        pts = [bn.g1_mul(bn.G1_GEN, rng.randrange(R)) for _ in range(n)]
        sc = [rng.randrange(R) for _ in range(n)]
        if n >= 33:
            pts[5] = pts[4]; pts[7] = None; sc[0] = 0; sc[1] = 1; sc[2] = R - 1
        msm.append(dict(curve="g1", points=hx(b"".join(map(bn.g1_to_bytes, pts))), scalars=hx(b"".join(map(bn.fr_to_bytes, sc))),
                        out=hx(bn.g1_to_bytes(bn.g1_msm(pts, sc)))))
    pts = [bn.g2_mul(bn.G2_GEN, rng.randrange(R)) for _ in range(5)]
    sc = [rng.randrange(R) for _ in range(5)]
    msm.append(dict(curve="g2", points=hx(b"".join(map(bn.g2_to_bytes, pts))), scalars=hx(b"".join(map(bn.fr_to_bytes, sc))),
                    out=hx(bn.g2_to_bytes(bn.g2_msm(pts, sc)))))
    out["msm"] = msm
    nt = []
    for log_n in (1, 2, 10):
        seed = 100 + log_n
        r2 = random.Random(seed)
        v = [r2.randrange(R) for _ in range(1 << log_n)]
        for inv in (False, True):
            for co in (False, True):
                res = b"".join(bn.fr_to_bytes(x) for x in ntt.ntt(v, inv, co))
                nt.append(dict(log_n=log_n, seed=seed, inverse=inv, coset=co, sha256=hashlib.sha256(res).hexdigest(),
                               out=hx(res) if log_n <= 2 else None))
    out["ntt"] = nt
    leaves = [rng.randrange(R) for _ in range(4)]
    t = mimc7.MerkleTree(2)
    for l in leaves:
        t.insert(l)
    sib = [rng.randrange(R) for _ in range(3)]
    out["mimc7"] = dict(
        c1=str(mimc7.CONSTANTS[1]), c90=str(mimc7.CONSTANTS[90]), hash_1_2=str(mimc7.mimc7_hash(1, 2)),
        multi_hash_1_2=str(mimc7.multi_hash([1, 2])), multi_hash_1_2_3_4=str(mimc7.multi_hash([1, 2, 3, 4])),
        tree4_leaves=[str(x) for x in leaves], tree4_root=str(t.root()),
        path=dict(leaf=str(leaves[0]), siblings=[str(x) for x in sib], bits=5,
                  nodes=[str(x) for x in mimc7.merkle_path_nodes(leaves[0], sib, [1, 0, 1])]))
    # one full proof, depth-1 withdraw circuit, everything injected
    depth = 1
    cs = build_r1cs(depth)
    tox = [rng.randrange(1, R) for _ in range(5)]
    pk, vk = g16.setup(cs, *tox)
    nul, sec, rec, sb, r, s = (rng.randrange(R) for _ in range(6))
    w = witness(nul, sec, rec, [sb], [1])
    proof = g16.prove(cs, pk, w, r, s)
    assert g16.verify(vk, w[1:4], proof)
    pk_bytes = (b"".join(map(bn.g1_to_bytes, pk["a"])) + b"".join(map(bn.g1_to_bytes, pk["b1"])) + b"".join(map(bn.g2_to_bytes, pk["b2"]))
                + b"".join(map(bn.g1_to_bytes, pk["l"])) + b"".join(map(bn.g1_to_bytes, pk["h"])))
    out["groth16"] = dict(
        depth=depth, toxic=[str(x) for x in tox], nullifier=str(nul), secret=str(sec), recipient=str(rec), sibling=str(sb), bits=1,
        r=str(r), s=str(s), public=[str(x) for x in w[1:4]], proof=hx(g16.proof_to_bytes(proof)),
        vk=dict(alpha1=hx(bn.g1_to_bytes(vk["alpha1"])), beta2=hx(bn.g2_to_bytes(vk["beta2"])), gamma2=hx(bn.g2_to_bytes(vk["gamma2"])),
                delta2=hx(bn.g2_to_bytes(vk["delta2"])), ic=hx(b"".join(map(bn.g1_to_bytes, vk["ic"])))),
        pk_queries_sha256=hashlib.sha256(pk_bytes).hexdigest(),
        witness_sha256=hashlib.sha256(b"".join(map(bn.fr_to_bytes, w))).hexdigest())
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote vectors.json")


if __name__ == "__main__":
    main()
