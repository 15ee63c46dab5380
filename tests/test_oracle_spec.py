"""The pure-Python spec against public known answers and its own algebraic invariants (CPU only)."""
import random

from oracle import bn254 as bn
from oracle import groth16 as g16
from oracle import mimc7, ntt, pairing
from oracle.keccak import keccak256
from oracle.withdraw_circuit import Layout, build_r1cs, witness, V_NHASH, V_ROOT

R, P = bn.R, bn.P


def test_keccak_known_answers():
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"


def test_mimc7_circomlib_known_answers():
    # first constants listed in circomlib's mimc7.circom and the circomlibjs mimc7 test vector hash(1, 2)
    assert mimc7.CONSTANTS[0] == 0
    assert mimc7.CONSTANTS[1] == 20888961410941983456478427210666206549300505294776164667214940546594746570981
    assert mimc7.CONSTANTS[2] == 15265126113435022738560151911929040668591755459209400716467504685752745317193
    assert len(mimc7.CONSTANTS) == 91
    assert hex(mimc7.mimc7_hash(1, 2)) == "0x176c6eefc3fdf8d6136002d8e6f7a885bbd1c4e3957b93ddc1ec3ae7859f1a08"


def test_reference_field_conventions():
    # /root/reference/.../babyjubjub/mod.rs:7-11: modulus, generator 7, little-endian repr
    assert R == 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert bn.fr_to_bytes(1) == b"\x01" + bytes(31)
    assert pow(bn.FR_GENERATOR, (R - 1) // 2, R) == R - 1          # 7 is a non-residue
    w = bn.root_of_unity(28)
    assert pow(w, 1 << 28, R) == 1 and pow(w, 1 << 27, R) == R - 1
    # BabyJubJub constants (mod.rs:174-189): BASE on a*x^2+y^2 = 1+d*x^2*y^2
    bx = 5299619240641551281634865583518297030282874472190772894086521144482721001553
    by = 16950150798460657717958625567821834550301663161624707787222815936182638968203
    assert (bn.BJJ_A * bx * bx + by * by) % R == (1 + bn.BJJ_D * bx * bx * by * by) % R


def test_curves_and_pairing():
    assert bn.g1_on_curve(bn.G1_GEN) and bn.g2_on_curve(bn.G2_GEN)
    assert bn.g1_mul(bn.G1_GEN, R) is None and bn.g2_mul(bn.G2_GEN, R - 1) == bn.g2_neg(bn.G2_GEN)
    assert bn.g2_add(bn.g2_mul(bn.G2_GEN, R - 1), bn.G2_GEN) is None
    e1 = pairing.pairing(bn.G2_GEN, bn.G1_GEN)
    assert e1 != pairing.F12_ONE and pairing.f12_pow(e1, R) == pairing.F12_ONE
    assert pairing.pairing(bn.g2_mul(bn.G2_GEN, 5), bn.g1_mul(bn.G1_GEN, 7)) == pairing.f12_pow(e1, 35)


def test_ntt_against_naive_dft():
    rng = random.Random(1)
    for log_n in (1, 2, 4):
        v = [rng.randrange(R) for _ in range(1 << log_n)]
        assert ntt.ntt(v) == ntt.dft_naive(v, bn.root_of_unity(log_n))
        assert ntt.ntt(ntt.ntt(v), inverse=True) == v
        assert ntt.ntt(ntt.ntt(v, coset=True), inverse=True, coset=True) == v
        g = bn.root_of_unity(log_n + 1)
        omega = bn.root_of_unity(log_n)
        ev = [sum(c * pow(g * pow(omega, k, R), j, R) for j, c in enumerate(v)) % R for k in range(1 << log_n)]
        assert ntt.ntt(v, coset=True) == ev


def test_merkle_tree_and_paths():
    t = mimc7.MerkleTree(4)
    for i in range(5):
        t.insert(i + 1)
    s, b = t.path(3)
    assert mimc7.merkle_path_nodes(4, s, b)[-1] == t.root()


def test_withdraw_r1cs_and_groth16_small():
    rng = random.Random(5)
    depth, nr = 2, 3
    cs = build_r1cs(depth, nr)
    L = Layout(depth, nr)
    assert cs.n_constraints == L.n_constraints and cs.n_vars == L.n_vars
    sib = [rng.randrange(R) for _ in range(depth)]
    w = witness(11, 22, 33, sib, [1, 0], nr)
    assert cs.is_satisfied(w)
    assert w[V_NHASH] == mimc7.multi_hash([11], 1, nr)
    cm = mimc7.multi_hash([11, 22], 0, nr)
    assert w[V_ROOT] == mimc7.merkle_path_nodes(cm, sib, [1, 0], nr)[-1]
    bad = list(w); bad[V_ROOT] = (bad[V_ROOT] + 1) % R
    assert not cs.is_satisfied(bad)
    pk, vk = g16.setup(cs, *[rng.randrange(1, R) for _ in range(5)])
    proof = g16.prove(cs, pk, w, rng.randrange(R), rng.randrange(R))
    assert g16.verify(vk, w[1:4], proof)
    assert not g16.verify(vk, [w[1], w[2], (w[3] + 1) % R], proof)
    assert g16.proof_from_bytes(g16.proof_to_bytes(proof)) == proof
    w2 = list(w); w2[5] = (w2[5] + 1) % R       # unsatisfying witness -> proof must not verify
    assert not g16.verify(vk, w2[1:4], g16.prove(cs, pk, w2, 1, 2))


def test_full_size_circuit_shape():
    cs = build_r1cs(32)
    assert (cs.n_constraints, cs.n_vars, cs.n_pub) == (24488, 24524, 3)
    assert g16.domain_log(cs.n_constraints, cs.n_pub) == 15
