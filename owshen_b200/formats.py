"""Proof / verifying-key encodings for consumers outside this library: the EVM
precompile layout of EIP-197 (what a Solidity Groth16 verifier next to
/root/reference/contracts/src/Owshen.sol:66-78 would take as calldata) and snarkjs-style JSON.
Pure byte shuffling of the library's own 256-byte proofs and OGVK blobs -- no arithmetic happens here.

Library layout (include/owshen_b200.h): little-endian 32-byte coordinates, G2 = x.c0 || x.c1 || y.c0 || y.c1.
EIP-197 / Solidity: big-endian 32-byte words, G2 = x.c1 || x.c0 || y.c1 || y.c0 (imaginary part first)."""
import json
import struct


def _ints(b: bytes):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def _be(x: int) -> bytes:
    return x.to_bytes(32, "big")


def proof_points(proof: bytes):
    """(A, B, C) as integer coordinates: A = (x, y), B = ((x0, x1), (y0, y1)), C = (x, y)."""
    assert len(proof) == 256
    v = _ints(proof)
    return (v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7])


def proof_to_eip197(proof: bytes) -> bytes:
    """a[2] || b[2][2] || c[2] as 8 big-endian words, the argument order of the usual Solidity verifyProof."""
    (ax, ay), ((bx0, bx1), (by0, by1)), (cx, cy) = proof_points(proof)
    return b"".join(_be(x) for x in (ax, ay, bx1, bx0, by1, by0, cx, cy))


def proof_from_eip197(data: bytes) -> bytes:
    assert len(data) == 256
    w = [int.from_bytes(data[i:i + 32], "big") for i in range(0, 256, 32)]
    ax, ay, bx1, bx0, by1, by0, cx, cy = w
    return b"".join(x.to_bytes(32, "little") for x in (ax, ay, bx0, bx1, by0, by1, cx, cy))


def public_inputs_to_eip197(public_inputs: bytes) -> bytes:
    return b"".join(_be(x) for x in _ints(public_inputs))


def proof_to_snarkjs(proof: bytes) -> dict:
    (ax, ay), ((bx0, bx1), (by0, by1)), (cx, cy) = proof_points(proof)
    s = str
    return {"pi_a": [s(ax), s(ay), "1"], "pi_b": [[s(bx0), s(bx1)], [s(by0), s(by1)], ["1", "0"]],
            "pi_c": [s(cx), s(cy), "1"], "protocol": "groth16", "curve": "bn128"}


def parse_vk(vk: bytes) -> dict:
    """OGVK v1 blob -> integer coordinates."""
    assert vk[:4] == b"OGVK"
    ver, n_pub = struct.unpack("<II", vk[4:12])
    assert ver == 1 and len(vk) == 12 + 64 + 3 * 128 + 64 * (n_pub + 1)
    v = _ints(vk[12:])
    g2 = lambda o: ((v[o], v[o + 1]), (v[o + 2], v[o + 3]))
    ic = [(v[14 + 2 * i], v[15 + 2 * i]) for i in range(n_pub + 1)]
    return {"n_pub": n_pub, "alpha1": (v[0], v[1]), "beta2": g2(2), "gamma2": g2(6), "delta2": g2(10), "ic": ic}


def vk_to_snarkjs(vk: bytes) -> dict:
    p = parse_vk(vk)
    s = str
    g1 = lambda q: [s(q[0]), s(q[1]), "1"]
    g2 = lambda q: [[s(q[0][0]), s(q[0][1])], [s(q[1][0]), s(q[1][1])], ["1", "0"]]
    return {"protocol": "groth16", "curve": "bn128", "nPublic": p["n_pub"], "vk_alpha_1": g1(p["alpha1"]),
            "vk_beta_2": g2(p["beta2"]), "vk_gamma_2": g2(p["gamma2"]), "vk_delta_2": g2(p["delta2"]),
            "IC": [g1(q) for q in p["ic"]]}


def to_json(obj: dict) -> str:
    return json.dumps(obj, indent=1)
