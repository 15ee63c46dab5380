"""ctypes binding of oracle/cpu/liboracle.so (the C port of this package's spec).

TEST INFRASTRUCTURE: the checker and the timed CPU baseline, never the product.
Builds the library with ``make -C oracle/cpu`` if it is missing.
"""
import ctypes as C
import os
import subprocess

from . import bn254, mimc7

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu")
_LIB = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


def build():
    subprocess.run(["make", "-C", _DIR, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oc_prover_new.restype = C.c_void_p
        _LIB.oc_withdraw_n_vars.restype = C.c_uint32
        cts = b"".join(bn254.fr_to_bytes(c) for c in mimc7.CONSTANTS)
        assert _LIB.oc_mimc7_set_constants(cts, len(mimc7.CONSTANTS)) == 0
    return _LIB


def set_mimc_rounds(n_rounds):
    cts = b"".join(bn254.fr_to_bytes(c) for c in mimc7.CONSTANTS[:n_rounds])
    assert lib().oc_mimc7_set_constants(cts, n_rounds) == 0


def _buf(n):
    return C.create_string_buffer(n)


def _check(rc):
    if rc != 0:
        raise ValueError(f"oracle C port returned {rc}")


def frs(xs) -> bytes:
    return b"".join(bn254.fr_to_bytes(x) for x in xs)


def fqs(xs) -> bytes:
    return b"".join(bn254.fq_to_bytes(x) for x in xs)


def unfr(b: bytes):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def field_binop(name, a: bytes, b: bytes) -> bytes:
    n = len(a) // 32
    out = _buf(len(a))
    _check(getattr(lib(), name)(a, b, out, C.c_uint64(n)))
    return out.raw


def field_inv(name, a: bytes) -> bytes:
    out = _buf(len(a))
    _check(getattr(lib(), name)(a, out, C.c_uint64(len(a) // 32)))
    return out.raw


def g1_mul(pt: bytes, k: bytes) -> bytes:
    out = _buf(64); _check(lib().oc_g1_mul(pt, k, out)); return out.raw


def g2_mul(pt: bytes, k: bytes) -> bytes:
    out = _buf(128); _check(lib().oc_g2_mul(pt, k, out)); return out.raw


def g1_add(a: bytes, b: bytes) -> bytes:
    out = _buf(64); _check(lib().oc_g1_add(a, b, out)); return out.raw


def g2_add(a: bytes, b: bytes) -> bytes:
    out = _buf(128); _check(lib().oc_g2_add(a, b, out)); return out.raw


def g1_msm(points: bytes, scalars: bytes) -> bytes:
    n = len(scalars) // 32
    assert len(points) == 64 * n
    out = _buf(64); _check(lib().oc_g1_msm(points, scalars, C.c_uint64(n), out)); return out.raw


def g2_msm(points: bytes, scalars: bytes) -> bytes:
    n = len(scalars) // 32
    assert len(points) == 128 * n
    out = _buf(128); _check(lib().oc_g2_msm(points, scalars, C.c_uint64(n), out)); return out.raw


def g1_fixed_mul_batch(base: bytes, scalars: bytes) -> bytes:
    n = len(scalars) // 32
    out = _buf(64 * n); _check(lib().oc_g1_fixed_mul_batch(base, scalars, C.c_uint64(n), out)); return out.raw


def g2_fixed_mul_batch(base: bytes, scalars: bytes) -> bytes:
    n = len(scalars) // 32
    out = _buf(128 * n); _check(lib().oc_g2_fixed_mul_batch(base, scalars, C.c_uint64(n), out)); return out.raw


def ntt(data: bytes, inverse=False, coset=False) -> bytes:
    n = len(data) // 32
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    buf = C.create_string_buffer(data, len(data))
    _check(lib().oc_ntt(buf, log_n, int(inverse), int(coset)))
    return buf.raw


def mimc7_hash(x: int, k: int) -> int:
    out = _buf(32)
    _check(lib().oc_mimc7_hash(bn254.fr_to_bytes(x), bn254.fr_to_bytes(k), out))
    return int.from_bytes(out.raw, "little")


def mimc7_multi_hash(xs, key=0) -> int:
    out = _buf(32)
    _check(lib().oc_mimc7_multi_hash(frs(xs), len(xs), bn254.fr_to_bytes(key), out))
    return int.from_bytes(out.raw, "little")


def merkle_paths(leaves: bytes, siblings: bytes, path_bits, depth: int) -> bytes:
    n = len(leaves) // 32
    bits = (C.c_uint32 * n)(*path_bits)
    out = _buf(n * (depth + 1) * 32)
    _check(lib().oc_mimc7_merkle_paths(leaves, siblings, bits, n, depth, out))
    return out.raw


def withdraw_n_vars(depth: int) -> int:
    return lib().oc_withdraw_n_vars(depth)


def withdraw_witness(nullifiers: bytes, secrets: bytes, recipients: bytes, siblings: bytes, path_bits, depth) -> bytes:
    n = len(nullifiers) // 32
    bits = (C.c_uint32 * n)(*path_bits)
    out = _buf(n * withdraw_n_vars(depth) * 32)
    _check(lib().oc_withdraw_witness(nullifiers, secrets, recipients, siblings, bits, n, depth, out))
    return out.raw


def _u32arr(xs):
    return (C.c_uint32 * len(xs))(*xs)


class Prover:
    """C-port Groth16 prover over an R1CS (CSR from the Python spec) and a proving key in bytes."""

    def __init__(self, cs, pk_bytes: dict):
        from .groth16 import domain_log
        self.n_vars, self.n_pub = cs.n_vars, cs.n_pub
        self.log_m = domain_log(cs.n_constraints, cs.n_pub)
        ap, ai, av = cs.csr("A")
        bp, bi, bv = cs.csr("B")
        p = pk_bytes
        self._h = lib().oc_prover_new(
            cs.n_constraints, cs.n_vars, cs.n_pub, self.log_m,
            _u32arr(ap), _u32arr(ai), frs(av), _u32arr(bp), _u32arr(bi), frs(bv),
            p["alpha1"], p["beta1"], p["beta2"], p["delta1"], p["delta2"],
            p["a"], p["b1"], p["b2"], p["l"], p["h"])
        if not self._h:
            raise ValueError("oc_prover_new failed")

    def h_evals(self, witness: bytes) -> bytes:
        out = _buf(32 << self.log_m)
        _check(lib().oc_prover_h_evals(C.c_void_p(self._h), witness, out))
        return out.raw

    def prove(self, witness: bytes, r: int, s: int) -> bytes:
        out = _buf(256)
        _check(lib().oc_prover_prove(C.c_void_p(self._h), witness, bn254.fr_to_bytes(r), bn254.fr_to_bytes(s), out))
        return out.raw

    def prove_batch(self, witnesses: bytes, rs: bytes) -> bytes:
        n = len(rs) // 64
        out = _buf(256 * n)
        _check(lib().oc_prover_prove_batch(C.c_void_p(self._h), witnesses, rs, n, out))
        return out.raw

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oc_prover_free(C.c_void_p(self._h))
            self._h = None


def setup_bytes(cs, tau, alpha, beta, gamma, delta):
    """pk/vk in boundary bytes, computed with the C fixed-base multiplier from the
    Python spec's exponents (oracle/groth16.py: setup_scalars)."""
    from .groth16 import setup_scalars
    s = setup_scalars(cs, tau, alpha, beta, gamma, delta)
    g1 = bn254.g1_to_bytes(bn254.G1_GEN)
    g2 = bn254.g2_to_bytes(bn254.G2_GEN)
    one = lambda base, k, f: f(base, frs([k]))
    pk = dict(
        log_m=s["log_m"], n_vars=s["n_vars"], n_pub=s["n_pub"],
        alpha1=one(g1, s["alpha"], g1_fixed_mul_batch), beta1=one(g1, s["beta"], g1_fixed_mul_batch),
        beta2=one(g2, s["beta"], g2_fixed_mul_batch), delta1=one(g1, s["delta"], g1_fixed_mul_batch),
        delta2=one(g2, s["delta"], g2_fixed_mul_batch),
        a=g1_fixed_mul_batch(g1, frs(s["a"])), b1=g1_fixed_mul_batch(g1, frs(s["b"])),
        b2=g2_fixed_mul_batch(g2, frs(s["b"])), l=g1_fixed_mul_batch(g1, frs(s["l"])),
        h=g1_fixed_mul_batch(g1, frs(s["h"])),
    )
    vk = dict(alpha1=pk["alpha1"], beta2=pk["beta2"], gamma2=one(g2, s["gamma"], g2_fixed_mul_batch),
              delta2=pk["delta2"], ic=g1_fixed_mul_batch(g1, frs(s["ic"])))
    return pk, vk
