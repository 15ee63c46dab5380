"""BN254 (alt_bn128) fields and groups as pure-Python big-integer arithmetic.

EXTERNAL standard (EIP-196/197), not from the reference, except the scalar
field: modulus, generator 7 and little-endian 32-byte encoding follow
/root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs:7-11.

Everything here is written for obviousness, not speed; it is the slow half of
the two-implementation oracle (the fast half is oracle/cpu/*.c).
"""

# --- fields ---------------------------------------------------------------
# Fr: scalar field, babyjubjub/mod.rs:8
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
# Fq: base field
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR_GENERATOR = 7          # babyjubjub/mod.rs:9
FR_TWO_ADICITY = 28
assert (R - 1) % (1 << FR_TWO_ADICITY) == 0 and ((R - 1) >> FR_TWO_ADICITY) & 1
assert P % 4 == 3


def fr_to_bytes(x: int) -> bytes:
    """Canonical little-endian 32 bytes (babyjubjub/mod.rs:10)."""
    return (x % R).to_bytes(32, "little")


def fq_to_bytes(x: int) -> bytes:
    return (x % P).to_bytes(32, "little")


def fr_from_bytes(b: bytes) -> int:
    x = int.from_bytes(b, "little")
    if x >= R:
        raise ValueError("non-canonical Fr encoding")
    return x


def fq_from_bytes(b: bytes) -> int:
    x = int.from_bytes(b, "little")
    if x >= P:
        raise ValueError("non-canonical Fq encoding")
    return x


def root_of_unity(log_n: int) -> int:
    """Primitive 2^log_n-th root of unity, omega = 7^((r-1)/2^log_n)."""
    assert 0 <= log_n <= FR_TWO_ADICITY
    return pow(FR_GENERATOR, (R - 1) >> log_n, R)


# Fq2 = Fq[i]/(i^2+1), elements are (c0, c1)
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_muls(a, s): return (a[0] * s % P, a[1] * s % P)
def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)
F2_ZERO = (0, 0)
F2_ONE = (1, 0)

# --- curves -----------------------------------------------------------------
# G1: y^2 = x^3 + 3 over Fq, generator (1, 2).  Points: None (infinity) or (x, y).
G1_B = 3
G1_GEN = (1, 2)
# G2: y^2 = x^3 + 3/(9+i) over Fq2
G2_B = f2_mul((3, 0), f2_inv((9, 1)))
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


class _Fq:
    """Uniform field-op table so one set of curve routines serves G1 and G2."""
    zero, one = 0, 1
    add = staticmethod(lambda a, b: (a + b) % P)
    sub = staticmethod(lambda a, b: (a - b) % P)
    mul = staticmethod(lambda a, b: a * b % P)
    neg = staticmethod(lambda a: (-a) % P)
    inv = staticmethod(lambda a: pow(a, -1, P))
    muls = staticmethod(lambda a, s: a * s % P)
    b = G1_B


class _Fq2:
    zero, one = F2_ZERO, F2_ONE
    add = staticmethod(f2_add)
    sub = staticmethod(f2_sub)
    mul = staticmethod(f2_mul)
    neg = staticmethod(f2_neg)
    inv = staticmethod(f2_inv)
    muls = staticmethod(f2_muls)
    b = G2_B


def _on_curve(F, pt):
    if pt is None:
        return True
    x, y = pt
    return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)


def _neg(F, pt):
    return None if pt is None else (pt[0], F.neg(pt[1]))


def _add(F, p, q):
    """Affine addition with every special case (the obviously-correct form)."""
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        lam = F.mul(F.muls(F.mul(x1, x1), 3), F.inv(F.muls(y1, 2)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


# Jacobian internals for scalar multiplication (no inversion per step)
def _jdbl(F, p):
    X, Y, Z = p
    if Z == F.zero:
        return p
    A = F.mul(X, X); B = F.mul(Y, Y); C = F.mul(B, B)
    t = F.add(X, B)
    D = F.muls(F.sub(F.sub(F.mul(t, t), A), C), 2)
    E = F.muls(A, 3); Fv = F.mul(E, E)
    X3 = F.sub(Fv, F.muls(D, 2))
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muls(C, 8))
    Z3 = F.muls(F.mul(Y, Z), 2)
    return (X3, Y3, Z3)


def _jadd_affine(F, p, q):
    """Jacobian p + affine q (q not infinity)."""
    X1, Y1, Z1 = p
    if Z1 == F.zero:
        return (q[0], q[1], F.one)
    Z1Z1 = F.mul(Z1, Z1)
    U2 = F.mul(q[0], Z1Z1)
    S2 = F.mul(F.mul(q[1], Z1), Z1Z1)
    if U2 == X1:
        if S2 == Y1:
            return _jdbl(F, p)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, X1); HH = F.mul(H, H); HHH = F.mul(H, HH)
    r = F.sub(S2, Y1)
    V = F.mul(X1, HH)
    X3 = F.sub(F.sub(F.mul(r, r), HHH), F.muls(V, 2))
    Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.mul(Y1, HHH))
    Z3 = F.mul(Z1, H)
    return (X3, Y3, Z3)


def _jaffine(F, p):
    X, Y, Z = p
    if Z == F.zero:
        return None
    zi = F.inv(Z); zi2 = F.mul(zi, zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


def _mul(F, pt, k):
    if pt is None:
        return None
    k %= R
    acc = (F.one, F.one, F.zero)
    for bit in bin(k)[2:] if k else "":
        acc = _jdbl(F, acc)
        if bit == "1":
            acc = _jadd_affine(F, acc, pt)
    return _jaffine(F, acc)


def g1_on_curve(p): return _on_curve(_Fq, p)
def g1_neg(p): return _neg(_Fq, p)
def g1_add(p, q): return _add(_Fq, p, q)
def g1_mul(p, k): return _mul(_Fq, p, k)
def g2_on_curve(p): return _on_curve(_Fq2, p)
def g2_neg(p): return _neg(_Fq2, p)
def g2_add(p, q): return _add(_Fq2, p, q)
def g2_mul(p, k): return _mul(_Fq2, p, k)


def _msm(F, points, scalars):
    """Naive sum of k_i * P_i -- the definition, used to check every fast MSM."""
    acc = None
    for pt, k in zip(points, scalars):
        acc = _add(F, acc, _mul(F, pt, k))
    return acc


def g1_msm(points, scalars): return _msm(_Fq, points, scalars)
def g2_msm(points, scalars): return _msm(_Fq2, points, scalars)


# --- byte formats at the C-ABI boundary -------------------------------------
# G1 affine: x || y, 32-byte LE each; infinity = 64 zero bytes.
# G2 affine: x.c0 || x.c1 || y.c0 || y.c1; infinity = 128 zero bytes.
def g1_to_bytes(p) -> bytes:
    if p is None:
        return bytes(64)
    return fq_to_bytes(p[0]) + fq_to_bytes(p[1])


def g1_from_bytes(b: bytes):
    assert len(b) == 64
    if b == bytes(64):
        return None
    return (fq_from_bytes(b[:32]), fq_from_bytes(b[32:]))


def g2_to_bytes(p) -> bytes:
    if p is None:
        return bytes(128)
    (x0, x1), (y0, y1) = p
    return fq_to_bytes(x0) + fq_to_bytes(x1) + fq_to_bytes(y0) + fq_to_bytes(y1)


def g2_from_bytes(b: bytes):
    assert len(b) == 128
    if b == bytes(128):
        return None
    v = [fq_from_bytes(b[i:i + 32]) for i in range(0, 128, 32)]
    return ((v[0], v[1]), (v[2], v[3]))


# --- BabyJubJub constants, the reference's only curve (mod.rs:174-189) -------
BJJ_A = 168700
BJJ_D = 168696
