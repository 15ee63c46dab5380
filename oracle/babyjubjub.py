"""BabyJubJub EdDSA-style signatures -- the ONE piece of this repo whose algorithm the reference defines:
a restatement of /root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs, function by function.
(SURVEY.md section 8f.3 / row a7.  Dead code at runtime in the reference, but it is its only big-integer path.)

Parity status: the reference's own tests (babyjubjub/tests.rs:3-51) are algebraic identities, not byte
vectors, and no Rust toolchain exists here, so parity is pinned by re-running those identities against this
restatement (tests/test_babyjubjub.py) -- not by reference outputs.

Points are (x, y) tuples over Fr; the curve is a x^2 + y^2 = 1 + d x^2 y^2 with A = 168700, D = 168696
(mod.rs:174-176).  `hash` is the reference's placeholder product (mod.rs:202-204); `hash_kind=1` swaps in
MultiMiMC7 (the "real hash" SURVEY 8f.3 asks for) and is NOT reference behaviour.
"""
from .bn254 import R
from . import mimc7

A = 168700                       # mod.rs:175
D = 168696                       # mod.rs:176
BASE = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
        16950150798460657717958625567821834550301663161624707787222815936182638968203)   # mod.rs:177-183
ORDER = 21888242871839275222246405745257275088614511777268538073601725287587578984328      # mod.rs:185-188
ZERO = (0, 1)                    # PointAffine::zero(), mod.rs:53-55


class CannotInvert(Exception):
    """anyhow!("Cannot invert") / ("Cannot take sqrt") in the reference"""


def _inv(x):
    if x % R == 0:
        raise CannotInvert("Cannot invert")
    return pow(x, -1, R)


def is_on_curve(p):              # mod.rs:47-49
    x, y = p
    return (y * y + A * x * x) % R == (1 + D * x * x % R * y * y) % R


def double(p):                   # mod.rs:56-67
    x, y = p
    xx = _inv((A * x * x + y * y) % R)
    yy = _inv((2 - A * x * x - y * y) % R)
    return (2 * x * y * xx % R, (y * y - A * x * x) * yy % R)


def add(p, q):                   # mod.rs:28-43 (equal points fall through to double, as there)
    if p == q:
        return double(p)
    x1, y1 = p
    x2, y2 = q
    t = D * x1 * x2 % R * y1 * y2 % R
    xx = _inv((1 + t) % R)
    yy = _inv((1 - t) % R)
    return ((x1 * y2 + y1 * x2) * xx % R, (y1 * y2 - A * x1 * x2) * yy % R)


# projective (mod.rs:117-172): Z == 0 is the "empty accumulator" sentinel, not a curve point
P_ZERO = (0, 1, 0)


def p_double(p):                 # mod.rs:152-164
    X, Y, Z = p
    if Z == 0:
        return P_ZERO
    b = (X + Y) ** 2 % R; c = X * X % R; d = Y * Y % R
    e = A * c % R; f = (e + d) % R; h = Z * Z % R
    j = (f - 2 * h) % R
    return ((b - c - d) * j % R, f * (e - d) % R, f * j % R)


def p_to_affine(p):              # mod.rs:165-171
    X, Y, Z = p
    if Z == 0:
        return ZERO
    zi = _inv(Z)
    return (X * zi % R, Y * zi % R)


def p_add(p, q):                 # mod.rs:118-140
    if p[2] == 0:
        return q
    if q[2] == 0:
        return p
    if p_to_affine(p) == p_to_affine(q):
        return p_double(p)
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    a = Z1 * Z2 % R; b = a * a % R; c = X1 * X2 % R; d = Y1 * Y2 % R
    e = D * c % R * d % R; f = (b - e) % R; g = (b + e) % R
    return (a * f % R * ((X1 + Y1) * (X2 + Y2) - c - d) % R, a * g % R * (d - A * c) % R, f * g % R)


def multiply(p, k):              # mod.rs:68-78: MSB-first double-and-add over the 256 bits of the LE repr
    k %= R
    acc = P_ZERO
    pp = (p[0], p[1], 1)
    for i in range(255, -1, -1):
        acc = p_double(acc)
        if (k >> i) & 1:
            acc = p_add(acc, pp)
    return p_to_affine(acc)


def compress(p):                 # mod.rs:82-84
    return (p[0], p[1] & 1)


def _sqrt(a):
    """Tonelli-Shanks in Fr (2-adicity 28); either root is fine, decompress fixes the parity."""
    a %= R
    if a == 0:
        return 0
    if pow(a, (R - 1) // 2, R) != 1:
        raise CannotInvert("Cannot take sqrt")
    s, t = 28, (R - 1) >> 28
    z = pow(7, t, R)
    x = pow(a, (t + 1) // 2, R); b = pow(a, t, R); m = s
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % R; i += 1
        w = pow(z, 1 << (m - i - 1), R)
        x = x * w % R; z = w * w % R; b = b * z % R; m = i
    return x


def decompress(c):               # mod.rs:88-98
    x, odd = c
    inv = _inv((1 - D * x * x) % R)
    y = _sqrt(inv * (1 - A * x * x) % R)
    if (y & 1) != odd:
        y = (-y) % R
    return (x, y)


def hash_placeholder(inp):       # mod.rs:202-204: product of the inputs
    out = 1
    for v in inp:
        out = out * v % R
    return out


def _hash(inp, hash_kind):
    return hash_placeholder(inp) if hash_kind == 0 else mimc7.multi_hash(inp, 0)


def to_pub(sk):                  # mod.rs:207-209
    return compress(multiply(BASE, sk))


def sign(sk, randomness, message, hash_kind=0):      # mod.rs:210-237
    pk = decompress(to_pub(sk))
    r = _hash([randomness, message], hash_kind)
    rr = multiply(BASE, r)
    h = _hash([rr[0], rr[1], pk[0], pk[1], message], hash_kind)
    s = (r + h * sk) % ORDER
    if s >= R:
        raise ValueError("Invalid repr")             # Fp::from_repr rejects s >= r (ORDER > r)
    return (rr, s)


def verify(pk_compressed, message, sig, hash_kind=0):   # mod.rs:99-115
    pk = decompress(pk_compressed)
    rr, s = sig
    if not is_on_curve(pk) or not is_on_curve(rr):
        return False
    h = _hash([rr[0], rr[1], pk[0], pk[1], message], hash_kind)
    sb = multiply(BASE, s)
    r_plus_ha = add(multiply(pk, h), rr)
    return r_plus_ha == sb
