"""Optimal-ate pairing on BN254, pure Python (slow, obviously-correct form).

EXTERNAL standard (EIP-197 semantics), not from the reference.  Fq12 is held as
Fq[w]/(w^12 - 18 w^6 + 82) (so that w^6 = 9 + i), G2 points are untwisted into
Fq12 and the Miller loop runs over affine Fq12 points.  Used only to check
Groth16 proofs in tests; the product has its own tower-field C++ pairing.
"""
from .bn254 import P, R, G1_GEN, G2_GEN, g1_neg

_MODC = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]   # w^12 = 18 w^6 - 82
ATE_LOOP = 29793968203157093288                    # 6u+2, u = 4965661367192848881
LOG_ATE = 63


def _f12(c):
    return [x % P for x in c]


F12_ONE = _f12([1] + [0] * 11)
F12_ZERO = [0] * 12


def f12_add(a, b): return [(x + y) % P for x, y in zip(a, b)]
def f12_sub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def f12_neg(a): return [(-x) % P for x in a]
def f12_scalar(a, s): return [x * s % P for x in a]


def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):
        top = t[k]
        if top:
            t[k - 6] += 18 * top
            t[k - 12] -= 82 * top
    return [x % P for x in t[:12]]


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_rounded_div(a, b):
    dega, degb = _poly_deg(a), _poly_deg(b)
    temp = list(a)
    o = [0] * len(a)
    binv = pow(b[degb], -1, P)
    for i in range(dega - degb, -1, -1):
        q = temp[degb + i] * binv % P
        o[i] = (o[i] + q) % P
        for c in range(degb + 1):
            temp[c + i] = (temp[c + i] - q * b[c]) % P
    return o[: _poly_deg(o) + 1]


def f12_inv(a):
    """Extended Euclid in Fq[w] against the modulus polynomial."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [c % P for c in _MODC] + [1]
    while _poly_deg(low):
        r = _poly_rounded_div(high, low)
        r += [0] * (13 - len(r))
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % P
                new[i + j] = (new[i + j] - low[i] * r[j]) % P
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, P)
    return [c * inv0 % P for c in lm[:12]]


def f12_pow(a, e):
    out = F12_ONE
    for bit in bin(e)[2:]:
        out = f12_mul(out, out)
        if bit == "1":
            out = f12_mul(out, a)
    return out


def _twist(q):
    """G2 affine over Fq2 -> affine over Fq12 on y^2 = x^3 + 3."""
    (x0, x1), (y0, y1) = q
    xc = [(x0 - 9 * x1) % P, x1]
    yc = [(y0 - 9 * y1) % P, y1]
    nx = _f12([xc[0]] + [0] * 5 + [xc[1]] + [0] * 5)
    ny = _f12([yc[0]] + [0] * 5 + [yc[1]] + [0] * 5)
    w2 = _f12([0, 0, 1] + [0] * 9)
    w3 = _f12([0, 0, 0, 1] + [0] * 8)
    return (f12_mul(nx, w2), f12_mul(ny, w3))


def _cast_g1(p):
    return (_f12([p[0]] + [0] * 11), _f12([p[1]] + [0] * 11))


def _pt_double(p):
    x, y = p
    lam = f12_mul(f12_scalar(f12_mul(x, x), 3), f12_inv(f12_scalar(y, 2)))
    nx = f12_sub(f12_mul(lam, lam), f12_scalar(x, 2))
    ny = f12_sub(f12_mul(lam, f12_sub(x, nx)), y)
    return (nx, ny)


def _pt_add(p, q):
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if y1 == y2:
            return _pt_double(p)
        raise ValueError("unexpected infinity in Miller loop")
    lam = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    nx = f12_sub(f12_sub(f12_mul(lam, lam), x1), x2)
    ny = f12_sub(f12_mul(lam, f12_sub(x1, nx)), y1)
    return (nx, ny)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    if y1 == y2:
        m = f12_mul(f12_scalar(f12_mul(x1, x1), 3), f12_inv(f12_scalar(y1, 2)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    return f12_sub(xt, x1)


def miller_loop(q2, p1):
    """Miller loop only (no final exponentiation). Either argument infinity -> 1."""
    if q2 is None or p1 is None:
        return F12_ONE
    Q = _twist(q2)
    Pt = _cast_g1(p1)
    Rp = Q
    f = F12_ONE
    for i in range(LOG_ATE, -1, -1):
        f = f12_mul(f12_mul(f, f), _linefunc(Rp, Rp, Pt))
        Rp = _pt_double(Rp)
        if ATE_LOOP & (1 << i):
            f = f12_mul(f, _linefunc(Rp, Q, Pt))
            Rp = _pt_add(Rp, Q)
    Q1 = (f12_pow(Q[0], P), f12_pow(Q[1], P))
    nQ2 = (f12_pow(Q1[0], P), f12_neg(f12_pow(Q1[1], P)))
    f = f12_mul(f, _linefunc(Rp, Q1, Pt))
    Rp = _pt_add(Rp, Q1)
    f = f12_mul(f, _linefunc(Rp, nQ2, Pt))
    return f


def final_exponentiate(f):
    return f12_pow(f, (P ** 12 - 1) // R)


def pairing(q2, p1):
    return final_exponentiate(miller_loop(q2, p1))


def pairing_product_is_one(pairs):
    """pairs: [(G1 point, G2 point), ...]; True iff prod e(P_i, Q_i) == 1."""
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(q2, p1))
    return final_exponentiate(f) == F12_ONE
