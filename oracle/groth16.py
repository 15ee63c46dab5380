"""Groth16 over BN254, pure Python: setup with injected toxic waste, prove with
injected (r, s), verify via the pairing.  EXTERNAL standard (Groth, EUROCRYPT
2016), not from the reference; conventions frozen in DESIGN.md section 4:

 * domain size m = 2^k >= n_constraints + n_pub + 1; rows n_constraints+i
   (i = 0..n_pub) of A hold the input-consistency constraint x_i * 0 = 0;
 * on the domain c_j := a_j * b_j (the witness is assumed satisfying);
 * h is carried in the Lagrange basis of the coset g*D, g = omega_{2m}:
   H_query[j] = [ L_j(tau/g) * Z(tau) / (-2 delta) ]_1, so the prover computes
   d_j = (a*b - c)(g omega^j) with 3 iNTT + 3 coset NTT and one size-m MSM;
 * proof bytes = A (G1, 64) || B (G2, 128) || C (G1, 64), affine, 32-byte LE.
"""
from .bn254 import (R, G1_GEN, G2_GEN, g1_add, g1_mul, g1_neg, g2_add, g2_mul, g1_msm, g2_msm,
                    g1_to_bytes, g2_to_bytes, g1_from_bytes, g2_from_bytes, root_of_unity)
from .ntt import ntt
from .pairing import pairing_product_is_one


def domain_log(n_constraints, n_pub):
    need = n_constraints + n_pub + 1
    k = 0
    while (1 << k) < need:
        k += 1
    return k


def lagrange_at(tau, log_m, shift=1):
    """[L_j(tau/shift)] for the size-2^log_m domain, closed form."""
    m = 1 << log_m
    omega = root_of_unity(log_m)
    x = tau * pow(shift, -1, R) % R
    zx = (pow(x, m, R) - 1) * pow(m, -1, R) % R
    out, wj = [], 1
    for _ in range(m):
        out.append(zx * wj % R * pow((x - wj) % R, -1, R) % R)
        wj = wj * omega % R
    return out


def qap_at_tau(cs, tau, log_m):
    """u_i(tau), v_i(tau), w_i(tau) for every variable i."""
    Lg = lagrange_at(tau, log_m)
    u = [0] * cs.n_vars; v = [0] * cs.n_vars; w = [0] * cs.n_vars
    for j in range(cs.n_constraints):
        for i, c in cs.A[j].items(): u[i] = (u[i] + c * Lg[j]) % R
        for i, c in cs.B[j].items(): v[i] = (v[i] + c * Lg[j]) % R
        for i, c in cs.C[j].items(): w[i] = (w[i] + c * Lg[j]) % R
    for i in range(cs.n_pub + 1):
        u[i] = (u[i] + Lg[cs.n_constraints + i]) % R
    return u, v, w


def setup_scalars(cs, tau, alpha, beta, gamma, delta):
    """Every pk/vk element as an exponent of the generators (so a fast
    fixed-base multiplier -- oracle/cpu or the GPU -- can be checked against it)."""
    log_m = domain_log(cs.n_constraints, cs.n_pub)
    m = 1 << log_m
    u, v, w = qap_at_tau(cs, tau, log_m)
    dinv, ginv = pow(delta, -1, R), pow(gamma, -1, R)
    k = [(beta * u[i] + alpha * v[i] + w[i]) % R for i in range(cs.n_vars)]
    g = root_of_unity(log_m + 1)
    zt = (pow(tau, m, R) - 1) % R
    hfac = zt * pow((-2 * delta) % R, -1, R) % R
    Lc = lagrange_at(tau, log_m, shift=g)
    return dict(
        log_m=log_m, n_vars=cs.n_vars, n_pub=cs.n_pub,
        alpha=alpha % R, beta=beta % R, gamma=gamma % R, delta=delta % R,
        a=u, b=v,
        l=[k[i] * dinv % R for i in range(cs.n_pub + 1, cs.n_vars)],
        ic=[k[i] * ginv % R for i in range(cs.n_pub + 1)],
        h=[x * hfac % R for x in Lc],
    )


def setup(cs, tau, alpha, beta, gamma, delta):
    s = setup_scalars(cs, tau, alpha, beta, gamma, delta)
    G1 = lambda k: g1_mul(G1_GEN, k)
    G2 = lambda k: g2_mul(G2_GEN, k)
    pk = dict(
        log_m=s["log_m"], n_vars=s["n_vars"], n_pub=s["n_pub"],
        alpha1=G1(s["alpha"]), beta1=G1(s["beta"]), beta2=G2(s["beta"]),
        delta1=G1(s["delta"]), delta2=G2(s["delta"]),
        a=[G1(x) for x in s["a"]], b1=[G1(x) for x in s["b"]], b2=[G2(x) for x in s["b"]],
        l=[G1(x) for x in s["l"]], h=[G1(x) for x in s["h"]],
    )
    vk = dict(alpha1=pk["alpha1"], beta2=pk["beta2"], gamma2=G2(s["gamma"]), delta2=pk["delta2"],
              ic=[G1(x) for x in s["ic"]])
    return pk, vk


def abc_evals(cs, wit, log_m):
    m = 1 << log_m
    a = [0] * m; b = [0] * m
    for j in range(cs.n_constraints):
        a[j] = sum(c * wit[i] for i, c in cs.A[j].items()) % R
        b[j] = sum(c * wit[i] for i, c in cs.B[j].items()) % R
    for i in range(cs.n_pub + 1):
        a[cs.n_constraints + i] = wit[i]
    c = [x * y % R for x, y in zip(a, b)]
    return a, b, c


def h_evals(cs, wit, log_m):
    """d_j = (a*b - c)(g omega^j): the scalars of the H-query MSM."""
    a, b, c = abc_evals(cs, wit, log_m)
    ac, bc, cc = (ntt(ntt(x, inverse=True), coset=True) for x in (a, b, c))
    return [(x * y - z) % R for x, y, z in zip(ac, bc, cc)]


def prove(cs, pk, wit, r, s):
    d = h_evals(cs, wit, pk["log_m"])
    A = g1_add(g1_add(pk["alpha1"], g1_msm(pk["a"], wit)), g1_mul(pk["delta1"], r))
    B2 = g2_add(g2_add(pk["beta2"], g2_msm(pk["b2"], wit)), g2_mul(pk["delta2"], s))
    B1 = g1_add(g1_add(pk["beta1"], g1_msm(pk["b1"], wit)), g1_mul(pk["delta1"], s))
    C = g1_msm(pk["l"], wit[cs.n_pub + 1:])
    C = g1_add(C, g1_msm(pk["h"], d))
    C = g1_add(C, g1_mul(A, s))
    C = g1_add(C, g1_mul(B1, r))
    C = g1_add(C, g1_neg(g1_mul(pk["delta1"], r * s % R)))
    return (A, B2, C)


def proof_to_bytes(proof) -> bytes:
    A, B, C = proof
    return g1_to_bytes(A) + g2_to_bytes(B) + g1_to_bytes(C)


def proof_from_bytes(b: bytes):
    assert len(b) == 256
    return (g1_from_bytes(b[:64]), g2_from_bytes(b[64:192]), g1_from_bytes(b[192:]))


def verify(vk, public_inputs, proof) -> bool:
    A, B, C = proof
    if A is None or B is None or C is None:
        return False
    assert len(public_inputs) + 1 == len(vk["ic"])
    acc = vk["ic"][0]
    for x, pt in zip(public_inputs, vk["ic"][1:]):
        acc = g1_add(acc, g1_mul(pt, x))
    return pairing_product_is_one([
        (g1_neg(A), B), (vk["alpha1"], vk["beta2"]), (acc, vk["gamma2"]), (C, vk["delta2"])])
