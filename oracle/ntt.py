"""Number-theoretic transform over BN254 Fr (pure Python).

Convention (follows the reference's generator 7, babyjubjub/mod.rs:9):
omega_n = 7^((r-1)/n); forward out[k] = sum_j in[j] * omega^(j*k), natural
order in and out; inverse divides by n.  ``coset`` evaluates on g*omega^k with
g = omega_{2n} (the shift the Groth16 prover uses, g^n = -1).
"""
from .bn254 import R, root_of_unity


def dft_naive(vals, omega):
    n = len(vals)
    return [sum(v * pow(omega, j * k, R) for j, v in enumerate(vals)) % R for k in range(n)]


def _fft(vals, omega):
    n = len(vals)
    if n == 1:
        return list(vals)
    ev = _fft(vals[0::2], omega * omega % R)
    od = _fft(vals[1::2], omega * omega % R)
    out = [0] * n
    w = 1
    for k in range(n // 2):
        t = w * od[k] % R
        out[k] = (ev[k] + t) % R
        out[k + n // 2] = (ev[k] - t) % R
        w = w * omega % R
    return out


def ntt(vals, inverse=False, coset=False):
    n = len(vals)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    omega = root_of_unity(log_n)
    g = root_of_unity(log_n + 1)
    if not inverse:
        v = list(vals)
        if coset:
            v = [x * pow(g, j, R) % R for j, x in enumerate(v)]
        return _fft(v, omega)
    out = _fft(vals, pow(omega, -1, R))
    ninv = pow(n, -1, R)
    out = [x * ninv % R for x in out]
    if coset:
        ginv = pow(g, -1, R)
        out = [x * pow(ginv, j, R) % R for j, x in enumerate(out)]
    return out
