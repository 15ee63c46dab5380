"""The frozen privacy-pool *withdraw* statement as an R1CS, plus its witness map.

Not defined by the reference: this file IS the spec
that DESIGN.md section 3 describes, the product's C++ builder
(owshen_b200/csrc/withdraw_circuit.cuh) must reproduce it entry for entry.

Statement (public: root, nullifier_hash, recipient):
  I know (nullifier, secret, siblings[depth], bits[depth]) such that
    nullifier_hash = MultiMiMC7([nullifier], key=1)
    commitment     = MultiMiMC7([nullifier, secret], key=0)
    root           = Merkle root reached from commitment along (siblings, bits)
                     with node = MultiMiMC7([left, right], key=0)
  and recipient is bound by recipient^2 = recipient_sq.

Variable layout (index -> meaning), PERM = 4*n_rounds variables (t2,t4,t6,t7 per round):
  0 ONE | 1 root | 2 nullifier_hash | 3 recipient            (public, n_pub = 3)
  4 nullifier | 5 secret | 6 recipient_sq
  7 .. 7+PERM                      nullifier-hash permutation
  then commitment block            perm1[PERM] perm2[PERM] out
  then per level l                 sibling bit left perm1[PERM] perm2[PERM] out
Constraint order: recipient; nullifier-hash perm rounds, its output; commitment
perm1, perm2, output; per level: boolean, select, perm1, perm2, output; root
equality.  Each round is (t)(t)=t2, (t2)(t2)=t4, (t4)(t2)=t6, (t6)(t)=t7.
A linear combination is a dict var -> coefficient (mod r), zero terms dropped;
the CSR export sorts terms by variable index.
"""
from .bn254 import R
from .mimc7 import CONSTANTS, N_ROUNDS

N_PUB = 3
V_ONE, V_ROOT, V_NHASH, V_RECIP, V_NULL, V_SECRET, V_RSQ = range(7)
V_NH_PERM = 7


def lc_add(*lcs):
    out = {}
    for lc in lcs:
        for v, c in lc.items():
            n = (out.get(v, 0) + c) % R
            if n:
                out[v] = n
            else:
                out.pop(v, None)
    return out


def lc_scale(lc, s):
    return {v: c * s % R for v, c in lc.items() if c * s % R}


def lc_eval(lc, w):
    return sum(c * w[v] for v, c in lc.items()) % R


class Layout:
    def __init__(self, depth: int, n_rounds: int = N_ROUNDS):
        self.depth, self.n_rounds = depth, n_rounds
        self.perm = 4 * n_rounds
        self.cm_base = V_NH_PERM + self.perm
        self.cm_out = self.cm_base + 2 * self.perm
        self.lvl_base = self.cm_out + 1
        self.lvl_size = 3 + 2 * self.perm + 1
        self.n_vars = self.lvl_base + depth * self.lvl_size
        per_perm = 4 * n_rounds
        self.n_constraints = 1 + (per_perm + 1) + (2 * per_perm + 1) + depth * (2 + 2 * per_perm + 1) + 1

    def level(self, l):
        b = self.lvl_base + l * self.lvl_size
        return dict(sib=b, bit=b + 1, left=b + 2, perm1=b + 3, perm2=b + 3 + self.perm,
                    out=b + 3 + 2 * self.perm)


class R1CS:
    def __init__(self, n_vars, n_pub):
        self.n_vars, self.n_pub = n_vars, n_pub
        self.A, self.B, self.C = [], [], []

    def add(self, a, b, c):
        self.A.append(a); self.B.append(b); self.C.append(c)

    @property
    def n_constraints(self):
        return len(self.A)

    def is_satisfied(self, w):
        for a, b, c in zip(self.A, self.B, self.C):
            if lc_eval(a, w) * lc_eval(b, w) % R != lc_eval(c, w):
                return False
        return True

    def csr(self, which):
        """(row_ptr, col_idx, coeffs) with terms sorted by variable index."""
        M = {"A": self.A, "B": self.B, "C": self.C}[which]
        ptr, idx, val = [0], [], []
        for lc in M:
            for v in sorted(lc):
                idx.append(v); val.append(lc[v])
            ptr.append(len(idx))
        return ptr, idx, val


def _perm_constraints(cs, x_lc, k_lc, base, n_rounds):
    """MiMC7 permutation rounds; returns the LC of perm(x,k)+k (= hash(x,k))."""
    prev = x_lc
    for i in range(n_rounds):
        t2, t4, t6, t7 = base + 4 * i, base + 4 * i + 1, base + 4 * i + 2, base + 4 * i + 3
        t = lc_add(prev, k_lc, {V_ONE: CONSTANTS[i]})
        cs.add(t, t, {t2: 1})
        cs.add({t2: 1}, {t2: 1}, {t4: 1})
        cs.add({t4: 1}, {t2: 1}, {t6: 1})
        cs.add({t6: 1}, t, {t7: 1})
        prev = {t7: 1}
    return lc_add(prev, k_lc)


def _hash2_constraints(cs, left_lc, right_lc, perm1, perm2, out, n_rounds):
    h1 = _perm_constraints(cs, left_lc, {}, perm1, n_rounds)         # key r0 = 0
    r1 = lc_add(left_lc, h1)                                         # r1 = r0 + x0 + hash(x0, r0)
    h2 = _perm_constraints(cs, right_lc, r1, perm2, n_rounds)
    r2 = lc_add(r1, right_lc, h2)
    cs.add(r2, {V_ONE: 1}, {out: 1})


def build_r1cs(depth: int, n_rounds: int = N_ROUNDS) -> R1CS:
    L = Layout(depth, n_rounds)
    cs = R1CS(L.n_vars, N_PUB)
    cs.add({V_RECIP: 1}, {V_RECIP: 1}, {V_RSQ: 1})
    # nullifier_hash = MultiMiMC7([nullifier], key=1) = 1 + nullifier + hash(nullifier, 1)
    h = _perm_constraints(cs, {V_NULL: 1}, {V_ONE: 1}, V_NH_PERM, n_rounds)
    cs.add(lc_add({V_ONE: 1}, {V_NULL: 1}, h), {V_ONE: 1}, {V_NHASH: 1})
    _hash2_constraints(cs, {V_NULL: 1}, {V_SECRET: 1}, L.cm_base, L.cm_base + L.perm, L.cm_out, n_rounds)
    cur = L.cm_out
    for l in range(depth):
        v = L.level(l)
        cs.add({v["bit"]: 1}, lc_add({v["bit"]: 1}, {V_ONE: R - 1}), {})
        # left = cur + bit*(sib - cur)
        cs.add({v["bit"]: 1}, lc_add({v["sib"]: 1}, {cur: R - 1}), lc_add({v["left"]: 1}, {cur: R - 1}))
        right = lc_add({v["sib"]: 1}, {cur: 1}, {v["left"]: R - 1})
        _hash2_constraints(cs, {v["left"]: 1}, right, v["perm1"], v["perm2"], v["out"], n_rounds)
        cur = v["out"]
    cs.add(lc_add({cur: 1}, {V_ROOT: R - 1}), {V_ONE: 1}, {})
    assert cs.n_constraints == L.n_constraints
    return cs


def _perm_witness(w, x, k, base, n_rounds):
    r = x
    for i in range(n_rounds):
        t = (r + k + CONSTANTS[i]) % R
        t2 = t * t % R; t4 = t2 * t2 % R; t6 = t4 * t2 % R; t7 = t6 * t % R
        w[base + 4 * i: base + 4 * i + 4] = [t2, t4, t6, t7]
        r = t7
    return (r + k) % R


def _hash2_witness(w, left, right, perm1, perm2, out, n_rounds):
    r1 = (left + _perm_witness(w, left, 0, perm1, n_rounds)) % R
    r2 = (r1 + right + _perm_witness(w, right, r1, perm2, n_rounds)) % R
    w[out] = r2
    return r2


def witness(nullifier, secret, recipient, siblings, bits, n_rounds: int = N_ROUNDS):
    """Full assignment (list of n_vars ints).  root / nullifier_hash are derived."""
    depth = len(siblings)
    L = Layout(depth, n_rounds)
    w = [0] * L.n_vars
    w[V_ONE] = 1
    w[V_RECIP] = recipient % R
    w[V_NULL] = nullifier % R
    w[V_SECRET] = secret % R
    w[V_RSQ] = w[V_RECIP] * w[V_RECIP] % R
    w[V_NHASH] = (1 + w[V_NULL] + _perm_witness(w, w[V_NULL], 1, V_NH_PERM, n_rounds)) % R
    cur = _hash2_witness(w, w[V_NULL], w[V_SECRET], L.cm_base, L.cm_base + L.perm, L.cm_out, n_rounds)
    for l in range(depth):
        v = L.level(l)
        sib, bit = siblings[l] % R, bits[l] & 1
        left, right = (sib, cur) if bit else (cur, sib)
        w[v["sib"]], w[v["bit"]], w[v["left"]] = sib, bit, left
        cur = _hash2_witness(w, left, right, v["perm1"], v["perm2"], v["out"], n_rounds)
    w[V_ROOT] = cur
    return w
