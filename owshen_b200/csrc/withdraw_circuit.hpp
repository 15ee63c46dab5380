// owshen_b200/csrc/withdraw_circuit.hpp -- host-side builder of the withdraw statement's R1CS
// (DESIGN.md section 3).  The reference defines no circuit; this is the
// product's own definition and tests/test_circuit_parity.py checks it entry for entry against the
// independently written oracle/withdraw_circuit.py through og_withdraw_r1cs_export.
//
//   public : root, nullifier_hash, recipient
//   private: nullifier, secret, siblings[depth], bits[depth]
//   nullifier_hash = MultiMiMC7([nullifier], key 1); commitment = MultiMiMC7([nullifier, secret], 0);
//   root = Merkle root from commitment with node = MultiMiMC7([left, right], 0); recipient^2 bound.
#pragma once
#include <map>
#include <vector>
#include "host_math.hpp"
#include "mimc.cuh"

namespace og {

typedef std::map<uint32_t, Fr> LC;   // variable -> coefficient (Montgomery), zero terms erased, sorted by variable

inline void lc_add_term(LC& lc, uint32_t var, const Fr& coeff) {
    auto it = lc.find(var);
    if (it == lc.end()) { if (!coeff.is_zero()) lc[var] = coeff; return; }
    Fr s = it->second + coeff;
    if (s.is_zero()) lc.erase(it); else it->second = s;
}
inline LC lc_sum(std::initializer_list<const LC*> parts) {
    LC out;
    for (const LC* p : parts) for (auto& kv : *p) lc_add_term(out, kv.first, kv.second);
    return out;
}
inline LC lc_var(uint32_t v) { LC l; l[v] = Fr::one(); return l; }
inline LC lc_neg_var(uint32_t v) { LC l; l[v] = Fr::one().neg(); return l; }

struct Csr {
    std::vector<uint32_t> row_ptr{0}, col;
    std::vector<Fr> val;
    void push_row(const LC& lc) {
        for (auto& kv : lc) { col.push_back(kv.first); val.push_back(kv.second); }
        row_ptr.push_back((uint32_t)col.size());
    }
    uint32_t rows() const { return (uint32_t)row_ptr.size() - 1; }
};

struct R1cs {
    uint32_t n_vars = 0, n_pub = 0;
    Csr A, B, C;
    void add(const LC& a, const LC& b, const LC& c) { A.push_row(a); B.push_row(b); C.push_row(c); }
    uint32_t n_constraints() const { return A.rows(); }
};

struct WithdrawBuilder {
    R1cs cs;
    Fr consts[MIMC_ROUNDS];
    uint32_t n_rounds;

    // MiMC7 permutation rounds over x with key k; returns the LC of hash(x, k) = perm + k
    LC perm(const LC& x, const LC& k, uint32_t base) {
        LC prev = x;
        for (uint32_t i = 0; i < n_rounds; i++) {
            uint32_t t2 = base + 4 * i, t4 = t2 + 1, t6 = t2 + 2, t7 = t2 + 3;
            LC c; if (!consts[i].is_zero()) c[0] = consts[i];
            LC t = lc_sum({&prev, &k, &c});
            cs.add(t, t, lc_var(t2));
            cs.add(lc_var(t2), lc_var(t2), lc_var(t4));
            cs.add(lc_var(t4), lc_var(t2), lc_var(t6));
            cs.add(lc_var(t6), t, lc_var(t7));
            prev = lc_var(t7);
        }
        return lc_sum({&prev, &k});
    }
    void hash2(const LC& left, const LC& right, uint32_t perm1, uint32_t perm2, uint32_t out) {
        LC zero;
        LC h1 = perm(left, zero, perm1);
        LC r1 = lc_sum({&left, &h1});
        LC h2 = perm(right, r1, perm2);
        LC r2 = lc_sum({&r1, &right, &h2});
        cs.add(r2, lc_var(0), lc_var(out));
    }

    static R1cs build(uint32_t depth, uint32_t n_rounds = MIMC_ROUNDS) {
        WithdrawBuilder b;
        b.n_rounds = n_rounds;
        mimc7_round_constants(b.consts);
        WithdrawLayout L = WithdrawLayout::make(depth, n_rounds);
        b.cs.n_vars = L.n_vars;
        b.cs.n_pub = WITHDRAW_N_PUB;
        const uint32_t V_ONE = 0, V_ROOT = 1, V_NHASH = 2, V_RECIP = 3, V_NULL = 4, V_SECRET = 5, V_RSQ = 6, V_NH_PERM = 7;
        b.cs.add(lc_var(V_RECIP), lc_var(V_RECIP), lc_var(V_RSQ));
        {
            LC h = b.perm(lc_var(V_NULL), lc_var(V_ONE), V_NH_PERM);
            LC one = lc_var(V_ONE), nu = lc_var(V_NULL);
            b.cs.add(lc_sum({&one, &nu, &h}), lc_var(V_ONE), lc_var(V_NHASH));
        }
        b.hash2(lc_var(V_NULL), lc_var(V_SECRET), L.cm_base, L.cm_base + L.perm, L.cm_out);
        uint32_t cur = L.cm_out;
        for (uint32_t l = 0; l < depth; l++) {
            uint32_t base = L.lvl_base + l * L.lvl_size;
            uint32_t sib = base, bit = base + 1, left = base + 2, p1 = base + 3, p2 = base + 3 + L.perm, out = base + 3 + 2 * L.perm;
            LC vbit = lc_var(bit), m1 = lc_neg_var(V_ONE), vsib = lc_var(sib), ncur = lc_neg_var(cur), vleft = lc_var(left), vcur = lc_var(cur);
            LC nleft = lc_neg_var(left);
            b.cs.add(vbit, lc_sum({&vbit, &m1}), LC());
            b.cs.add(vbit, lc_sum({&vsib, &ncur}), lc_sum({&vleft, &ncur}));
            LC right = lc_sum({&vsib, &vcur, &nleft});
            b.hash2(vleft, right, p1, p2, out);
            cur = out;
        }
        LC vcur = lc_var(cur), nroot = lc_neg_var(V_ROOT);
        b.cs.add(lc_sum({&vcur, &nroot}), lc_var(V_ONE), LC());
        return b.cs;
    }
};

inline uint32_t groth16_domain_log(uint32_t n_constraints, uint32_t n_pub) {
    uint32_t need = n_constraints + n_pub + 1, k = 0;
    while ((1u << k) < need) k++;
    return k;
}

}  // namespace og
