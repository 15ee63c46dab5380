/* oracle/cpu/oracle.c -- the fast half of the CPU oracle and the timed CPU baseline.
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load this library; the product never does.
 *
 * PARITY UNPINNED against the reference: OwshenNetwork/owshen @ c7b1f00 has no
 * Groth16 / MSM / NTT / MiMC code.  This port follows the
 * Python spec in oracle/*.py function for function and is differential-tested
 * against it (tests/test_oracle_*.py); the only reference-fixed convention is the
 * Fr field and its little-endian bytes (babyjubjub/mod.rs:7-11).
 *
 * Every entry point takes canonical little-endian bytes (32 B field elements,
 * G1 = x||y 64 B, G2 = x.c0||x.c1||y.c0||y.c1 128 B, all-zero = infinity) and
 * returns 0 on success, negative on malformed input. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "mont.h"

#define CT_F fe
#define CT_(n) fq_##n
#define CT_P(n) g1_##n
#define CT_BYTES 32
#include "curve_tmpl.h"
#undef CT_F
#undef CT_
#undef CT_P
#undef CT_BYTES

#define CT_F fe2
#define CT_(n) fq2_##n
#define CT_P(n) g2_##n
#define CT_BYTES 64
#include "curve_tmpl.h"
#undef CT_F
#undef CT_
#undef CT_P
#undef CT_BYTES

#define OC_EINVAL (-1)
#define OC_ENOMEM (-2)

int oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- element-wise field ops (differential tests vs python ints) ---------- */
#define VEC_BINOP(name, pfx, op)                                                         \
    int name(const uint8_t *a, const uint8_t *b, uint8_t *out, uint64_t n) {             \
        for (uint64_t i = 0; i < n; i++) {                                               \
            fe x, y, z;                                                                  \
            if (pfx##_from_bytes(&x, a + 32 * i) | pfx##_from_bytes(&y, b + 32 * i)) return OC_EINVAL; \
            pfx##_##op(&z, &x, &y);                                                      \
            pfx##_to_bytes(out + 32 * i, &z);                                            \
        }                                                                                \
        return 0;                                                                        \
    }
VEC_BINOP(oc_fr_mul, fr, mul)
VEC_BINOP(oc_fr_add, fr, add)
VEC_BINOP(oc_fr_sub, fr, sub)
VEC_BINOP(oc_fq_mul, fq, mul)
VEC_BINOP(oc_fq_add, fq, add)
VEC_BINOP(oc_fq_sub, fq, sub)
int oc_fr_inv(const uint8_t *a, uint8_t *out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        fe x, z;
        if (fr_from_bytes(&x, a + 32 * i)) return OC_EINVAL;
        fr_inv(&z, &x);
        fr_to_bytes(out + 32 * i, &z);
    }
    return 0;
}
int oc_fq_inv(const uint8_t *a, uint8_t *out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        fe x, z;
        if (fq_from_bytes(&x, a + 32 * i)) return OC_EINVAL;
        fq_inv(&z, &x);
        fq_to_bytes(out + 32 * i, &z);
    }
    return 0;
}

/* ---- curve entry points ---------------------------------------------------- */
static int scalar_from_bytes(uint64_t s[4], const uint8_t *b) {
    memcpy(s, b, 32);
    return fe_geq(s, FR_MOD) ? OC_EINVAL : 0;
}

#define CURVE_API(G, NB)                                                                        \
    int oc_##G##_mul(const uint8_t *pt, const uint8_t *scalar, uint8_t *out) {                 \
        G##_aff p, r; G##_jac j; uint64_t s[4];                                                \
        if (G##_aff_from_bytes(&p, pt) || scalar_from_bytes(s, scalar)) return OC_EINVAL;      \
        if (p.inf) { memset(out, 0, NB); return 0; }                                           \
        G##_scalar_mul(&j, &p, s);                                                             \
        G##_jac_to_aff(&r, &j);                                                                \
        G##_aff_to_bytes(out, &r);                                                             \
        return 0;                                                                              \
    }                                                                                          \
    int oc_##G##_add(const uint8_t *a, const uint8_t *b, uint8_t *out) {                       \
        G##_aff p, q, r; G##_jac j;                                                            \
        if (G##_aff_from_bytes(&p, a) || G##_aff_from_bytes(&q, b)) return OC_EINVAL;          \
        G##_jac_set_inf(&j);                                                                   \
        G##_jac_madd(&j, &j, &p);                                                              \
        G##_jac_madd(&j, &j, &q);                                                              \
        G##_jac_to_aff(&r, &j);                                                                \
        G##_aff_to_bytes(out, &r);                                                             \
        return 0;                                                                              \
    }                                                                                          \
    int oc_##G##_msm(const uint8_t *pts, const uint8_t *scalars, uint64_t n, uint8_t *out) {   \
        G##_aff *P = (G##_aff *)malloc(sizeof(G##_aff) * (n ? n : 1));                         \
        uint64_t(*S)[4] = (uint64_t(*)[4])malloc(32 * (n ? n : 1));                            \
        if (!P || !S) { free(P); free(S); return OC_ENOMEM; }                                  \
        int bad = 0;                                                                           \
        _Pragma("omp parallel for reduction(|:bad)")                                           \
        for (uint64_t i = 0; i < n; i++)                                                       \
            bad |= G##_aff_from_bytes(&P[i], pts + (uint64_t)NB * i) | scalar_from_bytes(S[i], scalars + 32 * i); \
        if (bad) { free(P); free(S); return OC_EINVAL; }                                       \
        G##_jac j; G##_aff r;                                                                  \
        G##_msm(&j, P, (const uint64_t(*)[4])S, n);                                            \
        G##_jac_to_aff(&r, &j);                                                                \
        G##_aff_to_bytes(out, &r);                                                             \
        free(P); free(S);                                                                      \
        return 0;                                                                              \
    }                                                                                          \
    /* out[i] = scalars[i] * base, 8-bit fixed windows (setup-time helper) */                  \
    int oc_##G##_fixed_mul_batch(const uint8_t *base, const uint8_t *scalars, uint64_t n, uint8_t *out) { \
        G##_aff b;                                                                             \
        if (G##_aff_from_bytes(&b, base)) return OC_EINVAL;                                    \
        if (b.inf) { memset(out, 0, (uint64_t)NB * n); return 0; }                             \
        G##_aff *tab = (G##_aff *)malloc(sizeof(G##_aff) * 32 * 255);                          \
        if (!tab) return OC_ENOMEM;                                                            \
        G##_jac cur; cur.x = b.x; cur.y = b.y; G##_jac_set_inf(&cur);                          \
        G##_jac_madd(&cur, &cur, &b);                                                          \
        for (int w = 0; w < 32; w++) {                                                         \
            G##_aff wb; G##_jac_to_aff(&wb, &cur);                                             \
            G##_jac acc; G##_jac_set_inf(&acc);                                                \
            for (int d = 1; d < 256; d++) {                                                    \
                G##_jac_madd(&acc, &acc, &wb);                                                 \
                G##_jac_to_aff(&tab[w * 255 + d - 1], &acc);                                   \
            }                                                                                  \
            for (int k = 0; k < 8; k++) G##_jac_dbl(&cur, &cur);                               \
        }                                                                                      \
        int bad = 0;                                                                           \
        _Pragma("omp parallel for reduction(|:bad) schedule(static)")                          \
        for (uint64_t i = 0; i < n; i++) {                                                     \
            uint64_t s[4];                                                                     \
            if (scalar_from_bytes(s, scalars + 32 * i)) { bad = 1; continue; }                 \
            G##_jac acc; G##_jac_set_inf(&acc);                                                \
            for (int w = 0; w < 32; w++) {                                                     \
                uint32_t d = (uint32_t)(s[w >> 3] >> ((w & 7) * 8)) & 255;                     \
                if (d) G##_jac_madd(&acc, &acc, &tab[w * 255 + d - 1]);                        \
            }                                                                                  \
            G##_aff r; G##_jac_to_aff(&r, &acc);                                               \
            G##_aff_to_bytes(out + (uint64_t)NB * i, &r);                                      \
        }                                                                                      \
        free(tab);                                                                             \
        return bad ? OC_EINVAL : 0;                                                            \
    }
CURVE_API(g1, 64)
CURVE_API(g2, 128)

int oc_g1_on_curve(const uint8_t *pt) {
    g1_aff p; fe l, r, b;
    if (g1_aff_from_bytes(&p, pt)) return OC_EINVAL;
    if (p.inf) return 1;
    fq_sqr(&l, &p.y); fq_sqr(&r, &p.x); fq_mul(&r, &r, &p.x);
    memcpy(b.l, G1_B_M, 32); fq_add(&r, &r, &b);
    return fe_eq(&l, &r);
}
int oc_g2_on_curve(const uint8_t *pt) {
    g2_aff p; fe2 l, r, b;
    if (g2_aff_from_bytes(&p, pt)) return OC_EINVAL;
    if (p.inf) return 1;
    fq2_sqr(&l, &p.y); fq2_sqr(&r, &p.x); fq2_mul(&r, &r, &p.x);
    memcpy(b.c0.l, G2_B_C0, 32); memcpy(b.c1.l, G2_B_C1, 32); fq2_add(&r, &r, &b);
    return fq2_eq(&l, &r);
}

/* ---- NTT over Fr (oracle/ntt.py) ------------------------------------------- */
static void fr_root_of_unity(fe *w, uint32_t log_n) {
    /* 7^((r-1) >> log_n), generator 7 per babyjubjub/mod.rs:9 */
    uint64_t e[4] = {FR_MOD[0] - 1, FR_MOD[1], FR_MOD[2], FR_MOD[3]};
    for (uint32_t k = 0; k < log_n; k++) {
        for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
        e[3] >>= 1;
    }
    fe seven, r2, t = {{7, 0, 0, 0}};
    memcpy(r2.l, FR_R2, 32);
    fr_mul(&seven, &t, &r2);
    fr_pow(w, &seven, e);
}

/* in-place NTT on Montgomery-form data, natural order in and out */
static void ntt_mont(fe *a, uint32_t log_n, int inverse, int coset) {
    size_t n = (size_t)1 << log_n;
    fe omega, g;
    fr_root_of_unity(&omega, log_n);
    fr_root_of_unity(&g, log_n + 1);
    if (inverse) { fr_inv(&omega, &omega); fr_inv(&g, &g); }
    if (coset && !inverse) {
        fe p; fr_one(&p);
        for (size_t i = 0; i < n; i++) { fr_mul(&a[i], &a[i], &p); fr_mul(&p, &p, &g); }
    }
    for (size_t i = 0; i < n; i++) {              /* bit reversal */
        size_t j = 0;
        for (uint32_t b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < j) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    fe *tw = (fe *)malloc(sizeof(fe) * (n / 2 ? n / 2 : 1));
    fr_one(&tw[0]);
    for (size_t i = 1; i < n / 2; i++) fr_mul(&tw[i], &tw[i - 1], &omega);
    for (uint32_t s = 1; s <= log_n; s++) {
        size_t half = (size_t)1 << (s - 1), step = n >> s;
        #pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t k = 0; k < n / 2; k++) {
            size_t blk = k / half, j = k % half;
            size_t i0 = blk * 2 * half + j, i1 = i0 + half;
            fe t, u = a[i0];
            fr_mul(&t, &a[i1], &tw[j * step]);
            fr_add(&a[i0], &u, &t);
            fr_sub(&a[i1], &u, &t);
        }
    }
    free(tw);
    if (inverse) {
        fe ninv, nn = {{(uint64_t)n, 0, 0, 0}}, r2;
        memcpy(r2.l, FR_R2, 32);
        fr_mul(&nn, &nn, &r2);
        fr_inv(&ninv, &nn);
        fe p = ninv;
        for (size_t i = 0; i < n; i++) {
            fr_mul(&a[i], &a[i], &p);
            if (coset) fr_mul(&p, &p, &g);
        }
    }
}

int oc_ntt(uint8_t *data, uint32_t log_n, int inverse, int coset) {
    if (log_n > 27) return OC_EINVAL;
    size_t n = (size_t)1 << log_n;
    fe *a = (fe *)malloc(sizeof(fe) * n);
    if (!a) return OC_ENOMEM;
    for (size_t i = 0; i < n; i++)
        if (fr_from_bytes(&a[i], data + 32 * i)) { free(a); return OC_EINVAL; }
    ntt_mont(a, log_n, inverse, coset);
    for (size_t i = 0; i < n; i++) fr_to_bytes(data + 32 * i, &a[i]);
    free(a);
    return 0;
}

/* ---- MiMC7 (oracle/mimc7.py) ------------------------------------------------ */
#define MIMC_MAX_ROUNDS 128
static fe g_mimc_c[MIMC_MAX_ROUNDS];
static uint32_t g_mimc_rounds = 0;

/* constants are derived by keccak in oracle/mimc7.py and handed over as bytes */
int oc_mimc7_set_constants(const uint8_t *c, uint32_t n_rounds) {
    if (n_rounds == 0 || n_rounds > MIMC_MAX_ROUNDS) return OC_EINVAL;
    for (uint32_t i = 0; i < n_rounds; i++)
        if (fr_from_bytes(&g_mimc_c[i], c + 32 * i)) return OC_EINVAL;
    g_mimc_rounds = n_rounds;
    return 0;
}

/* hash(x,k) = perm(x,k) + k; optionally records t2,t4,t6,t7 of every round */
static void mimc7_hash_m(fe *out, const fe *x, const fe *k, fe *trace) {
    fe r = *x, t, t2, t4, t6;
    for (uint32_t i = 0; i < g_mimc_rounds; i++) {
        fr_add(&t, &r, k);
        fr_add(&t, &t, &g_mimc_c[i]);
        fr_sqr(&t2, &t);
        fr_sqr(&t4, &t2);
        fr_mul(&t6, &t4, &t2);
        fr_mul(&r, &t6, &t);
        if (trace) { trace[4 * i] = t2; trace[4 * i + 1] = t4; trace[4 * i + 2] = t6; trace[4 * i + 3] = r; }
    }
    fr_add(out, &r, k);
}
static void mimc7_hash2_m(fe *out, const fe *l, const fe *r, fe *trace1, fe *trace2) {
    fe zero, h, r1, r2;
    fr_zero(&zero);
    mimc7_hash_m(&h, l, &zero, trace1);
    fr_add(&r1, l, &h);
    mimc7_hash_m(&h, r, &r1, trace2);
    fr_add(&r2, &r1, r);
    fr_add(out, &r2, &h);
}

int oc_mimc7_hash(const uint8_t *x, const uint8_t *k, uint8_t *out) {
    fe a, b, o;
    if (!g_mimc_rounds || fr_from_bytes(&a, x) || fr_from_bytes(&b, k)) return OC_EINVAL;
    mimc7_hash_m(&o, &a, &b, NULL);
    fr_to_bytes(out, &o);
    return 0;
}
int oc_mimc7_multi_hash(const uint8_t *xs, uint32_t n, const uint8_t *key, uint8_t *out) {
    fe r, x, h;
    if (!g_mimc_rounds || fr_from_bytes(&r, key)) return OC_EINVAL;
    for (uint32_t i = 0; i < n; i++) {
        if (fr_from_bytes(&x, xs + 32 * i)) return OC_EINVAL;
        mimc7_hash_m(&h, &x, &r, NULL);
        fr_add(&r, &r, &x);
        fr_add(&r, &r, &h);
    }
    fr_to_bytes(out, &r);
    return 0;
}

/* out_nodes: n_paths * (depth+1) * 32; path_bits[p] bit l = 1 -> current node is the right child */
int oc_mimc7_merkle_paths(const uint8_t *leaves, const uint8_t *siblings, const uint32_t *path_bits,
                          uint32_t n_paths, uint32_t depth, uint8_t *out_nodes) {
    if (!g_mimc_rounds || depth > 32) return OC_EINVAL;
    int bad = 0;
    #pragma omp parallel for reduction(|:bad) schedule(static)
    for (uint32_t p = 0; p < n_paths; p++) {
        fe cur, sib, nxt;
        uint8_t *o = out_nodes + (uint64_t)p * (depth + 1) * 32;
        if (fr_from_bytes(&cur, leaves + 32ULL * p)) { bad = 1; continue; }
        fr_to_bytes(o, &cur);
        for (uint32_t l = 0; l < depth; l++) {
            if (fr_from_bytes(&sib, siblings + ((uint64_t)p * depth + l) * 32)) { bad = 1; break; }
            if ((path_bits[p] >> l) & 1) mimc7_hash2_m(&nxt, &sib, &cur, NULL, NULL);
            else mimc7_hash2_m(&nxt, &cur, &sib, NULL, NULL);
            cur = nxt;
            fr_to_bytes(o + 32 * (l + 1), &cur);
        }
    }
    return bad ? OC_EINVAL : 0;
}

/* ---- withdraw-circuit witness (oracle/withdraw_circuit.py: witness()) -------- */
uint32_t oc_withdraw_n_vars(uint32_t depth) {
    uint32_t perm = 4 * g_mimc_rounds;
    return 7 + perm + (2 * perm + 1) + depth * (3 + 2 * perm + 1);
}

static void withdraw_witness_m(fe *w, const fe *nullifier, const fe *secret, const fe *recipient,
                               const fe *sibs, uint32_t bits, uint32_t depth) {
    uint32_t perm = 4 * g_mimc_rounds;
    fe one, h, cur;
    fr_one(&one);
    w[0] = one; w[3] = *recipient; w[4] = *nullifier; w[5] = *secret;
    fr_sqr(&w[6], recipient);
    mimc7_hash_m(&h, nullifier, &one, w + 7);
    fr_add(&h, &h, nullifier);
    fr_add(&w[2], &h, &one);
    uint32_t cm = 7 + perm;
    mimc7_hash2_m(&cur, nullifier, secret, w + cm, w + cm + perm);
    w[cm + 2 * perm] = cur;
    uint32_t base = cm + 2 * perm + 1, lsz = 3 + 2 * perm + 1;
    for (uint32_t l = 0; l < depth; l++) {
        fe *v = w + base + l * lsz;
        uint32_t bit = (bits >> l) & 1;
        fe left = bit ? sibs[l] : cur, right = bit ? cur : sibs[l];
        v[0] = sibs[l];
        if (bit) v[1] = one; else fr_zero(&v[1]);
        v[2] = left;
        mimc7_hash2_m(&cur, &left, &right, v + 3, v + 3 + perm);
        v[3 + 2 * perm] = cur;
    }
    w[1] = cur;
}

/* inputs per proof: nullifier, secret, recipient (32 B each), siblings depth*32, path_bits u32 */
int oc_withdraw_witness(const uint8_t *nullifiers, const uint8_t *secrets, const uint8_t *recipients,
                        const uint8_t *siblings, const uint32_t *path_bits, uint32_t n, uint32_t depth,
                        uint8_t *out) {
    if (!g_mimc_rounds || depth > 32) return OC_EINVAL;
    uint32_t nv = oc_withdraw_n_vars(depth);
    int bad = 0;
    #pragma omp parallel for reduction(|:bad) schedule(static)
    for (uint32_t p = 0; p < n; p++) {
        fe nu, se, re, sib[32];
        fe *w = (fe *)malloc(sizeof(fe) * nv);
        int b = fr_from_bytes(&nu, nullifiers + 32ULL * p) | fr_from_bytes(&se, secrets + 32ULL * p) |
                fr_from_bytes(&re, recipients + 32ULL * p);
        for (uint32_t l = 0; l < depth; l++) b |= fr_from_bytes(&sib[l], siblings + ((uint64_t)p * depth + l) * 32);
        if (b) { bad = 1; free(w); continue; }
        withdraw_witness_m(w, &nu, &se, &re, sib, path_bits[p], depth);
        for (uint32_t i = 0; i < nv; i++) fr_to_bytes(out + ((uint64_t)p * nv + i) * 32, &w[i]);
        free(w);
    }
    return bad ? OC_EINVAL : 0;
}

/* ---- Groth16 prover (oracle/groth16.py: prove()) ------------------------------ */
typedef struct {
    uint32_t n_constraints, n_vars, n_pub, log_m;
    uint32_t *a_ptr, *a_idx, *b_ptr, *b_idx;
    fe *a_val, *b_val;
    g1_aff alpha1, beta1, delta1;
    g2_aff beta2, delta2;
    g1_aff *qa, *qb1, *ql, *qh;
    g2_aff *qb2;
} oc_prover;

static void *dup_mem(const void *p, size_t n) { void *r = malloc(n ? n : 1); if (r && n) memcpy(r, p, n); return r; }

void oc_prover_free(void *hp) {
    oc_prover *h = (oc_prover *)hp;
    if (!h) return;
    free(h->a_ptr); free(h->a_idx); free(h->b_ptr); free(h->b_idx); free(h->a_val); free(h->b_val);
    free(h->qa); free(h->qb1); free(h->ql); free(h->qh); free(h->qb2);
    free(h);
}

/* CSR of A and B (coefficients canonical bytes); pk query arrays as affine bytes:
 * qa[n_vars], qb1[n_vars], qb2[n_vars], ql[n_vars-n_pub-1], qh[2^log_m] */
void *oc_prover_new(uint32_t n_constraints, uint32_t n_vars, uint32_t n_pub, uint32_t log_m,
                    const uint32_t *a_ptr, const uint32_t *a_idx, const uint8_t *a_val,
                    const uint32_t *b_ptr, const uint32_t *b_idx, const uint8_t *b_val,
                    const uint8_t *alpha1, const uint8_t *beta1, const uint8_t *beta2,
                    const uint8_t *delta1, const uint8_t *delta2,
                    const uint8_t *qa, const uint8_t *qb1, const uint8_t *qb2,
                    const uint8_t *ql, const uint8_t *qh) {
    oc_prover *h = (oc_prover *)calloc(1, sizeof(oc_prover));
    if (!h) return NULL;
    size_t m = (size_t)1 << log_m, n_priv = n_vars - n_pub - 1;
    if (n_constraints + n_pub + 1 > m) { free(h); return NULL; }
    h->n_constraints = n_constraints; h->n_vars = n_vars; h->n_pub = n_pub; h->log_m = log_m;
    uint32_t annz = a_ptr[n_constraints], bnnz = b_ptr[n_constraints];
    h->a_ptr = dup_mem(a_ptr, 4 * (n_constraints + 1)); h->a_idx = dup_mem(a_idx, 4 * (size_t)annz);
    h->b_ptr = dup_mem(b_ptr, 4 * (n_constraints + 1)); h->b_idx = dup_mem(b_idx, 4 * (size_t)bnnz);
    h->a_val = malloc(sizeof(fe) * (annz + 1)); h->b_val = malloc(sizeof(fe) * (bnnz + 1));
    h->qa = malloc(sizeof(g1_aff) * n_vars); h->qb1 = malloc(sizeof(g1_aff) * n_vars);
    h->qb2 = malloc(sizeof(g2_aff) * n_vars); h->ql = malloc(sizeof(g1_aff) * (n_priv + 1));
    h->qh = malloc(sizeof(g1_aff) * m);
    int bad = 0;
    for (uint32_t i = 0; i < annz; i++) bad |= fr_from_bytes(&h->a_val[i], a_val + 32ULL * i);
    for (uint32_t i = 0; i < bnnz; i++) bad |= fr_from_bytes(&h->b_val[i], b_val + 32ULL * i);
    bad |= g1_aff_from_bytes(&h->alpha1, alpha1) | g1_aff_from_bytes(&h->beta1, beta1) |
           g1_aff_from_bytes(&h->delta1, delta1) | g2_aff_from_bytes(&h->beta2, beta2) |
           g2_aff_from_bytes(&h->delta2, delta2);
    for (uint32_t i = 0; i < n_vars; i++) {
        bad |= g1_aff_from_bytes(&h->qa[i], qa + 64ULL * i) | g1_aff_from_bytes(&h->qb1[i], qb1 + 64ULL * i) |
               g2_aff_from_bytes(&h->qb2[i], qb2 + 128ULL * i);
    }
    for (size_t i = 0; i < n_priv; i++) bad |= g1_aff_from_bytes(&h->ql[i], ql + 64ULL * i);
    for (size_t i = 0; i < m; i++) bad |= g1_aff_from_bytes(&h->qh[i], qh + 64ULL * i);
    if (bad) { oc_prover_free(h); return NULL; }
    return h;
}

/* d_j = (a*b - c)(g w^j), Montgomery form; wit in Montgomery form */
static void h_evals_m(const oc_prover *h, const fe *wit, fe *d) {
    size_t m = (size_t)1 << h->log_m;
    fe *a = (fe *)calloc(m, sizeof(fe)), *b = (fe *)calloc(m, sizeof(fe)), *c = (fe *)calloc(m, sizeof(fe));
    for (uint32_t j = 0; j < h->n_constraints; j++) {
        fe acc, t;
        fr_zero(&acc);
        for (uint32_t k = h->a_ptr[j]; k < h->a_ptr[j + 1]; k++) { fr_mul(&t, &h->a_val[k], &wit[h->a_idx[k]]); fr_add(&acc, &acc, &t); }
        a[j] = acc;
        fr_zero(&acc);
        for (uint32_t k = h->b_ptr[j]; k < h->b_ptr[j + 1]; k++) { fr_mul(&t, &h->b_val[k], &wit[h->b_idx[k]]); fr_add(&acc, &acc, &t); }
        b[j] = acc;
    }
    for (uint32_t i = 0; i <= h->n_pub; i++) a[h->n_constraints + i] = wit[i];
    for (size_t j = 0; j < m; j++) fr_mul(&c[j], &a[j], &b[j]);
    ntt_mont(a, h->log_m, 1, 0); ntt_mont(a, h->log_m, 0, 1);
    ntt_mont(b, h->log_m, 1, 0); ntt_mont(b, h->log_m, 0, 1);
    ntt_mont(c, h->log_m, 1, 0); ntt_mont(c, h->log_m, 0, 1);
    for (size_t j = 0; j < m; j++) { fe t; fr_mul(&t, &a[j], &b[j]); fr_sub(&d[j], &t, &c[j]); }
    free(a); free(b); free(c);
}

static void fr_to_scalar(uint64_t s[4], const fe *a) {
    uint8_t b[32]; fr_to_bytes(b, a); memcpy(s, b, 32);
}

static int prove_one(const oc_prover *h, const uint8_t *witness, const uint8_t *r_b, const uint8_t *s_b, uint8_t *out) {
    size_t m = (size_t)1 << h->log_m, nv = h->n_vars, n_priv = nv - h->n_pub - 1;
    fe *wit = (fe *)malloc(sizeof(fe) * nv), *d = (fe *)malloc(sizeof(fe) * m);
    uint64_t(*ws)[4] = (uint64_t(*)[4])malloc(32 * nv);
    uint64_t(*ds)[4] = (uint64_t(*)[4])malloc(32 * m);
    uint64_t r[4], s[4], rs[4];
    int bad = scalar_from_bytes(r, r_b) | scalar_from_bytes(s, s_b);
    for (size_t i = 0; i < nv; i++) {
        bad |= fr_from_bytes(&wit[i], witness + 32 * i);
        memcpy(ws[i], witness + 32 * i, 32);
    }
    if (bad) { free(wit); free(d); free(ws); free(ds); return OC_EINVAL; }
    h_evals_m(h, wit, d);
    for (size_t j = 0; j < m; j++) fr_to_scalar(ds[j], &d[j]);
    { fe rm, sm, t; fr_from_bytes(&rm, r_b); fr_from_bytes(&sm, s_b); fr_mul(&t, &rm, &sm); fr_to_scalar(rs, &t); }

    g1_jac A, B1, C, T;
    g2_jac B2, T2;
    g1_msm(&A, h->qa, (const uint64_t(*)[4])ws, nv);
    g1_jac_madd(&A, &A, &h->alpha1);
    g1_scalar_mul(&T, &h->delta1, r); g1_jac_add(&A, &A, &T);
    g2_msm(&B2, h->qb2, (const uint64_t(*)[4])ws, nv);
    g2_jac_madd(&B2, &B2, &h->beta2);
    g2_scalar_mul(&T2, &h->delta2, s); g2_jac_add(&B2, &B2, &T2);
    g1_msm(&B1, h->qb1, (const uint64_t(*)[4])ws, nv);
    g1_jac_madd(&B1, &B1, &h->beta1);
    g1_scalar_mul(&T, &h->delta1, s); g1_jac_add(&B1, &B1, &T);
    g1_msm(&C, h->ql, (const uint64_t(*)[4])(ws + h->n_pub + 1), n_priv);
    g1_msm(&T, h->qh, (const uint64_t(*)[4])ds, m); g1_jac_add(&C, &C, &T);
    g1_aff Aa, B1a, Ca; g2_aff B2a;
    g1_jac_to_aff(&Aa, &A); g1_jac_to_aff(&B1a, &B1);
    g1_scalar_mul(&T, &Aa, s); g1_jac_add(&C, &C, &T);
    g1_scalar_mul(&T, &B1a, r); g1_jac_add(&C, &C, &T);
    g1_scalar_mul(&T, &h->delta1, rs);
    fq_neg(&T.y, &T.y); g1_jac_add(&C, &C, &T);
    g1_jac_to_aff(&Ca, &C); g2_jac_to_aff(&B2a, &B2);
    g1_aff_to_bytes(out, &Aa); g2_aff_to_bytes(out + 64, &B2a); g1_aff_to_bytes(out + 192, &Ca);
    free(wit); free(d); free(ws); free(ds);
    return 0;
}

int oc_prover_h_evals(void *hp, const uint8_t *witness, uint8_t *out) {
    oc_prover *h = (oc_prover *)hp;
    size_t m = (size_t)1 << h->log_m;
    fe *wit = (fe *)malloc(sizeof(fe) * h->n_vars), *d = (fe *)malloc(sizeof(fe) * m);
    int bad = 0;
    for (size_t i = 0; i < h->n_vars; i++) bad |= fr_from_bytes(&wit[i], witness + 32 * i);
    if (!bad) { h_evals_m(h, wit, d); for (size_t j = 0; j < m; j++) fr_to_bytes(out + 32 * j, &d[j]); }
    free(wit); free(d);
    return bad ? OC_EINVAL : 0;
}

/* one proof, OpenMP inside the MSMs */
int oc_prover_prove(void *hp, const uint8_t *witness, const uint8_t *r, const uint8_t *s, uint8_t *out256) {
    return prove_one((const oc_prover *)hp, witness, r, s, out256);
}

/* n proofs, one per OpenMP thread (each single-threaded inside): the CPU baseline's shape.
 * witnesses: n * n_vars * 32; rs: n * 64 (r || s); out: n * 256 */
int oc_prover_prove_batch(void *hp, const uint8_t *witnesses, const uint8_t *rs, uint32_t n, uint8_t *out) {
    const oc_prover *h = (const oc_prover *)hp;
    int bad = 0;
    #pragma omp parallel for reduction(|:bad) schedule(dynamic, 1)
    for (uint32_t i = 0; i < n; i++)
        bad |= prove_one(h, witnesses + (uint64_t)i * h->n_vars * 32, rs + 64ULL * i, rs + 64ULL * i + 32,
                         out + 256ULL * i) != 0;
    return bad ? OC_EINVAL : 0;
}
