/* oracle/cpu/mont.h -- 4x64-bit Montgomery arithmetic for the BN254 base and
 * scalar fields.  TEST INFRASTRUCTURE (CPU oracle / CPU baseline), never linked
 * into the product.  Field definition follows the reference only for Fr
 * (src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs:7-11: modulus, [u64;4]
 * Montgomery limbs, little-endian repr); Fq is the public alt_bn128 base field.
 * PARITY UNPINNED: the reference has no Groth16 path (see oracle/__init__.py). */
#ifndef ORACLE_MONT_H
#define ORACLE_MONT_H
#include <stdint.h>
#include <string.h>
#include "constants.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

static inline int fe_geq(const uint64_t a[4], const uint64_t m[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > m[i]) return 1;
        if (a[i] < m[i]) return 0;
    }
    return 1;
}
static inline void fe_sub_raw(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - (uint64_t)br;
        r[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static inline uint64_t fe_add_raw(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }

static inline void mont_mul(fe *r, const fe *a, const fe *b, const uint64_t m[4], uint64_t inv) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * inv;
        c = (u128)q * m[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)q * m[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || fe_geq(t, m)) fe_sub_raw(r->l, t, m);
    else memcpy(r->l, t, 32);
}
static inline void mont_add(fe *r, const fe *a, const fe *b, const uint64_t m[4]) {
    uint64_t t[4];
    uint64_t c = fe_add_raw(t, a->l, b->l);
    if (c || fe_geq(t, m)) fe_sub_raw(r->l, t, m);
    else memcpy(r->l, t, 32);
}
static inline void mont_sub(fe *r, const fe *a, const fe *b, const uint64_t m[4]) {
    uint64_t t[4];
    if (fe_geq(a->l, b->l)) fe_sub_raw(r->l, a->l, b->l);
    else { fe_add_raw(t, a->l, m); fe_sub_raw(r->l, t, b->l); }
}
static inline void mont_neg(fe *r, const fe *a, const uint64_t m[4]) {
    if (fe_is_zero(a)) *r = *a; else fe_sub_raw(r->l, m, a->l);
}

#define DEFINE_FIELD(pfx, MOD, INV, ONE, R2)                                                   \
    static inline void pfx##_mul(fe *r, const fe *a, const fe *b) { mont_mul(r, a, b, MOD, INV); } \
    static inline void pfx##_sqr(fe *r, const fe *a) { mont_mul(r, a, a, MOD, INV); }            \
    static inline void pfx##_add(fe *r, const fe *a, const fe *b) { mont_add(r, a, b, MOD); }    \
    static inline void pfx##_sub(fe *r, const fe *a, const fe *b) { mont_sub(r, a, b, MOD); }    \
    static inline void pfx##_neg(fe *r, const fe *a) { mont_neg(r, a, MOD); }                    \
    static inline void pfx##_dbl(fe *r, const fe *a) { mont_add(r, a, a, MOD); }                 \
    static inline void pfx##_one(fe *r) { memcpy(r->l, ONE, 32); }                               \
    static inline void pfx##_zero(fe *r) { memset(r->l, 0, 32); }                                \
    /* canonical little-endian bytes <-> Montgomery form; returns -1 if >= modulus */         \
    static inline int pfx##_from_bytes(fe *r, const uint8_t *b) {                                \
        fe t, r2;                                                                              \
        memcpy(t.l, b, 32);                                                                    \
        if (fe_geq(t.l, MOD)) return -1;                                                       \
        memcpy(r2.l, R2, 32);                                                                  \
        mont_mul(r, &t, &r2, MOD, INV);                                                        \
        return 0;                                                                              \
    }                                                                                          \
    static inline void pfx##_to_bytes(uint8_t *b, const fe *a) {                                 \
        fe one = {{1, 0, 0, 0}}, t;                                                            \
        mont_mul(&t, a, &one, MOD, INV);                                                       \
        memcpy(b, t.l, 32);                                                                    \
    }                                                                                          \
    static inline void pfx##_pow(fe *r, const fe *a, const uint64_t e[4]) {                      \
        fe acc; memcpy(acc.l, ONE, 32);                                                        \
        for (int i = 255; i >= 0; i--) {                                                       \
            mont_mul(&acc, &acc, &acc, MOD, INV);                                              \
            if ((e[i >> 6] >> (i & 63)) & 1) mont_mul(&acc, &acc, a, MOD, INV);                \
        }                                                                                      \
        *r = acc;                                                                              \
    }                                                                                          \
    static inline void pfx##_inv(fe *r, const fe *a) { /* Fermat: a^(m-2) */                     \
        uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};                                  \
        pfx##_pow(r, a, e);                                                                    \
    }

DEFINE_FIELD(fq, FQ_MOD, FQ_INV, FQ_ONE, FQ_R2)
DEFINE_FIELD(fr, FR_MOD, FR_INV, FR_ONE, FR_R2)

/* Fq2 = Fq[i]/(i^2+1) */
typedef struct { fe c0, c1; } fe2;
static inline void fq2_add(fe2 *r, const fe2 *a, const fe2 *b) { fq_add(&r->c0, &a->c0, &b->c0); fq_add(&r->c1, &a->c1, &b->c1); }
static inline void fq2_sub(fe2 *r, const fe2 *a, const fe2 *b) { fq_sub(&r->c0, &a->c0, &b->c0); fq_sub(&r->c1, &a->c1, &b->c1); }
static inline void fq2_dbl(fe2 *r, const fe2 *a) { fq_dbl(&r->c0, &a->c0); fq_dbl(&r->c1, &a->c1); }
static inline void fq2_neg(fe2 *r, const fe2 *a) { fq_neg(&r->c0, &a->c0); fq_neg(&r->c1, &a->c1); }
static inline void fq2_mul(fe2 *r, const fe2 *a, const fe2 *b) {
    fe t0, t1, s0, s1, m;
    fq_mul(&t0, &a->c0, &b->c0);
    fq_mul(&t1, &a->c1, &b->c1);
    fq_add(&s0, &a->c0, &a->c1);
    fq_add(&s1, &b->c0, &b->c1);
    fq_mul(&m, &s0, &s1);
    fq_sub(&m, &m, &t0);
    fq_sub(&r->c1, &m, &t1);
    fq_sub(&r->c0, &t0, &t1);
}
static inline void fq2_sqr(fe2 *r, const fe2 *a) {
    fe s, d, m;
    fq_add(&s, &a->c0, &a->c1);
    fq_sub(&d, &a->c0, &a->c1);
    fq_mul(&m, &a->c0, &a->c1);
    fq_mul(&r->c0, &s, &d);
    fq_dbl(&r->c1, &m);
}
static inline void fq2_one(fe2 *r) { fq_one(&r->c0); fq_zero(&r->c1); }
static inline void fq2_zero(fe2 *r) { fq_zero(&r->c0); fq_zero(&r->c1); }
static inline int fq2_is_zero(const fe2 *a) { return fe_is_zero(&a->c0) && fe_is_zero(&a->c1); }
static inline int fq2_eq(const fe2 *a, const fe2 *b) { return fe_eq(&a->c0, &b->c0) && fe_eq(&a->c1, &b->c1); }
static inline void fq2_inv(fe2 *r, const fe2 *a) {
    fe n, t, ninv;
    fq_sqr(&n, &a->c0);
    fq_sqr(&t, &a->c1);
    fq_add(&n, &n, &t);
    fq_inv(&ninv, &n);
    fq_mul(&r->c0, &a->c0, &ninv);
    fq_mul(&t, &a->c1, &ninv);
    fq_neg(&r->c1, &t);
}
static inline int fq2_from_bytes(fe2 *r, const uint8_t *b) {
    return fq_from_bytes(&r->c0, b) | fq_from_bytes(&r->c1, b + 32);
}
static inline void fq2_to_bytes(uint8_t *b, const fe2 *a) { fq_to_bytes(b, &a->c0); fq_to_bytes(b + 32, &a->c1); }
/* Fq "as a degree-1 extension" helpers so curve_tmpl.h can be instantiated for both */
static inline int fq_is_zero(const fe *a) { return fe_is_zero(a); }
static inline int fq_eq(const fe *a, const fe *b) { return fe_eq(a, b); }
#endif
