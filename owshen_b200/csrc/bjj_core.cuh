// owshen_b200/csrc/bjj_core.cuh -- BabyJubJub point arithmetic and the per-signature verification logic,
// host+device so tests can run the exact code without a GPU.  Follows the reference's
// /root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs (line numbers at each function).
#pragma once
#include "fp.cuh"

#if defined(__CUDA_ARCH__)
#define OG_BJJ_FN static __device__ __noinline__
#else
#define OG_BJJ_FN static inline
#endif

namespace og {

struct BjjPoint { Fr x, y, z; };     // z == 0: the reference's "empty accumulator" sentinel

OG_HD Fr bjj_a() { return Fr::from_u32(168700); }
OG_HD Fr bjj_d() { return Fr::from_u32(168696); }

OG_BJJ_FN void bjj_double(BjjPoint* p, const Fr* A) {
    if (p->z.is_zero()) return;
    Fr b = (p->x + p->y).sqr(), c = p->x.sqr(), d = p->y.sqr();
    Fr e = *A * c, f = e + d, h = p->z.sqr();
    Fr j = f - h.dbl();
    p->x = (b - c - d) * j;
    p->y = f * (e - d);
    p->z = f * j;
}

// unified addition (complete on this curve: a is a square, d is not), so the reference's
// "equal points -> double" branch needs no special case
OG_BJJ_FN void bjj_add(BjjPoint* p, const BjjPoint* q, const Fr* A, const Fr* D) {
    if (p->z.is_zero()) { *p = *q; return; }
    if (q->z.is_zero()) return;
    Fr a = p->z * q->z, b = a.sqr(), c = p->x * q->x, d = p->y * q->y;
    Fr e = *D * c * d, f = b - e, g = b + e;
    Fr x3 = a * f * ((p->x + p->y) * (q->x + q->y) - c - d);
    Fr y3 = a * g * (d - *A * c);
    p->x = x3; p->y = y3; p->z = f * g;
}

OG_BJJ_FN void bjj_mul(BjjPoint* out, const BjjPoint* base, const Fr* k, const Fr* A, const Fr* D) {
    uint32_t s[8];
    k->to_canonical(s);
    BjjPoint acc{Fr::zero(), Fr::one(), Fr::zero()};
    for (int i = 255; i >= 0; i--) {
        bjj_double(&acc, A);
        if ((s[i >> 5] >> (i & 31)) & 1) bjj_add(&acc, base, A, D);
    }
    *out = acc;
}

OG_HD bool bjj_on_curve(const Fr& x, const Fr& y, const Fr& A, const Fr& D) {
    Fr xx = x.sqr(), yy = y.sqr();
    return yy + A * xx == Fr::one() + D * xx * yy;
}

// Tonelli-Shanks (r - 1 = 2^28 t, non-residue 7); false if a is a non-residue
OG_BJJ_FN bool fr_sqrt(Fr* out, const Fr* a) {
    if (a->is_zero()) { *out = *a; return true; }
    // t = (r - 1) >> 28
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = FrParams::mod(i);
    e[0] -= 1;
    uint32_t t[8];
    for (int i = 0; i < 8; i++) t[i] = (e[i] >> 28) | (i < 7 ? e[i + 1] << 4 : 0);
    uint32_t half[8];                                  // (t + 1) / 2 ; t is odd
    {
        uint32_t c = 1;
        for (int i = 0; i < 8; i++) { uint32_t v = t[i] + c; c = (v < c) ? 1 : 0; half[i] = v; }
        for (int i = 0; i < 8; i++) half[i] = (half[i] >> 1) | (i < 7 ? half[i + 1] << 31 : 0);
    }
    Fr z = Fr::from_u32(7).pow(t);
    Fr x = a->pow(half), b = a->pow(t);
    uint32_t m = 28;
    while (b != Fr::one()) {
        uint32_t i = 0;
        Fr b2 = b;
        while (b2 != Fr::one()) { b2 = b2.sqr(); i++; if (i == m) return false; }
        Fr w = z;
        for (uint32_t k = 0; k + i + 1 < m; k++) w = w.sqr();
        x = x * w; z = w.sqr(); b = b * z; m = i;
    }
    *out = x;
    return true;
}

OG_HD bool fr_is_odd(const Fr& v) { uint32_t c[8]; v.to_canonical(c); return c[0] & 1; }


// status: 1 verifies, 0 does not, 2 = Err in the reference (public key does not decompress).  `h_mimc` is only
// read when hash_kind == 1 (the caller computes MultiMiMC7([R.x, R.y, pk.x, pk.y, msg]) after decompression via cb).
template <class HashFn>
OG_HD uint8_t bjj_verify_one(const Fr& x, bool pk_odd, const Fr& msg, const Fr& rx, const Fr& ry, const Fr& s, const Fr& base_x,
                             const Fr& base_y, HashFn hash5) {
    const Fr A = bjj_a(), D = bjj_d(), one = Fr::one();
    // decompress (mod.rs:88-98)
    Fr xx = x.sqr();
    Fr den = one - D * xx;
    if (den.is_zero()) return 2;
    Fr y2 = den.inv() * (one - A * xx), y;
    if (!fr_sqrt(&y, &y2)) return 2;
    if (fr_is_odd(y) != pk_odd) y = y.neg();
    // verify (mod.rs:99-115)
    if (!bjj_on_curve(x, y, A, D) || !bjj_on_curve(rx, ry, A, D)) return 0;
    Fr in[5] = {rx, ry, x, y, msg};
    Fr h = hash5(in);
    BjjPoint base{base_x, base_y, one}, pk{x, y, one}, rr{rx, ry, one}, sb, ha;
    bjj_mul(&sb, &base, &s, &A, &D);
    bjj_mul(&ha, &pk, &h, &A, &D);
    bjj_add(&ha, &rr, &A, &D);
    // affine equality by cross-multiplication; an empty accumulator is the affine point (0, 1)
    if (sb.z.is_zero()) sb = BjjPoint{Fr::zero(), one, one};
    if (ha.z.is_zero()) ha = BjjPoint{Fr::zero(), one, one};
    bool eq = (sb.x * ha.z == ha.x * sb.z) && (sb.y * ha.z == ha.y * sb.z);
    return eq ? 1 : 0;
}

}  // namespace og
