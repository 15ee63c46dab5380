// owshen_b200/csrc/fp.cuh -- 256-bit prime-field arithmetic for sm_100a, 8 x 32-bit limbs in registers.
//
// Montgomery form with R = 2^256.  The multiplier is the even/odd split CIOS: the running sum is
// kept in two staggered 8-limb arrays so every 32x32->64 partial product lands in a (lo, hi) pair
// of ONE array and each row is a single carry chain of mad.lo.cc / madc.hi.cc pairs, which ptxas
// fuses into IMAD.WIDE.U32(.X).  No tensor cores: this is modular big-integer work.
//
// Field definition: Fr follows /root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs:7-11
// (modulus, little-endian canonical bytes); Fq is the public alt_bn128 base field.
//
// The same source compiles for the host (carry flag emulated in `CC`) so that tests/ can run the
// exact limb algorithm without a GPU; the product never executes the host path for proving.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OG_HD __host__ __device__ __forceinline__
#else
#define OG_HD inline
#endif

namespace og {

struct CC { uint32_t c = 0; };  // host-side carry/borrow flag; the device uses the PTX CC register

#if defined(__CUDA_ARCH__)
#define OG_ASM asm volatile
OG_HD uint32_t add_cc(uint32_t a, uint32_t b, CC&) { uint32_t r; OG_ASM("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OG_HD uint32_t addc_cc(uint32_t a, uint32_t b, CC&) { uint32_t r; OG_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OG_HD uint32_t addc(uint32_t a, uint32_t b, CC&) { uint32_t r; OG_ASM("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OG_HD uint32_t sub_cc(uint32_t a, uint32_t b, CC&) { uint32_t r; OG_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OG_HD uint32_t subc_cc(uint32_t a, uint32_t b, CC&) { uint32_t r; OG_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OG_HD uint32_t subc(uint32_t a, uint32_t b, CC&) { uint32_t r; OG_ASM("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c, CC&) { uint32_t r; OG_ASM("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c, CC&) { uint32_t r; OG_ASM("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c, CC&) { uint32_t r; OG_ASM("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OG_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c, CC&) { uint32_t r; OG_ASM("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OG_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
OG_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
OG_HD uint32_t add_cc(uint32_t a, uint32_t b, CC& cc) { uint64_t t = (uint64_t)a + b; cc.c = (uint32_t)(t >> 32); return (uint32_t)t; }
OG_HD uint32_t addc_cc(uint32_t a, uint32_t b, CC& cc) { uint64_t t = (uint64_t)a + b + cc.c; cc.c = (uint32_t)(t >> 32); return (uint32_t)t; }
OG_HD uint32_t addc(uint32_t a, uint32_t b, CC& cc) { return a + b + cc.c; }
OG_HD uint32_t sub_cc(uint32_t a, uint32_t b, CC& cc) { uint64_t t = (uint64_t)a - b; cc.c = (uint32_t)(t >> 63); return (uint32_t)t; }
OG_HD uint32_t subc_cc(uint32_t a, uint32_t b, CC& cc) { uint64_t t = (uint64_t)a - b - cc.c; cc.c = (uint32_t)(t >> 63); return (uint32_t)t; }
OG_HD uint32_t subc(uint32_t a, uint32_t b, CC& cc) { return a - b - cc.c; }
OG_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
OG_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
OG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c, CC& cc) { uint64_t t = (uint64_t)mul_lo(a, b) + c; cc.c = (uint32_t)(t >> 32); return (uint32_t)t; }
OG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c, CC& cc) { uint64_t t = (uint64_t)mul_lo(a, b) + c + cc.c; cc.c = (uint32_t)(t >> 32); return (uint32_t)t; }
OG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c, CC& cc) { uint64_t t = (uint64_t)mul_hi(a, b) + c + cc.c; cc.c = (uint32_t)(t >> 32); return (uint32_t)t; }
OG_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c, CC& cc) { return mul_hi(a, b) + c + cc.c; }
#endif

// ---- field parameters ---------------------------------------------------------------------------
// mod(i)/r2(i)/one(i) are constexpr functions so that fully unrolled loops fold them to immediates.
struct FqParams {  // alt_bn128 base field p
    static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
    OG_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    OG_HD static constexpr uint32_t r2(int i) {   // 2^512 mod p
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    OG_HD static constexpr uint32_t one(int i) {  // 2^256 mod p
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
};
struct FrParams {  // BN254 scalar field r  (reference: babyjubjub/mod.rs:8)
    static constexpr uint32_t INV = 0xefffffffu;
    OG_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    OG_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    OG_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
};

// ---- raw limb routines ----------------------------------------------------------------------------
// r = (r >= p) ? r - p : r     (r < 2p)
template <class P>
OG_HD void final_sub(uint32_t* r) {
    uint32_t t[8];
    CC cc;
    t[0] = sub_cc(r[0], P::mod(0), cc);
#pragma unroll
    for (int j = 1; j < 8; j++) t[j] = subc_cc(r[j], P::mod(j), cc);
    uint32_t borrow = subc(0u, 0u, cc);  // 0xffffffff when r < p
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = borrow ? r[j] : t[j];
}

// One interleaved Montgomery row.  E is the array aligned at limb 0, O the one aligned at limb 1;
// on entry (when !first) O is the previous row's E: its limb 0 is dead (zero by construction), O[1] sits
// at limb 0 (the "orphan") and O[2..7] at limbs 1..6.
//  * every a*b chain starts with mad.lo.cc so that ptxas fuses the (lo, hi) pairs into IMAD.WIDE.U32(.X);
//  * the orphan is folded while forming s = E[0] + orphan: its carry enters the q*p chain on O (limb 1),
//    and because s + lo(q*p0) == 0 (mod 2^32) the low product is never computed: its carry is (s != 0),
//    injected with add.cc(s, 0xffffffff) into the q*p chain on E, which starts at limb 1.
// No limb ripples through a whole array, so successive rows overlap and the dependency chain per product
// is short (this kernel family is latency-bound at 4 warps/scheduler; see profiles/).
template <class P>
OG_HD void mont_row(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, bool first) {
    CC cc;
    uint32_t orphan = 0;
    if (first) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            E[j] = mul_lo(a[j], bi);
            E[j + 1] = mul_hi(a[j], bi);
            O[j] = mul_lo(a[j + 1], bi);
            O[j + 1] = mul_hi(a[j + 1], bi);
        }
    } else {
        orphan = O[1];
        // O' = (O >> 2 limbs) + a_odd * bi
        O[0] = mad_lo_cc(a[1], bi, O[2], cc);
        O[1] = madc_hi_cc(a[1], bi, O[3], cc);
#pragma unroll
        for (int j = 2; j < 6; j += 2) {
            O[j] = madc_lo_cc(a[j + 1], bi, O[j + 2], cc);
            O[j + 1] = madc_hi_cc(a[j + 1], bi, O[j + 3], cc);
        }
        O[6] = madc_lo_cc(a[7], bi, 0u, cc);
        O[7] = madc_hi(a[7], bi, 0u, cc);
        // E += a_even * bi ; carry lands on limb 8 = O[7]
        E[0] = mad_lo_cc(a[0], bi, E[0], cc);
        E[1] = madc_hi_cc(a[0], bi, E[1], cc);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            E[j] = madc_lo_cc(a[j], bi, E[j], cc);
            E[j + 1] = madc_hi_cc(a[j], bi, E[j + 1], cc);
        }
        O[7] = addc(O[7], 0u, cc);
    }
    uint32_t s = add_cc(E[0], orphan, cc);            // carry -> limb 1
    uint32_t q = mul_lo(s, P::INV);
    O[0] = madc_lo_cc(P::mod(1), q, O[0], cc);
    O[1] = madc_hi_cc(P::mod(1), q, O[1], cc);
#pragma unroll
    for (int j = 2; j < 6; j += 2) {
        O[j] = madc_lo_cc(P::mod(j + 1), q, O[j], cc);
        O[j + 1] = madc_hi_cc(P::mod(j + 1), q, O[j + 1], cc);
    }
    O[6] = madc_lo_cc(P::mod(7), q, O[6], cc);
    O[7] = madc_hi(P::mod(7), q, O[7], cc);           // T < 2^288: no carry out of limb 8
    (void)add_cc(s, 0xffffffffu, cc);                 // carry = (s != 0) = carry of s + lo(q*p0)
    E[1] = madc_hi_cc(P::mod(0), q, E[1], cc);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        E[j] = madc_lo_cc(P::mod(j), q, E[j], cc);
        E[j + 1] = madc_hi_cc(P::mod(j), q, E[j + 1], cc);
    }
    O[7] = addc(O[7], 0u, cc);
    E[0] = 0;                                         // dead from here on
}

// r = a * b * 2^-256 mod p   (a, b < p; r may alias a or b)
template <class P>
OG_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t E[8], O[8];
    mont_row<P>(E, O, a, b[0], true);
    mont_row<P>(O, E, a, b[1], false);
    mont_row<P>(E, O, a, b[2], false);
    mont_row<P>(O, E, a, b[3], false);
    mont_row<P>(E, O, a, b[4], false);
    mont_row<P>(O, E, a, b[5], false);
    mont_row<P>(E, O, a, b[6], false);
    mont_row<P>(O, E, a, b[7], false);
    CC cc;
    r[0] = add_cc(E[0], O[1], cc);
#pragma unroll
    for (int j = 1; j < 7; j++) r[j] = addc_cc(E[j], O[j + 1], cc);
    r[7] = addc(E[7], 0u, cc);
    final_sub<P>(r);
}

// ---- the field element type -----------------------------------------------------------------------
template <class P>
struct alignas(32) Fp {
    uint32_t l[8];

    OG_HD static Fp zero() { Fp r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
    OG_HD static Fp one() { Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
        return r; }
    OG_HD bool is_zero() const { uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= l[i];
        return o == 0; }
    OG_HD bool operator==(const Fp& b) const { uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= l[i] ^ b.l[i];
        return o == 0; }
    OG_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    OG_HD friend Fp operator*(const Fp& a, const Fp& b) { Fp r; mont_mul<P>(r.l, a.l, b.l); return r; }
    OG_HD Fp sqr() const { Fp r; mont_mul<P>(r.l, l, l); return r; }

    OG_HD friend Fp operator+(const Fp& a, const Fp& b) {
        Fp r; CC cc;
        r.l[0] = add_cc(a.l[0], b.l[0], cc);
#pragma unroll
        for (int j = 1; j < 7; j++) r.l[j] = addc_cc(a.l[j], b.l[j], cc);
        r.l[7] = addc(a.l[7], b.l[7], cc);   // p < 2^254: no carry out
        final_sub<P>(r.l);
        return r;
    }
    OG_HD friend Fp operator-(const Fp& a, const Fp& b) {
        Fp r; CC cc;
        r.l[0] = sub_cc(a.l[0], b.l[0], cc);
#pragma unroll
        for (int j = 1; j < 8; j++) r.l[j] = subc_cc(a.l[j], b.l[j], cc);
        uint32_t borrow = subc(0u, 0u, cc);
        r.l[0] = add_cc(r.l[0], P::mod(0) & borrow, cc);
#pragma unroll
        for (int j = 1; j < 7; j++) r.l[j] = addc_cc(r.l[j], P::mod(j) & borrow, cc);
        r.l[7] = addc(r.l[7], P::mod(7) & borrow, cc);
        return r;
    }
    OG_HD Fp neg() const { return zero() - *this; }
    OG_HD Fp dbl() const { return *this + *this; }

    // canonical integer (little-endian limbs) <-> Montgomery form
    OG_HD static Fp from_canonical(const uint32_t* c) {
        Fp t, r2;
#pragma unroll
        for (int i = 0; i < 8; i++) { t.l[i] = c[i]; r2.l[i] = P::r2(i); }
        return t * r2;
    }
    OG_HD void to_canonical(uint32_t* c) const {
        Fp o = zero(); o.l[0] = 1;
        Fp t = *this * o;
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = t.l[i];
    }
    OG_HD static bool canonical_lt_mod(const uint32_t* c) {
        for (int i = 7; i >= 0; i--) {
            if (c[i] < P::mod(i)) return true;
            if (c[i] > P::mod(i)) return false;
        }
        return false;
    }
    OG_HD static Fp from_u32(uint32_t v) { uint32_t c[8] = {v, 0, 0, 0, 0, 0, 0, 0}; return from_canonical(c); }

    // x^e, e given as canonical 8-limb integer (not constant time; used for inversion / roots only)
    OG_HD Fp pow(const uint32_t* e) const {
        Fp acc = one();
        for (int i = 255; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * *this;
        }
        return acc;
    }
    OG_HD Fp inv() const {  // Fermat: x^(p-2); inv(0) = 0
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = P::mod(i);
        e[0] -= 2;          // p is odd and p mod 2^32 >= 2 for both fields
        return pow(e);
    }
};

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

// ---- Fq2 = Fq[i]/(i^2+1) ----------------------------------------------------------------------------
struct Fq2 {
    Fq c0, c1;
    OG_HD static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    OG_HD static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    OG_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    OG_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    OG_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
    OG_HD friend Fq2 operator+(const Fq2& a, const Fq2& b) { return Fq2{a.c0 + b.c0, a.c1 + b.c1}; }
    OG_HD friend Fq2 operator-(const Fq2& a, const Fq2& b) { return Fq2{a.c0 - b.c0, a.c1 - b.c1}; }
    OG_HD static Fq2 mul_inl(const Fq2& a, const Fq2& b) {     // Karatsuba, 3 Fq muls
        Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
        Fq m = (a.c0 + a.c1) * (b.c0 + b.c1);
        return Fq2{t0 - t1, m - t0 - t1};
    }
    OG_HD static Fq2 sqr_inl(const Fq2& a) {                    // 2 Fq muls
        Fq m = a.c0 * a.c1;
        return Fq2{(a.c0 + a.c1) * (a.c0 - a.c1), m + m};
    }
    OG_HD friend Fq2 operator*(const Fq2& a, const Fq2& b);
    OG_HD Fq2 sqr() const;
    OG_HD Fq2 neg() const { return Fq2{c0.neg(), c1.neg()}; }
    OG_HD Fq2 dbl() const { return Fq2{c0.dbl(), c1.dbl()}; }
    OG_HD Fq2 inv() const {
        Fq n = (c0.sqr() + c1.sqr()).inv();
        return Fq2{c0 * n, (c1 * n).neg()};
    }
    OG_HD Fq2 conj() const { return Fq2{c0, c1.neg()}; }
    OG_HD Fq2 mul_fq(const Fq& s) const { return Fq2{c0 * s, c1 * s}; }
};

// Fq2 products are inlined.  Kernels keep ptxas time sane by calling the out-of-line group operations of
// ec.cuh (xyzz_*_ni) everywhere except in the bucket-accumulation inner loop (a G2 group operation is
// ~10k instructions; inlining several of them into one kernel once cost 30 minutes of ptxas).
#if defined(__CUDA_ARCH__) && defined(OG_FP_MUL_CALL)
// Translation units whose kernels would inline dozens of Fq2 products per group operation keep ONE copy of
// the Fq2 multiplier / squarer (arguments and result travel in registers): the fully inlined G2 bucket kernel
// overflowed the instruction cache (ncu: 30 % of warp samples in "no_instructions", profiles/).
static __device__ __noinline__ Fq2 fq2_mul_call(Fq2 a, Fq2 b) { return Fq2::mul_inl(a, b); }
static __device__ __noinline__ Fq2 fq2_sqr_call(Fq2 a) { return Fq2::sqr_inl(a); }
OG_HD Fq2 operator*(const Fq2& a, const Fq2& b) { return fq2_mul_call(a, b); }
OG_HD Fq2 Fq2::sqr() const { return fq2_sqr_call(*this); }
#else
OG_HD Fq2 operator*(const Fq2& a, const Fq2& b) { return Fq2::mul_inl(a, b); }
OG_HD Fq2 Fq2::sqr() const { return Fq2::sqr_inl(*this); }
#endif

}  // namespace og
