// owshen_b200/csrc/ec.cuh -- short-Weierstrass (a = 0) group arithmetic in extended Jacobian
// ("XYZZ") coordinates, templated on the coordinate field (Fq -> G1, Fq2 -> G2).
//
// x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2; infinity is ZZ == 0.  Affine infinity is (0, 0), which
// is off-curve because b != 0 and matches the all-zero boundary encoding (include/owshen_b200.h).
// Formulas are the public EFD ones (madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1) with
// every exceptional case handled, because Pippenger buckets do meet P+P, P+(-P) and infinity.
// No counterpart in the reference: its only curve is BabyJubJub.
#pragma once
#include "fp.cuh"

namespace og {

template <class F>
struct Affine {
    F x, y;
    OG_HD static Affine inf() { return Affine{F::zero(), F::zero()}; }
    OG_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    OG_HD Affine neg() const { return Affine{x, y.neg()}; }
    OG_HD bool operator==(const Affine& o) const { return x == o.x && y == o.y; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;

    OG_HD static XYZZ inf() { return XYZZ{F::one(), F::one(), F::zero(), F::zero()}; }
    OG_HD bool is_inf() const { return zz.is_zero(); }
    OG_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
    OG_HD XYZZ neg() const { return XYZZ{x, y.neg(), zz, zzz}; }

    // 2 * (affine p), p finite
    OG_HD static XYZZ dbl_affine(const Affine<F>& p) {
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        F y3 = m * (s - x3) - w * p.y;
        return XYZZ{x3, y3, v, w};
    }

    OG_HD XYZZ dbl() const {
        if (is_inf()) return *this;
        F u = y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        F y3 = m * (s - x3) - w * y;
        return XYZZ{x3, y3, v * zz, w * zzz};
    }

    // this += affine q   (8M + 2S in the generic case)
    OG_HD void madd(const Affine<F>& q) {
        if (q.is_inf()) return;
        if (is_inf()) { x = q.x; y = q.y; zz = F::one(); zzz = F::one(); return; }
        F u2 = q.x * zz;
        F s2 = q.y * zzz;
        F p = u2 - x;
        F r = s2 - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(q);
            else *this = inf();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q1 = x * pp;
        F x3 = r.sqr() - ppp - q1.dbl();
        y = r * (q1 - x3) - y * ppp;
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += o   (12M + 2S in the generic case)
    OG_HD void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        F u1 = x * o.zz;
        F u2 = o.x * zz;
        F s1 = y * o.zzz;
        F s2 = o.y * zzz;
        F p = u2 - u1;
        F r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = inf();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q1 = u1 * pp;
        F x3 = r.sqr() - ppp - q1.dbl();
        y = r * (q1 - x3) - s1 * ppp;
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }

    OG_HD Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        F t = (zz * zzz).inv();
        return Affine<F>{x * (t * zzz), y * (t * zz)};
    }

    // k * p, k a canonical 8-limb integer (plain MSB-first double-and-add)
    OG_HD static XYZZ mul(const Affine<F>& p, const uint32_t* k) {
        XYZZ acc = inf();
        for (int i = 255; i >= 0; i--) {
            acc = acc.dbl();
            if ((k[i >> 5] >> (i & 31)) & 1) acc.madd(p);
        }
        return acc;
    }
};

// Out-of-line copies for kernels where a group operation is not the inner loop (reductions, table
// builds, finalisation): one body per field instead of one per call site keeps ptxas time sane.
#if defined(__CUDACC__)
template <class F> __device__ __noinline__ void xyzz_add_ni(XYZZ<F>* a, const XYZZ<F>* b) { a->add(*b); }
template <class F> __device__ __noinline__ void xyzz_madd_ni(XYZZ<F>* a, const Affine<F>* b) { a->madd(*b); }
template <class F> __device__ __noinline__ void xyzz_dbl_ni(XYZZ<F>* a) { *a = a->dbl(); }
template <class F> __device__ __noinline__ void xyzz_to_affine_ni(Affine<F>* r, const XYZZ<F>* a) { *r = a->to_affine(); }
#endif

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

}  // namespace og
