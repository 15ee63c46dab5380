// owshen_b200/csrc/mimc_core.cuh -- the MiMC7 permutation as one dependent chain with LAZY reduction (host + device).
//
// A chain of 91 x 4 dependent products runs on one thread, and a lone warp pays for every instruction it issues
// (cycles per round ~ 4 x IMAD.WIDE + 2 x everything else), so the 2 modular additions and 4 final
// subtractions of a fully reduced round (112 of its ~950 instructions) are replaced by two plain additions and ONE
// conditional subtraction of 2p:
//     r in [0, 2p)  ->  t = r + k + c_i < 4p  ->  t in [0, 2p)  ->  t^2, t^3, t^4 < 1.76 p  ->  r' = t^3 t^4 < 1.5 p
// (bounds at fp.cuh: mont_mul_lazy; 4p < 2^256 because p < 0.19 * 2^256).  The result is reduced to [0, p) at the end, so the
// value is the circomlib one bit for bit.  `c(i)` returns the round constant (Montgomery form, < p); KZERO skips the + k.
// Not in the reference; tests: tests/test_host_limbs.py runs this very code on the host.
#pragma once
#include "fp.cuh"

namespace og {

constexpr int MIMC7_ROUNDS = 91;

template <bool KZERO, class CFn>
OG_HD Fr mimc7_perm_lazy(const Fr& x, const Fr& k, CFn c) {      // x, k < p; returns perm(x, k) + k in [0, p)
    Fr r = x;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < MIMC7_ROUNDS; i++) {
        Fr t = Fr::add_raw(r, c(i));
        if (!KZERO) t = Fr::add_raw(t, k);
        t = t.reduce_4p_to_2p();
        Fr t2 = t.sqr_lazy();
        Fr t3 = Fr::mul_lazy(t2, t);          // t3 and t4 are independent: two products deep instead of three
        Fr t4 = t2.sqr_lazy();
        r = Fr::mul_lazy(t3, t4);
    }
    if (!KZERO) r = Fr::add_raw(r, k).reduce_4p_to_2p();
    return r.reduce_2p_to_p();
}

// MultiMiMC7([l, r], key 0): r1 = l + hash(l, 0); out = r1 + r + hash(r, r1)      (l, r < p; result < p)
template <class CFn>
OG_HD Fr mimc7_hash2_lazy(const Fr& l, const Fr& r, CFn c) {
    Fr r1 = l + mimc7_perm_lazy<true>(l, Fr::zero(), c);
    return r1 + r + mimc7_perm_lazy<false>(r, r1, c);
}

}  // namespace og
