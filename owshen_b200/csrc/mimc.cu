// owshen_b200/csrc/mimc.cu -- MiMC7 (circomlib flavour) on sm_100a: 2-to-1 node hash, batched Merkle
// paths (BASELINE config 2), level-by-level tree build, and the witness generator of the withdraw
// statement (every t^2, t^4, t^6, t^7 of every round is a circuit variable).
//
// Not in the reference (its only field "hash" is a placeholder product,
// /root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs:202-204); the algorithm is the
// published circomlib one, see DESIGN.md section 2.  One thread owns one sequential hash chain:
// a path is 32 levels x 2 permutations x 91 rounds x 4 multiplications that each depend on the
// previous one, so the parallelism is across paths, the arithmetic stays in registers and the
// kernel is bound by the integer multiply-add pipe, not by HBM.
#include "common.cuh"
#include "host_math.hpp"
#include "mimc.cuh"
#include "mimc_core.cuh"
#include <stdlib.h>

namespace og {

__constant__ uint32_t c_mimc[MIMC_ROUNDS * 8];   // round constants, Montgomery form

__device__ __forceinline__ Fr mimc_c(int i) {
    Fr r;
#pragma unroll
    for (int j = 0; j < 8; j++) r.l[j] = c_mimc[i * 8 + j];
    return r;
}

// hash(x, k) = perm(x, k) + k.  TRACE: also store t2,t4,t6,t7 of every round to trace[4*i..].
template <bool TRACE>
__device__ __forceinline__ Fr mimc7_hash(const Fr& x, const Fr& k, Fr* trace) {
    Fr r = x;
#pragma unroll 1
    for (int i = 0; i < MIMC_ROUNDS; i++) {
        Fr t = r + k + mimc_c(i);
        Fr t2 = t.sqr();
        if (TRACE) {
            Fr t4 = t2.sqr();
            Fr t6 = t4 * t2;
            r = t6 * t;
            trace[4 * i] = t2; trace[4 * i + 1] = t4; trace[4 * i + 2] = t6; trace[4 * i + 3] = r;
        } else {
            // same value, shallower dependency chain: t3 and t4 are independent
            Fr t3 = t2 * t;
            Fr t4 = t2.sqr();
            r = t3 * t4;
        }
    }
    return r + k;
}

// MultiMiMC7([l, r], key 0): r1 = l + hash(l, 0); out = r1 + r + hash(r, r1).  The TRACE form (witness generation) stores
// every fully reduced intermediate; the plain form is the lazy chain of mimc_core.cuh.
template <bool TRACE>
__device__ __forceinline__ Fr mimc7_hash2(const Fr& l, const Fr& r, Fr* trace1, Fr* trace2) {
    if (!TRACE) return mimc7_hash2_lazy(l, r, [](int i) { return mimc_c(i); });
    Fr r1 = l + mimc7_hash<TRACE>(l, Fr::zero(), trace1);
    return r1 + r + mimc7_hash<TRACE>(r, r1, trace2);
}

__global__ void __launch_bounds__(64) k_hash2(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right,
                                              uint64_t n, uint8_t* __restrict__ out, int* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr l = load_canonical<Fr>(left + 32 * i, flag);
    Fr r = load_canonical<Fr>(right + 32 * i, flag);
    store_canonical(out + 32 * i, mimc7_hash2<false>(l, r, nullptr, nullptr));
}

// one level of a full tree: out[i] = hash2(in[2i], in[2i+1]), Montgomery-form in and out
__global__ void __launch_bounds__(64) k_tree_level(const Fr* __restrict__ in, uint64_t n_out, Fr* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    out[i] = mimc7_hash2<false>(in[2 * i], in[2 * i + 1], nullptr, nullptr);
}

// one level of an APPEND to a fixed-depth sparse tree: parents [p0, p0 + n_out) of the dirty children [c0, c0 + n_in).
// A child left of the dirty range is the stored boundary node of this level (at most one: index c0 - 1), a child right
// of it is the empty-subtree root of the level (append-only: nothing exists to the right of the new leaves).
__global__ void __launch_bounds__(64) k_tree_append_level(const Fr* __restrict__ in, uint64_t c0, uint64_t n_in, uint64_t p0, uint64_t n_out,
                                                          Fr left_boundary, Fr zero, Fr* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    uint64_t l = 2 * (p0 + i), r = l + 1;
    Fr a = l < c0 ? left_boundary : (l < c0 + n_in ? in[l - c0] : zero);
    Fr b = r < c0 ? left_boundary : (r < c0 + n_in ? in[r - c0] : zero);
    out[i] = mimc7_hash2<false>(a, b, nullptr, nullptr);
}

__global__ void __launch_bounds__(128) k_to_mont(const uint8_t* __restrict__ in, uint64_t n, Fr* __restrict__ out, int* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = load_canonical<Fr>(in + 32 * i, flag);
}
__global__ void __launch_bounds__(128) k_from_mont(const Fr* __restrict__ in, uint64_t n, uint8_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_canonical(out + 32 * i, in[i]);
}

// BASELINE config 2.  One thread per path; 32 threads per CTA so that 4096 paths spread over
// 128 SMs instead of piling 4 warps onto 32 of them (the chain is latency-bound per warp).
__global__ void __launch_bounds__(32) k_merkle_paths(const uint8_t* __restrict__ leaves, const uint8_t* __restrict__ siblings,
                                                     const uint32_t* __restrict__ path_bits, uint32_t n_paths, uint32_t depth,
                                                     uint8_t* __restrict__ out_nodes, int* flag) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    Fr cur = load_canonical<Fr>(leaves + 32ull * p, flag);
    uint8_t* o = out_nodes + (uint64_t)p * (depth + 1) * 32;
    store_canonical(o, cur);
    uint32_t bits = path_bits[p];
    const uint8_t* sp = siblings + (uint64_t)p * depth * 32;
#pragma unroll 1
    for (uint32_t l = 0; l < depth; l++) {
        Fr sib = load_canonical<Fr>(sp + 32 * l, flag);
        bool right = (bits >> l) & 1;
        Fr a = right ? sib : cur;
        Fr b = right ? cur : sib;
        cur = mimc7_hash2<false>(a, b, nullptr, nullptr);
        store_canonical(o + 32 * (l + 1), cur);
    }
}

// Witness of the withdraw statement, layout of DESIGN.md section 3 (== oracle/withdraw_circuit.py).
// One thread per proof; W is [batch][n_vars] in Montgomery form.
__global__ void __launch_bounds__(32) k_withdraw_witness(WithdrawLayout L, uint32_t w_stride, const uint8_t* __restrict__ nullifiers,
                                                         const uint8_t* __restrict__ secrets, const uint8_t* __restrict__ recipients,
                                                         const uint8_t* __restrict__ siblings, const uint32_t* __restrict__ path_bits,
                                                         uint32_t batch, Fr* __restrict__ W, int* flag) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    Fr* w = W + (uint64_t)p * w_stride;
    Fr nu = load_canonical<Fr>(nullifiers + 32ull * p, flag);
    Fr se = load_canonical<Fr>(secrets + 32ull * p, flag);
    Fr re = load_canonical<Fr>(recipients + 32ull * p, flag);
    Fr one = Fr::one();
    w[0] = one; w[3] = re; w[4] = nu; w[5] = se;
    w[6] = re.sqr();
    // nullifier_hash = MultiMiMC7([nullifier], key 1) = 1 + nullifier + hash(nullifier, 1)
    w[2] = one + nu + mimc7_hash<true>(nu, one, w + 7);
    Fr cur = mimc7_hash2<true>(nu, se, w + L.cm_base, w + L.cm_base + L.perm);
    w[L.cm_out] = cur;
    uint32_t bits = path_bits[p];
    const uint8_t* sp = siblings + (uint64_t)p * L.depth * 32;
#pragma unroll 1
    for (uint32_t l = 0; l < L.depth; l++) {
        Fr* v = w + L.lvl_base + l * L.lvl_size;
        Fr sib = load_canonical<Fr>(sp + 32 * l, flag);
        bool right = (bits >> l) & 1;
        Fr a = right ? sib : cur;
        Fr b = right ? cur : sib;
        v[0] = sib;
        v[1] = right ? one : Fr::zero();
        v[2] = a;
        cur = mimc7_hash2<true>(a, b, v + 3, v + 3 + L.perm);
        v[3 + 2 * L.perm] = cur;
    }
    w[1] = cur;
}

// ---- host side ------------------------------------------------------------------------------------
void mimc_constants_host(Fr* out91) { mimc7_round_constants(out91); }

int32_t mimc_init(og_ctx* ctx) {
    Fr c[MIMC_ROUNDS];
    mimc7_round_constants(c);
    OG_CUDA(ctx, cudaMemcpyToSymbol(c_mimc, c, sizeof(c)));
    return OG_OK;
}

int32_t mimc_hash2_dev(og_ctx* ctx, const uint8_t* d_l, const uint8_t* d_r, uint64_t n, uint8_t* d_out) {
    if (n == 0) return OG_OK;
    OG_LAUNCH(ctx, k_hash2, (unsigned)((n + 63) / 64), 64, 0, d_l, d_r, n, d_out, ctx->d_flag);
    return OG_OK;
}

int32_t mimc_merkle_paths_dev(og_ctx* ctx, const uint8_t* d_leaves, const uint8_t* d_siblings, const uint32_t* d_bits,
                              uint32_t n_paths, uint32_t depth, uint8_t* d_out) {
    if (n_paths == 0) return OG_OK;
    OG_LAUNCH(ctx, k_merkle_paths, (n_paths + 31) / 32, 32, 0, d_leaves, d_siblings, d_bits, n_paths, depth, d_out, ctx->d_flag);
    return OG_OK;
}

int32_t mimc_to_mont_dev(og_ctx* ctx, const uint8_t* d_in, uint64_t n, Fr* d_out) {
    if (n == 0) return OG_OK;
    OG_LAUNCH(ctx, k_to_mont, (unsigned)((n + 127) / 128), 128, 0, d_in, n, d_out, ctx->d_flag);
    return OG_OK;
}
int32_t mimc_from_mont_dev(og_ctx* ctx, const Fr* d_in, uint64_t n, uint8_t* d_out) {
    if (n == 0) return OG_OK;
    OG_LAUNCH(ctx, k_from_mont, (unsigned)((n + 127) / 128), 128, 0, d_in, n, d_out);
    return OG_OK;
}

// levels: Montgomery-form buffer holding n + n/2 + ... + 1 elements, level 0 already filled
int32_t mimc_tree_build_dev(og_ctx* ctx, Fr* d_levels, uint64_t n_leaves) {
    Fr* in = d_levels;
    for (uint64_t n = n_leaves; n > 1; n >>= 1) {
        Fr* out = in + n;
        OG_LAUNCH(ctx, k_tree_level, (unsigned)((n / 2 + 63) / 64), 64, 0, in, n / 2, out);
        in = out;
    }
    return OG_OK;
}

// d_nodes: Montgomery buffer, level 0 (the n new leaves) already filled; levels 1..depth are appended behind it.
// h_aux (host): 2*depth Montgomery elements = left boundary per level, then empty-subtree root per level (kernel arguments).
int32_t mimc_tree_append_dev(og_ctx* ctx, uint32_t depth, uint64_t start, uint64_t n, const Fr* h_aux, Fr* d_nodes) {
    Fr* in = d_nodes;
    uint64_t c0 = start, n_in = n;
    for (uint32_t l = 0; l < depth; l++) {
        uint64_t p0 = c0 >> 1, p1 = (c0 + n_in - 1) >> 1, n_out = p1 - p0 + 1;
        Fr* out = in + n_in;
        OG_LAUNCH(ctx, k_tree_append_level, (unsigned)((n_out + 63) / 64), 64, 0, in, c0, n_in, p0, n_out, h_aux[l], h_aux[depth + l], out);
        in = out; c0 = p0; n_in = n_out;
    }
    return OG_OK;
}

int32_t withdraw_witness_strided_dev(og_ctx* ctx, const WithdrawLayout& L, uint32_t w_stride, const uint8_t* d_null, const uint8_t* d_sec,
                                     const uint8_t* d_rec, const uint8_t* d_sib, const uint32_t* d_bits, uint32_t batch, Fr* d_W) {
    if (batch == 0) return OG_OK;
    if (w_stride < L.n_vars) return OG_E_INVALID;
    OG_LAUNCH(ctx, k_withdraw_witness, (batch + 31) / 32, 32, 0, L, w_stride, d_null, d_sec, d_rec, d_sib, d_bits, batch, d_W, ctx->d_flag);
    return OG_OK;
}

}  // namespace og

#include "bjj_impl.cuh"
