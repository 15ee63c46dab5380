// owshen_b200/csrc/mimc.cuh -- declarations of the MiMC7 module and the variable layout of the
// withdraw statement (must equal oracle/withdraw_circuit.py: Layout).
#pragma once
#include "common.cuh"

namespace og {

constexpr uint32_t WITHDRAW_N_PUB = 3;

struct WithdrawLayout {
    uint32_t depth, perm, cm_base, cm_out, lvl_base, lvl_size, n_vars, n_constraints;
    static WithdrawLayout make(uint32_t depth, uint32_t n_rounds = 91) {
        WithdrawLayout L;
        L.depth = depth;
        L.perm = 4 * n_rounds;
        L.cm_base = 7 + L.perm;
        L.cm_out = L.cm_base + 2 * L.perm;
        L.lvl_base = L.cm_out + 1;
        L.lvl_size = 3 + 2 * L.perm + 1;
        L.n_vars = L.lvl_base + depth * L.lvl_size;
        L.n_constraints = 1 + (L.perm + 1) + (2 * L.perm + 1) + depth * (2 + 2 * L.perm + 1) + 1;
        return L;
    }
};

int32_t mimc_hash2_dev(og_ctx* ctx, const uint8_t* d_l, const uint8_t* d_r, uint64_t n, uint8_t* d_out);
int32_t mimc_merkle_paths_dev(og_ctx* ctx, const uint8_t* d_leaves, const uint8_t* d_siblings, const uint32_t* d_bits,
                              uint32_t n_paths, uint32_t depth, uint8_t* d_out);
int32_t mimc_to_mont_dev(og_ctx* ctx, const uint8_t* d_in, uint64_t n, Fr* d_out);
int32_t mimc_from_mont_dev(og_ctx* ctx, const Fr* d_in, uint64_t n, uint8_t* d_out);
int32_t mimc_tree_build_dev(og_ctx* ctx, Fr* d_levels, uint64_t n_leaves);
int32_t mimc_tree_append_dev(og_ctx* ctx, uint32_t depth, uint64_t start, uint64_t n, const Fr* h_aux, Fr* d_nodes);
// W rows are w_stride elements apart (the prover keeps two extra scalars after every witness)
int32_t withdraw_witness_strided_dev(og_ctx* ctx, const WithdrawLayout& L, uint32_t w_stride, const uint8_t* d_null, const uint8_t* d_sec,
                                     const uint8_t* d_rec, const uint8_t* d_sib, const uint32_t* d_bits, uint32_t batch, Fr* d_W);

// BabyJubJub batch verification (bjj_impl.cuh); out[i] in {0, 1, 2 = public key does not decompress}
int32_t bjj_verify_dev(og_ctx* ctx, const uint8_t* d_pk_x, const uint8_t* d_pk_odd, const uint8_t* d_msgs, const uint8_t* d_sigs,
                       uint32_t n, int hash_kind, uint8_t* d_out);

// batch of PrivateKey::to_pub + sign; status[i] in {1, 2 = "Invalid repr" in the reference}
int32_t bjj_sign_dev(og_ctx* ctx, const uint8_t* d_sk, const uint8_t* d_rnd, const uint8_t* d_msgs, uint32_t n, int hash_kind,
                     uint8_t* d_pk_x, uint8_t* d_pk_odd, uint8_t* d_sigs, uint8_t* d_status);

}  // namespace og
