/* include/owshen_b200.h -- C ABI of the B200-native Groth16 backend (libowshen_b200.so).
 *
 * WHAT THIS REPLACES IN THE REFERENCE.  OwshenNetwork/owshen @ c7b1f00 has no prover, no FFI and no
 * plugin interface for proving, so there is no reference binding to cite
 * per entry point; this header DEFINES the boundary that BASELINE.json's north_star asks for
 * ("Rust prove()/verify()/MerkleTree ... through a thin C-ABI/FFI layer").  The conventions it
 * inherits from the reference are:
 *   - field elements: BN254 Fr, canonical little-endian 32 bytes
 *     (/root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs:7-11);
 *   - fallible calls return a status instead of panicking, mirroring `anyhow::Result<T>`
 *     (/root/reference/src/blockchain/mod.rs:11); byte blobs are caller-owned buffers
 *     (the reference passes owned `Vec<u8>`, e.g. src/types/tx/custom.rs:258-287);
 *   - the natural call site is a service handler shaped like withdraw_handler
 *     (/root/reference/src/services/api_services/withdraw.rs:27-71); INTEGRATION.md shows the
 *     Rust `extern "C"` stub a maintainer would add there.
 *
 * Formats.  Fr / Fq: 32 B little-endian canonical (values >= modulus are rejected with
 * OG_E_ENCODING).  G1 affine: x || y (64 B).  G2 affine: x.c0 || x.c1 || y.c0 || y.c1 (128 B).
 * The point at infinity is the all-zero encoding.  Proof: A (G1) || B (G2) || C (G1) = 256 B.
 *
 * Threading.  An og_ctx owns one CUDA device and one stream; calls on one ctx must not overlap,
 * different ctxs are independent.  Host-pointer entry points copy H2D/D2H themselves and return
 * after the result is in the caller's buffer.  `_dev` entry points take DEVICE pointers, enqueue
 * on the ctx stream and return without synchronising (use og_sync / og_timer_*).
 *
 * There is NO CPU fallback: without a CUDA device og_init fails with OG_E_NO_DEVICE and nothing
 * else can be called (og_groth16_verify is a host function by design: three pairings).
 */
#ifndef OWSHEN_B200_H
#define OWSHEN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct og_ctx og_ctx;
typedef struct og_pk og_pk;

enum {
    OG_OK = 0,
    OG_E_INVALID = -1,     /* bad argument (null pointer, size out of range) */
    OG_E_ENCODING = -2,    /* non-canonical field element / malformed blob */
    OG_E_NO_DEVICE = -3,   /* no usable CUDA device: the library has no CPU path */
    OG_E_CUDA = -4,        /* a CUDA runtime call failed; see og_last_error */
    OG_E_NOMEM = -5,
    OG_E_VERIFY = -6       /* og_groth16_verify: well-formed proof that does not verify */
};

int32_t og_abi_version(void);
const char* og_strerror(int32_t code);
const char* og_last_error(const og_ctx* ctx);

/* ---- context ------------------------------------------------------------------------------- */
int32_t og_init(int32_t device, og_ctx** out);
void og_free(og_ctx* ctx);
int32_t og_sync(og_ctx* ctx);
/* the CUDA stream (cudaStream_t) every `_dev` entry point of ctx enqueues on: a host that orders other work against
 * the library without a host synchronisation (the NCCL all-gather of the sharded MSM) wraps it, e.g.
 * torch.cuda.ExternalStream(ptr) */
int32_t og_stream(og_ctx* ctx, void** out_cuda_stream);
/* CUDA-event timer on the ctx stream (bench.py times kernels with these, not torch events) */
int32_t og_timer_start(og_ctx* ctx);
int32_t og_timer_stop(og_ctx* ctx, float* ms);
/* number of this library's kernel launches enqueued on ctx since creation */
uint64_t og_launch_count(const og_ctx* ctx);
/* per-kernel timing: CUDA events around every launch while enabled; og_profile_dump synchronises and
 * writes "kernel,launches,total_ms" lines accumulated since the previous dump */
int32_t og_profile(og_ctx* ctx, int32_t enable);
int32_t og_profile_dump(og_ctx* ctx, char* buf, uint64_t cap);
/* integer-pipe micro-benchmark: achieved 32-bit multiply-add lane-ops per second */
int32_t og_imad_peak(og_ctx* ctx, double* mad_per_s, double* wide_mad_per_s);
/* same plus the rate of 32x32->64 multiply-adds issued as mad.lo.cc/madc.hi.cc carry chains (the shape of
 * a Montgomery row, IMAD.WIDE.U32.X): the honest roofline denominator of the field multiplier */
int32_t og_int_pipe_peaks(og_ctx* ctx, double* mad_per_s, double* wide_mad_per_s, double* carry_chain_wide_per_s);

/* SM cycles per dependent Fr multiply(+add) for one warp alone on a scheduler, and per iteration of two
 * independent chains: the latency that bounds the sequential MiMC chains (planning probe) */
int32_t og_mul_latency(og_ctx* ctx, double* cycles_dependent, double* cycles_two_chains);
/* planning probe for a hybrid multiplier: rates4 = {52x52-bit FP64-pipe products/s alone, 32x32-bit carry-chain
 * multiply-adds/s alone, and both rates when the two kinds run interleaved in every warp} */
int32_t og_hybrid_probe(og_ctx* ctx, double* rates4);
/* FP64 fused multiply-adds per second (planning probe: the FP64 pipe is idle in every kernel of this library) */
int32_t og_fp64_peak(og_ctx* ctx, double* dfma_per_s);

/* ---- element-wise field ops (parity probes for the limb arithmetic) ------------------------- */
/* field: 0 = Fq, 1 = Fr; op: 0 = mul, 1 = add, 2 = sub */
int32_t og_field_op(og_ctx* ctx, int32_t field, int32_t op, const uint8_t* a, const uint8_t* b,
                    uint64_t n, uint8_t* out);

/* ---- MiMC7 / Merkle (BASELINE config 2) ------------------------------------------------------ */
/* MiMC7 round constants as derived on the host (keccak chain from "mimc"): 91 * 32 B */
int32_t og_mimc7_constants(uint8_t* out, uint32_t* n_rounds);
/* out[i] = MultiMiMC7([left[i], right[i]], key = 0) */
int32_t og_mimc7_hash2(og_ctx* ctx, const uint8_t* left, const uint8_t* right, uint64_t n, uint8_t* out);
/* out_nodes: n_paths * (depth+1) * 32 B, node 0 = leaf ... node depth = root.
 * path_bits[p] bit l = 1: the running node is the RIGHT child at level l.  depth <= 32. */
int32_t og_mimc7_merkle_paths(og_ctx* ctx, const uint8_t* leaves, const uint8_t* siblings,
                              const uint32_t* path_bits, uint32_t n_paths, uint32_t depth,
                              uint8_t* out_nodes);
int32_t og_mimc7_merkle_paths_dev(og_ctx* ctx, const uint8_t* d_leaves, const uint8_t* d_siblings,
                                  const uint32_t* d_path_bits, uint32_t n_paths, uint32_t depth,
                                  uint8_t* d_out_nodes);
/* full tree build: levels[0] = leaves (n = 2^depth_built padded by the caller), returns all
 * levels concatenated: sum_{l=0..log2 n} (n >> l) * 32 B */
int32_t og_mimc7_merkle_build(og_ctx* ctx, const uint8_t* leaves, uint64_t n_leaves_pow2, uint8_t* out_levels);
/* append to a fixed-depth sparse tree (the incremental builder of SURVEY.md 8f.2): all nodes of levels 1..depth
 * touched by inserting n leaves at index `start`, in ONE call.  left_boundary[l] (32 B per level, l < depth) is the
 * stored node (l, (start >> l) - 1) when (start >> l) is odd (ignored otherwise); zeros[l] is the root of an empty
 * subtree of height l.  out_nodes: level 1 first, level l holds ((start+n-1)>>l) - (start>>l) + 1 nodes. */
int32_t og_mimc7_merkle_append(og_ctx* ctx, uint32_t depth, uint64_t start, const uint8_t* leaves, uint64_t n,
                               const uint8_t* left_boundary, const uint8_t* zeros, uint8_t* out_nodes);

/* ---- BabyJubJub EdDSA-style batch verification -------------------------------------------
 * Replaces a loop over PointCompressed::verify, /root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/
 * mod.rs:99-115 (with decompress :88-98, multiply :68-78, hash :202-204).  Per signature: pk_x 32 B and one
 * is_odd byte (PointCompressed, mod.rs:19), message 32 B, signature = R.x || R.y || s (96 B).
 * hash_kind 0 = the reference's placeholder product hash, 1 = MultiMiMC7 (not reference behaviour).
 * out_status[i]: 1 verifies, 0 does not, 2 = the reference would return Err (pk does not decompress). */
int32_t og_bjj_verify_batch(og_ctx* ctx, const uint8_t* pk_x, const uint8_t* pk_is_odd, const uint8_t* messages,
                            const uint8_t* signatures, uint32_t n, int32_t hash_kind, uint8_t* out_status);
int32_t og_bjj_verify_batch_dev(og_ctx* ctx, const uint8_t* d_pk_x, const uint8_t* d_pk_is_odd, const uint8_t* d_messages,
                                const uint8_t* d_signatures, uint32_t n, int32_t hash_kind, uint8_t* d_out_status);
/* Batch of PrivateKey::to_pub + PrivateKey::sign (mod.rs:206-237): per key 32 B secret scalar, 32 B randomness, 32 B message
 * -> compressed public key (x, is_odd), signature R.x || R.y || s.  out_status[i]: 1 = written; 2 = the reference returns
 * Err("Invalid repr") because s = (r + h a) mod ORDER does not fit the field (ORDER > r, mod.rs:222-233). */
int32_t og_bjj_sign_batch(og_ctx* ctx, const uint8_t* secret_keys, const uint8_t* randomness, const uint8_t* messages,
                          uint32_t n, int32_t hash_kind, uint8_t* out_pk_x, uint8_t* out_pk_is_odd,
                          uint8_t* out_signatures, uint8_t* out_status);
int32_t og_bjj_sign_batch_dev(og_ctx* ctx, const uint8_t* d_secret_keys, const uint8_t* d_randomness, const uint8_t* d_messages,
                              uint32_t n, int32_t hash_kind, uint8_t* d_out_pk_x, uint8_t* d_out_pk_is_odd,
                              uint8_t* d_out_signatures, uint8_t* d_out_status);

/* ---- MSM (BASELINE configs 3 and 5) ----------------------------------------------------------- */
int32_t og_msm_g1(og_ctx* ctx, const uint8_t* points, const uint8_t* scalars, uint64_t n, uint8_t* out64);
int32_t og_msm_g2(og_ctx* ctx, const uint8_t* points, const uint8_t* scalars, uint64_t n, uint8_t* out128);
int32_t og_msm_g1_dev(og_ctx* ctx, const uint8_t* d_points, const uint8_t* d_scalars, uint64_t n, uint8_t* d_out64);
int32_t og_msm_g2_dev(og_ctx* ctx, const uint8_t* d_points, const uint8_t* d_scalars, uint64_t n, uint8_t* d_out128);
/* out_points[i] = scalars[i] * G for the standard generators (fixed-base windows on the GPU) */
int32_t og_g1_generator_mul(og_ctx* ctx, const uint8_t* scalars, uint64_t n, uint8_t* out_points64);
int32_t og_g2_generator_mul(og_ctx* ctx, const uint8_t* scalars, uint64_t n, uint8_t* out_points128);
int32_t og_g1_generator_mul_dev(og_ctx* ctx, const uint8_t* d_scalars, uint64_t n, uint8_t* d_out_points64);
int32_t og_g2_generator_mul_dev(og_ctx* ctx, const uint8_t* d_scalars, uint64_t n, uint8_t* d_out_points128);
/* plain sums of affine points: the local step after the multi-GPU all-gather of partial MSMs */
int32_t og_g1_sum(og_ctx* ctx, const uint8_t* points, uint64_t n, uint8_t* out64);
int32_t og_g2_sum(og_ctx* ctx, const uint8_t* points, uint64_t n, uint8_t* out128);
int32_t og_g1_sum_dev(og_ctx* ctx, const uint8_t* d_points, uint64_t n, uint8_t* d_out64);
int32_t og_g2_sum_dev(og_ctx* ctx, const uint8_t* d_points, uint64_t n, uint8_t* d_out128);

/* ---- NTT over Fr -------------------------------------------------------------------------------- */
/* `batch` independent transforms of 2^log_n elements, contiguous, natural order in and out.
 * omega = 7^((r-1)/2^log_n); coset = 1 evaluates on / interpolates from g*omega^k, g = omega_{2n}. */
int32_t og_ntt(og_ctx* ctx, uint8_t* data, uint32_t log_n, uint32_t batch, int32_t inverse, int32_t coset);
int32_t og_ntt_dev(og_ctx* ctx, uint8_t* d_data, uint32_t log_n, uint32_t batch, int32_t inverse, int32_t coset);

/* ---- the withdraw statement ---------------------------------------------- */
int32_t og_withdraw_r1cs_info(uint32_t depth, uint32_t* n_constraints, uint32_t* n_vars,
                              uint32_t* n_pub, uint32_t* log_m);
/* CSR of matrix `which` (0 = A, 1 = B, 2 = C); pass NULL arrays to query nnz only */
int32_t og_withdraw_r1cs_export(uint32_t depth, int32_t which, uint32_t* row_ptr, uint32_t* col_idx,
                                uint8_t* coeffs, uint64_t* nnz);
/* full assignments (batch * n_vars * 32 B) computed on the GPU */
int32_t og_withdraw_witness(og_ctx* ctx, uint32_t depth, const uint8_t* nullifiers, const uint8_t* secrets,
                            const uint8_t* recipients, const uint8_t* siblings, const uint32_t* path_bits,
                            uint32_t batch, uint8_t* witnesses);

/* ---- Groth16 ------------------------------------------------------------------------------------ */
/* Development ("toxic waste in the clear") setup for the withdraw statement, computed on the GPU.
 * toxic = tau || alpha || beta || gamma || delta (5 * 32 B).  Writes serialized pk / vk blobs;
 * call with pk_out == NULL to get the sizes. */
int32_t og_groth16_setup_withdraw(og_ctx* ctx, uint32_t depth, const uint8_t* toxic160,
                                  uint8_t* pk_out, uint64_t* pk_len, uint8_t* vk_out, uint64_t* vk_len);
/* parse a pk blob, upload it and build the fixed-base window tables in HBM */
int32_t og_load_pk(og_ctx* ctx, const uint8_t* pk_bytes, uint64_t len, og_pk** out);
void og_free_pk(og_pk* pk);
int32_t og_pk_info(const og_pk* pk, uint32_t* n_vars, uint32_t* n_pub, uint32_t* log_m, uint32_t* depth);

/* batch of proofs from full witnesses (batch * n_vars * 32 B); rs = batch * (r || s) */
int32_t og_groth16_prove(og_ctx* ctx, const og_pk* pk, const uint8_t* witnesses, uint32_t batch,
                         const uint8_t* rs, uint8_t* proofs);
/* batch of withdraw proofs from the secret inputs: witness generation (MiMC7 Merkle paths) runs on
 * the GPU too.  public_out (optional): batch * 3 * 32 B = root, nullifier_hash, recipient. */
int32_t og_groth16_prove_withdraw(og_ctx* ctx, const og_pk* pk, const uint8_t* nullifiers,
                                  const uint8_t* secrets, const uint8_t* recipients, const uint8_t* siblings,
                                  const uint32_t* path_bits, uint32_t batch, const uint8_t* rs,
                                  uint8_t* proofs, uint8_t* public_out);
/* same with every buffer already in HBM; no synchronisation */
int32_t og_groth16_prove_withdraw_dev(og_ctx* ctx, const og_pk* pk, const uint8_t* d_nullifiers,
                                      const uint8_t* d_secrets, const uint8_t* d_recipients,
                                      const uint8_t* d_siblings, const uint32_t* d_path_bits, uint32_t batch,
                                      const uint8_t* d_rs, uint8_t* d_proofs, uint8_t* d_public_out);
/* debug/parity probe: the H-query scalars d_j = (a*b - c)(g w^j) for one witness, 2^log_m * 32 B */
int32_t og_groth16_h_evals(og_ctx* ctx, const og_pk* pk, const uint8_t* witness, uint8_t* out);

/* host-side verifier (3 pairings + n_pub scalar multiplications): OG_OK or OG_E_VERIFY */
int32_t og_groth16_verify(const uint8_t* vk, uint64_t vk_len, const uint8_t* public_inputs,
                          uint32_t n_pub, const uint8_t* proof256);

#ifdef __cplusplus
}
#endif
#endif
