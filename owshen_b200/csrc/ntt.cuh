// owshen_b200/csrc/ntt.cuh -- interface of the batched NTT module.
#pragma once
#include "common.cuh"

namespace og {
// `batch` contiguous transforms of 2^log_n Montgomery-form elements, in place; tmp: batch*n scratch
// (only touched when log_n > 10).
int32_t ntt_mont_dev(og_ctx* ctx, Fr* data, Fr* tmp, uint32_t log_n, uint32_t batch, int inverse, int coset);
// builds the twiddle table of size 2^log_n on the current stream if it does not exist yet (call before forking lanes)
int32_t ntt_prepare(og_ctx* ctx, uint32_t log_n);
void ntt_free_tables(og_ctx* ctx);
}  // namespace og
