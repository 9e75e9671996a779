// Generic bf16 tcgen05 linear layer for wide GNCore layers (internal interface, see tc_gemm.cu).
#pragma once
#include "kernels.cuh"

// true when launch_linear_tc can run this layer (row sources multiples of 64 wide, output a multiple of 128, enough rows)
bool tc_lin_supported(const LinArgs& a);
// same contract as launch_linear_fp32, bf16 operands / fp32 accumulation on the tensor cores
int launch_linear_tc(gnb_ctx* ctx, const LinArgs& a);
void tc_lin_cache_free(void* cache);
// drops (and frees) the packed weights of one model from a context's cache (gnb_model_destroy)
void tc_lin_cache_evict(void* cache, uint64_t model_id);

// fused GNFeedForward + residuals, y = x + h + W2 relu(W1 LN2(x) + b1) + b2, for hidden width 256 or 384 (hidden activation on chip)
bool tc_ffn256_supported(int64_t R, int d);
int launch_ffn256_tc(gnb_ctx* ctx, int64_t R, const gnb_ffn_params& f, const gnb_ln_params& ln2, const float* x, const float* h, float* y, int d);
