// Launch interfaces of the compute kernels (internal).
#pragma once
#include "common.cuh"

// One direct (un-gathered) row source of a fused linear layer:
//   contribution[r][n] = sum_k  LN?(x[r][:])[k] * W[k][n]
// W is a k-major sub-block of a Flux Dense.weight ((out,in) column-major == [in][out]),
// i.e. rows [koff, koff+d) of the full matrix: W = Wfull + koff*ldw.
struct LinSrc {
  const float* x;      // [R][ldx]
  int d;               // width of this source (number of k rows)
  int ldx;
  const float* W;      // [d][ldw]
  const float* gamma;  // LayerNorm affine, nullptr = no LayerNorm on this source
  const float* beta;
  float eps;
  int eps_mode;
  int x_bf16;          // tensor-core path only: x is a bf16 [R][ldx] matrix (the hidden activation of an FFN), no LayerNorm
};

// Row-gathered addend of the epilogue: out[r][:] += a[idx ? idx[r] : r][:]
struct LinAdd {
  const float* a;
  const int32_t* idx;
  int lda;
};

struct LinArgs {
  int64_t R;      // rows
  int Nout;       // output width
  int ldw;        // leading dim of every W block (= out dim of the Dense layer)
  int nsrc;
  LinSrc src[3];
  const float* bias;  // [Nout] or nullptr
  int nadd;
  LinAdd add[4];
  int relu;
  float* out;     // [R][ldo]
  int ldo;
  int out_bf16;   // tensor-core path only: out is a bf16 [R][ldo] matrix
};

// out = act( sum_s LN_s(x_s) W_s + bias + sum_j add_j[idx_j] )     fp32 CUDA cores
int launch_linear_fp32(gnb_ctx* ctx, const LinArgs& a);

// dispatch: bf16 tcgen05 kernel (tc_gemm.cu) when the forward runs in a tensor-core precision mode and the layer is wide
// enough, else the fp32 kernel
int launch_linear(gnb_ctx* ctx, const LinArgs& a);

// out[s][:] = sum_{r in [ptr[s], ptr[s+1])} x[r][:]  in ascending r (deterministic)
int launch_segsum(gnb_ctx* ctx, const float* x, int D, const int32_t* ptr, int64_t S, float* out);




// shared between model.cu (fp32 orchestration) and tc.cu (graph-level pieces of the tensor path)
LinSrc mk_src(const float* x, int d, const float* W, const gnb_ln_params* ln);
// y = (x + h) + W2 relu(W1 LN2(x) + b1) + b2   on the fp32 CUDA-core path
int run_ffn_residual_fp32(gnb_ctx* ctx, int64_t R, int d, const gnb_ffn_params& f, const gnb_ln_params& ln2,
                          const float* x, const float* h, float* y);
