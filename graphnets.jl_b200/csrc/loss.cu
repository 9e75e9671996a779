// Loss over the compact views: Flux.logitcrossentropy(flatunpaddednf(y), flatunpaddednf(t)) as the reference's training example
// uses it (examples/sort/sort.jl:76-78) = mean over the rows (real nodes / active edges) of  -sum_d t[r][d] * logsoftmax(y[r])[d].
// flatunpaddednf / flatunpaddedef (src/views.jl:80-98) are free in the compact layout, so the loss reads the model output
// exactly once, straight from the (D, R) matrices the forward wrote.  HBM-bound; deterministic (fixed-order two-stage sum).
#include "common.cuh"

namespace {

constexpr int LOSS_THREADS = 256;

// one warp per row; lane-strided columns; per-block partial = ordered sum of its warps' rows
__global__ void __launch_bounds__(LOSS_THREADS) k_xent_rows(const float* __restrict__ x, const float* __restrict__ t, int D, int64_t R,
                                                          float* __restrict__ per_row, double* __restrict__ block_part) {
  __shared__ double wsum[LOSS_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps = (int64_t)gridDim.x * (LOSS_THREADS / 32);
  double acc = 0.0;      // this warp's rows, ascending
  for (int64_t r = (int64_t)blockIdx.x * (LOSS_THREADS / 32) + warp; r < R; r += warps) {
    const float* xr = x + (size_t)r * D;
    const float* tr = t + (size_t)r * D;
    float mx = -INFINITY;
    for (int d = lane; d < D; d += 32) mx = fmaxf(mx, xr[d]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f, tx = 0.f, ts = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float xv = xr[d], tv = tr[d];
      se += expf(xv - mx);
      tx = fmaf(tv, xv, tx);
      ts += tv;
    }
    for (int o = 16; o; o >>= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, o);
      tx += __shfl_xor_sync(0xffffffffu, tx, o);
      ts += __shfl_xor_sync(0xffffffffu, ts, o);
    }
    const float lse = mx + logf(se);
    const float l = ts * lse - tx;      // -sum_d t_d (x_d - lse)
    if (lane == 0 && per_row) per_row[r] = l;
    acc += (double)l;
  }
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; w++) s += wsum[w];
    block_part[blockIdx.x] = s;
  }
}
__global__ void k_xent_finish(const double* __restrict__ block_part, int nb, int64_t R, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nb; i++) s += block_part[i];
    out[0] = (float)(s / (double)(R > 0 ? R : 1));      // agg = mean (Flux default)
  }
}

// d loss / d logits of the mean cross-entropy:  (sum_d t[r][d]) softmax(x[r]) - t[r], scaled by scale / R; one warp per row
__global__ void __launch_bounds__(LOSS_THREADS) k_xent_bwd(const float* __restrict__ x, const float* __restrict__ t, int D, int64_t R, float scale,
                                                         float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x) >> 5;
  if (r >= R) return;
  const float* xr = x + (size_t)r * D;
  const float* tr = t + (size_t)r * D;
  float mx = -INFINITY;
  for (int d = lane; d < D; d += 32) mx = fmaxf(mx, xr[d]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f, ts = 0.f;
  for (int d = lane; d < D; d += 32) { se += expf(xr[d] - mx); ts += tr[d]; }
  for (int o = 16; o; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); ts += __shfl_xor_sync(0xffffffffu, ts, o); }
  const float c = scale / (float)R;
  for (int d = lane; d < D; d += 32) dx[(size_t)r * D + d] = c * (ts * expf(xr[d] - mx) / se - tr[d]);
}

}  // namespace

extern "C" int gnb_logit_cross_entropy_bwd(gnb_ctx* ctx, const float* logits, const float* targets, int D, int64_t R, float scale,
                                           float* dlogits) {
  GNB_CHECK(ctx && D > 0 && R >= 0, "gnb_logit_cross_entropy_bwd: bad arguments");
  if (R == 0) return GNB_OK;
  GNB_CHECK(logits && targets && dlogits, "gnb_logit_cross_entropy_bwd: null matrix");
  GNB_CUDA(cudaSetDevice(ctx->device));
  Launch L(ctx, "xent_bwd", 12.0 * R * D, 0);
  k_xent_bwd<<<ceil_div(R * 32, LOSS_THREADS), LOSS_THREADS, 0, ctx->stream>>>(logits, targets, D, R, scale, dlogits);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

extern "C" int gnb_logit_cross_entropy(gnb_ctx* ctx, const float* logits, const float* targets, int D, int64_t R, float* loss,
                                       float* per_row) {
  GNB_CHECK(ctx && loss, "gnb_logit_cross_entropy: null argument");
  GNB_CHECK(D > 0 && R >= 0, "gnb_logit_cross_entropy: need D > 0 and R >= 0");
  GNB_CHECK(R == 0 || (logits && targets), "gnb_logit_cross_entropy: null feature matrix");
  GNB_CUDA(cudaSetDevice(ctx->device));
  int64_t nb64 = (R + LOSS_THREADS / 32 - 1) / (LOSS_THREADS / 32);
  const int cap = ctx->sm_count * 8;
  const int nb = (int)(nb64 < 1 ? 1 : (nb64 > cap ? cap : nb64));
  double* part = nullptr;
  // the partial sums live in the staging arena of the context (not the forward's scratch arena: the logits may live there)
  int rc = GNB_OK;
  ctx->staging.reset();
  part = arena_ptr<double>(ctx->staging, nb, &rc);
  if (rc != GNB_OK) return rc;
  {
    Launch L(ctx, "xent_rows", 8.0 * R * D, 0);
    k_xent_rows<<<nb, LOSS_THREADS, 0, ctx->stream>>>(logits, targets, D, R, per_row, part);
  }
  k_xent_finish<<<1, 32, 0, ctx->stream>>>(part, nb, R, loss);
  ctx->launches++;
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
