// Primitive operators of the TRAINING step (backward pass of the GNBlock / GNCore forward, fp32 CUDA cores).
//
// The reference differentiates (m::GNBlock)(x) (src/gnblock.jl:63-69) with Zygote inside Flux.withgradient
// (examples/sort/sort.jl:122-132).  Every piece of that forward is linear except LayerNorm and relu, so its adjoint is built from
//   * the forward's own fused linear kernel (dX = dY W^T, with the same "project per node / per graph, gather-add in the
//     epilogue" trick: the adjoint of a gather is a segmented sum, taken BEFORE the GEMM by linearity)         gnb_op_linear
//   * deterministic segmented sums over the receiver-sorted index, and over a sender-sorted permutation       gnb_op_segsum
//   * weight gradients dW = X^T dY as a two-stage, atomic-free reduction over row chunks                      gnb_op_wgrad / gnb_op_colsum
//   * LayerNorm forward / backward for the three eps conventions of the forward                               gnb_op_layernorm(_bwd)
//   * relu mask, gather-add, transpose, AdamW                                                                  small element-wise kernels
// host-side orchestration: graphnets.jl_b200/train.py; oracle: torch float64 autograd of oracle/gn_oracle.py's formulation.
#include "kernels.cuh"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ float ln_rstd_t(float var, float eps, int mode) {
  if (mode == GNB_EPS_SQRT_VAR_EPS2) return 1.0f / sqrtf(var + eps * eps);
  if (mode == GNB_EPS_STD_PLUS_EPS) return 1.0f / (sqrtf(var) + eps);
  return 1.0f / sqrtf(var + eps);
}
// d rstd / d var
__device__ __forceinline__ float ln_drstd_t(float var, float eps, int mode) {
  if (mode == GNB_EPS_SQRT_VAR_EPS2) { const float t = var + eps * eps; return -0.5f / (t * sqrtf(t)); }
  if (mode == GNB_EPS_STD_PLUS_EPS) { const float s = sqrtf(fmaxf(var, 1e-30f)); return -0.5f / (s * (s + eps) * (s + eps)); }
  const float t = var + eps;
  return -0.5f / (t * sqrtf(t));
}

// one warp per row
__global__ void k_ln_fwd(const float* __restrict__ x, int64_t R, int D, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float eps, int mode, float* __restrict__ y) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* xr = x + (size_t)r * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s += xr[k];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)D;
  float q = 0.f;
  for (int k = lane; k < D; k += 32) { const float t = xr[k] - mu; q += t * t; }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rs = ln_rstd_t(q / (float)D, eps, mode);
  for (int k = lane; k < D; k += 32) y[(size_t)r * D + k] = (xr[k] - mu) * rs * gamma[k] + beta[k];
}

// y = gamma (x - mu) s(var) + beta;  g = dL/dy.   dx += s (g gamma - mean(g gamma)) + (2 s'/D) (x - mu) sum(g gamma (x - mu));
// gxhat = g (x - mu) s   (its column sum is d gamma, the column sum of g is d beta)
__global__ void k_ln_bwd(const float* __restrict__ x, const float* __restrict__ g, int64_t R, int D, const float* __restrict__ gamma,
                         float eps, int mode, float* __restrict__ dx, float* __restrict__ gxhat) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* xr = x + (size_t)r * D;
  const float* gr = g + (size_t)r * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s += xr[k];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)D;
  float q = 0.f, a = 0.f, b = 0.f;      // sum (x-mu)^2, sum g gamma, sum g gamma (x-mu)
  for (int k = lane; k < D; k += 32) {
    const float t = xr[k] - mu, gg = gr[k] * gamma[k];
    q += t * t; a += gg; b += gg * t;
  }
  for (int o = 16; o; o >>= 1) {
    q += __shfl_xor_sync(0xffffffffu, q, o);
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const float var = q / (float)D;
  const float rs = ln_rstd_t(var, eps, mode), drs = ln_drstd_t(var, eps, mode);
  const float ma = a / (float)D, c2 = 2.f * drs * b / (float)D;
  for (int k = lane; k < D; k += 32) {
    const float t = xr[k] - mu;
    dx[(size_t)r * D + k] += rs * (gr[k] * gamma[k] - ma) + c2 * t;
    gxhat[(size_t)r * D + k] = gr[k] * t * rs;
  }
}

// ---- dW[k][n] (+)= sum_r X[idx ? idx[r] : r][k] dY[r][n]: stage 1 = partial sums per row chunk (grid.z), stage 2 = ordered reduction
// 64 x 64 output tile per CTA, 4 x 4 per thread, 16 rows per step (LDS.128 : FFMA = 2 : 16)
constexpr int WG_T = 64, WG_R = 16;
__global__ void __launch_bounds__(256) k_wgrad_part(const float* __restrict__ X, int ldx, int K, const int32_t* __restrict__ idx,
                                                    const float* __restrict__ dY, int ldy, int N, int64_t R, int64_t rows_per_chunk,
                                                    float* __restrict__ part /*[chunks][K][N]*/) {
  __shared__ __align__(16) float Xs[2][WG_R][WG_T];
  __shared__ __align__(16) float Ys[2][WG_R][WG_T];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16: k = k0 + 4 ty + i, n = n0 + 4 tx + j
  const int k0 = blockIdx.x * WG_T, n0 = blockIdx.y * WG_T;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_chunk;
  const int64_t r_end = r_begin + rows_per_chunk < R ? r_begin + rows_per_chunk : R;
  const int lc = threadIdx.x & 63, lr = threadIdx.x >> 6;      // loader: column lc, rows lr, lr + 4, ...
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  // software pipelined: the rows of step s+1 are fetched into registers while step s is multiplied out of shared memory
  float px[WG_R / 4], py[WG_R / 4];
  auto gload = [&](int64_t rb) {
#pragma unroll
    for (int q = 0; q < WG_R / 4; q++) {
      const int64_t r = rb + lr + 4 * q;
      float xv = 0.f, yv = 0.f;
      if (r < r_end) {
        const int64_t xr = idx ? (int64_t)idx[r] : r;
        if (k0 + lc < K) xv = X[(size_t)xr * ldx + k0 + lc];
        if (n0 + lc < N) yv = dY[(size_t)r * ldy + n0 + lc];
      }
      px[q] = xv; py[q] = yv;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int q = 0; q < WG_R / 4; q++) { Xs[buf][lr + 4 * q][lc] = px[q]; Ys[buf][lr + 4 * q][lc] = py[q]; }
  };
  gload(r_begin);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int64_t rb = r_begin; rb < r_end; rb += WG_R, buf ^= 1) {
    const bool more = rb + WG_R < r_end;
    if (more) gload(rb + WG_R);
#pragma unroll
    for (int rr = 0; rr < WG_R; rr++) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Xs[buf][rr][4 * ty]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Ys[buf][rr][4 * tx]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int k = k0 + 4 * ty + i, n = n0 + 4 * tx + j;
      if (k < K && n < N) part[((size_t)blockIdx.z * K + k) * N + n] = acc[i][j];
    }
}
__global__ void k_reduce_parts(const float* __restrict__ part, int chunks, int K, int N, float* __restrict__ out, int ldo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i % N;
  float s = 0.f;
  for (int c = 0; c < chunks; c++) s += part[(size_t)c * K * N + i];      // fixed order: deterministic
  out[(size_t)k * ldo + n] += s;
}
// column sums: part[chunk][d]
__global__ void __launch_bounds__(256) k_colsum_part(const float* __restrict__ X, int ldx, int D, int64_t R, int64_t rows_per_chunk,
                                                     float* __restrict__ part) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + tx;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r_end = r_begin + rows_per_chunk < R ? r_begin + rows_per_chunk : R;
  float s = 0.f;
  if (d < D)
    for (int64_t r = r_begin + ty; r < r_end; r += 8) s += X[(size_t)r * ldx + d];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && d < D) {
    float t = 0.f;
    for (int j = 0; j < 8; j++) t += red[j][tx];
    part[(size_t)blockIdx.y * D + d] = t;
  }
}

__global__ void k_relu_mask(float* __restrict__ t, const float* __restrict__ h, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(h[i] > 0.f)) t[i] = 0.f;
}
// out[r][:] = (a ? a[r][:] : 0) + (b1 ? b1[i1 ? i1[r] : r][:] : 0) + (b2 ? b2[i2 ? i2[r] : r][:] : 0)
__global__ void k_gather_add(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b1, const int32_t* __restrict__ i1,
                             const float* __restrict__ b2, const int32_t* __restrict__ i2, int64_t R, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * D) return;
  const int64_t r = i / D;
  const int c = (int)(i % D);
  float v = a ? a[i] : 0.f;
  if (b1) v += b1[(size_t)(i1 ? (int64_t)i1[r] : r) * D + c];
  if (b2) v += b2[(size_t)(i2 ? (int64_t)i2[r] : r) * D + c];
  out[i] = v;
}
// out[s][:] = sum over p in [ptr[s], ptr[s+1]) of x[perm ? perm[p] : p][:], ascending p (deterministic); one warp per segment
__global__ void k_segsum_perm(const float* __restrict__ x, int D, const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm, int64_t S,
                              float* __restrict__ out) {
  const int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (seg >= S) return;
  const int64_t p0 = ptr[seg], p1 = ptr[seg + 1];
  for (int c = lane; c < D; c += 32) {
    float s = 0.f;
    for (int64_t p = p0; p < p1; p++) s += x[(size_t)(perm ? (int64_t)perm[p] : p) * D + c];
    out[(size_t)seg * D + c] = s;
  }
}
__global__ void k_transpose(const float* __restrict__ in, int rows, int cols, int ld_in, float* __restrict__ out) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    t[j][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * rows + r] = t[threadIdx.x][j];
  }
}
// decoupled weight decay (AdamW, examples/sort/sort.jl:118 uses Flux.Optimise.AdamW)
__global__ void k_adamw(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                        float b1, float b2, float eps, float wd, float c1, float c2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] = p[i] - lr * ((mi / c1) / (sqrtf(vi / c2) + eps) + wd * p[i]);
}

// ---- dW = X^T dY on the tensor cores (bf16 operands, fp32 accumulation in TMEM): one 128 x 128 output tile per CTA, the
// reduction (rows) split over grid.z chunks; per 64-row slab eight producer warps transpose X[rows][k0 .. k0+128) and
// dY[rows][n0 .. n0+128) into K-major 128B-swizzled bf16 operand slabs (lane = operand row m, eight consecutive reduction rows
// -> one 16 B chunk: the global loads of a warp are 128 B coalesced, the shared stores conflict-free), warp 8 issues 4 UMMAs per
// slab through a 3-stage ring, warps 0-3 drain the accumulator into the partial-sum buffer (summed in a fixed order afterwards).
constexpr int TW_STAGES = 3, TW_SLAB = 16384;
constexpr int TW_SMEM = TW_STAGES * 2 * TW_SLAB + 16 * 8 + 16 + 1024;
constexpr int TW_THREADS = 9 * 32;
__global__ void __launch_bounds__(TW_THREADS, 2) k_tc_wgrad(const float* __restrict__ X, int ldx, int K, const int32_t* __restrict__ idx,
                                                           const float* __restrict__ dY, int ldy, int N, int64_t R, int64_t rows_per_chunk,
                                                           float* __restrict__ part, WatchArgs wd) {
  using namespace tcx;
  extern __shared__ uint8_t tw_raw[];
  const uint32_t raw = smem_u32(tw_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = tw_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TW_STAGES * 2 * TW_SLAB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };      // FULL[s] = s, EMPTY[s] = 3 + s, DONE = 6
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  bool wd_dead = false;
  if (tid == 0) {
    for (int s2 = 0; s2 < TW_STAGES; s2++) { mbar_init(BAR(s2), 8); mbar_init(BAR(3 + s2), 1); }
    mbar_init(BAR(6), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_chunk;
  const int64_t r_end = r_begin + rows_per_chunk < R ? r_begin + rows_per_chunk : R;
  const int nslab = r_end > r_begin ? (int)((r_end - r_begin + 63) / 64) : 0;
  if (warp < 8) {
    // ---- producers: warp w builds chunks (8 reduction rows each) w of both slabs: items (operand, m group of 32) x 4
    for (int sl = 0; sl < nslab && !wd_dead; sl++) {
      const int st = sl % TW_STAGES;
      mbar_wait_w(BAR(3 + st), ((sl / TW_STAGES) & 1) ^ 1, wd_dead, wd);
      uint8_t* A = sm + st * 2 * TW_SLAB;
      uint8_t* Bm = A + TW_SLAB;
      const int64_t rb = r_begin + (int64_t)sl * 64 + 8 * warp;      // this warp's 8 reduction rows (chunk `warp` of the slab)
      int64_t xr[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int64_t r = rb + j;
        xr[j] = r < r_end ? (idx ? (int64_t)__ldg(idx + r) : r) : -1;
      }
#pragma unroll
      for (int mg = 0; mg < 4; mg++) {
        const int m = 32 * mg + lane;      // operand row: column k0 + m of X, column n0 + m of dY
        float xv[8], yv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const bool ok = xr[j] >= 0;
          xv[j] = (ok && k0 + m < K) ? __ldg(X + (size_t)xr[j] * ldx + k0 + m) : 0.f;
          yv[j] = (ok && n0 + m < N) ? __ldg(dY + (size_t)(rb + j) * ldy + n0 + m) : 0.f;
        }
        const uint32_t off = (uint32_t)(m * 128 + ((warp ^ (m & 7)) << 4));      // sw_off(m, 8 warp): chunk `warp` of row m
        *reinterpret_cast<uint4*>(A + off) = make_uint4(pack_bf16(xv[0], xv[1]), pack_bf16(xv[2], xv[3]), pack_bf16(xv[4], xv[5]), pack_bf16(xv[6], xv[7]));
        *reinterpret_cast<uint4*>(Bm + off) = make_uint4(pack_bf16(yv[0], yv[1]), pack_bf16(yv[2], yv[3]), pack_bf16(yv[4], yv[5]), pack_bf16(yv[6], yv[7]));
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(st));
    }
    // ---- epilogue (warps 0-3, TMEM lane quadrant = warp): accumulator -> part[chunk][k][n]
    if (warp < 4 && nslab > 0) {
      mbar_wait_w(BAR(6), 0, wd_dead, wd);
      tc_fence_after();
      const int m = 32 * warp + lane;
      const uint32_t lane_base = ((uint32_t)(32 * warp)) << 16;
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        uint32_t v[32];
        TC_LD32(tmem + lane_base + 32 * c, v);
        tc_wait_ld();
        if (k0 + m < K) {
          float* o = part + ((size_t)blockIdx.z * K + k0 + m) * N + n0 + 32 * c;
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (n0 + 32 * c + j < N) o[j] = __uint_as_float(v[j]);
        }
      }
    } else if (warp < 4) {      // empty chunk: zeros
      const int m = 32 * warp + lane;
      if (k0 + m < K)
        for (int j = 0; j < 128; j++)
          if (n0 + j < N) part[((size_t)blockIdx.z * K + k0 + m) * N + n0 + j] = 0.f;
    }
  } else {
    // ---- MMA issuer
    for (int sl = 0; sl < nslab && !wd_dead; sl++) {
      const int st = sl % TW_STAGES;
      mbar_wait_w(BAR(st), (sl / TW_STAGES) & 1, wd_dead, wd);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = umma_desc(base + st * 2 * TW_SLAB), bd = umma_desc(base + st * 2 * TW_SLAB + TW_SLAB);
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) mma_ss(tmem, ad + 2 * k4, bd + 2 * k4, IDESC, (sl > 0 || k4 > 0) ? 1u : 0u);
        tc_commit(BAR(3 + st));
        if (sl == nslab - 1) tc_commit(BAR(6));
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int chunks_for(int64_t R) {
  int64_t c = (R + 2047) / 2048;
  return (int)(c < 1 ? 1 : (c > 256 ? 256 : c));
}
int train_ws(gnb_ctx* ctx, size_t bytes, float** out) {
  if (ctx->train_ws_bytes < bytes) {
    if (ctx->train_ws) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->train_ws); ctx->train_ws = nullptr; ctx->train_ws_bytes = 0; }
    if (cudaMalloc(&ctx->train_ws, bytes) != cudaSuccess) {
      cudaGetLastError();
      gnb_set_error("training workspace: cudaMalloc(%zu) failed", bytes);
      return GNB_ERR_OOM;
    }
    ctx->train_ws_bytes = bytes;
  }
  *out = (float*)ctx->train_ws;
  return GNB_OK;
}

}  // namespace

extern "C" int gnb_op_linear(gnb_ctx* ctx, const gnb_lin_args* p) {
  GNB_CHECK(ctx && p && p->nsrc >= 1 && p->nsrc <= 3 && p->nadd >= 0 && p->nadd <= 4, "gnb_op_linear: bad arguments");
  GNB_CUDA(cudaSetDevice(ctx->device));
  LinArgs a{};
  a.R = p->R; a.Nout = p->Nout; a.ldw = p->ldw; a.nsrc = p->nsrc; a.bias = p->bias; a.nadd = p->nadd; a.relu = p->relu; a.out = p->out; a.ldo = p->ldo;
  for (int s = 0; s < p->nsrc; s++) {
    a.src[s].x = p->src[s].x; a.src[s].d = p->src[s].d; a.src[s].ldx = p->src[s].ldx; a.src[s].W = p->src[s].W;
    a.src[s].gamma = p->src[s].gamma; a.src[s].beta = p->src[s].beta; a.src[s].eps = p->src[s].eps; a.src[s].eps_mode = p->src[s].eps_mode;
  }
  for (int j = 0; j < p->nadd; j++) { a.add[j].a = p->add[j].a; a.add[j].idx = p->add[j].idx; a.add[j].lda = p->add[j].lda; }
  if ((p->precision == GNB_PREC_BF16 || p->precision == GNB_PREC_AUTO) && tc_lin_supported(a)) {
    // bf16 operands / fp32 accumulation on the tensor cores (csrc/tc_gemm.cu), weights packed per call (they change every step)
    ctx->tc_lin_nocache = true;
    const int rc = launch_linear_tc(ctx, a);
    ctx->tc_lin_nocache = false;
    return rc;
  }
  return launch_linear_fp32(ctx, a);
}
extern "C" int gnb_op_segsum(gnb_ctx* ctx, const float* x, int D, const int32_t* ptr, int64_t S, const int32_t* perm, float* out) {
  GNB_CHECK(ctx && x && ptr && out && D > 0, "gnb_op_segsum: bad arguments");
  if (S <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  if (!perm) return launch_segsum(ctx, x, D, ptr, S, out);
  Launch L(ctx, "train_segsum");
  k_segsum_perm<<<ceil_div(S * 32, 256), 256, 0, ctx->stream>>>(x, D, ptr, perm, S, out);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_layernorm(gnb_ctx* ctx, const float* x, int64_t R, int D, const float* gamma, const float* beta, float eps, int eps_mode,
                                float* y) {
  GNB_CHECK(ctx && x && y && gamma && beta && D > 0, "gnb_op_layernorm: bad arguments");
  if (R <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  Launch L(ctx, "train_ln");
  k_ln_fwd<<<ceil_div(R * 32, 256), 256, 0, ctx->stream>>>(x, R, D, gamma, beta, eps, eps_mode, y);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_layernorm_bwd(gnb_ctx* ctx, const float* x, const float* g, int64_t R, int D, const float* gamma, float eps, int eps_mode,
                                    float* dx, float* gxhat) {
  GNB_CHECK(ctx && x && g && gamma && dx && gxhat && D > 0, "gnb_op_layernorm_bwd: bad arguments");
  if (R <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  Launch L(ctx, "train_ln_bwd");
  k_ln_bwd<<<ceil_div(R * 32, 256), 256, 0, ctx->stream>>>(x, g, R, D, gamma, eps, eps_mode, dx, gxhat);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_wgrad(gnb_ctx* ctx, const float* X, int ldx, int K, const int32_t* idx, const float* dY, int ldy, int N, int64_t R,
                            float* dW, int ldw, int precision) {
  GNB_CHECK(ctx && X && dY && dW && K > 0 && N > 0, "gnb_op_wgrad: bad arguments");
  if (R <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  // (narrow blocks too - e.g. the decoder's 768 x 3 edge Dense over all E rows: a mostly empty 128 x 128 tile on the tensor cores
  // still beats the fp32 reduction by an order of magnitude)
  if ((precision == GNB_PREC_BF16 || precision == GNB_PREC_AUTO) && R >= 4096 && (int64_t)K * N >= 256) {
    // tensor cores: enough row chunks to fill the GPU with (K / 128) x (N / 128) x chunks CTAs, two CTAs per SM
    const int tiles = ceil_div(K, 128) * ceil_div(N, 128);
    int64_t chunks = (2 * (int64_t)ctx->sm_count + tiles - 1) / tiles;
    const int64_t max_chunks = (R + 511) / 512;      // at least 512 rows (8 slabs) per chunk
    chunks = chunks < 1 ? 1 : (chunks > max_chunks ? max_chunks : chunks);
    int64_t rpc = (R + chunks - 1) / chunks;
    rpc = (rpc + 63) / 64 * 64;
    chunks = (R + rpc - 1) / rpc;
    float* part = nullptr;
    GNB_TRY(train_ws(ctx, (size_t)chunks * K * N * sizeof(float), &part));
    if (ctx_first(ctx, ONCE_TC_WGRAD)) GNB_CUDA(cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
    {
      Launch L(ctx, "train_wgrad_tc", 4.0 * R * (K + N), 2.0 * R * K * N);
      k_tc_wgrad<<<dim3((unsigned)ceil_div(K, 128), (unsigned)ceil_div(N, 128), (unsigned)chunks), TW_THREADS, TW_SMEM, ctx->stream>>>(
          X, ldx, K, idx, dY, ldy, N, R, rpc, part, ctx_watch(ctx));
      GNB_CUDA(cudaGetLastError());
    }
    Launch L2(ctx, "train_reduce");
    k_reduce_parts<<<ceil_div((int64_t)K * N, 256), 256, 0, ctx->stream>>>(part, (int)chunks, K, N, dW, ldw);
    GNB_CUDA(cudaGetLastError());
    return GNB_OK;
  }
  const int chunks = chunks_for(R);
  const int64_t rpc = (R + chunks - 1) / chunks;
  float* part = nullptr;
  GNB_TRY(train_ws(ctx, (size_t)chunks * K * N * sizeof(float), &part));
  {
    Launch L(ctx, "train_wgrad", 4.0 * R * (K + N), 2.0 * R * K * N);
    k_wgrad_part<<<dim3((unsigned)ceil_div(K, WG_T), (unsigned)ceil_div(N, WG_T), (unsigned)chunks), 256, 0, ctx->stream>>>(X, ldx, K, idx, dY, ldy, N, R, rpc, part);
    GNB_CUDA(cudaGetLastError());
  }
  Launch L2(ctx, "train_reduce");
  k_reduce_parts<<<ceil_div((int64_t)K * N, 256), 256, 0, ctx->stream>>>(part, chunks, K, N, dW, ldw);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_colsum(gnb_ctx* ctx, const float* X, int ldx, int D, int64_t R, float* out) {
  GNB_CHECK(ctx && X && out && D > 0, "gnb_op_colsum: bad arguments");
  if (R <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  const int chunks = chunks_for(R);
  const int64_t rpc = (R + chunks - 1) / chunks;
  float* part = nullptr;
  GNB_TRY(train_ws(ctx, (size_t)chunks * D * sizeof(float), &part));
  {
    Launch L(ctx, "train_colsum");
    k_colsum_part<<<dim3((unsigned)ceil_div(D, 32), (unsigned)chunks), 256, 0, ctx->stream>>>(X, ldx, D, R, rpc, part);
    GNB_CUDA(cudaGetLastError());
  }
  Launch L2(ctx, "train_reduce");
  k_reduce_parts<<<ceil_div((int64_t)D, 256), 256, 0, ctx->stream>>>(part, chunks, 1, D, out, D);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_relu_mask(gnb_ctx* ctx, float* t, const float* h, int64_t n) {
  GNB_CHECK(ctx && t && h, "gnb_op_relu_mask: bad arguments");
  if (n <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  Launch L(ctx, "train_elementwise");
  k_relu_mask<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(t, h, n);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_gather_add(gnb_ctx* ctx, float* out, const float* a, const float* b1, const int32_t* idx1, const float* b2,
                                 const int32_t* idx2, int64_t R, int D) {
  GNB_CHECK(ctx && out && D > 0, "gnb_op_gather_add: bad arguments");
  if (R <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  Launch L(ctx, "train_elementwise");
  k_gather_add<<<ceil_div(R * D, 256), 256, 0, ctx->stream>>>(out, a, b1, idx1, b2, idx2, R, D);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_transpose(gnb_ctx* ctx, const float* in, int rows, int cols, int ld_in, float* out) {
  GNB_CHECK(ctx && in && out && rows > 0 && cols > 0, "gnb_op_transpose: bad arguments");
  GNB_CUDA(cudaSetDevice(ctx->device));
  Launch L(ctx, "train_elementwise");
  k_transpose<<<dim3((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), dim3(32, 8), 0, ctx->stream>>>(in, rows, cols, ld_in, out);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
extern "C" int gnb_op_adamw(gnb_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, int step) {
  GNB_CHECK(ctx && p && g && m && v && step >= 1, "gnb_op_adamw: bad arguments");
  if (n <= 0) return GNB_OK;
  GNB_CUDA(cudaSetDevice(ctx->device));
  Launch L(ctx, "train_adamw");
  k_adamw<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.f - powf(beta1, (float)step),
                                                     1.f - powf(beta2, (float)step));
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
// device pointers of the lowered index (for host-side orchestration of the backward pass; owned by the graph)
extern "C" int gnb_graph_device_index(const gnb_graph* g, const int32_t** edge_src, const int32_t** edge_dst, const int32_t** edge_graph,
                                      const int32_t** node_graph, const int32_t** graph_edge_ptr, const int32_t** graph_node_ptr,
                                      const int32_t** node_in_ptr) {
  GNB_CHECK(g, "gnb_graph_device_index: null graph");
  if (edge_src) *edge_src = g->edge_src;
  if (edge_dst) *edge_dst = g->edge_dst;
  if (edge_graph) *edge_graph = g->edge_graph;
  if (node_graph) *node_graph = g->node_graph;
  if (graph_edge_ptr) *graph_edge_ptr = g->graph_edge_ptr;
  if (graph_node_ptr) *graph_node_ptr = g->graph_node_ptr;
  if (node_in_ptr) *node_in_ptr = g->node_in_ptr;
  return GNB_OK;
}
