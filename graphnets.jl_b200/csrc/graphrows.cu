// Graph-level (B rows) stages of a GNCore with 128-wide features, fp32 CUDA cores.
//
// A batch has few graphs (B << N << E), so these stages are latency- not throughput-bound: a CTA owns
// RT = 8 graphs and all output columns, keeps its rows in shared memory and streams the weights from L2.
// One launch per GNCore (k_graph_post, which also emits the per-graph rows of the next core; k_graph_pre only
// runs in front of the first core of a chain) replaces ten small GEMM / segmented-sum launches:
//   k_graph_pre :  P_ue = LN1(u) W_eu + c_e          per-graph row of the edge update   (src/edgefninput.jl:6)
//                  P_un = LN1(u) W_nu + c_n          per-graph row of the node update   (src/nodefninput.jl:5)
//   k_graph_post:  s_e = sum of the graph's edges (src/graphfninput.jl:3) and s_v = sum of its updated nodes (:4), both by
//                  linearity from the partial rows the aggregate / node kernels emit (a handful of rows per graph)
//                  h_u = W_g [s_e ; s_v ; LN1(u)] + b_g                                 (src/gnblock.jl:67)
//                  y_u = (u + h_u) + W2 relu(W1 LN2(u) + b1) + b2                       (src/gncore.jl:56-68)
#include "kernels.cuh"
#include "graphrows.cuh"
#include <stdlib.h>

namespace {

constexpr int H = 128;
constexpr int RT = 8;   // graphs per CTA

__device__ __forceinline__ float ln_rstd(float var, float eps, int mode) {
  if (mode == GNB_EPS_SQRT_VAR_EPS2) return 1.0f / sqrtf(var + eps * eps);
  if (mode == GNB_EPS_STD_PLUS_EPS) return 1.0f / (sqrtf(var) + eps);
  return 1.0f / sqrtf(var + eps);
}

// LayerNorm of RT rows held in xs[RT][H] (warp w normalises rows w, w+4): out = gamma (x - mu) rstd + beta
__device__ __forceinline__ void ln_rows(const float (*xs)[H], float (*out)[H], const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float eps, int mode, int warp, int lane) {
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  for (int r = warp; r < RT; r += 4) {
    const float4 v = *reinterpret_cast<const float4*>(&xs[r][4 * lane]);
    float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * (1.0f / H);
    const float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
    float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = ln_rstd(q * (1.0f / H), eps, mode);
    *reinterpret_cast<float4*>(&out[r][4 * lane]) =
        make_float4(dx * rs * g.x + b.x, dy * rs * g.y + b.y, dz * rs * g.z + b.z, dw * rs * g.w + b.w);
  }
}

// acc[r] += sum_k xs[r][k] * W[k*ldw + n]   for k in [0, K)   (W row-major [k][n], coalesced over the CTA)
template <int K>
__device__ __forceinline__ void gemv_rows(const float (*xs)[K], const float* __restrict__ W, int ldw, int n, float* acc) {
#pragma unroll 4
  for (int k = 0; k < K; k += 4) {
    const float w0 = __ldg(W + (size_t)(k + 0) * ldw + n), w1 = __ldg(W + (size_t)(k + 1) * ldw + n);
    const float w2 = __ldg(W + (size_t)(k + 2) * ldw + n), w3 = __ldg(W + (size_t)(k + 3) * ldw + n);
#pragma unroll
    for (int r = 0; r < RT; r++) {
      const float4 x = *reinterpret_cast<const float4*>(&xs[r][k]);
      acc[r] = fmaf(x.x, w0, acc[r]);
      acc[r] = fmaf(x.y, w1, acc[r]);
      acc[r] = fmaf(x.z, w2, acc[r]);
      acc[r] = fmaf(x.w, w3, acc[r]);
    }
  }
}

__global__ void __launch_bounds__(128) k_graph_pre(const GraphPreArgs a) {
  __shared__ __align__(16) float xs[RT][H];
  __shared__ __align__(16) float xa[RT][H];
  const int n = threadIdx.x, warp = n >> 5, lane = n & 31;
  const int64_t g0 = (int64_t)blockIdx.x * RT;
  for (int r = 0; r < RT; r++) {
    const int64_t g = g0 + r < a.B ? g0 + r : a.B - 1;
    xs[r][n] = a.xg[(size_t)g * H + n];
  }
  __syncthreads();
  ln_rows(xs, xa, a.gamma, a.beta, a.eps, a.eps_mode, warp, lane);
  __syncthreads();
  float ae[RT], an[RT];
  const float ce = a.ce[n], cn = a.cn[n];
#pragma unroll
  for (int r = 0; r < RT; r++) { ae[r] = ce; an[r] = cn; }
  gemv_rows<H>(xa, a.Weu, H, n, ae);
  gemv_rows<H>(xa, a.Wnu, H, n, an);
#pragma unroll
  for (int r = 0; r < RT; r++) {
    if (g0 + r < a.B) {
      a.Pue[(size_t)(g0 + r) * H + n] = ae[r];
      a.Pun[(size_t)(g0 + r) * H + n] = an[r];
    }
  }
}

// K-slice of gemv_rows: acc[r] += sum_{k in [k0, k0 + KS)} xs[r][k] * W[k*ldw + n]     (xs rows have stride LD floats)
template <int KS, int LD, int RT>
__device__ __forceinline__ void gemv_slice(const float* xs, int k0, const float* __restrict__ W, int ldw, int n, float* acc) {
#pragma unroll 4
  for (int k = k0; k < k0 + KS; k += 4) {
    const float w0 = __ldg(W + (size_t)(k + 0) * ldw + n), w1 = __ldg(W + (size_t)(k + 1) * ldw + n);
    const float w2 = __ldg(W + (size_t)(k + 2) * ldw + n), w3 = __ldg(W + (size_t)(k + 3) * ldw + n);
#pragma unroll
    for (int r = 0; r < RT; r++) {
      const float4 x = *reinterpret_cast<const float4*>(xs + r * LD + k);
      acc[r] = fmaf(x.x, w0, acc[r]);
      acc[r] = fmaf(x.y, w1, acc[r]);
      acc[r] = fmaf(x.z, w2, acc[r]);
      acc[r] = fmaf(x.w, w3, acc[r]);
    }
  }
}

// LayerNorm of row r held in xs[r][H] by one warp: out = gamma (x - mu) rstd + beta
__device__ __forceinline__ void ln_row(const float* xs, float* out, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float eps, int mode, int lane) {
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 v = *reinterpret_cast<const float4*>(xs + 4 * lane);
  float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s * (1.0f / H);
  const float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
  float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rs = ln_rstd(q * (1.0f / H), eps, mode);
  *reinterpret_cast<float4*>(out + 4 * lane) = make_float4(dx * rs * g.x + b.x, dy * rs * g.y + b.y, dz * rs * g.z + b.z, dw * rs * g.w + b.w);
}

// 512 threads: thread (n = tid & 127, ks = tid >> 7).  The stages are latency-bound chains of L2 weight loads, so every GEMV is
// split 4 ways along K (partials combined in a fixed order through shared memory: deterministic) and the node sums 4 ways over
// the graphs of the CTA.
constexpr int GP_THREADS = 512;
template <int RT>
__global__ void __launch_bounds__(GP_THREADS, 2) k_graph_post(const GraphPostArgs a) {
  // dynamic shared memory (RT x 11 H floats: 45 KB at RT = 8, 90 KB at RT = 16)
  extern __shared__ __align__(16) float gp_sm[];
  float(*xs)[H] = reinterpret_cast<float(*)[H]>(gp_sm);                          // u
  float(*cat)[3 * H] = reinterpret_cast<float(*)[3 * H]>(gp_sm + RT * H);        // [s_e | s_v | LN1(u)]   (s_v slot first holds gamma . sum v^)
  float(*xb)[H] = reinterpret_cast<float(*)[H]>(gp_sm + RT * 4 * H);             // LN2(u); later y_u, LN1'(y_u)
  float(*svg)[H] = reinterpret_cast<float(*)[H]>(gp_sm + RT * 5 * H);            // sum of the node addends
  float(*seg)[H] = reinterpret_cast<float(*)[H]>(gp_sm + RT * 6 * H);            // sum of the edge addends
  float(*hid)[4 * H] = reinterpret_cast<float(*)[4 * H]>(gp_sm + RT * 7 * H);    // FFN hidden; before that the scratch of the K-split reductions
  float(*red)[RT][H] = reinterpret_cast<float(*)[RT][H]>(&hid[0][0]);      // [4][RT][H] == sizeof(hid)
  const int tid = threadIdx.x, n = tid & (H - 1), ks = tid >> 7, warp = tid >> 5, lane = tid & 31;
  const int64_t g0 = (int64_t)blockIdx.x * RT;
  // ---- ordered sums of the graph's partial rows (per (16-node block, graph) run: a handful per graph): deterministic, no
  // atomics; group ks owns RT / 4 graphs of the CTA
#pragma unroll 1
  for (int r = ks * (RT / 4); r < (ks + 1) * (RT / 4); r++) {
    const int64_t g = g0 + r < a.B ? g0 + r : a.B - 1;
    xs[r][n] = a.xg[(size_t)g * H + n];
    const int p0 = a.graph_npart_ptr[g], p1 = a.graph_npart_ptr[g + 1];
    float se = 0.f, sg = 0.f, sh = 0.f, sn = 0.f;
    for (int p = p0; p < p1; p++) {
      se += a.SEpart[(size_t)p * H + n]; sg += a.SGpart[(size_t)p * H + n];
      sh += a.Vpart[(size_t)p * H + n]; sn += a.Npart[(size_t)p * H + n];
    }
    // sums of the normalised rows, scaled by the LayerNorm scales the packed bf16 weights carry; the fp32 weights follow below
    cat[r][n] = se * a.g1e[n];
    seg[r][n] = sg;
    cat[r][H + n] = sh * a.g1n[n];
    svg[r][n] = sn;
  }
  __syncthreads();
  for (int i = warp; i < 2 * RT; i += GP_THREADS / 32) {
    if (i < RT) ln_row(xs[i], &cat[i][2 * H], a.g1, a.b1ln, a.eps1, a.eps_mode1, lane);
    else ln_row(xs[i - RT], xb[i - RT], a.g2, a.b2ln, a.eps2, a.eps_mode2, lane);
  }
  // ---- s_e = W_ee (gamma_e . sum ê) + sum of the edge addends;  s_v = W_nv (gamma_n . sum v^) + sum of the node addends
  {
    float t[RT], te[RT];
#pragma unroll
    for (int r = 0; r < RT; r++) { t[r] = 0.f; te[r] = 0.f; }
    gemv_slice<H / 4, 3 * H, RT>(&cat[0][0], ks * (H / 4), a.Wee, H, n, te);
    gemv_slice<H / 4, 3 * H, RT>(&cat[0][H], ks * (H / 4), a.Wnv, H, n, t);
#pragma unroll
    for (int r = 0; r < RT; r++) red[ks][r][n] = t[r];
    __syncthreads();      // also publishes the LayerNorm rows
    if (ks == 0) {
#pragma unroll
      for (int r = 0; r < RT; r++) cat[r][H + n] = svg[r][n] + (((red[0][r][n] + red[1][r][n]) + red[2][r][n]) + red[3][r][n]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RT; r++) red[ks][r][n] = te[r];
    __syncthreads();
    if (ks == 0) {
#pragma unroll
      for (int r = 0; r < RT; r++) cat[r][n] = seg[r][n] + (((red[0][r][n] + red[1][r][n]) + red[2][r][n]) + red[3][r][n]);
    }
    __syncthreads();
  }
  // ---- h_u = W_g [s_e ; s_v ; LN1(u)] + b_g          (K = 3H, slice = 96)
  float hu[RT];
  {
    float t[RT];
#pragma unroll
    for (int r = 0; r < RT; r++) t[r] = 0.f;
    gemv_slice<3 * H / 4, 3 * H, RT>(&cat[0][0], ks * (3 * H / 4), a.Wg, H, n, t);
#pragma unroll
    for (int r = 0; r < RT; r++) red[ks][r][n] = t[r];
    __syncthreads();
    const float bg = a.bg[n];
#pragma unroll
    for (int r = 0; r < RT; r++) hu[r] = bg + (((red[0][r][n] + red[1][r][n]) + red[2][r][n]) + red[3][r][n]);
    __syncthreads();      // red (== hid) is rewritten below
  }
  // ---- FFN hidden: relu(W1 LN2(u) + b1), 4H wide: thread (n, ks) owns hidden unit ks H + n
  {
    float hc[RT];
    const float b1 = a.b1[ks * H + n];
#pragma unroll
    for (int r = 0; r < RT; r++) hc[r] = b1;
    gemv_slice<H, H, RT>(&xb[0][0], 0, a.W1 + ks * H, 4 * H, n, hc);
#pragma unroll
    for (int r = 0; r < RT; r++) hid[r][ks * H + n] = fmaxf(hc[r], 0.f);
  }
  __syncthreads();
  {
    float t[RT];
#pragma unroll
    for (int r = 0; r < RT; r++) t[r] = 0.f;
    gemv_slice<H, 4 * H, RT>(&hid[0][0], ks * H, a.W2, H, n, t);
    float(*red2)[RT][H] = reinterpret_cast<float(*)[RT][H]>(&cat[0][0]);      // [3][RT][H] == sizeof(cat): slices 1..3
    if (ks > 0) {
#pragma unroll
      for (int r = 0; r < RT; r++) red2[ks - 1][r][n] = t[r];
    }
    __syncthreads();
    if (ks == 0) {
      const float b2 = a.b2[n];
#pragma unroll
      for (int r = 0; r < RT; r++) {
        const float f = b2 + (((t[r] + red2[0][r][n]) + red2[1][r][n]) + red2[2][r][n]);
        const float y = (xs[r][n] + hu[r]) + f;
        if (g0 + r < a.B) a.yg[(size_t)(g0 + r) * H + n] = y;
        xb[r][n] = y;
      }
    }
  }
  if (a.next_Pue == nullptr) return;      // uniform
  // ---- per-graph rows of the NEXT core (k_graph_pre fused): P_ue = LN1'(y_u) W_eu' + c_e', P_un likewise
  __syncthreads();
  for (int i = warp; i < RT; i += GP_THREADS / 32) ln_row(xb[i], xs[i], a.next_gamma, a.next_beta, a.next_eps, a.next_eps_mode, lane);
  __syncthreads();
  {
    float te[RT];
#pragma unroll
    for (int r = 0; r < RT; r++) te[r] = 0.f;
    // groups 0, 1: the two K halves of P_ue; groups 2, 3: of P_un
    const float* W = ks < 2 ? a.next_Weu : a.next_Wnu;
    gemv_slice<H / 2, H, RT>(&xs[0][0], (ks & 1) * (H / 2), W, H, n, te);
#pragma unroll
    for (int r = 0; r < RT; r++) red[ks][r][n] = te[r];
    __syncthreads();
    if ((ks & 1) == 0) {
      const float c = ks == 0 ? a.next_ce[n] : a.next_cn[n];
      float* out = ks == 0 ? a.next_Pue : a.next_Pun;
#pragma unroll
      for (int r = 0; r < RT; r++)
        if (g0 + r < a.B) out[(size_t)(g0 + r) * H + n] = c + (red[ks][r][n] + red[ks + 1][r][n]);
    }
  }
}

}  // namespace

int launch_graph_pre(gnb_ctx* ctx, const GraphPreArgs& a) {
  if (a.B <= 0) return GNB_OK;
  Launch L(ctx, "graph_pre", 4.0 * a.B * 3 * H + 8.0 * H * H, 4.0 * a.B * H * H);
  k_graph_pre<<<ceil_div(a.B, RT), 128, 0, ctx->stream>>>(a);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

int launch_graph_post(gnb_ctx* ctx, const GraphPostArgs& a, int64_t N) {
  if (a.B <= 0) return GNB_OK;
  (void)N;
  Launch L(ctx, "graph_post", 8.0 * a.B * H + 4.0 * 12 * H * H, 24.0 * a.B * H * H);
  // 16 graphs per CTA: 256 CTAs for 4096 graphs = ONE wave at two resident CTAs per SM and half the weight traffic from L2
  // (8 graphs per CTA: 512 CTAs = 1.7 waves); small batches keep 8 so that every SM has work
  static const int rt_env = getenv("GNB_GRAPH_RT") ? atoi(getenv("GNB_GRAPH_RT")) : 0;
  const int rt = rt_env ? rt_env : (a.B >= 16 * ctx->sm_count ? 16 : 8);
  if (ctx_first(ctx, ONCE_GRAPH_POST)) {
    GNB_CUDA(cudaFuncSetAttribute(k_graph_post<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 11 * H * 4));
    GNB_CUDA(cudaFuncSetAttribute(k_graph_post<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 11 * H * 4));
  }
  if (rt == 16) k_graph_post<16><<<ceil_div(a.B, 16), GP_THREADS, 16 * 11 * H * 4, ctx->stream>>>(a);
  else k_graph_post<8><<<ceil_div(a.B, 8), GP_THREADS, 8 * 11 * H * 4, ctx->stream>>>(a);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
