// fp32 CUDA-core path: fused linear layer + deterministic segmented sums.
//
// One kernel template covers every Dense of the reference forward:
//   out = act( sum_s LN_s(x_s) W_s + bias + sum_j add_j[idx_j] )
// The reference builds [e | v_src | v_dst | u] with dense batched_mul gathers and a
// materialised vcat (src/edgefninput.jl:1-8) and then applies one Dense (src/gnblock.jl:65).
// Dense is linear, so W [e;v_s;v_r;u] = W_e e + (W_s v)[src] + (W_r v)[dst] + (W_u u)[graph]:
// nodes / graphs are projected once (N, B << E rows) and the gather becomes an indexed ADD in
// the epilogue of the per-edge GEMM - the concat never exists.  LayerNorm (src/gngraphnorm.jl)
// is applied on the fly while the A tile is staged in shared memory.
#include "kernels.cuh"

namespace {

constexpr int TK = 16;      // K tile; two shared-memory buffers of TK rows

__device__ __forceinline__ float ln_rstd(float var, float eps, int mode) {
  if (mode == GNB_EPS_SQRT_VAR_EPS2) return 1.0f / sqrtf(var + eps * eps);
  if (mode == GNB_EPS_STD_PLUS_EPS) return 1.0f / (sqrtf(var) + eps);
  return 1.0f / sqrtf(var + eps);
}

// Thread tile: QM x QN quads of 4x4; quads are strided by TM/QM rows and TN/QN columns so
// every shared-memory read is a conflict-free LDS.128.
#ifndef GNB_LIN_MINB
#define GNB_LIN_MINB 2      /* <= 128 registers: two CTAs per SM (one CTA of 153 registers measured 40 % slower) */
#endif
template <int TM, int TN, int QM, int QN>
__global__ void __launch_bounds__((TM / (4 * QM)) * (TN / (4 * QN)), GNB_LIN_MINB)
k_linear(const LinArgs a) {
  constexpr int TXN = TN / (4 * QN);
  constexpr int TYN = TM / (4 * QM);
  constexpr int NT = TXN * TYN;
  constexpr int LDA = TM + 4;
  constexpr int LDB = TN + 4;
  __shared__ __align__(16) float As[2 * TK * LDA];
  __shared__ __align__(16) float Ws[2 * TK * LDB];
  __shared__ float s_mu[3][TM];
  __shared__ float s_rs[3][TM];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const int col0 = blockIdx.y * TN;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;

  // ---- LayerNorm statistics of this CTA's rows (two-pass, fp32) -------------------------
  for (int s = 0; s < a.nsrc; s++) {
    const LinSrc& S = a.src[s];
    if (S.gamma == nullptr || S.d == 0) continue;
    if (S.d <= 32) {
      // narrow rows (config 1 / 2 widths): one row per THREAD, the row in registers - one round trip to memory for the whole
      // tile.  (A warp per row costs TM / NW dependent round trips x 2 passes: 16 us of a 27 us launch at config 2.)
      for (int r = tid; r < TM; r += NT) {
        const int64_t row = row0 + r;
        float mu = 0.f, rs = 0.f;
        if (row < a.R) {
          const float* xr = S.x + (size_t)row * S.ldx;
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; k++) v[k] = k < S.d ? xr[k] : 0.f;
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < 32; k++) sum += v[k];
          mu = sum / (float)S.d;
          float sq = 0.f;
#pragma unroll
          for (int k = 0; k < 32; k++) {
            const float t = k < S.d ? v[k] - mu : 0.f;
            sq += t * t;
          }
          rs = ln_rstd(sq / (float)S.d, S.eps, S.eps_mode);
        }
        s_mu[s][r] = mu;
        s_rs[s][r] = rs;
      }
      continue;
    }
    if ((S.d & 3) == 0 && S.d <= 512 && (S.ldx & 3) == 0 && ((((uintptr_t)S.x) & 15) == 0)) {
      // wide rows: 8 lanes per row, FOUR rows of a warp in flight, the row in registers (float4 chunks c = 8 i + l8), one pass
      // over memory and 3 shuffle levels per statistic.  (A warp per row with two dependent passes cost 2 TM / NW round trips:
      // 19 k cycles per CTA at 128-wide layers - more than the whole K loop.)
      const int rr = lane >> 3, l8 = lane & 7;
      const int nchunk = S.d >> 2;
      for (int r4 = warp * 4; r4 < TM; r4 += NW * 4) {
        const int r = r4 + rr;
        const int64_t row = row0 + r;
        const bool ok = r < TM && row < a.R;
        const float4* xr = reinterpret_cast<const float4*>(S.x + (size_t)(ok ? row : row0) * S.ldx);
        float4 v[16];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int c = 8 * i + l8;
          v[i] = (ok && c < nchunk) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mu = sum / (float)S.d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (8 * i + l8 < nchunk) {
            const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
            sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
          }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (l8 == 0 && r < TM) {
          s_mu[s][r] = ok ? mu : 0.f;
          s_rs[s][r] = ok ? ln_rstd(sq / (float)S.d, S.eps, S.eps_mode) : 0.f;
        }
      }
      continue;
    }
    for (int r = warp; r < TM; r += NW) {
      int64_t row = row0 + r;
      float mu = 0.f, rs = 0.f;
      if (row < a.R) {
        const float* xr = S.x + (size_t)row * S.ldx;
        float sum = 0.f;
        for (int k = lane; k < S.d; k += 32) sum += xr[k];
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mu = sum / (float)S.d;
        float sq = 0.f;
        for (int k = lane; k < S.d; k += 32) {
          float t = xr[k] - mu;
          sq += t * t;
        }
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        rs = ln_rstd(sq / (float)S.d, S.eps, S.eps_mode);
      }
      if (lane == 0) {
        s_mu[s][r] = mu;
        s_rs[s][r] = rs;
      }
    }
  }
  __syncthreads();

  float acc[QM * 4][QN * 4];
#pragma unroll
  for (int i = 0; i < QM * 4; i++)
#pragma unroll
    for (int j = 0; j < QN * 4; j++) acc[i][j] = 0.f;

  // ---- main loop over K tiles of all sources, software pipelined: the global loads of tile t+1 (into registers, LayerNorm
  // applied on the way) are in flight while tile t is multiplied out of shared memory; two shared-memory buffers, one
  // __syncthreads per tile.  (The un-pipelined loop ran at 22 TFLOP/s = 30 % of the fp32 FMA peak at 128-wide layers.)
  int ntile[3] = {0, 0, 0}, T = 0;
  for (int s = 0; s < a.nsrc; s++) { ntile[s] = (a.src[s].d + TK - 1) / TK; T += ntile[s]; }
  constexpr int A_PER = (TM * (TK / 4) + NT - 1) / NT;
  constexpr int W_PER = (TK * (TN / 4) + NT - 1) / NT;
  float4 ra[A_PER], rw[W_PER];
  auto gload = [&](int t) {
    int s = 0;
    while (t >= ntile[s]) { t -= ntile[s]; s++; }
    const int k0 = t * TK;
    const LinSrc& S = a.src[s];
    const bool has_ln = S.gamma != nullptr;
    const bool xvec = ((S.ldx & 3) == 0) && ((((uintptr_t)S.x) & 15) == 0);
    const bool wvec = ((a.ldw & 3) == 0) && ((((uintptr_t)S.W) & 15) == 0) && ((col0 & 3) == 0);
#pragma unroll
    for (int i = 0; i < A_PER; i++) {
      const int idx = tid + i * NT;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (idx < TM * (TK / 4)) {
        const int r = idx / (TK / 4);
        const int k = k0 + (idx % (TK / 4)) * 4;
        const int64_t row = row0 + r;
        if (row < a.R && k < S.d) {
          const float* p = S.x + (size_t)row * S.ldx + k;
          if (xvec && k + 3 < S.d) {
            const float4 t4 = *reinterpret_cast<const float4*>(p);
            v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
              if (k + q < S.d) v[q] = p[q];
          }
          if (has_ln) {
            const float mu = s_mu[s][r], rs = s_rs[s][r];
#pragma unroll
            for (int q = 0; q < 4; q++)
              if (k + q < S.d) v[q] = (v[q] - mu) * rs * S.gamma[k + q] + S.beta[k + q];
          }
        }
      }
      ra[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
#pragma unroll
    for (int i = 0; i < W_PER; i++) {
      const int idx = tid + i * NT;
      float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < TK * (TN / 4)) {
        const int kk = idx / (TN / 4);
        const int n = (idx % (TN / 4)) * 4;
        const int k = k0 + kk;
        if (k < S.d) {
          const float* p = S.W + (size_t)k * a.ldw + col0 + n;
          if (wvec && col0 + n + 3 < a.Nout) {
            t4 = *reinterpret_cast<const float4*>(p);
          } else {
            if (col0 + n + 0 < a.Nout) t4.x = p[0];
            if (col0 + n + 1 < a.Nout) t4.y = p[1];
            if (col0 + n + 2 < a.Nout) t4.z = p[2];
            if (col0 + n + 3 < a.Nout) t4.w = p[3];
          }
        }
      }
      rw[i] = t4;
    }
  };
  auto sstore = [&](int buf) {
    float* Ab = As + buf * (TK * LDA);
    float* Wb = Ws + buf * (TK * LDB);
#pragma unroll
    for (int i = 0; i < A_PER; i++) {
      const int idx = tid + i * NT;
      if (idx < TM * (TK / 4)) {
        const int r = idx / (TK / 4), kk = (idx % (TK / 4)) * 4;      // stored transposed: As[k][r]
        Ab[(kk + 0) * LDA + r] = ra[i].x; Ab[(kk + 1) * LDA + r] = ra[i].y;
        Ab[(kk + 2) * LDA + r] = ra[i].z; Ab[(kk + 3) * LDA + r] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < W_PER; i++) {
      const int idx = tid + i * NT;
      if (idx < TK * (TN / 4)) {
        const int kk = idx / (TN / 4), n = (idx % (TN / 4)) * 4;
        *reinterpret_cast<float4*>(&Wb[kk * LDB + n]) = rw[i];
      }
    }
  };
  if (T > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int t = 0; t < T; t++) {
    if (t + 1 < T) gload(t + 1);
    const float* Ab = As + (t & 1) * (TK * LDA);
    const float* Wb = Ws + (t & 1) * (TK * LDB);
#pragma unroll
    for (int kk = 0; kk < TK; kk++) {
      float av[QM * 4], bv[QN * 4];
#pragma unroll
      for (int q = 0; q < QM; q++) {
        const float4 t4 = *reinterpret_cast<const float4*>(&Ab[kk * LDA + q * (TM / QM) + ty * 4]);
        av[q * 4 + 0] = t4.x; av[q * 4 + 1] = t4.y; av[q * 4 + 2] = t4.z; av[q * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int q = 0; q < QN; q++) {
        const float4 t4 = *reinterpret_cast<const float4*>(&Wb[kk * LDB + q * (TN / QN) + tx * 4]);
        bv[q * 4 + 0] = t4.x; bv[q * 4 + 1] = t4.y; bv[q * 4 + 2] = t4.z; bv[q * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int i = 0; i < QM * 4; i++)
#pragma unroll
        for (int j = 0; j < QN * 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < T) sstore((t + 1) & 1);
    __syncthreads();
  }

  // ---- epilogue: bias + gathered addends + activation + store ---------------------------
  const bool ovec = ((a.ldo & 3) == 0) && ((((uintptr_t)a.out) & 15) == 0);
#pragma unroll
  for (int qi = 0; qi < QM; qi++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int r = qi * (TM / QM) + ty * 4 + i;
      int64_t row = row0 + r;
      if (row >= a.R) continue;
      const float* addrow[4];
      for (int j = 0; j < a.nadd; j++) {
        int64_t ar = a.add[j].idx ? (int64_t)a.add[j].idx[row] : row;
        addrow[j] = a.add[j].a + (size_t)ar * a.add[j].lda;
      }
#pragma unroll
      for (int qj = 0; qj < QN; qj++) {
        int n = col0 + qj * (TN / QN) + tx * 4;
        if (n >= a.Nout) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = acc[qi * 4 + i][qj * 4 + j];
        const bool full = n + 3 < a.Nout;
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (n + j < a.Nout) v[j] += a.bias[n + j];
        }
        for (int t = 0; t < a.nadd; t++) {
          const float* p = addrow[t] + n;
          if (full && ((a.add[t].lda & 3) == 0) && ((((uintptr_t)a.add[t].a) & 15) == 0)) {
            float4 q = *reinterpret_cast<const float4*>(p);
            v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (n + j < a.Nout) v[j] += p[j];
          }
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 4; j++) v[j] = fmaxf(v[j], 0.f);
        }
        float* o = a.out + (size_t)row * a.ldo + n;
        if (full && ovec) {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (n + j < a.Nout) o[j] = v[j];
        }
      }
    }
  }
}

// One warp per segment, lanes over feature columns, rows in ascending order.
__global__ void k_segsum(const float* __restrict__ x, int D, const int32_t* __restrict__ ptr, int64_t S,
                         float* __restrict__ out) {
  int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (seg >= S) return;
  int64_t r0 = ptr[seg], r1 = ptr[seg + 1];
  if (((D & 3) == 0) && ((((uintptr_t)x) & 15) == 0) && ((((uintptr_t)out) & 15) == 0)) {
    int D4 = D >> 2;
    for (int c = lane; c < D4; c += 32) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* p = reinterpret_cast<const float4*>(x) + c;
      int64_t r = r0;
      for (; r + 3 < r1; r += 4) {
        float4 t0 = p[(size_t)r * D4], t1 = p[(size_t)(r + 1) * D4];
        float4 t2 = p[(size_t)(r + 2) * D4], t3 = p[(size_t)(r + 3) * D4];
        s.x += t0.x; s.y += t0.y; s.z += t0.z; s.w += t0.w;
        s.x += t1.x; s.y += t1.y; s.z += t1.z; s.w += t1.w;
        s.x += t2.x; s.y += t2.y; s.z += t2.z; s.w += t2.w;
        s.x += t3.x; s.y += t3.y; s.z += t3.z; s.w += t3.w;
      }
      for (; r < r1; r++) {
        float4 t = p[(size_t)r * D4];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      reinterpret_cast<float4*>(out)[(size_t)seg * D4 + c] = s;
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      float s = 0.f;
      for (int64_t r = r0; r < r1; r++) s += x[(size_t)r * D + c];
      out[(size_t)seg * D + c] = s;
    }
  }
}

}  // namespace

int launch_linear_fp32(gnb_ctx* ctx, const LinArgs& a) {
  if (a.R <= 0 || a.Nout <= 0) return GNB_OK;
  // algorithmic traffic: every source row and gathered addend read once, the output written once,
  // the weights once; flops = 2 R K Nout
  double K = 0, bytes = 0;
  for (int s = 0; s < a.nsrc; s++) K += a.src[s].d;
  bytes = 4.0 * ((double)a.R * (K + (double)a.Nout * (1 + a.nadd)) + K * a.Nout);
  for (int j = 0; j < a.nadd; j++)
    if (a.add[j].idx) bytes += 4.0 * a.R;
  Launch L(ctx, a.Nout > 64 ? "linear_fp32_128x128" : (a.Nout > 16 ? "linear_fp32_64x64" : "linear_fp32_256x16"), bytes,
           2.0 * a.R * K * a.Nout);
  if (a.Nout > 64) {
    dim3 grid((unsigned)ceil_div(a.R, 128), (unsigned)ceil_div(a.Nout, 128));
    k_linear<128, 128, 2, 2><<<grid, 256, 0, ctx->stream>>>(a);
  } else if (a.Nout > 16) {
    dim3 grid((unsigned)ceil_div(a.R, 64), (unsigned)ceil_div(a.Nout, 64));
    k_linear<64, 64, 1, 1><<<grid, 256, 0, ctx->stream>>>(a);
  } else {
    dim3 grid((unsigned)ceil_div(a.R, 256), (unsigned)ceil_div(a.Nout, 16));
    k_linear<256, 16, 1, 1><<<grid, 256, 0, ctx->stream>>>(a);
  }
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

int launch_segsum(gnb_ctx* ctx, const float* x, int D, const int32_t* ptr, int64_t S, float* out) {
  if (S <= 0 || D <= 0) return GNB_OK;
  Launch L(ctx, "segsum_fp32", 0, 0);   // row count is data dependent; bytes accounted by the caller's model
  k_segsum<<<ceil_div(S * 32, 256), 256, 0, ctx->stream>>>(x, D, ptr, S, out);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
