// fp32 CUDA-core kernels for GNBlocks whose input OR output features are narrow (encoder / decoder
// blocks, src/gnblock.jl:63-69).  These layers are HBM-bound: the wide side is streamed once, coalesced,
// the narrow side lives in registers / shared memory.
//
//   k_wide   : out[r][0..N) = sum_p  in_p[idx_p ? idx_p[r] : r][0..d_p) . W_p  (+ bias)      sum_p d_p <= 32
//              the [e | v_src | v_dst | u] concat (src/edgefninput.jl:1-8) is assembled per row in shared
//              memory from the raw narrow inputs and never reaches HBM; one warp computes 8 rows x 128 columns
//              per step with the weight block resident in shared memory.
//   k_zsum   : Z[v] = [ sum_{e->v} e_e ; sum_{e->v} v_src(e) ; deg v_v ; deg u_g ; deg ]   (receiver CSR, ordered)
//              by linearity of Dense:  sum_{e->v} (W_e z_e + b_e) = [W_e ; b_e] Z[v]  - the narrow inputs are
//              aggregated, then transformed, so the wide h_e is never re-read for the edge->node sum
//              (src/nodefninput.jl:3).
//   k_narrow : out[r][0..No) = bias + sum_s x_s[r][0..K_s) . W_s + sum_j add_j[idx_j[r]][0..No)     No <= 8
//              lane owns 4 consecutive k of a row (512 B coalesced row loads), butterfly reduction per output.
#include "kernels.cuh"
#include "smallk.cuh"

namespace {

constexpr int WK_MAX = 32;      // max concatenated input width of k_wide
constexpr int WIDE_ROWS = 32;   // rows staged per warp step
constexpr int WIDE_WARPS = 8;
constexpr int WIDE_MAXP = 5;    // == the size of WideArgs::pc

constexpr int WIDE_MINB = 3;      // resident CTAs per SM (<= 85 registers per thread; measured: 2 and 4 are slower)
__global__ void __launch_bounds__(WIDE_WARPS * 32, WIDE_MINB) k_wide(const WideArgs a) {
  extern __shared__ __align__(16) float sm_w[];
  // layout: Ws[K4/2][2][32][4] (column block of this CTA, k pairs interleaved) | per-warp zin[WIDE_ROWS][K4]
  const int K4 = a.K4;                       // total input width rounded up to a multiple of 4
  float* Ws = sm_w;
  float* zin_all = sm_w + (size_t)K4 * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col0 = blockIdx.y * 128;
  // ---- weight block: rows = concatenated k, zero padded; four independent loads in flight per thread
  for (int i0 = tid; i0 < K4 * 128; i0 += 4 * blockDim.x) {
    float w[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * blockDim.x;
      const int k = i >> 7, n = i & 127;
      w[u] = 0.f;
      if (i < K4 * 128 && col0 + n < a.Nout) {
        int kk = k;
        for (int p = 0; p < a.np; p++) {
          if (kk < a.pc[p].d) { w[u] = __ldg(a.pc[p].W + (size_t)kk * a.ldw + col0 + n); break; }
          kk -= a.pc[p].d;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * blockDim.x;
      const int k = i >> 7, n = i & 127;
      // per k pair two planes of [lane][4]: plane c/2 holds (col c, k) (col c, k+1) (col c+1, k) (col c+1, k+1) of the lane's
      // column quad - every 16 B weight load of the loop below is contiguous across the lanes (conflict-free)
      if (i < K4 * 128) Ws[((((k >> 1) * 2 + ((n & 3) >> 1)) * 32 + (n >> 2)) << 2) + ((n & 1) << 1) + (k & 1)] = w[u];
    }
  }
  __syncthreads();
  float* zin = zin_all + (size_t)warp * WIDE_ROWS * K4;
  const int n = col0 + 4 * lane;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.bias) {
    if (n + 0 < a.Nout) bias.x = a.bias[n + 0];
    if (n + 1 < a.Nout) bias.y = a.bias[n + 1];
    if (n + 2 < a.Nout) bias.z = a.bias[n + 2];
    if (n + 3 < a.Nout) bias.w = a.bias[n + 3];
  }
  const int64_t nblocks = (a.R + WIDE_ROWS - 1) / WIDE_ROWS;
  for (int64_t blk = (int64_t)blockIdx.x * WIDE_WARPS + warp; blk < nblocks; blk += (int64_t)gridDim.x * WIDE_WARPS) {
    const int64_t row0 = blk * WIDE_ROWS;
    const int rows = (int)((a.R - row0) < WIDE_ROWS ? (a.R - row0) : WIDE_ROWS);
    // ---- assemble the concat rows of this block in shared memory: lane r gathers the pieces of row r.  All index loads are
    // issued together, then the first 8 values of EVERY piece, before anything is stored: two dependent memory round trips per
    // block (a load -> store loop over k costs one round trip per k: the loop bounds are run-time values, nothing is hoisted).
    // Lanes past the end of the batch gather a valid row (the last one); their outputs are never stored.
    __syncwarp();
    {
      float* zr = zin + lane * K4;
      int64_t rr = row0 + lane;
      rr = rr < a.R ? rr : a.R - 1;
      const float* sp[WIDE_MAXP];
#pragma unroll
      for (int p = 0; p < WIDE_MAXP; p++) {
        sp[p] = nullptr;
        if (p < a.np) {
          const int64_t g = a.pc[p].idx ? (int64_t)__ldg(a.pc[p].idx + rr) : rr;
          sp[p] = a.pc[p].x + (size_t)g * a.pc[p].ldx;
        }
      }
      float t[WIDE_MAXP][8];
#pragma unroll
      for (int p = 0; p < WIDE_MAXP; p++)
#pragma unroll
        for (int j = 0; j < 8; j++) t[p][j] = (p < a.np && j < a.pc[p].d) ? __ldg(sp[p] + j) : 0.f;
      int koff = 0;
#pragma unroll
      for (int p = 0; p < WIDE_MAXP; p++) {
        if (p < a.np) {
          const int d = a.pc[p].d;
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (j < d) zr[koff + j] = t[p][j];
          for (int k0 = 8; k0 < d; k0 += 8) {      // pieces wider than 8: further batches of 8
            float u[8];
#pragma unroll
            for (int j = 0; j < 8; j++) u[j] = k0 + j < d ? __ldg(sp[p] + k0 + j) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (k0 + j < d) zr[koff + k0 + j] = u[j];
          }
          koff += d;
        }
      }
      for (int k = koff; k < K4; k++) zr[k] = 0.f;
    }
    __syncwarp();
    // ---- 8 rows x (4 columns per lane) per step
#pragma unroll 1
    for (int r0 = 0; r0 < rows; r0 += 8) {
      // plain FFMA: a packed FFMA2 occupies the FMA pipe for two cycles, so it only saves issue slots, and its k-pair
      // accumulators cost 32 more registers (measured: 178 us per launch with FFMA2 at two CTAs per SM, 172 us this way)
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[j][c] = 0.f;
      for (int k = 0; k < K4; k += 4) {
        const float* wp = Ws + (k >> 1) * 256 + 4 * lane;
        const float4 wa = *reinterpret_cast<const float4*>(wp);            // (c0: k, k+1) (c1: k, k+1)
        const float4 wb = *reinterpret_cast<const float4*>(wp + 128);      // (c2: k, k+1) (c3: k, k+1)
        const float4 wc = *reinterpret_cast<const float4*>(wp + 256);      // (c0: k+2, k+3) (c1: k+2, k+3)
        const float4 wd = *reinterpret_cast<const float4*>(wp + 384);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float4 z = *reinterpret_cast<const float4*>(zin + (r0 + j) * K4 + k);   // broadcast
          acc[j][0] = fmaf(z.w, wc.y, fmaf(z.z, wc.x, fmaf(z.y, wa.y, fmaf(z.x, wa.x, acc[j][0]))));
          acc[j][1] = fmaf(z.w, wc.w, fmaf(z.z, wc.z, fmaf(z.y, wa.w, fmaf(z.x, wa.z, acc[j][1]))));
          acc[j][2] = fmaf(z.w, wd.y, fmaf(z.z, wd.x, fmaf(z.y, wb.y, fmaf(z.x, wb.x, acc[j][2]))));
          acc[j][3] = fmaf(z.w, wd.w, fmaf(z.z, wd.z, fmaf(z.y, wb.w, fmaf(z.x, wb.z, acc[j][3]))));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (r0 + j < rows) {
          float* o = a.out + (size_t)(row0 + r0 + j) * a.ldo + n;
          const float4 v = make_float4(acc[j][0] + bias.x, acc[j][1] + bias.y, acc[j][2] + bias.z, acc[j][3] + bias.w);
          if (n + 3 < a.Nout) {
            *reinterpret_cast<float4*>(o) = v;
          } else {
            if (n + 0 < a.Nout) o[0] = v.x;
            if (n + 1 < a.Nout) o[1] = v.y;
            if (n + 2 < a.Nout) o[2] = v.z;
          }
        }
      }
    }
  }
}

// One thread per (group of ZS_NPT consecutive nodes, column of Z), groups packed back to back along the thread index.  The in-edges
// of consecutive nodes are one contiguous edge range (receiver-sorted), so a thread walks that range ONCE with independent
// loads and adds every value to the node that owns the edge - rows in ascending order per node (deterministic); the kernel is
// a chain of dependent memory round trips per thread (offsets -> [sender index ->] values), so fewer, longer threads with more
// loads in flight take fewer waves (one node per thread: 99 us at N = 262 144).
constexpr int ZS_NPT = 4;
__global__ void __launch_bounds__(256) k_zsum(const ZsumArgs a) {
  const int o_s = a.de, o_v = a.de + a.dn, o_u = a.de + 2 * a.dn, o_d = a.de + 2 * a.dn + a.dg;
  const int kz = o_d + 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t vq = t / kz;
  const int k = (int)(t - vq * kz);
  const int64_t v0 = vq * ZS_NPT;
  if (v0 >= a.N) return;
  int b[ZS_NPT + 1];
#pragma unroll
  for (int i = 0; i <= ZS_NPT; i++) b[i] = a.node_in_ptr[v0 + i < a.N ? v0 + i : a.N];
  float s[ZS_NPT];
#pragma unroll
  for (int i = 0; i < ZS_NPT; i++) s[i] = 0.f;
  if (k < o_v) {
    const bool gathered = k >= o_s;
    const float* base = gathered ? a.nf + (k - o_s) : a.ef + k;
    const int ld = gathered ? a.dn : a.de;
#pragma unroll 4
    for (int e = b[0]; e < b[ZS_NPT]; e++) {
      const int64_t row = gathered ? (int64_t)__ldg(a.edge_src + e) : (int64_t)e;
      const float val = __ldg(base + (size_t)row * ld);
#pragma unroll
      for (int i = 0; i < ZS_NPT; i++) s[i] += (e >= b[i] && e < b[i + 1]) ? val : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < ZS_NPT; i++) {
      if (v0 + i < a.N) {
        const float deg = (float)(b[i + 1] - b[i]);
        if (k < o_u) s[i] = deg * a.nf[(size_t)(v0 + i) * a.dn + (k - o_v)];
        else if (k < o_d) s[i] = deg * a.gf[(size_t)a.node_graph[v0 + i] * a.dg + (k - o_u)];
        else s[i] = deg;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ZS_NPT; i++)
    if (v0 + i < a.N) a.Z[(size_t)(v0 + i) * kz + k] = s[i];
}

// lane owns k = 4*lane + 128*c (+0..3); one row per warp step, 4 rows in flight.
template <int NO>
__global__ void __launch_bounds__(256) k_narrow(const NarrowArgs a) {
  __shared__ __align__(16) float Wsm[NARROW_KMAX * NO];   // [k][NO] of all direct sources, concatenated
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int ktot = 0;
  for (int s = 0; s < a.nsrc; s++) {
    const int d = a.src[s].d;
    for (int i = tid; i < d * NO; i += blockDim.x) {
      const int k = i / NO, j = i % NO;
      float w = 0.f;
      if (j < a.No) w = (a.src[s].W2 && j >= a.src[s].n1) ? a.src[s].W2[(size_t)k * a.ldw + (j - a.src[s].n1)] : a.src[s].W[(size_t)k * a.ldw + j];
      Wsm[(ktot + k) * NO + j] = w;
    }
    ktot += d;
  }
  __syncthreads();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 4; r0 < a.R; r0 += warps * 4) {
    float acc[4][NO];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < NO; j++) acc[u][j] = 0.f;
    int kbase = 0;
    for (int s = 0; s < a.nsrc; s++) {
      const NarrowSrc& S = a.src[s];
      const bool vec = ((S.ldx & 3) == 0) && ((((uintptr_t)S.x) & 15) == 0);
      for (int k = 4 * lane; k < S.d; k += 128) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int64_t r = r0 + u < a.R ? r0 + u : a.R - 1;
          if (vec && k + 3 < S.d) {
            x[u] = __ldg(reinterpret_cast<const float4*>(S.x + (size_t)r * S.ldx + k));
          } else {
            const float* p = S.x + (size_t)r * S.ldx + k;
            x[u].x = p[0];
            x[u].y = k + 1 < S.d ? p[1] : 0.f;
            x[u].z = k + 2 < S.d ? p[2] : 0.f;
            x[u].w = k + 3 < S.d ? p[3] : 0.f;
          }
        }
        const float* w = Wsm + (size_t)(kbase + k) * NO;
        const bool t1 = k + 1 < S.d, t2 = k + 2 < S.d, t3 = k + 3 < S.d;
#pragma unroll
        for (int j = 0; j < NO; j++) {
          const float w0 = w[j], w1 = t1 ? w[NO + j] : 0.f, w2 = t2 ? w[2 * NO + j] : 0.f, w3 = t3 ? w[3 * NO + j] : 0.f;
#pragma unroll
          for (int u = 0; u < 4; u++)
            acc[u][j] = fmaf(x[u].w, w3, fmaf(x[u].z, w2, fmaf(x[u].y, w1, fmaf(x[u].x, w0, acc[u][j]))));
        }
      }
      kbase += S.d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int j = 0; j < NO; j++) acc[u][j] += __shfl_xor_sync(0xffffffffu, acc[u][j], o);
    // lanes 0 .. 4*No-1 finish one (row, output) each: bias + gathered addends
    const int u = lane / NO, j = lane % NO;
    if (u < 4 && j < a.No && r0 + u < a.R) {
      float v = 0.f;
#pragma unroll
      for (int uu = 0; uu < 4; uu++)
#pragma unroll
        for (int jj = 0; jj < NO; jj++)
          if (uu == u && jj == j) v = acc[uu][jj];
      const int64_t r = r0 + u;
      if (a.bias) v += a.bias[j];
      for (int t = 0; t < a.nadd; t++) {
        const int64_t ar = a.add[t].idx ? (int64_t)a.add[t].idx[r] : r;
        v += a.add[t].a[(size_t)ar * a.add[t].lda + j];
      }
      a.out[(size_t)r * a.ldo + j] = v;
    }
  }
}


// Thread-per-row variant for one WIDE direct source (d % 4 == 0) plus optional small sources (d <= 16):
// a warp stages 32 rows x 128 columns coalesced into a swizzled shared-memory tile (16 B chunk c of row r at
// chunk c ^ (r & 7): conflict-free for the row-wise writes and for the column-wise reads), then lane r owns row r
// and accumulates its NO outputs with the weights broadcast from shared memory - no cross-lane reduction.
constexpr int N2_WARPS = 4;
template <int NO>
__global__ void __launch_bounds__(N2_WARPS * 32) k_narrow2(const NarrowArgs a) {
  extern __shared__ __align__(16) float sm_n2[];
  float* Wsm = sm_n2;                                       // [K][NO] of all direct sources, concatenated
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int ktot = 0;
  for (int s = 0; s < a.nsrc; s++) {
    const int d = a.src[s].d;
    for (int i = tid; i < d * NO; i += blockDim.x) {
      const int k = i / NO, j = i % NO;
      float w = 0.f;
      if (j < a.No) w = (a.src[s].W2 && j >= a.src[s].n1) ? a.src[s].W2[(size_t)k * a.ldw + (j - a.src[s].n1)] : a.src[s].W[(size_t)k * a.ldw + j];
      Wsm[(ktot + k) * NO + j] = w;
    }
    ktot += d;
  }
  __syncthreads();
  const int kw_pad = (ktot * NO + 3) / 4 * 4;
  float* tile = sm_n2 + kw_pad + (size_t)warp * 32 * 128;   // [32 rows][128 floats], swizzled
  const NarrowSrc& S0 = a.src[0];
  const int64_t nblocks = (a.R + 31) / 32;
  for (int64_t blk = (int64_t)blockIdx.x * N2_WARPS + warp; blk < nblocks; blk += (int64_t)gridDim.x * N2_WARPS) {
    const int64_t row0 = blk * 32;
    float acc[NO];
#pragma unroll
    for (int j = 0; j < NO; j++) acc[j] = 0.f;
    for (int c0 = 0; c0 < S0.d; c0 += 128) {
      const int cw = S0.d - c0 < 128 ? S0.d - c0 : 128;     // columns of this chunk (multiple of 4)
      __syncwarp();
#pragma unroll
      for (int r8 = 0; r8 < 32; r8 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          int64_t r = row0 + r8 + u;
          r = r < a.R ? r : a.R - 1;
          v[u] = (4 * lane < cw) ? __ldg(reinterpret_cast<const float4*>(S0.x + (size_t)r * S0.ldx + c0) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int r = r8 + u;
          *reinterpret_cast<float4*>(tile + r * 128 + ((lane ^ (r & 7)) << 2)) = v[u];
        }
      }
      __syncwarp();
      const float* myrow = tile + lane * 128;
      const float* w = Wsm + (size_t)c0 * NO;
      for (int c = 0; c < (cw >> 2); c++) {
        const float4 x = *reinterpret_cast<const float4*>(myrow + ((c ^ (lane & 7)) << 2));
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          if (NO == 4) {
            const float4 ww = *reinterpret_cast<const float4*>(w + (4 * c + kk) * 4);
            acc[0] = fmaf(xs[kk], ww.x, acc[0]); acc[1] = fmaf(xs[kk], ww.y, acc[1]);
            acc[2] = fmaf(xs[kk], ww.z, acc[2]); acc[3] = fmaf(xs[kk], ww.w, acc[3]);
          } else {
#pragma unroll
            for (int j = 0; j < NO; j++) acc[j] = fmaf(xs[kk], w[(4 * c + kk) * NO + j], acc[j]);
          }
        }
      }
    }
    const int64_t r = row0 + lane;
    if (r < a.R) {
      int kbase = S0.d;
      for (int s = 1; s < a.nsrc; s++) {
        const NarrowSrc& S = a.src[s];
        for (int k = 0; k < S.d; k++) {
          const float x = S.x[(size_t)r * S.ldx + k];
#pragma unroll
          for (int j = 0; j < NO; j++) acc[j] = fmaf(x, Wsm[(kbase + k) * NO + j], acc[j]);
        }
        kbase += S.d;
      }
#pragma unroll
      for (int j = 0; j < NO; j++) {
        if (j < a.No) {
          float v = acc[j];
          if (a.bias) v += a.bias[j];
          for (int t = 0; t < a.nadd; t++) {
            const int64_t ar = a.add[t].idx ? (int64_t)a.add[t].idx[r] : r;
            v += a.add[t].a[(size_t)ar * a.add[t].lda + j];
          }
          a.out[(size_t)r * a.ldo + j] = v;
        }
      }
    }
  }
}

}  // namespace

int launch_wide(gnb_ctx* ctx, const WideArgs& a0) {
  if (a0.R <= 0 || a0.Nout <= 0) return GNB_OK;
  WideArgs a = a0;
  int K = 0;
  for (int p = 0; p < a.np; p++) K += a.pc[p].d;
  GNB_CHECK(K > 0 && K <= WK_MAX, "launch_wide: concatenated input width %d not in 1..%d", K, WK_MAX);
  GNB_CHECK((a.ldo & 3) == 0 && ((uintptr_t)a.out & 15) == 0, "launch_wide: output must be 16-byte aligned rows");
  a.K4 = (K + 3) / 4 * 4;
  const size_t smem = ((size_t)a.K4 * 128 + (size_t)WIDE_WARPS * WIDE_ROWS * a.K4) * sizeof(float);
  if (ctx_first(ctx, ONCE_WIDE)) {
    GNB_CUDA(cudaFuncSetAttribute(k_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (WK_MAX * 128 + WIDE_WARPS * WIDE_ROWS * WK_MAX) * 4));
  }
  const int64_t nblocks = (a.R + WIDE_ROWS - 1) / WIDE_ROWS;
  int64_t gx = (nblocks + WIDE_WARPS - 1) / WIDE_WARPS;
  const int64_t cap = (int64_t)ctx->sm_count * WIDE_MINB;     // persistent over row blocks, one wave of resident CTAs: the weight block is loaded once per CTA
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)ceil_div(a.Nout, 128));
  double bytes = 4.0 * a.R * (K + a.Nout) + 4.0 * K * a.Nout;
  for (int p = 0; p < a.np; p++) if (a.pc[p].idx) bytes += 4.0 * a.R;
  Launch L(ctx, "wide_fp32", bytes, 2.0 * a.R * K * a.Nout);
  k_wide<<<grid, WIDE_WARPS * 32, smem, ctx->stream>>>(a);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

int launch_zsum(gnb_ctx* ctx, const ZsumArgs& a) {
  if (a.N <= 0) return GNB_OK;
  GNB_CHECK(a.de + 2 * a.dn + a.dg + 1 <= 32, "launch_zsum: aggregated input width > 32");
  Launch L(ctx, "zsum_fp32", 0, 0);
  k_zsum<<<ceil_div(ceil_div(a.N, ZS_NPT) * (a.de + 2 * a.dn + a.dg + 1), 256), 256, 0, ctx->stream>>>(a);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

int launch_narrow(gnb_ctx* ctx, const NarrowArgs& a_in) {
  NarrowArgs a = a_in;
  if (a.R <= 0 || a.No <= 0) return GNB_OK;
  int K = 0;
  for (int s = 0; s < a.nsrc; s++) K += a.src[s].d;
  GNB_CHECK(a.No <= 8 && K <= NARROW_KMAX, "launch_narrow: No %d > 8 or K %d > %d", a.No, K, NARROW_KMAX);
  double bytes = 4.0 * a.R * (K + a.No * (1 + a.nadd));
  Launch L(ctx, "narrow_fp32", bytes, 2.0 * a.R * K * a.No);
  // wide first source: thread-per-row through a swizzled staging tile; otherwise the lane-per-k kernel.  Every source carries
  // its own weight rows, so the order of the sources is free: the widest goes first (the node decoder lists its 3-wide edge
  // aggregate before the 128-wide node rows)
  for (int s = 1; s < a.nsrc; s++)
    if (a.src[s].d > a.src[0].d) { const NarrowSrc t = a.src[0]; a.src[0] = a.src[s]; a.src[s] = t; }
  bool small_rest = true;
  for (int s = 1; s < a.nsrc; s++) small_rest = small_rest && a.src[s].d <= 16;
  const NarrowSrc& S0 = a.src[0];
  if (a.nsrc >= 1 && S0.d >= 64 && (S0.d & 3) == 0 && (S0.ldx & 3) == 0 && ((uintptr_t)S0.x & 15) == 0 && small_rest) {
    const int NOp = a.No <= 4 ? 4 : 8;
    const size_t smem = ((size_t)(K * NOp + 3) / 4 * 4 + (size_t)N2_WARPS * 32 * 128) * sizeof(float);
    if (ctx_first(ctx, ONCE_NARROW2)) {
      GNB_CUDA(cudaFuncSetAttribute(k_narrow2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (NARROW_KMAX * 4 + N2_WARPS * 32 * 128) * 4));
      GNB_CUDA(cudaFuncSetAttribute(k_narrow2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (NARROW_KMAX * 8 + N2_WARPS * 32 * 128) * 4));
    }
    int64_t gx = ((a.R + 31) / 32 + N2_WARPS - 1) / N2_WARPS;
    const int64_t cap = (int64_t)ctx->sm_count * 3;
    if (gx > cap) gx = cap;
    if (NOp == 4) k_narrow2<4><<<(unsigned)gx, N2_WARPS * 32, smem, ctx->stream>>>(a);
    else k_narrow2<8><<<(unsigned)gx, N2_WARPS * 32, smem, ctx->stream>>>(a);
  } else {
    int64_t gx = (a.R + 31) / 32;            // 8 warps x 4 rows per CTA step
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (gx > cap) gx = cap;
    if (a.No <= 4) k_narrow<4><<<(unsigned)gx, 256, 0, ctx->stream>>>(a);
    else k_narrow<8><<<(unsigned)gx, 256, 0, ctx->stream>>>(a);
  }
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
