// Generic bf16 tcgen05 linear layer for WIDE GNCore layers (hidden width 256, 384, ...: every width that is a multiple of 64
// on the K side and of 128 on the N side), same contract as the fp32 kernel it replaces (fp32.cu::k_linear):
//
//   out = act( sum_s LN_s(x_s) W_s + bias + sum_j add_j[idx_j] )
//
// i.e. every Dense of src/gnblock.jl:65-67 and src/gnfeedforward.jl:27-31 after the linearity split of the concat
// (src/edgefninput.jl:1-8): direct row sources are concatenated along K, gathered node / graph projections are added in the
// epilogue.  The 128-wide cores have their own fully fused kernel (tc_edge.cu); this one trades fusion for generality: one
// launch per Dense, activations make an HBM round trip between launches.
//
// Work item = (128-row tile, group of NG <= 2 output blocks of 128 columns); a persistent CTA walks the row tiles and, inside
// a tile, the column groups (the LayerNorm statistics of the tile are computed once and kept in shared memory).  18 warps:
//   warps 4-11 A producers : fp32 rows -> LayerNorm (two-pass statistics, affine applied here) -> bf16 -> 128B-swizzled K-major
//                            64-wide slabs (16 KB) through a 6-stage ring; K <= 256: the slabs of a row tile are produced
//                            once and reused by all column groups
//   warp 17    W loader    : cp.async.bulk of pre-packed bf16 slabs (NG x 16 KB per K step) through a 3-stage ring
//   warp 16    MMA issuer  : per K step 4 UMMAs (K = 16) per output block into TMEM (2 accumulator sets x 256 columns)
//   warps 0-3, 12-15 drain : accumulator fragments -> + bias + gathered addends -> relu -> sector-exact fp32 stores
#include "tc_ptx.cuh"
#include "tc_gemm.cuh"
#include <map>
#include <mutex>
#include <string>
#include <stdlib.h>
#include <tuple>

using namespace tcx;

namespace {

#define mbar_wait(b, p) mbar_wait_w((b), (p), wd_dead, a.wd)      /* `wd`: the kernel's Watch (tc_ptx.cuh) */

#ifndef GNB_LIN_L2_HINTS
#define GNB_LIN_L2_HINTS 0      /* 1: k_tc_lin with gathered addend tables evict_last, streamed rows evict_first after their last read, fp32 output stored streaming.  Measured neutral at hidden 256 (523 vs 523 us per launch) and 5 % slower at 384: off */
#endif
#ifndef GNB_LIN_PREFETCH
#define GNB_LIN_PREFETCH 0      /* 1: k_tc_lin producers prefetch the rows of their next tile into L2.  Measured SLOWER (cfg5: 541 vs 518 us per launch, profiles/r02_summary.md): the kernel is bound by memory throughput, not by load latency */
#endif
constexpr int SLAB = KB_BYTES;            // 128 rows x 64 k bf16
constexpr int NA = 6, NW = 3;             // ring depths (A: 6 x 16 KB, W: 3 x 32 KB)
constexpr int G_OFF_A = 0;
constexpr int G_OFF_W = NA * SLAB;                    // stages of 2 slabs
constexpr int G_OFF_STAT = G_OFF_W + NW * 2 * SLAB;   // float2 stats[3][128]
constexpr int G_OFF_BAR = G_OFF_STAT + 3 * 128 * 8;
constexpr int G_SMEM = G_OFF_BAR + 32 * 8 + 16 + 1024;
constexpr int G_THREADS = 18 * 32;
constexpr int W_MMA = 16, W_LOAD = 17;      // warps 0-3 and 12-15 drain (TMEM lane quadrant = warp % 4), 4-11 A producers
enum { GB_WFULL = 0, GB_WEMPTY = 3, GB_AFULL = 6, GB_AEMPTY = 12, GB_ACCFULL = 18, GB_ACCFREE = 20 };

struct GemmArgs {
  int64_t R;
  int Nout, ldo;
  int nsrc;
  const float* x[3]; int ldx[3]; int d[3]; int xbf[3];      // xbf: the source is bf16 (ldx in elements)
  const float* gamma[3]; const float* beta[3]; float eps[3]; int eps_mode[3];
  int KS;                         // total K / 64
  const __nv_bfloat16* wpack;     // [n group][K step][block in group][8192 elements]
  const float* bias;
  int nadd;
  const float* add[4]; const int32_t* add_idx[4]; int lda[4];
  int relu, out_bf16;
  float* out;
  int num_tiles, ngroups, NG;     // NG output blocks per group (1 or 2)
  int resident;                   // the A slabs of a row tile are produced once and reused by every column group (KS <= NA - 2)
  WatchArgs wd;                   // kernel watchdog (tc_ptx.cuh)
};

// LayerNorm statistics of the 16 rows of one producer warp, two-pass in fp32 (src/gngraphnorm.jl:19-26): 16 lanes per row, the
// row stays in registers for both passes; NP row pairs (= 2 NP rows) of NV float4 per lane are loaded before anything is reduced.
template <int NP, int NV>
__device__ __forceinline__ void tile_stats(const float* xs, int ldx, int nv, float inv_d, int64_t row0, int64_t R, int hr, int c16,
                                           float2* st, float eps, int eps_mode, uint64_t pol) {
#pragma unroll 1
  for (int r0 = 0; r0 < 16; r0 += 2 * NP) {
    float4 v[NP][NV];
#pragma unroll
    for (int p = 0; p < NP; p++) {
      int64_t row = row0 + r0 + 2 * p + hr;
      row = row < R ? row : R - 1;
      const float* xr = xs + (size_t)row * ldx;
#pragma unroll
      for (int j = 0; j < NV; j++) v[p][j] = j < nv ? (GNB_LIN_L2_HINTS ? ld_hint(reinterpret_cast<const float4*>(xr + 64 * j), pol) : __ldg(reinterpret_cast<const float4*>(xr + 64 * j))) : f4zero();
    }
    float sum[NP], sq[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < NV; j++) t += (v[p][j].x + v[p][j].y) + (v[p][j].z + v[p][j].w);
      sum[p] = t;
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1)
#pragma unroll
      for (int p = 0; p < NP; p++) sum[p] += __shfl_xor_sync(0xffffffffu, sum[p], o);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const float mu = sum[p] * inv_d;
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < NV; j++) {
        if (j < nv) {
          const float dx = v[p][j].x - mu, dy = v[p][j].y - mu, dz = v[p][j].z - mu, dw = v[p][j].w - mu;
          t += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
      }
      sq[p] = t;
      sum[p] = mu;
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1)
#pragma unroll
      for (int p = 0; p < NP; p++) sq[p] += __shfl_xor_sync(0xffffffffu, sq[p], o);
    if (c16 == 0) {
#pragma unroll
      for (int p = 0; p < NP; p++) st[r0 + 2 * p + hr] = make_float2(sum[p], ln_rstd(sq[p] * inv_d, eps, eps_mode));
    }
  }
}

__global__ void __launch_bounds__(G_THREADS, 1) k_tc_lin(const GemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float2* stats = reinterpret_cast<float2*>(sm + G_OFF_STAT);      // (mean, rstd) per source and row of the tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + G_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  bool wd_dead = false;      // kernel watchdog (tc_ptx.cuh)
  if (tid == 0) {
    for (int i = 0; i < NW; i++) { mbar_init(BAR(GB_WFULL + i), 1); mbar_init(BAR(GB_WEMPTY + i), 1); }
    for (int i = 0; i < NA; i++) { mbar_init(BAR(GB_AFULL + i), 8); mbar_init(BAR(GB_AEMPTY + i), 1); }
    for (int i = 0; i < 2; i++) { mbar_init(BAR(GB_ACCFULL + i), 1); mbar_init(BAR(GB_ACCFREE + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int KS = a.KS, NG = a.NG;
  // every CTA walks the K slabs starting at a different offset: the CTAs run in lock step, and without the rotation all of
  // them stream the same weight slab from the same L2 slice at the same time
  const int rot = (int)((blockIdx.x * 5u) % (uint32_t)KS);

  if (warp == W_LOAD) {
    // ===================================================== weight loader
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x) {
      for (int ng = 0; ng < a.ngroups; ng++) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)ng * KS * NG * SLAB;
#pragma unroll 1
        for (int ks = 0; ks < KS; ks++, it++) {
          const uint32_t st = it % NW, ph = (it / NW) & 1;
          mbar_wait(BAR(GB_WEMPTY + st), ph ^ 1);
          if (elect_one()) {
            const uint32_t nb = (uint32_t)(NG * SLAB);
            mbar_expect_tx(BAR(GB_WFULL + st), nb);
            const int kr = ks + rot < KS ? ks + rot : ks + rot - KS;
            bulk_g2s(base + G_OFF_W + st * 2 * SLAB, src + (size_t)kr * NG * SLAB, nb, BAR(GB_WFULL + st));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================================================== MMA issuer
    uint32_t ita = 0, itw = 0, item = 0, tl = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x, tl++) {
      for (int ng = 0; ng < a.ngroups; ng++, item++) {
        const uint32_t buf = item & 1, aph = (item >> 1) & 1;
        mbar_wait(BAR(GB_ACCFREE + buf), aph ^ 1);
        tc_fence_after();
        if (a.resident) ita = tl * KS;      // every column group walks the same slabs of the tile
        const bool first = !a.resident || ng == 0, last = !a.resident || ng == a.ngroups - 1;
#pragma unroll 1
        for (int ks = 0; ks < KS; ks++, ita++, itw++) {
          const uint32_t as = ita % NA, pa = (ita / NA) & 1, ws = itw % NW, pw = (itw / NW) & 1;
          if (first) mbar_wait(BAR(GB_AFULL + as), pa);
          mbar_wait(BAR(GB_WFULL + ws), pw);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = umma_desc(base + G_OFF_A + as * SLAB);
            for (int nb = 0; nb < NG; nb++) {
              const uint64_t wd = umma_desc(base + G_OFF_W + ws * 2 * SLAB + nb * SLAB);
              const uint32_t D = tmem + buf * 256 + nb * 128;
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) mma_ss(D, ad + 2 * k4, wd + 2 * k4, IDESC, (ks > 0 || k4 > 0) ? 1u : 0u);
            }
            if (last) tc_commit(BAR(GB_AEMPTY + as));
            tc_commit(BAR(GB_WEMPTY + ws));
            if (ks == KS - 1) tc_commit(BAR(GB_ACCFULL + buf));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================== A producers: warp pw owns rows 16 pw .. 16 pw + 15 of the tile
    const int pw = warp - 4;
    const int hr = lane >> 4, c16 = lane & 15;      // two rows per load instruction, 16 lanes x 16 B per row slab
    const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();      // statistics pass, then the last read
    uint32_t it = 0;
    // Experiment (GNB_LIN_PREFETCH=1, off: measured slower): the rows of this warp's NEXT tile -> L2 (one bulk prefetch per
    // contiguous source), so that the statistics pass below loads at L2 instead of HBM latency
    auto prefetch_rows = [&](int t) {
      const int64_t r0 = (int64_t)t * TM + 16 * pw;
      if (t < a.num_tiles && r0 < a.R && elect_one()) {
        const int64_t nr = a.R - r0 < 16 ? a.R - r0 : 16;
        for (int s = 0; s < a.nsrc; s++) {
          const size_t esz = a.xbf[s] ? 2 : 4;
          if (a.ldx[s] != a.d[s] || ((a.d[s] * esz) & 15)) continue;      // rows not contiguous (a column slice of a wider buffer)
          bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(a.x[s]) + (size_t)r0 * a.ldx[s] * esz, (uint32_t)(nr * a.d[s] * esz));
        }
      }
      __syncwarp();
    };
    if (GNB_LIN_PREFETCH) prefetch_rows(blockIdx.x);
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x) {
      const int64_t row0 = (int64_t)tile * TM + 16 * pw;
      if (GNB_LIN_PREFETCH) prefetch_rows(tile + gridDim.x);
      // ---- LayerNorm statistics of this warp's rows, two-pass in fp32 (src/gngraphnorm.jl:19-26), kept for all column groups
      // (tile_stats above)
      for (int s = 0; s < a.nsrc; s++) {
        if (a.gamma[s] == nullptr) continue;
        const int d = a.d[s], nv = d >> 6;      // float4 per lane and row (d <= 512: nv <= 8)
        float2* st = stats + s * 128 + 16 * pw;
        // rows up to 256 wide: 8 rows of loads in flight per warp, wider: 4 (the same 64 registers)
        if (nv <= 4) tile_stats<4, 4>(a.x[s] + 4 * c16, a.ldx[s], nv, 1.0f / (float)d, row0, a.R, hr, c16, st, a.eps[s], a.eps_mode[s], pol_last);
        else tile_stats<2, 8>(a.x[s] + 4 * c16, a.ldx[s], nv, 1.0f / (float)d, row0, a.R, hr, c16, st, a.eps[s], a.eps_mode[s], pol_last);
      }
      __syncwarp();
      const int npass = a.resident ? 1 : a.ngroups;
      for (int ng = 0; ng < npass; ng++) {
        float4 vn[8];      // the loads of the next slab fly while the current one is converted
        // K slab ks of this CTA's (rotated) order -> (source, offset inside the source)
        auto locate = [&](int ks, int& ss, int& kk) {
          const int kr = ks + rot < KS ? ks + rot : ks + rot - KS;
          ss = 0; kk = kr * 64;
          while (kk >= a.d[ss]) { kk -= a.d[ss]; ss++; }
        };
        auto issue = [&](int ss, int kk) {
          if (a.xbf[ss]) {
            const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(a.x[ss]) + kk + 8 * (lane & 7);
#pragma unroll
            for (int u = 0; u < 4; u++) {
              int64_t row = row0 + 4 * u + (lane >> 3);
              row = row < a.R ? row : a.R - 1;
              vn[u] = GNB_LIN_L2_HINTS ? ld_hint(reinterpret_cast<const float4*>(xs + (size_t)row * a.ldx[ss]), pol_first) : __ldg(reinterpret_cast<const float4*>(xs + (size_t)row * a.ldx[ss]));
            }
          } else {
            const float* xs = a.x[ss] + kk + 4 * c16;
#pragma unroll
            for (int u = 0; u < 8; u++) {
              int64_t row = row0 + 2 * u + hr;
              row = row < a.R ? row : a.R - 1;
              vn[u] = GNB_LIN_L2_HINTS ? ld_hint(reinterpret_cast<const float4*>(xs + (size_t)row * a.ldx[ss]), pol_first) : __ldg(reinterpret_cast<const float4*>(xs + (size_t)row * a.ldx[ss]));
            }
          }
        };
        int s, koff;
        locate(0, s, koff);
        issue(s, koff);
#pragma unroll 1
        for (int ks = 0; ks < KS; ks++, it++) {
          locate(ks, s, koff);
          const uint32_t st = it % NA, ph = (it / NA) & 1;
          const bool ln = a.gamma[s] != nullptr, bf = a.xbf[s] != 0;
          float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ln) {
            g4 = __ldg(reinterpret_cast<const float4*>(a.gamma[s] + koff + 4 * c16));
            b4 = __ldg(reinterpret_cast<const float4*>(a.beta[s] + koff + 4 * c16));
          }
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = vn[u];
          if (ks + 1 < KS) {
            int s2, k2;
            locate(ks + 1, s2, k2);
            issue(s2, k2);
          }
          uint8_t* A = sm + G_OFF_A + st * SLAB;
          mbar_wait(BAR(GB_AEMPTY + st), ph ^ 1);
          if (bf) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int r = 16 * pw + 4 * u + (lane >> 3);
              float4 t = v[u];
              if (row0 + 4 * u + (lane >> 3) >= a.R) t = f4zero();
              *reinterpret_cast<float4*>(A + r * 128 + (((lane & 7) ^ (r & 7)) << 4)) = t;
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const int r = 16 * pw + 2 * u + hr;
              float4 t = v[u];
              if (ln) {
                const float2 ms = stats[s * 128 + r];
                t.x = (t.x - ms.x) * ms.y * g4.x + b4.x;
                t.y = (t.y - ms.x) * ms.y * g4.y + b4.y;
                t.z = (t.z - ms.x) * ms.y * g4.z + b4.z;
                t.w = (t.w - ms.x) * ms.y * g4.w + b4.w;
              }
              if (row0 + 2 * u + hr >= a.R) t = f4zero();
              uint2 pk;
              pk.x = pack_bf16(t.x, t.y);
              pk.y = pack_bf16(t.z, t.w);
              *reinterpret_cast<uint2*>(A + sw_off(r, 4 * c16)) = pk;
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(GB_AFULL + st));
        }
      }
    }
  } else if (warp < 4 || (warp >= 12 && warp < 16)) {
    // ===================================================== drain (TMEM lane quadrant = warp % 4; the two warps of a quadrant
    // split the 64-column chunks of the column group), accumulator-fragment layout: every 4 lanes own one 32 B sector of a row
    const int dq = warp & 3, dh = warp >= 12 ? 1 : 0;
    const uint64_t pol_add = l2_policy_evict_last();      // the gathered node / graph tables are re-read by every edge of the node / graph
    const uint32_t lane_base = ((uint32_t)(dq * 32)) << 16;
    const int qr = lane >> 2, cq = 2 * (lane & 3);
    uint32_t item = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x) {
      const int64_t row0 = (int64_t)tile * TM + 32 * dq;
      // the 4 rows of this lane: 16 hh + 8 h2 + qr
      int32_t arow[4][4];      // gathered addend rows (row counts fit 31 bits: checked by the lowering)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        int64_t r = row0 + 16 * (k >> 1) + 8 * (k & 1) + qr;
        r = r < a.R ? r : a.R - 1;
#pragma unroll
        for (int j = 0; j < 4; j++) arow[j][k] = (j < a.nadd && a.add_idx[j]) ? __ldg(a.add_idx[j] + r) : (int32_t)r;
      }
      for (int ng = 0; ng < a.ngroups; ng++, item++) {
        const uint32_t buf = item & 1, ph = (item >> 1) & 1;
        mbar_wait(BAR(GB_ACCFULL + buf), ph);
        tc_fence_after();
        const int nch = NG;      // 64-column chunks of this warp: [dh * NG, dh * NG + NG)
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
          const int ch = dh * NG + c;
          const int col = ng * NG * 128 + 64 * ch + cq;      // + 8 n
          float2 bias[8];
#pragma unroll
          for (int n = 0; n < 8; n++) bias[n] = a.bias ? __ldg(reinterpret_cast<const float2*>(a.bias + col + 8 * n)) : make_float2(0.f, 0.f);
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            uint32_t dreg[32];
            TC_LD_FRAG64(tmem + buf * 256 + lane_base + ((uint32_t)(16 * hh) << 16) + 64 * ch, dreg);
            tc_wait_ld();
            if (c == nch - 1 && hh == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(BAR(GB_ACCFREE + buf));
            }
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
              const int k = 2 * hh + h2;
              const int64_t row = row0 + 16 * hh + 8 * h2 + qr;
              float2 v[8];
#pragma unroll
              for (int n = 0; n < 8; n++)
                v[n] = make_float2(__uint_as_float(dreg[4 * n + 2 * h2]) + bias[n].x, __uint_as_float(dreg[4 * n + 2 * h2 + 1]) + bias[n].y);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                if (j >= a.nadd) break;
                const float* ap = a.add[j] + (size_t)(uint32_t)arow[j][k] * a.lda[j] + col;
#pragma unroll
                for (int n = 0; n < 8; n++) {
                  const float2 t = GNB_LIN_L2_HINTS ? ld_hint2(reinterpret_cast<const float2*>(ap + 8 * n), pol_add) : __ldg(reinterpret_cast<const float2*>(ap + 8 * n));
                  v[n].x += t.x; v[n].y += t.y;
                }
              }
              if (a.relu) {
#pragma unroll
                for (int n = 0; n < 8; n++) { v[n].x = fmaxf(v[n].x, 0.f); v[n].y = fmaxf(v[n].y, 0.f); }
              }
              if (row < a.R) {
                if (a.out_bf16) {
                  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + (size_t)row * a.ldo + col;
#pragma unroll
                  for (int n = 0; n < 8; n++) *reinterpret_cast<uint32_t*>(o + 8 * n) = pack_bf16(v[n].x, v[n].y);
                } else {
                  float* o = a.out + (size_t)row * a.ldo + col;
#pragma unroll
                  for (int n = 0; n < 8; n++) {
                    if (GNB_LIN_L2_HINTS) __stcs(reinterpret_cast<float2*>(o + 8 * n), v[n]);
                    else *reinterpret_cast<float2*>(o + 8 * n) = v[n];
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// packed[((ng * KS + ks) * NG + nb) * 8192 + swizzled(n, k)] = bf16(W[(ks*64 + k) * ldw + (ng*NG + nb)*128 + n]),
// W given per source as k-major row blocks (LinSrc::W)
struct PackSrc { const float* W[3]; int d[3]; int nsrc; };
__global__ void k_pack_lin(PackSrc ps, int ldw, int KS, int NG, int ngroups, __nv_bfloat16* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over ngroups * KS * NG * 8192
  const int64_t total = (int64_t)ngroups * KS * NG * 8192;
  if (idx >= total) return;
  const int e = (int)(idx & 8191);
  int64_t slab = idx >> 13;
  const int nb = (int)(slab % NG); slab /= NG;
  const int ks = (int)(slab % KS);
  const int ng = (int)(slab / KS);
  const int k = e >> 7, n = e & 127;      // element (n, k) of the slab, k in [0, 64)
  int kk = ks * 64 + k, s = 0;
  while (kk >= ps.d[s]) { kk -= ps.d[s]; s++; }
  const float w = ps.W[s][(size_t)kk * ldw + (size_t)(ng * NG + nb) * 128 + n];
  dst[(idx >> 13) * 8192 + (sw_off(n, k) >> 1)] = __float2bfloat16_rn(w);
}

struct PackKey {
  uint64_t model;
  const float* W[3];
  int d[3];
  int Nout, ldw, ng;
  bool operator<(const PackKey& o) const {
    return std::tie(model, W[0], W[1], W[2], d[0], d[1], d[2], Nout, ldw, ng) <
           std::tie(o.model, o.W[0], o.W[1], o.W[2], o.d[0], o.d[1], o.d[2], o.Nout, o.ldw, o.ng);
  }
};
struct PackCache { std::mutex mu; std::map<PackKey, __nv_bfloat16*> m; };      // mu: gnb_model_destroy evicts from any thread

// =====================================================================================================
// Fused GNFeedForward + residuals for hidden width FH = 256 or 384 (src/gnfeedforward.jl:27-31, src/gncore.jl:56-59):
//   y = x + h + W2 relu(W1 LN2(x) + b1) + b2
// One 128-row tile per pass, persistent CTA, the 4H = 1024 wide hidden activation never leaves the SM (the generic path above
// writes / reads it through HBM: 4 KB per row).  Per tile 8 hidden chunks of 128 units, software pipelined:
//   up(c)  : Hd[c & 1] (TMEM, 128 fp32 columns) = A . W1[:, chunk c]          4 K slabs x 4 UMMAs (SS)
//   conv(c): Hd -> + b1 -> relu -> bf16, in place (DRAIN warps)               the A operand of the down projection (TS)
//   down(c): D (TMEM, 256 fp32 columns) += relu(.)[chunk c] . W2[chunk c, :]   2 output blocks x 8 UMMAs (TS)
// issued as up0 up1 down0 up2 down1 ... so that conv(c) runs under up(c+1).  TMEM: D 256 + Hd 2 x 128 = 512 columns.
// FH = 384: D takes 384 columns, so there is ONE hidden buffer (up(c) conv(c) down(c) in sequence) and one A stage (96 KB).
// 14 warps: 0-3 DRAIN (conversion), 4-7 A producers (LayerNorm -> bf16 slabs, next tile), 8-11 EPI (D + b2 + x + h -> y,
// fragment layout), 12 MMA issuer (two slabs = 8 UMMAs per iteration), 13 weight loader (16 KB slabs, 5-stage ring: W1 slabs of chunk c, then W2 slabs of chunk c-1).
// =====================================================================================================
// CL2 (default): the two CTAs of a cluster (an SM pair) run ONE 256-row cta_group::2 UMMA stream; each CTA holds its own 128-row A
// tile and the N half (64 of 128 rows, 8 KB) of every weight slab - half the weight bytes per SM (the kernel streams 1 MB of
// slabs per 128-row tile, which bound it at ~13 TB/s of L2 -> SM traffic: profiles/r01_wide_ncu.json) and a ring twice as deep
// in slabs.  Only the leader CTA issues UMMAs; its barriers collect the arrivals of both CTAs, tcgen05.commit multicasts the
// completions to both (same protocol as k_edge5, tc_edge.cu).
constexpr int F_RING_BYTES = 5 * SLAB;                    // weight ring: 5 x 16 KB slabs, or 10 x 8 KB half slabs (CL2); consumed two slabs at a time
constexpr int F_THREADS = 14 * 32;
#ifndef GNB_FFN_L2_HINTS
#define GNB_FFN_L2_HINTS 1      /* L2 eviction priorities for the two reads of x (ncu: the second one missed L2: 2.2 GB of DRAM reads per edge launch) */
#endif
#ifndef GNB_FFN_Y_STREAM
#define GNB_FFN_Y_STREAM 0      /* 1: final y rows stored streaming (evict-first); measured neutral (1047 vs 1051 us per launch) */
#endif
#define FFN_LDX(p) (GNB_FFN_L2_HINTS ? ld_hint((p), pol_last) : __ldg(p))
#ifndef GNB_FFN_EPI_PREFETCH
#define GNB_FFN_EPI_PREFETCH 0      /* 1: EPI warps prefetch the x / h rows of their next tile into L2.  Measured SLOWER (cfg5: 1175 vs 1129 us per launch) */
#endif
#ifndef GNB_FFN_EPI_ROWS
#define GNB_FFN_EPI_ROWS 2          /* rows in flight per EPI warp in phase 2 at hidden 256 (4 measured slower still: 1232 us; hidden 384 always 2) */
#endif
enum { FB_WFULL = 0, FB_WEMPTY = 10, FB_AFULL = 20, FB_AEMPTY = 22, FB_HIDFULL = 24, FB_HSREADY = 26, FB_ACCFULL = 28, FB_ACCFREE = 29 };
// Block order of a tile, step s -> 2 * chunk + (0: up projection, 1: down projection).  Two hidden buffers: up and down blocks
// go in PAIRS  up(2p) up(2p+1) down(2p) down(2p+1): conv(2p) runs under up(2p+1), conv(2p+1) under down(2p), and the tensor pipe
// switches between SS and TS operand mode 8 times per tile instead of 16 (~435 cycles each, tools/micro/hwprobe.cu T5).
// One hidden buffer (FH = 384): up(c) down(c) in sequence.
template <int NHD>
__device__ __forceinline__ int ffn_step(int s) {
  if (NHD == 2) { const int p = s >> 2, r = s & 3; return 2 * (2 * p + (r & 1)) + (r >> 1); }
  return s;      // 2 c + {0, 1}
}
template <int FH> struct FfnCfg {
  static constexpr int KS = FH / 64;               // K slabs of the up projection
  static constexpr int CH = 4 * FH / 128;          // hidden chunks
  static constexpr int NB = FH / 128;              // output blocks of the down projection
  static constexpr int NHD = FH <= 256 ? 2 : 1;    // hidden buffers in TMEM (D = FH columns, 512 in total)
  static constexpr int NAS = FH <= 256 ? 2 : 1;    // A tile stages in shared memory
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = NAS * KS * SLAB;
  static constexpr int OFF_STAT = OFF_W + F_RING_BYTES;     // float2 stats[128]
  static constexpr int OFF_B1 = OFF_STAT + 128 * 8;         // float b1[4 FH]
  static constexpr int OFF_BAR = OFF_B1 + 4 * FH * 4;
  static constexpr int SMEM = OFF_BAR + 32 * 8 + 16 + 1024;
};

struct FfnArgs {
  const float* x;       // [R][256]
  const float* h;       // [R][256] block output (second residual addend)
  float* y;             // [R][256]
  int64_t R;
  int num_tiles;
  const float *gamma, *beta; float eps; int eps_mode;      // LN2
  const __nv_bfloat16* w1;      // [8 chunks][4 K slabs][8192]
  const __nv_bfloat16* w2;      // [2 output blocks][16 K slabs][8192]
  const float *b1, *b2;
  WatchArgs wd;                 // kernel watchdog (tc_ptx.cuh)
};

template <int FH, bool CL2>
__global__ void __launch_bounds__(F_THREADS, 1) k_tc_ffn(const FfnArgs a) {
  using Cfg = FfnCfg<FH>;
  constexpr int F_NW = CL2 ? 10 : 5;                       // ring stages
  constexpr uint32_t WST = CL2 ? SLAB / 2 : SLAB;          // bytes per stage: this CTA's share of one slab
  constexpr uint32_t NARR = CL2 ? 8u : 4u;                 // arrivals of a 4-warp role on a leader barrier
  constexpr int F_H = FH, F_KS = Cfg::KS, F_CH = Cfg::CH, F_NB = Cfg::NB, NHD = Cfg::NHD, NAS = Cfg::NAS;
  constexpr int F_OFF_A = Cfg::OFF_A, F_OFF_W = Cfg::OFF_W, F_OFF_STAT = Cfg::OFF_STAT, F_OFF_B1 = Cfg::OFF_B1, F_OFF_BAR = Cfg::OFF_BAR;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float2* stats = reinterpret_cast<float2*>(sm + F_OFF_STAT);
  float* sB1 = reinterpret_cast<float*>(sm + F_OFF_B1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + F_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  bool wd_dead = false;      // kernel watchdog (tc_ptx.cuh)
  const uint32_t rank = CL2 ? cluster_ctarank() : 0u;      // == blockIdx.x & 1
  if (tid == 0) {
    // pair leader: WFULL[st] also counts the peer's "my half has landed" relay, so the MMA warp waits on ONE barrier per slab
    for (int i = 0; i < F_NW; i++) { mbar_init(BAR(FB_WFULL + i), (CL2 && rank == 0) ? 2 : 1); mbar_init(BAR(FB_WEMPTY + i), 1); }
    for (int i = 0; i < 2; i++) {
      mbar_init(BAR(FB_AFULL + i), NARR); mbar_init(BAR(FB_AEMPTY + i), 1);
      mbar_init(BAR(FB_HIDFULL + i), 1); mbar_init(BAR(FB_HSREADY + i), NARR);
    }
    mbar_init(BAR(FB_ACCFULL), 1); mbar_init(BAR(FB_ACCFREE), NARR);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    if (CL2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = tid; i < 4 * F_H; i += F_THREADS) sB1[i] = a.b1[i];
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();      // the peer's barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  // a 4-warp role's arrival on a barrier the MMA warp (leader CTA) waits on
  auto ARRIVE_LEADER = [&](int i) {
    __syncwarp();
    if (lane == 0) {
      if (CL2) mbar_arrive_cluster(map_to_cta(BAR(i), 0));
      else mbar_arrive(BAR(i));
    }
  };
#define TILE_OK(t) ((t) - (int)rank < a.num_tiles && !wd_dead)   /* both CTAs of a pair run the same number of passes */
  const uint32_t tmem = *tmem_slot;
  const uint32_t Dt = tmem, Hd0 = tmem + F_H;      // D: columns [0, FH); Hd[b]: [FH + 128 b, +128)
  // the hidden chunks are independent terms of the down projection: every CTA starts at a different one, so that the
  // lock-stepped CTAs do not stream the same weight slab from the same L2 slice at the same time
  const int crot = (int)((CL2 ? blockIdx.x >> 1 : blockIdx.x) % F_CH);      // (the same for both CTAs of a pair)
  auto rotc = [&](int c) { const int t = c + crot; return t < F_CH ? t : t - F_CH; };

  if (warp == 13) {
    // ===================================================== weight loader: slabs in the order the MMA warp consumes them
    uint32_t it = 0;
    auto load = [&](const __nv_bfloat16* src) {
      const uint32_t st = it % F_NW, ph = (it / F_NW) & 1;
      it++;
      mbar_wait(BAR(FB_WEMPTY + st), ph ^ 1);
      if (elect_one()) {
        // CL2: this CTA's N half of the slab = rows [64 rank, 64 rank + 64), 8 KB contiguous in the swizzled slab image
        mbar_expect_tx(BAR(FB_WFULL + st), WST);
        bulk_g2s(base + F_OFF_W + st * WST, reinterpret_cast<const uint8_t*>(src) + (CL2 ? rank * WST : 0u), WST, BAR(FB_WFULL + st));
      }
      __syncwarp();
    };
    for (int tile = blockIdx.x; TILE_OK(tile); tile += gridDim.x) {
#pragma unroll 1
      for (int s2 = 0; s2 < F_CH * 2; s2++) {      // the MMA warp's block order (ffn_step below)
        const int c = ffn_step<NHD>(s2) >> 1;
        if ((ffn_step<NHD>(s2) & 1) == 0) {
          for (int ks = 0; ks < F_KS; ks++) load(a.w1 + (size_t)(rotc(c) * F_KS + ks) * 8192);
        } else {
          for (int nb = 0; nb < F_NB; nb++)
            for (int kh = 0; kh < 2; kh++) load(a.w2 + (size_t)(nb * (4 * F_H / 64) + 2 * rotc(c) + kh) * 8192);
        }
      }
    }
  } else if (warp == 12) {
    // ===================================================== MMA issuer: two weight slabs (8 UMMAs) per iteration - the issue
    // overhead of a barrier wait + fence + election per 4 UMMAs is more than the 256 cycles those UMMAs take
    uint32_t it = 0, tl = 0, nhid[2] = {0, 0};
    uint64_t wd0 = 0, wd1 = 0;
    uint32_t ws0 = 0, ws1 = 0;
    auto get_w2 = [&]() {
      ws0 = it % F_NW; ws1 = (it + 1) % F_NW;
      const uint32_t p0 = (it / F_NW) & 1, p1 = ((it + 1) / F_NW) & 1;
      it += 2;
      mbar_wait(BAR(FB_WFULL + ws0), p0);      // CL2: own half (expect_tx) + the peer's relay arrive
      mbar_wait(BAR(FB_WFULL + ws1), p1);
      wd0 = umma_desc(base + F_OFF_W + ws0 * WST);
      wd1 = umma_desc(base + F_OFF_W + ws1 * WST);
    };
    auto COMMIT = [&](int i) {
      if (CL2) tc_commit2(BAR(i));
      else tc_commit(BAR(i));
    };
    constexpr uint32_t ID = CL2 ? IDESC2 : IDESC;
    auto MMA_SS = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) {
      if (CL2) mma_ss2(d, ad, bd, ID, acc);
      else mma_ss(d, ad, bd, ID, acc);
    };
    auto MMA_TS = [&](uint32_t d, uint32_t at, uint64_t bd, uint32_t acc) {
      if (CL2) mma_ts2(d, at, bd, ID, acc);
      else mma_ts(d, at, bd, ID, acc);
    };
    if (CL2 && rank != 0) {
      // peer CTA: relay "my half of the slab has landed" to the leader, in ring order
      constexpr int SLABS_PER_TILE = F_CH * (F_KS + 2 * F_NB);
      for (int tile = blockIdx.x; TILE_OK(tile); tile += gridDim.x) {
#pragma unroll 1
        for (int s2 = 0; s2 < SLABS_PER_TILE; s2++, it++) {
          const uint32_t st = it % F_NW, ph = (it / F_NW) & 1;
          mbar_wait(BAR(FB_WFULL + st), ph);
          if (lane == 0) mbar_arrive_cluster(map_to_cta(BAR(FB_WFULL + st), 0));
          __syncwarp();
        }
      }
    } else
    for (int tile = blockIdx.x; TILE_OK(tile); tile += gridDim.x, tl++) {
      const uint32_t st = tl % NAS;
      mbar_wait(BAR(FB_AFULL + st), (tl / NAS) & 1);
#pragma unroll 1
      for (int s2 = 0; s2 < F_CH * 2; s2++) {
        const int c = ffn_step<NHD>(s2) >> 1;
        const bool is_up = (ffn_step<NHD>(s2) & 1) == 0;
        if (is_up) {      // up projection of chunk c: K slabs two at a time
          const uint32_t Hd = Hd0 + 128 * (c % NHD);
#pragma unroll 1
          for (int kp = 0; kp < F_KS / 2; kp++) {
            get_w2();
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad = umma_desc(base + F_OFF_A + (st * F_KS + 2 * kp) * SLAB);
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) MMA_SS(Hd, ad + 2 * k4, wd0 + 2 * k4, (kp > 0 || k4 > 0) ? 1u : 0u);
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) MMA_SS(Hd, ad + (SLAB >> 4) + 2 * k4, wd1 + 2 * k4, 1u);
              COMMIT(FB_WEMPTY + ws0);
              COMMIT(FB_WEMPTY + ws1);
              if (kp == F_KS / 2 - 1) {
                COMMIT(FB_HIDFULL + (c % NHD));
                if (c == F_CH - 1) COMMIT(FB_AEMPTY + st);
              }
            }
            __syncwarp();
          }
        }
        if (!is_up) {  // down projection of chunk c: one output block (both K halves) per iteration
          const int cc = c, hb = cc % NHD;
          const uint32_t Hd = Hd0 + 128 * hb;
          mbar_wait(BAR(FB_HSREADY + hb), nhid[hb] & 1);      // conversion of this chunk done
          nhid[hb]++;
          if (cc == 0) mbar_wait(BAR(FB_ACCFREE), (tl & 1) ^ 1);      // D drained by the epilogue of the previous tile
#pragma unroll 1
          for (int nb = 0; nb < F_NB; nb++) {
            get_w2();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) MMA_TS(Dt + 128 * nb, Hd + 8 * k4, wd0 + 2 * k4, (cc > 0 || k4 > 0) ? 1u : 0u);
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) MMA_TS(Dt + 128 * nb, Hd + 32 + 8 * k4, wd1 + 2 * k4, 1u);
              COMMIT(FB_WEMPTY + ws0);
              COMMIT(FB_WEMPTY + ws1);
              if (cc == F_CH - 1 && nb == F_NB - 1) COMMIT(FB_ACCFULL);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 4) {
    // ===================================================== DRAIN: hidden chunk fp32 -> + b1 -> relu -> bf16, in place in TMEM
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    auto cvt32 = [&](const uint32_t (&v)[32], const float* bias32, uint32_t dst) {
      uint32_t p[16];
      const float4* bb = reinterpret_cast<const float4*>(bias32);
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const float4 b = bb[t];
        p[2 * t] = pack_bf16_relu(__uint_as_float(v[4 * t]) + b.x, __uint_as_float(v[4 * t + 1]) + b.y);
        p[2 * t + 1] = pack_bf16_relu(__uint_as_float(v[4 * t + 2]) + b.z, __uint_as_float(v[4 * t + 3]) + b.w);
      }
      TC_ST16(dst, p);
    };
    uint32_t nh[2] = {0, 0};
    for (int tile = blockIdx.x; TILE_OK(tile); tile += gridDim.x) {
#pragma unroll 1
      for (int c = 0; c < F_CH; c++) {
        const int hb = c % NHD;
        mbar_wait(BAR(FB_HIDFULL + hb), nh[hb] & 1);
        nh[hb]++;
        tc_fence_after();
        uint32_t va[32], vb[32];
        const float* bias = sB1 + rotc(c) * 128;
        const uint32_t t0 = Hd0 + 128 * hb + lane_base;
        TC_LD32(t0, va);
        TC_LD32(t0 + 32, vb);
        tc_wait_ld();
        cvt32(va, bias, t0);
        TC_LD32(t0 + 64, va);
        cvt32(vb, bias + 32, t0 + 16);
        TC_LD32(t0 + 96, vb);
        tc_wait_ld();
        cvt32(va, bias + 64, t0 + 32);
        cvt32(vb, bias + 96, t0 + 48);
        tc_wait_st();
        tc_fence_before();
        ARRIVE_LEADER(FB_HSREADY + hb);
      }
    }
  } else if (warp < 8) {
    // ===================================================== A producers (next tile): LayerNorm -> bf16 slabs, warp q rows 32 q .. + 31
    const int q = warp - 4;
    const int hr = lane >> 4, c16 = lane & 15;
    const uint64_t pol_last = l2_policy_evict_last();
    uint32_t tl = 0;
    for (int tile = blockIdx.x; TILE_OK(tile); tile += gridDim.x, tl++) {
      const int64_t row0 = (int64_t)tile * TM + 32 * q;
      // (x is read again by the EPI warps one to two tile periods later: evict_last here, evict_first there)
      // two-pass statistics, 16 lanes per row, the row in registers (4 float4 per lane), 2 row pairs in flight
      const float* xs = a.x + 4 * c16;
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 4) {
        float4 v[2][F_KS];
#pragma unroll
        for (int p = 0; p < 2; p++) {
          int64_t row = row0 + r0 + 2 * p + hr;
          row = row < a.R ? row : a.R - 1;
#pragma unroll
          for (int j = 0; j < F_KS; j++) v[p][j] = FFN_LDX(reinterpret_cast<const float4*>(xs + (size_t)row * F_H + 64 * j));
        }
        float sum[2], sq[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
          float t = 0.f;
#pragma unroll
          for (int j = 0; j < F_KS; j++) t += (v[p][j].x + v[p][j].y) + (v[p][j].z + v[p][j].w);
          sum[p] = t;
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
#pragma unroll
          for (int p = 0; p < 2; p++) sum[p] += __shfl_xor_sync(0xffffffffu, sum[p], o);
#pragma unroll
        for (int p = 0; p < 2; p++) {
          const float mu = sum[p] * (1.0f / F_H);
          float t = 0.f;
#pragma unroll
          for (int j = 0; j < F_KS; j++) {
            const float dx = v[p][j].x - mu, dy = v[p][j].y - mu, dz = v[p][j].z - mu, dw = v[p][j].w - mu;
            t += (dx * dx + dy * dy) + (dz * dz + dw * dw);
          }
          sq[p] = t;
          sum[p] = mu;
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
#pragma unroll
          for (int p = 0; p < 2; p++) sq[p] += __shfl_xor_sync(0xffffffffu, sq[p], o);
        if (c16 == 0) {
#pragma unroll
          for (int p = 0; p < 2; p++) stats[32 * q + r0 + 2 * p + hr] = make_float2(sum[p], ln_rstd(sq[p] * (1.0f / F_H), a.eps, a.eps_mode));
        }
      }
      __syncwarp();
      const uint32_t st = tl % NAS;
      mbar_wait(BAR(FB_AEMPTY + st), ((tl / NAS) & 1) ^ 1);
#pragma unroll 1
      for (int ks = 0; ks < F_KS; ks++) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma + 64 * ks + 4 * c16));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.beta + 64 * ks + 4 * c16));
        uint8_t* A = sm + F_OFF_A + (st * F_KS + ks) * SLAB;
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 16) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            int64_t row = row0 + i0 + 2 * u + hr;
            row = row < a.R ? row : a.R - 1;
            v[u] = FFN_LDX(reinterpret_cast<const float4*>(xs + (size_t)row * F_H + 64 * ks));
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int r = 32 * q + i0 + 2 * u + hr;
            const float2 ms = stats[r];
            float4 t = v[u];
            t.x = (t.x - ms.x) * ms.y * g4.x + b4.x;
            t.y = (t.y - ms.x) * ms.y * g4.y + b4.y;
            t.z = (t.z - ms.x) * ms.y * g4.z + b4.z;
            t.w = (t.w - ms.x) * ms.y * g4.w + b4.w;
            if (row0 + i0 + 2 * u + hr >= a.R) t = f4zero();
            uint2 pk;
            pk.x = pack_bf16(t.x, t.y);
            pk.y = pack_bf16(t.z, t.w);
            *reinterpret_cast<uint2*>(A + sw_off(r, 4 * c16)) = pk;
          }
        }
      }
      fence_async_smem();
      ARRIVE_LEADER(FB_AFULL + st);
    }
  } else if (warp < 12) {
    // ===================================================== EPI (TMEM lane quadrant = warp % 4): y = D + b2 + x + h
    // The accumulator is single buffered (TMEM: D 256 + 2 x 128 hidden columns), so the window in which it is busy is serial
    // with the next tile's down projections.  Phase 1 keeps it short: TMEM -> registers -> y (raw accumulator, fragment-layout
    // sector-exact stores, no loads), then D is released.  Phase 2 runs beside the next tile's MMAs: the same warp re-reads its
    // 32 rows from L2 as coalesced 1 KB rows and adds x, h and b2 (4 rows in flight).
    const int dq = warp & 3;
    const uint32_t lane_base = ((uint32_t)(dq * 32)) << 16;
    const int qr = lane >> 2, cq = 2 * (lane & 3);
    const uint64_t pol_first = l2_policy_evict_first();
    float4 b2v[F_NB];
#pragma unroll
    for (int s2 = 0; s2 < F_NB; s2++) b2v[s2] = __ldg(reinterpret_cast<const float4*>(a.b2) + 32 * s2 + lane);
    uint32_t tl = 0;
    // Experiment (GNB_FFN_EPI_PREFETCH=1, off): phase 2 is a chain of dependent memory round trips per group of rows (ncu: x is no
    // longer in L2 by then and h has not been touched at all), so an elected lane asks for the x and h rows of this warp's NEXT
    // tile as two bulk L2 prefetches while the current tile is drained.  It made the kernel 4 % slower: the memory system, not
    // the latency of these loads, is what bounds it (the extra requests only add contention / early evictions).
    auto prefetch_rows = [&](int t) {
      const int64_t r0 = (int64_t)t * TM + 32 * dq;
      if (t - (int)rank < a.num_tiles && r0 < a.R && elect_one()) {
        const int64_t nr = a.R - r0 < 32 ? a.R - r0 : 32;
        bulk_prefetch_l2(a.x + (size_t)r0 * F_H, (uint32_t)(nr * F_H * sizeof(float)));
        bulk_prefetch_l2(a.h + (size_t)r0 * F_H, (uint32_t)(nr * F_H * sizeof(float)));
      }
      __syncwarp();
    };
    if (GNB_FFN_EPI_PREFETCH) prefetch_rows(blockIdx.x);
    for (int tile = blockIdx.x; TILE_OK(tile); tile += gridDim.x, tl++) {
      const int64_t row0 = (int64_t)tile * TM + 32 * dq;
      if (GNB_FFN_EPI_PREFETCH) prefetch_rows(tile + gridDim.x);
      mbar_wait(BAR(FB_ACCFULL), tl & 1);
      tc_fence_after();
#pragma unroll 1
      for (int st2 = 0; st2 < 2 * (F_H / 64); st2++) {      // (64-column chunk, row half)
        const int ch = st2 >> 1, hh = st2 & 1;
        uint32_t dreg[32];
        TC_LD_FRAG64(Dt + lane_base + ((uint32_t)(16 * hh) << 16) + 64 * ch, dreg);
        tc_wait_ld();
        if (st2 == 2 * (F_H / 64) - 1) {
          tc_fence_before();
          ARRIVE_LEADER(FB_ACCFREE);
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
          const int64_t row = row0 + 16 * hh + 8 * h2 + qr;
          if (row < a.R) {
            float* o = a.y + (size_t)row * F_H + 64 * ch + cq;
#pragma unroll
            for (int n = 0; n < 8; n++) *reinterpret_cast<float2*>(o + 8 * n) = make_float2(__uint_as_float(dreg[4 * n + 2 * h2]), __uint_as_float(dreg[4 * n + 2 * h2 + 1]));
          }
        }
      }
      __syncwarp();      // the rows below were written by other lanes of this warp
      const int64_t left = a.R - row0;
      const int rows = left < 0 ? 0 : (left > 32 ? 32 : (int)left);
      constexpr int ER = F_NB == 2 ? GNB_FFN_EPI_ROWS : 2;      // rows in flight per warp
#pragma unroll 1
      for (int i0 = 0; i0 < rows; i0 += ER) {
        float4 yy[ER][F_NB], xx[ER][F_NB], hh2[ER][F_NB];      // ER rows x F_NB 512-byte segments in flight
#pragma unroll
        for (int u = 0; u < ER; u++) {
          const int64_t r = row0 + (i0 + u < rows ? i0 + u : rows - 1);
          const size_t o = (size_t)r * F_H + 4 * lane;
#pragma unroll
          for (int s2 = 0; s2 < F_NB; s2++) {
            yy[u][s2] = __ldcg(reinterpret_cast<const float4*>(a.y + o + 128 * s2));
            xx[u][s2] = GNB_FFN_L2_HINTS ? ld_stream_hint(a.x + o + 128 * s2, pol_first) : __ldg(reinterpret_cast<const float4*>(a.x + o + 128 * s2));      // last read of x
            hh2[u][s2] = GNB_FFN_L2_HINTS ? ld_stream_hint(a.h + o + 128 * s2, pol_first) : __ldg(reinterpret_cast<const float4*>(a.h + o + 128 * s2));     // only read of h
          }
        }
#pragma unroll
        for (int u = 0; u < ER; u++) {
          if (i0 + u < rows) {
            const size_t o = (size_t)(row0 + i0 + u) * F_H + 4 * lane;
#pragma unroll
            for (int s2 = 0; s2 < F_NB; s2++) {
              const float4 X = xx[u][s2], Hh = hh2[u][s2], Y = yy[u][s2], Bv = b2v[s2];
              const float4 out4 = make_float4(((X.x + Hh.x) + Y.x) + Bv.x, ((X.y + Hh.y) + Y.y) + Bv.y, ((X.z + Hh.z) + Y.z) + Bv.z, ((X.w + Hh.w) + Y.w) + Bv.w);
              if (GNB_FFN_Y_STREAM) __stcs(reinterpret_cast<float4*>(a.y + o + 128 * s2), out4);
              else *reinterpret_cast<float4*>(a.y + o + 128 * s2) = out4;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();      // the peer may still be reading this CTA's shared / tensor memory
  if (warp == 12) {
    if (CL2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
#undef TILE_OK
}

// packed bf16 weight slabs of one Dense, built on first use and cached per (model, weight block, group size)
static int get_pack(gnb_ctx* ctx, const PackKey& key, const PackSrc& ps, int ldw, int K, int Nout, int NG, const __nv_bfloat16** out) {
  if (ctx->tc_lin_nocache) {
    // per-call pack (training step): the slabs live in a grow-only scratch of the context; the next call overwrites them in stream
    // order, after the GEMM that reads them
    const size_t bytes = (size_t)K * Nout * sizeof(__nv_bfloat16);
    if (ctx->pack_ws_bytes < bytes) {
      if (ctx->pack_ws) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->pack_ws); ctx->pack_ws = nullptr; ctx->pack_ws_bytes = 0; }
      if (cudaMalloc(&ctx->pack_ws, bytes) != cudaSuccess) {
        cudaGetLastError();
        gnb_set_error("tc_gemm: cudaMalloc(%zu) for the per-call weight pack failed", bytes);
        return GNB_ERR_OOM;
      }
      ctx->pack_ws_bytes = bytes;
    }
    k_pack_lin<<<(unsigned)ceil_div((int64_t)K * Nout, 256), 256, 0, ctx->stream>>>(ps, ldw, K / 64, NG, Nout / 128 / NG, (__nv_bfloat16*)ctx->pack_ws);
    GNB_CUDA(cudaGetLastError());
    ctx->launches++;
    *out = (const __nv_bfloat16*)ctx->pack_ws;
    return GNB_OK;
  }
  if (!ctx->lin_cache) ctx->lin_cache = new PackCache();
  PackCache* cache = static_cast<PackCache*>(ctx->lin_cache);
  std::lock_guard<std::mutex> lk(cache->mu);
  auto itc = cache->m.find(key);
  if (itc == cache->m.end()) {
    __nv_bfloat16* dst = nullptr;
    const size_t elems = (size_t)K * Nout;
    if (cudaMalloc((void**)&dst, elems * sizeof(__nv_bfloat16)) != cudaSuccess) {
      cudaGetLastError();
      gnb_set_error("tc_gemm: cudaMalloc(%zu) for packed weights failed", elems * 2);
      return GNB_ERR_OOM;
    }
    k_pack_lin<<<(unsigned)ceil_div((int64_t)elems, 256), 256, 0, ctx->stream>>>(ps, ldw, K / 64, NG, Nout / 128 / NG, dst);
    GNB_CUDA(cudaGetLastError());
    ctx->launches++;
    itc = cache->m.emplace(key, dst).first;
  }
  *out = itc->second;
  return GNB_OK;
}

}  // namespace

void tc_lin_cache_evict(void* cache, uint64_t model_id) {
  if (!cache) return;
  PackCache* c = static_cast<PackCache*>(cache);
  std::lock_guard<std::mutex> lk(c->mu);
  for (auto it = c->m.begin(); it != c->m.end();) {
    if (it->first.model == model_id) { cudaFree(it->second); it = c->m.erase(it); }
    else ++it;
  }
}

void tc_lin_cache_free(void* cache) {
  if (!cache) return;
  PackCache* c = static_cast<PackCache*>(cache);
  for (auto& kv : c->m) cudaFree(kv.second);
  delete c;
}

bool tc_lin_supported(const LinArgs& a) {
  if (a.R < 256 || a.R > 0x7fffffffLL || a.Nout < 128 || (a.Nout & 127) || a.nsrc < 1) return false;
  int K = 0;
  for (int s = 0; s < a.nsrc; s++) {
    if (a.src[s].d <= 0 || (a.src[s].d & 63) || (a.src[s].ldx & (a.src[s].x_bf16 ? 7 : 3))) return false;
    if (a.src[s].x_bf16 && a.src[s].gamma) return false;
    if (a.src[s].gamma && ((a.src[s].d & 127) || a.src[s].d > 512)) return false;
    K += a.src[s].d;
  }
  if (K < 128 || (a.ldo & 1)) return false;
  for (int j = 0; j < a.nadd; j++)
    if (a.add[j].lda & 1) return false;
  return true;
}

int launch_linear_tc(gnb_ctx* ctx, const LinArgs& a) {
  if (a.R <= 0) return GNB_OK;
  if (ctx_first(ctx, ONCE_TC_LIN)) GNB_CUDA(cudaFuncSetAttribute(k_tc_lin, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
  GemmArgs g{};
  g.R = a.R; g.Nout = a.Nout; g.ldo = a.ldo; g.nsrc = a.nsrc;
  int K = 0;
  PackKey key{};
  key.model = ctx->cur_model_id; key.Nout = a.Nout; key.ldw = a.ldw; key.ng = (a.Nout / 128) % 2 == 0 ? 2 : 1;
  PackSrc ps{};
  ps.nsrc = a.nsrc;
  for (int s = 0; s < 3; s++) { ps.d[s] = 1 << 30; }
  for (int s = 0; s < a.nsrc; s++) {
    g.x[s] = a.src[s].x; g.ldx[s] = a.src[s].ldx; g.d[s] = a.src[s].d; g.xbf[s] = a.src[s].x_bf16;
    g.gamma[s] = a.src[s].gamma; g.beta[s] = a.src[s].beta; g.eps[s] = a.src[s].eps; g.eps_mode[s] = a.src[s].eps_mode;
    key.W[s] = a.src[s].W; key.d[s] = a.src[s].d;
    ps.W[s] = a.src[s].W; ps.d[s] = a.src[s].d;
    K += a.src[s].d;
  }
  for (int s = a.nsrc; s < 3; s++) g.d[s] = 1 << 30;
  g.KS = K / 64;
  const int nblk = a.Nout / 128;
  g.NG = (nblk % 2 == 0) ? 2 : 1;
  g.ngroups = nblk / g.NG;
  g.num_tiles = (int)ceil_div(a.R, TM);
  g.resident = (g.KS <= NA - 2 && g.ngroups > 1) ? 1 : 0;
  g.bias = a.bias; g.nadd = a.nadd; g.relu = a.relu; g.out = a.out; g.out_bf16 = a.out_bf16;
  for (int j = 0; j < a.nadd; j++) { g.add[j] = a.add[j].a; g.add_idx[j] = a.add[j].idx; g.lda[j] = a.add[j].lda; }
  GNB_TRY(get_pack(ctx, key, ps, a.ldw, K, a.Nout, g.NG, &g.wpack));
  double bytes = 2.0 * K * a.Nout + 4.0 * (double)a.R * a.Nout * a.nadd + (a.out_bf16 ? 2.0 : 4.0) * a.R * a.Nout;
  for (int s = 0; s < a.nsrc; s++) bytes += (a.src[s].x_bf16 ? 2.0 : 4.0) * a.R * a.src[s].d;
  // profile tag per layer shape (interned: Launch keeps the pointer)
  static std::mutex names_mu;
  static std::map<std::pair<int, int>, std::string> names;      // std::map: element addresses are stable
  std::unique_lock<std::mutex> nlk(names_mu);
  auto nit = names.find({K, a.Nout});
  if (nit == names.end()) nit = names.emplace(std::make_pair(K, a.Nout), "tc_linear_k" + std::to_string(K) + "_n" + std::to_string(a.Nout)).first;
  const char* shape_name = nit->second.c_str();
  nlk.unlock();
  Launch L(ctx, ctx->profiling && getenv("GNB_PROFILE_SHAPES") ? shape_name : "tc_linear", bytes, 2.0 * a.R * K * a.Nout);
  const int grid = g.num_tiles < ctx->sm_count ? g.num_tiles : ctx->sm_count;
  g.wd = ctx_watch(ctx);
  k_tc_lin<<<grid, G_THREADS, G_SMEM, ctx->stream>>>(g);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

bool tc_ffn256_supported(int64_t R, int d) { return (d == 256 || d == 384) && R >= 256 && R <= 0x7fffffffLL; }

template <int FH>
static int launch_ffn_t(gnb_ctx* ctx, int once_key, int64_t R, const gnb_ffn_params& f, const gnb_ln_params& ln2, const float* x, const float* h, float* y) {
  if (ctx_first(ctx, once_key)) {
    GNB_CUDA(cudaFuncSetAttribute(k_tc_ffn<FH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnCfg<FH>::SMEM));
    GNB_CUDA(cudaFuncSetAttribute(k_tc_ffn<FH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnCfg<FH>::SMEM));
  }
  FfnArgs a{};
  a.x = x; a.h = h; a.y = y; a.R = R; a.num_tiles = (int)ceil_div(R, TM);
  a.gamma = ln2.gamma; a.beta = ln2.beta; a.eps = ln2.eps; a.eps_mode = ln2.eps_mode;
  a.b1 = f.b1; a.b2 = f.b2;
  a.wd = ctx_watch(ctx);
  PackKey k1{}, k2{};
  PackSrc p1{}, p2{};
  for (int s = 0; s < 3; s++) { p1.d[s] = p2.d[s] = 1 << 30; }
  k1.model = k2.model = ctx->cur_model_id;
  k1.W[0] = f.W1; k1.d[0] = FH; k1.Nout = 4 * FH; k1.ldw = 4 * FH; k1.ng = 1;
  k2.W[0] = f.W2; k2.d[0] = 4 * FH; k2.Nout = FH; k2.ldw = FH; k2.ng = 1;
  p1.nsrc = p2.nsrc = 1;
  p1.W[0] = f.W1; p1.d[0] = FH;
  p2.W[0] = f.W2; p2.d[0] = 4 * FH;
  GNB_TRY(get_pack(ctx, k1, p1, 4 * FH, FH, 4 * FH, 1, &a.w1));
  GNB_TRY(get_pack(ctx, k2, p2, FH, 4 * FH, FH, 1, &a.w2));
  // canonical work of the reference FFN: 16 d^2 flop per row, x / h read and y written once
  Launch L(ctx, FH == 256 ? "tc_ffn256" : "tc_ffn384", 4.0 * 3 * FH * R, 16.0 * FH * FH * R);
  // GNB_FFN_CTA_PAIR=0: one CTA per tile stream (cta_group::1).  Read per launch so that a test can run both instantiations.
  const char* pair_env = getenv("GNB_FFN_CTA_PAIR");
  const bool cl2 = (pair_env ? atoi(pair_env) : 1) && ctx->sm_count >= 2;
  if (cl2) {
    const int pairs = (a.num_tiles + 1) / 2, max_clusters = ctx->sm_count / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (pairs < max_clusters ? pairs : max_clusters)); cfg.blockDim = dim3(F_THREADS);
    cfg.dynamicSmemBytes = FfnCfg<FH>::SMEM; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GNB_CUDA(cudaLaunchKernelEx(&cfg, k_tc_ffn<FH, true>, a));
  } else {
    const int grid = a.num_tiles < ctx->sm_count ? a.num_tiles : ctx->sm_count;
    k_tc_ffn<FH, false><<<grid, F_THREADS, FfnCfg<FH>::SMEM, ctx->stream>>>(a);
  }
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

int launch_ffn256_tc(gnb_ctx* ctx, int64_t R, const gnb_ffn_params& f, const gnb_ln_params& ln2, const float* x, const float* h, float* y, int d) {
  if (R <= 0) return GNB_OK;
  if (d == 256) return launch_ffn_t<256>(ctx, ONCE_TC_FFN, R, f, ln2, x, h, y);
  return launch_ffn_t<384>(ctx, ONCE_TC_FFN384, R, f, ln2, x, h, y);
}
