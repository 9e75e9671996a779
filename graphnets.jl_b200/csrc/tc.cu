// tcgen05 / TMEM / bulk-async (TMA engine) bf16 path for GNCore layers with hidden width 128.
//
// Algebra (all exact rewrites of src/gnblock.jl:63-69 + src/gncore.jl:56-59 by linearity of Dense):
//   h_e   = We_e' ê + Ps[src] + Pr[dst] + Pu[g]           ê = LayerNorm-normalised edge row, affine folded into W
//   y_e   = x_e + h_e + W2 relu(W1' ê + b1') + b2          one TMEM accumulator: D = ê We_e' + relu(.) W2
//   agg_v = sum_{e->v} h_e = We_e' (sum ê) + sum (Ps+Pr+Pu)   "aggregate, then transform": the edge kernel only
//                                                          emits per-32-row partial sums of ê and of the gathered
//                                                          addends (deterministic order, no atomics)
//
// k_core (persistent, warp specialised, one CTA per SM, a PAIR of 128-row tiles in lock step so that every
// streamed weight block feeds two UMMA tiles):
//   warps 0-3 / 4-7   drain + epilogue group of sub-tile 0 / 1 (TMEM lane quadrant = warp & 3)
//   warps 8-15        prologue, one pair AHEAD of the MMAs: coalesced row loads, LayerNorm, bf16 A operand into
//                     128B-swizzled K-major smem (double buffered), gathered addends + residual written as y0,
//                     partial receiver sums
//   warp 16           MMA issuer: converged warp, one elected lane, descriptors in uniform registers
//   warp 17           weight loader: cp.async.bulk of pre-swizzled 16 KB half blocks through a 5-stage ring
// TMEM per sub-tile: D (128 cols, fp32 accumulator) | Hd (128 cols: FFN hidden chunk fp32, rewritten IN PLACE as
// bf16 by the drain warps and consumed as the TMEM A operand of the down-projection; the 4H hidden activation
// never touches shared or global memory).
#include "tc_ptx.cuh"
#include "tc.cuh"
#include "graphrows.cuh"
#include "tc_edge.cuh"

using namespace tcx;

namespace {

// =====================================================================================================
// Projection kernel: out = A . W for NBLK weight blocks (N = 128 NBLK), fp32 out; A = LayerNorm-normalised rows of x (affine
// folded into W) -> P_s | P_r of the nodes.  `addend` (optional) is added to one 128-column block of the output.
// 10 warps: 0-3 drain, 4-7 A-operand producers, 8 MMA issuer, 9 weight loader.  One 128-row tile per iteration.
// =====================================================================================================
enum { SRC_LN = 0 };

struct ProjArgs {
  const float* x;                 // [R][H]
  const float* addend;            // optional [*][H], added to output columns [add_col0, add_col0 + H)
  const int32_t* addend_idx;      // row of `addend` per output row (nullptr: the row itself)
  int add_col0;
  float* out;                     // [R][nblk*H] fp32, or (out_bf16) the same matrix in bf16
  int out_bf16;
  int64_t R;
  int num_tiles;
  int nblk;                       // weight blocks (1 or 2)
  const __nv_bfloat16* wpack;
  float eps;
  int eps_mode;
  WatchArgs wd;                   // kernel watchdog (tc_ptx.cuh)
  unsigned int dbg_drain_delay_ns;   // test hook: stall the drain warps per tile (gnb_ctx::dbg_proj_drain_delay_ns)
};

// smem: A[2] (32 KB each) | W[2 NBLK] (32 KB each, resident: NBLK bf16 "hi" blocks, then their NBLK "lo" blocks) | barriers
// Weights are split  W = hi + lo  (hi = bf16(W), lo = bf16(W - hi)) and every output block is D = A.hi + A.lo: the rounding
// error of a bf16 WEIGHT is the same for every row, so it does not average out in the sums over hundreds of edges / nodes that
// feed the graph update - it was the dominant (coherent) error of the graph features (tools/dec_sensitivity.py).  The rounding
// of the ACTIVATIONS is independent per row and stays; node-level kernels are cheap (N rows), so the second MMA is free.
__host__ __device__ constexpr int pj_off_w() { return 2 * BLK_BYTES; }
__host__ __device__ constexpr int pj_off_misc(int nblk) { return (2 + 2 * nblk) * BLK_BYTES; }
__host__ __device__ constexpr int pj_smem(int nblk) { return pj_off_misc(nblk) + 256 + 1024; }
// Barrier protocol (every waiter sits inside the back-pressure loop of the barrier it waits on, so a 1-bit phase parity can
// never alias - no producer can complete phase n + 1 of a barrier before every waiter has observed phase n):
//   AFULL[st]   producers (128 arrivals) -> MMA warp          next phase needs AEMPTY[st] <- MMA warp after it saw AFULL
//   AEMPTY[st]  MMA commit -> producers                       next phase needs AFULL[st]  <- producers after they saw AEMPTY
//   ASEEN[st]   MMA warp (after it saw AFULL) -> drain warps  next phase needs ACCFREE[st] <- drain warps after they saw ASEEN
//   OUTDONE[st] MMA commit -> drain warps                     next phase needs ACCFREE[st] <- drain warps after they saw OUTDONE
//   ACCFREE[st] drain warps (128 arrivals) -> MMA warp        next phase needs OUTDONE[st] <- MMA warp after it saw ACCFREE
// Round 1 let the drain warps wait on AFULL directly (their addends could come from the producers): they are NOT in AFULL's loop -
// the producers only need the MMA warp to refill a stage - so a drain warp delayed by more than one tile found AFULL two
// phases ahead, read the aliased parity as "not yet complete" and waited forever (the dead-lock of the two-context mode,
// reproduced with GNB_DEBUG_PROJ_DRAIN_DELAY_NS under -DGNB_OLD_PROJ_PROTOCOL, tests/test_gpu_watchdog.py).
enum { PB_WFULL = 0, PB_AFULL = 2, PB_AEMPTY = 4, PB_OUTDONE = 6, PB_ACCFREE = 8, PB_ASEEN = 10 };
constexpr int PJ_THREADS = 10 * 32;

// Pipelined over tiles: warps 4-7 build the bf16 A operand of tile t+1 (double buffered) while warp 8 issues the MMAs
// of tile t and warps 0-3 drain the accumulator of tile t-1 (TMEM double buffered); weights stay resident.
template <int SRC, int NBLK>
__global__ void __launch_bounds__(PJ_THREADS, NBLK == 1 ? 2 : 1) k_tc_proj(const ProjArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t sW = base + pj_off_w();
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + pj_off_misc(NBLK));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  bool wd_dead = false;      // kernel watchdog (tc_ptx.cuh)
#define mbar_wait(b, p) mbar_wait_w((b), (p), wd_dead, a.wd)
  constexpr uint32_t TCOLS = 256u * NBLK;     // 2 accumulator sets
  if (tid == 0) {
    mbar_init(BAR(PB_WFULL), 1); mbar_init(BAR(PB_WFULL + 1), 1);
    for (int s = 0; s < 2; s++) {
      mbar_init(BAR(PB_AFULL + s), 128); mbar_init(BAR(PB_AEMPTY + s), 1);
      mbar_init(BAR(PB_OUTDONE + s), 1); mbar_init(BAR(PB_ACCFREE + s), 128);
      mbar_init(BAR(PB_ASEEN + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 9) {
    if (elect_one()) {
      for (int b = 0; b < NBLK; b++) {      // barrier b: hi and lo block of output block b
        mbar_expect_tx(BAR(PB_WFULL + b), 2 * BLK_BYTES);
        bulk_g2s(sW + b * BLK_BYTES, reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)b * BLK_BYTES, BLK_BYTES, BAR(PB_WFULL + b));
        bulk_g2s(sW + (NBLK + b) * BLK_BYTES, reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)(NBLK + b) * BLK_BYTES, BLK_BYTES, BAR(PB_WFULL + b));
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    uint32_t tl = 0;
    for (int b = 0; b < NBLK; b++) mbar_wait(BAR(PB_WFULL + b), 0);
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x, tl++) {
      const uint32_t st = tl & 1, ph = (tl >> 1) & 1;
      mbar_wait(BAR(PB_AFULL + st), ph);
      mbar_wait(BAR(PB_ACCFREE + st), ph ^ 1);
      tc_fence_after();
      if (lane == 0) mbar_arrive(BAR(PB_ASEEN + st));      // release: forwards whatever the producers published to the drain warps
      __syncwarp();
      if (elect_one()) {
        const uint64_t adesc = umma_desc(base + st * BLK_BYTES);
        for (int b = 0; b < NBLK; b++) {
          const uint64_t w = umma_desc(sW + b * BLK_BYTES), wl = umma_desc(sW + (NBLK + b) * BLK_BYTES);
          issue_ss(tmem + (st * NBLK + b) * 128, adesc, w, w + (KB_BYTES >> 4), false);
          issue_ss(tmem + (st * NBLK + b) * 128, adesc, wl, wl + (KB_BYTES >> 4), true);
        }
        tc_commit(BAR(PB_AEMPTY + st));
        tc_commit(BAR(PB_OUTDONE + st));
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================================================== A-operand producers, one tile ahead
    const int q = warp - 4;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x, tl++) {
      const uint32_t st = tl & 1, ph = (tl >> 1) & 1;
      const int64_t row0 = (int64_t)tile * TM + 32 * q;
      const int64_t left = a.R - row0;
      const int wrows = left < 0 ? 0 : (left > 32 ? 32 : (int)left);
      uint8_t* A = sm + st * BLK_BYTES;
      {
        // 8 lanes per row (tc_edge.cu LN warps): 4 rows x 128 B per load instruction, 3 shuffle levels per statistic
        const int rr = lane >> 3, l8 = lane & 7;
        const float* xbase = a.x + 4 * l8;
        const uint32_t a_lane = (uint32_t)((32 * q) * 128 + (l8 & 1) * 8);
        const uint32_t a_chunk = (uint32_t)(l8 >> 1);
        float4 xa[2][4];
        auto issue = [&](int i0) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            int64_t r = row0 + i0 + 4 * h + rr;
            r = r < a.R ? r : a.R - 1;
            const float4* p = reinterpret_cast<const float4*>(xbase + (size_t)r * H);
#pragma unroll
            for (int j = 0; j < 4; j++) xa[h][j] = __ldg(p + 8 * j);
          }
        };
        issue(0);
        mbar_wait(BAR(PB_AEMPTY + st), ph ^ 1);
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 8) {
          float4 xc[2][4];
          float s[2];
#pragma unroll
          for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int j = 0; j < 4; j++) xc[h][j] = xa[h][j];
            const float4 t = f4add(f4add(xc[h][0], xc[h][1]), f4add(xc[h][2], xc[h][3]));
            s[h] = (t.x + t.y) + (t.z + t.w);
          }
          if (i0 + 8 < 32) issue(i0 + 8);
#pragma unroll
          for (int o = 1; o < 8; o <<= 1)
#pragma unroll
            for (int h = 0; h < 2; h++) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const float mu = s[h] * (1.0f / H);
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              xc[h][j].x -= mu; xc[h][j].y -= mu; xc[h][j].z -= mu; xc[h][j].w -= mu;
              t += (xc[h][j].x * xc[h][j].x + xc[h][j].y * xc[h][j].y) + (xc[h][j].z * xc[h][j].z + xc[h][j].w * xc[h][j].w);
            }
            s[h] = t;
          }
#pragma unroll
          for (int o = 1; o < 8; o <<= 1)
#pragma unroll
            for (int h = 0; h < 2; h++) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int i = i0 + 4 * h + rr;
            const float rs = (i < wrows) ? ln_rstd(s[h] * (1.0f / H), a.eps, a.eps_mode) : 0.f;
            const uint32_t rowoff = a_lane + (uint32_t)i * 128;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              uint2 pk;
              pk.x = pack_bf16(xc[h][j].x * rs, xc[h][j].y * rs);
              pk.y = pack_bf16(xc[h][j].z * rs, xc[h][j].w * rs);
              *reinterpret_cast<uint2*>(A + (j >> 1) * KB_BYTES + rowoff + (((a_chunk + 4 * (j & 1)) ^ (uint32_t)(i & 7)) << 4)) = pk;
            }
          }
        }
      }
      fence_async_smem();
      mbar_arrive(BAR(PB_AFULL + st));
    }
  } else {
    // ===================================================== accumulator drain (TMEM lane quadrant = warp), fragment layout:
    // every 4 lanes write one full 32 B sector
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const int ld = NBLK * H;
    const int q = lane >> 2, cq = 2 * (lane & 3);
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x, tl++) {
      const uint32_t st = tl & 1, ph = (tl >> 1) & 1;
      const int64_t row0 = (int64_t)tile * TM;
      const int rows = (int)((a.R - row0) < TM ? (a.R - row0) : TM);
      if (a.dbg_drain_delay_ns) __nanosleep(a.dbg_drain_delay_ns);      // test hook
      // the drain warps may prefetch their addends once the A tile is complete
#ifdef GNB_OLD_PROJ_PROTOCOL
      mbar_wait(BAR(PB_AFULL + st), ph);      // round-1 protocol, kept only to reproduce its dead-lock
#else
      mbar_wait(BAR(PB_ASEEN + st), ph);
#endif
      int64_t ar[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        int64_t r = row0 + warp * 32 + q + 8 * k;
        r = r < a.R ? r : a.R - 1;
        ar[k] = (a.addend && a.addend_idx) ? (int64_t)__ldg(a.addend_idx + r) : r;
      }
      float2 t[2][8];
      auto fetch = [&](int stp, int h2, float2 (&dst)[8]) {
        const int hh = stp & 1, ch = stp >> 1;
        const bool has = a.addend != nullptr && 64 * ch >= a.add_col0;
        const float* ad = a.addend + (size_t)ar[2 * hh + h2] * H + (has ? (64 * ch - a.add_col0) : 0) + cq;
#pragma unroll
        for (int n = 0; n < 8; n++) dst[n] = has ? *reinterpret_cast<const float2*>(ad + 8 * n) : make_float2(0.f, 0.f);
      };
      fetch(0, 0, t[0]);
      fetch(0, 1, t[1]);
      mbar_wait(BAR(PB_OUTDONE + st), ph);
      tc_fence_after();
#pragma unroll 1
      for (int stp = 0; stp < 4 * NBLK; stp++) {
        const int hh = stp & 1, ch = stp >> 1;
        uint32_t d[32];
        TC_LD_FRAG64(tmem + (st * NBLK) * 128 + lane_base + ((uint32_t)(16 * hh) << 16) + 64 * ch, d);
        tc_wait_ld();
        if (stp == 4 * NBLK - 1) {
          tc_fence_before();
          mbar_arrive(BAR(PB_ACCFREE + st));
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
          const int r = warp * 32 + 16 * hh + q + 8 * h2;
          float2 v[8];
#pragma unroll
          for (int n = 0; n < 8; n++)
            v[n] = make_float2(__uint_as_float(d[4 * n + 2 * h2]) + t[h2][n].x, __uint_as_float(d[4 * n + 2 * h2 + 1]) + t[h2][n].y);
          if (stp + 1 < 4 * NBLK) fetch(stp + 1, h2, t[h2]);     // next step's addends fly during this step's stores
          if (r < rows) {
            if (a.out_bf16) {      // rows consumed as gathered addends by the fused edge kernel: half the bytes to write and to gather
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + (size_t)(row0 + r) * ld + 64 * ch + cq;
#pragma unroll
              for (int n = 0; n < 8; n++) *reinterpret_cast<uint32_t*>(o + 8 * n) = pack_bf16(v[n].x, v[n].y);
            } else {
              float* o = a.out + (size_t)(row0 + r) * ld + 64 * ch + cq;
#pragma unroll
              for (int n = 0; n < 8; n++) *reinterpret_cast<float2*>(o + 8 * n) = v[n];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
#undef mbar_wait
}

// =====================================================================================================
// Node aggregate + its projection in one pass (replaces round 1's tc_agg + tc_agg_proj launches):
//   P_agg[v] = W_na agg_v,   agg_v = W_ee' (sum_{e->v} ê_e) + sum_{e->v} (P_s[src] + P_r'[dst])          (src/nodefninput.jl:3)
//            = (sum ê) F + (sum G) W_na            F = W_ee' W_na folded (fp32) when the model is packed
// A = [ bf16(sum of the node's E_part rows) | bf16(sum of its G_part rows) ]  (K = 256), one 128 x 128 accumulator; the weights
// are split hi + lo like k_tc_proj's (4 resident blocks), so there is room for ONE A stage only.
// agg itself is never materialised: the graph update only needs sum_v agg_v, which by linearity is
// (sum_v sum ê) W_ee' + sum_v sum G - so the producers also emit the ordered sums of both operands per (16-node block,
// graph) run (SE_part, SG_part; index node_gpart) and k_graph_post finishes them in fp32.
// 14 warps: 0-3 drain, 4-11 producers (16 rows each = one cut of node_gpart), 12 MMA issuer, 13 weight loader.
// Barriers as k_tc_proj (no ASEEN: the drain warps read nothing the producers write).
// =====================================================================================================
struct Agg2Args {
  const float* Epart;             // [n_parts][H]  (part_bf16: bf16 rows)
  const float* Gpart;             // [n_parts][H]
  int part_bf16;
  const int32_t* part_ptr;        // [R+1] node -> its partial rows [part_ptr[v], part_ptr[v+1])
  const int32_t* gpart;           // [R]   node -> (16-node block, graph) run id (gnb_graph::node_gpart)
  float* SEpart;                  // out [n_nparts][H]
  float* SGpart;                  // out [n_nparts][H]
  float* out;                     // out [R][H]  P_agg
  int64_t R;
  int num_tiles;
  const __nv_bfloat16* wpack;     // 4 blocks: F hi, W_na hi, F lo, W_na lo (hi / lo weight split, see k_tc_proj)
  WatchArgs wd;
};
constexpr int AG_THREADS = 14 * 32;
constexpr int AG_STAGE = 2 * BLK_BYTES;                      // A0 | A1
constexpr int AG_NSTAGE = 1;                                 // one A stage: the hi / lo weight split (k_tc_proj) needs 4 resident blocks
constexpr int AG_OFF_W = AG_NSTAGE * AG_STAGE;
constexpr int AG_OFF_MISC = AG_OFF_W + 4 * BLK_BYTES;
constexpr int AG_SMEM = AG_OFF_MISC + 256 + 1024;
enum { AB_WFULL = 0, AB_AFULL = 1, AB_AEMPTY = 3, AB_OUTDONE = 5, AB_ACCFREE = 7 };

template <bool PBF>      // PBF: the partial rows are bf16 (Agg2Args::part_bf16)
__global__ void __launch_bounds__(AG_THREADS, 1) k_tc_agg2(const Agg2Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t sW = base + AG_OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + AG_OFF_MISC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  bool wd_dead = false;      // kernel watchdog (tc_ptx.cuh)
#define mbar_wait(b, p) mbar_wait_w((b), (p), wd_dead, a.wd)
  if (tid == 0) {
    mbar_init(BAR(AB_WFULL), 1);
    mbar_init(BAR(AB_AFULL), 8); mbar_init(BAR(AB_AEMPTY), 1);
    for (int s = 0; s < 2; s++) { mbar_init(BAR(AB_OUTDONE + s), 1); mbar_init(BAR(AB_ACCFREE + s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 13) {
    if (elect_one()) {
      mbar_expect_tx(BAR(AB_WFULL), 4 * BLK_BYTES);      // F hi, W_na hi, F lo, W_na lo
      for (int b = 0; b < 4; b++)
        bulk_g2s(sW + b * BLK_BYTES, reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)b * BLK_BYTES, BLK_BYTES, BAR(AB_WFULL));
    }
    __syncwarp();
  } else if (warp == 12) {
    uint32_t tl = 0;
    mbar_wait(BAR(AB_WFULL), 0);
    const uint64_t w0 = umma_desc(sW), w1 = umma_desc(sW + BLK_BYTES), w0l = umma_desc(sW + 2 * BLK_BYTES), w1l = umma_desc(sW + 3 * BLK_BYTES);
    const uint64_t a0 = umma_desc(base), a1 = umma_desc(base + BLK_BYTES);      // the single A stage
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x, tl++) {
      const uint32_t st = tl & 1, ph = (tl >> 1) & 1;      // accumulator (TMEM) double buffer
      mbar_wait(BAR(AB_AFULL), tl & 1);
      mbar_wait(BAR(AB_ACCFREE + st), ph ^ 1);
      tc_fence_after();
      if (elect_one()) {
        issue_ss(tmem + st * 128, a0, w0, w0 + (KB_BYTES >> 4), false);
        issue_ss(tmem + st * 128, a0, w0l, w0l + (KB_BYTES >> 4), true);
        issue_ss(tmem + st * 128, a1, w1, w1 + (KB_BYTES >> 4), true);
        issue_ss(tmem + st * 128, a1, w1l, w1l + (KB_BYTES >> 4), true);
        tc_commit(BAR(AB_AEMPTY));
        tc_commit(BAR(AB_OUTDONE + st));
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================================================== producers: warp pw owns rows 16 pw .. 16 pw + 15 of the tile
    const int pw = warp - 4;
    // this lane's 4 columns of partial row p (fp32 rows of 512 B, or bf16 rows of 256 B)
    auto ldE = [&](size_t p) {
      if constexpr (PBF) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(a.Epart) + p * (H / 4) + lane); return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u)); }
      else return __ldg(reinterpret_cast<const float4*>(a.Epart) + p * (H / 4) + lane);
    };
    auto ldG = [&](size_t p) {
      if constexpr (PBF) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(a.Gpart) + p * (H / 4) + lane); return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u)); }
      else return __ldg(reinterpret_cast<const float4*>(a.Gpart) + p * (H / 4) + lane);
    };
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x, tl++) {
      const int64_t row0 = (int64_t)tile * TM + 16 * pw;
      const int64_t left = a.R - row0;
      const int rows = left < 0 ? 0 : (left > 16 ? 16 : (int)left);
      int p0 = 0, p1 = 0, my_gp = -1;
      if (lane < rows) {
        p0 = __ldg(a.part_ptr + row0 + lane);
        p1 = __ldg(a.part_ptr + row0 + lane + 1);
        my_gp = __ldg(a.gpart + row0 + lane);
      }
      const int nxt = __shfl_down_sync(0xffffffffu, my_gp, 1);
      const uint32_t endmask = __ballot_sync(0xffffffffu, lane < rows && (lane == rows - 1 || nxt != my_gp));
      int gp = __shfl_sync(0xffffffffu, my_gp, 0);
      uint8_t* A0 = sm;
      uint8_t* A1 = A0 + BLK_BYTES;
      float4 accE = f4zero(), accG = f4zero();
      mbar_wait(BAR(AB_AEMPTY), (tl & 1) ^ 1);      // the MMAs of the previous tile have read the (single) A stage
#pragma unroll 1
      for (int i0 = 0; i0 < 16; i0 += 4) {
        // partial rows of consecutive nodes are consecutive in memory (parts are numbered in edge order, edges are
        // receiver-sorted): up to 2 partial rows per node are fetched unconditionally, 4 nodes (16 loads) in flight
        float4 s[4], t[4], s2[4], t2[4];
        int q0[4], q1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          q0[u] = __shfl_sync(0xffffffffu, p0, i0 + u);
          q1[u] = __shfl_sync(0xffffffffu, p1, i0 + u);
          const bool h0 = q0[u] < q1[u], h1 = q0[u] + 1 < q1[u];
          s[u] = h0 ? ldE((size_t)q0[u]) : f4zero();
          t[u] = h1 ? ldE((size_t)q0[u] + 1) : f4zero();
          s2[u] = h0 ? ldG((size_t)q0[u]) : f4zero();
          t2[u] = h1 ? ldG((size_t)q0[u] + 1) : f4zero();
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u;
          s[u] = f4add(s[u], t[u]);
          s2[u] = f4add(s2[u], t2[u]);
          for (int p = q0[u] + 2; p < q1[u]; p++) {
            s[u] = f4add(s[u], ldE((size_t)p));
            s2[u] = f4add(s2[u], ldG((size_t)p));
          }
          uint2 pk;
          pk.x = pack_bf16(s[u].x, s[u].y); pk.y = pack_bf16(s[u].z, s[u].w);
          *reinterpret_cast<uint2*>(A0 + sw_off(16 * pw + i, 4 * lane)) = pk;      // rows past the end are zero (no parts)
          pk.x = pack_bf16(s2[u].x, s2[u].y); pk.y = pack_bf16(s2[u].z, s2[u].w);
          *reinterpret_cast<uint2*>(A1 + sw_off(16 * pw + i, 4 * lane)) = pk;
          // ordered sums per (16-node block, graph) run, in fp32
          accE = f4add(accE, s[u]);
          accG = f4add(accG, s2[u]);
          const bool fl = (endmask >> i) & 1u;      // warp-uniform
          if (fl) {
            *(reinterpret_cast<float4*>(a.SEpart + (size_t)gp * H) + lane) = accE;
            *(reinterpret_cast<float4*>(a.SGpart + (size_t)gp * H) + lane) = accG;
            accE = f4zero(); accG = f4zero();
            gp++;
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(AB_AFULL));
    }
  } else {
    // ===================================================== accumulator drain (TMEM lane quadrant = warp), fragment layout:
    // every 4 lanes write one full 32 B sector
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const int q = lane >> 2, cq = 2 * (lane & 3);
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles && !wd_dead; tile += gridDim.x, tl++) {
      const uint32_t st = tl & 1, ph = (tl >> 1) & 1;
      const int64_t row0 = (int64_t)tile * TM;
      const int rows = (int)((a.R - row0) < TM ? (a.R - row0) : TM);
      mbar_wait(BAR(AB_OUTDONE + st), ph);
      tc_fence_after();
#pragma unroll 1
      for (int stp = 0; stp < 4; stp++) {
        const int hh = stp & 1, ch = stp >> 1;
        uint32_t d[32];
        TC_LD_FRAG64(tmem + st * 128 + lane_base + ((uint32_t)(16 * hh) << 16) + 64 * ch, d);
        tc_wait_ld();
        if (stp == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(AB_ACCFREE + st));
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
          const int r = warp * 32 + 16 * hh + q + 8 * h2;
          if (r < rows) {
            float* o = a.out + (size_t)(row0 + r) * H + 64 * ch + cq;
#pragma unroll
            for (int n = 0; n < 8; n++)
              *reinterpret_cast<float2*>(o + 8 * n) = make_float2(__uint_as_float(d[4 * n + 2 * h2]), __uint_as_float(d[4 * n + 2 * h2 + 1]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
#undef mbar_wait
}

// F[k][n] = gamma[k] * sum_j A[k][j] * B[j][n]     (A = rows [0,H) of the edge Dense, B = rows [0,H) of the node Dense; all
// k-major H x H fp32): the fold of W_ee' W_na
__global__ void k_fold_matmul(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ gamma,
                              float* __restrict__ F) {
  const int k = blockIdx.x, n = threadIdx.x;      // H x H
  float s = 0.f;
  for (int j = 0; j < H; j++) s = fmaf(A[(size_t)k * H + j], B[(size_t)j * H + n], s);
  F[(size_t)k * H + n] = s * gamma[k];
}

// ------------------------------------------------------------------ weight packing
// dst block (bf16, swizzled smem image): B[n][k] = W[(n0+n) + ldw*(k0+k)] * (gamma ? gamma[k] : 1);  lo: the residual
// bf16(w - bf16(w)) of the hi / lo weight split
__global__ void k_pack_block(const float* __restrict__ W, int ldw, int n0, int k0, const float* __restrict__ gamma,
                             __nv_bfloat16* __restrict__ dst, int lo) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;   // 128 x 128
  if (idx >= 128 * 128) return;
  int k = idx >> 7, n = idx & 127;
  float w = W[(size_t)(n0 + n) + (size_t)ldw * (k0 + k)];
  if (gamma) w *= gamma[k];
  uint32_t off = sw_off(n, k);
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  dst[off >> 1] = lo ? __float2bfloat16_rn(w - __bfloat162float(hi)) : hi;
}
// out[n] = (b ? b[n] : 0) + sum_k W[(n0+n) + ldw*(k0+k)] * beta[k]     (LayerNorm shift folded into a bias)
__global__ void k_fold_bias(const float* __restrict__ W, int ldw, int n0, int k0, int K, const float* __restrict__ beta,
                            const float* __restrict__ b, int nout, float* __restrict__ out, int accumulate) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nout) return;
  float s = accumulate ? out[n] : (b ? b[n] : 0.f);
  for (int k = 0; k < K; k++) s += W[(size_t)(n0 + n) + (size_t)ldw * (k0 + k)] * beta[k];
  out[n] = s;
}

// Bias slabs of the fused kernel (tc_edge.cu): the FFN biases are added by the tensor core as one extra K = 16 step
// ones[128][16] . slab[n][16]^T with slab[n][0] = bf16(b[n]), slab[n][1] = bf16(b[n] - slab[n][0]) (hi / lo split: exact to 2^-17),
// zeros elsewhere.  Five slabs of 128 rows: the four hidden chunks of b1' (LN2 shift folded in) and b2.  Layout: K-major,
// NO swizzle, 8 x 16 B core matrices: element (n, k) at byte (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2.
__global__ void k_pack_bias(const float* __restrict__ b1f /*[512]*/, const float* __restrict__ b2 /*[128]*/, __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // 640 rows
  if (i >= 640) return;
  const float b = i < 512 ? b1f[i] : b2[i - 512];
  const int slab = i >> 7, n = i & 127;
  __nv_bfloat16* o = dst + (size_t)slab * 2048 + (n >> 3) * 128 + (n & 7) * 8;
  const __nv_bfloat16 hi = __float2bfloat16_rn(b);
  o[0] = hi;
  o[1] = __float2bfloat16_rn(b - __bfloat162float(hi));
}

}  // namespace

static unsigned long long* g_tc_dbg = nullptr;
extern "C" int gnb_debug_tc_timing(unsigned long long* out, int n) {
  // first call (out == nullptr): allocate + enable;  later calls copy the stamps of the last edge launch out
  if (!g_tc_dbg) {
    if (cudaMalloc((void**)&g_tc_dbg, 148 * 18 * 32 * 8) != cudaSuccess) return GNB_ERR_OOM;
    cudaMemset(g_tc_dbg, 0, 148 * 18 * 32 * 8);
  }
  if (out) {
    if (n > 148 * 18 * 32) n = 148 * 18 * 32;
    if (cudaMemcpy(out, g_tc_dbg, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return GNB_ERR_CUDA;
  }
  return GNB_OK;
}

struct TcCorePack {
  __nv_bfloat16* w = nullptr;   // [proj: Ps Pr hi, Ps Pr lo | agg2: F W_na hi, F W_na lo | edge 9 | node 9] blocks of 32 KB
  float* f = nullptr;           // folded fp32 vectors: cu_e[128] cu_n[128] b1f_e[512] b1f_n[512]
  const __nv_bfloat16 *w_proj, *w_agg, *w_edge, *w_node;
  float *cu_e, *cu_n, *b1f_e, *b1f_n;
  __nv_bfloat16* bias = nullptr;      // bias slabs of the fused kernel (k_pack_bias): [edge 5 x 4 KB | node 5 x 4 KB]
  const __nv_bfloat16 *bias_e, *bias_n;
};

bool tc_core_supported(int de, int dn, int dg) { return de == H && dn == H && dg == H; }

void tc_core_pack_free(TcCorePack* p) {
  if (!p) return;
  if (p->w) cudaFree(p->w);
  if (p->f) cudaFree(p->f);
  if (p->bias) cudaFree(p->bias);
  delete p;
}

int tc_core_pack(gnb_ctx* ctx, const gnb_block_params& blk, const gnb_ffn_params* ffn, const gnb_ln_params* ln1,
                 const gnb_ln_params* ln2, TcCorePack** out) {
  *out = nullptr;
  // LN1 and LN2 share one normalised A tile (their affine parts are folded into the weights), which
  // needs identical eps conventions; otherwise the layer stays on the fp32 path.
  for (int i = 0; i < 2; i++)
    if (ln1[i].eps != ln2[i].eps || ln1[i].eps_mode != ln2[i].eps_mode) return GNB_OK;
  TcCorePack* p = new TcCorePack();
  const size_t nblk = 4 + 4 + 9 + 9;
  if (cudaMalloc((void**)&p->w, nblk * BLK_BYTES) != cudaSuccess ||
      cudaMalloc((void**)&p->f, (128 + 128 + 512 + 512) * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void**)&p->bias, 2 * TC_BIAS_PACK_BYTES) != cudaSuccess) {
    cudaGetLastError();
    tc_core_pack_free(p);
    gnb_set_error("tc_core_pack: cudaMalloc failed");
    return GNB_ERR_OOM;
  }
  __nv_bfloat16* w = p->w;
  const size_t BE = BLK_BYTES / 2;   // elements per block
  p->w_proj = w; p->w_agg = w + 4 * BE; p->w_edge = w + 8 * BE; p->w_node = w + 17 * BE;
  p->cu_e = p->f; p->cu_n = p->f + 128; p->b1f_e = p->f + 256; p->b1f_n = p->f + 768;
  cudaStream_t st = ctx->stream;
  auto pack = [&](const float* W, int ldw, int n0, int k0, const float* gamma, __nv_bfloat16* dst, int lo = 0) {
    k_pack_block<<<64, 256, 0, st>>>(W, ldw, n0, k0, gamma, dst, lo);
  };
  // We: (128, 4*128) input rows [e | v_src | v_dst | u]; Wn: (128, 3*128) input rows [agg | v | u]
  const float *g1e = ln1[0].gamma, *g1n = ln1[1].gamma, *b1e = ln1[0].beta, *b1n = ln1[1].beta;
  for (int lo = 0; lo < 2; lo++) {
    pack(blk.We, H, 0, H, g1n, w + (2 * lo + 0) * BE, lo);          // P_s
    pack(blk.We, H, 0, 2 * H, g1n, w + (2 * lo + 1) * BE, lo);      // P_r
  }
  // agg2 blocks: F = W_ee' W_na (folded in fp32, then rounded once) and W_na (aggregate rows, no LayerNorm)
  float* Ftmp = nullptr;
  if (cudaMalloc((void**)&Ftmp, (size_t)H * H * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    tc_core_pack_free(p);
    gnb_set_error("tc_core_pack: cudaMalloc failed");
    return GNB_ERR_OOM;
  }
  k_fold_matmul<<<H, H, 0, st>>>(blk.We, blk.Wn, g1e, Ftmp);
  for (int lo = 0; lo < 2; lo++) {
    pack(Ftmp, H, 0, 0, nullptr, w + (4 + 2 * lo) * BE, lo);
    pack(blk.Wn, H, 0, 0, nullptr, w + (5 + 2 * lo) * BE, lo);
  }
  // fused kernels: block order documented at k_core
  // both fused kernels (tc_edge.cu) index the blocks as  W1_0 W_blk W2_0 W1_1 W2_1 W1_2 W2_2 W1_3 W2_3
  for (int kind = 0; kind < 2; kind++) {
    __nv_bfloat16* dst = w + (kind == 0 ? 8 : 17) * BE;
    for (int c = 0; c < 4; c++) {
      const int i1 = c == 0 ? 0 : 2 * c + 1;      // W1_c
      const int i2 = 2 * c + 2;                   // W2_c
      pack(ffn[kind].W1, 4 * H, c * H, 0, ln2[kind].gamma, dst + i1 * BE);   // W1 (4H, H): hidden unit c*128+n
      pack(ffn[kind].W2, H, 0, c * H, nullptr, dst + i2 * BE);               // W2 (H, 4H): k = hidden index
    }
    if (kind == 0) pack(blk.We, H, 0, 0, g1e, dst + 1 * BE);
    else pack(blk.Wn, H, 0, H, g1n, dst + 1 * BE);
  }
  // constants folded into the per-graph rows / FFN bias
  k_fold_bias<<<1, 128, 0, st>>>(blk.We, H, 0, 0, H, b1e, blk.be, H, p->cu_e, 0);
  k_fold_bias<<<1, 128, 0, st>>>(blk.We, H, 0, H, H, b1n, nullptr, H, p->cu_e, 1);
  k_fold_bias<<<1, 128, 0, st>>>(blk.We, H, 0, 2 * H, H, b1n, nullptr, H, p->cu_e, 1);
  k_fold_bias<<<1, 128, 0, st>>>(blk.Wn, H, 0, H, H, b1n, blk.bn, H, p->cu_n, 0);
  k_fold_bias<<<4, 128, 0, st>>>(ffn[0].W1, 4 * H, 0, 0, H, ln2[0].beta, ffn[0].b1, 4 * H, p->b1f_e, 0);
  k_fold_bias<<<4, 128, 0, st>>>(ffn[1].W1, 4 * H, 0, 0, H, ln2[1].beta, ffn[1].b1, 4 * H, p->b1f_n, 0);
  p->bias_e = p->bias; p->bias_n = p->bias + TC_BIAS_PACK_BYTES / 2;
  cudaMemsetAsync(p->bias, 0, 2 * TC_BIAS_PACK_BYTES, st);
  k_pack_bias<<<5, 128, 0, st>>>(p->b1f_e, ffn[0].b2, p->bias);
  k_pack_bias<<<5, 128, 0, st>>>(p->b1f_n, ffn[1].b2, p->bias + TC_BIAS_PACK_BYTES / 2);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(Ftmp);
  if (e != cudaSuccess) {
    gnb_set_error("tc_core_pack: %s", cudaGetErrorString(e));
    tc_core_pack_free(p);
    return GNB_ERR_CUDA;
  }
  *out = p;
  return GNB_OK;
}

template <int SRC, int NBLK>
static int launch_proj(gnb_ctx* ctx, const ProjArgs& a, const char* name, double flops, double bytes) {
  if (a.num_tiles <= 0) return GNB_OK;
  if (ctx_first(ctx, NBLK == 2 ? ONCE_PROJ_LN2 : ONCE_PROJ_OTHER))
    GNB_CUDA(cudaFuncSetAttribute(k_tc_proj<SRC, NBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, pj_smem(NBLK)));
  const int per_sm = NBLK == 1 ? 2 : 1;
  const int grid = a.num_tiles < per_sm * ctx->sm_count ? a.num_tiles : per_sm * ctx->sm_count;
  Launch L(ctx, name, bytes, flops);
  ProjArgs aw = a;
  aw.wd = ctx_watch(ctx);
  aw.dbg_drain_delay_ns = ctx->dbg_proj_drain_delay_ns;
  k_tc_proj<SRC, NBLK><<<grid, PJ_THREADS, pj_smem(NBLK), ctx->stream>>>(aw);
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}

int tc_core_forward(gnb_ctx* ctx, const gnb_graph* g, const TcCorePack* pk, const gnb_block_params& blk,
                    const gnb_ffn_params* ffn, const gnb_ln_params* ln1, const gnb_ln_params* ln2, const float* xe,
                    const float* xn, const float* xg, float* ye, float* yn, float* yg, TcPreRows pre_in, TcNextCore next,
                    TcPreRows pre_out, TcDecFuse dec) {
  const int64_t E = g->E, N = g->N, B = g->B;
  int rc = GNB_OK;
  const size_t nparts = (size_t)(g->n_parts > 0 ? g->n_parts : 1);
  float* Pue = pre_in.Pue ? pre_in.Pue : arena_ptr<float>(ctx->arena, (size_t)B * H, &rc);
  float* Pun = pre_in.Pue ? pre_in.Pun : arena_ptr<float>(ctx->arena, (size_t)B * H, &rc);
  // P_s | P_r' of the nodes in bf16: the fused edge kernel gathers two of these rows per EDGE, so their size sets its L2 -> SM
  // traffic and how many rows of loads fit in its registers; the rounding is independent per node and element (it averages out in
  // every sum, unlike the rounding of a weight - see the hi / lo split of the packed weights)
  // (GNB_PSR_FP32=1 keeps them in fp32: A/B toggle)
  static const bool psr_bf16 = getenv("GNB_PSR_FP32") == nullptr || atoi(getenv("GNB_PSR_FP32")) == 0;
  const size_t psr_es = psr_bf16 ? 2 : 4;
  uint8_t* Psr = arena_ptr<uint8_t>(ctx->arena, (size_t)N * 2 * H * psr_es, &rc);
  // partial rows of the edge kernel (sums over <= 16 rows, written by k_edge5, read once by k_tc_agg2) in bf16: they are 0.33 GB
  // of the kernel's 2.8 GB of DRAM traffic per launch in fp32; like P_s | P_r' the rounding is independent per row and element.
  // (GNB_PART_FP32=1: fp32, A/B toggle)
  static const bool part_bf16 = getenv("GNB_PART_FP32") == nullptr || atoi(getenv("GNB_PART_FP32")) == 0;
  float* Epart = reinterpret_cast<float*>(arena_ptr<uint8_t>(ctx->arena, nparts * H * (part_bf16 ? 2 : 4), &rc));
  float* Gpart = reinterpret_cast<float*>(arena_ptr<uint8_t>(ctx->arena, nparts * H * (part_bf16 ? 2 : 4), &rc));
  float* Pagg = arena_ptr<float>(ctx->arena, (size_t)N * H, &rc);
  const size_t nnparts = (size_t)(g->n_nparts > 0 ? g->n_nparts : 1);
  float* SEpart = arena_ptr<float>(ctx->arena, nnparts * H, &rc);
  float* SGpart = arena_ptr<float>(ctx->arena, nnparts * H, &rc);
  float* Vpart = arena_ptr<float>(ctx->arena, nnparts * H, &rc);
  float* Npart = arena_ptr<float>(ctx->arena, nnparts * H, &rc);
  if (rc != GNB_OK) return rc;

  // per-graph rows (fp32 CUDA cores, B rows): P_ue = W_eu LN1(gf) + be + folded LN shifts, P_un likewise
  if (!pre_in.Pue) {
    GraphPreArgs ga{};
    ga.xg = xg; ga.B = B; ga.gamma = ln1[2].gamma; ga.beta = ln1[2].beta; ga.eps = ln1[2].eps; ga.eps_mode = ln1[2].eps_mode;
    ga.Weu = blk.We + (size_t)3 * H * H; ga.Wnu = blk.Wn + (size_t)2 * H * H;
    ga.ce = pk->cu_e; ga.cn = pk->cu_n; ga.Pue = Pue; ga.Pun = Pun;
    GNB_TRY(launch_graph_pre(ctx, ga));
  }
  const double HH = (double)H * H;
  {  // node projections P_s | P_r
    ProjArgs a{};
    a.x = xn; a.out = reinterpret_cast<float*>(Psr); a.out_bf16 = psr_bf16; a.R = N; a.num_tiles = ceil_div(N, TM); a.nblk = 2; a.wpack = pk->w_proj;
    a.addend = Pue; a.addend_idx = g->node_graph; a.add_col0 = H;   // P_r += P_u[graph of the node]
    a.eps = ln1[1].eps; a.eps_mode = ln1[1].eps_mode;
    GNB_TRY((launch_proj<SRC_LN, 2>(ctx, a, "tc_node_proj", 2.0 * N * 2 * HH, 4.0 * N * H + (double)psr_es * N * 2 * H)));
  }
  {  // edges: GNBlock edge update + FFN + residual; partial receiver sums of the inputs
    EdgeArgs a{};
    a.x = xe; a.y = ye; a.R = E; a.num_tiles = ceil_div(E, TM); a.wpack = pk->w_edge;
    a.bias_pack = pk->bias_e; a.eps = ln1[0].eps; a.eps_mode = ln1[0].eps_mode;
    a.add1 = Psr; a.idx1 = g->edge_src; a.ld1 = 2 * H; a.add2 = Psr + H * psr_es; a.idx2 = g->edge_dst; a.ld2 = 2 * H; a.add_bf16 = psr_bf16;
    static const bool pf = getenv("GNB_EDGE_PPREFETCH") != nullptr && atoi(getenv("GNB_EDGE_PPREFETCH")) != 0;      // experiment toggle
    if (pf) { a.pf_row_graph = g->edge_graph; a.pf_graph_ptr = g->graph_node_ptr; }
    a.part = g->edge_part; a.Epart = Epart; a.Gpart = Gpart; a.part_bf16 = part_bf16; a.dbg = g_tc_dbg; a.wd = ctx_watch(ctx);
    a.decW = dec.W4; a.dec_out = dec.partial;      // fused narrow decoder: y_e is not stored (ye may be nullptr)
    // canonical work of the reference's edge update + edge FFN: 24 H^2 flops and
    // 8H bytes of features + 12 B of index per edge
    GNB_TRY(launch_edge5(ctx, a, "tc_edge_core", 24.0 * HH * E, (8.0 * H + 12.0) * E));
  }
  // edge -> node aggregate (src/nodefninput.jl:3) and its node-update projection in one pass, by linearity:
  // P_agg = W_na [ W_ee' (sum ê) + sum (Ps + Pr + Pu) ] = (sum ê) F + (sum G) W_na;  the per-graph sums of both operands go to
  // k_graph_post (agg itself is never materialised)
  {
    Agg2Args a{};
    a.Epart = Epart; a.Gpart = Gpart; a.part_bf16 = part_bf16; a.part_ptr = g->node_part_ptr; a.gpart = g->node_gpart;
    a.SEpart = SEpart; a.SGpart = SGpart; a.out = Pagg; a.R = N; a.num_tiles = ceil_div(N, TM); a.wpack = pk->w_agg;
    a.wd = ctx_watch(ctx);
    if (a.num_tiles > 0) {
      if (ctx_first(ctx, ONCE_AGG2)) {
        GNB_CUDA(cudaFuncSetAttribute(k_tc_agg2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AG_SMEM));
        GNB_CUDA(cudaFuncSetAttribute(k_tc_agg2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AG_SMEM));
      }
      const int grid = a.num_tiles < ctx->sm_count ? a.num_tiles : ctx->sm_count;
      Launch L(ctx, "tc_agg", 4.0 * (2.0 * nparts + N) * H, 2.0 * N * 2 * HH);
      if (part_bf16) k_tc_agg2<true><<<grid, AG_THREADS, AG_SMEM, ctx->stream>>>(a);
      else k_tc_agg2<false><<<grid, AG_THREADS, AG_SMEM, ctx->stream>>>(a);
      GNB_CUDA(cudaGetLastError());
    }
  }
  {  // nodes: same fused kernel; the node -> graph sum of the block output h_v = W_nv' v^ + P_agg + P_un[g] is taken by
     // linearity over the partial sums of v^ (V_part) and of the addends (N_part), so h_v is never materialised
    EdgeArgs a{};
    a.x = xn; a.y = yn; a.R = N; a.num_tiles = ceil_div(N, TM); a.wpack = pk->w_node;
    a.bias_pack = pk->bias_n; a.eps = ln1[1].eps; a.eps_mode = ln1[1].eps_mode;
    a.add1 = Pagg; a.idx1 = nullptr; a.ld1 = H; a.add2 = Pun; a.idx2 = g->node_graph; a.ld2 = H; a.add_bf16 = 0;
    a.part = g->node_gpart; a.Epart = Vpart; a.Gpart = Npart; a.part_bf16 = 0; a.dbg = nullptr; a.wd = ctx_watch(ctx);
    GNB_TRY(launch_edge5(ctx, a, "tc_node_core", 20.0 * HH * N, (8.0 * H + 8.0) * N));
  }
  // graphs (B rows, fp32 CUDA cores): sums over the graph's nodes, graph update, graph FFN + residual
  {
    GraphPostArgs ga{};
    ga.xg = xg; ga.B = B;
    ga.graph_npart_ptr = g->graph_npart_ptr; ga.Vpart = Vpart; ga.Npart = Npart; ga.SEpart = SEpart; ga.SGpart = SGpart;
    ga.Wee = blk.We; ga.g1e = ln1[0].gamma;
    ga.Wnv = blk.Wn + (size_t)H * H; ga.g1n = ln1[1].gamma;
    ga.g1 = ln1[2].gamma; ga.b1ln = ln1[2].beta; ga.eps1 = ln1[2].eps; ga.eps_mode1 = ln1[2].eps_mode;
    ga.g2 = ln2[2].gamma; ga.b2ln = ln2[2].beta; ga.eps2 = ln2[2].eps; ga.eps_mode2 = ln2[2].eps_mode;
    ga.Wg = blk.Wg; ga.bg = blk.bg; ga.W1 = ffn[2].W1; ga.b1 = ffn[2].b1; ga.W2 = ffn[2].W2; ga.b2 = ffn[2].b2;
    ga.yg = yg;
    if (next.pk && pre_out.Pue) {      // the per-graph rows of the next core, from y_u, in the same launch
      ga.next_gamma = next.ln1[2].gamma; ga.next_beta = next.ln1[2].beta; ga.next_eps = next.ln1[2].eps; ga.next_eps_mode = next.ln1[2].eps_mode;
      ga.next_Weu = next.blk->We + (size_t)3 * H * H; ga.next_Wnu = next.blk->Wn + (size_t)2 * H * H;
      ga.next_ce = next.pk->cu_e; ga.next_cn = next.pk->cu_n;
      ga.next_Pue = pre_out.Pue; ga.next_Pun = pre_out.Pun;
    }
    GNB_TRY(launch_graph_post(ctx, ga, N));
  }
  return GNB_OK;
}
