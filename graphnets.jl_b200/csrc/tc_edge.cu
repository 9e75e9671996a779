// Fused GNCore EDGE kernel for 128-wide features (tcgen05 / TMEM / bulk-async weights), generation 5.
//
//   y_e = x_e + [ W_ee' ê + P_s[src] + P_r'[dst] ] + [ W2 relu(W1' ê + b1') + b2 ]         ê = LayerNorm-normalised row
//         (src/gnblock.jl:65 by linearity of Dense, src/gnfeedforward.jl:27-31, src/gncore.jl:56-68)
//   E_part, G_part = ordered partial sums of ê and of the gathered addends per (32-row block, receiver) run
//         (the edge -> node aggregation of src/nodefninput.jl:3, "aggregate, then transform")
//
// One 128-row tile per pass; persistent CTA, one per SM, 16 warps (512 threads x 128 registers: the register file is split
// per scheduler, 4 x 16 K, so a 17th warp would cap every thread at 96 registers and spill all three data roles):
//   warps 8-11  DRAIN : FFN hidden chunk TMEM fp32 -> +b1 -> relu -> bf16 -> TMEM (in place, the A operand of the
//                       down-projection); finished accumulator TMEM -> swizzled shared-memory staging tile
//   warps 0-3   LN    : one tile AHEAD of the MMAs: coalesced row loads (8 rows in flight per warp), LayerNorm with
//                       transposed butterfly reductions, bf16 A operand into 128B-swizzled K-major shared memory,
//                       partial sums E_part
//   warps 4-7, 12, 13  OUT : one tile BEHIND, in 16-row slices (8 per tile, slice k = 8 pass + s goes to OUT warp k mod 6):
//                       staging row + x row + gathered P_s / P_r' rows, all 512 B coalesced, -> y (streaming stores), partial
//                       sums G_part.  The role is bound by the latency of its gathered loads (4 rows in flight per warp), so it
//                       gets every warp the register budget leaves (round 1: four warps, 12.7 k cycles per tile)
//   warp 14     MMA issuer (one elected lane), warp 15 weight loader (cp.async.bulk, 5 x 16 KB ring; also prefetches the x rows
//               of the pass after next into L2 with one bulk prefetch)
// TMEM (512 columns): D accumulator (128) | Hd[3] hidden chunk buffers (3 x 128), used round-robin by the stream of hidden
// chunks (4 per tile), so that up to three up-projections are in flight ahead of the conversion warps: the MMA warp never
// waits for a single conversion round trip (round 1 / early round 2: D[2] + Hd[2], where the tensor pipe idled half of
// every tile behind the serial chain  up(c) -> convert -> down(c) -> up(c+2)).  The accumulator is single buffered: it is
// drained (TMEM -> staging tile) while the first three up-projections of the next tile run.
// MMA order per tile: up0 up1 up2 blk dn0 up3 dn1 dn2 dn3 (SEQ_PACKED below).
#include "tc_ptx.cuh"
#include "tc_edge.cuh"
#include <stdlib.h>
#include <type_traits>

using namespace tcx;

namespace {

constexpr int HALF_BYTES = KB_BYTES;   // weight ring stage = one 64-wide K half of a block (16 KB)
// shared memory (CL2: per CTA of the pair).  The FFN biases are added by the tensor core: one extra K = 16 UMMA per up block /
// per GNBlock block,  ONES[128][16] (column 0 = 1) . BIAS slab (row n = [hi(b_n), lo(b_n), 0 ...]), so that the conversion warps
// (the serial link of the MMA chain) only have to relu + pack, and the OUT warps do not add b2.
constexpr int E_OFF_A = 0;                                   // 2 stages x 32 KB
constexpr int E_OFF_STG = 2 * BLK_BYTES;                     // 64 KB: [4 column groups][128 rows][128 B], 16B chunks XOR (row & 7)
constexpr int E_OFF_ONES = E_OFF_STG + 65536;                // 4 KB: A operand of the bias step (K-major, no swizzle, 8 x 16 B core matrices)
constexpr int E_OFF_BIAS = E_OFF_ONES + 4096;                // 5 bias slabs: this CTA's N rows (CL2: 64 of 128 -> 2 KB per slab)
__host__ __device__ constexpr int e_bias_bytes(bool cl2) { return 5 * (cl2 ? 2048 : 4096); }
__host__ __device__ constexpr int e_nws(bool cl2) { return cl2 ? 5 : 4; }       // weight ring stages of 16 KB (CL2: one block per stage; else two stages per block)
__host__ __device__ constexpr int e_off_w(bool cl2) { return E_OFF_BIAS + e_bias_bytes(cl2); }
__host__ __device__ constexpr int e_off_misc(bool cl2) { return e_off_w(cl2) + e_nws(cl2) * HALF_BYTES; }
constexpr int E_MISC = 40 * 8 + 16;      // barriers[40] (29 used), tmem slot
__host__ __device__ constexpr int e_smem(bool cl2) { return e_off_misc(cl2) + E_MISC + 1024; }
static_assert(e_off_w(true) % 1024 == 0 && e_off_w(false) % 1024 == 0, "swizzled weight stages need 1024 B alignment");
static_assert(e_smem(true) <= 232448 && e_smem(false) <= 232448, "shared memory budget (227 KB per CTA)");
// K-major operand WITHOUT swizzle: core matrices of 8 rows x 16 B; LBO = 128 B between the two core matrices of a K = 16 step,
// SBO = 256 B between 8-row groups (cute::UMMA::SmemDescriptor, LayoutType::INTERLEAVE)
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
}
#ifndef GNB_OUT_WARPS
#define GNB_OUT_WARPS 6
#endif
constexpr int OUT_WARPS = GNB_OUT_WARPS;      // 6: 16 warps x 128 registers;  8: 18 warps x 96 registers (one slice per OUT warp)
static_assert(OUT_WARPS == 6 || OUT_WARPS == 8, "OUT warps: 4-7 + 12,13 or 4-7 + 12-15");
constexpr int E_WARPS = 10 + OUT_WARPS;
constexpr int W_MMA = 8 + OUT_WARPS, W_LOAD = 9 + OUT_WARPS;
constexpr int OUT_ROWS = GNB_PART_ROWS;      // rows per OUT slice == cut of the partial-row index (lower.cu)
constexpr int OUT_SLICES = TM / OUT_ROWS;
static_assert(OUT_ROWS == 16 && OUT_SLICES == 8, "8 slices of 16 rows per tile");
constexpr int E_THREADS = E_WARPS * 32;
enum { EB_WFULL = 0, EB_WEMPTY = 5, EB_AFULL = 10, EB_AEMPTY = 12, EB_HIDFULL = 14, EB_HSREADY = 17, EB_OUTDONE = 20,
       EB_ACCFREE = 21, EB_STGFULL = 24, EB_STGEMPTY = 32 };      // STGFULL / STGEMPTY: one pair per 16-row slice of the staging tile

// block ids inside the packed edge weights (tc.cu::tc_core_pack): W1_0 W_blk W2_0 W1_1 W2_1 W1_2 W2_2 W1_3 W2_3
constexpr int PK_W1_0 = 0, PK_BLK = 1, PK_W2_0 = 2, PK_W1_1 = 3, PK_W2_1 = 4, PK_W1_2 = 5, PK_W2_2 = 6, PK_W1_3 = 7, PK_W2_3 = 8;
// issue order of a tile, one nibble per block:  up0 up1 up2 blk dn0 up3 dn1 dn2 dn3.  Three up-projections fill the three hidden
// buffers while the conversion warps work; up3 re-uses the buffer of chunk 0 right behind dn0.  Four SS <-> TS operand-mode
// switches of the tensor pipe per tile (~435 cycles each, profiles/r01_hwprobe.log T5).
constexpr unsigned long long SEQ_PACKED = (unsigned long long)PK_W1_0 | ((unsigned long long)PK_W1_1 << 4) | ((unsigned long long)PK_W1_2 << 8) |
    ((unsigned long long)PK_BLK << 12) | ((unsigned long long)PK_W2_0 << 16) | ((unsigned long long)PK_W1_3 << 20) |
    ((unsigned long long)PK_W2_1 << 24) | ((unsigned long long)PK_W2_2 << 28) | ((unsigned long long)PK_W2_3 << 32);
constexpr int B_BLK = 3, B_LASTA = 5;      // the GNBlock block is the first write of the accumulator; up3 is the last read of the A tile
#define IS_DN(b) (((b) == 4) | ((b) >= 6))
#define IS_UP(b) (((b) <= 2) | ((b) == 5))
#define CHUNK_OF(b) ((b) <= 2 ? (b) : ((b) == 4 ? 0 : ((b) == 5 ? 3 : (b) - 5)))      /* hidden chunk (0-3) of an up / down block */


#ifdef GNB_TC_TIMING
#define EDBG(k)                                                                                     \
  do {                                                                                              \
    if (a.dbg != nullptr && tl == 4 && lane == 0)                                                   \
      a.dbg[((size_t)blockIdx.x * 18 + warp) * 32 + (k)] = (unsigned long long)clock64();           \
  } while (0)
#else
#define EDBG(k) do { } while (0)
#endif

// packed fp32x2 (FADD2 / FMUL2 / FFMA2 on sm_100): two lanes of a float4 per instruction
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)), hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 adds4(float4 a, float s) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), ss), hi = __fadd2_rn(make_float2(a.z, a.w), ss);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 muls4(float4 a, float s) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __fmul2_rn(make_float2(a.x, a.y), ss), hi = __fmul2_rn(make_float2(a.z, a.w), ss);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// four bf16 (two packed words) -> fp32
__device__ __forceinline__ float4 bf4(uint2 v) {
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
}
// 1 / LayerNorm denominator without branches: e_add = eps^2 | 0 | eps and plus = 0 | eps | 0 for the three conventions
__device__ __forceinline__ float rstd_nb(float var, float e_add, float plus) {
  const float t = var + e_add;
  const float r = rsqrtf(fmaxf(t, 1e-38f));
  return plus > 0.f ? __frcp_rn(fmaf(t, r, plus)) : r;      // sqrt(t) = t * rsqrt(t)
}

// x is read twice by a CTA, two to three tile periods apart (LayerNorm warps, then the residual of the OUT warps).  ncu showed
// 45 % of the second reads missing L2 (0.48 GB of DRAM reads per launch): the first read (and the prefetch) mark the lines
// evict_last, the second one evict_first.
#ifndef GNB_X_L2_HINTS
#define GNB_X_L2_HINTS 1
#endif
#ifndef GNB_EDGE_XPREFETCH
#define GNB_EDGE_XPREFETCH 1      /* the weight loader's bulk L2 prefetch of the x rows two passes ahead; without it 851 vs 842 us per edge launch, forward 6.14 vs 6.03 ms (same box) */
#endif

// CL2: the two CTAs of a cluster (an SM pair) run one 256-row MMA stream (cta_group::2): each CTA holds half of every weight
// block (N split) - half the weight bytes per SM and a ring twice as deep in blocks - and its own 128-row tile otherwise.
// sum over the 32 lanes of 16 per-lane values at once (transposed butterfly: 8 + 4 + 2 + 1 + 1 shuffles instead of 80); lane l
// returns the total of value 8 b4 + 4 b3 + 2 b2 + b1 (b_i = bit i of l; both lanes of a pair hold the same total)
__device__ __forceinline__ float warp_sum16(const float (&v)[16], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
  float w[8], z[4], y2[2];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float send = h16 ? v[j] : v[j + 8], keep = h16 ? v[j + 8] : v[j];
    w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float send = h8 ? w[j] : w[j + 4], keep = h8 ? w[j + 4] : w[j];
    z[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const float send = h4 ? z[j] : z[j + 2], keep = h4 ? z[j + 2] : z[j];
    y2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = h2 ? y2[0] : y2[1], keep = h2 ? y2[1] : y2[0];
  float t = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

// DEC: fused narrow decoder - the OUT warps do not store y but y . decW (EdgeArgs::decW)
// PBF: the gathered addend rows are bf16 (EdgeArgs::add_bf16: the edges' P_s | P_r'), else fp32 (the nodes' P_agg, P_un)
template <bool CL2, bool DEC, bool PBF>
__global__ void __launch_bounds__(E_THREADS, 1) k_edge5(const __grid_constant__ EdgeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  constexpr int NWS = e_nws(CL2);
  const uint32_t sW = base + e_off_w(CL2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + e_off_misc(CL2));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = CL2 ? cluster_ctarank() : 0u;      // == blockIdx.x & 1 (clusters are consecutive block pairs)
  constexpr uint32_t NARR = CL2 ? 8u : 4u;                  // one arrival per warp (lane 0 after __syncwarp) on the barriers the 4-warp groups feed

  if (tid == 0) {
    // pair leader: WFULL[st] also counts the peer's "my half has landed" relay, so the MMA warp waits on ONE barrier per block
    for (int i = 0; i < NWS; i++) { mbar_init(BAR(EB_WFULL + i), (CL2 && rank == 0) ? 2 : 1); mbar_init(BAR(EB_WEMPTY + i), 1); }
    for (int s = 0; s < 2; s++) { mbar_init(BAR(EB_AFULL + s), NARR); mbar_init(BAR(EB_AEMPTY + s), 1); }
    for (int s = 0; s < 3; s++) { mbar_init(BAR(EB_HIDFULL + s), 1); mbar_init(BAR(EB_HSREADY + s), NARR); }
    mbar_init(BAR(EB_OUTDONE), 1); mbar_init(BAR(EB_ACCFREE), NARR);
    for (int i = 0; i < OUT_SLICES; i++) { mbar_init(BAR(EB_STGFULL + i), 1); mbar_init(BAR(EB_STGEMPTY + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    if (CL2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  {
    // ONES: element (row r, k = 0) = 1.0, everything else 0;  BIAS: this CTA's rows of the five packed slabs
    uint4* ones = reinterpret_cast<uint4*>(sm + E_OFF_ONES);      // 256 x 16 B: chunk (r >> 3) * 16 + (k >> 3) * 8 + (r & 7)
    for (int i = tid; i < 256; i += E_THREADS) ones[i] = make_uint4(((i >> 3) & 1) ? 0u : 0x3F80u, 0u, 0u, 0u);
    constexpr int SLAB = CL2 ? 2048 : 4096;
    const uint4* src = reinterpret_cast<const uint4*>(a.bias_pack);
    uint4* dstb = reinterpret_cast<uint4*>(sm + E_OFF_BIAS);
    for (int i = tid; i < 5 * SLAB / 16; i += E_THREADS) {
      const int slab = i / (SLAB / 16), j = i % (SLAB / 16);
      dstb[i] = __ldg(src + slab * (TC_BIAS_SLAB_BYTES / 16) + (CL2 ? (int)rank * (SLAB / 16) : 0) + j);
    }
    fence_async_smem();      // generic-proxy writes -> visible to the UMMAs (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();      // the peer's barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  // leader-side barriers fed by both CTAs: arrive through the cluster address of CTA 0
  // warp-level arrive: every lane has fenced its own writes; __syncwarp orders them before lane 0's release
  auto ARRIVE_LEADER = [&](int i) {
    __syncwarp();
    if (lane == 0) {
      if (CL2) mbar_arrive_cluster(map_to_cta(BAR(i), 0));
      else mbar_arrive(BAR(i));
    }
  };
  auto ARRIVE_LOCAL = [&](int i) {
    __syncwarp();
    if (lane == 0) mbar_arrive(BAR(i));
  };
  bool wd_dead = false;      // kernel watchdog (tc_ptx.cuh)
#define mbar_wait(b, p) mbar_wait_w((b), (p), wd_dead, a.wd)
#define mbar_wait_cluster(b, p) mbar_wait_w((b), (p), wd_dead, a.wd)      /* barriers that receive arrivals from the peer CTA */
#define TILE_OK(t) ((t) - (int)rank < a.num_tiles && !wd_dead)   /* both CTAs of a pair run the same number of passes */
  const uint32_t tmem = *tmem_slot;
  const uint32_t Hd0 = tmem + 128;      // hidden buffer i at Hd0 + 128 i; the accumulator at tmem
  const int grid = gridDim.x;
  const int first_tile = (int)blockIdx.x - (int)rank;      // both CTAs of a pair run the same number of passes
  const int npass = first_tile < a.num_tiles ? (a.num_tiles - first_tile + grid - 1) / grid : 0;

  if (warp == W_LOAD) {
    // ===================================================== weight loader
    uint32_t it = 0;
    auto load_block = [&](int blk) {
      if (CL2) {
        // one ring stage = this CTA's N half of the block: rows [64 rank, 64 rank + 64) of both 64-wide K halves
        const uint32_t st = it % NWS, ph = (it / NWS) & 1;
        it++;
        mbar_wait(BAR(EB_WEMPTY + st), ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(BAR(EB_WFULL + st), HALF_BYTES);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)blk * BLK_BYTES + (size_t)rank * 8192;
          bulk_g2s(sW + st * HALF_BYTES, src, 8192, BAR(EB_WFULL + st));
          bulk_g2s(sW + st * HALF_BYTES + 8192, src + HALF_BYTES, 8192, BAR(EB_WFULL + st));
        }
        __syncwarp();
        return;
      }
      for (int half = 0; half < 2; half++, it++) {
        const uint32_t st = it % NWS, ph = (it / NWS) & 1;
        mbar_wait(BAR(EB_WEMPTY + st), ph ^ 1);
        if (elect_one()) {
          const uint32_t nb = (uint32_t)HALF_BYTES;
          mbar_expect_tx(BAR(EB_WFULL + st), nb);
          bulk_g2s(sW + st * HALF_BYTES, reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)blk * BLK_BYTES + (size_t)half * HALF_BYTES,
                   nb, BAR(EB_WFULL + st));
        }
        __syncwarp();
      }
    };
    for (int tile = blockIdx.x; TILE_OK(tile); tile += grid) {
      // the x rows of the pass after next -> L2: the LayerNorm warps (one pass ahead of the MMAs) then load at L2 latency
      // instead of HBM latency
      const int64_t pre0 = ((int64_t)tile + 2 * (int64_t)grid) * TM;
      if (GNB_EDGE_XPREFETCH && pre0 < a.R && elect_one()) {
        const int64_t prows = a.R - pre0 < TM ? a.R - pre0 : TM;
        if (GNB_X_L2_HINTS)
          asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(a.x + (size_t)pre0 * H),
                       "r"((uint32_t)(prows * H * sizeof(float))), "l"(l2_policy_evict_last()) : "memory");
        else bulk_prefetch_l2(a.x + (size_t)pre0 * H, (uint32_t)(prows * H * sizeof(float)));
      }
      __syncwarp();
#pragma unroll 1
      for (int b = 0; b < 9; b++) load_block((SEQ_PACKED >> (4 * b)) & 15);
    }
  } else if (warp == W_MMA) {
    // ===================================================== MMA issuer (whole warp converged, one lane issues)
    uint32_t it = 0, tl = 0;
    uint64_t w0 = 0, w1 = 0;
    uint32_t st0 = 0, st1 = 0;
    if (CL2 && rank != 0) {
      // peer CTA: relay "my half of the weight block has landed" to the leader, in ring order
      for (int tile = blockIdx.x; TILE_OK(tile); tile += grid) {
#pragma unroll 1
        for (int b = 0; b < 9; b++, it++) {
          const uint32_t st = it % NWS, ph = (it / NWS) & 1;
          mbar_wait(BAR(EB_WFULL + st), ph);
          if (lane == 0) mbar_arrive_cluster(map_to_cta(BAR(EB_WFULL + st), 0));
          __syncwarp();
        }
      }
    } else {
    // ring position kept as (stage, phase) counters: no division in the issue loop
    uint32_t rst = 0, rph = 0;
    auto ring_next = [&]() { rst = rst + 1 == NWS ? 0 : rst + 1; rph ^= (rst == 0) ? 1u : 0u; };
    auto get_w = [&]() {
      if (CL2) {
        st0 = st1 = rst;
        mbar_wait_cluster(BAR(EB_WFULL + rst), rph);      // own half (expect_tx) + the peer's relay arrive
        w0 = umma_desc(sW + rst * HALF_BYTES);
        ring_next();
        return;
      }
      st0 = rst;
      mbar_wait(BAR(EB_WFULL + rst), rph);
      w0 = umma_desc(sW + rst * HALF_BYTES);
      ring_next();
      st1 = rst;
      mbar_wait(BAR(EB_WFULL + rst), rph);
      w1 = umma_desc(sW + rst * HALF_BYTES);
      ring_next();
    };
    auto COMMIT = [&](int i) {
      if (CL2) tc_commit2(BAR(i));
      else tc_commit(BAR(i));
    };
    constexpr uint32_t SLAB = CL2 ? 2048 : 4096;
    const uint64_t ones_desc = umma_desc_nosw(base + E_OFF_ONES), bias_desc = umma_desc_nosw(base + E_OFF_BIAS);
    // position of the next up / down block in the stream of hidden chunks: buffer (chunk mod 3) and use count parity
    uint32_t ub = 0, uq = 0, db = 0, dq = 0;
    for (int tile = blockIdx.x; TILE_OK(tile); tile += grid, tl++) {
      const uint32_t st = tl & 1, uph = (tl >> 1) & 1;
      const uint64_t adesc = umma_desc(base + E_OFF_A + st * BLK_BYTES);
      const uint32_t D = tmem;
      // A compact loop, NOT unrolled: the nine blocks unrolled are ~29 KB of straight-line code that every pass streams through
      // the SM's instruction cache, which the four other roles' loops need (ncu: instruction-cache hit rate 70 %, "no
      // instruction" the second largest stall reason of the kernel).  All branches are warp-uniform.
      int upc = 0;      // hidden chunk of the next up block (its bias slab)
#pragma unroll 1
      for (int b = 0; b < 9; b++) {
        EDBG(b);
        get_w();
        const bool is_dn = IS_DN(b), is_up = IS_UP(b);
        if (b == 0) mbar_wait_cluster(BAR(EB_AFULL + st), uph);                      // A tile of this pass
        if (b == B_BLK) mbar_wait_cluster(BAR(EB_ACCFREE), (tl & 1) ^ 1);            // accumulator of the previous tile drained
        if (is_dn) mbar_wait_cluster(BAR(EB_HSREADY + db), dq & 1);                  // hidden chunk converted to bf16
        tc_fence_after();
        if (elect_one()) {
          if (is_dn) {                                                             // D += relu(.)[chunk] W2_c
            const uint32_t Hd = Hd0 + 128 * db;
            if (CL2) issue_ts2(D, Hd, w0, true);
            else issue_ts(D, Hd, w0, w1, true);
          } else if (b == B_BLK) {                                                 // GNBlock GEMM (+ b2): first write of D
            if (CL2) { issue_ss2(D, adesc, w0, false); mma_ss2(D, ones_desc, bias_desc + 4 * (SLAB >> 4), IDESC2, 1u); }
            else { issue_ss(D, adesc, w0, w1, false); mma_ss(D, ones_desc, bias_desc + 4 * (SLAB >> 4), IDESC, 1u); }
          } else {                                                                 // FFN up-projection chunk (+ b1')
            const int c = upc;
            const uint32_t Hd = Hd0 + 128 * ub;
            if (CL2) { issue_ss2(Hd, adesc, w0, false); mma_ss2(Hd, ones_desc, bias_desc + c * (SLAB >> 4), IDESC2, 1u); }
            else { issue_ss(Hd, adesc, w0, w1, false); mma_ss(Hd, ones_desc, bias_desc + c * (SLAB >> 4), IDESC, 1u); }
            COMMIT(EB_HIDFULL + ub);
          }
          if (b == B_LASTA) COMMIT(EB_AEMPTY + st);                                // last read of the A tile
          if (b == 8) COMMIT(EB_OUTDONE);                                          // accumulator complete
          COMMIT(EB_WEMPTY + st0);
          if (!CL2) COMMIT(EB_WEMPTY + st1);
        }
        __syncwarp();
        if (is_up) { ub = ub == 2 ? 0 : ub + 1; uq += ub == 0 ? 1u : 0u; upc++; }
        if (is_dn) { db = db == 2 ? 0 : db + 1; dq += db == 0 ? 1u : 0u; }
      }
      EDBG(9);
    }
    }
  } else if (warp >= 8 && warp < 12) {
    // ===================================================== DRAIN warps (TMEM lane quadrant = warp)
    const int dq = warp - 8;                 // TMEM lane quadrant (== warp % 4)
    const uint32_t lane_base = ((uint32_t)(dq * 32)) << 16;
    uint32_t tl = 0;
    // TMEM fp32 -> relu -> bf16 pairs -> TMEM, in place (the bias is already in the accumulator: the MMA warp adds it with a
    // K = 16 step).  Two 32-column loads are kept in flight; the store of columns [16j,16j+16) only overwrites columns whose
    // load has completed.  (tcgen05.ld + wait is ~22 cycles even under a full UMMA queue: tools/micro/tmem_lat.cu)
    auto cvt32 = [&](const uint32_t (&v)[32], uint32_t dst) {
      uint32_t p[16];
#pragma unroll
      for (int t = 0; t < 16; t++) p[t] = pack_bf16_relu(__uint_as_float(v[2 * t]), __uint_as_float(v[2 * t + 1]));
      TC_ST16(dst, p);
    };
    auto convert = [&](uint32_t Hd) {
      uint32_t va[32], vb[32];
      const uint32_t t0 = Hd + lane_base;
      TC_LD32(t0, va);
      TC_LD32(t0 + 32, vb);
      tc_wait_ld();
      cvt32(va, t0);
      TC_LD32(t0 + 64, va);
      cvt32(vb, t0 + 16);
      TC_LD32(t0 + 96, vb);
      tc_wait_ld();
      cvt32(va, t0 + 32);
      cvt32(vb, t0 + 48);
      tc_wait_st();
      tc_fence_before();
    };
    uint32_t cb = 0, cq = 0;      // position in the stream of hidden chunks: buffer, use count
    auto conv = [&](int c) {
      mbar_wait(BAR(EB_HIDFULL + cb), cq & 1);
      EDBG(1 + 2 * c);
      tc_fence_after();
      convert(Hd0 + 128 * cb);
      ARRIVE_LEADER(EB_HSREADY + cb);
      EDBG(2 + 2 * c);
      cb = cb == 2 ? 0 : cb + 1; cq += cb == 0 ? 1u : 0u;
    };
    for (tl = 0; (int)tl < npass && !wd_dead; tl++) {
      EDBG(0);
#pragma unroll 1
      for (int c = 0; c < 4; c++) conv(c);      // one copy of the conversion code (instruction cache)
      // ---------------- finished accumulator -> staging tile (row r = this thread; 4 column groups of 32 fp32)
      mbar_wait(BAR(EB_OUTDONE), tl & 1);
      EDBG(9);
      tc_fence_after();
      // the staging tile is handed over per 16-row slice: this warp's rows are slices 2 dq and 2 dq + 1, free as soon as the
      // OUT warps have drained those two slices of the previous tile (not the whole tile)
      mbar_wait(BAR(EB_STGEMPTY + 2 * dq), (tl & 1) ^ 1);
      mbar_wait(BAR(EB_STGEMPTY + 2 * dq + 1), (tl & 1) ^ 1);
      EDBG(10);
      const uint32_t D = tmem;
      const int r = dq * 32 + lane;
      uint8_t* srow = sm + E_OFF_STG + r * 128;
#ifdef GNB_ABL_DRAIN_NOSTG
      tc_fence_before();
      ARRIVE_LEADER(EB_ACCFREE);
      if (a.R < 0)
#endif
      {
        uint32_t va[32], vb[32];
        auto put = [&](const uint32_t (&v)[32], int g) {
#pragma unroll
          for (int j = 0; j < 8; j++)
            *reinterpret_cast<uint4*>(srow + g * 16384 + ((j ^ (r & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        };
        TC_LD32(D + lane_base, va);
        TC_LD32(D + lane_base + 32, vb);
        tc_wait_ld();
        put(va, 0);
        TC_LD32(D + lane_base + 64, va);
        put(vb, 1);
        TC_LD32(D + lane_base + 96, vb);
        tc_wait_ld();
        tc_fence_before();          // TMEM fully read: the accumulator goes back to the MMA warp
        ARRIVE_LEADER(EB_ACCFREE);
        put(va, 2);
        put(vb, 3);
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(BAR(EB_STGFULL + 2 * dq)); mbar_arrive(BAR(EB_STGFULL + 2 * dq + 1)); }
      EDBG(11);
    }
  } else if (warp < 4) {
    // ===================================================== LN warps (0-3), one tile ahead of the MMAs
    // 8 lanes per row: lane (rr = lane >> 3, l8 = lane & 7) holds columns 4 l8 + 32 j .. +3 (j = 0..3) of row i0 + rr, so
    // one load instruction covers 4 rows x 128 contiguous bytes and a row statistic needs 3 shuffle levels instead of 5.
    const int q = warp & 3;
    const int rr = lane >> 3, l8 = lane & 7;
    const float* xbase = a.x + 4 * l8;
    // swizzled A operand, element (row 32q + i, k = 4 l8 + 32 j): byte (k >> 6) * 16 KB + row * 128 + ((chunk ^ (row & 7)) << 4) + (l8 & 1) * 8
    const uint32_t a_lane = (uint32_t)((32 * q) * 128 + (l8 & 1) * 8);
    const uint32_t a_chunk = (uint32_t)(l8 >> 1);            // + 4 (j & 1)
    // partial-sum pass: lane owns columns 4 lane .. 4 lane + 3 of every row of the slice
    const uint32_t p_lane = (uint32_t)((lane >> 4) * KB_BYTES + (32 * q) * 128 + (lane & 1) * 8);
    const uint32_t p_chunk = (uint32_t)((lane & 15) >> 1);
    const float e_add = a.eps_mode == GNB_EPS_SQRT_VAR_EPS2 ? a.eps * a.eps : (a.eps_mode == GNB_EPS_STD_PLUS_EPS ? 0.f : a.eps);
    const float e_plus = a.eps_mode == GNB_EPS_STD_PLUS_EPS ? a.eps : 0.f;
    const uint64_t pol_last = l2_policy_evict_last();
    uint32_t tl = 0;
    for (int tile = blockIdx.x; TILE_OK(tile); tile += grid, tl++) {
      const uint32_t stage = tl & 1, aph = (tl >> 1) & 1;
      const int64_t row0 = (int64_t)tile * TM + 32 * q;
      const int64_t left = a.R - row0;
      const int rows = left < 0 ? 0 : (left > 32 ? 32 : (int)left);
      int my_pid = -1;
      if (lane < rows) my_pid = __ldg(a.part + row0 + lane);
      const int nxt = __shfl_down_sync(0xffffffffu, my_pid, 1);
      const uint32_t endmask = __ballot_sync(0xffffffffu, lane < rows && (lane == rows - 1 || nxt != my_pid));
      int pid = __shfl_sync(0xffffffffu, my_pid, 0);
      EDBG(0);
      if (q == 0 && a.pf_row_graph != nullptr && left > 0) {
        // the OUT warps gather P_s[src] / P_r'[dst] of this tile two tile periods from now; every sender / receiver lies in the
        // contiguous node range of the graphs the tile touches: pull that range into L2 now (this warp idles at AEMPTY anyway;
        // issued from the weight loader warp instead, the dependent index loads delayed the weight ring: profiles/r02_summary.md)
        const int64_t t0 = (int64_t)tile * TM, tl_ = a.R - t0 < TM ? a.R - t0 : TM;
        if (lane == 0) {
          const int g0 = __ldg(a.pf_row_graph + t0), g1 = __ldg(a.pf_row_graph + t0 + tl_ - 1);
          const int64_t n0 = __ldg(a.pf_graph_ptr + g0), n1 = __ldg(a.pf_graph_ptr + g1 + 1);
          const int64_t rowb = (int64_t)a.ld1 * (a.add_bf16 ? 2 : 4);
          int64_t nb = (n1 - n0) * rowb;
          nb = nb > 131072 ? 131072 : nb;
          if (nb > 0) bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(a.add1) + (size_t)n0 * rowb, (uint32_t)nb);
        }
        __syncwarp();
      }
      float4 xa[2][4];   // two 4-row steps in flight
      auto issue = [&](int i0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          int64_t r = row0 + i0 + 4 * h + rr;
          r = r < a.R ? r : a.R - 1;      // rows past the end re-read the last row; masked below
          const float4* p = reinterpret_cast<const float4*>(xbase + (size_t)r * H);
#pragma unroll
          for (int j = 0; j < 4; j++) xa[h][j] = GNB_X_L2_HINTS ? ld_hint(p + 8 * j, pol_last) : __ldg(p + 8 * j);
        }
      };
#ifndef GNB_ABL_LN_NONE
      issue(0);
#endif
      mbar_wait(BAR(EB_AEMPTY + stage), aph ^ 1);
      EDBG(1);
      uint8_t* A = sm + E_OFF_A + stage * BLK_BYTES;
#ifdef GNB_ABL_LN_NONE
      if (a.R < 0)
#endif
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += 8) {
        float4 xc[2][4];
        float s[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
          for (int j = 0; j < 4; j++) xc[h][j] = xa[h][j];
          const float4 t = add4(add4(xc[h][0], xc[h][1]), add4(xc[h][2], xc[h][3]));
          s[h] = (t.x + t.y) + (t.z + t.w);
        }
        if (i0 + 8 < 32) issue(i0 + 8);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
          for (int h = 0; h < 2; h++) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float nmu = -s[h] * (1.0f / H);
          float2 t2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            xc[h][j] = adds4(xc[h][j], nmu);
            t2 = __ffma2_rn(make_float2(xc[h][j].x, xc[h][j].y), make_float2(xc[h][j].x, xc[h][j].y), t2);
            t2 = __ffma2_rn(make_float2(xc[h][j].z, xc[h][j].w), make_float2(xc[h][j].z, xc[h][j].w), t2);
          }
          s[h] = t2.x + t2.y;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
          for (int h = 0; h < 2; h++) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int i = i0 + 4 * h + rr;
          const float rs = (i < rows) ? rstd_nb(s[h] * (1.0f / H), e_add, e_plus) : 0.f;
          const uint32_t rowoff = a_lane + (uint32_t)i * 128;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float4 xh = muls4(xc[h][j], rs);
            uint2 pk;
            pk.x = pack_bf16(xh.x, xh.y);
            pk.y = pack_bf16(xh.z, xh.w);
            *reinterpret_cast<uint2*>(A + (j >> 1) * KB_BYTES + rowoff + (((a_chunk + 4 * (j & 1)) ^ (uint32_t)(i & 7)) << 4)) = pk;
          }
        }
        EDBG(2 + (i0 >> 3));
      }
      // ---- ordered partial sums per (32-row block, receiver) run, taken over the bf16 operand the MMAs consume
      __syncwarp();
#ifdef GNB_ABL_LN_NOPART
      if (a.R < 0)
#endif
      {
        float4 acc = f4zero();
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 16) {
          uint2 v[16];     // 16 independent shared-memory loads in flight, then the ordered sums
#pragma unroll
          for (int u = 0; u < 16; u++) {
            const int i = i0 + u;
            v[u] = *reinterpret_cast<const uint2*>(A + p_lane + i * 128 + ((p_chunk ^ (uint32_t)(i & 7)) << 4));
          }
#pragma unroll
          for (int u = 0; u < 16; u++) {
            const int i = i0 + u;
            acc = add4(acc, make_float4(__uint_as_float(v[u].x << 16), __uint_as_float(v[u].x & 0xffff0000u),
                                        __uint_as_float(v[u].y << 16), __uint_as_float(v[u].y & 0xffff0000u)));
            const bool fl = (endmask >> i) & 1u;
            if (fl) {
              if (a.part_bf16) *(reinterpret_cast<uint2*>(a.Epart) + (size_t)pid * (H / 4) + lane) = make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
              else *(reinterpret_cast<float4*>(a.Epart + (size_t)pid * H) + lane) = acc;
            }
            acc.x = fl ? 0.f : acc.x; acc.y = fl ? 0.f : acc.y; acc.z = fl ? 0.f : acc.z; acc.w = fl ? 0.f : acc.w;
            pid += fl ? 1 : 0;
          }
        }
      }
      fence_async_smem();
      ARRIVE_LEADER(EB_AFULL + stage);
      EDBG(6);
    }
  } else {
    // ===================================================== OUT warps (4-7, 12, 13), one tile behind the MMAs.  A tile is drained in
    // 8 slices of 16 rows (= one cut of the partial-row index); slice k = 8 pass + s is taken by OUT warp k mod 6.
    // The role is bound by the latency of its three global loads per row (x for the residual, gathered P_s[src], P_r'[dst]), so
    // the loads of DEPTH rows are kept in flight in a rolling window that runs across slice boundaries: the indices of the next
    // slice are fetched while the current one is processed, and its first rows are issued from the tail of the current slice.
#ifndef GNB_OUT_DEPTH
#define GNB_OUT_DEPTH 4
#endif
    constexpr int DEPTH = GNB_OUT_DEPTH;
    static_assert(OUT_ROWS % DEPTH == 0, "the rolling window keeps its slot numbering across slices only if DEPTH divides the slice");
    const int ow = warp < 8 ? warp - 4 : warp - 8;
    const uint64_t pol_xf = l2_policy_evict_first();
    const float* xbase = a.x + 4 * lane;
    // this lane's 4 columns of a gathered row: 16 B of an fp32 row, 8 B of a bf16 row
    const uint8_t* base1 = reinterpret_cast<const uint8_t*>(a.add1) + (PBF ? 8 : 16) * lane;
    const uint8_t* base2 = reinterpret_cast<const uint8_t*>(a.add2) + (PBF ? 8 : 16) * lane;
    const size_t ldb1 = (size_t)a.ld1 * (PBF ? 2 : 4), ldb2 = (size_t)a.ld2 * (PBF ? 2 : 4);
    const uint32_t s_lane0 = (uint32_t)(E_OFF_STG + (lane >> 3) * 16384);
    const uint32_t s_chunk = (uint32_t)(lane & 7);
    const int total = OUT_SLICES * npass;
    // per-lane row attributes of a slice (lane i < 16 holds row i): gather indices and partial-row id
    auto slice_row0 = [&](int k) { return ((int64_t)blockIdx.x + (int64_t)(k >> 3) * grid) * TM + OUT_ROWS * (k & (OUT_SLICES - 1)); };
    auto load_attr = [&](int64_t row0, int& i1, int& i2, int& pid) {
      i1 = 0; i2 = 0; pid = -1;
      if (lane < OUT_ROWS && row0 + lane < a.R) {
        i1 = a.idx1 ? __ldg(a.idx1 + row0 + lane) : (int)(row0 + lane);
        i2 = __ldg(a.idx2 + row0 + lane);
        pid = __ldg(a.part + row0 + lane);
      }
    };
    // (a fully unrolled 16-row rolling window was measured slower than this compact 4-row-group loop: 0.96 vs 0.86 ms per launch,
    // instruction-cache misses)
    // Four rows of loads are in flight per warp.  The window runs ACROSS slices: the gather indices of this warp's next slice are
    // fetched at the start of the current one and its first group is issued from the last group of the current one, so a slice
    // starts with its loads already in flight (without this every slice paid two full load latencies before its first row).
    float4 wdec[4];      // DEC: rows 4 lane .. 4 lane + 3 of decW (4 outputs each)
    if (DEC) {
#pragma unroll
      for (int c = 0; c < 4; c++) wdec[c] = __ldg(reinterpret_cast<const float4*>(a.decW) + 4 * lane + c);
    }
    // NG groups of four rows of loads in flight per warp: two where the registers allow it (bf16 gathered rows, no fused decoder)
#ifndef GNB_OUT_NG
#define GNB_OUT_NG 1      /* 2 (eight rows in flight) measured 2 % slower: the role is not bound by its loads in flight (profiles/r02_summary.md) */
#endif
    constexpr int NG = (PBF && !DEC) ? GNB_OUT_NG : 1;
    constexpr int WIN = 4 * NG;
    static_assert(OUT_ROWS % WIN == 0, "the window keeps its slot numbering across slices only if it divides the slice");
    float4 xa[WIN];
    typename std::conditional<PBF, uint2, float4>::type pa[WIN], pb[WIN];
    auto issue1 = [&](int u, int64_t rb, int src_i1, int src_i2, int i) {
      const int i1 = __shfl_sync(0xffffffffu, src_i1, i), i2 = __shfl_sync(0xffffffffu, src_i2, i);
      int64_t r = rb + i;
      r = r < a.R ? r : a.R - 1;
#ifdef GNB_ABL_OUT_NOX      /* GNB_ABL_*: timing-only ablations (results wrong), tools/ab_edge.sh */
      xa[u] = f4zero();
#else
      xa[u] = GNB_X_L2_HINTS ? ld_stream_hint(xbase + (size_t)r * H, pol_xf) : ld_stream(xbase + (size_t)r * H);      // second and last read of x
#endif
#ifdef GNB_ABL_OUT_NOGATHER
      if constexpr (PBF) { pa[u] = make_uint2(0u, 0u); pb[u] = pa[u]; } else { pa[u] = f4zero(); pb[u] = pa[u]; }
      return;
#endif
      if constexpr (PBF) {
        pa[u] = __ldg(reinterpret_cast<const uint2*>(base1 + (size_t)i1 * ldb1));
        pb[u] = __ldg(reinterpret_cast<const uint2*>(base2 + (size_t)i2 * ldb2));
      } else {
        pa[u] = __ldg(reinterpret_cast<const float4*>(base1 + (size_t)i1 * ldb1));
        pb[u] = __ldg(reinterpret_cast<const float4*>(base2 + (size_t)i2 * ldb2));
      }
    };
    auto gathered = [&](int u) {      // g = add1[idx1] + add2[idx2] of window slot u, fp32
      if constexpr (PBF) return add4(bf4(pa[u]), bf4(pb[u]));
      else return add4(pa[u], pb[u]);
    };
    uint64_t pol_ef = 0;      // L2 evict-first policy of the TMA stores (y is not read again by this kernel)
    if (!DEC) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_ef));
    int c_i1 = 0, c_i2 = 0, c_pid = -1, n_i1 = 0, n_i2 = 0, n_pid = -1;
    if (ow < total) {
      load_attr(slice_row0(ow), c_i1, c_i2, c_pid);
#pragma unroll
      for (int u = 0; u < WIN; u++) issue1(u, slice_row0(ow), c_i1, c_i2, u);
    }
    for (int k = ow; k < total && !wd_dead; k += OUT_WARPS) {
      const uint32_t tl = (uint32_t)k >> 3;
      const int sl = k & (OUT_SLICES - 1);
      const uint32_t s_lane = s_lane0 + (uint32_t)((OUT_ROWS * sl) * 128);
      const int64_t row0 = slice_row0(k);
      const int64_t left = a.R - row0;
      const int rows = left < 0 ? 0 : (left > OUT_ROWS ? OUT_ROWS : (int)left);
      const bool has_next = k + OUT_WARPS < total;      // warp-uniform
      const int64_t nrow0 = slice_row0(has_next ? k + OUT_WARPS : k);
      if (has_next) load_attr(nrow0, n_i1, n_i2, n_pid);
      const int nxt = __shfl_down_sync(0xffffffffu, c_pid, 1);
      const uint32_t endmask = __ballot_sync(0xffffffffu, lane < rows && (lane == rows - 1 || nxt != c_pid));
      int pid = __shfl_sync(0xffffffffu, c_pid, 0);
      EDBG(0);
      mbar_wait(BAR(EB_STGFULL + sl), tl & 1);
      EDBG(1);
      float4 acc = f4zero();
#ifdef GNB_ABL_OUT_NONE
      if (a.R < 0)
#endif
#pragma unroll 1
      for (int j0 = 0; j0 < OUT_ROWS; j0 += WIN) {
#pragma unroll
      for (int hg = 0; hg < NG; hg++) {      // group hg of the window: slots 4 hg .. 4 hg + 3
        const int i0 = j0 + 4 * hg;
        float4 d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u;
          d[u] = *reinterpret_cast<const float4*>(sm + s_lane + i * 128 + ((s_chunk ^ (uint32_t)(i & 7)) << 4));
        }
        // where the refills of this group come from: WIN rows further down this slice, or the head of this warp's next slice
        const bool cur = i0 + WIN < OUT_ROWS;
        const int64_t rb = cur ? row0 : nrow0;
        const int s1 = cur ? c_i1 : n_i1, s2 = cur ? c_i2 : n_i2;
        const int ib = cur ? i0 + WIN : i0 + WIN - OUT_ROWS;
        const bool refill = cur || has_next;      // warp-uniform
        float od[16];      // DEC: this lane's share of the 4 decoder outputs of the 4 rows of the group
#pragma unroll
        for (int uu = 0; uu < 4; uu++) {
          const int u = 4 * hg + uu;      // window slot
          const int i = i0 + uu;
          const float4 g = gathered(u);
          const float4 y = add4(add4(xa[u], g), d[uu]);      // (b2 is in the accumulator)
          if (DEC) {
            od[4 * uu + 0] = fmaf(y.w, wdec[3].x, fmaf(y.z, wdec[2].x, fmaf(y.y, wdec[1].x, y.x * wdec[0].x)));
            od[4 * uu + 1] = fmaf(y.w, wdec[3].y, fmaf(y.z, wdec[2].y, fmaf(y.y, wdec[1].y, y.x * wdec[0].y)));
            od[4 * uu + 2] = fmaf(y.w, wdec[3].z, fmaf(y.z, wdec[2].z, fmaf(y.y, wdec[1].z, y.x * wdec[0].z)));
            od[4 * uu + 3] = fmaf(y.w, wdec[3].w, fmaf(y.z, wdec[2].w, fmaf(y.y, wdec[1].w, y.x * wdec[0].w)));
          } else if (i < rows) {
#if !defined(GNB_ABL_OUT_NOSTORE) && !defined(GNB_ABL_OUT_NOY)
            if (a.y_tma) *reinterpret_cast<float4*>(sm + s_lane + i * 128 + ((s_chunk ^ (uint32_t)(i & 7)) << 4)) = y;      // in place of d
            else __stcs(reinterpret_cast<float4*>(a.y + (size_t)(row0 + i) * H) + lane, y);
#else
            if (y.x == 123.456f) a.y[0] = y.y;
#endif
          }
          acc = add4(acc, g);
          const bool fl = (endmask >> i) & 1u;
#if !defined(GNB_ABL_OUT_NOSTORE) && !defined(GNB_ABL_OUT_NOGP)
          if (fl) {
            if (a.part_bf16) *(reinterpret_cast<uint2*>(a.Gpart) + (size_t)pid * (H / 4) + lane) = make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
            else *(reinterpret_cast<float4*>(a.Gpart + (size_t)pid * H) + lane) = acc;
          }
#elif defined(GNB_ABL_OUT_NOGP)
          if (fl && acc.x == 123.456f) a.Gpart[0] = acc.y;
#endif
          acc.x = fl ? 0.f : acc.x; acc.y = fl ? 0.f : acc.y; acc.z = fl ? 0.f : acc.z; acc.w = fl ? 0.f : acc.w;
          pid += fl ? 1 : 0;
          if (refill) issue1(u, rb, s1, s2, ib + uu);
        }
        if (DEC) {
          // 16 totals (4 rows x 4 outputs) land on the even lanes: value 8 b4 + 4 b3 + 2 b2 + b1 = 4 u + j, i.e. 16 consecutive floats
          const float t = warp_sum16(od, lane);
          const int vi = lane >> 1;
          if ((lane & 1) == 0 && i0 + (vi >> 2) < rows) a.dec_out[(size_t)(row0 + i0) * 4 + vi] = t;
        }
      }
      }
      if (!DEC && a.y_tma) {
        // the slice now holds y in the staging layout = four 16 x 128 B boxes in the 128B-swizzle pattern of the tensor map
        fence_async_smem();      // this thread's generic-proxy writes -> visible to the TMA engine (async proxy)
        __syncwarp();
        if (lane == 0 && rows > 0) {
          const uint32_t s0 = base + (uint32_t)(E_OFF_STG + (OUT_ROWS * sl) * 128);
#pragma unroll
          for (int g = 0; g < 4; g++)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                         ::"l"(reinterpret_cast<uint64_t>(&a.ymap)), "r"(32 * g), "r"((int)row0), "r"(s0 + (uint32_t)g * 16384u), "l"(pol_ef)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the slice may be refilled once the engine has read it
        }
        __syncwarp();
      }
      ARRIVE_LOCAL(EB_STGEMPTY + sl);
      EDBG(2);
      c_i1 = n_i1; c_i2 = n_i2; c_pid = n_pid;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();      // the peer may still be reading this CTA's shared / tensor memory
  if (warp == W_MMA) {
    if (CL2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
#undef TILE_OK
#undef mbar_wait
#undef mbar_wait_cluster
}

}  // namespace

namespace {
template <bool CL2, bool DEC, bool PBF>
int launch_edge5_t(gnb_ctx* ctx, const EdgeArgs& a, int grid) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(E_THREADS); cfg.dynamicSmemBytes = e_smem(CL2); cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  GNB_CUDA(cudaLaunchKernelEx(&cfg, k_edge5<CL2, DEC, PBF>, a));
  return GNB_OK;
}
template <bool CL2, bool DEC, bool PBF>
int edge5_smem_attr() {
  GNB_CUDA(cudaFuncSetAttribute(k_edge5<CL2, DEC, PBF>, cudaFuncAttributeMaxDynamicSharedMemorySize, e_smem(CL2)));
  return GNB_OK;
}
}  // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}

int launch_edge5(gnb_ctx* ctx, const EdgeArgs& a_in, const char* name, double flops, double bytes) {
  if (a_in.num_tiles <= 0) return GNB_OK;
  EdgeArgs a = a_in;
  a.y_tma = 0;
  static const bool tma_env = getenv("GNB_EDGE_NO_TMA_STORE") == nullptr;      // A/B toggle
  if (a.decW == nullptr && a.y != nullptr && tma_env && encode_tiled_fn()) {
    // y [R][128] fp32: box = 32 columns x 16 rows (one column group of one OUT slice), 128 B swizzle == the staging layout
    const cuuint64_t dims[2] = {128, (cuuint64_t)a.R};
    const cuuint64_t strides[1] = {512};
    const cuuint32_t box[2] = {32, (cuuint32_t)OUT_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled_fn()(&a.ymap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.y, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    a.y_tma = r == CUDA_SUCCESS ? 1 : 0;
  }
  // GNB_EDGE_CTA_PAIR=0: one CTA per tile stream (cta_group::1).  Read per launch so that a test can run both instantiations.
  const char* pair_env = getenv("GNB_EDGE_CTA_PAIR");
  const bool cl2 = (pair_env ? atoi(pair_env) : 1) && ctx->sm_count >= 2;
  const bool dec = a.decW != nullptr, pbf = a.add_bf16 != 0;
  int grid;
  if (cl2) {
    static const int cap_env = getenv("GNB_EDGE_MAX_PAIRS") ? atoi(getenv("GNB_EDGE_MAX_PAIRS")) : 0;      // experiment: fewer SMs
    const int pairs = (a.num_tiles + 1) / 2, max_clusters = (cap_env > 0 && cap_env < ctx->sm_count / 2) ? cap_env : ctx->sm_count / 2;
    grid = 2 * (pairs < max_clusters ? pairs : max_clusters);
  } else {
    grid = a.num_tiles < ctx->sm_count ? a.num_tiles : ctx->sm_count;
  }
  if (ctx_first(ctx, ONCE_EDGE5)) {
    GNB_TRY((edge5_smem_attr<false, false, false>())); GNB_TRY((edge5_smem_attr<false, false, true>()));
    GNB_TRY((edge5_smem_attr<false, true, false>())); GNB_TRY((edge5_smem_attr<false, true, true>()));
    GNB_TRY((edge5_smem_attr<true, false, false>())); GNB_TRY((edge5_smem_attr<true, false, true>()));
    GNB_TRY((edge5_smem_attr<true, true, false>())); GNB_TRY((edge5_smem_attr<true, true, true>()));
  }
  Launch L(ctx, name, bytes, flops);
  const int sel = (cl2 ? 4 : 0) | (dec ? 2 : 0) | (pbf ? 1 : 0);
  switch (sel) {
    case 0: GNB_TRY((launch_edge5_t<false, false, false>(ctx, a, grid))); break;
    case 1: GNB_TRY((launch_edge5_t<false, false, true>(ctx, a, grid))); break;
    case 2: GNB_TRY((launch_edge5_t<false, true, false>(ctx, a, grid))); break;
    case 3: GNB_TRY((launch_edge5_t<false, true, true>(ctx, a, grid))); break;
    case 4: GNB_TRY((launch_edge5_t<true, false, false>(ctx, a, grid))); break;
    case 5: GNB_TRY((launch_edge5_t<true, false, true>(ctx, a, grid))); break;
    case 6: GNB_TRY((launch_edge5_t<true, true, false>(ctx, a, grid))); break;
    default: GNB_TRY((launch_edge5_t<true, true, true>(ctx, a, grid))); break;
  }
  GNB_CUDA(cudaGetLastError());
  return GNB_OK;
}
