// PTX helpers shared by the tcgen05 kernels (tc.cu, tc_edge.cu): mbarrier, bulk async copies (TMA engine),
// tcgen05 MMA / TMEM load-store wrappers, UMMA descriptors, the 128B-swizzled K-major operand layout.
#pragma once
#include <cuda_bf16.h>
#include "kernels.cuh"

namespace tcx {

constexpr int H = 128;             // feature width handled by this path
constexpr int TM = 128;            // rows per tile (UMMA M)
constexpr int BLK_BYTES = 32768;   // one 128x128 bf16 operand block
constexpr int KB_BYTES = 16384;    // one 64-wide K half of a block: 128 rows x 128 B

enum { MODE_EDGE = 0, MODE_NODE = 1 };

// ------------------------------------------------------------------ PTX helpers
static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Kernel watchdog.  Every mbarrier wait is a bounded wait: if a wait does not complete within `limit_ns` (a protocol
// dead-lock), the waiting warp raises the device-global flag, marks itself dead and stops waiting; every other warp of
// every CTA sees the flag at its next slow-path check and does the same, so the kernel drains and EXITS instead of
// hanging or trapping (a trap leaves a sticky error that kills the CUDA context).  The host reads the flag at its
// next synchronisation point and returns GNB_ERR_TIMEOUT; the results of that forward are invalid.
// All waits are executed by whole, converged warps; the votes keep the outcome warp-uniform.
// (WatchArgs, the kernel-argument half, lives in common.cuh)
// Per-thread state is one predicate (`dead`); the flag pointer and the limit stay in the kernel-parameter constant bank and
// are only read on the slow path, so the watchdog costs the hot loops no registers.
static __device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// try_wait with a suspend-time hint: the thread may sleep in hardware until the phase completes (or the hint expires) instead of
// returning to the polling loop (ncu: 20 % of all instructions of k_edge5 were barrier polls).  OFF by default (0): measured, it
// changes nothing for k_edge5, and a waiter can over-sleep - harmless for the product's protocols, but it made the round-1
// k_tc_proj protocol (test-only build, tests/test_gpu_watchdog.py) alias its 1-bit phase without any artificial stall.
#ifndef GNB_MBAR_HINT_NS
#define GNB_MBAR_HINT_NS 0
#endif
static __device__ __forceinline__ uint32_t mbar_test_hint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"((uint32_t)GNB_MBAR_HINT_NS)
      : "memory");
  return ok;
}
static __device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// slow path; returns true when the watchdog fired
static __device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, int* flag, unsigned long long limit_ns) {      // one copy per kernel (instruction cache)
  const unsigned long long t0 = gtimer_ns();
  for (uint32_t spin = 1;; spin++) {
    if (__all_sync(0xffffffffu, GNB_MBAR_HINT_NS > 0 ? mbar_test_hint(bar, parity) : mbar_test(bar, parity))) return false;
    if ((spin & 63u) == 0u) {
      int dead = 0;
      if ((threadIdx.x & 31) == 0) {
        dead = *reinterpret_cast<volatile int*>(flag);
        if (!dead && gtimer_ns() - t0 > limit_ns) {
          atomicExch(flag, 1);
          dead = 1;
        }
      }
      if (__shfl_sync(0xffffffffu, dead, 0)) return true;
    }
  }
}
static __device__ __forceinline__ void mbar_wait_w(uint32_t bar, uint32_t parity, bool& dead, const WatchArgs& wa) {
  if (dead) return;
  if (__all_sync(0xffffffffu, mbar_test(bar, parity))) return;
  dead = mbar_wait_slow(bar, parity, wa.flag, wa.limit_ns);
}
static __device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// one instruction pulls `bytes` (multiple of 16) of global memory into L2 (no registers, no shared memory)
static __device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// ---- cache-policy loads (L1 no-allocate for read-once rows, L2 eviction priority for rows a kernel reads twice)
static __device__ __forceinline__ float4 ld_stream(const float* p) {   // read-once data: do not keep it in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
static __device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
static __device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
static __device__ __forceinline__ float4 ld_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
static __device__ __forceinline__ float2 ld_hint2(const float2* p, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
static __device__ __forceinline__ float4 ld_stream_hint(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}

static __device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// One lane of a CONVERGED warp.  tcgen05.mma / commit take their operands from uniform registers; issuing
// them from divergent code (if (lane == 0)) makes ptxas wrap every instruction in an ELECT/branch loop,
// which costs ~2x the tensor-pipe time of a 128x128x16 UMMA (tools/micro/hwprobe.cu, T3).
static __device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem] . B[smem]
static __device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem, bf16 pairs packed per 32-bit column] . B[smem]
static __device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
#define TC_LD32(taddr, r)                                                                                             \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                       \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"   \
      "%29,%30,%31}, [%32];"                                                                                          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                     \
      : "r"(taddr)                                                                                                    \
      : "memory")
// 16 lanes x 64 fp32 columns in the accumulator-fragment layout (verified by tools/micro/hwprobe.cu, T1):
//   r[4n + 2h + c] of lane l = TMEM[lane base + l/4 + 8h][col base + 8n + 2(l%4) + c]
// four consecutive lanes cover one 32-byte sector of a row: sector-exact global loads / stores from registers.
#define TC_LD_FRAG64(taddr, r)                                                                                        \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "                                                                       \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"   \
      "%29,%30,%31}, [%32];"                                                                                          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                     \
      : "r"(taddr)                                                                                                    \
      : "memory")
// 32 lanes x 16 columns store (thread i writes row lane base + i)
#define TC_ST16(taddr, r)                                                                                             \
  asm volatile(                                                                                                       \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::"r"( \
          r[0]),                                                                                                      \
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),  \
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)                                          \
      : "memory")
static __device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
static __device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 | LBO (unused for swizzled K-major, 1) | SBO = 1024 B between 8-row groups |
// version 1 (Blackwell) | layout type 2 (SWIZZLE_128B)
static __device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// byte offset of element (row r, k) inside a 128-row K-major SW128 operand block (K <= 128)
static __device__ __forceinline__ uint32_t sw_off(int r, int k) {
  return (uint32_t)((k >> 6) * KB_BYTES + r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
}

static __device__ __forceinline__ float ln_rstd(float var, float eps, int mode) {
  if (mode == GNB_EPS_SQRT_VAR_EPS2) return rsqrtf(var + eps * eps);
  if (mode == GNB_EPS_STD_PLUS_EPS) return 1.0f / (sqrtf(var) + eps);
  return rsqrtf(var + eps);
}
static __device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// relu fused into the conversion: {lo, hi} = bf16(max(lo,0)), bf16(max(hi,0))
static __device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
static __device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
static __device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static __device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// one 128x128x128 block = 8 UMMAs of K=16; A block = two 64-wide K halves 16 KB apart, B = two ring stages.
// Call from ONE elected lane of a converged warp.
static __device__ __forceinline__ void issue_ss(uint32_t d_tmem, uint64_t adesc, uint64_t w0, uint64_t w1, bool accumulate) {
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ss(d_tmem, adesc + 2 * ks, w0 + 2 * ks, IDESC, (accumulate || ks > 0) ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ss(d_tmem, adesc + (KB_BYTES >> 4) + 2 * ks, w1 + 2 * ks, IDESC, 1u);
}
// A operand in TMEM: bf16 pairs, 8 columns per K=16 step
static __device__ __forceinline__ void issue_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t w0, uint64_t w1, bool accumulate) {
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ts(d_tmem, a_tmem + 8 * ks, w0 + 2 * ks, IDESC, (accumulate || ks > 0) ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ts(d_tmem, a_tmem + 32 + 8 * ks, w1 + 2 * ks, IDESC, 1u);
}


// ------------------------------------------------------------------ 2-CTA (cta_group::2) helpers
static __device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
static __device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
static __device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
static __device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default .release.cta like CUTLASS' ClusterBarrier::arrive: what is published across the pair is tensor memory (ordered by
  // tcgen05.fence) and shared memory consumed through the async proxy (ordered by fence.proxy.async); a cluster-scope release
  // would add a full memory barrier (~1.2k cycles per arrive, measured)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// arrive::one on the barrier at this offset in both CTAs of the pair once all prior tcgen05 ops of this thread completed
static __device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
static __device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
static __device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// kind::f16 instruction descriptor of the pair MMA: M = 256 (128 rows per CTA), N = 128
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
// one 256x128x128 block of the CTA pair: A = each CTA's own 128-row tile (two 64-wide K halves 16 KB apart), B = each CTA's
// 64-row N half (two K halves 8 KB apart inside a 16 KB ring stage)
static __device__ __forceinline__ void issue_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t w, bool accumulate) {
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ss2(d_tmem, adesc + 2 * ks, w + 2 * ks, IDESC2, (accumulate || ks > 0) ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ss2(d_tmem, adesc + (KB_BYTES >> 4) + 2 * ks, w + (8192 >> 4) + 2 * ks, IDESC2, 1u);
}
static __device__ __forceinline__ void issue_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t w, bool accumulate) {
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ts2(d_tmem, a_tmem + 8 * ks, w + 2 * ks, IDESC2, (accumulate || ks > 0) ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < 4; ks++) mma_ts2(d_tmem, a_tmem + 32 + 8 * ks, w + (8192 >> 4) + 2 * ks, IDESC2, 1u);
}

}  // namespace tcx
