// Latency of tcgen05.ld / tcgen05.st issued by an epilogue warp WHILE the tensor pipe runs a queue of UMMAs (sm_100a):
//   T6  ld(32x32b.x32)+wait latency: idle pipe / unthrottled MMA queue (64 blocks enqueued back to back) / throttled queue
//       (the issuer waits for the commit of block b-1 before issuing block b+1: at most 2 blocks queued)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_lat tools/micro/tmem_lat.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__host__ __device__ inline uint32_t sw_off(int r, int k, int rows) {   // K-major SW128 block with `rows` rows, K <= 128
  return (uint32_t)((k >> 6) * (rows * 128) + r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
}
#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define LD256x8(taddr, r) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define ST32(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" \
  :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr) : "memory")
#define ST16(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" \
  :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr) : "memory")
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) { uint32_t d; asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }


// mode 0: no MMAs;  1: unthrottled queue;  2: at most `depth` blocks in flight
template <bool TS>
__global__ void __launch_bounds__(192, 1) k_lat(int mode, int depth, int nblocks, int nsamp, long long* out /*[grid][4 warps][nsamp][2]*/, long long* out_total) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 131072);      // bars[0..7]: block-complete ring
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 16);
  volatile int* flags = reinterpret_cast<volatile int*>(bars + 18);      // [0] started, [1] stop
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (tid == 0) { for (int i = 0; i < 8; i++) mbar_init(smem_u32(bars + i), 1); flags[0] = 0; flags[1] = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 4) {
    const uint32_t sA = base, sB = base + 32768;
    constexpr uint32_t idesc = make_idesc(128, 128);
    const uint64_t dA = umma_desc(sA), dB = umma_desc(sB);
    if (lane == 0) flags[0] = 1;
    long long t0 = clock64();
    if (mode != 0) {
      for (int b = 0; b < nblocks; b++) {
        if (mode == 2 && b >= depth) mbar_wait(smem_u32(bars + ((b - depth) & 7)), ((b - depth) >> 3) & 1);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            const uint32_t off = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
            if (TS) mma_ts(tmem + 256, tmem + 384 + 8 * ks, dB + (off >> 4), idesc, 1);
            else mma_ss(tmem + 256, dA + (off >> 4), dB + (off >> 4), idesc, 1);
          }
          tc_commit(smem_u32(bars + (b & 7)));
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(bars + ((nblocks - 1) & 7)), ((nblocks - 1) >> 3) & 1);
    } else {
      while (!flags[1]) { }
    }
    long long t1 = clock64();
    if (lane == 0) { flags[1] = 1; out_total[blockIdx.x] = t1 - t0; }
  } else if (warp < 4) {
    while (!flags[0]) { }
    uint32_t v[32];
    long long* o = out + ((size_t)(blockIdx.x * 4 + warp) * nsamp) * 2;
    uint32_t keep = 0;
    for (int s = 0; s < nsamp; s++) {
      if (mode != 0 && flags[1]) { if (lane == 0) { o[2 * s] = -1; o[2 * s + 1] = -1; } continue; }
      long long a0 = clock64();
      LD32(tmem + ((uint32_t)(warp * 32) << 16) + 0, v);        // columns 0..31: not touched by the MMAs (D = 256.., A(TS) = 384..)
      ld_wait();
      long long a1 = clock64();
      keep += v[0] + v[31];
      ST32(tmem + ((uint32_t)(warp * 32) << 16) + 32, v);
      st_wait();
      long long a2 = clock64();
      if (lane == 0) { o[2 * s] = a1 - a0; o[2 * s + 1] = a2 - a1; }
      __nanosleep(100);
    }
    if (keep == 0x12345678u) out_total[0] = keep;
    if (mode == 0 && warp == 0 && lane == 0) flags[1] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <bool TS>
void run(const char* name, int mode, int depth, long long* dO, long long* dT) {
  const int grid = 148, nb = 64, ns = 48;
  CK(cudaFuncSetAttribute(k_lat<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 133000));
  CK(cudaMemset(dO, 0xff, (size_t)grid * 4 * ns * 2 * 8));
  k_lat<TS><<<grid, 192, 133000>>>(mode, depth, nb, ns, dO, dT);
  CK(cudaDeviceSynchronize());
  std::vector<long long> o((size_t)grid * 4 * ns * 2), t(grid);
  CK(cudaMemcpy(o.data(), dO, o.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(t.data(), dT, grid * 8, cudaMemcpyDeviceToHost));
  double sl = 0, ss = 0; long long ml = 0, ms = 0, nl = 1ll << 60, n = 0;
  for (size_t i = 0; i < o.size(); i += 2) {
    if (o[i] < 0) continue;
    sl += o[i]; ss += o[i + 1]; ml = o[i] > ml ? o[i] : ml; ms = o[i + 1] > ms ? o[i + 1] : ms; nl = o[i] < nl ? o[i] : nl; n++;
  }
  long long mx = 0; for (auto v : t) mx = v > mx ? v : mx;
  printf("T6: %-44s: ld+wait min %lld avg %.0f max %lld | st+wait avg %.0f max %lld | %lld samples | MMA stream %lld cycles (%.0f per block)\n",
         name, nl, sl / n, ml, ss / n, ms, n, mx, (double)mx / nb);
}

int main() {
  long long *dO, *dT;
  CK(cudaMalloc(&dO, 148 * 4 * 64 * 2 * 8)); CK(cudaMalloc(&dT, 148 * 8));
  for (int rep = 0; rep < 2; rep++) {
    run<false>("idle tensor pipe", 0, 0, dO, dT);
    run<false>("SS queue unthrottled (64 blocks)", 1, 0, dO, dT);
    run<false>("SS queue, <= 1 block in flight", 2, 1, dO, dT);
    run<false>("SS queue, <= 2 blocks in flight", 2, 2, dO, dT);
    run<false>("SS queue, <= 3 blocks in flight", 2, 3, dO, dT);
    run<true>("TS queue unthrottled (64 blocks)", 1, 0, dO, dT);
    run<true>("TS queue, <= 2 blocks in flight", 2, 2, dO, dT);
  }
  return 0;
}
