// Hardware probe for the design of the fused GNCore kernel (sm_100a):
//   T1  register layout of tcgen05.ld.16x256b
//   T2  tcgen05.mma with the A operand in TMEM (bf16 packed, written in place over fp32 columns)
//   T3  tensor-pipe time per 128x128x128 block: SS vs TS, N=128 vs N=256, with shared-memory contention
//   T4  tcgen05.ld throughput (32x32b.x32 and 16x256b.x8) with 4 / 8 warps
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hwprobe tools/micro/hwprobe.cu
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

// ------------------------------------------------------------------------------------------------ T1 + T2
// out_layout[tid*16 + i]  : 16x256b.x2 registers of warp 0 / second half etc.
// out_d[m*128 + n]        : D = relu(Hf32) (bf16, in place) . W^T    (TS MMA)
__global__ void __launch_bounds__(160, 1) k_probe12(const __nv_bfloat16* __restrict__ Wg /*[128 n][128 k]*/, float* out_layout, float* out_d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(smem_u32(bars), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B operand: W[n][k] -> K-major SW128 image
  for (int i = tid; i < 128 * 128; i += blockDim.x) {
    int n = i >> 7, k = i & 127;
    *reinterpret_cast<__nv_bfloat16*>(sm + sw_off(n, k, 128)) = Wg[i];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp < 4) {
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const int row = warp * 32 + lane;
    // fp32 "hidden" values H[row][c] = ((row*3 + c*5) % 11) - 4   in columns [0,128)
    for (int j = 0; j < 4; j++) {
      uint32_t r[32];
      for (int i = 0; i < 32; i++) r[i] = __float_as_uint((float)(((row * 3 + (32 * j + i) * 5) % 11) - 4));
      ST32(tmem + lane_base + 32 * j, r);
    }
    st_wait();
    // T1: read back columns [0,16) of lanes 0..15 and 16..31 of this warp with 16x256b.x2
    if (warp == 0) {
      uint32_t a[8], b[8];
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "r"(tmem + 0) : "memory");
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(tmem + (16u << 16)) : "memory");
      ld_wait();
      for (int i = 0; i < 8; i++) { out_layout[lane * 16 + i] = __uint_as_float(a[i]); out_layout[lane * 16 + 8 + i] = __uint_as_float(b[i]); }
    }
    __syncwarp();
    // in place: relu -> bf16 pairs -> columns [16j, 16j+16)
    for (int j = 0; j < 4; j++) {
      uint32_t r[32], p[16];
      LD32(tmem + lane_base + 32 * j, r);
      ld_wait();
      for (int i = 0; i < 16; i++) p[i] = pack_bf16_relu(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      ST16(tmem + lane_base + 16 * j, p);
    }
    st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4 && lane == 0) {
    tc_fence_after();
    for (int ks = 0; ks < 8; ks++) {
      uint32_t off = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
      mma_ts(tmem + 128, tmem + 8 * ks, umma_desc(base + off), make_idesc(128, 128), ks > 0);
    }
    tc_commit(smem_u32(bars));
  }
  if (warp < 4) {
    mbar_wait(smem_u32(bars), 0);
    tc_fence_after();
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const int row = warp * 32 + lane;
    for (int j = 0; j < 4; j++) {
      uint32_t r[32];
      LD32(tmem + 128 + lane_base + 32 * j, r);
      ld_wait();
      for (int i = 0; i < 32; i++) out_d[row * 128 + 32 * j + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------ T3
// compile-time variants so the issue loop is nothing but UTCHMMA with precomputed uniform descriptors
template <bool TS, int N, bool HAMMER>
__global__ void __launch_bounds__(192, 1) k_probe3(int nblocks, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (tid == 0) { mbar_init(smem_u32(bars), 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
    constexpr uint32_t idesc = make_idesc(128, N);
    constexpr uint32_t khalf = (uint32_t)N * 128;
    const uint64_t dA = umma_desc(sA), dB = umma_desc(sB);
    long long t0 = clock64();
    for (int b = 0; b < nblocks; b++) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          constexpr int dummy = 0; (void)dummy;
          const uint32_t offA = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
          const uint32_t offB = (uint32_t)((ks >> 2) * khalf + (ks & 3) * 32);
          if (TS) mma_ts(tmem + 256, tmem + 8 * ks, dB + (offB >> 4), idesc, 1);
          else mma_ss(tmem + 256, dA + (offA >> 4), dB + (offB >> 4), idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(smem_u32(bars));
    __syncwarp();
    mbar_wait(smem_u32(bars), 0);
    long long t1 = clock64();
    if (lane == 0) { *stop = 1; out_cycles[blockIdx.x] = t1 - t0; }
  } else if (warp < 4 && HAMMER) {
    float4* p = reinterpret_cast<float4*>(sm + 98304) + tid;
    float4 acc = make_float4(0, 0, 0, 0);
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 16; i++) { float4 v = p[i * 128]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      p[0] = acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <bool TS, int N, bool HAMMER>
void run3(const char* name, long long* dC) {
  CK(cudaFuncSetAttribute(k_probe3<TS, N, HAMMER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 133000));
  for (int rep = 0; rep < 2; rep++) {
    int nb = 64, grid = 148;
    k_probe3<TS, N, HAMMER><<<grid, 192, 133000>>>(nb, dC);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(grid);
    CK(cudaMemcpy(c.data(), dC, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0, mn = 1ll << 60; for (auto v : c) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
    double per = (double)mx / (nb * 8);
    printf("T3: %-28s: %lld..%lld cycles / %d blocks -> %.1f cycles per MMA (K=16), %.0f per 128x128x128-equivalent\n", name, mn, mx, nb, per, per * 8 * 128 / N);
  }
}

// ------------------------------------------------------------------------------------------------ T4
// mode 0: 32x32b.x32, mode 1: 16x256b.x8 (two 16-lane halves)   nwarps in {4, 8, 16}
__global__ void __launch_bounds__(512, 1) k_probe4(int mode, int iters, long long* out_cycles, float* sink) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
  const uint32_t colbase = (uint32_t)((warp >> 2) * 128) & 511u;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint32_t r[32];
        LD32(tmem + lane_base + colbase + 32 * j, r);
        ld_wait();
        acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        uint32_t r[32], q[32];
        LD256x8(tmem + lane_base + colbase + 64 * j, r);
        LD256x8(tmem + lane_base + (16u << 16) + colbase + 64 * j, q);
        ld_wait();
        acc += __uint_as_float(r[0]) + __uint_as_float(q[31]);
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (tid == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.f) sink[0] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}


// ------------------------------------------------------------------------------------------------ T5
// per 128x128x128 block: 8 UMMAs + NC commits to rotating mbarriers (the fused kernel's issue pattern)
template <int NC, int ALT>
__global__ void __launch_bounds__(192, 1) k_probe5(int nblocks, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (tid == 0) { for (int i = 0; i < 12; i++) mbar_init(smem_u32(bars + i), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
    long long t0 = clock64();
    for (int b = 0; b < nblocks; b++) {
      if (elect_one()) {
        // ALT bits: 1 = alternate accumulator per block, 2 = TS for every other pair of blocks, 4 = overwrite (accumulate=0) at block start,
        //           8 = all TS, 16 = N=256 accumulators side by side (single MMA writes 256 columns)
        const uint32_t D = tmem + (((ALT & 1) && (b & 1)) ? 128u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          const uint32_t off = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
          const uint32_t acc = ((ALT & 4) && ks == 0) ? 0u : 1u;
          if ((ALT & 8) || ((ALT & 2) && (b & 2))) mma_ts(D, tmem + 256 + 8 * ks, dB + (off >> 4), idesc, acc);
          else mma_ss(D, dA + (off >> 4), dB + (off >> 4), idesc, acc);
        }
#pragma unroll
        for (int c = 0; c < NC; c++) tc_commit(smem_u32(bars + 1 + ((b + c) % 8)));
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(smem_u32(bars));
    __syncwarp();
    mbar_wait(smem_u32(bars), 0);
    long long t1 = clock64();
    if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int NC, int ALT>
void run5(const char* name, long long* dC) {
  CK(cudaFuncSetAttribute(k_probe5<NC, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 133000));
  int nb = 64, grid = 148;
  k_probe5<NC, ALT><<<grid, 192, 133000>>>(nb, dC);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dC, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
  printf("T5: %-40s: %.0f cycles per 128x128x128 block\n", name, (double)mx / nb);
}

int main() {
  // ---------------- T1 + T2
  {
    std::vector<__nv_bfloat16> W(128 * 128);
    std::vector<float> Wf(128 * 128);
    for (int n = 0; n < 128; n++) for (int k = 0; k < 128; k++) { float v = (float)(((n * 5 + k) % 7) - 3); Wf[n * 128 + k] = v; W[n * 128 + k] = __float2bfloat16(v); }
    __nv_bfloat16* dW; float *dL, *dD;
    CK(cudaMalloc(&dW, W.size() * 2)); CK(cudaMalloc(&dL, 32 * 16 * 4)); CK(cudaMalloc(&dD, 128 * 128 * 4));
    CK(cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_probe12, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000));
    k_probe12<<<1, 160, 36000>>>(dW, dL, dD);
    CK(cudaDeviceSynchronize());
    std::vector<float> L(32 * 16), D(128 * 128);
    CK(cudaMemcpy(L.data(), dL, L.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    // decode layout: value = ((row*3 + col*5) % 11) - 4 is ambiguous, so use a table search over (row<32, col<16)
    printf("T1: 16x256b.x2 register -> (row, col) for lanes 0..7 (first 16-lane half), regs 0..7\n");
    auto val = [](int row, int col) { return (float)(((row * 3 + col * 5) % 11) - 4); };
    // check the hypothesis: reg[4*n + 2*h + c] of lane l = H[row = l/4 + 8*h][col = 8*n + 2*(l%4) + c]
    int bad = 0;
    for (int l = 0; l < 32; l++) for (int half = 0; half < 2; half++) for (int n = 0; n < 2; n++) for (int h = 0; h < 2; h++) for (int c = 0; c < 2; c++) {
      float got = L[l * 16 + half * 8 + n * 4 + h * 2 + c];
      float exp = val(16 * half + l / 4 + 8 * h, 8 * n + 2 * (l % 4) + c);
      if (got != exp) bad++;
    }
    printf("T1: hypothesis reg[4n+2h+c] = H[l/4 + 8h (+16 for the lane+16 address)][8n + 2(l%%4) + c]: %s (%d mismatches)\n", bad ? "WRONG" : "CONFIRMED", bad);
    if (bad) { for (int l = 0; l < 8; l++) { printf("  lane %d:", l); for (int i = 0; i < 16; i++) printf(" %g", L[l * 16 + i]); printf("\n"); } }
    // T2 reference
    double maxerr = 0;
    for (int m = 0; m < 128; m++) for (int n = 0; n < 128; n++) {
      double s = 0;
      for (int k = 0; k < 128; k++) { float h = val(m, k); h = h > 0 ? h : 0; s += (double)h * Wf[n * 128 + k]; }
      maxerr = fmax(maxerr, fabs(s - D[m * 128 + n]));
    }
    printf("T2: TS MMA (A = in-place bf16 relu of fp32 TMEM columns): max abs err %.3g  -> %s\n", maxerr, maxerr == 0 ? "EXACT" : "MISMATCH");
  }
  // ---------------- T3
  {
    long long* dC; CK(cudaMalloc(&dC, 148 * 8));
    run3<false, 128, false>("SS N=128", dC);
    run3<false, 256, false>("SS N=256", dC);
    run3<false, 64, false>("SS N=64", dC);
    run3<true, 128, false>("TS N=128", dC);
    run3<true, 256, false>("TS N=256", dC);
    run3<true, 64, false>("TS N=64", dC);
    run3<false, 128, true>("SS N=128 + smem traffic", dC);
    run3<true, 128, true>("TS N=128 + smem traffic", dC);
    run3<true, 64, true>("TS N=64 + smem traffic", dC);
  }
  // ---------------- T5
  {
    long long* dC; CK(cudaMalloc(&dC, 148 * 8));
    run5<1, 0>("SS same D", dC);
    run5<1, 1>("SS alternate D per block", dC);
    run5<1, 4>("SS same D, overwrite at block start", dC);
    run5<1, 5>("SS alternate D + overwrite", dC);
    run5<1, 8>("TS same D", dC);
    run5<1, 9>("TS alternate D", dC);
    run5<1, 2>("SS/TS switching every 2 blocks, same D", dC);
    run5<1, 3>("SS/TS switching + alternate D", dC);
    run5<1, 7>("SS/TS switching + alternate D + overwrite", dC);
  }
  // ---------------- T4
  {
    long long* dC; float* dS; CK(cudaMalloc(&dC, 148 * 8)); CK(cudaMalloc(&dS, 4));
    for (int grid : {1, 148}) for (int mode = 0; mode < 2; mode++) for (int nw : {4, 8, 16}) {
      int iters = 200;
      k_probe4<<<grid, nw * 32, 0>>>(mode, iters, dC, dS);
      CK(cudaDeviceSynchronize());
      std::vector<long long> c(grid);
      CK(cudaMemcpy(c.data(), dC, grid * 8, cudaMemcpyDeviceToHost));
      long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
      double bytes = (double)nw * iters * 32 * 128 * 4;
      printf("T4: grid %3d %-10s %2d warps: %lld cycles, %.1f B/clk/SM (%.0f cycles per 64 KB tile)\n", grid, mode ? "16x256b.x8" : "32x32b.x32", nw, mx, bytes / mx, 65536.0 / (bytes / mx));
    }
  }
  return 0;
}
