// Microbenchmark of the prologue / epilogue load patterns of k_tc_pair (persistent CTAs, few warps).
#include <cstdio>
#include <cuda_runtime.h>
template <int UNROLL>
__global__ void __launch_bounds__(320, 1) k(const float* __restrict__ x, long R, float* out, int nwarps, int smem_dummy) {
  extern __shared__ float sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps) return;
  int l8 = lane & 7, rsub = lane >> 3;
  float acc = 0.f;
  long ntiles = R / 128;
  int wpt = 4;  // warps per 128-row tile
  int tiles_per_iter = nwarps / wpt;
  for (long t0 = (long)blockIdx.x * tiles_per_iter; t0 < ntiles; t0 += (long)gridDim.x * tiles_per_iter) {
    long tile = t0 + warp / wpt;
    if (tile >= ntiles) break;
    long row0 = tile * 128 + (warp % wpt) * 32;
#pragma unroll UNROLL
    for (int i = 0; i < 8; i++) {
      long r = row0 + i * 4 + rsub;
      const float4* p = reinterpret_cast<const float4*>(x + r * 128) + l8;
      float4 v0 = __ldg(p), v1 = __ldg(p + 8), v2 = __ldg(p + 16), v3 = __ldg(p + 24);
      float s = v0.x + v1.x + v2.x + v3.x;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      acc += s;
    }
  }
  if (acc == 123.f) out[0] = acc;
}
int main() {
  long R = 2097152;
  float *x, *o;
  cudaMalloc(&x, R * 128 * 4); cudaMalloc(&o, 4); cudaMemset(x, 0, R * 128 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int smem : {0, 200 * 1024}) {
    for (int nw : {8, 10}) {
      for (int u = 0; u < 3; u++) {
        auto run = [&](int un) {
          if (un == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<2><<<148, 320, smem>>>(x, R, o, nw == 10 ? 8 : nw, 0); }
          if (un == 4) { cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<4><<<148, 320, smem>>>(x, R, o, nw == 10 ? 8 : nw, 0); }
          if (un == 8) { cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<8><<<148, 320, smem>>>(x, R, o, nw == 10 ? 8 : nw, 0); }
        };
        int un = u == 0 ? 2 : (u == 1 ? 4 : 8);
        run(un);
        cudaEventRecord(a); run(un); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("smem %6d  warps %d unroll %d: %.3f ms  %.0f GB/s  %s\n", smem, nw, un, ms, R * 512.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
