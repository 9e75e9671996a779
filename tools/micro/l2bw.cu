// L2 read bandwidth microbenchmark: every CTA streams the same L2-resident buffer with 128-bit loads.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const float4* __restrict__ p, size_t n4, int iters, float* out) {
  float4 acc = make_float4(0, 0, 0, 0);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; it++) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * 4) {
      float4 a = __ldcg(p + i);
      float4 b = i + stride < n4 ? __ldcg(p + i + stride) : a;
      float4 c = i + 2 * stride < n4 ? __ldcg(p + i + 2 * stride) : a;
      float4 d = i + 3 * stride < n4 ? __ldcg(p + i + 3 * stride) : a;
      acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
    }
  }
  if (acc.x == 123.456f) out[0] = acc.y;
}
int main() {
  for (size_t mb : {8, 32, 64, 96, 512}) {
    size_t bytes = mb << 20, n4 = bytes / 16;
    float4* p; float* o;
    cudaMalloc(&p, bytes); cudaMalloc(&o, 4); cudaMemset(p, 0, bytes);
    for (int bs : {256, 1024}) {
      int grid = 148 * (2048 / bs);
      int iters = mb >= 512 ? 4 : 40;
      rd<<<grid, bs>>>(p, n4, 2, o);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a); rd<<<grid, bs>>>(p, n4, iters, o); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("buf %4zu MB  block %4d grid %5d : %.1f GB/s\n", mb, bs, grid, (double)bytes * iters / ms / 1e6);
    }
    cudaFree(p); cudaFree(o);
  }
  return 0;
}
