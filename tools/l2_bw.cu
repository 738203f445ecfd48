// Microbenchmark: aggregate L2 -> SM read bandwidth for an L2-resident working set (design input:
// can the normal operator afford to deliver A to the SMs twice per apply?)
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__global__ void __launch_bounds__(512) rd(const float4* __restrict__ A, size_t n4, int reps, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ldg_stream(A + i + u * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
int main() {
  float4* A; float* out;
  size_t maxb = (size_t)8 << 30;
  cudaMalloc(&A, maxb); cudaMalloc(&out, 4); cudaMemset(A, 0, maxb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 4096}) {
    size_t bytes = mb << 20; size_t n4 = bytes / 16;
    int reps = (int)(((size_t)16 << 30) / bytes); if (reps < 2) reps = 2;
    for (int grid : {148 * 2, 148 * 4}) {
      rd<<<grid, 512>>>(A, n4, 2, out);
      cudaEventRecord(e0);
      rd<<<grid, 512>>>(A, n4, reps, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("working set %5zu MB grid=%d: %.0f GB/s (%s)\n", mb, grid, (double)bytes * reps / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
