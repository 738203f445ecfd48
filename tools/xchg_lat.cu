// Microbenchmark: latency of one grid-wide all-reduce of a 32-float vector between 148 persistent CTAs
// through L2 (the y-panel exchange of the one-pass normal operator), for two protocols:
//   P1: slot store -> __threadfence -> atomic counter -> poll counter -> gather slots
//   P2: flag-in-data (LL): {value, tag} 8-byte units, no fence / counter; gather polls the data itself
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) { uint4 r; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ void st_cg_u4(uint4* p, uint4 v) { asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

__global__ void __launch_bounds__(512, 1) p1(float4* slots, unsigned* counter, int rounds, float* out) {
  const int grid = gridDim.x, cta = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float4 gsm[4][8];
  __shared__ float4 y[8];
  float acc0 = 0.f;
  for (int t = 0; t < rounds; ++t) {
    const int b = t & 15;
    if (warp == 0) {
      if (lane < 8) slots[((size_t)b * grid + cta) * 8 + lane] = make_float4(t + lane + acc0 * 1e-30f, 1.f, 2.f, 3.f);
      __threadfence(); __syncwarp();
      if (lane == 0) atomicAdd(&counter[b * 32], 1u);
    }
    if (warp < 4) {
      const unsigned target = (unsigned)grid * (unsigned)(t / 16 + 1);
      if (lane == 0) while ((int)(ld_acq(&counter[b * 32]) - target) < 0) {}
      __syncwarp();
      const int r4 = lane & 7, g0 = warp * 4 + (lane >> 3);
      float4 v[10];
#pragma unroll
      for (int u = 0; u < 10; ++u) { int c = g0 + 16 * u; v[u] = c < grid ? __ldcg(&slots[((size_t)b * grid + c) * 8 + r4]) : make_float4(0, 0, 0, 0); }
      float4 a = v[0];
#pragma unroll
      for (int u = 1; u < 10; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
      for (int o = 8; o < 32; o <<= 1) { a.x += __shfl_xor_sync(~0u, a.x, o); a.y += __shfl_xor_sync(~0u, a.y, o); a.z += __shfl_xor_sync(~0u, a.z, o); a.w += __shfl_xor_sync(~0u, a.w, o); }
      if (lane < 8) gsm[warp][lane] = a;
    }
    __syncthreads();
    if (threadIdx.x < 8) { float4 s = gsm[0][threadIdx.x]; for (int w = 1; w < 4; ++w) { s.x += gsm[w][threadIdx.x].x; } y[threadIdx.x] = s; }
    __syncthreads();
    acc0 += y[0].x;
  }
  if (threadIdx.x == 0 && cta == 0) out[0] = acc0;
}

__global__ void __launch_bounds__(512, 1) p2(uint4* slots, int rounds, unsigned tag_base, float* out) {
  const int grid = gridDim.x, cta = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float2 gsm[16][16];
  float acc0 = 0.f;
  for (int t = 0; t < rounds; ++t) {
    const int b = t & 15;
    const unsigned tag = tag_base + t + 1;
    if (warp == 0 && lane < 16) st_cg_u4(&slots[((size_t)b * grid + cta) * 16 + lane], make_uint4(__float_as_uint((float)(t + lane) + acc0 * 1e-30f), tag, __float_as_uint(1.f), tag));
    const int ep = lane & 15, grp = warp * 2 + (lane >> 4);
    uint4 v[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) { int c = grp + 32 * u; v[u] = c < grid ? ld_cg_u4(&slots[((size_t)b * grid + c) * 16 + ep]) : make_uint4(0, tag, 0, tag); }
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      int c = grp + 32 * u;
      while (v[u].y != tag || v[u].w != tag) v[u] = ld_cg_u4(&slots[((size_t)b * grid + c) * 16 + ep]);
      a.x += __uint_as_float(v[u].x); a.y += __uint_as_float(v[u].z);
    }
    a.x += __shfl_xor_sync(~0u, a.x, 16); a.y += __shfl_xor_sync(~0u, a.y, 16);
    if (lane < 16) gsm[warp][lane] = a;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < 16; ++w) s += gsm[w][0].x;
    acc0 += s;
    __syncthreads();
  }
  if (threadIdx.x == 0 && cta == 0) out[0] = acc0;
}

int main() {
  int grid = 148, rounds = 4000;
  float4* slots; unsigned* counter; float* out; uint4* slots2;
  cudaMalloc(&slots, sizeof(float4) * 16 * grid * 8); cudaMalloc(&counter, 16 * 32 * 4); cudaMalloc(&out, 4);
  cudaMalloc(&slots2, sizeof(uint4) * 16 * grid * 16); cudaMemset(slots2, 0, sizeof(uint4) * 16 * grid * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(counter, 0, 16 * 32 * 4);
    void* a1[] = {&slots, &counter, &rounds, &out};
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)p1, dim3(grid), dim3(512), a1, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("P1 counter+fence : %.3f us per all-reduce round (%s)\n", ms * 1e3 / rounds, cudaGetErrorString(cudaGetLastError()));
    unsigned tb = 1 + rep * (rounds + 1);
    void* a2[] = {&slots2, &rounds, &tb, &out};
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)p2, dim3(grid), dim3(512), a2, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("P2 flag-in-data  : %.3f us per all-reduce round (%s)\n", ms * 1e3 / rounds, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
