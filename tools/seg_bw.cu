// Microbenchmark: HBM read bandwidth when a column-major matrix is streamed panel by panel,
// i.e. as contiguous segments of SEG bytes separated by the column stride (design input for
// the one-pass normal-operator kernel).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
template <int LPC>
__global__ void __launch_bounds__(512, 1) seg_read(const float4* __restrict__ A, long long ldv, int panels, int n, int cpw, float* out) {
  constexpr int NSEG = 32 / LPC;
  const int lane = threadIdx.x & 31, seg = lane / LPC, li = lane % LPC;
  const long long gw = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const long long c0 = gw * cpw, c1 = min((long long)n, c0 + cpw);
  float acc = 0.f;
  for (int p = 0; p < panels; ++p) {
    const float4* base = A + (long long)p * LPC + li;
    for (long long c = c0 + seg; c < c1; c += NSEG * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        long long col = c + (long long)u * NSEG;
        v[u] = col < c1 ? ldg_stream(base + col * ldv) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
int main(int argc, char** argv) {
  const long long m = 16384, n = 65536;   // float32
  const long long ldv = m / 4;
  float4* A; float* out;
  cudaMalloc(&A, m * n * 4); cudaMalloc(&out, 4);
  cudaMemset(A, 0, m * n * 4);
  int sms = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps_per_cta : {16}) {
    for (int ctas_per_sm : {1, 2, 4}) {
      int grid = sms * ctas_per_sm; int block = warps_per_cta * 32 / ctas_per_sm; if (block < 128) block = 128;
      long long W = (long long)grid * (block / 32);
      int cpw = (int)((n + W - 1) / W);
      auto run = [&](int lpc) {
        int panels = (int)(ldv / lpc);
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          switch (lpc) {
            case 4: seg_read<4><<<grid, block>>>(A, ldv, panels, (int)n, cpw, out); break;
            case 8: seg_read<8><<<grid, block>>>(A, ldv, panels, (int)n, cpw, out); break;
            case 16: seg_read<16><<<grid, block>>>(A, ldv, panels, (int)n, cpw, out); break;
            case 32: seg_read<32><<<grid, block>>>(A, ldv, panels, (int)n, cpw, out); break;
          }
          cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("seg=%4d B  grid=%d block=%d cpw=%d : %.3f ms  %.0f GB/s  (%s)\n", lpc * 16, grid, block, cpw, ms, m * n * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      };
      for (int lpc : {4, 8, 16, 32}) run(lpc);
    }
  }
  return 0;
}
