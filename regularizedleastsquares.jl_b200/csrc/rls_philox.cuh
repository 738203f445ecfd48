// rls_philox.cuh — counter-based Philox4x32-10 synthetic data.  The same function is
// restated in oracle/philox.py; both produce bit-identical floats (integer-exact
// distributions, two individually rounded float multiplies).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RLS_HD __host__ __device__ __forceinline__
#else
#define RLS_HD inline
#endif

struct Philox4 { uint32_t c[4]; };

RLS_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
  return o;
}

// value for (seed; element index idx, stream, component)
//   counter = (idx_lo, idx_hi, stream_lo*2+component, stream_hi) ; key = (seed_lo, seed_hi)
//   UNIFORM01: (c0 >> 8) * 2^-24 * scale
//   IH4      : ((c0>>8)+(c1>>8)+(c2>>8)+(c3>>8) - 2^25) -> float (rn) * (sqrt(3)*2^-24) * scale
RLS_HD float philox_value(uint64_t seed, uint64_t idx, uint64_t stream, uint32_t component, int dist, float scale) {
  Philox4 r = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(stream * 2ull + component),
                            (uint32_t)(stream >> 31), (uint32_t)seed, (uint32_t)(seed >> 32));
  if (dist == 0) {
    float u = (float)(r.c[0] >> 8) * 5.9604644775390625e-08f;  // exact
#ifdef __CUDA_ARCH__
    return __fmul_rn(u, scale);
#else
    return u * scale;
#endif
  }
  int32_t s = (int32_t)((r.c[0] >> 8) + (r.c[1] >> 8) + (r.c[2] >> 8) + (r.c[3] >> 8)) - (1 << 25);
  const float K = 1.0323827126512697e-07f;  // float32(sqrt(3) * 2^-24)
#ifdef __CUDA_ARCH__
  return __fmul_rn(__fmul_rn(__int2float_rn(s), K), scale);
#else
  return ((float)s * K) * scale;
#endif
}
