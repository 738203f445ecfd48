// rls_normal_tma.cu — one-HBM-pass normal operator  g = A'(A x)  as a TMA-fed two-phase
// streaming kernel (the roofline-defining kernel of the package).
//
// Persistent cooperative kernel, one CTA per SM; every CTA owns a fixed range of columns
// (multiples of 32).  A row panel (PR rows x all columns) is used twice:
//   phase 1  y_p = A_p x   : the CTA's [PR x cols] tile streams from HBM through a shared-memory
//                            ring (TMA cp.async.bulk.tensor + mbarrier complete_tx) and is
//                            reduced against x; the PR-vector of partial sums is exchanged
//                            between CTAs through per-CTA slots in L2 (all-gather with a fixed
//                            summation order => deterministic, bit-identical on every CTA);
//   phase 2  g += A_p' y_p : D panels later the same tile streams through the ring again — now
//                            an L2 hit (phase-1 loads carry an evict_last policy, phase-2 loads
//                            evict_first) — and is reduced against y_p.
// HBM therefore sees A once; the lag D hides the latency of the y exchange completely, and
// because nothing has to stay resident in shared memory the panels can be tall enough
// (256-byte column segments) for full HBM efficiency.  g lives in registers for the whole
// launch and is written once.  Warp roles: 16 compute warps, 1 TMA producer, 1 sender
// (CTA-level y reduction + slot publish), 4 gatherers.  Every wait is bounded: a time-out
// raises an abort flag instead of hanging the GPU.  Out-of-range rows / columns are
// zero-filled by TMA, so there is no tail masking.  Float32 and interleaved ComplexF32.
#include <cuda.h>

#include <type_traits>

#include "rls_common.cuh"

namespace {

constexpr int T_NCW = 16;              // compute warps
constexpr int T_NGW = 2;               // gatherer warps (cooperate on every panel; the lag hides their latency)
constexpr int T_CT = T_NCW * 32;
constexpr int T_THREADS = (T_NCW + 2 + T_NGW) * 32;
constexpr int T_BOXC = 32;             // columns per TMA box
constexpr int T_NY = 8;                // y ring in shared memory (lag <= T_NY - 1)
constexpr int T_NB = 16;               // slot / counter ring in global memory (lag <= (T_NB - 2) / 2)
constexpr int T_MAXLAG = 6;
constexpr int T_MAXI = 4;              // max float4 sweeps of the compute warps over a stage
constexpr int T_GLD = 10;              // slot loads in flight per gatherer lane
constexpr unsigned T_SPIN_LIMIT = 4000000u;

struct TmaWs {
  float4* slots;        // [T_NB][grid][16]
  unsigned* counter;    // [T_NB] arrival counters, 128 bytes apart (zeroed before every launch)
  int* abort_flag;
};
constexpr int T_CSTRIDE = 32;          // uints between counters (one L2 line each)

struct TmaArgs {
  const void* x;
  void* g;
  TmaWs ws;
  long long m, n;       // rows / columns (elements)
  int panels;           // number of row panels
  int nblk;             // total 32-column blocks
  int sb;               // boxes per stage
  int nstages;          // ring depth S
  int lag;              // D: phase 2 of panel p runs during step p + D
  int stage_bytes;
  int use_hint;
  int p2_ldg;           // 1: phase 2 re-reads the panel with plain 128-bit loads from L2 (TMA carries A only once)
  const void* A;
  long long ld;
  const int* gate;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait; returns false (and raises the abort flag) on time-out
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity, int* abort_flag) {
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > T_SPIN_LIMIT) { *abort_flag = 1; return false; }
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4shfl_xor(float4 a, int o) {
  return make_float4(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o),
                     __shfl_xor_sync(0xffffffffu, a.z, o), __shfl_xor_sync(0xffffffffu, a.w, o));
}

__device__ __forceinline__ float4 ldg_l2(const float4* p, unsigned long long pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
template <typename T>
__device__ __forceinline__ float4 mask_rows4(float4 v, long long row0, long long m) {
  constexpr int VEC = Elem<T>::vec;
  if (row0 + VEC <= m) return v;
  float t[4] = {v.x, v.y, v.z, v.w};
  constexpr int FPE = 4 / VEC;
#pragma unroll
  for (int e = 0; e < VEC; ++e)
    if (row0 + e >= m)
      for (int f = 0; f < FPE; ++f) t[e * FPE + f] = 0.f;
  return make_float4(t[0], t[1], t[2], t[3]);
}

template <typename T> __device__ __forceinline__ void fma_y(float4& acc, float4 a, T x);
template <> __device__ __forceinline__ void fma_y<float>(float4& acc, float4 a, float x) {
  acc.x = fmaf(a.x, x, acc.x); acc.y = fmaf(a.y, x, acc.y); acc.z = fmaf(a.z, x, acc.z); acc.w = fmaf(a.w, x, acc.w);
}
template <> __device__ __forceinline__ void fma_y<float2>(float4& acc, float4 a, float2 x) {
  acc.x = fmaf(a.x, x.x, acc.x); acc.x = fmaf(-a.y, x.y, acc.x);
  acc.y = fmaf(a.x, x.y, acc.y); acc.y = fmaf(a.y, x.x, acc.y);
  acc.z = fmaf(a.z, x.x, acc.z); acc.z = fmaf(-a.w, x.y, acc.z);
  acc.w = fmaf(a.z, x.y, acc.w); acc.w = fmaf(a.w, x.x, acc.w);
}
template <typename T> __device__ __forceinline__ void fma_g(T& acc, float4 a, float4 y);
template <> __device__ __forceinline__ void fma_g<float>(float& acc, float4 a, float4 y) {
  acc = fmaf(a.x, y.x, acc); acc = fmaf(a.y, y.y, acc); acc = fmaf(a.z, y.z, acc); acc = fmaf(a.w, y.w, acc);
}
template <> __device__ __forceinline__ void fma_g<float2>(float2& acc, float4 a, float4 y) {
  acc.x = fmaf(a.x, y.x, acc.x); acc.x = fmaf(a.y, y.y, acc.x); acc.y = fmaf(a.x, y.y, acc.y); acc.y = fmaf(-a.y, y.x, acc.y);
  acc.x = fmaf(a.z, y.z, acc.x); acc.x = fmaf(a.w, y.w, acc.x); acc.y = fmaf(a.z, y.w, acc.y); acc.y = fmaf(-a.w, y.z, acc.y);
}

// shared-memory carve-up (dynamic): [stages][x][ypart][ysm][gsm][barriers]
template <typename T, int LPC>
struct SmemLayout {
  int stage_bytes, nstages, xcols;
  __host__ __device__ size_t off_x() const { return (size_t)stage_bytes * nstages; }
  __host__ __device__ size_t off_ypart() const { return off_x() + (((size_t)xcols * sizeof(T) + 127) & ~(size_t)127); }
  __host__ __device__ size_t off_ysm() const { return off_ypart() + sizeof(float4) * 2 * T_NCW * LPC; }
  __host__ __device__ size_t off_gsm() const { return off_ysm() + sizeof(float4) * T_NY * LPC; }
  __host__ __device__ size_t off_bars() const { return off_gsm() + sizeof(float4) * 2 * T_NGW * LPC; }
  __host__ __device__ size_t total() const { return off_bars() + sizeof(uint64_t) * (2 * 16 + 4 + T_NY + 4) + 64; }
};

// LPC = float4 per column segment (8: 128-byte, 16: 256-byte segments); NJ = max column chunks per panel
template <typename T, int LPC, int NJ>
__global__ void __launch_bounds__(T_THREADS, 1) normal_tma_kernel(const __grid_constant__ CUtensorMap tmap, TmaArgs a) {
  if (a.gate && *a.gate) return;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NGRP = 32 / LPC;          // columns covered by one warp-wide float4 read
  constexpr int BOX_BYTES = T_BOXC * LPC * 16;
  constexpr int PRF = LPC * 4;            // floats of y per panel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int S = a.nstages, D = a.lag, P = a.panels;

  // this CTA's column blocks
  const int blk0 = (int)(((long long)cta * a.nblk) / grid);
  const int blk1 = (int)(((long long)(cta + 1) * a.nblk) / grid);
  const int nb = blk1 - blk0;
  const int nch = (nb + a.sb - 1) / a.sb;               // stages (column chunks) per panel and phase
  const long long col0 = (long long)blk0 * T_BOXC;

  SmemLayout<T, LPC> L{a.stage_bytes, S, ((a.nblk + grid - 1) / grid + 1) * T_BOXC};
  uint8_t* stage_base = smem;
  T* xs = reinterpret_cast<T*>(smem + L.off_x());
  float4* ypart = reinterpret_cast<float4*>(smem + L.off_ypart());   // [2][NCW][LPC]
  float4* ysm = reinterpret_cast<float4*>(smem + L.off_ysm());       // [T_NY][LPC]
  float4* gsm = reinterpret_cast<float4*>(smem + L.off_gsm());       // [2][T_NGW][LPC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bars());
  uint64_t* full = bars;              // [16]
  uint64_t* empty = bars + 16;        // [16]
  uint64_t* yp_full = bars + 32;      // [2]
  uint64_t* yp_free = bars + 34;      // [2]
  uint64_t* yready = bars + 36;       // [T_NY]
  uint64_t* gdone = bars + 36 + T_NY; // [2] gatherer partials written
  uint64_t* gfree = gdone + 2;        // [2] gatherer partials consumed
  int* abort_flag = a.ws.abort_flag;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], T_NCW); }
    for (int k = 0; k < 2; ++k) { mbar_init(&yp_full[k], T_NCW); mbar_init(&yp_free[k], 1); }
    for (int k = 0; k < T_NY; ++k) mbar_init(&yready[k], 1);
    for (int k = 0; k < 2; ++k) { mbar_init(&gdone[k], T_NGW); mbar_init(&gfree[k], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
    for (int q = threadIdx.x; q < nb * T_BOXC; q += T_THREADS) {
      long long col = col0 + q;
      xs[q] = (col < a.n) ? x[col] : T{};
    }
  }
  __syncthreads();

  if (warp == T_NCW) {
    // ================================ TMA producer ====================================
    // issue order == consumption order: [phase-1 chunks of panel t][phase-2 chunks of panel t-D]
    if (lane == 0 && nb > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      unsigned long long pol_keep = 0, pol_drop = 0;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
      long long k = 0;
      for (int t = 0; t < P + D; ++t) {
        for (int ph = 0; ph < 2; ++ph) {
          const int pnl = ph == 0 ? t : t - D;
          if (pnl < 0 || pnl >= P) continue;
          if (ph == 1 && a.p2_ldg) continue;
          for (int j = 0; j < nch; ++j, ++k) {
            const int s = (int)(k % S);
            const unsigned use = (unsigned)(k / S);
            if (use > 0 && !mbar_wait(&empty[s], (use - 1) & 1, abort_flag)) return;
            const int bx0 = j * a.sb;
            const int nbx = min(a.sb, nb - bx0);
            mbar_expect_tx(&full[s], (unsigned)(nbx * BOX_BYTES));
            uint8_t* dst = stage_base + (size_t)s * a.stage_bytes;
            for (int bx = 0; bx < nbx; ++bx) {
              if (a.use_hint)
                tma_load_2d_hint(dst + (size_t)bx * BOX_BYTES, &tmap, pnl * PRF, (blk0 + bx0 + bx) * T_BOXC, &full[s],
                                 ph == 0 ? pol_keep : pol_drop);
              else
                tma_load_2d(dst + (size_t)bx * BOX_BYTES, &tmap, pnl * PRF, (blk0 + bx0 + bx) * T_BOXC, &full[s]);
            }
          }
        }
      }
    }
    return;
  }
  if (warp == T_NCW + 1) {
    // ================================ sender ==========================================
    for (int t = 0; t < P; ++t) {
      const int pb = t & 1;
      if (!mbar_wait(&yp_full[pb], (unsigned)(t >> 1) & 1, abort_flag)) return;
      const int b = t % T_NB;
      if (lane < LPC) {
        float4 s = ypart[(pb * T_NCW + 0) * LPC + lane];
#pragma unroll
        for (int w = 1; w < T_NCW; ++w) s = f4add(s, ypart[(pb * T_NCW + w) * LPC + lane]);
        a.ws.slots[((size_t)b * grid + cta) * 16 + lane] = s;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        atomicAdd(&a.ws.counter[b * T_CSTRIDE], 1u);
        mbar_arrive(&yp_free[pb]);
      }
    }
    return;
  }
  if (warp >= T_NCW + 2) {
    // ================================ gatherers =======================================
    // All T_NGW warps work on every panel: the grid's slots are dealt to T_NGW*NGRP lane groups,
    // every lane issues its loads in one batch, partial sums are combined in a fixed order
    // (lane order, shuffle tree, warp order) => deterministic and bit-identical on every CTA.
    const int gw = warp - (T_NCW + 2);
    const int gl = lane / LPC, r4 = lane % LPC;
    constexpr int NG = T_NGW * NGRP;
    const int g0 = gw * NGRP + gl;
    for (int t = 0; t < P; ++t) {
      const int b = t % T_NB;
      const unsigned target = (unsigned)grid * (unsigned)(t / T_NB + 1);   // counters are zeroed before every launch
      unsigned ok = 1;
      if (lane == 0) {
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(&a.ws.counter[b * T_CSTRIDE]) - target) < 0) {
          if (++spins > T_SPIN_LIMIT || *((volatile int*)abort_flag)) { ok = 0; break; }
          __nanosleep(100);
        }
        if (!ok) *abort_flag = 1;
      }
      ok = __shfl_sync(0xffffffffu, ok, 0);
      __syncwarp();
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const float4* __restrict__ sl = a.ws.slots + (size_t)b * grid * 16 + r4;
        for (int base = g0; base < grid; base += NG * T_GLD) {
          float4 v[T_GLD];
#pragma unroll
          for (int u = 0; u < T_GLD; ++u) {
            const int c = base + u * NG;
            v[u] = (c < grid) ? __ldcg(sl + (size_t)c * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < T_GLD; ++u) acc = f4add(acc, v[u]);
        }
#pragma unroll
        for (int o = LPC; o < 32; o <<= 1) acc = f4add(acc, f4shfl_xor(acc, o));
      }
      const int pb = t & 1;
      if (t >= 2 && !mbar_wait(&gfree[pb], (unsigned)((t >> 1) - 1) & 1, abort_flag)) return;
      if (lane < LPC) gsm[(pb * T_NGW + gw) * LPC + lane] = acc;
      __syncwarp();
      if (lane == 0) mbar_arrive(&gdone[pb]);
      if (gw == 0) {
        if (!mbar_wait(&gdone[pb], (unsigned)(t >> 1) & 1, abort_flag)) return;
        if (lane < LPC) {
          float4 s = gsm[(pb * T_NGW + 0) * LPC + lane];
#pragma unroll
          for (int w = 1; w < T_NGW; ++w) s = f4add(s, gsm[(pb * T_NGW + w) * LPC + lane]);
          ysm[(t % T_NY) * LPC + lane] = s;
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(&gfree[pb]); mbar_arrive(&yready[t % T_NY]); }
      }
    }
    return;
  }

  // ================================== compute warps =====================================
  const int r4 = lane % LPC, cg = lane / LPC;
  T gacc[NJ][T_MAXI];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int i = 0; i < T_MAXI; ++i) gacc[j][i] = T{};

  bool dead = false;
  long long k = 0;                       // running stage index, same order as the producer
  if (a.p2_ldg) {
    // ---- hybrid: phase 1 from the TMA ring (HBM), phase 2 with 128-bit loads that hit L2 ----
    constexpr int VEC = Elem<T>::vec;
    const float4* __restrict__ Av = reinterpret_cast<const float4*>(a.A);
    const long long ldv = a.ld / VEC;
    unsigned long long pol = 0;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int t = 0; t < P + D; ++t) {
      const int q = t - D;
      const bool do1 = t < P, do2 = q >= 0;
      float4 yacc = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const long long row0 = ((long long)q * LPC + r4) * VEC;           // first row of this lane's slice of panel q
      const bool rvalid = do2 && row0 < a.m;
      bool have_y = false;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (j < nch) {
          const int nbx = min(a.sb, nb - j * a.sb);
          const int nelem = nbx * T_BOXC * LPC;
          // (a) issue the phase-2 loads of chunk j (panel q) — they fly while phase 1 of chunk j runs
          float4 v2[T_MAXI];
#pragma unroll
          for (int i = 0; i < T_MAXI; ++i) {
            const int e = (i * T_NCW + warp) * 32 + lane;
            const long long col = col0 + (long long)j * a.sb * T_BOXC + (i * T_NCW + warp) * NGRP + cg;
            v2[i] = (rvalid && e < nelem && col < a.n) ? ldg_l2(Av + col * ldv + ((long long)q * LPC + r4), pol)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          // (b) phase 1 of chunk j (panel t) from the ring
          if (do1) {
            const int s = (int)(k % S);
            if (!dead && !mbar_wait(&full[s], (unsigned)(k / S) & 1, abort_flag)) dead = true;
            ++k;
            const float4* __restrict__ tile = reinterpret_cast<const float4*>(stage_base + (size_t)s * a.stage_bytes);
#pragma unroll
            for (int i = 0; i < T_MAXI; ++i) {
              const int e = (i * T_NCW + warp) * 32 + lane;
              if (e < nelem) {
                const int c = j * a.sb * T_BOXC + (i * T_NCW + warp) * NGRP + cg;
                fma_y<T>(yacc, tile[e], xs[c]);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
          }
          // (c) consume the phase-2 loads
          if (do2) {
            if (!have_y) {
              if (!dead && !mbar_wait(&yready[q % T_NY], (unsigned)(q / T_NY) & 1, abort_flag)) dead = true;
              y4 = ysm[(q % T_NY) * LPC + r4];
              have_y = true;
            }
            const bool partial = rvalid && row0 + VEC > a.m;
#pragma unroll
            for (int i = 0; i < T_MAXI; ++i) {
              float4 v = v2[i];
              if (partial) v = mask_rows4<T>(v, row0, a.m);
              fma_g<T>(gacc[j][i], v, y4);
            }
          }
        }
      }
      if (do1) {
#pragma unroll
        for (int o = LPC; o < 32; o <<= 1) yacc = f4add(yacc, f4shfl_xor(yacc, o));
        const int pb = t & 1;
        if (t >= 2 && !dead && !mbar_wait(&yp_free[pb], (unsigned)((t >> 1) - 1) & 1, abort_flag)) dead = true;
        if (lane < LPC) ypart[(pb * T_NCW + warp) * LPC + lane] = yacc;
        __syncwarp();
        if (lane == 0) mbar_arrive(&yp_full[pb]);
      }
    }
  } else {
  for (int t = 0; t < P + D; ++t) {
    if (t < P) {
      // ---- phase 1: y_t partial over this warp's share of the CTA's columns (HBM stream) ----
      float4 yacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (j < nch) {
          const int s = (int)(k % S);
          if (!dead && !mbar_wait(&full[s], (unsigned)(k / S) & 1, abort_flag)) dead = true;
          ++k;
          const int nbx = min(a.sb, nb - j * a.sb);
          const int nelem = nbx * T_BOXC * LPC;                 // float4 elements in this stage
          const float4* __restrict__ tile = reinterpret_cast<const float4*>(stage_base + (size_t)s * a.stage_bytes);
#pragma unroll
          for (int i = 0; i < T_MAXI; ++i) {
            const int e = (i * T_NCW + warp) * 32 + lane;
            if (e < nelem) {
              const int c = j * a.sb * T_BOXC + (i * T_NCW + warp) * NGRP + cg;   // column within the CTA
              fma_y<T>(yacc, tile[e], xs[c]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
      }
#pragma unroll
      for (int o = LPC; o < 32; o <<= 1) yacc = f4add(yacc, f4shfl_xor(yacc, o));
      const int pb = t & 1;
      if (t >= 2 && !dead && !mbar_wait(&yp_free[pb], (unsigned)((t >> 1) - 1) & 1, abort_flag)) dead = true;
      if (lane < LPC) ypart[(pb * T_NCW + warp) * LPC + lane] = yacc;
      __syncwarp();
      if (lane == 0) mbar_arrive(&yp_full[pb]);
    }
    if (t >= D) {
      // ---- phase 2: g += A_q' y_q, the tile comes back through the ring from L2 ----
      const int q = t - D;
      if (!dead && !mbar_wait(&yready[q % T_NY], (unsigned)(q / T_NY) & 1, abort_flag)) dead = true;
      const float4 y4 = ysm[(q % T_NY) * LPC + r4];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (j < nch) {
          const int s = (int)(k % S);
          if (!dead && !mbar_wait(&full[s], (unsigned)(k / S) & 1, abort_flag)) dead = true;
          ++k;
          const int nbx = min(a.sb, nb - j * a.sb);
          const int nelem = nbx * T_BOXC * LPC;
          const float4* __restrict__ tile = reinterpret_cast<const float4*>(stage_base + (size_t)s * a.stage_bytes);
#pragma unroll
          for (int i = 0; i < T_MAXI; ++i) {
            const int e = (i * T_NCW + warp) * 32 + lane;
            if (e < nelem) fma_g<T>(gacc[j][i], tile[e], y4);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
      }
    }
  }
  }
  // ---- write g: reduce each column's partial over its LPC lanes ----
  T* __restrict__ g = reinterpret_cast<T*>(a.g);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
#pragma unroll
    for (int i = 0; i < T_MAXI; ++i) {
      T s = gacc[j][i];
#pragma unroll
      for (int o = LPC / 2; o > 0; o >>= 1) {
        if constexpr (Elem<T>::is_complex) {
          s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
          s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        } else {
          s += __shfl_xor_sync(0xffffffffu, s, o);
        }
      }
      if (j < nch && r4 == 0) {
        const int nbx = min(a.sb, nb - j * a.sb);
        const int cst = (i * T_NCW + warp) * NGRP + cg;         // column within the stage
        const long long col = col0 + (long long)j * a.sb * T_BOXC + cst;
        if (cst < nbx * T_BOXC && col < a.n) g[col] = s;
      }
    }
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (encode_tiled_fn)p;
  return fn;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace

struct TmaPlan {
  rls_ctx_s* ctx = nullptr;
  rls_mat_s* A = nullptr;
  CUtensorMap tmap;
  int lpc = 0, nj = 0, grid = 0, sb = 0, nstages = 0, lag = 0, stage_bytes = 0, panels = 0, nblk = 0, use_hint = 1, p2_ldg = 0;
  const void* kernel = nullptr;
  size_t smem_bytes = 0;
  TmaWs ws{};
  void* ws_mem = nullptr;
};

void rls_tma_plan_destroy(TmaPlan* p) {
  if (!p) return;
  if (p->ws_mem) cudaFree(p->ws_mem);
  delete p;
}

template <typename T, int LPC>
static const void* pick_tma_kernel(int nj) {
  if (nj <= 1) return (const void*)normal_tma_kernel<T, LPC, 1>;
  if (nj <= 2) return (const void*)normal_tma_kernel<T, LPC, 2>;
  return (const void*)normal_tma_kernel<T, LPC, 4>;
}

template <typename T, int LPC>
static int32_t tma_configure(TmaPlan* p) {
  rls_ctx_s* c = p->ctx;
  rls_mat_s* A = p->A;
  constexpr int BOX_BYTES = T_BOXC * LPC * 16;
  const int grid = c->sm_count;
  const int nblk = (int)((A->n + T_BOXC - 1) / T_BOXC);
  const int nbmax = (nblk + grid - 1) / grid;                 // boxes of the widest CTA
  int dev_smem = 0;
  RLS_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  int sb = env_int("RLS_TMA_STAGE_KB", 32) * 1024 / BOX_BYTES;   // ~32 KB stages
  if (sb < 1) sb = 1;
  if (sb > nbmax) sb = nbmax;
  while ((sb * T_BOXC * LPC + T_CT - 1) / T_CT > T_MAXI) --sb;   // at most T_MAXI sweeps per stage
  int nch = (nbmax + sb - 1) / sb;
  if (nch > 4) {
    rls_set_error("one-pass(TMA): %lld columns over %d SMs need more than 4 stages per panel", (long long)A->n, grid);
    return RLS_ERR_UNSUPPORTED;
  }
  const int stage_bytes = sb * BOX_BYTES;
  SmemLayout<T, LPC> L{stage_bytes, 0, (nbmax + 1) * T_BOXC};
  int S = 16;
  for (; S >= 2; --S) {
    L.nstages = S;
    if (L.total() <= (size_t)dev_smem) break;
  }
  if (S < 2) {
    rls_set_error("one-pass(TMA): shared memory cannot hold two stages of %d bytes", stage_bytes);
    return RLS_ERR_UNSUPPORTED;
  }
  L.nstages = S;
  const int PR = LPC * Elem<T>::vec;
  p->panels = (int)((A->m + PR - 1) / PR);
  // lag: keep (lag + 1) panels inside ~40 % of L2, and give the y exchange at least ~8 us
  const double panel_bytes = (double)PR * (double)A->n * sizeof(T);
  const double l2 = (double)(c->l2_bytes ? c->l2_bytes : ((size_t)96 << 20));
  int lag = (int)(0.40 * l2 / panel_bytes) - 1;
  if (lag > T_MAXLAG) lag = T_MAXLAG;
  if (lag < 1) lag = 1;
  lag = std::max(1, std::min(T_MAXLAG, env_int("RLS_TMA_LAG", lag)));
  p->grid = grid; p->sb = sb; p->nstages = S; p->stage_bytes = stage_bytes; p->nblk = nblk; p->nj = nch; p->lag = lag;
  p->use_hint = env_int("RLS_TMA_HINT", 1);
  p->p2_ldg = env_int("RLS_TMA_P2LDG", 0);   // experimental hybrid (phase 2 by plain loads); off: measured slower
  p->smem_bytes = L.total();
  p->kernel = pick_tma_kernel<T, LPC>(nch);
  RLS_CUDA(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
  int per_sm = 0;
  RLS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p->kernel, T_THREADS, p->smem_bytes));
  if (per_sm < 1) {
    rls_set_error("one-pass(TMA): kernel does not fit on an SM (smem %zu)", p->smem_bytes);
    return RLS_ERR_UNSUPPORTED;
  }
  // tensor map over the Float32 view of A: dim0 = rows*floats-per-element, dim1 = columns
  encode_tiled_fn enc = get_encode();
  if (!enc) { rls_set_error("cuTensorMapEncodeTiled is not available from this driver"); return RLS_ERR_UNSUPPORTED; }
  const int fpe = Elem<T>::is_complex ? 2 : 1;
  cuuint64_t gdim[2] = {(cuuint64_t)A->m * fpe, (cuuint64_t)A->n};
  cuuint64_t gstr[1] = {(cuuint64_t)A->ld * sizeof(T)};
  cuuint32_t box[2] = {(cuuint32_t)(LPC * 4), (cuuint32_t)T_BOXC};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&p->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A->d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, LPC >= 16 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rls_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return RLS_ERR_UNSUPPORTED; }
  return RLS_OK;
}

int32_t rls_tma_plan_create(rls_ctx_s* c, rls_mat_s* A, TmaPlan** out) {
  if (A->m == 0 || A->n == 0) { rls_set_error("one-pass(TMA): empty matrix"); return RLS_ERR_UNSUPPORTED; }
  const size_t es = rls_elem_size(A->dtype);
  if (((uintptr_t)A->d % 16) != 0 || ((size_t)A->ld * es) % 16 != 0) {
    rls_set_error("one-pass(TMA): matrix base and column stride must be 16-byte aligned");
    return RLS_ERR_UNSUPPORTED;
  }
  TmaPlan* p = new TmaPlan();
  p->ctx = c;
  p->A = A;
  // 256-byte column segments reach full HBM read bandwidth, 128-byte ones ~83 % (tools/seg_bw.cu)
  p->lpc = env_int("RLS_TMA_LPC", 16);
  if (p->lpc != 8 && p->lpc != 16) p->lpc = 16;
  int32_t s;
  if (A->dtype == RLS_C32) s = p->lpc == 16 ? tma_configure<float2, 16>(p) : tma_configure<float2, 8>(p);
  else s = p->lpc == 16 ? tma_configure<float, 16>(p) : tma_configure<float, 8>(p);
  if (s != RLS_OK) { delete p; return s; }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  size_t o_slots = take(sizeof(float4) * T_NB * p->grid * 16);
  size_t o_cnt = take(sizeof(unsigned) * T_NB * T_CSTRIDE);
  size_t o_ab = take(sizeof(int));
  if (cudaMalloc(&p->ws_mem, off) != cudaSuccess) { delete p; rls_set_error("cudaMalloc failed for the one-pass workspace"); return RLS_ERR_NOMEM; }
  cudaMemsetAsync(p->ws_mem, 0, off, c->stream);
  char* b = (char*)p->ws_mem;
  p->ws.slots = (float4*)(b + o_slots);
  p->ws.counter = (unsigned*)(b + o_cnt);
  p->ws.abort_flag = (int*)(b + o_ab);
  *out = p;
  return RLS_OK;
}

int32_t rls_tma_apply(TmaPlan* p, const void* x, void* g, const int* gate) {
  rls_ctx_s* c = p->ctx;
  // the arrival counters are zeroed by a stream-ordered memset before every launch, so a launch that is
  // gated off on the device (done() already true) leaves nothing behind for the next one
  RLS_CUDA(cudaMemsetAsync(p->ws.counter, 0, sizeof(unsigned) * T_NB * T_CSTRIDE, c->stream));
  TmaArgs a;
  a.x = x; a.g = g; a.ws = p->ws;
  a.m = p->A->m; a.n = p->A->n;
  a.panels = p->panels; a.nblk = p->nblk; a.sb = p->sb; a.nstages = p->nstages; a.lag = p->lag; a.stage_bytes = p->stage_bytes;
  a.use_hint = p->use_hint;
  a.p2_ldg = p->p2_ldg;
  a.A = p->A->d;
  a.ld = p->A->ld;
  a.gate = gate;
  void* args[] = {(void*)&p->tmap, (void*)&a};
  RLS_CUDA(cudaLaunchCooperativeKernel(p->kernel, dim3(p->grid), dim3(T_THREADS), args, p->smem_bytes, c->stream));
  c->launches++;
  return RLS_OK;
}

int32_t rls_tma_check_abort(TmaPlan* p) {
  int flag = 0;
  RLS_CUDA(cudaMemcpyAsync(&flag, p->ws.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (flag) {
    rls_set_error("one-pass(TMA) normal operator timed out inside the panel pipeline (abort flag set)");
    return RLS_ERR_CUDA;
  }
  return RLS_OK;
}

void rls_tma_describe(TmaPlan* p, char* buf, int len) {
  snprintf(buf, len, "onepass/tma: grid=%d segment=%dB panels=%d lag=%d chunks/panel<=%d boxes/stage=%d stage=%dB stages=%d smem=%zuB hint=%d phase2=%s",
           p->grid, p->lpc * 16, p->panels, p->lag, p->nj, p->sb, p->stage_bytes, p->nstages, p->smem_bytes, p->use_hint,
           p->p2_ldg ? "ldg(L2)" : "tma(L2)");
}
