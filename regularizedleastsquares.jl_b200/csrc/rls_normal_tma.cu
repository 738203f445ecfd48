// rls_normal_tma.cu — single-HBM-pass normal operator  g = A'(A x)  with TMA-staged,
// shared-memory-resident row panels and a flag-in-data all-reduce of y (the roofline-defining
// kernel of the package).
//
// Persistent cooperative kernel, one CTA per SM.  Every CTA owns a fixed range of columns
// (multiples of 32).  A row panel (32 floats of every column = 128-byte segments) is used twice:
//   phase 1  y_p = A_p x   : the CTA's [32 x cols] tile is brought into a shared-memory ring by a
//                            TMA producer thread (cp.async.bulk.tensor + mbarrier complete_tx) and
//                            reduced against x by 14 compute warps;
//   exchange               : the CTA's 32 partial sums are published to its slot in L2 as
//                            {value, tag} 8-byte units (one 16-byte store per lane, no fence, no
//                            counter); every CTA gathers all slots (5-6 loads per lane, retried
//                            until the tags match) and sums them in a fixed order => deterministic
//                            and bit-identical on every CTA.  Measured 1.4 us per all-reduce against
//                            5.1 us for the slot/fence/atomic-counter protocol (tools/xchg_lat.cu);
//   phase 2  g += A_p' y_p : one panel later the SAME shared-memory tile is read again and then
//                            released to the producer.
// A is therefore delivered from HBM to the SMs exactly once (no L2 re-read); g lives in registers
// for the whole launch and is written once.  Every wait is bounded: a time-out raises an abort flag
// instead of hanging the GPU.  Out-of-range rows / columns are zero-filled by TMA (no tail masking).
// Float32 and interleaved ComplexF32.
#include <cuda.h>

#include <type_traits>

#include "rls_common.cuh"

namespace {

constexpr int T_NCW = 14;              // compute warps (+1 producer = 13 -> register allocation rounds to 16 warps: 128 regs/thread)
constexpr int T_CT = T_NCW * 32;       // compute threads
constexpr int T_THREADS = T_CT + 32;   // + the TMA producer warp
constexpr int T_BOXC = 32;             // columns per TMA box
constexpr int T_LPC = 8;               // float4 per column segment: 128-byte segments, 32 floats per panel column
constexpr int T_PRF = T_LPC * 4;       // floats of y per panel
constexpr int T_NB = 16;               // slot ring in global memory (lag 2 needs >= 6)
constexpr int T_MAXI = 4;              // max float4 sweeps of the compute warps over a stage
constexpr int T_GLD = 6;               // slot loads in flight per lane (grid <= 2*T_NCW*T_GLD per batch)
constexpr unsigned T_SPIN_LIMIT = 4000000u;

struct TmaWs {
  uint4* slots;         // [T_NB][grid][16] : {y[2e], tag, y[2e+1], tag}
  int* abort_flag;
};

struct TmaArgs {
  const void* x;
  void* g;
  TmaWs ws;
  long long m, n;       // rows / columns (elements)
  int panels;           // number of row panels
  int nblk;             // total 32-column blocks
  int sb;               // boxes per stage
  int nstages;          // ring depth S
  int stage_bytes;
  unsigned tag_base;    // tags of this launch: tag_base + panel + 1
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
// L2-coherent 16-byte accesses for the flag-in-data exchange (each 8-byte {value, tag} unit is single-copy atomic)
__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_cg_u4(uint4* p, uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(T_CT) : "memory"); }

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4shfl_xor(float4 a, int o) {
  return make_float4(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o),
                     __shfl_xor_sync(0xffffffffu, a.z, o), __shfl_xor_sync(0xffffffffu, a.w, o));
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

// shared-memory carve-up (dynamic): [stages][x][ypart][gsm][barriers]
template <typename T>
struct SmemLayout {
  int stage_bytes, nstages, xcols;
  __host__ __device__ size_t off_x() const { return (size_t)stage_bytes * nstages; }
  __host__ __device__ size_t off_ypart() const { return off_x() + (((size_t)xcols * sizeof(T) + 127) & ~(size_t)127); }
  __host__ __device__ size_t off_gsm() const { return off_ypart() + sizeof(float4) * T_NCW * T_LPC; }
  __host__ __device__ size_t off_bars() const { return off_gsm() + sizeof(float4) * T_NCW * T_LPC; }
  __host__ __device__ size_t total() const { return off_bars() + sizeof(uint64_t) * (2 * 16) + 64; }
};

// NJ = max column chunks (stages) per panel for this instantiation
template <typename T, int NJ>
__global__ void __launch_bounds__(T_THREADS, 1) normal_tma_kernel(const __grid_constant__ CUtensorMap tmap, TmaArgs a) {
  if (a.gate && *a.gate) return;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int LPC = T_LPC;
  constexpr int NGRP = 32 / LPC;          // columns covered by one warp-wide float4 read
  constexpr int BOX_BYTES = T_BOXC * LPC * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int S = a.nstages, P = a.panels;

  // this CTA's column blocks
  const int blk0 = (int)(((long long)cta * a.nblk) / grid);
  const int blk1 = (int)(((long long)(cta + 1) * a.nblk) / grid);
  const int nb = blk1 - blk0;
  const int nch = (nb + a.sb - 1) / a.sb;               // stages (column chunks) per panel
  const long long col0 = (long long)blk0 * T_BOXC;

  SmemLayout<T> L{a.stage_bytes, S, ((a.nblk + grid - 1) / grid + 1) * T_BOXC};
  uint8_t* stage_base = smem;
  T* xs = reinterpret_cast<T*>(smem + L.off_x());
  float4* ypart = reinterpret_cast<float4*>(smem + L.off_ypart());   // [NCW][LPC]  per-warp partial y of the current panel
  float4* gsm = reinterpret_cast<float4*>(smem + L.off_gsm());       // [NCW][LPC]  per-warp partial of the gathered y
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bars());
  uint64_t* full = bars;              // [16]
  uint64_t* empty = bars + 16;        // [16]
  int* abort_flag = a.ws.abort_flag;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], T_NCW); }
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
    if (lane == 0 && nb > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      long long k = 0;
      for (int t = 0; t < P; ++t) {
        for (int j = 0; j < nch; ++j, ++k) {
          const int s = (int)(k % S);
          const unsigned use = (unsigned)(k / S);
          if (use > 0 && !mbar_wait(&empty[s], (use - 1) & 1, abort_flag)) return;
          const int bx0 = j * a.sb;
          const int nbx = min(a.sb, nb - bx0);
          mbar_expect_tx(&full[s], (unsigned)(nbx * BOX_BYTES));
          uint8_t* dst = stage_base + (size_t)s * a.stage_bytes;
          for (int bx = 0; bx < nbx; ++bx)
            tma_load_2d(dst + (size_t)bx * BOX_BYTES, &tmap, t * T_PRF, (blk0 + bx0 + bx) * T_BOXC, &full[s]);
        }
      }
    }
    return;
  }

  // ================================== compute warps =====================================
  // step t:  (1) issue the gather loads for panel q = t-1 (they fly during phase 1)
  //          (2) phase 1 of panel t from the ring (tiles stay resident)
  //          (3) CTA reduction, publish {value, tag} pairs of panel t to this CTA's slot
  //          (4) validate / sum the gathered slots of panel q in a fixed order
  //          (5) phase 2 of panel q from the same tiles, release them to the producer
  const int r4 = lane % LPC, cg = lane / LPC;
  const int ep = lane & 15;                       // element pair (2 floats of y) this lane gathers
  const int grp = warp * 2 + (lane >> 4);         // CTA residue class (mod NG) this lane gathers
  constexpr int NG = 2 * T_NCW;
  T gacc[NJ][T_MAXI];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int i = 0; i < T_MAXI; ++i) gacc[j][i] = T{};

  bool dead = false;
  for (int t = 0; t <= P; ++t) {
    const int q = t - 1;
    const bool do1 = t < P, do2 = q >= 0;
    // ---- (1) gather loads for panel q ----
    uint4 gl[T_GLD];
    const unsigned want = a.tag_base + (unsigned)q + 1u;
    const uint4* __restrict__ sl = a.ws.slots + ((size_t)(q & (T_NB - 1)) * grid) * 16 + ep;
    if (do2) {
#pragma unroll
      for (int u = 0; u < T_GLD; ++u) {
        const int c = grp + NG * u;
        gl[u] = (c < grid) ? ld_cg_u4(sl + (size_t)c * 16) : make_uint4(0u, want, 0u, want);
      }
    }
    // ---- (2) phase 1 of panel t ----
    if (do1) {
      float4 yacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (j < nch) {
          const long long k = (long long)t * nch + j;
          const int s = (int)(k % S);
          if (!dead && !mbar_wait(&full[s], (unsigned)(k / S) & 1, abort_flag)) dead = true;
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
        }
      }
#pragma unroll
      for (int o = LPC; o < 32; o <<= 1) yacc = f4add(yacc, f4shfl_xor(yacc, o));
      if (lane < LPC) ypart[warp * LPC + lane] = yacc;
    }
    bar_compute();                                               // ypart(t) complete
    // ---- (3) CTA reduction in warp order, publish: one 16-byte store per lane, no fence ----
    if (do1 && warp == 0 && lane < 16) {
      const float2* yp2 = reinterpret_cast<const float2*>(ypart);
      float2 s2 = yp2[lane];
#pragma unroll
      for (int w = 1; w < T_NCW; ++w) { float2 v = yp2[w * (LPC * 2) + lane]; s2.x += v.x; s2.y += v.y; }
      const unsigned tag = a.tag_base + (unsigned)t + 1u;
      st_cg_u4(a.ws.slots + ((size_t)(t & (T_NB - 1)) * grid + cta) * 16 + lane,
               make_uint4(__float_as_uint(s2.x), tag, __float_as_uint(s2.y), tag));
    }
    if (do2) {
      // ---- (4) validate the gathered slots (retry until every unit carries panel q's tag), fixed-order sum ----
      float2 acc = make_float2(0.f, 0.f);
      unsigned spins = 0;
#pragma unroll
      for (int u = 0; u < T_GLD; ++u) {
        const int c = grp + NG * u;
        while (!dead && (gl[u].y != want || gl[u].w != want)) {
          if (++spins > T_SPIN_LIMIT || *((volatile int*)abort_flag)) { *abort_flag = 1; dead = true; break; }
          gl[u] = ld_cg_u4(sl + (size_t)c * 16);
        }
        acc.x += __uint_as_float(gl[u].x);
        acc.y += __uint_as_float(gl[u].z);
      }
      for (int base = NG * T_GLD; base < grid; base += NG * T_GLD) {   // grids wider than NG*T_GLD CTAs
        for (int u = 0; u < T_GLD; ++u) {
          const int c = base + grp + NG * u;
          if (c >= grid) break;
          uint4 v = ld_cg_u4(sl + (size_t)c * 16);
          while (!dead && (v.y != want || v.w != want)) {
            if (++spins > T_SPIN_LIMIT) { *abort_flag = 1; dead = true; break; }
            v = ld_cg_u4(sl + (size_t)c * 16);
          }
          acc.x += __uint_as_float(v.x);
          acc.y += __uint_as_float(v.z);
        }
      }
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
      if (lane < 16) reinterpret_cast<float2*>(gsm)[warp * 16 + lane] = acc;
    }
    bar_compute();                                               // gsm(q) complete; ypart may be rewritten
    if (do2) {
      // ---- (5) y_q = sum over warps (fixed order), phase 2: g += A_q' y_q, release the tiles ----
      float4 y4 = gsm[r4];
#pragma unroll
      for (int w = 1; w < T_NCW; ++w) y4 = f4add(y4, gsm[w * LPC + r4]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (j < nch) {
          const long long k = (long long)q * nch + j;
          const int s = (int)(k % S);
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
  int nj = 0, grid = 0, sb = 0, nstages = 0, stage_bytes = 0, panels = 0, nblk = 0;
  const void* kernel = nullptr;
  size_t smem_bytes = 0;
  TmaWs ws{};
  void* ws_mem = nullptr;
  size_t ws_bytes = 0;
  unsigned tag_next = 1;
};

void rls_tma_plan_destroy(TmaPlan* p) {
  if (!p) return;
  if (p->ws_mem) cudaFree(p->ws_mem);
  delete p;
}

template <typename T>
static const void* pick_tma_kernel(int nj) {
  if (nj <= 1) return (const void*)normal_tma_kernel<T, 1>;
  if (nj <= 2) return (const void*)normal_tma_kernel<T, 2>;
  return (const void*)normal_tma_kernel<T, 4>;
}

template <typename T>
static int32_t tma_configure(TmaPlan* p) {
  rls_ctx_s* c = p->ctx;
  rls_mat_s* A = p->A;
  constexpr int BOX_BYTES = T_BOXC * T_LPC * 16;
  const int grid = c->sm_count;
  const int nblk = (int)((A->n + T_BOXC - 1) / T_BOXC);
  const int nbmax = (nblk + grid - 1) / grid;                 // boxes of the widest CTA
  int dev_smem = 0;
  RLS_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  int sb = env_int("RLS_TMA_STAGE_KB", 28) * 1024 / BOX_BYTES;   // ~28 KB stages
  if (sb < 1) sb = 1;
  if (sb > nbmax) sb = nbmax;
  int nch = (nbmax + sb - 1) / sb;
  if (nch > 4) { nch = 4; sb = (nbmax + nch - 1) / nch; }
  if ((sb * T_BOXC * T_LPC + T_CT - 1) / T_CT > T_MAXI) {
    rls_set_error("one-pass(TMA): %lld columns over %d SMs exceed the per-CTA tile budget", (long long)A->n, grid);
    return RLS_ERR_UNSUPPORTED;
  }
  const int stage_bytes = sb * BOX_BYTES;
  SmemLayout<T> L{stage_bytes, 0, (nbmax + 1) * T_BOXC};
  int S = 16;
  for (; S >= 2; --S) {
    L.nstages = S;
    if (L.total() <= (size_t)dev_smem) break;
  }
  // two resident panels (phase 2 runs one panel behind phase 1) and at least one prefetch stage
  if (S < 2 * nch + 1) {
    rls_set_error("one-pass(TMA): shared memory cannot hold the panel pipeline for %lld columns per SM", (long long)A->n / grid);
    return RLS_ERR_UNSUPPORTED;
  }
  L.nstages = S;
  p->grid = grid; p->sb = sb; p->nstages = S; p->stage_bytes = stage_bytes; p->nblk = nblk; p->nj = nch;
  p->smem_bytes = L.total();
  const int PR = T_LPC * Elem<T>::vec;
  p->panels = (int)((A->m + PR - 1) / PR);
  p->kernel = pick_tma_kernel<T>(nch);
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
  cuuint32_t box[2] = {(cuuint32_t)T_PRF, (cuuint32_t)T_BOXC};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&p->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A->d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  int32_t s = (A->dtype == RLS_C32) ? tma_configure<float2>(p) : tma_configure<float>(p);
  if (s != RLS_OK) { delete p; return s; }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  size_t o_slots = take(sizeof(uint4) * T_NB * p->grid * 16);
  size_t o_ab = take(sizeof(int));
  if (cudaMalloc(&p->ws_mem, off) != cudaSuccess) { delete p; rls_set_error("cudaMalloc failed for the one-pass workspace"); return RLS_ERR_NOMEM; }
  p->ws_bytes = off;
  cudaMemsetAsync(p->ws_mem, 0, off, c->stream);   // tag 0 is never used by a launch
  char* b = (char*)p->ws_mem;
  p->ws.slots = (uint4*)(b + o_slots);
  p->ws.abort_flag = (int*)(b + o_ab);
  *out = p;
  return RLS_OK;
}

int32_t rls_tma_apply(TmaPlan* p, const void* x, void* g, const int* gate) {
  rls_ctx_s* c = p->ctx;
  if (p->tag_next > 0xffffffffu - (unsigned)(p->panels + 2)) {   // tag wrap: restart from a clean slate
    RLS_CUDA(cudaMemsetAsync(p->ws.slots, 0, sizeof(uint4) * T_NB * p->grid * 16, c->stream));
    p->tag_next = 1;
  }
  TmaArgs a;
  a.x = x; a.g = g; a.ws = p->ws;
  a.m = p->A->m; a.n = p->A->n;
  a.panels = p->panels; a.nblk = p->nblk; a.sb = p->sb; a.nstages = p->nstages; a.stage_bytes = p->stage_bytes;
  a.tag_base = p->tag_next;   // unique per launch, so a launch that is gated off on the device leaves nothing behind
  a.gate = gate;
  p->tag_next += (unsigned)p->panels + 1u;
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
  snprintf(buf, len, "onepass/tma: smem-resident panels, flag-in-data exchange; grid=%d segment=128B panels=%d chunks/panel<=%d boxes/stage=%d stage=%dB stages=%d smem=%zuB",
           p->grid, p->panels, p->nj, p->sb, p->stage_bytes, p->nstages, p->smem_bytes);
}
