// rls_normal_tma.cu — single-HBM-pass normal operator  g = A'(A x)  with TMA-staged,
// shared-memory-resident row panels (the roofline-defining kernel of the package).
//
// Persistent cooperative kernel, one CTA per SM.  Every CTA owns a fixed range of columns
// (multiples of 32).  A row panel (PR rows x all columns) is consumed in two phases:
//   phase 1  y_p = A_p x   : the CTA's [PR x cols] tile is brought into a shared-memory ring
//                            by TMA (cp.async.bulk.tensor, mbarrier complete_tx) and reduced
//                            against x; the PR-vector of partial sums is exchanged between
//                            CTAs through per-CTA slots in L2 (all-gather, fixed summation
//                            order => deterministic and identical on every CTA);
//   phase 2  g += A_p' y_p : the SAME shared-memory tile is read again D panels later and
//                            then released to the TMA producer.
// A is therefore read from HBM exactly once and never re-read from L2; g lives in
// registers for the whole launch and is written once.  Warp roles: 16 compute warps,
// 1 TMA producer, 1 sender (CTA-level y reduction + slot publish), 2 gatherers.
// Every wait is bounded: a time-out raises an abort flag instead of hanging the GPU.
//
// Works for any column-major (m, n, ld) that TMA can describe (16-byte aligned base and
// column stride); out-of-range rows / columns are zero-filled by TMA, so there is no tail
// masking.  Float32 and interleaved ComplexF32.
#include <cuda.h>

#include "rls_common.cuh"

namespace {

constexpr int T_NCW = 16;              // compute warps
constexpr int T_THREADS = (T_NCW + 4) * 32;
constexpr int T_BOXC = 32;             // columns per TMA box
constexpr int T_NY = 8;                // y ring (shared memory) and slot ring (global)
constexpr int T_MAXJ = 4;              // max column chunks (stages) per panel
constexpr int T_MAXI = 4;              // max float4 sweeps of the compute warps over a stage
constexpr unsigned T_SPIN_LIMIT = 4000000u;

struct TmaWs {
  float4* slots;        // [T_NY][grid][LPCmax=16]
  unsigned* counter;    // [T_NY] monotonic arrival counters
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
  int lag;              // D
  int stage_bytes;
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

// shared-memory carve-up (dynamic): [stages][x][ypart][ysm][barriers]
template <typename T, int LPC>
struct SmemLayout {
  int stage_bytes, nstages, xcols;
  __host__ __device__ size_t off_x() const { return (size_t)stage_bytes * nstages; }
  __host__ __device__ size_t off_ypart() const { return off_x() + (((size_t)xcols * sizeof(T) + 127) & ~(size_t)127); }
  __host__ __device__ size_t off_ysm() const { return off_ypart() + sizeof(float4) * 2 * T_NCW * LPC; }
  __host__ __device__ size_t off_bars() const { return off_ysm() + sizeof(float4) * T_NY * LPC; }
  __host__ __device__ size_t total(int) const { return off_bars() + sizeof(uint64_t) * (2 * 16 + 4 + T_NY) + 64; }
};

template <typename T, int LPC>
__global__ void __launch_bounds__(T_THREADS, 1) normal_tma_kernel(const __grid_constant__ CUtensorMap tmap, TmaArgs a) {
  if (a.gate && *a.gate) return;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NGRP = 32 / LPC;          // columns covered by one warp-wide float4 read
  constexpr int BOX_BYTES = T_BOXC * LPC * 16;
  constexpr int PRF = LPC * 4;            // floats per column segment (panel rows x floats per element)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int S = a.nstages, D = a.lag, P = a.panels;

  // this CTA's column blocks
  const int blk0 = (int)(((long long)cta * a.nblk) / grid);
  const int blk1 = (int)(((long long)(cta + 1) * a.nblk) / grid);
  const int nb = blk1 - blk0;
  const int nch = (nb + a.sb - 1) / a.sb;               // stages (column chunks) per panel
  const long long col0 = (long long)blk0 * T_BOXC;

  SmemLayout<T, LPC> L{a.stage_bytes, S, 0};
  {
    const int nbmax = (a.nblk + grid - 1) / grid + 1;
    L.xcols = nbmax * T_BOXC;
  }
  uint8_t* stage_base = smem;
  T* xs = reinterpret_cast<T*>(smem + L.off_x());
  float4* ypart = reinterpret_cast<float4*>(smem + L.off_ypart());   // [2][NCW][LPC]
  float4* ysm = reinterpret_cast<float4*>(smem + L.off_ysm());       // [T_NY][LPC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bars());
  uint64_t* full = bars;              // [16]
  uint64_t* empty = bars + 16;        // [16]
  uint64_t* yp_full = bars + 32;      // [2]
  uint64_t* yp_free = bars + 34;      // [2]
  uint64_t* yready = bars + 36;       // [T_NY]
  int* abort_flag = a.ws.abort_flag;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], T_NCW); }
    for (int k = 0; k < 2; ++k) { mbar_init(&yp_full[k], T_NCW); mbar_init(&yp_free[k], 1); }
    for (int k = 0; k < T_NY; ++k) mbar_init(&yready[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stage x for this CTA's columns
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
            tma_load_2d(dst + (size_t)bx * BOX_BYTES, &tmap, t * PRF, (blk0 + bx0 + bx) * T_BOXC, &full[s]);
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
      const int b = t % T_NY;
      if (lane < LPC) {
        float4 s = ypart[(pb * T_NCW + 0) * LPC + lane];
#pragma unroll
        for (int w = 1; w < T_NCW; ++w) s = f4add(s, ypart[(pb * T_NCW + w) * LPC + lane]);
        a.ws.slots[((size_t)b * grid + cta) * 16 + lane] = s;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        atomicAdd(&a.ws.counter[b], 1u);
        mbar_arrive(&yp_free[pb]);
      }
    }
    return;
  }
  if (warp >= T_NCW + 2) {
    // ================================ gatherers =======================================
    const int gl = lane / LPC, r4 = lane % LPC;
    for (int t = warp - (T_NCW + 2); t < P; t += 2) {
      const int b = t % T_NY;
      const unsigned target = (unsigned)grid * (unsigned)(t / T_NY + 1);   // counters are zeroed before every launch
      unsigned ok = 1;
      if (lane == 0) {
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(&a.ws.counter[b]) - target) < 0) {
          if (++spins > T_SPIN_LIMIT || *((volatile int*)abort_flag)) { ok = 0; break; }
          __nanosleep(20);
        }
        if (!ok) *abort_flag = 1;
      }
      ok = __shfl_sync(0xffffffffu, ok, 0);
      __syncwarp();
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const float4* __restrict__ sl = a.ws.slots + (size_t)b * grid * 16 + r4;
        int c = gl;
        for (; c + 3 * NGRP < grid; c += 4 * NGRP) {   // 4 independent loads in flight, fixed order of adds
          float4 v0 = __ldcg(sl + (size_t)c * 16), v1 = __ldcg(sl + (size_t)(c + NGRP) * 16);
          float4 v2 = __ldcg(sl + (size_t)(c + 2 * NGRP) * 16), v3 = __ldcg(sl + (size_t)(c + 3 * NGRP) * 16);
          acc = f4add(f4add(f4add(f4add(acc, v0), v1), v2), v3);
        }
        for (; c < grid; c += NGRP) acc = f4add(acc, __ldcg(sl + (size_t)c * 16));
#pragma unroll
        for (int o = LPC; o < 32; o <<= 1) acc = f4add(acc, f4shfl_xor(acc, o));
      }
      if (lane < LPC) ysm[(t % T_NY) * LPC + lane] = acc;
      __syncwarp();
      if (lane == 0) mbar_arrive(&yready[t % T_NY]);
    }
    return;
  }

  // ================================== compute warps =====================================
  const int r4 = lane % LPC, cg = lane / LPC;
  T gacc[T_MAXJ][T_MAXI];
#pragma unroll
  for (int j = 0; j < T_MAXJ; ++j)
#pragma unroll
    for (int i = 0; i < T_MAXI; ++i) gacc[j][i] = T{};

  bool dead = false;
  for (int t = 0; t < P + D && !dead; ++t) {
    if (t < P) {
      // ---- phase 1: y_t partial over this warp's share of the CTA's columns ----
      float4 yacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < T_MAXJ; ++j) {
        if (j < nch && !dead) {
          const long long k = (long long)t * nch + j;
          const int s = (int)(k % S);
          if (!mbar_wait(&full[s], (unsigned)(k / S) & 1, abort_flag)) dead = true;
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
      if (dead) break;
#pragma unroll
      for (int o = LPC; o < 32; o <<= 1) yacc = f4add(yacc, f4shfl_xor(yacc, o));
      const int pb = t & 1;
      if (t >= 2 && !mbar_wait(&yp_free[pb], (unsigned)((t >> 1) - 1) & 1, abort_flag)) break;
      if (lane < LPC) ypart[(pb * T_NCW + warp) * LPC + lane] = yacc;
      __syncwarp();
      if (lane == 0) mbar_arrive(&yp_full[pb]);
    }
    if (t >= D) {
      // ---- phase 2: g += A_q' y_q from the same shared-memory tiles, then release them ----
      const int q = t - D;
      if (!mbar_wait(&yready[q % T_NY], (unsigned)(q / T_NY) & 1, abort_flag)) break;
      const float4 y4 = ysm[(q % T_NY) * LPC + r4];
#pragma unroll
      for (int j = 0; j < T_MAXJ; ++j) {
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
  for (int j = 0; j < T_MAXJ; ++j) {
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
  int lpc = 0, grid = 0, sb = 0, nstages = 0, lag = 0, stage_bytes = 0, panels = 0, nblk = 0;
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
static int32_t tma_configure(TmaPlan* p) {
  auto kern = normal_tma_kernel<T, LPC>;
  rls_ctx_s* c = p->ctx;
  rls_mat_s* A = p->A;
  constexpr int BOX_BYTES = T_BOXC * LPC * 16;
  const int grid = c->sm_count;
  const int nblk = (int)((A->n + T_BOXC - 1) / T_BOXC);
  const int nbmax = (nblk + grid - 1) / grid;                 // boxes of the widest CTA
  int dev_smem = 0;
  RLS_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  // boxes per stage: ~28 KB stages by default
  int sb = env_int("RLS_TMA_STAGE_KB", 28) * 1024 / BOX_BYTES;
  if (sb < 1) sb = 1;
  if (sb > nbmax) sb = nbmax;
  int nch = (nbmax + sb - 1) / sb;
  if (nch > T_MAXJ) { nch = T_MAXJ; sb = (nbmax + nch - 1) / nch; }
  // sweeps of the 16 compute warps over one stage
  if ((sb * T_BOXC * LPC + T_NCW * 32 - 1) / (T_NCW * 32) > T_MAXI) {
    rls_set_error("one-pass(TMA): %lld columns over %d SMs exceed the per-CTA tile budget", (long long)A->n, grid);
    return RLS_ERR_UNSUPPORTED;
  }
  const int stage_bytes = sb * BOX_BYTES;
  SmemLayout<T, LPC> L{stage_bytes, 0, (nbmax + 1) * T_BOXC};
  int S = 16;
  for (; S >= 2; --S) {
    L.nstages = S;
    if (L.total(0) <= (size_t)dev_smem) break;
  }
  const int lag = (S - 1) / nch - 1;
  if (S < 2 || lag < 1) {
    rls_set_error("one-pass(TMA): shared memory cannot hold two panels of %lld columns per SM", (long long)A->n / grid);
    return RLS_ERR_UNSUPPORTED;
  }
  L.nstages = S;
  p->grid = grid; p->sb = sb; p->nstages = S; p->stage_bytes = stage_bytes; p->nblk = nblk;
  p->lag = std::min(lag, env_int("RLS_TMA_LAG", T_NY - 2));
  if (p->lag < 1) p->lag = 1;
  p->smem_bytes = L.total(0);
  const int vec = Elem<T>::vec;
  const int PR = LPC * vec;
  p->panels = (int)((A->m + PR - 1) / PR);
  RLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
  int per_sm = 0;
  RLS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T_THREADS, p->smem_bytes));
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
  // 256-byte column segments reach full HBM read bandwidth, 128-byte ones ~83 % (tools/seg_bw.cu);
  // the smaller panel needs half the on-chip window, so it wins when the exchange latency dominates.
  p->lpc = env_int("RLS_TMA_LPC", 8);
  if (p->lpc != 8 && p->lpc != 16) p->lpc = 8;
  int32_t s;
  if (A->dtype == RLS_C32) s = p->lpc == 16 ? tma_configure<float2, 16>(p) : tma_configure<float2, 8>(p);
  else s = p->lpc == 16 ? tma_configure<float, 16>(p) : tma_configure<float, 8>(p);
  if (s != RLS_OK) { delete p; return s; }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  size_t o_slots = take(sizeof(float4) * T_NY * p->grid * 16);
  size_t o_cnt = take(sizeof(unsigned) * T_NY);
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
  TmaArgs a;
  a.x = x; a.g = g; a.ws = p->ws;
  // the arrival counters are zeroed by a stream-ordered memset before every launch, so a launch that is
  // gated off on the device (done() already true) leaves nothing behind for the next one
  RLS_CUDA(cudaMemsetAsync(p->ws.counter, 0, sizeof(unsigned) * T_NY, c->stream));
  a.m = p->A->m; a.n = p->A->n;
  a.panels = p->panels; a.nblk = p->nblk; a.sb = p->sb; a.nstages = p->nstages; a.lag = p->lag; a.stage_bytes = p->stage_bytes;
  a.gate = gate;
  void* args[] = {(void*)&p->tmap, (void*)&a};
  const void* fn;
  if (p->A->dtype == RLS_C32) fn = p->lpc == 16 ? (const void*)normal_tma_kernel<float2, 16> : (const void*)normal_tma_kernel<float2, 8>;
  else fn = p->lpc == 16 ? (const void*)normal_tma_kernel<float, 16> : (const void*)normal_tma_kernel<float, 8>;
  RLS_CUDA(cudaLaunchCooperativeKernel(fn, dim3(p->grid), dim3(T_THREADS), args, p->smem_bytes, c->stream));
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

void rls_tma_describe(TmaPlan* p, int* lpc, int* stages, int* lag, int* sb, size_t* smem) {
  *lpc = p->lpc; *stages = p->nstages; *lag = p->lag; *sb = p->sb; *smem = p->smem_bytes;
}
