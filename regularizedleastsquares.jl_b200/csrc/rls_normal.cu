// rls_normal.cu — the normal operator  res = A'(A x)  (mul!(res, AHA, x): FISTA.jl:152,
// POGM.jl:181, OptISTA.jl:182, CGNR.jl:151, cg! inside ADMM.jl:244, Utils.jl:278).
//
// Three forms:
//   TWOPASS : gemv_n then gemv_c                      -> 2 HBM sweeps over A
//   ONEPASS : the panel kernel below                  -> 1 HBM sweep over A (+1 L2 re-read)
//   GRAM    : dense G = A'A, one gemv over G          -> n^2 bytes (reference default form)
// With a communicator on the context, A is this rank's row shard and the n-vector of
// partial sums is combined with one allreduce.
//
// ---- the one-pass kernel ---------------------------------------------------------------
// g = sum over row panels p of A_p'(A_p x).  A panel (PR = LPC*VEC rows, all n columns)
// is streamed from HBM once for y_p = A_p x ("phase 1") and re-read D panels later from L2
// for g += A_p' y_p ("phase 2").  Every compute warp owns a fixed set of columns for the
// whole launch, so its partial g lives in registers and is written exactly once; the only
// cross-CTA exchange is the PR-vector y_p, reduced deterministically through per-CTA slots
// and a two-level ticket tree by a dedicated communication warp per CTA, and published
// with a release flag that phase 2 acquires (bounded spin; a time-out raises an abort flag
// instead of hanging the GPU).  Persistent cooperative grid: 2 CTAs x (8 compute + 1 comm)
// warps per SM.
#include "rls_common.cuh"

// TMA / shared-memory-resident implementation (rls_normal_tma.cu)
struct TmaPlan;
int32_t rls_tma_plan_create(rls_ctx_s* c, rls_mat_s* A, TmaPlan** out);
int32_t rls_tma_apply(TmaPlan* p, const void* x, void* g, const int* gate);
int32_t rls_tma_check_abort(TmaPlan* p);
void rls_tma_plan_destroy(TmaPlan* p);
void rls_tma_describe(TmaPlan* p, char* buf, int len);

namespace {

constexpr int OP_CWARPS = 8;                       // compute warps per CTA
constexpr int OP_THREADS = (OP_CWARPS + 1) * 32;   // + 1 communication warp
constexpr int OP_NBUF = 6;                         // y-panel ring (>= lag + 2)
constexpr int OP_GROUP = 24;                       // CTAs per first-level reduction group
constexpr unsigned OP_SPIN_LIMIT = 4000000u;       // ~ seconds; then abort instead of hanging

struct OnepassWs {
  float4* slots;      // [OP_NBUF][grid][LPCmax]   per-CTA partial y panels
  float4* gslots;     // [OP_NBUF][ngroups][LPCmax]
  float4* ypanel;     // [OP_NBUF][LPCmax]
  unsigned* gticket;  // [OP_NBUF][ngroups]
  unsigned* tticket;  // [OP_NBUF]
  unsigned* flag;     // [OP_NBUF]
  int* abort_flag;    // [1]
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ldg_plain(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_hint(const float4* p, unsigned long long pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4shfl_xor(float4 a, int o) {
  return make_float4(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o),
                     __shfl_xor_sync(0xffffffffu, a.z, o), __shfl_xor_sync(0xffffffffu, a.w, o));
}

// y-partial FMA: acc(4 floats = VEC rows) += a(16 bytes of a column) * x_col
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
// g-partial FMA: acc += conj(a) . y over the lane's VEC rows
template <typename T> __device__ __forceinline__ void fma_g(T& acc, float4 a, float4 y);
template <> __device__ __forceinline__ void fma_g<float>(float& acc, float4 a, float4 y) {
  acc = fmaf(a.x, y.x, acc); acc = fmaf(a.y, y.y, acc); acc = fmaf(a.z, y.z, acc); acc = fmaf(a.w, y.w, acc);
}
template <> __device__ __forceinline__ void fma_g<float2>(float2& acc, float4 a, float4 y) {
  acc.x = fmaf(a.x, y.x, acc.x); acc.x = fmaf(a.y, y.y, acc.x); acc.y = fmaf(a.x, y.y, acc.y); acc.y = fmaf(-a.y, y.x, acc.y);
  acc.x = fmaf(a.z, y.z, acc.x); acc.x = fmaf(a.w, y.w, acc.x); acc.y = fmaf(a.z, y.w, acc.y); acc.y = fmaf(-a.w, y.z, acc.y);
}

template <typename T>
__device__ __forceinline__ float4 mask_slice(float4 v, int64_t row0, int64_t m) {
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

struct OnepassArgs {
  const void* A;
  int64_t ld, m, n;
  const void* x;
  void* g;
  OnepassWs ws;
  unsigned tag_base;   // tags of this launch are tag_base + panel + 1
  int lag;             // D: phase 2 of panel p runs during step p + D
  int cols_per_warp;   // contiguous columns owned by each compute warp
  int use_hint;        // 1: phase-2 loads carry an L2 evict_first policy
  const int* gate;
};

// LPC lanes cooperate on one 16*LPC-byte column segment; a warp covers 32/LPC columns per
// load instruction; each lane keeps partial g for up to MAXC of its slot's columns.
template <typename T, int LPC, int MAXC>
__global__ void __launch_bounds__(OP_THREADS, 2) normal_onepass_kernel(OnepassArgs a) {
  if (a.gate && *a.gate) return;
  constexpr int VEC = Elem<T>::vec;
  constexpr int NSEG = 32 / LPC;
  constexpr int PR = LPC * VEC;  // panel rows
  constexpr int BK = MAXC < 8 ? MAXC : 8;
  __shared__ float4 ybuf[2][OP_CWARPS][LPC];
  __shared__ T xs[OP_CWARPS * NSEG * MAXC];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = a.m, n = a.n;
  const int P = (int)((m + PR - 1) / PR);
  const int grid = gridDim.x;
  const int ngroups = (grid + OP_GROUP - 1) / OP_GROUP;
  const OnepassWs& ws = a.ws;

  if (warp == OP_CWARPS) {
    // ------------------------- communication warp -------------------------------------
    const int group = blockIdx.x / OP_GROUP;
    const int gfirst = group * OP_GROUP;
    const int gsize = min(OP_GROUP, grid - gfirst);
    for (int t = 0; t < P; ++t) {
      __syncthreads();  // barrier A(t): compute warps have written ybuf[t&1]
      const int b = t % OP_NBUF;
      if (lane < LPC) {
        float4 s = ybuf[t & 1][0][lane];
#pragma unroll
        for (int w = 1; w < OP_CWARPS; ++w) s = f4add(s, ybuf[t & 1][w][lane]);
        ws.slots[((size_t)b * grid + blockIdx.x) * LPC + lane] = s;
      }
      __threadfence();
      __syncwarp();
      unsigned last = 0;
      if (lane == 0) last = (atomicAdd(&ws.gticket[b * ngroups + group], 1u) == (unsigned)gsize - 1) ? 1u : 0u;
      last = __shfl_sync(0xffffffffu, last, 0);
      if (!last) continue;
      __threadfence();
      if (lane < LPC) {
        float4 s = __ldcg(&ws.slots[((size_t)b * grid + gfirst) * LPC + lane]);
        for (int c = 1; c < gsize; ++c) s = f4add(s, __ldcg(&ws.slots[((size_t)b * grid + gfirst + c) * LPC + lane]));
        ws.gslots[((size_t)b * ngroups + group) * LPC + lane] = s;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        ws.gticket[b * ngroups + group] = 0u;
        last = (atomicAdd(&ws.tticket[b], 1u) == (unsigned)ngroups - 1) ? 1u : 0u;
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (!last) continue;
      __threadfence();
      if (lane < LPC) {
        float4 s = __ldcg(&ws.gslots[((size_t)b * ngroups) * LPC + lane]);
        for (int gidx = 1; gidx < ngroups; ++gidx) s = f4add(s, __ldcg(&ws.gslots[((size_t)b * ngroups + gidx) * LPC + lane]));
        ws.ypanel[(size_t)b * LPC + lane] = s;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        ws.tticket[b] = 0u;
        st_release(&ws.flag[b], a.tag_base + (unsigned)t + 1u);
      }
    }
    return;
  }

  // --------------------------- compute warps ------------------------------------------
  const int seg = lane / LPC, li = lane % LPC;
  const int64_t gw = (int64_t)blockIdx.x * OP_CWARPS + warp;
  const int64_t c0 = gw * a.cols_per_warp;
  const int64_t cend = min(n, c0 + a.cols_per_warp);
  const int64_t ldv = a.ld / VEC;
  const float4* __restrict__ Av = reinterpret_cast<const float4*>(a.A);
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);

  // stage this warp's x values in shared memory: xs[warp][k*NSEG + seg]
  T* xw = xs + warp * (NSEG * MAXC);
  for (int q = lane; q < NSEG * MAXC; q += 32) {
    int64_t col = c0 + q;
    xw[q] = (col < cend) ? x[col] : Elem<T>::zero();
  }
  __syncwarp();

  unsigned long long pol = 0;
  if (a.use_hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));

  T gacc[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) gacc[k] = Elem<T>::zero();

  const int D = a.lag;
  bool aborted = false;
  for (int t = 0; t < P + D; ++t) {
    if (t < P) {
      // ---- phase 1: y_t partial over this warp's columns (HBM stream) ----
      const int64_t row0 = (int64_t)t * PR + (int64_t)li * VEC;
      const bool rvalid = row0 < m;
      const float4* __restrict__ base = Av + (row0 / VEC);
      float4 yacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int kb = 0; kb < MAXC; kb += BK) {   // BK loads in flight per lane, then their FMAs
        float4 v[BK];
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          const int64_t col = c0 + (kb + k) * NSEG + seg;
          v[k] = (rvalid && col < cend) ? ldg_plain(base + col * ldv) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) fma_y<T>(yacc, v[k], xw[(kb + k) * NSEG + seg]);
      }
      if (rvalid && row0 + VEC > m) yacc = mask_slice<T>(yacc, row0, m);
#pragma unroll
      for (int o = LPC; o < 32; o <<= 1) yacc = f4add(yacc, f4shfl_xor(yacc, o));
      if (seg == 0) ybuf[t & 1][warp][li] = yacc;
      __syncthreads();  // barrier A(t)
    }
    if (t >= D && !aborted) {
      // ---- phase 2: g += A_q' y_q for q = t - D (L2 re-read) ----
      const int q = t - D;
      const int b = q % OP_NBUF;
      const unsigned want = a.tag_base + (unsigned)q + 1u;
      unsigned ok = 1;
      if (lane == 0) {
        unsigned spins = 0;
        while (ld_acquire(&ws.flag[b]) != want) {
          if (++spins > OP_SPIN_LIMIT || *((volatile int*)ws.abort_flag)) { ok = 0; break; }
          __nanosleep(64);
        }
        if (!ok) *ws.abort_flag = 1;
      }
      ok = __shfl_sync(0xffffffffu, ok, 0);
      __syncwarp();
      if (!ok) { aborted = true; continue; }
      const int64_t row0 = (int64_t)q * PR + (int64_t)li * VEC;
      const bool rvalid = row0 < m;
      float4 y4 = __ldcg(&ws.ypanel[(size_t)b * LPC + li]);
      if (!rvalid) y4 = make_float4(0.f, 0.f, 0.f, 0.f);
      else if (row0 + VEC > m) y4 = mask_slice<T>(y4, row0, m);
      const float4* __restrict__ base = Av + (row0 / VEC);
      const bool partial = rvalid && row0 + VEC > m;  // padding rows of a wrapped matrix may hold anything
#pragma unroll
      for (int kb = 0; kb < MAXC; kb += BK) {
        float4 v[BK];
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          const int64_t col = c0 + (kb + k) * NSEG + seg;
          const bool p = rvalid && col < cend;
          if (a.use_hint) v[k] = p ? ldg_hint(base + col * ldv, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
          else v[k] = p ? ldg_plain(base + col * ldv) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (partial) v[k] = mask_slice<T>(v[k], row0, m);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) fma_g<T>(gacc[kb + k], v[k], y4);
      }
    }
  }
  // ---- write g: reduce each column's partial over the LPC lanes of its slot ----
  T* __restrict__ g = reinterpret_cast<T*>(a.g);
#pragma unroll
  for (int k = 0; k < MAXC; ++k) {
    T s = gacc[k];
#pragma unroll
    for (int o = LPC / 2; o > 0; o >>= 1) {
      if constexpr (Elem<T>::is_complex) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
      } else {
        s += __shfl_xor_sync(0xffffffffu, s, o);
      }
    }
    const int64_t col = c0 + k * NSEG + seg;
    if (li == 0 && col < cend) g[col] = s;
  }
}

// simple tiled Gram build  G = A'A  (one-time precompute; CUDA cores for now — the
// tcgen05 split-precision GEMM is scheduled work)
constexpr int GT = 32;
template <typename T>
__global__ void __launch_bounds__(GT* GT) gram_kernel(const T* __restrict__ A, int64_t rs, int64_t cs, int64_t m, int64_t n, T* __restrict__ G) {
  __shared__ T sa[GT][GT + 1];
  __shared__ T sb[GT][GT + 1];
  const int tx = threadIdx.x % GT, ty = threadIdx.x / GT;
  const int64_t i0 = (int64_t)blockIdx.y * GT, j0 = (int64_t)blockIdx.x * GT;
  double accr = 0.0, acci = 0.0;
  for (int64_t k0 = 0; k0 < m; k0 += GT) {
    // coalesced along rows (k): thread (tx = k offset, ty = column offset)
    int64_t k = k0 + tx;
    sa[ty][tx] = (k < m && i0 + ty < n) ? A[k * rs + (i0 + ty) * cs] : Elem<T>::zero();
    sb[ty][tx] = (k < m && j0 + ty < n) ? A[k * rs + (j0 + ty) * cs] : Elem<T>::zero();
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < GT; ++kk) Elem<T>::dotc(sa[ty][kk], sb[tx][kk], accr, acci);
    __syncthreads();
  }
  const int64_t i = i0 + ty, j = j0 + tx;
  if (i < n && j < n) {
    if constexpr (Elem<T>::is_complex) G[i + j * n] = make_float2((float)accr, (float)acci);
    else G[i + j * n] = (float)accr;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct rls_normal_s {
  rls_ctx_s* ctx;
  rls_mat_s* A;      // borrowed
  int32_t form;      // resolved form
  rls_vec_s* ytmp;   // m-vector for the two-pass form
  rls_vec_s* gpart = nullptr;  // n-vector: this rank's partial A_i'(A_i x) before the allreduce
  rls_mat_s* G;      // Gram matrix
  bool own_G = true; // false when the caller supplied AHA as a matrix (FISTA(; AHA=...), FISTA.jl:55)
  int64_t n_ = 0;
  int32_t dtype_ = 0;
  // one-pass: TMA/shared-memory-resident kernel when supported, else the L2-lag kernel below
  TmaPlan* tma = nullptr;
  RowPlan* row = nullptr;   // row-major A: cluster kernel plan (owned by the matrix)
  bool gram_on_tensor_cores = false;
  TcBatchPlan* tc = nullptr;  // tensor-core plan of the multi-RHS apply (created on first use for a given K)
  int tc_K = 0;
  TcBatchPlan* tc_adj = nullptr;  // Gram form: plan on A for the K back-projections A'b_k of a multi-RHS init! (tc runs on G)
  int tc_adj_K = 0;
  OnepassWs ws{};
  void* ws_mem = nullptr;
  int op_grid = 0, op_lpc = 0, op_maxc = 0, op_cpw = 0, op_lag = 2, op_hint = 1;
  unsigned tag_next = 1;
  // matrix-free form (RLS_NORMAL_MATRIXFREE): res = AHA x through a function — a structured operator of rls_linop.cu or a
  // caller-supplied callback working on device pointers on the context's stream
  int32_t (*mf_apply)(void* user, const void* x, void* out, void* stream) = nullptr;
  void (*mf_release)(void* user) = nullptr;   // drops what the operator holds on `user` when it goes
  void* mf_user = nullptr;
  rls_vec_s* mf_tmp = nullptr;                // n-vector: result of a gated apply before the gated copy
  char mf_name[96] = {0};
  std::atomic<int> refs{1};   // the creator + every solver that borrows the operator
};

typedef void (*onepass_fn)(OnepassArgs);

template <typename T>
static onepass_fn pick_onepass(int lpc, int maxc) {
#define RLS_OP_CASE(L, M) if (lpc == L && maxc == M) return normal_onepass_kernel<T, L, M>;
  RLS_OP_CASE(8, 4) RLS_OP_CASE(8, 8) RLS_OP_CASE(8, 16)
  RLS_OP_CASE(16, 4) RLS_OP_CASE(16, 8) RLS_OP_CASE(16, 16)
  RLS_OP_CASE(32, 4) RLS_OP_CASE(32, 8) RLS_OP_CASE(32, 16)
#undef RLS_OP_CASE
  return nullptr;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// choose (LPC, MAXC) for n columns over `warps` compute warps; returns false when unsupported
static bool plan_onepass(rls_normal_s* op) {
  rls_ctx_s* c = op->ctx;
  rls_mat_s* A = op->A;
  onepass_fn probe = A->dtype == RLS_C32 ? pick_onepass<float2>(16, 16) : pick_onepass<float>(16, 16);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe, OP_THREADS, 0) != cudaSuccess || per_sm < 1) return false;
  if (per_sm > 2) per_sm = 2;
  int grid = c->sm_count * per_sm;
  int64_t warps = (int64_t)grid * OP_CWARPS;
  int cpw = (int)((A->n + warps - 1) / warps);
  if (cpw < 1) cpw = 1;
  // panel bytes = LPC*16*n ; keep (lag+1) panels comfortably inside L2
  int lag = env_int("RLS_ONEPASS_LAG", 2);
  if (lag < 1) lag = 1;
  if (lag > OP_NBUF - 2) lag = OP_NBUF - 2;
  const double l2 = (double)(c->l2_bytes ? c->l2_bytes : ((size_t)96 << 20));
  int lpc_force = env_int("RLS_ONEPASS_LPC", 0);
  int best_lpc = 0, best_maxc = 0;
  const int lpcs[3] = {32, 16, 8};
  for (int lpc : lpcs) {
    if (lpc_force && lpc != lpc_force) continue;
    int nseg = 32 / lpc;
    int ncs = (cpw + nseg - 1) / nseg;
    int maxc = ncs <= 4 ? 4 : (ncs <= 8 ? 8 : (ncs <= 16 ? 16 : 0));
    if (!maxc) continue;
    double panel = (double)lpc * 16.0 * (double)A->n;
    if (!lpc_force && panel * (lag + 1) > 0.45 * l2) continue;
    best_lpc = lpc;
    best_maxc = maxc;
    break;
  }
  if (!best_lpc) return false;
  op->op_grid = grid;
  op->op_lpc = best_lpc;
  op->op_maxc = best_maxc;
  op->op_cpw = cpw;
  op->op_lag = lag;
  op->op_hint = env_int("RLS_ONEPASS_HINT", 1);
  return true;
}

static int32_t alloc_onepass_ws(rls_normal_s* op) {
  const int grid = op->op_grid;
  const int ngroups = (grid + OP_GROUP - 1) / OP_GROUP;
  const size_t lpc = 32;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  size_t o_slots = take(sizeof(float4) * OP_NBUF * grid * lpc);
  size_t o_gslots = take(sizeof(float4) * OP_NBUF * ngroups * lpc);
  size_t o_yp = take(sizeof(float4) * OP_NBUF * lpc);
  size_t o_gt = take(sizeof(unsigned) * OP_NBUF * ngroups);
  size_t o_tt = take(sizeof(unsigned) * OP_NBUF);
  size_t o_fl = take(sizeof(unsigned) * OP_NBUF);
  size_t o_ab = take(sizeof(int));
  RLS_CUDA(cudaMalloc(&op->ws_mem, off));
  RLS_CUDA(cudaMemsetAsync(op->ws_mem, 0, off, op->ctx->stream));
  char* b = (char*)op->ws_mem;
  op->ws.slots = (float4*)(b + o_slots);
  op->ws.gslots = (float4*)(b + o_gslots);
  op->ws.ypanel = (float4*)(b + o_yp);
  op->ws.gticket = (unsigned*)(b + o_gt);
  op->ws.tticket = (unsigned*)(b + o_tt);
  op->ws.flag = (unsigned*)(b + o_fl);
  op->ws.abort_flag = (int*)(b + o_ab);
  return RLS_OK;
}

static int32_t build_gram(rls_normal_s* op) {
  RlsNvtxRange nvtx("rls: A'*A (Gram build)");
  rls_mat_s* A = op->A;
  rls_ctx_s* c = op->ctx;
  RLS_TRY(rls_mat_create_layout(c, A->dtype, A->n, A->n, nullptr, A->n, RLS_LAYOUT_COLMAJOR, &op->G));
  // dense n x n with ld == n (padded ld would break the symmetric indexing in gram_kernel)
  RLS_CHECK_ARG(op->G->ld == A->n, "Gram form needs n to be a multiple of %d", A->dtype == RLS_C32 ? 2 : 4);
  // tensor cores (FP32-accurate split-precision tcgen05 GEMM) when A is row-major; RLS_GRAM_CUDA_CORES=1 keeps the
  // tiled CUDA-core kernel (the comparison baseline)
  const char* cc = getenv("RLS_GRAM_CUDA_CORES");
  if (A->layout == RLS_LAYOUT_ROWMAJOR && !(cc && atoi(cc) != 0)) {
    int32_t st = rls_tc_gram(A, op->G);
    if (st == RLS_OK) {
      op->gram_on_tensor_cores = true;
      if (c->nranks > 1) RLS_TRY(rls_allreduce_raw(c, op->G->d, A->n * A->n * (A->dtype == RLS_C32 ? 2 : 1)));
      return RLS_OK;
    }
    if (st != RLS_ERR_UNSUPPORTED && st != RLS_ERR_NOMEM) return st;
  }
  dim3 grid((unsigned)((A->n + GT - 1) / GT), (unsigned)((A->n + GT - 1) / GT));
  const int64_t rs = A->layout == RLS_LAYOUT_ROWMAJOR ? A->ld : 1, cs = A->layout == RLS_LAYOUT_ROWMAJOR ? 1 : A->ld;
  if (A->dtype == RLS_C32)
    gram_kernel<float2><<<grid, GT * GT, 0, c->stream>>>((const float2*)A->d, rs, cs, A->m, A->n, (float2*)op->G->d);
  else
    gram_kernel<float><<<grid, GT * GT, 0, c->stream>>>((const float*)A->d, rs, cs, A->m, A->n, (float*)op->G->d);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  if (c->nranks > 1) RLS_TRY(rls_allreduce_raw(c, op->G->d, A->n * A->n * (A->dtype == RLS_C32 ? 2 : 1)));
  return RLS_OK;
}

extern "C" int32_t rls_normal_create(rls_mat_t A, int32_t form, rls_normal_t* out) {
  RLS_CHECK_ARG(A && out, "NULL argument");
  RLS_CHECK_ARG(form >= RLS_NORMAL_TWOPASS && form <= RLS_NORMAL_AUTO, "unknown normal-operator form %d", form);
  RlsDeviceGuard g(A->ctx->device);
  rls_normal_s* op = new rls_normal_s();
  op->ctx = A->ctx;
  op->A = A;
  op->ytmp = nullptr;
  op->G = nullptr;
  op->n_ = A->n;
  op->dtype_ = A->dtype;
  const double bytes = (double)A->m * (double)A->n * (double)rls_elem_size(A->dtype);
  if (A->layout == RLS_LAYOUT_ROWMAJOR && form != RLS_NORMAL_GRAM) {
    // rows contiguous: the streaming cluster kernel of rls_rowstream.cu sweeps A once (ONEPASS, also what AUTO
    // resolves to) or runs as two sweeps y = A x, g = A' y (TWOPASS, kept for comparison)
    op->row = rls_mat_rowplan(A);
    if (!op->row) { delete op; return RLS_ERR_UNSUPPORTED; }
    if (form == RLS_NORMAL_AUTO) form = RLS_NORMAL_ONEPASS;
  }
  if (form == RLS_NORMAL_AUTO) {
    // Measured on B200 (profiles/): the two-sweep form runs at the HBM read roofline (7.3 TB/s), the
    // one-pass panel kernels currently reach the same wall time with half the DRAM traffic.  AUTO
    // therefore keeps the two-sweep form unless RLS_AUTO_ONEPASS=1 asks for the panel kernel on
    // matrices far larger than L2.
    const char* want = getenv("RLS_AUTO_ONEPASS");
    const bool big = bytes >= 4.0 * (double)(A->ctx->l2_bytes ? A->ctx->l2_bytes : ((size_t)64 << 20));
    form = (want && atoi(want) != 0 && big) ? RLS_NORMAL_ONEPASS : RLS_NORMAL_TWOPASS;
  }
  if (form == RLS_NORMAL_ONEPASS && !op->row) {
    const char* impl = getenv("RLS_ONEPASS_IMPL");
    const bool want_l2 = impl && strcmp(impl, "l2") == 0;
    if (!want_l2 && rls_tma_plan_create(A->ctx, A, &op->tma) != RLS_OK) {
      op->tma = nullptr;
      if (getenv("RLS_VERBOSE")) fprintf(stderr, "[rls] TMA one-pass unavailable, using the L2-lag kernel: %s\n", rls_last_error());
    }
    if (!op->tma) {
      if (!plan_onepass(op)) {
        delete op;
        rls_set_error("one-pass normal operator does not support n=%lld on this device", (long long)A->n);
        return RLS_ERR_UNSUPPORTED;
      }
      int32_t s = alloc_onepass_ws(op);
      if (s != RLS_OK) { delete op; return s; }
    }
  }
  op->form = form;
  rls_mat_retain(A);          // from here on rls_normal_destroy / release undoes everything
  rls_ctx_retain(op->ctx);
  if (form == RLS_NORMAL_GRAM) {
    int32_t s = build_gram(op);
    if (s != RLS_OK) { rls_normal_destroy(op); return s; }
  } else {
    int32_t s = rls_vec_create_internal(A->ctx, A->dtype, A->m, &op->ytmp);
    if (s != RLS_OK) { rls_normal_destroy(op); return s; }
  }
  *out = op;
  return RLS_OK;
}

void rls_normal_retain(rls_normal_t op) { if (op) op->refs.fetch_add(1); }

void rls_normal_release(rls_normal_t op) {
  if (!op || op->refs.fetch_sub(1) != 1) return;
  rls_ctx_s* c = op->ctx;
  rls_mat_s* A = op->A;
  rls_mat_s* Gborrowed = (op->G && !op->own_G) ? op->G : nullptr;
  {
    RlsDeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    if (op->ytmp) rls_vec_destroy(op->ytmp);
    if (op->gpart) rls_vec_destroy(op->gpart);
    if (op->G && op->own_G) rls_mat_destroy(op->G);
    if (op->ws_mem) cudaFree(op->ws_mem);
    if (op->tma) rls_tma_plan_destroy(op->tma);
    if (op->tc) rls_tc_batch_destroy(op->tc);
    if (op->tc_adj) rls_tc_batch_destroy(op->tc_adj);
    if (op->mf_tmp) rls_vec_destroy(op->mf_tmp);
    if (op->mf_release) op->mf_release(op->mf_user);
    delete op;
  }
  rls_mat_release(A);
  rls_mat_release(Gborrowed);
  rls_ctx_release(c);
}

extern "C" int32_t rls_normal_destroy(rls_normal_t op) {
  rls_normal_release(op);   // solvers that borrow the operator keep it until they go
  return RLS_OK;
}

// human-readable description of the kernel plan behind this operator (diagnostics / bench config)
extern "C" int32_t rls_normal_describe(rls_normal_t op, char* buf, int32_t len) {
  RLS_CHECK_ARG(op && buf && len > 0, "NULL argument");
  if (op->form == RLS_NORMAL_MATRIXFREE) snprintf(buf, len, "matrix-free: %s", op->mf_name);
  else if (op->form == RLS_NORMAL_ONEPASS && op->row) rls_rowpass_describe(op->row, buf, len);
  else if (op->row) snprintf(buf, len, "twopass/rowmajor: gemv_n + gemv_c (cluster kernels)");
  else if (op->form == RLS_NORMAL_ONEPASS && op->tma) rls_tma_describe(op->tma, buf, len);
  else if (op->form == RLS_NORMAL_ONEPASS)
    snprintf(buf, len, "onepass/l2: grid=%d lanes/column=%d cols/lane<=%d cols/warp=%d lag=%d hint=%d", op->op_grid, op->op_lpc,
             op->op_maxc, op->op_cpw, op->op_lag, op->op_hint);
  else if (op->form == RLS_NORMAL_GRAM)
    snprintf(buf, len, "gram: dense %lldx%lld%s", (long long)op->n_, (long long)op->n_,
             op->gram_on_tensor_cores ? " (built on tensor cores: tcgen05 kind::tf32 x3 split)" : "");
  else snprintf(buf, len, "twopass: gemv_n + gemv_c");
  return RLS_OK;
}

extern "C" int32_t rls_normal_form(rls_normal_t op, int32_t* form) {
  RLS_CHECK_ARG(op && form, "NULL argument");
  *form = op->form;
  return RLS_OK;
}

// AHA supplied directly as a dense n x n matrix (createLinearSolver(S; AHA=G), FISTA.jl:55)
extern "C" int32_t rls_normal_from_gram(rls_mat_t G, rls_normal_t* out) {
  RLS_CHECK_ARG(G && out, "NULL argument");
  RLS_CHECK_ARG(G->m == G->n, "AHA must be square, got %lldx%lld", (long long)G->m, (long long)G->n);
  rls_normal_s* op = new rls_normal_s();
  op->ctx = G->ctx;
  op->A = nullptr;
  op->ytmp = nullptr;
  op->G = G;
  op->own_G = false;
  op->form = RLS_NORMAL_GRAM;
  op->n_ = G->n;
  op->dtype_ = G->dtype;
  rls_mat_retain(G);
  rls_ctx_retain(op->ctx);
  *out = op;
  return RLS_OK;
}

// matrix-free AHA: apply(user, x, out, stream) must enqueue out = AHA x on `stream` (device pointers, n elements)
int32_t rls_normal_from_function(rls_ctx_s* ctx, int32_t dtype, int64_t n, int32_t (*apply)(void*, const void*, void*, void*),
                                 void (*release)(void*), void* user, const char* name, rls_normal_t* out) {
  RLS_CHECK_ARG(ctx && apply && out && n >= 0, "bad argument");
  RLS_CHECK_ARG(dtype == RLS_F32 || dtype == RLS_C32, "unsupported element type %d", dtype);
  rls_normal_s* op = new rls_normal_s();
  op->ctx = ctx;
  op->A = nullptr;
  op->ytmp = nullptr;
  op->G = nullptr;
  op->form = RLS_NORMAL_MATRIXFREE;
  op->n_ = n;
  op->dtype_ = dtype;
  op->mf_apply = apply;
  op->mf_release = release;
  op->mf_user = user;
  snprintf(op->mf_name, sizeof(op->mf_name), "%s", name ? name : "callback");
  rls_ctx_retain(ctx);
  RlsDeviceGuard g(ctx->device);
  int32_t st = rls_vec_create_internal(ctx, dtype, n, &op->mf_tmp);
  if (st != RLS_OK) { op->mf_release = nullptr; rls_normal_release(op); return st; }
  *out = op;
  return RLS_OK;
}

// createLinearSolver(S, A_matrix_free): AHA as a C callback (a Julia @cfunction around any LinearOperator — FFTW/CUFFT
// plans, NFFT, Radon, ...: docs/src/literate/howto/normal_operator.jl, examples/computed_tomography.jl:23).  The callback
// receives DEVICE pointers and the CUDA stream it must enqueue its work on; it returns 0 on success.
extern "C" int32_t rls_normal_from_callback(rls_ctx_t ctx, int32_t dtype, int64_t n, rls_apply_fn fn, void* user, rls_normal_t* out) {
  RLS_CHECK_ARG(ctx && fn && out, "NULL argument");
  return rls_normal_from_function(ctx, dtype, n, fn, nullptr, user, "callback", out);
}

int32_t rls_normal_shape(rls_normal_t op, int64_t* n, int32_t* dtype) {
  if (n) *n = op->n_;
  if (dtype) *dtype = op->dtype_;
  return RLS_OK;
}

rls_mat_s* rls_normal_matrix(rls_normal_t op) { return op->A; }

// May the applies of this operator be recorded into a CUDA graph?  Yes for the column-major two-sweep form and the Gram
// form on one rank (plain kernels whose arguments do not change from launch to launch); no for the L2-lag one-pass kernel
// (its exchange tags advance per launch), callbacks (anything may happen inside) and row shards (collectives).
bool rls_normal_graph_safe(rls_normal_t op) {
  if (!op || op->ctx->nranks > 1 || op->row || op->tma) return false;
  if (op->form == RLS_NORMAL_GRAM) return op->G != nullptr;
  return op->form == RLS_NORMAL_TWOPASS && op->A && op->A->layout == RLS_LAYOUT_COLMAJOR;
}

rls_ctx_s* rls_normal_ctx(rls_normal_t op) { return op->ctx; }

static int32_t launch_onepass(rls_normal_s* op, const void* x, void* res, const int* gate) {
  rls_ctx_s* c = op->ctx;
  rls_mat_s* A = op->A;
  const int vec = A->dtype == RLS_C32 ? 2 : 4;
  const int PR = op->op_lpc * vec;
  const unsigned P = (unsigned)((A->m + PR - 1) / PR);
  if (op->tag_next > 0xffffffffu - (P + 2)) {  // tag wrap: restart the epoch
    RLS_CUDA(cudaMemsetAsync(op->ws.flag, 0, sizeof(unsigned) * OP_NBUF, c->stream));
    op->tag_next = 1;
  }
  OnepassArgs a;
  a.A = A->d; a.ld = A->ld; a.m = A->m; a.n = A->n;
  a.x = x; a.g = res; a.ws = op->ws;
  a.tag_base = op->tag_next;
  a.lag = op->op_lag;
  a.cols_per_warp = op->op_cpw;
  a.use_hint = op->op_hint;
  a.gate = gate;
  op->tag_next += P + 1;
  onepass_fn fn = A->dtype == RLS_C32 ? pick_onepass<float2>(op->op_lpc, op->op_maxc) : pick_onepass<float>(op->op_lpc, op->op_maxc);
  RLS_CHECK_ARG(fn, "no one-pass kernel for LPC=%d MAXC=%d", op->op_lpc, op->op_maxc);
  void* args[] = {&a};
  RLS_CUDA(cudaLaunchCooperativeKernel((const void*)fn, dim3(op->op_grid), dim3(OP_THREADS), args, 0, c->stream));
  c->launches++;
  return RLS_OK;
}

__global__ void gated_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t nfloats, const int* gate) {
  if (gate && *gate) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nfloats; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int32_t rls_normal_apply_raw(rls_normal_t op, const void* x, void* res, const int* gate) {
  RlsNvtxRange nvtx("rls: mul!(res, AHA, x)");
  rls_ctx_s* c = op->ctx;
  if (op->form == RLS_NORMAL_GRAM) return rls_gemv_n_raw(op->G, x, res, gate);  // G already summed over ranks
  if (op->form == RLS_NORMAL_MATRIXFREE) {
    // the function knows nothing of the solver's done() gate: a gated apply lands in a scratch vector and is copied
    // into res by a gated kernel, so a finished solve is never disturbed.  The operator is global (not row-sharded).
    void* out = gate ? op->mf_tmp->d : res;
    const int32_t st = op->mf_apply(op->mf_user, x, out, (void*)c->stream);
    if (st != RLS_OK) {
      if (st > RLS_ERR_NOMEM || st < 0) { rls_set_error("matrix-free normal operator: the callback returned %d", (int)st); return RLS_ERR_INVALID; }
      return st;
    }
    if (gate) {
      gated_copy_kernel<<<c->sm_count, 256, 0, c->stream>>>((float*)res, (const float*)out, op->n_ * (op->dtype_ == RLS_C32 ? 2 : 1), gate);
      c->launches++;
      RLS_CUDA(cudaGetLastError());
    }
    return RLS_OK;
  }
  // row-sharded: kernels write this rank's partial into gpart, one sum-allreduce of the
  // n-vector over NVLink, then a (gated) copy into res.  A gated-off launch still joins the
  // collective so that ranks stay in lock-step, but never touches res.
  void* out = res;
  const int64_t nf_all = op->n_ * (op->dtype_ == RLS_C32 ? 2 : 1);
  if (c->nranks > 1) {
    // one-pass kernel on rows + peer-memory exchange: cluster partials -> (sum, NVLink all-reduce, gated write) in ONE kernel
    if (op->row && op->form == RLS_NORMAL_ONEPASS && op->A->m > 0 && rls_p2p_available(c, nf_all)) {
      const float* gp = nullptr; int64_t gs = 0; int ncl = 0;
      RLS_TRY(rls_rowpass_normal_deferred(op->row, x, nullptr, nullptr, nullptr, gate, &gp, &gs, &ncl));
      return rls_p2p_allreduce(c, gp, gs, ncl, nf_all, (float*)res, gate);
    }
    if (!op->gpart) RLS_TRY(rls_vec_create_internal(c, op->dtype_, op->n_, &op->gpart));
    out = op->gpart->d;
  }
  switch (op->form) {
    case RLS_NORMAL_TWOPASS:
      RLS_TRY(rls_gemv_n_raw(op->A, x, op->ytmp->d, gate));
      RLS_TRY(rls_gemv_c_raw(op->A, op->ytmp->d, out, gate));
      break;
    case RLS_NORMAL_ONEPASS:
      if (op->A->m == 0) { RLS_CUDA(cudaMemsetAsync(out, 0, op->A->n * rls_elem_size(op->A->dtype), c->stream)); break; }
      if (op->row) RLS_TRY(rls_rowpass_normal(op->row, x, out, gate));
      else if (op->tma) RLS_TRY(rls_tma_apply(op->tma, x, out, gate));
      else RLS_TRY(launch_onepass(op, x, out, gate));
      break;
    default:
      rls_set_error("bad normal-operator form");
      return RLS_ERR_INVALID;
  }
  if (c->nranks > 1) {
    const int64_t nf = nf_all;
    if (rls_p2p_available(c, nf)) return rls_p2p_allreduce(c, (const float*)out, 0, 1, nf, (float*)res, gate);
    RLS_TRY(rls_allreduce_raw(c, out, nf));
    gated_copy_kernel<<<c->sm_count, 256, 0, c->stream>>>((float*)res, (const float*)out, nf, gate);
    c->launches++;
    RLS_CUDA(cudaGetLastError());
  }
  return RLS_OK;
}

// res_k = AHA x_k for K right-hand sides.  Lazy forms on a row-major A: two tensor-core GEMMs that read A once
// each; Gram form: one tensor-core GEMM over G (rls_tc.cu); otherwise K single applies.
int32_t rls_normal_apply_batch_raw(rls_normal_t op, int K, const void* const* xs, void* const* outs, const int* const* gates) {
  const bool want = rls_env_flag("RLS_BATCH_TENSOR_CORES", true);
  if (want && op->form != RLS_NORMAL_GRAM && op->A && rls_tc_batch_supported(op->A, K)) {
    if (op->tc && op->tc_K != K) { rls_tc_batch_destroy(op->tc); op->tc = nullptr; }
    if (!op->tc) {
      int32_t st = rls_tc_batch_create(op->A, K, &op->tc);
      if (st != RLS_OK) op->tc = nullptr;
      op->tc_K = K;
    }
    if (op->tc) return rls_tc_batch_apply(op->tc, xs, outs, gates);
  }
  // Gram form (the reference's default AHA): one GEMM over G instead of K gemvs that re-read G K times
  if (want && op->form == RLS_NORMAL_GRAM && op->G && rls_env_flag("RLS_GRAM_BATCH_TENSOR_CORES", true) && rls_tc_gram_batch_supported(op->G, K)) {
    if (op->tc && op->tc_K != K) { rls_tc_batch_destroy(op->tc); op->tc = nullptr; }
    if (!op->tc) {
      int32_t st = rls_tc_gram_batch_create(op->G, K, &op->tc);
      if (st != RLS_OK) op->tc = nullptr;
      op->tc_K = K;
    }
    if (op->tc) return rls_tc_gram_batch_apply(op->tc, xs, outs, gates);
  }
  for (int k = 0; k < K; ++k) RLS_TRY(rls_normal_apply_raw(op, xs[k], outs[k], gates ? gates[k] : nullptr));
  return RLS_OK;
}

// outs_k = A' b_k for the K columns of a multi-RHS init!, as one tensor-core GEMM when the batch plan applies;
// *done = false leaves the back-projections to the caller (per-column gemv_c)
int32_t rls_normal_adjoint_batch_raw(rls_normal_t op, int K, const void* const* bs, void* const* outs, bool* done) {
  *done = false;
  if (!rls_env_flag("RLS_BATCH_TENSOR_CORES", true) || !op->A || !rls_tc_batch_supported(op->A, K)) return RLS_OK;
  // lazy forms share the plan of the batched apply; the Gram form's apply plan lives on G, so it keeps a second one on A
  TcBatchPlan** plan = op->form == RLS_NORMAL_GRAM ? &op->tc_adj : &op->tc;
  int* plan_K = op->form == RLS_NORMAL_GRAM ? &op->tc_adj_K : &op->tc_K;
  if (*plan && *plan_K != K) { rls_tc_batch_destroy(*plan); *plan = nullptr; }
  if (!*plan) {
    if (rls_tc_batch_create(op->A, K, plan) != RLS_OK) { *plan = nullptr; return RLS_OK; }
    *plan_K = K;
  }
  RLS_TRY(rls_tc_batch_adjoint(*plan, bs, outs));
  *done = true;
  return RLS_OK;
}

extern "C" int32_t rls_normal_apply_batch(rls_normal_t op, int32_t K, const rls_vec_t* xs, const rls_vec_t* outs) {
  RLS_CHECK_ARG(op && xs && outs && K >= 1, "bad argument");
  std::vector<const void*> xp(K);
  std::vector<void*> op_(K);
  for (int k = 0; k < K; ++k) {
    RLS_CHECK_ARG(xs[k] && outs[k], "NULL vector");
    RLS_CHECK_ARG(xs[k]->len == op->n_ && outs[k]->len == op->n_ && xs[k]->dtype == op->dtype_ && outs[k]->dtype == op->dtype_,
                  "normal_apply_batch: shape / dtype mismatch in column %d", k);
    RLS_CHECK_ARG(xs[k]->d != outs[k]->d, "normal_apply_batch: x and res must not alias");
    xp[k] = xs[k]->d; op_[k] = outs[k]->d;
  }
  RlsDeviceGuard g(op->ctx->device);
  RLS_TRY(rls_normal_apply_batch_raw(op, K, xp.data(), op_.data(), nullptr));
  if (op->tc) return rls_tc_batch_check_abort(op->tc);
  return RLS_OK;
}

// diagnostics (tools/tc_check.py): internal operands of the last batched apply
extern "C" int32_t rls_normal_batch_debug(rls_normal_t op, int32_t which, float* host, int64_t nfloats) {
  RLS_CHECK_ARG(op && op->tc && host, "no tensor-core batch plan");
  return rls_tc_batch_debug(op->tc, which, host, nfloats);
}

int32_t rls_normal_apply_deferred_raw(rls_normal_t op, const void* x, const float* xold, const float* th_old, const float* th,
                                      const int* gate, NormalPartials* np) {
  np->gpart = nullptr; np->gstride = 0; np->ncl = 0;
  if (!rls_env_flag("RLS_FUSE_ITERATION", true) || !op->row || op->form != RLS_NORMAL_ONEPASS || op->ctx->nranks > 1 || !op->A || op->A->m == 0 || op->A->n == 0)
    return RLS_OK;
  return rls_rowpass_normal_deferred(op->row, x, xold, th_old, th, gate, &np->gpart, &np->gstride, &np->ncl);
}

int32_t rls_normal_check_abort(rls_normal_t op) {
  RLS_TRY(rls_p2p_check_abort(op->ctx));
  if (op->tc) RLS_TRY(rls_tc_batch_check_abort(op->tc));          // multi-RHS plans: a timed-out barrier ends the GEMM, not the process
  if (op->tc_adj) RLS_TRY(rls_tc_batch_check_abort(op->tc_adj));
  if (op->row) return rls_rowpass_check_abort(op->row);
  if (op->form != RLS_NORMAL_ONEPASS) return RLS_OK;
  if (op->tma) return rls_tma_check_abort(op->tma);
  int flag = 0;
  RLS_CUDA(cudaMemcpyAsync(&flag, op->ws.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, op->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(op->ctx->stream));
  if (flag) {
    rls_set_error("one-pass normal operator timed out waiting for a panel reduction (abort flag set)");
    return RLS_ERR_CUDA;
  }
  return RLS_OK;
}

extern "C" int32_t rls_normal_apply(rls_normal_t op, rls_vec_t x, rls_vec_t res) {
  RLS_CHECK_ARG(op && x && res, "NULL argument");
  RLS_CHECK_ARG(x->len == op->n_ && res->len == op->n_, "normal_apply: operator is %lldx%lld", (long long)op->n_, (long long)op->n_);
  RLS_CHECK_ARG(x->dtype == op->dtype_ && res->dtype == op->dtype_, "normal_apply: dtype mismatch");
  RLS_CHECK_ARG(x->d != res->d, "normal_apply: x and res must not alias");
  RlsDeviceGuard g(op->ctx->device);
  return rls_normal_apply_raw(op, x->d, res->d, nullptr);
}

// ---- power iterations (Utils.jl:262-287) -------------------------------------------------
template <typename T>
__global__ void scale_real_kernel(T* x, int64_t n, float s) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = Elem<T>::divr(x[i], s);
}

extern "C" int32_t rls_power_iterations(rls_normal_t op, rls_vec_t b0, double rtol, int32_t maxiter, double* lambda_max) {
  RLS_CHECK_ARG(op && b0 && lambda_max, "NULL argument");
  RLS_CHECK_ARG(b0->len == op->n_ && b0->dtype == op->dtype_, "power_iterations: start vector shape/dtype mismatch");
  rls_ctx_s* c = op->ctx;
  RlsDeviceGuard g(c->device);
  rls_vec_s *b = nullptr, *bold = nullptr;
  RLS_TRY(rls_vec_create_internal(c, b0->dtype, b0->len, &b));
  RLS_TRY(rls_vec_create_internal(c, b0->dtype, b0->len, &bold));
  RLS_TRY(rls_vec_copy(b, b0));
  double lam = INFINITY;
  int32_t status = RLS_OK;
  for (int it = 0; it < maxiter; ++it) {
    double nrm = 0;
    if ((status = rls_vec_nrm2(b, &nrm)) != RLS_OK) break;
    int grid = (int)((b->len + 255) / 256);
    if (b->dtype == RLS_C32) scale_real_kernel<float2><<<grid, 256, 0, c->stream>>>((float2*)b->d, b->len, (float)nrm);
    else scale_real_kernel<float><<<grid, 256, 0, c->stream>>>((float*)b->d, b->len, (float)nrm);
    c->launches++;
    std::swap(b, bold);
    if ((status = rls_normal_apply_raw(op, bold->d, b->d, nullptr)) != RLS_OK) break;
    double d[2];
    if ((status = rls_vec_dot(bold, b, d)) != RLS_OK) break;
    double lam_old = lam;
    lam = sqrt(d[0] * d[0] + d[1] * d[1]);
    if (fabs(lam / lam_old - 1.0) < rtol) break;
  }
  rls_vec_destroy(b);
  rls_vec_destroy(bold);
  if (status == RLS_OK) status = rls_normal_check_abort(op);
  *lambda_max = lam;
  return status;
}
