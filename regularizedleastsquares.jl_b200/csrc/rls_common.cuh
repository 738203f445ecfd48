// rls_common.cuh — internal types and device helpers shared by all translation units
// of librls_b200.so (sm_100a only; no CPU fallback, no multi-backend dispatch).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/rls_b200.h"

// ------------------------------------------------------------------------------------
// error plumbing (never abort / throw across the ABI)
// ------------------------------------------------------------------------------------
void rls_set_error(const char* fmt, ...);

#define RLS_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      rls_set_error(__VA_ARGS__);                \
      return RLS_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define RLS_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      rls_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,   \
                    cudaGetErrorString(_e));                                                  \
      return RLS_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define RLS_TRY(expr)              \
  do {                             \
    int32_t _s = (expr);           \
    if (_s != RLS_OK) return _s;   \
  } while (0)

// ------------------------------------------------------------------------------------
// host-side handle structs
// ------------------------------------------------------------------------------------
constexpr int RLS_MAX_PEERS = 16;          // ranks of the one-shot NVLink all-reduce (one box)
constexpr int RLS_MAX_RED_BLOCKS = 4096;  // upper bound on grid size of reducing kernels
constexpr int RLS_MAX_ACC = 8;            // accumulators per reducing kernel

struct rls_ctx_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  size_t l2_bytes = 0, hbm_bytes = 0;
  int64_t launches = 0;
  // reduction scratch (single stream => kernels never overlap)
  double* red_partials = nullptr;  // [RLS_MAX_RED_BLOCKS * RLS_MAX_ACC]
  unsigned* red_ticket = nullptr;  // self-resetting ticket
  double* red_out = nullptr;       // [RLS_MAX_ACC] device result slots for host-returning calls
  double* red_out_host = nullptr;  // pinned mirror
  // gemv scratch (partials of the split-column y = A x), grown on demand
  void* gemv_scratch = nullptr;
  size_t gemv_scratch_bytes = 0;
  unsigned* gemv_tickets = nullptr;  // [4096]
  // singular-value thresholding scratch (partial Gram matrices and the q x q maps W, rls_svt.cu), grown on demand
  void* svt_scratch = nullptr;
  size_t svt_scratch_bytes = 0;
  uint64_t llr_calls = 0;            // LLR randshift: prox! calls so far (counter of the shift generator)
  // L2 flush
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // communicator (one rank per process)
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  // one-shot all-reduce over NVLink peer memory (rls_p2p.cu): exchange buffers shared through CUDA IPC
  float* peer_own = nullptr;
  float* peer_base[RLS_MAX_PEERS] = {};
  int64_t peer_cap = 0;
  int* peer_abort = nullptr;
  bool peer_ready = false;
  // Handles are freed by garbage collectors (Julia finalizers, Python weakrefs) in ARBITRARY order: every object that
  // points to a context / matrix / operator holds a reference, and rls_*_destroy only drops the caller's one — the
  // memory goes when the last holder has gone.
  std::atomic<int> refs{1};
};

struct rls_vec_s {
  rls_ctx_s* ctx;
  int32_t dtype;
  int64_t len;
  void* d;
  bool owned;
};

struct rls_mat_s {
  rls_ctx_s* ctx;
  int32_t dtype;
  int64_t m, n, ld;   // ld: element stride between columns (col-major) or between rows (row-major)
  void* d;
  bool owned;
  int32_t layout = RLS_LAYOUT_COLMAJOR;
  struct RowPlan* rowplan = nullptr;  // row-major matrices: kernel plan shared by gemv / normal operator
  std::atomic<int> refs{1};           // the creator + every operator / solver built on it
};

static inline size_t rls_elem_size(int32_t dtype) { return dtype == RLS_C32 ? 8 : 4; }

// NVTX ranges (header-only NVTX3: a no-op unless a tool is attached): one range per phase of the reference's call
// stack — init!, iterate, solve!, the normal-operator apply, the Gram build, a Kaczmarz sweep — so a timeline of
// nsys / ncu --nvtx shows the solver structure above the kernels (SURVEY §5 tracing hook).
struct RlsNvtxRange {
  explicit RlsNvtxRange(const char* name);
  ~RlsNvtxRange();
};

// RAII device guard: every entry point runs on its context's device
struct RlsDeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit RlsDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~RlsDeviceGuard() {
    if (prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  int target = -1;
};

void rls_ctx_retain(rls_ctx_s* c);
void rls_ctx_release(rls_ctx_s* c);     // frees the context when the last reference goes
void rls_mat_retain(rls_mat_s* A);
void rls_mat_release(rls_mat_s* A);
void rls_normal_retain(rls_normal_t op);
void rls_normal_release(rls_normal_t op);
int32_t rls_ensure_gemv_scratch(rls_ctx_s* ctx, size_t bytes);
// diagnostic switches read from the environment (re-read on every call: tests flip them between solves; a getenv
// costs ~100 ns against >= 0.3 ms per normal-operator apply)
bool rls_env_flag(const char* name, bool dflt);

// internal (non-ABI) helpers used across translation units
int32_t rls_vec_create_internal(rls_ctx_s* ctx, int32_t dtype, int64_t len, rls_vec_s** out);
int32_t rls_allreduce_raw(rls_ctx_s* ctx, void* buf, int64_t nfloats);  // float sum-allreduce in place
int32_t rls_comm_init_all(rls_ctx_s* const* ctxs, int n);                // ncclCommInitAll over the contexts of one process
int32_t rls_allreduce_f64_host(rls_ctx_s* c, double* vals, int n);    // host doubles, sum over ranks in place
bool rls_p2p_available(const rls_ctx_s* c, int64_t nfloats);
int32_t rls_p2p_allreduce(rls_ctx_s* c, const float* src, int64_t sstride, int nsrc, int64_t nf, float* res, const int* gate);
int32_t rls_p2p_check_abort(rls_ctx_s* c);
void rls_ctx_peer_release(rls_ctx_s* c);
int32_t rls_normal_apply_raw(rls_normal_t op, const void* x, void* res, const int* gate);
int32_t rls_gemv_n_raw(rls_mat_s* A, const void* x, void* y, const int* gate);
int32_t rls_gemv_c_raw(rls_mat_s* A, const void* y, void* g, const int* gate);
int32_t rls_normal_from_function(rls_ctx_s* ctx, int32_t dtype, int64_t n, int32_t (*apply)(void*, const void*, void*, void*),
                                 void (*release)(void*), void* user, const char* name, rls_normal_t* out);
int32_t rls_normal_shape(rls_normal_t op, int64_t* n, int32_t* dtype);
bool rls_normal_graph_safe(rls_normal_t op);   // the apply is a fixed sequence of plain kernel launches (stream capture)
rls_ctx_s* rls_normal_ctx(rls_normal_t op);
rls_mat_s* rls_normal_matrix(rls_normal_t op);
int32_t rls_normal_check_abort(rls_normal_t op);
// per-cluster partial results of a one-pass apply whose final sum is left to the consuming kernel (ncl == 0: none)
struct NormalPartials { const float* gpart; int64_t gstride; int ncl; };
// apply with the finish (and optionally the FISTA momentum x*c1 + xold*c2) fused into neighbouring kernels; returns
// with np->ncl == 0 and nothing launched when the operator cannot defer (not row-major one-pass, or row-sharded)
int32_t rls_normal_apply_deferred_raw(rls_normal_t op, const void* x, const float* xold, const float* th_old, const float* th,
                                      const int* gate, NormalPartials* np);
// row-major one-pass kernels (rls_rowstream.cu)
struct RowPlan;
int32_t rls_rowpass_plan_create(rls_ctx_s* c, rls_mat_s* A, RowPlan** out);
void rls_rowpass_plan_destroy(RowPlan* p);
int32_t rls_rowpass_normal(RowPlan* p, const void* x, void* res, const int* gate);
int32_t rls_rowpass_normal_deferred(RowPlan* p, const void* x, const float* xold, const float* th_old, const float* th, const int* gate,
                                    const float** gpart, int64_t* gstride, int* ncl);
int32_t rls_rowpass_gemv_n(RowPlan* p, const void* x, void* y, const int* gate);
int32_t rls_rowpass_gemv_c(RowPlan* p, const void* y, void* g, const int* gate);
int32_t rls_rowpass_check_abort(RowPlan* p);
void rls_rowpass_describe(RowPlan* p, char* buf, int len);
RowPlan* rls_mat_rowplan(rls_mat_s* A);  // lazily created, owned by the matrix
// tensor-core paths (rls_tc.cu): batched normal operator for K right-hand sides, Gram build
struct TcBatchPlan;
bool rls_tc_batch_supported(const rls_mat_s* A, int K);
int32_t rls_tc_batch_create(rls_mat_s* A, int K, TcBatchPlan** out);
void rls_tc_batch_destroy(TcBatchPlan* p);
int32_t rls_tc_batch_apply(TcBatchPlan* p, const void* const* xs, void* const* outs, const int* const* gates);
int32_t rls_tc_batch_check_abort(TcBatchPlan* p);
int32_t rls_tc_batch_adjoint(TcBatchPlan* p, const void* const* bs, void* const* outs);
int32_t rls_normal_adjoint_batch_raw(rls_normal_t op, int K, const void* const* bs, void* const* outs, bool* done);
int32_t rls_tc_batch_debug(TcBatchPlan* p, int which, float* host, int64_t nfloats);
int32_t rls_tc_gram(rls_mat_s* A, rls_mat_s* G);
bool rls_tc_gram_batch_supported(const rls_mat_s* G, int K);
int32_t rls_tc_gram_batch_create(rls_mat_s* G, int K, TcBatchPlan** out);
int32_t rls_tc_gram_batch_apply(TcBatchPlan* p, const void* const* xs, void* const* outs, const int* const* gates);
int32_t rls_normal_apply_batch_raw(rls_normal_t op, int K, const void* const* xs, void* const* outs, const int* const* gates);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------
// programmatic dependent launch: the kernels of one solver iteration are chained on the stream;
// launched with this attribute, the next kernel's CTAs are resident and parked at
// griddepcontrol.wait while the previous kernel drains, which removes most of the ~15 us
// kernel-boundary gap around the big cluster kernel (profiles/r01_launch_gap_probe.txt).
// Every kernel launched through rls_launch_pdl starts with pdl_prologue().
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool rls_pdl_enabled();
void rls_pdl_suppress(bool on);
// RLS_TRACE_EVENTS=1: CUDA events around selected launches, dumped (durations and gaps, ms) by rls_trace_dump()
bool rls_trace_enabled();
void rls_trace_begin(cudaStream_t st, const char* name);
void rls_trace_end(cudaStream_t st);
void rls_trace_dump();
template <typename... KArgs, typename... Args>
static inline cudaError_t rls_launch_pdl(cudaStream_t st, dim3 grid, dim3 block, void (*kernel)(KArgs...), Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = rls_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------
// Individually rounded float arithmetic (no FMA contraction): the reference's
// broadcasts round after every operation, and so do we.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

// element traits: T = float (Float32) or float2 (ComplexF32, interleaved)
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr bool is_complex = false;
  static constexpr int vec = 4;  // elements per 128-bit load
  __device__ static __forceinline__ float zero() { return 0.f; }
  __device__ static __forceinline__ float add(float a, float b) { return fadd(a, b); }
  __device__ static __forceinline__ float sub(float a, float b) { return fsub(a, b); }
  __device__ static __forceinline__ float scale(float a, float s) { return fmul(a, s); }   // a * real
  __device__ static __forceinline__ float divr(float a, float s) { return fdiv(a, s); }    // a / real
  __device__ static __forceinline__ float neg(float a) { return -a; }
  __device__ static __forceinline__ float abs(float a) { return fabsf(a); }
  __device__ static __forceinline__ double abs2(float a) { return (double)a * (double)a; }
  // conj(a)*b accumulated in double
  __device__ static __forceinline__ void dotc(float a, float b, double& re, double& im) { re += (double)a * (double)b; }
  __device__ static __forceinline__ bool isnan_(float a) { return isnan(a); }
};
template <> struct Elem<float2> {
  static constexpr bool is_complex = true;
  static constexpr int vec = 2;
  __device__ static __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
  __device__ static __forceinline__ float2 add(float2 a, float2 b) { return make_float2(fadd(a.x, b.x), fadd(a.y, b.y)); }
  __device__ static __forceinline__ float2 sub(float2 a, float2 b) { return make_float2(fsub(a.x, b.x), fsub(a.y, b.y)); }
  __device__ static __forceinline__ float2 scale(float2 a, float s) { return make_float2(fmul(a.x, s), fmul(a.y, s)); }
  __device__ static __forceinline__ float2 divr(float2 a, float s) { return make_float2(fdiv(a.x, s), fdiv(a.y, s)); }
  __device__ static __forceinline__ float2 neg(float2 a) { return make_float2(-a.x, -a.y); }
  // Base.hypot / glibc hypotf evaluate in double for single precision operands
  __device__ static __forceinline__ float abs(float2 a) {
    return (float)sqrt((double)a.x * (double)a.x + (double)a.y * (double)a.y);
  }
  __device__ static __forceinline__ double abs2(float2 a) { return (double)a.x * (double)a.x + (double)a.y * (double)a.y; }
  __device__ static __forceinline__ void dotc(float2 a, float2 b, double& re, double& im) {
    re += (double)a.x * (double)b.x + (double)a.y * (double)b.y;
    im += (double)a.x * (double)b.y - (double)a.y * (double)b.x;
  }
  __device__ static __forceinline__ bool isnan_(float2 a) { return isnan(a.x) || isnan(a.y); }
};

// Julia's Complex*Complex: (ar*br - ai*bi, ar*bi + ai*br), every op rounded
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fsub(fmul(a.x, b.x), fmul(a.y, b.y)), fadd(fmul(a.x, b.y), fmul(a.y, b.x)));
}
__device__ __forceinline__ float cmul(float a, float b) { return fmul(a, b); }

// Julia's generic Complex{T}/Complex{T} (Smith) division
__device__ __forceinline__ float2 cdiv(float2 a, float2 b) {
  if (fabsf(b.x) <= fabsf(b.y)) {
    float r = fdiv(b.x, b.y);
    float den = fadd(b.y, fmul(r, b.x));
    return make_float2(fdiv(fadd(fmul(a.x, r), a.y), den), fdiv(fsub(fmul(a.y, r), a.x), den));
  }
  float r = fdiv(b.y, b.x);
  float den = fadd(b.x, fmul(r, b.y));
  return make_float2(fdiv(fadd(a.x, fmul(a.y, r)), den), fdiv(fsub(a.y, fmul(a.x, r)), den));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid-wide reduction of NACC doubles with a last-block finaliser.
// Every block reduces its threads' values in a fixed tree, writes one slot per
// accumulator, takes a ticket; the last block sums the slots in a fixed order and
// thread 0 runs `fin(totals)`.  The ticket resets itself for the next launch.
// Requires blockDim.x == BLOCK (multiple of 32, <= 1024) and gridDim.x <= RLS_MAX_RED_BLOCKS.
template <int NACC, int BLOCK, typename Fin>
__device__ __forceinline__ void grid_reduce_finalize(double (&v)[NACC], double* __restrict__ partials,
                                                     unsigned* __restrict__ ticket, Fin fin) {
  constexpr int NW = BLOCK / 32;
  __shared__ double s_red[NACC][NW];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    double w = warp_sum(v[k]);
    if (lane == 0) s_red[k][warp] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += s_red[k][w];
      partials[(size_t)blockIdx.x * NACC + k] = t;
    }
    __threadfence();
    unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double tot[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += BLOCK) t += __ldcg(&partials[(size_t)b * NACC + k]);
    tot[k] = warp_sum(t);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NACC; ++k)
    if (lane == 0) s_red[k][warp] = tot[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double total[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += s_red[k][w];
      total[k] = t;
    }
    *ticket = 0u;
    fin(total);
  }
}
#endif  // __CUDACC__
