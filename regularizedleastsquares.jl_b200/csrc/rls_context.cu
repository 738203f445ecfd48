// rls_context.cu — contexts, device vectors / matrices, Philox synthetic data,
// BLAS-1 style reductions, timers and the NCCL communicator (dlopen'ed, so the
// single-GPU path has no NCCL dependency).
#include <nvtx3/nvToolsExt.h>
#include <dlfcn.h>
#include <stdarg.h>

#include "rls_common.cuh"
#include "rls_philox.cuh"

// ------------------------------------------------------------------------------------
// error string (thread local)
// ------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void rls_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* rls_last_error(void) { return g_err; }

struct TraceRec { const char* name; cudaEvent_t a, b; };
static std::vector<TraceRec> g_trace;
RlsNvtxRange::RlsNvtxRange(const char* name) { nvtxRangePushA(name); }
RlsNvtxRange::~RlsNvtxRange() { nvtxRangePop(); }

bool rls_trace_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("RLS_TRACE_EVENTS"); on = (e && atoi(e) != 0) ? 1 : 0; }
  return on != 0;
}
void rls_trace_begin(cudaStream_t st, const char* name) {
  if (!rls_trace_enabled() || g_trace.size() >= 4096) return;
  TraceRec r{name, nullptr, nullptr};
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_trace.push_back(r);
}
void rls_trace_end(cudaStream_t st) {
  if (!rls_trace_enabled() || g_trace.empty() || g_trace.size() > 4096) return;
  cudaEventRecord(g_trace.back().b, st);
}
void rls_trace_dump() {
  if (!rls_trace_enabled() || g_trace.empty()) return;
  cudaDeviceSynchronize();
  const size_t n = g_trace.size(), from = n > 24 ? n - 24 : 0;
  {
    float span = 0.f;
    cudaEventElapsedTime(&span, g_trace[0].a, g_trace[n - 1].b);
    double sum = 0, gaps = 0, worst = 0; size_t iw = 0;
    for (size_t i = 0; i < n; ++i) {
      float d = 0.f, g = 0.f;
      cudaEventElapsedTime(&d, g_trace[i].a, g_trace[i].b);
      if (i) cudaEventElapsedTime(&g, g_trace[i - 1].b, g_trace[i].a);
      sum += d; gaps += g;
      if (d + g > worst) { worst = d + g; iw = i; }
    }
    fprintf(stderr, "[trace] %zu records: span %.3f ms, sum of durations %.3f ms, sum of gaps %.3f ms, worst record #%zu %.4f ms\n", n, span, sum, gaps, iw, worst);
    for (size_t q = 0; q < 8 && n >= 16; ++q) {  // mean duration of the large records per eighth of the trace
      double m = 0; int cnt = 0;
      for (size_t i = q * n / 8; i < (q + 1) * n / 8; ++i) {
        float d = 0.f; cudaEventElapsedTime(&d, g_trace[i].a, g_trace[i].b);
        if (d > 0.1f) { m += d; ++cnt; }
      }
      if (cnt) fprintf(stderr, "[trace] eighth %zu: mean large-kernel duration %.4f ms (%d launches)\n", q, m / cnt, cnt);
    }
    for (size_t i = 0; i < n && i < 4; ++i) {
      float d = 0.f, g = 0.f;
      cudaEventElapsedTime(&d, g_trace[i].a, g_trace[i].b);
      if (i) cudaEventElapsedTime(&g, g_trace[i - 1].b, g_trace[i].a);
      fprintf(stderr, "[trace] #%zu %-28s %8.4f ms   gap before %8.4f ms\n", i, g_trace[i].name, d, g);
    }
  }
  for (size_t i = from; i < n; ++i) {
    float dur = 0.f, gap = 0.f;
    cudaEventElapsedTime(&dur, g_trace[i].a, g_trace[i].b);
    if (i > from) cudaEventElapsedTime(&gap, g_trace[i - 1].b, g_trace[i].a);
    fprintf(stderr, "[trace] %-28s %8.4f ms   gap before %8.4f ms\n", g_trace[i].name, dur, gap);
  }
  for (auto& r : g_trace) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_trace.clear();
}

bool rls_env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) != 0 : dflt;
}

static thread_local int rls_pdl_suppressed = 0;
void rls_pdl_suppress(bool on) { rls_pdl_suppressed += on ? 1 : -1; }   // stream capture: plain (fully serialising) graph edges
bool rls_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("RLS_PDL"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on != 0 && rls_pdl_suppressed == 0;
}
extern "C" int32_t rls_abi_version(void) { return RLS_B200_ABI_VERSION; }

extern "C" int32_t rls_device_count(int32_t* count) {
  RLS_CHECK_ARG(count, "count is NULL");
  int c = 0;
  RLS_CUDA(cudaGetDeviceCount(&c));
  *count = c;
  return RLS_OK;
}

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
extern "C" int32_t rls_ctx_create(int32_t device, rls_ctx_t* out) {
  RLS_CHECK_ARG(out, "out is NULL");
  int ndev = 0;
  RLS_CUDA(cudaGetDeviceCount(&ndev));
  RLS_CHECK_ARG(device >= 0 && device < ndev, "device %d out of range (%d visible)", device, ndev);
  RlsDeviceGuard g(device);
  cudaDeviceProp prop;
  RLS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    rls_set_error("device %d is sm_%d%d; librls_b200 is built for sm_100a only and has no fallback", device,
                  prop.major, prop.minor);
    return RLS_ERR_UNSUPPORTED;
  }
  rls_ctx_s* c = new rls_ctx_s();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->l2_bytes = (size_t)prop.l2CacheSize;
  c->hbm_bytes = prop.totalGlobalMem;
  RLS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  RLS_CUDA(cudaEventCreate(&c->ev0));
  RLS_CUDA(cudaEventCreate(&c->ev1));
  RLS_CUDA(cudaMalloc(&c->red_partials, sizeof(double) * RLS_MAX_RED_BLOCKS * RLS_MAX_ACC));
  RLS_CUDA(cudaMalloc(&c->red_ticket, sizeof(unsigned) * 4));
  RLS_CUDA(cudaMemset(c->red_ticket, 0, sizeof(unsigned) * 4));
  RLS_CUDA(cudaMalloc(&c->red_out, sizeof(double) * RLS_MAX_ACC));
  RLS_CUDA(cudaMallocHost(&c->red_out_host, sizeof(double) * RLS_MAX_ACC));
  RLS_CUDA(cudaMalloc(&c->gemv_tickets, sizeof(unsigned) * 4096));
  RLS_CUDA(cudaMemset(c->gemv_tickets, 0, sizeof(unsigned) * 4096));
  *out = c;
  return RLS_OK;
}

typedef int (*nccl_destroy_fn)(void*);
static void* g_nccl_lib = nullptr;

void rls_ctx_retain(rls_ctx_s* c) { if (c) c->refs.fetch_add(1); }

extern "C" int32_t rls_ctx_destroy(rls_ctx_t c) {
  if (!c) return RLS_OK;
  rls_ctx_release(c);   // vectors / matrices / operators / solvers still alive keep the context until they go
  return RLS_OK;
}

static int32_t ctx_free(rls_ctx_s* c) {
  RlsDeviceGuard g(c->device);
  rls_ctx_peer_release(c);
  cudaStreamSynchronize(c->stream);
  if (c->nccl_comm && g_nccl_lib) {
    nccl_destroy_fn f = (nccl_destroy_fn)dlsym(g_nccl_lib, "ncclCommDestroy");
    if (f) f(c->nccl_comm);
  }
  cudaFree(c->red_partials);
  cudaFree(c->red_ticket);
  cudaFree(c->red_out);
  cudaFreeHost(c->red_out_host);
  cudaFree(c->gemv_scratch);
  cudaFree(c->svt_scratch);
  cudaFree(c->gemv_tickets);
  cudaFree(c->flush_buf);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
  return RLS_OK;
}

void rls_ctx_release(rls_ctx_s* c) {
  if (c && c->refs.fetch_sub(1) == 1) ctx_free(c);
}

extern "C" int32_t rls_ctx_sync(rls_ctx_t c) {
  RLS_CHECK_ARG(c, "ctx is NULL");
  RlsDeviceGuard g(c->device);
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

extern "C" int32_t rls_ctx_device_info(rls_ctx_t c, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                                       int64_t* l2_bytes, int64_t* hbm_bytes) {
  RLS_CHECK_ARG(c, "ctx is NULL");
  if (sm_count) *sm_count = c->sm_count;
  if (cc_major) *cc_major = c->cc_major;
  if (cc_minor) *cc_minor = c->cc_minor;
  if (l2_bytes) *l2_bytes = (int64_t)c->l2_bytes;
  if (hbm_bytes) *hbm_bytes = (int64_t)c->hbm_bytes;
  return RLS_OK;
}

extern "C" int32_t rls_timer_start(rls_ctx_t c) {
  RLS_CHECK_ARG(c, "ctx is NULL");
  RlsDeviceGuard g(c->device);
  RLS_CUDA(cudaEventRecord(c->ev0, c->stream));
  return RLS_OK;
}

extern "C" int32_t rls_timer_stop(rls_ctx_t c, float* ms) {
  RLS_CHECK_ARG(c && ms, "NULL argument");
  RlsDeviceGuard g(c->device);
  RLS_CUDA(cudaEventRecord(c->ev1, c->stream));
  RLS_CUDA(cudaEventSynchronize(c->ev1));
  RLS_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return RLS_OK;
}

extern "C" int32_t rls_ctx_launch_count(rls_ctx_t c, int64_t* launches) {
  RLS_CHECK_ARG(c && launches, "NULL argument");
  *launches = c->launches;
  return RLS_OK;
}

__global__ void flush_kernel(float4* buf, size_t n4, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) buf[i] = make_float4(v, v, v, v);
}

extern "C" int32_t rls_ctx_flush_l2(rls_ctx_t c) {
  RLS_CHECK_ARG(c, "ctx is NULL");
  RlsDeviceGuard g(c->device);
  if (!c->flush_buf) {
    c->flush_bytes = (c->l2_bytes ? c->l2_bytes : ((size_t)128 << 20)) * 2;
    RLS_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  flush_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>((float4*)c->flush_buf, c->flush_bytes / 16, 1.0f);
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

int32_t rls_ensure_gemv_scratch(rls_ctx_s* c, size_t bytes) {
  if (c->gemv_scratch_bytes >= bytes) return RLS_OK;
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  if (c->gemv_scratch) RLS_CUDA(cudaFree(c->gemv_scratch));
  c->gemv_scratch = nullptr;
  c->gemv_scratch_bytes = 0;
  RLS_CUDA(cudaMalloc(&c->gemv_scratch, bytes));
  c->gemv_scratch_bytes = bytes;
  return RLS_OK;
}

// ------------------------------------------------------------------------------------
// NCCL (dlopen)
// ------------------------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };
typedef int (*nccl_get_uid_fn)(NcclUniqueId*);
typedef int (*nccl_init_rank_fn)(void**, int, NcclUniqueId, int);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);

static nccl_get_uid_fn p_ncclGetUniqueId = nullptr;
static nccl_init_rank_fn p_ncclCommInitRank = nullptr;
static nccl_allreduce_fn p_ncclAllReduce = nullptr;
static nccl_errstr_fn p_ncclGetErrorString = nullptr;

static int32_t load_nccl() {
  if (g_nccl_lib) return RLS_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl_lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl_lib) break;
  }
  if (!g_nccl_lib) {
    rls_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return RLS_ERR_COMM;
  }
  p_ncclGetUniqueId = (nccl_get_uid_fn)dlsym(g_nccl_lib, "ncclGetUniqueId");
  p_ncclCommInitRank = (nccl_init_rank_fn)dlsym(g_nccl_lib, "ncclCommInitRank");
  p_ncclAllReduce = (nccl_allreduce_fn)dlsym(g_nccl_lib, "ncclAllReduce");
  p_ncclGetErrorString = (nccl_errstr_fn)dlsym(g_nccl_lib, "ncclGetErrorString");
  if (!p_ncclGetUniqueId || !p_ncclCommInitRank || !p_ncclAllReduce) {
    rls_set_error("libnccl is missing required symbols");
    return RLS_ERR_COMM;
  }
  return RLS_OK;
}

#define RLS_NCCL(expr)                                                                          \
  do {                                                                                          \
    int _r = (expr);                                                                            \
    if (_r != 0) {                                                                              \
      rls_set_error("NCCL error %d at %s:%d: %s", _r, __FILE__, __LINE__,                       \
                    p_ncclGetErrorString ? p_ncclGetErrorString(_r) : "?");                     \
      return RLS_ERR_COMM;                                                                      \
    }                                                                                           \
  } while (0)

extern "C" int32_t rls_comm_unique_id(void* id128) {
  RLS_CHECK_ARG(id128, "id128 is NULL");
  RLS_TRY(load_nccl());
  NcclUniqueId id;
  RLS_NCCL(p_ncclGetUniqueId(&id));
  memcpy(id128, &id, 128);
  return RLS_OK;
}

extern "C" int32_t rls_ctx_comm_init(rls_ctx_t c, int32_t rank, int32_t nranks, const void* id128) {
  RLS_CHECK_ARG(c && id128, "NULL argument");
  RLS_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d / nranks %d", rank, nranks);
  RLS_CHECK_ARG(!c->nccl_comm, "communicator already initialised");
  RLS_TRY(load_nccl());
  RlsDeviceGuard g(c->device);
  NcclUniqueId id;
  memcpy(&id, id128, 128);
  RLS_NCCL(p_ncclCommInitRank(&c->nccl_comm, nranks, id, rank));
  c->rank = rank;
  c->nranks = nranks;
  return RLS_OK;
}

// one communicator over the contexts of ONE process (rls_group.cu): ncclCommInitAll, rank i = contexts[i]
int32_t rls_comm_init_all(rls_ctx_s* const* ctxs, int n) {
  RLS_TRY(load_nccl());
  typedef int (*nccl_init_all_fn)(void**, int, const int*);
  nccl_init_all_fn init_all = (nccl_init_all_fn)dlsym(g_nccl_lib, "ncclCommInitAll");
  RLS_CHECK_ARG(init_all, "libnccl has no ncclCommInitAll");
  std::vector<void*> comms(n, nullptr);
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    RLS_CHECK_ARG(!ctxs[i]->nccl_comm, "communicator already initialised");
    devs[i] = ctxs[i]->device;
  }
  RLS_NCCL(init_all(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) { ctxs[i]->nccl_comm = comms[i]; ctxs[i]->rank = i; ctxs[i]->nranks = n; }
  return RLS_OK;
}

extern "C" int32_t rls_ctx_comm_info(rls_ctx_t c, int32_t* rank, int32_t* nranks) {
  RLS_CHECK_ARG(c, "ctx is NULL");
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  return RLS_OK;
}

int32_t rls_allreduce_raw(rls_ctx_s* c, void* buf, int64_t nfloats) {
  if (c->nranks <= 1) return RLS_OK;
  RLS_CHECK_ARG(c->nccl_comm, "context has nranks>1 but no communicator");
  // ncclFloat32 = 7, ncclSum = 0
  RLS_NCCL(p_ncclAllReduce(buf, buf, (size_t)nfloats, 7, 0, c->nccl_comm, c->stream));
  return RLS_OK;
}

// Sum of n <= RLS_MAX_ACC host doubles over the ranks (in place): the scalars a row-sharded solve needs globally —
// ‖A‖_F² (SystemMatrixBasedNormalization, NormalizedRegularization.jl:47-58), ‖b‖₁ and length(b)
// (MeasurementBasedNormalization :40-43; ADMM's σ_abs = sqrt(length(b))·absTol, ADMM.jl:212).
int32_t rls_allreduce_f64_host(rls_ctx_s* c, double* vals, int n) {
  if (c->nranks <= 1) return RLS_OK;
  RLS_CHECK_ARG(c->nccl_comm, "context has nranks>1 but no communicator");
  RLS_CHECK_ARG(vals && n >= 1 && n <= RLS_MAX_ACC, "allreduce_f64: 1..%d values", RLS_MAX_ACC);
  RLS_CUDA(cudaStreamSynchronize(c->stream));  // red_out_host may still be the target of an earlier download
  memcpy(c->red_out_host, vals, sizeof(double) * n);
  RLS_CUDA(cudaMemcpyAsync(c->red_out, c->red_out_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  // ncclFloat64 = 8, ncclSum = 0
  RLS_NCCL(p_ncclAllReduce(c->red_out, c->red_out, (size_t)n, 8, 0, c->nccl_comm, c->stream));
  RLS_CUDA(cudaMemcpyAsync(c->red_out_host, c->red_out, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(vals, c->red_out_host, sizeof(double) * n);
  return RLS_OK;
}

extern "C" int32_t rls_ctx_allreduce_f64(rls_ctx_t c, double* vals, int32_t n) {
  RLS_CHECK_ARG(c && vals, "NULL argument");
  RlsDeviceGuard g(c->device);
  return rls_allreduce_f64_host(c, vals, n);
}

extern "C" int32_t rls_vec_allreduce(rls_vec_t v) {
  RLS_CHECK_ARG(v, "vec is NULL");
  RlsDeviceGuard g(v->ctx->device);
  return rls_allreduce_raw(v->ctx, v->d, v->len * (v->dtype == RLS_C32 ? 2 : 1));
}

// ------------------------------------------------------------------------------------
// vectors
// ------------------------------------------------------------------------------------
int32_t rls_vec_create_internal(rls_ctx_s* ctx, int32_t dtype, int64_t len, rls_vec_s** out) {
  RLS_CHECK_ARG(ctx && out, "NULL argument");
  RLS_CHECK_ARG(dtype == RLS_F32 || dtype == RLS_C32, "unsupported dtype %d (Float32 / ComplexF32 only)", dtype);
  RLS_CHECK_ARG(len >= 0, "negative length");
  RlsDeviceGuard g(ctx->device);
  rls_vec_s* v = new rls_vec_s{ctx, dtype, len, nullptr, true};
  size_t bytes = (size_t)(len > 0 ? len : 1) * rls_elem_size(dtype);
  bytes = (bytes + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(&v->d, bytes);
  if (e != cudaSuccess) {
    delete v;
    rls_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return RLS_ERR_NOMEM;
  }
  cudaMemsetAsync(v->d, 0, bytes, ctx->stream);
  rls_ctx_retain(ctx);
  *out = v;
  return RLS_OK;
}

extern "C" int32_t rls_vec_create(rls_ctx_t ctx, int32_t dtype, int64_t len, rls_vec_t* out) {
  return rls_vec_create_internal(ctx, dtype, len, out);
}

extern "C" int32_t rls_vec_destroy(rls_vec_t v) {
  if (!v) return RLS_OK;
  RlsDeviceGuard g(v->ctx->device);
  if (v->owned && v->d) {
    cudaStreamSynchronize(v->ctx->stream);
    cudaFree(v->d);
  }
  rls_ctx_s* c = v->ctx;
  delete v;
  rls_ctx_release(c);
  return RLS_OK;
}

extern "C" int32_t rls_vec_len(rls_vec_t v, int64_t* len, int32_t* dtype) {
  RLS_CHECK_ARG(v, "vec is NULL");
  if (len) *len = v->len;
  if (dtype) *dtype = v->dtype;
  return RLS_OK;
}

extern "C" int32_t rls_vec_upload(rls_vec_t v, const void* host, int64_t len) {
  RLS_CHECK_ARG(v && host, "NULL argument");
  RLS_CHECK_ARG(len == v->len, "length mismatch: vec %lld, host %lld", (long long)v->len, (long long)len);
  RlsDeviceGuard g(v->ctx->device);
  RLS_CUDA(cudaMemcpyAsync(v->d, host, (size_t)len * rls_elem_size(v->dtype), cudaMemcpyHostToDevice, v->ctx->stream));
  // pageable host memory: the call returns after staging; pinned memory: the caller keeps it alive until sync
  return RLS_OK;
}

extern "C" int32_t rls_vec_download(rls_vec_t v, void* host, int64_t len) {
  RLS_CHECK_ARG(v && host, "NULL argument");
  RLS_CHECK_ARG(len == v->len, "length mismatch: vec %lld, host %lld", (long long)v->len, (long long)len);
  RlsDeviceGuard g(v->ctx->device);
  RLS_CUDA(cudaMemcpyAsync(host, v->d, (size_t)len * rls_elem_size(v->dtype), cudaMemcpyDeviceToHost, v->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return RLS_OK;
}

extern "C" int32_t rls_vec_copy(rls_vec_t dst, rls_vec_t src) {
  RLS_CHECK_ARG(dst && src, "NULL argument");
  RLS_CHECK_ARG(dst->len == src->len && dst->dtype == src->dtype, "vec_copy shape/dtype mismatch");
  RlsDeviceGuard g(dst->ctx->device);
  RLS_CUDA(cudaMemcpyAsync(dst->d, src->d, (size_t)src->len * rls_elem_size(src->dtype), cudaMemcpyDeviceToDevice,
                           dst->ctx->stream));
  return RLS_OK;
}

extern "C" int32_t rls_vec_device_ptr(rls_vec_t v, void** ptr) {
  RLS_CHECK_ARG(v && ptr, "NULL argument");
  *ptr = v->d;
  return RLS_OK;
}

__global__ void fill_kernel(float2* x, int64_t n, float re, float im) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = make_float2(re, im);
}
__global__ void fill_kernel(float* x, int64_t n, float re) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = re;
}

extern "C" int32_t rls_vec_fill(rls_vec_t v, float re, float im) {
  RLS_CHECK_ARG(v, "vec is NULL");
  if (v->len == 0) return RLS_OK;
  RlsDeviceGuard g(v->ctx->device);
  int grid = (int)((v->len + 255) / 256);
  if (v->dtype == RLS_C32)
    fill_kernel<<<grid, 256, 0, v->ctx->stream>>>((float2*)v->d, v->len, re, im);
  else
    fill_kernel<<<grid, 256, 0, v->ctx->stream>>>((float*)v->d, v->len, re);
  v->ctx->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// ---- Philox fills --------------------------------------------------------------------
template <bool CPLX>
__global__ void vec_philox_kernel(float* x, int64_t n, uint64_t seed, uint64_t stream, int dist, float scale,
                                  int64_t offset) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t idx = (uint64_t)(offset + i);
  if (CPLX) {
    x[2 * i] = philox_value(seed, idx, stream, 0u, dist, scale);
    x[2 * i + 1] = philox_value(seed, idx, stream, 1u, dist, scale);
  } else {
    x[i] = philox_value(seed, idx, stream, 0u, dist, scale);
  }
}

extern "C" int32_t rls_vec_fill_philox(rls_vec_t v, uint64_t seed, uint64_t stream, int32_t dist, float scale,
                                       int64_t offset) {
  RLS_CHECK_ARG(v, "vec is NULL");
  RLS_CHECK_ARG(dist == RLS_DIST_UNIFORM01 || dist == RLS_DIST_IH4, "unknown distribution %d", dist);
  if (v->len == 0) return RLS_OK;
  RlsDeviceGuard g(v->ctx->device);
  int grid = (int)((v->len + 255) / 256);
  if (v->dtype == RLS_C32)
    vec_philox_kernel<true><<<grid, 256, 0, v->ctx->stream>>>((float*)v->d, v->len, seed, stream, dist, scale, offset);
  else
    vec_philox_kernel<false><<<grid, 256, 0, v->ctx->stream>>>((float*)v->d, v->len, seed, stream, dist, scale, offset);
  v->ctx->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// A[i,j]: counter index (row_offset+i) + j*m_global, stream 0.  One thread per 4 rows of a column.
template <bool CPLX>
__global__ void mat_philox_kernel(float* A, int64_t m, int64_t n, int64_t ld, uint64_t seed, int dist, float scale,
                                  int64_t row_offset, int64_t m_global) {
  const int64_t rows_per_blk = (int64_t)blockDim.x;
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
    for (int64_t i = (int64_t)blockIdx.x * rows_per_blk + threadIdx.x; i < m; i += (int64_t)gridDim.x * rows_per_blk) {
      uint64_t idx = (uint64_t)(row_offset + i) + (uint64_t)j * (uint64_t)m_global;
      if (CPLX) {
        float2 v = make_float2(philox_value(seed, idx, 0ull, 0u, dist, scale), philox_value(seed, idx, 0ull, 1u, dist, scale));
        ((float2*)A)[i + j * ld] = v;
      } else {
        A[i + j * ld] = philox_value(seed, idx, 0ull, 0u, dist, scale);
      }
    }
  }
}

// row-major storage: same counters (the VALUE of A[i,j] does not depend on the layout), threads run along a row
template <bool CPLX>
__global__ void mat_philox_rowmajor_kernel(float* A, int64_t m, int64_t n, int64_t ld, uint64_t seed, int dist, float scale,
                                           int64_t row_offset, int64_t m_global) {
  for (int64_t i = blockIdx.y; i < m; i += gridDim.y) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
      uint64_t idx = (uint64_t)(row_offset + i) + (uint64_t)j * (uint64_t)m_global;
      if (CPLX) {
        float2 v = make_float2(philox_value(seed, idx, 0ull, 0u, dist, scale), philox_value(seed, idx, 0ull, 1u, dist, scale));
        ((float2*)A)[i * ld + j] = v;
      } else {
        A[i * ld + j] = philox_value(seed, idx, 0ull, 0u, dist, scale);
      }
    }
  }
}

extern "C" int32_t rls_mat_fill_philox(rls_mat_t A, uint64_t seed, int32_t dist, float scale, int64_t row_offset,
                                       int64_t m_global) {
  RLS_CHECK_ARG(A, "mat is NULL");
  RLS_CHECK_ARG(dist == RLS_DIST_UNIFORM01 || dist == RLS_DIST_IH4, "unknown distribution %d", dist);
  RLS_CHECK_ARG(row_offset >= 0 && row_offset + A->m <= m_global, "row shard [%lld,%lld) outside global m=%lld",
                (long long)row_offset, (long long)(row_offset + A->m), (long long)m_global);
  if (A->m == 0 || A->n == 0) return RLS_OK;
  RlsDeviceGuard g(A->ctx->device);
  if (A->layout == RLS_LAYOUT_ROWMAJOR) {
    dim3 rgrid((unsigned)std::min<int64_t>((A->n + 255) / 256, 64), (unsigned)std::min<int64_t>(A->m, 16384));
    if (A->dtype == RLS_C32)
      mat_philox_rowmajor_kernel<true><<<rgrid, 256, 0, A->ctx->stream>>>((float*)A->d, A->m, A->n, A->ld, seed, dist, scale, row_offset, m_global);
    else
      mat_philox_rowmajor_kernel<false><<<rgrid, 256, 0, A->ctx->stream>>>((float*)A->d, A->m, A->n, A->ld, seed, dist, scale, row_offset, m_global);
    A->ctx->launches++;
    RLS_CUDA(cudaGetLastError());
    return RLS_OK;
  }
  dim3 grid((unsigned)std::min<int64_t>((A->m + 255) / 256, 64), (unsigned)std::min<int64_t>(A->n, 16384));
  if (A->dtype == RLS_C32)
    mat_philox_kernel<true><<<grid, 256, 0, A->ctx->stream>>>((float*)A->d, A->m, A->n, A->ld, seed, dist, scale, row_offset, m_global);
  else
    mat_philox_kernel<false><<<grid, 256, 0, A->ctx->stream>>>((float*)A->d, A->m, A->n, A->ld, seed, dist, scale, row_offset, m_global);
  A->ctx->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// ---- reductions ------------------------------------------------------------------------
constexpr int RED_BLOCK = 256;

static inline int red_grid(const rls_ctx_s* c, int64_t n) {
  int64_t g = (n + RED_BLOCK * 4 - 1) / (RED_BLOCK * 4);
  int64_t cap = (int64_t)c->sm_count * 8;
  if (g > cap) g = cap;
  if (g > RLS_MAX_RED_BLOCKS) g = RLS_MAX_RED_BLOCKS;
  if (g < 1) g = 1;
  return (int)g;
}

// mode 0: sum |x|^2 ; mode 1: sum |x| ; mode 2: conj(x).y
template <typename T, int MODE>
__global__ void __launch_bounds__(RED_BLOCK) reduce_kernel(const T* __restrict__ x, const T* __restrict__ y, int64_t n,
                                                           double* partials, unsigned* ticket, double* out) {
  double acc[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * RED_BLOCK) {
    T a = x[i];
    if (MODE == 0) acc[0] += Elem<T>::abs2(a);
    else if (MODE == 1) acc[0] += (double)Elem<T>::abs(a);
    else Elem<T>::dotc(a, y[i], acc[0], acc[1]);
  }
  grid_reduce_finalize<2, RED_BLOCK>(acc, partials, ticket, [=](double* t) { out[0] = t[0]; out[1] = t[1]; });
}

template <int MODE>
static int32_t reduce_to_host(rls_vec_s* x, rls_vec_s* y, double* out, int nout) {
  rls_ctx_s* c = x->ctx;
  RlsDeviceGuard g(c->device);
  if (x->len == 0) {
    for (int k = 0; k < nout; ++k) out[k] = 0.0;
    return RLS_OK;
  }
  int grid = red_grid(c, x->len);
  if (x->dtype == RLS_C32)
    reduce_kernel<float2, MODE><<<grid, RED_BLOCK, 0, c->stream>>>((const float2*)x->d, y ? (const float2*)y->d : nullptr, x->len, c->red_partials, c->red_ticket, c->red_out);
  else
    reduce_kernel<float, MODE><<<grid, RED_BLOCK, 0, c->stream>>>((const float*)x->d, y ? (const float*)y->d : nullptr, x->len, c->red_partials, c->red_ticket, c->red_out);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  RLS_CUDA(cudaMemcpyAsync(c->red_out_host, c->red_out, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < nout; ++k) out[k] = c->red_out_host[k];
  return RLS_OK;
}

extern "C" int32_t rls_vec_nrm2(rls_vec_t v, double* out) {
  RLS_CHECK_ARG(v && out, "NULL argument");
  double s = 0;
  RLS_TRY(reduce_to_host<0>(v, nullptr, &s, 1));
  *out = sqrt(s);
  return RLS_OK;
}

extern "C" int32_t rls_vec_asum(rls_vec_t v, double* out) {
  RLS_CHECK_ARG(v && out, "NULL argument");
  return reduce_to_host<1>(v, nullptr, out, 1);
}

extern "C" int32_t rls_vec_dot(rls_vec_t a, rls_vec_t b, double out[2]) {
  RLS_CHECK_ARG(a && b && out, "NULL argument");
  RLS_CHECK_ARG(a->len == b->len && a->dtype == b->dtype, "dot: shape/dtype mismatch");
  return reduce_to_host<2>(a, b, out, 2);
}

// ------------------------------------------------------------------------------------
// matrices
// ------------------------------------------------------------------------------------
extern "C" int32_t rls_mat_create_layout(rls_ctx_t ctx, int32_t dtype, int64_t m, int64_t n, const void* host, int64_t ld,
                                         int32_t layout, rls_mat_t* out) {
  RLS_CHECK_ARG(ctx && out, "NULL argument");
  RLS_CHECK_ARG(dtype == RLS_F32 || dtype == RLS_C32, "unsupported dtype %d (Float32 / ComplexF32 only)", dtype);
  RLS_CHECK_ARG(m >= 0 && n >= 0, "negative shape");
  RLS_CHECK_ARG(!host || ld >= m, "ld < m");
  RLS_CHECK_ARG(layout >= RLS_LAYOUT_COLMAJOR && layout <= RLS_LAYOUT_AUTO, "unknown layout %d", layout);
  if (layout == RLS_LAYOUT_AUTO) {
    // rows of at most 16 x 8192 floats fit the cluster decomposition of rls_rowstream.cu
    const int64_t nf = n * (dtype == RLS_C32 ? 2 : 1);
    // and pay off once A no longer sits in L2: below ~64 MB an iteration is launch-latency bound and the
    // two-kernel column-major path is the shorter one (measured on C1: 40 vs 56 us per CGNR iteration)
    const char* force = getenv("RLS_LAYOUT");
    const double bytes = (double)m * (double)n * (double)rls_elem_size(dtype);
    if (force && force[0] == 'c') layout = RLS_LAYOUT_COLMAJOR;
    else if (force && force[0] == 'r') layout = RLS_LAYOUT_ROWMAJOR;
    else layout = ((nf + 3) / 4 * 4 <= 16 * 8192 && bytes >= 64.0 * 1024 * 1024) ? RLS_LAYOUT_ROWMAJOR : RLS_LAYOUT_COLMAJOR;
  }
  RlsDeviceGuard g(ctx->device);
  // device leading dimension padded to a 16-byte multiple so every column (row) supports 128-bit
  // loads and bulk copies; the padding is zeroed once and never written again
  const bool rowmajor = layout == RLS_LAYOUT_ROWMAJOR;
  int64_t vec = dtype == RLS_C32 ? 2 : 4;
  int64_t fast = rowmajor ? n : m, slow = rowmajor ? m : n;
  int64_t dld = ((fast + vec - 1) / vec) * vec;
  if (dld == 0) dld = vec;
  rls_mat_s* A = new rls_mat_s{ctx, dtype, m, n, dld, nullptr, true};
  rls_ctx_retain(ctx);
  A->layout = layout;
  size_t bytes = (size_t)dld * (size_t)(slow > 0 ? slow : 1) * rls_elem_size(dtype);
  cudaError_t e = cudaMalloc(&A->d, bytes);
  if (e != cudaSuccess) {
    delete A;
    rls_ctx_release(ctx);
    rls_set_error("cudaMalloc(%zu) for %lldx%lld matrix failed: %s", bytes, (long long)m, (long long)n, cudaGetErrorString(e));
    return RLS_ERR_NOMEM;
  }
  if (dld != fast) cudaMemsetAsync(A->d, 0, bytes, ctx->stream);
  *out = A;
  if (host) {
    int32_t s = rls_mat_upload(A, host, ld);
    if (s != RLS_OK) {
      rls_mat_destroy(A);
      *out = nullptr;
      return s;
    }
  }
  return RLS_OK;
}

extern "C" int32_t rls_mat_create(rls_ctx_t ctx, int32_t dtype, int64_t m, int64_t n, const void* host, int64_t ld,
                                  rls_mat_t* out) {
  return rls_mat_create_layout(ctx, dtype, m, n, host, ld, RLS_LAYOUT_AUTO, out);
}

extern "C" int32_t rls_mat_layout(rls_mat_t A, int32_t* layout) {
  RLS_CHECK_ARG(A && layout, "NULL argument");
  *layout = A->layout;
  return RLS_OK;
}

RowPlan* rls_mat_rowplan(rls_mat_s* A) {
  if (A->layout != RLS_LAYOUT_ROWMAJOR) return nullptr;
  if (!A->rowplan && rls_rowpass_plan_create(A->ctx, A, &A->rowplan) != RLS_OK) A->rowplan = nullptr;
  return A->rowplan;
}

extern "C" int32_t rls_mat_wrap_device(rls_ctx_t ctx, int32_t dtype, int64_t m, int64_t n, void* dev, int64_t ld,
                                       rls_mat_t* out) {
  RLS_CHECK_ARG(ctx && out && dev, "NULL argument");
  RLS_CHECK_ARG(dtype == RLS_F32 || dtype == RLS_C32, "unsupported dtype %d", dtype);
  RLS_CHECK_ARG(m >= 0 && n >= 0 && ld >= m, "bad shape");
  *out = new rls_mat_s{ctx, dtype, m, n, ld, dev, false};
  rls_ctx_retain(ctx);
  return RLS_OK;
}

void rls_mat_retain(rls_mat_s* A) { if (A) A->refs.fetch_add(1); }

void rls_mat_release(rls_mat_s* A) {
  if (!A || A->refs.fetch_sub(1) != 1) return;
  rls_ctx_s* c = A->ctx;
  {
    RlsDeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    if (A->rowplan) rls_rowpass_plan_destroy(A->rowplan);
    if (A->owned && A->d) cudaFree(A->d);
    delete A;
  }
  rls_ctx_release(c);
}

extern "C" int32_t rls_mat_destroy(rls_mat_t A) {
  rls_mat_release(A);   // operators / solvers built on A keep it until they go
  return RLS_OK;
}

extern "C" int32_t rls_mat_shape(rls_mat_t A, int64_t* m, int64_t* n, int32_t* dtype) {
  RLS_CHECK_ARG(A, "mat is NULL");
  if (m) *m = A->m;
  if (n) *n = A->n;
  if (dtype) *dtype = A->dtype;
  return RLS_OK;
}

// out(r,c) at c*ldo + r  <-  in(r,c) at r*ldi + c   (32x32 tiles through shared memory)
template <typename T>
__global__ void __launch_bounds__(256) transpose_tiles_kernel(const T* __restrict__ in, int64_t ldi, T* __restrict__ out, int64_t ldo,
                                                             int64_t R, int64_t C) {
  __shared__ T tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t tiles_c = (C + 31) / 32, tiles_r = (R + 31) / 32;
  for (int64_t t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    for (int k = ty; k < 32; k += 8)
      if (r0 + k < R && c0 + tx < C) tile[k][tx] = in[(r0 + k) * ldi + c0 + tx];
    __syncthreads();
    for (int k = ty; k < 32; k += 8)
      if (c0 + k < C && r0 + tx < R) out[(c0 + k) * ldo + r0 + tx] = tile[tx][k];
    __syncthreads();
  }
}

template <typename T>
static void launch_transpose(rls_ctx_s* c, const void* in, int64_t ldi, void* out, int64_t ldo, int64_t R, int64_t C) {
  int64_t tiles = ((R + 31) / 32) * ((C + 31) / 32);
  int grid = (int)std::min<int64_t>(tiles, (int64_t)c->sm_count * 16);
  transpose_tiles_kernel<T><<<grid, 256, 0, c->stream>>>((const T*)in, ldi, (T*)out, ldo, R, C);
  c->launches++;
}

// host (column-major) <-> device row-major: column blocks are staged column-major on the device and
// transposed there, so the host side stays a plain strided copy
static int32_t rowmajor_transfer(rls_mat_s* A, void* host, int64_t ld, bool upload) {
  rls_ctx_s* c = A->ctx;
  const size_t es = rls_elem_size(A->dtype);
  int64_t chunk = std::max<int64_t>(32, ((int64_t)64 << 20) / std::max<int64_t>(1, (int64_t)(A->m * es)));
  chunk = std::min<int64_t>((chunk + 31) / 32 * 32, std::max<int64_t>(A->n, 1));
  void* stage = nullptr;
  cudaError_t e = cudaMalloc(&stage, (size_t)A->m * (size_t)chunk * es);
  if (e != cudaSuccess) { rls_set_error("cudaMalloc for the transpose staging buffer failed: %s", cudaGetErrorString(e)); return RLS_ERR_NOMEM; }
  int32_t status = RLS_OK;
  for (int64_t j0 = 0; j0 < A->n && status == RLS_OK; j0 += chunk) {
    const int64_t nc = std::min<int64_t>(chunk, A->n - j0);
    char* hp = (char*)host + (size_t)j0 * (size_t)ld * es;
    char* dp = (char*)A->d + (size_t)j0 * es;
    if (upload) {
      e = cudaMemcpy2DAsync(stage, (size_t)A->m * es, hp, (size_t)ld * es, (size_t)A->m * es, (size_t)nc, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) {
        // staged block: (r = column j, c = row i) at r*m + c  ->  A[i*ld + j]
        if (A->dtype == RLS_C32) launch_transpose<float2>(c, stage, A->m, dp, A->ld, nc, A->m);
        else launch_transpose<float>(c, stage, A->m, dp, A->ld, nc, A->m);
        e = cudaGetLastError();
      }
    } else {
      if (A->dtype == RLS_C32) launch_transpose<float2>(c, dp, A->ld, stage, A->m, A->m, nc);
      else launch_transpose<float>(c, dp, A->ld, stage, A->m, A->m, nc);
      e = cudaGetLastError();
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(hp, (size_t)ld * es, stage, (size_t)A->m * es, (size_t)A->m * es, (size_t)nc, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { rls_set_error("CUDA error in row-major transfer: %s", cudaGetErrorString(e)); status = RLS_ERR_CUDA; }
  }
  cudaFree(stage);
  return status;
}

extern "C" int32_t rls_mat_upload(rls_mat_t A, const void* host, int64_t ld) {
  RLS_CHECK_ARG(A && host, "NULL argument");
  RLS_CHECK_ARG(ld >= A->m, "ld < m");
  if (A->m == 0 || A->n == 0) return RLS_OK;
  RlsDeviceGuard g(A->ctx->device);
  if (A->layout == RLS_LAYOUT_ROWMAJOR) return rowmajor_transfer(A, const_cast<void*>(host), ld, true);
  size_t es = rls_elem_size(A->dtype);
  RLS_CUDA(cudaMemcpy2DAsync(A->d, (size_t)A->ld * es, host, (size_t)ld * es, (size_t)A->m * es, (size_t)A->n,
                             cudaMemcpyHostToDevice, A->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(A->ctx->stream));
  return RLS_OK;
}

extern "C" int32_t rls_mat_download(rls_mat_t A, void* host, int64_t ld) {
  RLS_CHECK_ARG(A && host, "NULL argument");
  RLS_CHECK_ARG(ld >= A->m, "ld < m");
  if (A->m == 0 || A->n == 0) return RLS_OK;
  RlsDeviceGuard g(A->ctx->device);
  if (A->layout == RLS_LAYOUT_ROWMAJOR) return rowmajor_transfer(A, host, ld, false);
  size_t es = rls_elem_size(A->dtype);
  RLS_CUDA(cudaMemcpy2DAsync(host, (size_t)ld * es, A->d, (size_t)A->ld * es, (size_t)A->m * es, (size_t)A->n,
                             cudaMemcpyDeviceToHost, A->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(A->ctx->stream));
  return RLS_OK;
}

// sum |A_ij|^2 over the m x n window (padding rows excluded)
template <typename T>
__global__ void __launch_bounds__(RED_BLOCK) frob2_kernel(const T* __restrict__ A, int64_t m, int64_t n, int64_t ld,
                                                          double* partials, unsigned* ticket, double* out) {
  double acc[1] = {0.0};
  const int64_t total = m * n;
  for (int64_t k = (int64_t)blockIdx.x * RED_BLOCK + threadIdx.x; k < total; k += (int64_t)gridDim.x * RED_BLOCK) {
    int64_t j = k / m, i = k - j * m;
    acc[0] += Elem<T>::abs2(A[i + j * ld]);
  }
  grid_reduce_finalize<1, RED_BLOCK>(acc, partials, ticket, [=](double* t) { out[0] = t[0]; });
}

extern "C" int32_t rls_mat_frob2(rls_mat_t A, double* out) {
  RLS_CHECK_ARG(A && out, "NULL argument");
  rls_ctx_s* c = A->ctx;
  RlsDeviceGuard g(c->device);
  if (A->m == 0 || A->n == 0) { *out = 0.0; return RLS_OK; }
  int grid = c->sm_count * 8;
  const int64_t fast = A->layout == RLS_LAYOUT_ROWMAJOR ? A->n : A->m, slow = A->layout == RLS_LAYOUT_ROWMAJOR ? A->m : A->n;
  if (A->dtype == RLS_C32)
    frob2_kernel<float2><<<grid, RED_BLOCK, 0, c->stream>>>((const float2*)A->d, fast, slow, A->ld, c->red_partials, c->red_ticket, c->red_out);
  else
    frob2_kernel<float><<<grid, RED_BLOCK, 0, c->stream>>>((const float*)A->d, fast, slow, A->ld, c->red_partials, c->red_ticket, c->red_out);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  RLS_CUDA(cudaMemcpyAsync(c->red_out_host, c->red_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  double s = c->red_out_host[0];
  // on a row shard this is the local partial; the host sums it across ranks
  *out = s;
  return RLS_OK;
}

extern "C" int32_t rls_mat_relayout(rls_mat_t A, int32_t layout, rls_mat_t* out) {
  RLS_CHECK_ARG(A && out, "NULL argument");
  RLS_CHECK_ARG(layout == RLS_LAYOUT_ROWMAJOR || layout == RLS_LAYOUT_COLMAJOR, "relayout: layout must be row- or column-major");
  rls_ctx_s* c = A->ctx;
  RlsDeviceGuard g(c->device);
  RlsNvtxRange nvtx("rls: device re-layout of A");
  rls_mat_s* B = nullptr;
  RLS_TRY(rls_mat_create_layout(c, A->dtype, A->m, A->n, nullptr, A->m, layout, &B));
  const size_t es = rls_elem_size(A->dtype);
  if (A->m > 0 && A->n > 0) {
    if (B->layout == A->layout) {
      const int64_t rows = A->layout == RLS_LAYOUT_ROWMAJOR ? A->m : A->n, width = A->layout == RLS_LAYOUT_ROWMAJOR ? A->n : A->m;
      cudaError_t e = cudaMemcpy2DAsync(B->d, (size_t)B->ld * es, A->d, (size_t)A->ld * es, (size_t)width * es, (size_t)rows,
                                        cudaMemcpyDeviceToDevice, c->stream);
      if (e != cudaSuccess) { rls_mat_destroy(B); rls_set_error("relayout: %s", cudaGetErrorString(e)); return RLS_ERR_CUDA; }
    } else if (A->layout == RLS_LAYOUT_COLMAJOR) {
      // in(r = column j, c = row i) at j*ld + i  ->  out at i*ldB + j
      if (A->dtype == RLS_C32) launch_transpose<float2>(c, A->d, A->ld, B->d, B->ld, A->n, A->m);
      else launch_transpose<float>(c, A->d, A->ld, B->d, B->ld, A->n, A->m);
    } else {
      // in(r = row i, c = column j) at i*ld + j  ->  out at j*ldB + i
      if (A->dtype == RLS_C32) launch_transpose<float2>(c, A->d, A->ld, B->d, B->ld, A->m, A->n);
      else launch_transpose<float>(c, A->d, A->ld, B->d, B->ld, A->m, A->n);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { rls_mat_destroy(B); rls_set_error("relayout: %s", cudaGetErrorString(e)); return RLS_ERR_CUDA; }
  }
  *out = B;
  return RLS_OK;
}
