// rls_linop.cu — matrix-free system operators of the documentation's examples:
//   SamplingOp(T; pattern, shape)        y = x[pattern]                (docs/src/literate/examples/compressed_sensing.jl:23)
//   FFTOp(T; shape, shift, unitary)      y = fftshift(fft(ifftshift(x))) / sqrt(N)   (LinearOperatorCollection.jl 2.x, cuFFT)
//   outer ∘ inner                        e.g. SamplingOp ∘ FFTOp = undersampled Fourier encoding (test/testSolvers.jl:67-82 builds
//                                        the same operator as a dense matrix)
// with mul!(y, A, x), mul!(x, adjoint(A), y) and the lazy normal operator A'A as an rls_normal_t, so that the fused solver
// iterations and proximal maps of this library run unchanged on problems whose A is never stored.
// Gather / scatter / shift are plain HBM-bound index kernels; the transform itself is cuFFT (a library call, loaded with
// dlopen on first use like NCCL — the product has no link-time dependency on it).
// LinearOperatorCollection is not vendored in /root/reference (compat bound only, Project.toml:33): FFTOp is restated from
// its published definition — centred unitary DFT; the adjoint is the inverse transform with the same shifts and factor.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "rls_common.cuh"

enum { LINOP_SAMPLING = 0, LINOP_FFT = 1, LINOP_COMPOSE = 2 };
constexpr int LINOP_MAXD = 4;

struct rls_linop_s {
  rls_ctx_s* ctx = nullptr;
  int kind = 0;
  int32_t dtype = RLS_C32;
  int64_t m = 0, n = 0;                 // y = A x: x has n elements, y has m
  // sampling
  int64_t* idx = nullptr;               // device, 0-based
  // fft
  int ndims = 0;
  int64_t shape[LINOP_MAXD] = {1, 1, 1, 1};
  int shift = 1, unitary = 1;
  int plan = 0;
  bool have_plan = false;
  float2* tmp = nullptr;                // n complex: the (shifted) transform buffer
  // compose
  rls_linop_s* outer = nullptr;
  rls_linop_s* inner = nullptr;
  void* mid = nullptr;                  // inner->m elements
  std::atomic<int> refs{1};
};

namespace {

// ---- cuFFT through dlopen ------------------------------------------------------------------------------------------
typedef int (*cufftPlanMany_t)(int*, int, int*, int*, int, int, int*, int, int, int, int);
typedef int (*cufftSetStream_t)(int, cudaStream_t);
typedef int (*cufftExecC2C_t)(int, float2*, float2*, int);
typedef int (*cufftDestroy_t)(int);
struct CufftApi {
  cufftPlanMany_t plan_many = nullptr;
  cufftSetStream_t set_stream = nullptr;
  cufftExecC2C_t exec_c2c = nullptr;
  cufftDestroy_t destroy = nullptr;
  bool ok = false;
};
constexpr int CUFFT_C2C_TYPE = 0x29, CUFFT_FWD = -1, CUFFT_INV = 1;

CufftApi* cufft_api() {
  static CufftApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  const char* names[] = {"libcufft.so.11", "libcufft.so.12", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return nullptr;
  api.plan_many = (cufftPlanMany_t)dlsym(h, "cufftPlanMany");
  api.set_stream = (cufftSetStream_t)dlsym(h, "cufftSetStream");
  api.exec_c2c = (cufftExecC2C_t)dlsym(h, "cufftExecC2C");
  api.destroy = (cufftDestroy_t)dlsym(h, "cufftDestroy");
  api.ok = api.plan_many && api.set_stream && api.exec_c2c && api.destroy;
  return api.ok ? &api : nullptr;
}

constexpr int LB = 256;
int lgrid(const rls_ctx_s* c, int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + LB - 1) / LB, (int64_t)c->sm_count * 8));
}

template <typename T>
__global__ void __launch_bounds__(LB) gather_kernel(const T* __restrict__ x, const int64_t* __restrict__ idx, int64_t m, T* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * LB + threadIdx.x; i < m; i += (int64_t)gridDim.x * LB) y[i] = x[idx[i]];
}
// x = 0 first (memset), then x[idx[i]] = y[i]; the pattern has no duplicates (checked at creation)
template <typename T>
__global__ void __launch_bounds__(LB) scatter_kernel(const T* __restrict__ y, const int64_t* __restrict__ idx, int64_t m, T* __restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * LB + threadIdx.x; i < m; i += (int64_t)gridDim.x * LB) x[idx[i]] = y[i];
}

struct ShiftGeom {
  int ndims;
  int64_t shape[LINOP_MAXD], add[LINOP_MAXD];
  int64_t n;
};
// dst[k] = scale * src[(k + add) mod shape] per dimension (column-major).  ifftshift: add = N ÷ 2; fftshift: add = (N+1) ÷ 2.
__global__ void __launch_bounds__(LB) shift_scale_kernel(const float2* __restrict__ src, float2* __restrict__ dst, ShiftGeom g, float scale) {
  for (int64_t k = (int64_t)blockIdx.x * LB + threadIdx.x; k < g.n; k += (int64_t)gridDim.x * LB) {
    int64_t rem = k, off = 0, stride = 1;
#pragma unroll
    for (int d = 0; d < LINOP_MAXD; ++d) {
      if (d < g.ndims) {
        const int64_t c = rem % g.shape[d];
        rem /= g.shape[d];
        int64_t s = c + g.add[d];
        if (s >= g.shape[d]) s -= g.shape[d];
        off += s * stride;
        stride *= g.shape[d];
      }
    }
    const float2 v = src[off];
    dst[k] = make_float2(__fmul_rn(v.x, scale), __fmul_rn(v.y, scale));
  }
}

int32_t fft_apply(rls_linop_s* L, const void* in, void* out, bool adjoint, cudaStream_t st) {
  CufftApi* api = cufft_api();
  if (!api) { rls_set_error("FFTOp: libcufft could not be loaded"); return RLS_ERR_UNSUPPORTED; }
  rls_ctx_s* c = L->ctx;
  if (L->n == 0) return RLS_OK;
  ShiftGeom gi{}, go{};
  gi.ndims = go.ndims = L->ndims;
  gi.n = go.n = L->n;
  for (int d = 0; d < L->ndims; ++d) {
    gi.shape[d] = go.shape[d] = L->shape[d];
    gi.add[d] = L->shift ? L->shape[d] / 2 : 0;          // ifftshift on the way in
    go.add[d] = L->shift ? (L->shape[d] + 1) / 2 : 0;    // fftshift on the way out
  }
  const float factor = L->unitary ? (float)(1.0 / sqrt((double)L->n)) : 1.f;
  // tmp = ifftshift(in); tmp = F tmp (in place); out = factor * fftshift(tmp)
  shift_scale_kernel<<<lgrid(c, L->n), LB, 0, st>>>((const float2*)in, L->tmp, gi, 1.f);
  if (api->set_stream(L->plan, st) != 0) { rls_set_error("cufftSetStream failed"); return RLS_ERR_CUDA; }
  const int r = api->exec_c2c(L->plan, L->tmp, L->tmp, adjoint ? CUFFT_INV : CUFFT_FWD);
  if (r != 0) { rls_set_error("cufftExecC2C failed with %d", r); return RLS_ERR_CUDA; }
  shift_scale_kernel<<<lgrid(c, L->n), LB, 0, st>>>(L->tmp, (float2*)out, go, factor);
  c->launches += 2;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

int32_t linop_apply(rls_linop_s* L, const void* in, void* out, bool adjoint, cudaStream_t st) {
  rls_ctx_s* c = L->ctx;
  switch (L->kind) {
    case LINOP_SAMPLING:
      if (!adjoint) {
        if (L->m == 0) return RLS_OK;
        if (L->dtype == RLS_C32) gather_kernel<float2><<<lgrid(c, L->m), LB, 0, st>>>((const float2*)in, L->idx, L->m, (float2*)out);
        else gather_kernel<float><<<lgrid(c, L->m), LB, 0, st>>>((const float*)in, L->idx, L->m, (float*)out);
        c->launches++;
      } else {
        if (L->n == 0) return RLS_OK;
        RLS_CUDA(cudaMemsetAsync(out, 0, (size_t)L->n * rls_elem_size(L->dtype), st));
        if (L->m == 0) return RLS_OK;
        if (L->dtype == RLS_C32) scatter_kernel<float2><<<lgrid(c, L->m), LB, 0, st>>>((const float2*)in, L->idx, L->m, (float2*)out);
        else scatter_kernel<float><<<lgrid(c, L->m), LB, 0, st>>>((const float*)in, L->idx, L->m, (float*)out);
        c->launches++;
      }
      RLS_CUDA(cudaGetLastError());
      return RLS_OK;
    case LINOP_FFT:
      return fft_apply(L, in, out, adjoint, st);
    case LINOP_COMPOSE:
      if (!adjoint) {   // y = outer (inner x)
        RLS_TRY(linop_apply(L->inner, in, L->mid, false, st));
        return linop_apply(L->outer, L->mid, out, false, st);
      }
      RLS_TRY(linop_apply(L->outer, in, L->mid, true, st));
      return linop_apply(L->inner, L->mid, out, true, st);
  }
  rls_set_error("unknown operator kind");
  return RLS_ERR_INVALID;
}

void linop_release(rls_linop_s* L) {
  if (!L || L->refs.fetch_sub(1) != 1) return;
  rls_ctx_s* c = L->ctx;
  {
    RlsDeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(L->idx);
    cudaFree(L->tmp);
    cudaFree(L->mid);
    if (L->have_plan) { CufftApi* api = cufft_api(); if (api) api->destroy(L->plan); }
  }
  linop_release(L->outer);
  linop_release(L->inner);
  delete L;
  rls_ctx_release(c);
}

// AHA x = A'(A x) for the normal operator built on a structured operator; `user` = { L, scratch m-vector }
struct LinopNormal {
  rls_linop_s* L;
  void* y;
};
int32_t linop_normal_apply(void* user, const void* x, void* out, void* stream) {
  LinopNormal* u = (LinopNormal*)user;
  // SamplingOp alone: A'A is a diagonal mask — still gather + scatter here (two n-sized passes), no special case needed
  RLS_TRY(linop_apply(u->L, x, u->y, false, (cudaStream_t)stream));
  return linop_apply(u->L, u->y, out, true, (cudaStream_t)stream);
}
void linop_normal_release(void* user) {
  LinopNormal* u = (LinopNormal*)user;
  if (!u) return;
  {
    RlsDeviceGuard g(u->L->ctx->device);
    cudaStreamSynchronize(u->L->ctx->stream);
    cudaFree(u->y);
  }
  linop_release(u->L);
  delete u;
}

}  // namespace

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
// SamplingOp(T; pattern, shape): pattern = 1-based linear indices into the n = prod(shape) elements, strictly increasing or
// not, but without duplicates (compressed_sensing.jl:19-23 sorts a random third of eachindex(image))
extern "C" int32_t rls_linop_sampling_create(rls_ctx_t ctx, int32_t dtype, int64_t n, int64_t npattern, const int64_t* pattern_1based,
                                             rls_linop_t* out) {
  RLS_CHECK_ARG(ctx && out && (npattern == 0 || pattern_1based), "NULL argument");
  RLS_CHECK_ARG(dtype == RLS_F32 || dtype == RLS_C32, "unsupported element type %d", dtype);
  RLS_CHECK_ARG(n >= 0 && npattern >= 0, "negative size");
  std::vector<int64_t> h((size_t)npattern);
  std::vector<char> seen((size_t)n, 0);
  for (int64_t i = 0; i < npattern; ++i) {
    const int64_t k = pattern_1based[i] - 1;
    RLS_CHECK_ARG(k >= 0 && k < n, "SamplingOp: pattern[%lld] = %lld is outside 1..%lld", (long long)i + 1, (long long)pattern_1based[i], (long long)n);
    RLS_CHECK_ARG(!seen[(size_t)k], "SamplingOp: index %lld appears twice in the pattern", (long long)pattern_1based[i]);
    seen[(size_t)k] = 1;
    h[(size_t)i] = k;
  }
  RlsDeviceGuard g(ctx->device);
  rls_linop_s* L = new rls_linop_s();
  L->ctx = ctx; L->kind = LINOP_SAMPLING; L->dtype = dtype; L->m = npattern; L->n = n;
  rls_ctx_retain(ctx);
  if (npattern > 0) {
    if (cudaMalloc(&L->idx, (size_t)npattern * 8) != cudaSuccess) { cudaGetLastError(); linop_release(L); rls_set_error("out of device memory"); return RLS_ERR_NOMEM; }
    if (cudaMemcpyAsync(L->idx, h.data(), (size_t)npattern * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { linop_release(L); rls_set_error("CUDA copy of the sampling pattern failed"); return RLS_ERR_CUDA; }
  }
  *out = L;
  return RLS_OK;
}

// FFTOp(ComplexF32; shape, shift=true, unitary=true) — the n-dimensional DFT of the column-major array of that shape
extern "C" int32_t rls_linop_fft_create(rls_ctx_t ctx, int32_t ndims, const int64_t* shape, int32_t shift, int32_t unitary, rls_linop_t* out) {
  RLS_CHECK_ARG(ctx && out && shape, "NULL argument");
  RLS_CHECK_ARG(ndims >= 1 && ndims <= LINOP_MAXD, "FFTOp: 1..%d dimensions", LINOP_MAXD);
  int64_t n = 1;
  for (int d = 0; d < ndims; ++d) {
    RLS_CHECK_ARG(shape[d] >= 1 && shape[d] <= 0x7fffffff, "FFTOp: bad extent %lld", (long long)shape[d]);
    n *= shape[d];
  }
  CufftApi* api = cufft_api();
  if (!api) { rls_set_error("FFTOp: libcufft could not be loaded"); return RLS_ERR_UNSUPPORTED; }
  RlsDeviceGuard g(ctx->device);
  rls_linop_s* L = new rls_linop_s();
  L->ctx = ctx; L->kind = LINOP_FFT; L->dtype = RLS_C32; L->m = n; L->n = n;
  L->ndims = ndims; L->shift = shift ? 1 : 0; L->unitary = unitary ? 1 : 0;
  for (int d = 0; d < ndims; ++d) L->shape[d] = shape[d];
  rls_ctx_retain(ctx);
  // cuFFT is row-major: the slowest dimension first = the column-major shape reversed; singleton dimensions dropped
  int dims[LINOP_MAXD], rank = 0;
  for (int d = ndims - 1; d >= 0; --d)
    if (shape[d] > 1) dims[rank++] = (int)shape[d];
  if (rank == 0) { dims[0] = 1; rank = 1; }
  if (rank > 3) { linop_release(L); rls_set_error("FFTOp: cuFFT plans have at most 3 non-singleton dimensions"); return RLS_ERR_UNSUPPORTED; }
  const int r = api->plan_many(&L->plan, rank, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2C_TYPE, 1);
  if (r != 0) { linop_release(L); rls_set_error("cufftPlanMany failed with %d", r); return RLS_ERR_CUDA; }
  L->have_plan = true;
  if (cudaMalloc(&L->tmp, (size_t)std::max<int64_t>(n, 1) * 8) != cudaSuccess) { cudaGetLastError(); linop_release(L); rls_set_error("out of device memory"); return RLS_ERR_NOMEM; }
  *out = L;
  return RLS_OK;
}

// A = outer ∘ inner (ProdOp / `*` of LinearOperators): y = outer (inner x)
extern "C" int32_t rls_linop_compose(rls_linop_t outer, rls_linop_t inner, rls_linop_t* out) {
  RLS_CHECK_ARG(outer && inner && out, "NULL argument");
  RLS_CHECK_ARG(outer->ctx == inner->ctx, "operators live on different contexts");
  RLS_CHECK_ARG(outer->n == inner->m && outer->dtype == inner->dtype, "compose: outer takes %lld elements, inner produces %lld (or the element types differ)",
                (long long)outer->n, (long long)inner->m);
  RlsDeviceGuard g(outer->ctx->device);
  rls_linop_s* L = new rls_linop_s();
  L->ctx = outer->ctx; L->kind = LINOP_COMPOSE; L->dtype = outer->dtype; L->m = outer->m; L->n = inner->n;
  rls_ctx_retain(L->ctx);
  if (cudaMalloc(&L->mid, (size_t)std::max<int64_t>(inner->m, 1) * rls_elem_size(L->dtype)) != cudaSuccess) {
    cudaGetLastError(); linop_release(L); rls_set_error("out of device memory"); return RLS_ERR_NOMEM;
  }
  L->outer = outer; L->inner = inner;
  outer->refs.fetch_add(1); inner->refs.fetch_add(1);
  *out = L;
  return RLS_OK;
}

extern "C" int32_t rls_linop_destroy(rls_linop_t L) {
  linop_release(L);
  return RLS_OK;
}

extern "C" int32_t rls_linop_shape(rls_linop_t L, int64_t* m, int64_t* n, int32_t* dtype) {
  RLS_CHECK_ARG(L, "NULL argument");
  if (m) *m = L->m;
  if (n) *n = L->n;
  if (dtype) *dtype = L->dtype;
  return RLS_OK;
}

// mul!(y, A, x)
extern "C" int32_t rls_linop_mul(rls_linop_t L, rls_vec_t x, rls_vec_t y) {
  RLS_CHECK_ARG(L && x && y, "NULL argument");
  RLS_CHECK_ARG(x->len == L->n && y->len == L->m && x->dtype == L->dtype && y->dtype == L->dtype, "mul!: operator is %lldx%lld", (long long)L->m, (long long)L->n);
  RLS_CHECK_ARG(x->d != y->d, "mul!: x and y must not alias");
  RlsDeviceGuard g(L->ctx->device);
  return linop_apply(L, x->d, y->d, false, L->ctx->stream);
}

// mul!(x, adjoint(A), y)
extern "C" int32_t rls_linop_mul_adjoint(rls_linop_t L, rls_vec_t y, rls_vec_t x) {
  RLS_CHECK_ARG(L && x && y, "NULL argument");
  RLS_CHECK_ARG(x->len == L->n && y->len == L->m && x->dtype == L->dtype && y->dtype == L->dtype, "mul!: operator is %lldx%lld", (long long)L->m, (long long)L->n);
  RLS_CHECK_ARG(x->d != y->d, "mul!: x and y must not alias");
  RlsDeviceGuard g(L->ctx->device);
  return linop_apply(L, y->d, x->d, true, L->ctx->stream);
}

// normalOperator(A) for a structured operator: the lazy A'(A x)
extern "C" int32_t rls_normal_from_linop(rls_linop_t L, rls_normal_t* out) {
  RLS_CHECK_ARG(L && out, "NULL argument");
  RlsDeviceGuard g(L->ctx->device);
  LinopNormal* u = new LinopNormal{L, nullptr};
  if (cudaMalloc(&u->y, (size_t)std::max<int64_t>(L->m, 1) * rls_elem_size(L->dtype)) != cudaSuccess) {
    cudaGetLastError(); delete u; rls_set_error("out of device memory"); return RLS_ERR_NOMEM;
  }
  L->refs.fetch_add(1);
  const char* nm = L->kind == LINOP_SAMPLING ? "A'A of SamplingOp (gather + scatter)" : L->kind == LINOP_FFT ? "A'A of FFTOp (cuFFT forward + inverse)"
                                                                                     : "A'A of a composed operator (inner, outer, outer', inner')";
  int32_t st = rls_normal_from_function(L->ctx, L->dtype, L->n, linop_normal_apply, linop_normal_release, u, nm, out);
  if (st != RLS_OK) linop_normal_release(u);
  return st;
}
