// rls_prox.cu — proximal maps as fused elementwise / stencil kernels:
//   L1 soft-threshold, L2 scaling, positivity / real projections (one elementwise kernel),
//   L21 group shrinkage (one thread per strided group),
//   TV by Fast Gradient Projection: per inner iteration one divergence(gather)+axpy kernel
//   and one gradient+clip+linear-combination kernel (ProxTV.jl:89-125), with the
//   reference's pq/rs/pqOld buffer rotation done by pointer on the host.
// All HBM/L2-bound elementwise work; nothing here is shaped into a GEMM.
#include "rls_prox.cuh"

namespace {
constexpr int PB = 256;

static inline int ew_grid(const rls_ctx_s* c, int64_t n) {
  int64_t g = (n + PB - 1) / PB;
  int64_t cap = (int64_t)c->sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename T>
__global__ void __launch_bounds__(PB) prox_ew_kernel(T* __restrict__ x, int64_t n, int kind, float lam,
                                                     const float* __restrict__ lam_dev, int proj_mask,
                                                     const int* __restrict__ gate) {
  if (gate && *gate) return;
  const float thr = lam_dev ? *lam_dev : lam;
  for (int64_t i = (int64_t)blockIdx.x * PB + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB) {
    T v = x[i];
    v = prox_elementwise(v, kind, thr);
    if (proj_mask) v = proj_elem(v, proj_mask);
    x[i] = v;
  }
}

// ProxL21.jl:30-35: L = length(x) ÷ slices; group j = x[j:L:end] (slices elements, one more for the first
// length(x) - L*slices groups when the division leaves a remainder) ; x *= max((g-λ)/g, 0) (0/0 -> NaN kept)
template <typename T>
__global__ void __launch_bounds__(PB) prox_l21_kernel(T* __restrict__ x, int64_t L, int64_t n, float lam,
                                                      const float* __restrict__ lam_dev, const int* __restrict__ gate) {
  if (gate && *gate) return;
  const float thr = lam_dev ? *lam_dev : lam;
  for (int64_t j = (int64_t)blockIdx.x * PB + threadIdx.x; j < L; j += (int64_t)gridDim.x * PB) {
    double s = 0.0;
    for (int64_t e = j; e < n; e += L) s += Elem<T>::abs2(x[e]);
    const float g = (float)sqrt(s);
    const float q = fdiv(fsub(g, thr), g);
    const float ff = isnan(q) ? q : fmaxf(q, 0.f);  // Julia's max(NaN,0) is NaN (0/0 group, quirk 7); fmaxf would hide it
    for (int64_t e = j; e < n; e += L) x[e] = Elem<T>::scale(x[e], ff);
  }
}

template <typename T>
__global__ void __launch_bounds__(PB) grad_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, GradGeom G,
                                                      const int* __restrict__ gate) {
  if (gate && *gate) return;
  for (int64_t e = (int64_t)blockIdx.x * PB + threadIdx.x; e < G.rows; e += (int64_t)gridDim.x * PB)
    out[e] = grad_fwd_elem(x, G, e);
}

// res[pix] = base[pix] (or 0); for each block k: res = a*(A_k' g_k) + res   (5-arg mul! of a vcat'd operator)
template <typename T>
__global__ void __launch_bounds__(PB) grad_t_axpy_kernel(const T* __restrict__ g, const T* base, T* res, float a,
                                                         const float* __restrict__ a_dev, float a_sign, GradGeom G,
                                                         const int* __restrict__ gate) {
  if (gate && *gate) return;
  const float av = a_sign * (a_dev ? *a_dev : a);
  for (int64_t p = (int64_t)blockIdx.x * PB + threadIdx.x; p < G.npix; p += (int64_t)gridDim.x * PB) {
    T acc = base ? base[p] : Elem<T>::zero();
    for (int k = 0; k < G.ndirs; ++k) acc = Elem<T>::add(Elem<T>::scale(grad_t_block(g, G, k, p), av), acc);
    res[p] = acc;
  }
}

// FGP dual step (ProxTV.jl:109-120), one thread per dual element:
//   v  = rs + (1/(8λ)) * (∇ xTmp)      (pq aliases rs' storage)
//   v /= max(1, |v|)
//   rs_new = t3*v - t2*pqOld           (written into pqTmp's storage)
template <typename T>
__global__ void __launch_bounds__(PB) tv_dual_kernel(const T* __restrict__ xtmp, T* rs_pq, const T* __restrict__ pq_old,
                                                     T* __restrict__ rs_new, float lam, const float* __restrict__ lam_dev,
                                                     float t2, float t3, GradGeom G, const int* __restrict__ gate) {
  if (gate && *gate) return;
  const float thr = lam_dev ? *lam_dev : lam;
  const float inv8 = fdiv(1.f, fmul(8.f, thr));
  for (int64_t e = (int64_t)blockIdx.x * PB + threadIdx.x; e < G.rows; e += (int64_t)gridDim.x * PB) {
    T gr = grad_fwd_elem(xtmp, G, e);
    T v = Elem<T>::add(Elem<T>::scale(gr, inv8), rs_pq[e]);
    v = Elem<T>::divr(v, fmaxf(1.f, Elem<T>::abs(v)));
    rs_pq[e] = v;
    rs_new[e] = Elem<T>::sub(Elem<T>::scale(v, t3), Elem<T>::scale(pq_old[e], t2));
  }
}

template <typename T>
int32_t tv_fgp(rls_ctx_s* c, T* x, const GradGeom& G, float lam, const float* lam_dev, int iters, const int* gate,
               TvWork* w) {
  const size_t es = sizeof(T);
  T* pq = (T*)w->buf[0];
  T* rs = (T*)w->buf[1];
  T* pqOld = (T*)w->buf[2];
  T* xtmp = (T*)w->xtmp;
  for (int b = 0; b < 3; ++b) RLS_CUDA(cudaMemsetAsync(w->buf[b], 0, (size_t)(G.rows > 0 ? G.rows : 1) * es, c->stream));
  const int gp = ew_grid(c, G.npix), gr = ew_grid(c, G.rows);
  float t = 1.f;
  for (int it = 0; it < iters; ++it) {
    T* pqTmp = pqOld;
    pqOld = pq;
    pq = rs;
    grad_t_axpy_kernel<T><<<gp, PB, 0, c->stream>>>(rs, x, xtmp, lam, lam_dev, -1.f, G, gate);
    float tOld = t;
    t = (1.f + sqrtf(1.f + 4.f * (tOld * tOld))) / 2.f;
    float t2 = (tOld - 1.f) / t;
    float t3 = 1.f + t2;
    rs = pqTmp;
    if (G.rows > 0) tv_dual_kernel<T><<<gr, PB, 0, c->stream>>>(xtmp, pq, pqOld, rs, lam, lam_dev, t2, t3, G, gate);
    c->launches += 2;
  }
  grad_t_axpy_kernel<T><<<gp, PB, 0, c->stream>>>(pq, x, x, lam, lam_dev, -1.f, G, gate);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------
// geometry / workspace
// ------------------------------------------------------------------------------------
int32_t rls_make_grad_geom(int32_t ndims, const int64_t* shape, int32_t ndirs, const int32_t* dims_1based, GradGeom* g) {
  RLS_CHECK_ARG(g && shape, "NULL argument");
  RLS_CHECK_ARG(ndims >= 1 && ndims <= RLS_MAX_TV_DIMS, "TV/GradientOp supports 1..%d dimensions, got %d", RLS_MAX_TV_DIMS, ndims);
  RLS_CHECK_ARG(ndirs >= 0 && ndirs <= RLS_MAX_TV_DIMS, "bad number of gradient directions %d", ndirs);
  memset(g, 0, sizeof(*g));
  g->ndims = ndims;
  g->ndirs = ndirs;
  int64_t st = 1;
  for (int d = 0; d < ndims; ++d) {
    RLS_CHECK_ARG(shape[d] >= 1, "shape[%d] = %lld", d, (long long)shape[d]);
    g->shape[d] = shape[d];
    g->stride[d] = st;
    st *= shape[d];
  }
  for (int d = ndims; d < RLS_MAX_TV_DIMS; ++d) { g->shape[d] = 1; g->stride[d] = st; }
  g->npix = st;
  int64_t off = 0;
  for (int k = 0; k < ndirs; ++k) {
    int a = dims_1based[k] - 1;
    RLS_CHECK_ARG(a >= 0 && a < ndims, "gradient direction %d outside 1..%d", dims_1based[k], ndims);
    g->dim[k] = a;
    g->off[k] = off;
    int64_t rs = 1;
    for (int d = 0; d < RLS_MAX_TV_DIMS; ++d) {
      g->rstride[k][d] = rs;
      int64_t sd = (d < ndims) ? g->shape[d] : 1;
      rs *= (d == a) ? (sd - 1) : sd;
    }
    off += rs;
  }
  for (int k = ndirs; k <= RLS_MAX_TV_DIMS; ++k) g->off[k] = off;
  g->rows = off;
  return RLS_OK;
}

int32_t rls_tv_work_ensure(rls_ctx_s* c, TvWork* w, int32_t dtype, const GradGeom& g) {
  if (w->buf[0] && w->rows == g.rows && w->npix == g.npix && w->dtype == dtype) return RLS_OK;
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  rls_tv_work_free(w);
  size_t es = rls_elem_size(dtype);
  for (int b = 0; b < 3; ++b) RLS_CUDA(cudaMalloc(&w->buf[b], (size_t)(g.rows > 0 ? g.rows : 1) * es));
  RLS_CUDA(cudaMalloc(&w->xtmp, (size_t)(g.npix > 0 ? g.npix : 1) * es));
  w->rows = g.rows;
  w->npix = g.npix;
  w->dtype = dtype;
  return RLS_OK;
}

void rls_tv_work_free(TvWork* w) {
  for (int b = 0; b < 3; ++b) {
    if (w->buf[b]) cudaFree(w->buf[b]);
    w->buf[b] = nullptr;
  }
  if (w->xtmp) cudaFree(w->xtmp);
  w->xtmp = nullptr;
  w->rows = w->npix = 0;
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
int32_t rls_proj_launch(rls_ctx_s* c, int32_t dtype, void* x, int64_t n, int proj_mask, const int* gate) {
  if (n == 0 || proj_mask == 0) return RLS_OK;
  if (dtype == RLS_C32)
    prox_ew_kernel<float2><<<ew_grid(c, n), PB, 0, c->stream>>>((float2*)x, n, RLS_REG_NONE, 0.f, nullptr, proj_mask, gate);
  else
    prox_ew_kernel<float><<<ew_grid(c, n), PB, 0, c->stream>>>((float*)x, n, RLS_REG_NONE, 0.f, nullptr, proj_mask, gate);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

int32_t rls_grad_fwd_launch(rls_ctx_s* c, int32_t dtype, const void* x, void* out, const GradGeom& g, const int* gate) {
  if (g.rows == 0) return RLS_OK;
  if (dtype == RLS_C32)
    grad_fwd_kernel<float2><<<ew_grid(c, g.rows), PB, 0, c->stream>>>((const float2*)x, (float2*)out, g, gate);
  else
    grad_fwd_kernel<float><<<ew_grid(c, g.rows), PB, 0, c->stream>>>((const float*)x, (float*)out, g, gate);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

int32_t rls_grad_t_axpy_launch(rls_ctx_s* c, int32_t dtype, const void* g, const void* base, void* res, float a,
                               const float* a_dev, float a_sign, const GradGeom& geom, const int* gate) {
  if (geom.npix == 0) return RLS_OK;
  if (dtype == RLS_C32)
    grad_t_axpy_kernel<float2><<<ew_grid(c, geom.npix), PB, 0, c->stream>>>((const float2*)g, (const float2*)base, (float2*)res, a, a_dev, a_sign, geom, gate);
  else
    grad_t_axpy_kernel<float><<<ew_grid(c, geom.npix), PB, 0, c->stream>>>((const float*)g, (const float*)base, (float*)res, a, a_dev, a_sign, geom, gate);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

int32_t rls_prox_launch(rls_ctx_s* c, int32_t dtype, void* x, int64_t n, const rls_reg_desc* reg, float lam,
                        const float* lam_dev, const int* gate, TvWork* tv) {
  if (n == 0) return RLS_OK;
  switch (reg->kind) {
    case RLS_REG_NONE:
      return RLS_OK;
    case RLS_REG_L1:
    case RLS_REG_L2:
      if (dtype == RLS_C32)
        prox_ew_kernel<float2><<<ew_grid(c, n), PB, 0, c->stream>>>((float2*)x, n, reg->kind, lam, lam_dev, 0, gate);
      else
        prox_ew_kernel<float><<<ew_grid(c, n), PB, 0, c->stream>>>((float*)x, n, reg->kind, lam, lam_dev, 0, gate);
      c->launches++;
      break;
    case RLS_REG_L21: {
      RLS_CHECK_ARG(reg->slices >= 1, "L21: slices must be >= 1");
      int64_t L = n / reg->slices;
      if (n == 0) return RLS_OK;
      // upstream x[i:0:end] throws (zero step) when there are more slices than elements
      RLS_CHECK_ARG(L >= 1, "L21: %lld slices exceed the %lld elements of x", (long long)reg->slices, (long long)n);
      if (dtype == RLS_C32)
        prox_l21_kernel<float2><<<ew_grid(c, L), PB, 0, c->stream>>>((float2*)x, L, n, lam, lam_dev, gate);
      else
        prox_l21_kernel<float><<<ew_grid(c, L), PB, 0, c->stream>>>((float*)x, L, n, lam, lam_dev, gate);
      c->launches++;
      break;
    }
    case RLS_REG_TV: {
      GradGeom G;
      RLS_TRY(rls_make_grad_geom(reg->tv_ndims, reg->tv_shape, reg->tv_ndirs, reg->tv_dims, &G));
      RLS_CHECK_ARG(G.npix == n, "TV shape has %lld pixels but x has %lld elements", (long long)G.npix, (long long)n);
      TvWork local;
      TvWork* w = tv ? tv : &local;
      RLS_TRY(rls_tv_work_ensure(c, w, dtype, G));
      int32_t s = (dtype == RLS_C32) ? tv_fgp<float2>(c, (float2*)x, G, lam, lam_dev, reg->tv_iterations, gate, w)
                                     : tv_fgp<float>(c, (float*)x, G, lam, lam_dev, reg->tv_iterations, gate, w);
      if (!tv) {
        cudaStreamSynchronize(c->stream);
        rls_tv_work_free(&local);
      }
      return s;
    }
    case RLS_REG_NUCLEAR:
      // svtShape travels in tv_shape[0..1] (ProxNuclear.jl:15-19)
      return rls_prox_nuclear_launch(c, dtype, x, n, reg->tv_shape[0], reg->tv_shape[1], lam, lam_dev, gate);
    case RLS_REG_LLR: {
      // shape in tv_shape, blockSize in tv_dims, flags in tv_iterations (ProxLLR.jl:20-29).  randshift draws one shift of
      // the patch grid per prox! call (:55); here from a counter-based generator seeded by reg->slices, so that a solve
      // is reproducible (the reference uses the global RNG)
      RLS_CHECK_ARG(reg->tv_ndims >= 1 && reg->tv_ndims <= RLS_MAX_TV_DIMS, "LLR: bad dimensionality %d", reg->tv_ndims);
      int64_t block[RLS_MAX_TV_DIMS], shift[RLS_MAX_TV_DIMS];
      for (int d = 0; d < reg->tv_ndims; ++d) { block[d] = reg->tv_dims[d]; shift[d] = 0; }
      if (reg->tv_iterations & RLS_LLR_RANDSHIFT) {
        uint64_t z = (uint64_t)reg->slices * 0x9E3779B97F4A7C15ull + (++c->llr_calls) * 0xBF58476D1CE4E5B9ull;
        for (int d = 0; d < reg->tv_ndims; ++d) {
          z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;   // splitmix64
          shift[d] = 1 + (int64_t)(z % (uint64_t)(block[d] > 0 ? block[d] : 1));                             // rand(1:blockSize[d])
        }
      }
      return rls_prox_llr_launch(c, dtype, x, n, reg->tv_ndims, reg->tv_shape, block, shift, (reg->tv_iterations & RLS_LLR_OVERLAPPING) ? 1 : 0,
                                 lam, lam_dev, gate);
    }
    default:
      rls_set_error("unknown regularization kind %d", reg->kind);
      return RLS_ERR_INVALID;
  }
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
static int32_t prox_simple(rls_vec_t x, int kind, float lam, int64_t slices) {
  RLS_CHECK_ARG(x, "vec is NULL");
  RlsDeviceGuard g(x->ctx->device);
  rls_reg_desc r;
  memset(&r, 0, sizeof(r));
  r.kind = kind;
  r.slices = slices;
  return rls_prox_launch(x->ctx, x->dtype, x->d, x->len, &r, lam, nullptr, nullptr, nullptr);
}

extern "C" int32_t rls_prox_l1(rls_vec_t x, float lambda) { return prox_simple(x, RLS_REG_L1, lambda, 1); }
extern "C" int32_t rls_prox_l2(rls_vec_t x, float lambda) { return prox_simple(x, RLS_REG_L2, lambda, 1); }
extern "C" int32_t rls_prox_l21(rls_vec_t x, float lambda, int64_t slices) {
  RLS_CHECK_ARG(slices >= 1, "slices must be >= 1");
  return prox_simple(x, RLS_REG_L21, lambda, slices);
}

extern "C" int32_t rls_prox_tv(rls_vec_t x, float lambda, int32_t ndims, const int64_t* shape, int32_t ndirs,
                               const int32_t* dims_1based, int32_t iterations_tv) {
  RLS_CHECK_ARG(x && shape && (ndirs == 0 || dims_1based), "NULL argument");
  RLS_CHECK_ARG(ndims >= 1 && ndims <= RLS_MAX_TV_DIMS && ndirs >= 0 && ndirs <= RLS_MAX_TV_DIMS, "bad TV dimensionality");
  RlsDeviceGuard g(x->ctx->device);
  rls_reg_desc r;
  memset(&r, 0, sizeof(r));
  r.kind = RLS_REG_TV;
  r.tv_ndims = ndims;
  r.tv_ndirs = ndirs;
  for (int d = 0; d < ndims; ++d) r.tv_shape[d] = shape[d];
  for (int k = 0; k < ndirs; ++k) r.tv_dims[k] = dims_1based[k];
  r.tv_iterations = iterations_tv;
  return rls_prox_launch(x->ctx, x->dtype, x->d, x->len, &r, lambda, nullptr, nullptr, nullptr);
}

extern "C" int32_t rls_prox_positive(rls_vec_t x) {
  RLS_CHECK_ARG(x, "vec is NULL");
  RlsDeviceGuard g(x->ctx->device);
  return rls_proj_launch(x->ctx, x->dtype, x->d, x->len, RLS_PROJ_POSITIVE, nullptr);
}

extern "C" int32_t rls_prox_real(rls_vec_t x) {
  RLS_CHECK_ARG(x, "vec is NULL");
  RlsDeviceGuard g(x->ctx->device);
  return rls_proj_launch(x->ctx, x->dtype, x->d, x->len, RLS_PROJ_REAL, nullptr);
}

extern "C" int32_t rls_grad_rows(int32_t ndims, const int64_t* shape, int32_t ndirs, const int32_t* dims_1based, int64_t* rows) {
  RLS_CHECK_ARG(rows, "rows is NULL");
  GradGeom G;
  RLS_TRY(rls_make_grad_geom(ndims, shape, ndirs, dims_1based, &G));
  *rows = G.rows;
  return RLS_OK;
}

extern "C" int32_t rls_grad_apply(rls_vec_t img, rls_vec_t out, int32_t ndims, const int64_t* shape, int32_t ndirs,
                                  const int32_t* dims_1based) {
  RLS_CHECK_ARG(img && out, "NULL argument");
  GradGeom G;
  RLS_TRY(rls_make_grad_geom(ndims, shape, ndirs, dims_1based, &G));
  RLS_CHECK_ARG(img->len == G.npix && out->len == G.rows && img->dtype == out->dtype, "grad_apply: shape/dtype mismatch");
  RlsDeviceGuard g(img->ctx->device);
  return rls_grad_fwd_launch(img->ctx, img->dtype, img->d, out->d, G, nullptr);
}

extern "C" int32_t rls_grad_apply_t(rls_vec_t gvec, rls_vec_t out, int32_t ndims, const int64_t* shape, int32_t ndirs,
                                    const int32_t* dims_1based) {
  RLS_CHECK_ARG(gvec && out, "NULL argument");
  GradGeom G;
  RLS_TRY(rls_make_grad_geom(ndims, shape, ndirs, dims_1based, &G));
  RLS_CHECK_ARG(out->len == G.npix && gvec->len == G.rows && gvec->dtype == out->dtype, "grad_apply_t: shape/dtype mismatch");
  RlsDeviceGuard g(gvec->ctx->device);
  return rls_grad_t_axpy_launch(gvec->ctx, gvec->dtype, gvec->d, nullptr, out->d, 1.f, nullptr, 1.f, G, nullptr);
}
