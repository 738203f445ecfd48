// rls_prox.cuh — per-element proximal maps and projections as device functions, shared by
// the standalone prox kernels and the fused solver epilogues.  Arithmetic follows the
// reference operation by operation (every op individually rounded).
#pragma once
#include <float.h>

#include "rls_common.cuh"

#ifdef __CUDACC__
// ProxL1.jl:18-22   x = max(|x|-λ,0) * (x+ε) / (|x|+ε), ε = eps(Float32) on the real part
__device__ __forceinline__ float prox_l1_elem(float x, float lam) {
  const float eps = FLT_EPSILON;
  float ax = fabsf(x);
  float mg = fmaxf(fsub(ax, lam), 0.f);
  return fdiv(fmul(mg, fadd(x, eps)), fadd(ax, eps));
}
__device__ __forceinline__ float2 prox_l1_elem(float2 x, float lam) {
  const float eps = FLT_EPSILON;
  float ax = Elem<float2>::abs(x);
  float mg = fmaxf(fsub(ax, lam), 0.f);
  float den = fadd(ax, eps);
  return make_float2(fdiv(fmul(mg, fadd(x.x, eps)), den), fdiv(fmul(mg, x.y), den));
}

// ProxL2.jl:18-21   x *= 1/(1+2λ), factor in Float64, product rounded to T
__device__ __forceinline__ double prox_l2_factor(float lam) { return 1.0 / (1.0 + 2.0 * (double)lam); }
__device__ __forceinline__ float prox_l2_elem(float x, double f) { return (float)((double)x * f); }
__device__ __forceinline__ float2 prox_l2_elem(float2 x, double f) {
  return make_float2((float)((double)x.x * f), (float)((double)x.y * f));
}

// ProxReal.jl:16-19 / ProxPositive.jl:16-20 via Utils.jl:114-144
__device__ __forceinline__ float proj_elem(float x, int mask) {
  if (mask & RLS_PROJ_POSITIVE) return (x < 0.f) ? 0.f : x;
  return x;
}
__device__ __forceinline__ float2 proj_elem(float2 x, int mask) {
  if (mask & (RLS_PROJ_REAL | RLS_PROJ_POSITIVE)) x.y = 0.f;          // enfReal! (Positive calls it first)
  if ((mask & RLS_PROJ_POSITIVE) && x.x < 0.f) x = make_float2(0.f, x.y);  // re<0 -> im*i (im is 0 here)
  return x;
}

// elementwise prox dispatch used in fused epilogues (kind is warp-uniform)
template <typename T>
__device__ __forceinline__ T prox_elementwise(T x, int kind, float thr) {
  if (kind == RLS_REG_L1) return prox_l1_elem(x, thr);
  if (kind == RLS_REG_L2) return prox_l2_elem(x, prox_l2_factor(thr));
  return x;
}

__host__ __device__ inline bool rls_reg_is_elementwise(int kind) {
  return kind == RLS_REG_NONE || kind == RLS_REG_L1 || kind == RLS_REG_L2;
}
#endif

// description of a GradientOp (forward differences without boundary rows, stacked per dim;
// LinearOperatorCollection GradientOp = vcat of one operator per direction)
struct GradGeom {
  int ndims;                        // image dimensionality
  int ndirs;                        // number of difference directions (blocks)
  int64_t shape[RLS_MAX_TV_DIMS];   // column-major image shape
  int64_t stride[RLS_MAX_TV_DIMS];  // image element strides
  int dim[RLS_MAX_TV_DIMS];         // 0-based direction of block k
  int64_t off[RLS_MAX_TV_DIMS + 1]; // start of block k in the stacked output
  int64_t rstride[RLS_MAX_TV_DIMS][RLS_MAX_TV_DIMS];  // strides of block k's reduced shape
  int64_t npix;
  int64_t rows;
};

#ifdef __CUDACC__
// ---- GradientOp stencils ---------------------------------------------------------------
// (A_k' g_k)[pix] = (c_a < N_a-1 ? g_k[idx] : 0) - (c_a > 0 ? g_k[idx - rstride_a] : 0)
template <typename T>
__device__ __forceinline__ T grad_t_block(const T* __restrict__ g, const GradGeom& G, int k, int64_t pix) {
  const int a = G.dim[k];
  int64_t idx = 0, ca = 0, rem = pix;
#pragma unroll
  for (int d = RLS_MAX_TV_DIMS - 1; d >= 0; --d) {
    if (d < G.ndims) {
      int64_t c = rem / G.stride[d];
      rem -= c * G.stride[d];
      idx += c * G.rstride[k][d];
      if (d == a) ca = c;
    }
  }
  T r = Elem<T>::zero();
  const T* gk = g + G.off[k];
  if (ca < G.shape[a] - 1) r = gk[idx];
  if (ca > 0) r = Elem<T>::sub(r, gk[idx - G.rstride[k][a]]);
  return r;
}

// element e of the stacked gradient: block k, reduced-shape index -> image pixel i, returns x[i]-x[i+e_a]
template <typename T>
__device__ __forceinline__ T grad_fwd_elem(const T* __restrict__ x, const GradGeom& G, int64_t e) {
  int k = 0;
#pragma unroll
  for (int q = 1; q < RLS_MAX_TV_DIMS; ++q)
    if (q < G.ndirs && e >= G.off[q]) k = q;
  const int a = G.dim[k];
  int64_t rem = e - G.off[k], pix = 0;
#pragma unroll
  for (int d = RLS_MAX_TV_DIMS - 1; d >= 0; --d) {
    if (d < G.ndims) {
      int64_t c = rem / G.rstride[k][d];
      rem -= c * G.rstride[k][d];
      pix += c * G.stride[d];
    }
  }
  return Elem<T>::sub(x[pix], x[pix + G.stride[a]]);
}

#endif

int32_t rls_make_grad_geom(int32_t ndims, const int64_t* shape, int32_t ndirs, const int32_t* dims_1based, GradGeom* g);

struct TvWork {
  void* buf[3] = {nullptr, nullptr, nullptr};  // pq / rs / pqOld rotation
  void* xtmp = nullptr;
  int64_t rows = 0, npix = 0;
  int32_t dtype = 0;
};
int32_t rls_tv_work_ensure(rls_ctx_s* ctx, TvWork* w, int32_t dtype, const GradGeom& g);
void rls_tv_work_free(TvWork* w);

// non-ABI launchers used by the solvers (all on ctx->stream).  The threshold is `lam`, or
// *lam_dev when lam_dev != NULL (device-resident thresholds of the whole-solve path).
int32_t rls_prox_launch(rls_ctx_s* ctx, int32_t dtype, void* x, int64_t n, const rls_reg_desc* reg, float lam,
                        const float* lam_dev, const int* gate, TvWork* tv);
// singular-value thresholding (rls_svt.cu): NuclearRegularization on the column-major rows x cols matrix, LLRRegularization
// on the blockSize patches of an image series (shift = circular shift of the patch grid, may be NULL)
int32_t rls_prox_nuclear_launch(rls_ctx_s* ctx, int32_t dtype, void* x, int64_t n, int64_t rows, int64_t cols, float lam, const float* lam_dev,
                                const int* gate);
int32_t rls_prox_llr_launch(rls_ctx_s* ctx, int32_t dtype, void* x, int64_t n, int32_t ndims, const int64_t* shape, const int64_t* block,
                            const int64_t* shift, int fully_overlapping, float lam, const float* lam_dev, const int* gate);
int32_t rls_proj_launch(rls_ctx_s* ctx, int32_t dtype, void* x, int64_t n, int proj_mask, const int* gate);
// out[e] = (Phi x)[e]
int32_t rls_grad_fwd_launch(rls_ctx_s* ctx, int32_t dtype, const void* x, void* out, const GradGeom& g, const int* gate);
// res = a*(A_k' g_k) + res accumulated block by block, starting from base (NULL = 0); a from device if a_dev
int32_t rls_grad_t_axpy_launch(rls_ctx_s* ctx, int32_t dtype, const void* g, const void* base, void* res, float a,
                               const float* a_dev, float a_sign, const GradGeom& geom, const int* gate);
