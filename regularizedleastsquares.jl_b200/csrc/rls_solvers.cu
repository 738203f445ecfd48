// rls_solvers.cu — iteration bodies of FISTA / POGM / OptISTA / CGNR / ADMM as fused
// elementwise epilogue kernels around the normal-operator apply.
//   reference: src/FISTA.jl:139-185, src/POGM.jl:173-237, src/OptISTA.jl:164-204,
//              src/CGNR.jl:143-178, src/ADMM.jl:230-322 (+ IterativeSolvers cg!).
// Every solver iteration is: [scalar pre-step] -> normal apply -> ONE fused elementwise
// kernel (residual, gradient step, norm/dot reductions, elementwise prox, projections,
// momentum/inertia) whose last block runs the scalar recurrences on the device.  All
// kernels are gated by the device-side done() flag, so a callback-free solve! is enqueued
// back to back with a single synchronisation at the end.
#include <chrono>
#include <map>

#include "rls_prox.cuh"
#include "rls_solver_state.cuh"

namespace {
constexpr int EB = 256;
#define EW_LOOP(i, n) for (int64_t i = (int64_t)blockIdx.x * EB + threadIdx.x; i < (n); i += (int64_t)gridDim.x * EB)

static inline int ew_grid(const rls_ctx_s* c, int64_t n) {
  int64_t g = (n + EB - 1) / EB;
  int64_t cap = (int64_t)c->sm_count * 8;
  if (g > cap) g = cap;
  if (g > RLS_MAX_RED_BLOCKS) g = RLS_MAX_RED_BLOCKS;
  if (g < 1) g = 1;
  return (int)g;
}

__global__ void scalar_kernel(DevState* S, int step, int arg, const int* gate) {
  pdl_prologue();
  if (gate && *gate) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    scalar_step(S, step, arg, t);
  }
}

template <typename T>
__global__ void __launch_bounds__(EB) copy_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n, const int* gate) {
  if (gate && *gate) return;
  EW_LOOP(i, n) dst[i] = src[i];
}

template <typename T>
__global__ void __launch_bounds__(EB) fill_gated_kernel(T* __restrict__ dst, T v, int64_t n, const int* gate) {
  if (gate && *gate) return;
  EW_LOOP(i, n) dst[i] = v;
}

// |x|^2 with a scalar step in the finaliser (init: norm_x0 / z0)
template <typename T>
__global__ void __launch_bounds__(EB) norm_step_kernel(const T* __restrict__ x, int64_t n, DevState* S, int step, int arg,
                                                        double* partials, unsigned* ticket, const int* gate) {
  if (gate && *gate) return;
  double acc[1] = {0.0};
  EW_LOOP(i, n) acc[0] += Elem<T>::abs2(x[i]);
  grid_reduce_finalize<1, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, step, arg, t); });
}

// ================================ FISTA ==============================================
template <typename T>
__global__ void __launch_bounds__(EB) fista_momentum_kernel(T* __restrict__ x, const T* __restrict__ xold, int64_t n,
                                                             const DevState* __restrict__ S) {
  pdl_prologue();
  if (S->done) return;
  // x holds xᵒˡᵈ after the pointer swap: x = x*(1-θᵒˡᵈ)/θ ; x += ((θᵒˡᵈ-1)/θ + 1)*xᵒˡᵈ   FISTA.jl:144-148
  const float c1 = fdiv(fsub(1.f, S->theta_old), S->theta);
  const float c2 = fadd(fdiv(fsub(S->theta_old, 1.f), S->theta), 1.f);
  EW_LOOP(i, n) x[i] = Elem<T>::add(Elem<T>::scale(x[i], c1), Elem<T>::scale(xold[i], c2));
}

// sum of the per-cluster partials of a deferred one-pass apply, in rowpass_finish_kernel's order
template <typename T> __device__ __forceinline__ T sum_partials(const NormalPartials& np, int64_t i);
template <> __device__ __forceinline__ float sum_partials<float>(const NormalPartials& np, int64_t i) {
  double s = 0.0;   // Float64 accumulation, index order: bit-identical to rowpass_finish_kernel
  for (int k = 0; k < np.ncl; ++k) s += (double)__ldcg(np.gpart + (size_t)k * np.gstride + i);
  return (float)s;
}
template <> __device__ __forceinline__ float2 sum_partials<float2>(const NormalPartials& np, int64_t i) {
  double sx = 0.0, sy = 0.0;
  for (int k = 0; k < np.ncl; ++k) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(np.gpart + (size_t)k * np.gstride) + i);
    sx += (double)t.x; sy += (double)t.y;
  }
  return make_float2((float)sx, (float)sy);
}

// PART 0: everything; PART 1: up to the gradient step (a non-elementwise prox follows);
// PART 2: projections + restart dot after that prox.
// Fused form (single GPU, row-major one-pass operator): np.ncl > 0 — AHA x arrives as per-cluster partials that this
// kernel adds up itself; fuse_mom — the momentum step FISTA.jl:144-148 is applied here (and, identically, inside the
// operator kernel's x load) instead of by fista_momentum_kernel.
template <typename T, int PART>
__global__ void __launch_bounds__(EB) fista_main_kernel(T* __restrict__ x, T* __restrict__ res, const T* __restrict__ x0,
                                                         const T* __restrict__ xold, int64_t n, DevState* S, int reg_kind,
                                                         double* partials, unsigned* ticket, NormalPartials np, int fuse_mom) {
  pdl_prologue();
  if (S->done) return;
  const float rho = S->rho, thr = thr_from(S, 0, S->rho);   // ρ*λ(reg)   FISTA.jl:164
  const int proj = S->proj_mask, restart = S->restart;
  const float c1 = fdiv(fsub(1.f, S->theta_old), S->theta);
  const float c2 = fadd(fdiv(fsub(S->theta_old, 1.f), S->theta), 1.f);
  double acc[2] = {0.0, 0.0};
  EW_LOOP(i, n) {
    T r, xv;
    if (PART != 2) {
      T xin = x[i];
      if (fuse_mom) xin = Elem<T>::add(Elem<T>::scale(xin, c1), Elem<T>::scale(xold[i], c2));   // :144-148
      const T ahax = np.ncl > 0 ? sum_partials<T>(np, i) : res[i];
      r = Elem<T>::sub(ahax, x0[i]);                   // res = AHA x - x₀        :152-153
      res[i] = r;
      xv = Elem<T>::sub(xin, Elem<T>::scale(r, rho));  // x -= ρ res              :154
      acc[0] += Elem<T>::abs2(r);
      if (PART == 1) { x[i] = xv; continue; }
      xv = prox_elementwise(xv, reg_kind, thr);        // prox!(reg, x, ρλ)       :164
    } else {
      r = res[i];
      xv = x[i];
    }
    if (proj) xv = proj_elem(xv, proj);                // :166-168
    x[i] = xv;
    if (restart) {                                     // real(res ⋅ (x - xᵒˡᵈ))  :172
      double im = 0.0;
      Elem<T>::dotc(r, Elem<T>::sub(xv, xold[i]), acc[1], im);
    }
  }
  const int step = PART == 0 ? STEP_FISTA_POST : (PART == 1 ? STEP_FISTA_GRAD : STEP_FISTA_TAIL);
  grid_reduce_finalize<2, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, step, 0, t); });
}

// ---- multi-RHS: the K per-column pre / post kernels of one batched FISTA iteration as ONE launch each --------
// blockIdx.y = column (lane); pointers travel as kernel parameters; every lane keeps its own done() gate, its own
// reduction slots (partials + lane*stride) and its own ticket, so each column's arithmetic — and therefore its
// iterates — are exactly those of the per-column kernels above.
constexpr int BATCH_MAXK = 128;
template <typename T>
struct FistaLanes {
  T* x[BATCH_MAXK];
  T* res[BATCH_MAXK];
  const T* x0[BATCH_MAXK];
  const T* xold[BATCH_MAXK];
  DevState* S[BATCH_MAXK];
};

template <typename T>
__global__ void __launch_bounds__(EB) fista_momentum_batch_kernel(const __grid_constant__ FistaLanes<T> L, int64_t n) {
  pdl_prologue();
  const int k = blockIdx.y;
  const DevState* S = L.S[k];
  if (S->done) return;
  T* __restrict__ x = L.x[k];
  const T* __restrict__ xold = L.xold[k];
  const float c1 = fdiv(fsub(1.f, S->theta_old), S->theta);
  const float c2 = fadd(fdiv(fsub(S->theta_old, 1.f), S->theta), 1.f);
  EW_LOOP(i, n) x[i] = Elem<T>::add(Elem<T>::scale(x[i], c1), Elem<T>::scale(xold[i], c2));
}

template <typename T>
__global__ void __launch_bounds__(EB) fista_main_batch_kernel(const __grid_constant__ FistaLanes<T> L, int64_t n, int reg_kind, double* partials,
                                                               int64_t pstride, unsigned* tickets) {
  pdl_prologue();
  const int k = blockIdx.y;
  DevState* S = L.S[k];
  if (S->done) return;
  T* __restrict__ x = L.x[k];
  T* __restrict__ res = L.res[k];
  const T* __restrict__ x0 = L.x0[k];
  const T* __restrict__ xold = L.xold[k];
  const float rho = S->rho, thr = thr_from(S, 0, S->rho);
  const int proj = S->proj_mask, restart = S->restart;
  double acc[2] = {0.0, 0.0};
  EW_LOOP(i, n) {
    const T r = Elem<T>::sub(res[i], x0[i]);
    res[i] = r;
    T xv = Elem<T>::sub(x[i], Elem<T>::scale(r, rho));
    acc[0] += Elem<T>::abs2(r);
    xv = prox_elementwise(xv, reg_kind, thr);
    if (proj) xv = proj_elem(xv, proj);
    x[i] = xv;
    if (restart) {
      double im = 0.0;
      Elem<T>::dotc(r, Elem<T>::sub(xv, xold[i]), acc[1], im);
    }
  }
  grid_reduce_finalize<2, EB>(acc, partials + (size_t)k * pstride, tickets + k, [=](double* t) { scalar_step(S, STEP_FISTA_POST, 0, t); });
}

// ================================ SplitBregman =======================================
// Bregman update (SplitBregman.jl:258-263), executed only when the device-side flag says so:
//   β_y .+= y ; mul!(β_y, AHA, x, -1, 1)        (ahax = AHA x from the gated operator apply)
template <typename T>
__global__ void __launch_bounds__(EB) sb_outer_kernel(T* __restrict__ beta_y, const T* __restrict__ y, const T* __restrict__ ahax,
                                                       int64_t n, const DevState* __restrict__ S) {
  if (S->done || S->sb_outer_gate) return;
  EW_LOOP(i, n) beta_y[i] = Elem<T>::sub(Elem<T>::add(beta_y[i], y[i]), ahax[i]);
}
//   z[i] = Φ_i x (identity regTrafo) ; u[i] .= 0
template <typename T>
__global__ void __launch_bounds__(EB) sb_reset_term_kernel(T* __restrict__ z, T* __restrict__ u, const T* __restrict__ x, int64_t n,
                                                            const DevState* __restrict__ S) {
  if (S->done || S->sb_outer_gate) return;
  EW_LOOP(i, n) { z[i] = x[i]; u[i] = Elem<T>::zero(); }
}

// ================================ POGM ===============================================
// bufX holds x on entry and y on exit; bufY holds y on entry and x on exit (the host swaps roles).
template <typename T, int PART>
__global__ void __launch_bounds__(EB) pogm_main_kernel(T* __restrict__ bufX, T* __restrict__ bufY, const T* __restrict__ x0,
                                                        T* __restrict__ xold, T* __restrict__ z, T* __restrict__ w,
                                                        T* __restrict__ res, int64_t n, DevState* S, int reg_kind,
                                                        double* partials, unsigned* ticket) {
  pdl_prologue();
  if (S->done) return;
  const float rho = S->rho, alpha = S->alpha, beta = S->beta, gamma = S->gamma, gamma_old = S->gamma_old, thr = S->thr;
  const int proj = S->proj_mask, restart = S->restart;
  const float c_y = fadd(fadd(1.f, alpha), beta);                        // (1 + α + β)
  const float rag = fdiv(fmul(rho, alpha), gamma_old);                   // ρα/γᵒˡᵈ
  const float c_xo = fadd(beta, rag);                                    // (β + ρα/γᵒˡᵈ)
  const float rg = fdiv(rho, gamma);                                     // ρ/γ
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  EW_LOOP(i, n) {
    T r, xs, xn, xp;
    if (PART != 2) {
      const T xo = bufX[i];
      xold[i] = xo;                                                      // :180
      r = Elem<T>::sub(res[i], x0[i]);                                   // :181-182
      res[i] = r;
      xs = Elem<T>::sub(xo, Elem<T>::scale(r, rho));                     // :183  (becomes y)
      acc[0] += Elem<T>::abs2(r);
      xn = Elem<T>::scale(bufY[i], -alpha);                              // :209-212
      xn = Elem<T>::add(xn, Elem<T>::scale(xs, c_y));
      xn = Elem<T>::sub(xn, Elem<T>::scale(xo, c_xo));
      xn = Elem<T>::add(xn, Elem<T>::scale(z[i], rag));
      z[i] = xn;                                                         // :213
      bufX[i] = xs;
      if (PART == 1) { bufY[i] = xn; continue; }
      xp = prox_elementwise(xn, reg_kind, thr);                          // :216
    } else {
      r = res[i]; xs = bufX[i]; xn = z[i]; xp = bufY[i];
    }
    if (proj) xp = proj_elem(xp, proj);                                  // :217-219
    bufY[i] = xp;
    if (restart) {                                                       // :223-231
      T wv = Elem<T>::add(w[i], Elem<T>::add(xs, Elem<T>::scale(Elem<T>::sub(xp, xn), rg)));
      double im = 0.0;
      Elem<T>::dotc(wv, xp, acc[1], im);
      Elem<T>::dotc(wv, xn, acc[2], im);
      Elem<T>::dotc(wv, r, acc[3], im);
      w[i] = Elem<T>::sub(Elem<T>::scale(Elem<T>::sub(xn, xp), rg), xs);
    }
  }
  const int step = PART == 1 ? STEP_POGM_GRAD : STEP_POGM_POST;
  const int arg = PART == 2 ? 1 : 0;
  grid_reduce_finalize<4, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, step, arg, t); });
}

// ================================ OptISTA ============================================
template <typename T, int PART>
__global__ void __launch_bounds__(EB) optista_main_kernel(T* __restrict__ x, T* __restrict__ y, T* __restrict__ z,
                                                           T* __restrict__ zold, const T* __restrict__ x0,
                                                           T* __restrict__ res, int64_t n, DevState* S, int reg_kind,
                                                           double* partials, unsigned* ticket) {
  pdl_prologue();
  if (S->done) return;
  const float alpha = S->alpha, beta = S->beta, gamma = S->gamma, thr = S->thr;
  const float rg = fmul(S->rho, gamma);                                  // ρ*γ
  const float c_z = fadd(fadd(1.f, alpha), beta);
  double acc[1] = {0.0};
  EW_LOOP(i, n) {
    T yp, yold;
    if (PART != 2) {
      zold[i] = z[i];                                                    // :180
      yold = y[i];
      z[i] = yold;                                                       // :181 (z holds yᵒˡᵈ)
      T r = Elem<T>::sub(res[i], x0[i]);                                 // :182-183
      res[i] = r;
      T yn = Elem<T>::sub(yold, Elem<T>::scale(r, rg));                  // :184
      acc[0] += Elem<T>::abs2(r);
      if (PART == 1) { y[i] = yn; continue; }
      yp = prox_elementwise(yn, reg_kind, thr);                          // :190 (no projections in OptISTA's loop)
    } else {
      yp = y[i];
      yold = z[i];
    }
    y[i] = yp;
    T zn = Elem<T>::divr(yold, -gamma);                                  // :195
    const T xv = x[i];
    zn = Elem<T>::add(zn, Elem<T>::add(xv, Elem<T>::divr(yp, gamma)));   // :196
    z[i] = zn;
    T xn = Elem<T>::scale(xv, -beta);                                    // :197-199
    xn = Elem<T>::add(xn, Elem<T>::scale(zn, c_z));
    xn = Elem<T>::sub(xn, Elem<T>::scale(zold[i], alpha));
    x[i] = xn;
  }
  const int step = PART == 1 ? STEP_OPTISTA_GRAD : STEP_OPTISTA_POST;
  const int arg = PART == 2 ? 1 : 0;
  grid_reduce_finalize<1, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, step, arg, t); });
}

// ================================ CGNR ===============================================
template <typename T>
__device__ __forceinline__ T scal(float2 s);
template <> __device__ __forceinline__ float scal<float>(float2 s) { return s.x; }
template <> __device__ __forceinline__ float2 scal<float2>(float2 s) { return s; }
__device__ __forceinline__ float negs(float a) { return -a; }
__device__ __forceinline__ float2 negs(float2 a) { return make_float2(-a.x, -a.y); }

template <typename T>
__global__ void __launch_bounds__(EB) cgnr_dot_kernel(const T* __restrict__ p, const T* __restrict__ v, int64_t n, DevState* S,
                                                       double* partials, unsigned* ticket) {
  pdl_prologue();
  if (S->done) return;
  double acc[2] = {0.0, 0.0};
  EW_LOOP(i, n) Elem<T>::dotc(p[i], v[i], acc[0], acc[1]);               // dot(pl, vl)   CGNR.jl:154
  grid_reduce_finalize<2, EB>(acc, partials, ticket,
                              [=](double* t) { scalar_step(S, STEP_CGNR_ALPHA, Elem<T>::is_complex ? 1 : 0, t); });
}

template <typename T>
__global__ void __launch_bounds__(EB) cgnr_update_kernel(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p,
                                                          const T* __restrict__ v, int64_t n, DevState* S,
                                                          double* partials, unsigned* ticket) {
  pdl_prologue();
  if (S->done) return;
  const T al = scal<T>(S->cg_alpha);
  const T nal = negs(al);
  const bool lam_pos = S->lam_is_f64[0] ? (S->lam64[0] > 0.0) : (S->lam[0] > 0.f);
  const float lam = S->lam_is_f64[0] ? (float)S->lam64[0] : S->lam[0];
  double acc[1] = {0.0};
  EW_LOOP(i, n) {
    const T pv = p[i];
    x[i] = Elem<T>::add(x[i], cmul(pv, al));                             // x += p*α          :163
    T rv = Elem<T>::add(r[i], cmul(v[i], nal));                          // x₀ += v*(-α)      :165
    if (lam_pos) rv = Elem<T>::add(rv, cmul(Elem<T>::scale(pv, -lam), al));  // x₀ += (p*-λ)*α :168
    r[i] = rv;
    acc[0] += Elem<T>::abs2(rv);
  }
  grid_reduce_finalize<1, EB>(acc, partials, ticket,
                              [=](double* t) { scalar_step(S, STEP_CGNR_BETA, Elem<T>::is_complex ? 1 : 0, t); });
}

template <typename T>
__global__ void __launch_bounds__(EB) cgnr_p_kernel(T* __restrict__ p, const T* __restrict__ r, int64_t n, DevState* S,
                                                     double* partials, unsigned* ticket) {
  pdl_prologue();
  if (S->done) return;
  const T be = scal<T>(S->cg_beta);
  double acc[1] = {0.0};
  EW_LOOP(i, n) {
    T pv = Elem<T>::add(cmul(p[i], be), r[i]);                           // rmul!(pl, β); pl += x₀   :173-174
    p[i] = pv;
    acc[0] += Elem<T>::abs2(pv);
  }
  grid_reduce_finalize<1, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, STEP_CGNR_POST, 0, t); });
}


// ================================ CGNR: one cooperative kernel per solve ==============================================
// L2-resident systems (BASELINE config C1: 1024 x 4096 ComplexF32 = 33.6 MB) spend a CGNR iteration in five dependent kernels:
// two sweeps over the matrix and three grid reductions whose results steer the next kernel (p.v -> alpha; r.r -> beta;
// |p|^2), 30 us per iteration even when the whole launch sequence is replayed from a CUDA graph (profiles/r02_c1_latency.txt).
// Here the whole solve! (CGNR.jl:143-185, every iteration until done()) is ONE cooperative launch: CTA c owns the columns
// [c*CN, (c+1)*CN) of the column-major A for the whole solve, and with them the entries p_j, x_j, r_j, v_j of the
// n-vectors (shared memory); per iteration
//   A  y_c = sum_{j own} A[:,j] p_j           partial m-vector of this CTA -> global          | grid barrier
//   B  y[i] = sum_c y_c[i]                     rows dealt out to the CTAs, fixed order         | grid barrier
//   C  v_j = A[:,j]' y  (own columns);  partials of p.v and |p|^2                             | grid barrier
//   D  alpha (every CTA, same fixed-order sum => bit-identical);  x_j += alpha p_j,  r_j -= alpha v_j (- lambda alpha p_j);
//      partial |r|^2                                                                         | grid barrier
//   E  beta, iteration += 1, done();  p_j = beta p_j + r_j
// i.e. four grid-wide exchanges per iteration and two sweeps over the L2-resident matrix.  The scalar recurrences are
// the SAME device code as the chained path (scalar_step: Float32, individually rounded, complex Smith division), evaluated
// redundantly by every CTA on a private copy of the DevState; CTA 0 writes the state back at the end.  The element updates are
// those of cgnr_update_kernel / cgnr_p_kernel.  What differs from the chained path is only the order of the sums inside
// A p, A' y and the three dot products (per CTA, then over the CTAs), i.e. rounding of the size the parity bound allows.
// The grid barrier is a monotone counter in global memory (cooperative launch guarantees co-residency); every wait is
// bounded (abort flag -> RLS_ERR_CUDA).
constexpr int PK_THREADS = 1024;
constexpr int PK_MAXROWS = 4;      // rows per thread: m <= 4096
constexpr int PK_MAXCN = 64;       // columns per CTA

struct PkArgs {
  const void* A; int64_t ld; int m, n, CN;
  int ncache;                     // own columns kept in shared memory for the whole solve (the others stream from L2)
  void *x, *r, *p, *v;            // n-vectors of the lane (V_X, V_X0, V_P, V_V)
  DevState* S;
  void* ypart;                    // [grid][m]
  void* y;                        // [m]
  double* dpart;                  // [2][grid][4]
  unsigned* bar;                  // monotone barrier counter (zeroed before the launch)
  int* abort_flag;
  int cap;
};

__device__ __forceinline__ void pk_barrier(unsigned* bar, unsigned& target, int* abort_flag, volatile int* s_abort) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if (clock64() - t0 > 2000000000ll) { *s_abort = 1; atomicExch(abort_flag, 1); break; }
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ T pk_ldcg(const T* p);
template <> __device__ __forceinline__ float pk_ldcg<float>(const float* p) { return __ldcg(p); }
template <> __device__ __forceinline__ float2 pk_ldcg<float2>(const float2* p) { return __ldcg(p); }
// FMA-contracted multiply-accumulate for the matrix sweeps (as a BLAS gemv would)
__device__ __forceinline__ void pk_fma(float& acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ __forceinline__ void pk_fma(float2& acc, float2 a, float2 b) {          // acc += a*b
  acc.x = fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x));
  acc.y = fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y));
}
__device__ __forceinline__ void pk_fmac(float& acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ __forceinline__ void pk_fmac(float2& acc, float2 a, float2 b) {         // acc += conj(a)*b
  acc.x = fmaf(a.x, b.x, fmaf(a.y, b.y, acc.x));
  acc.y = fmaf(a.x, b.y, fmaf(-a.y, b.x, acc.y));
}
__device__ __forceinline__ float pk_shfl(float v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ float2 pk_shfl(float2 v, int o) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}
template <typename T> __device__ __forceinline__ T pk_plain_add(T a, T b);
template <> __device__ __forceinline__ float pk_plain_add<float>(float a, float b) { return a + b; }
template <> __device__ __forceinline__ float2 pk_plain_add<float2>(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }

// fixed-order sum over the CTAs of NV doubles per CTA (dpart[c*4 + k]); every thread of warp 0 returns the totals
template <int NV>
__device__ __forceinline__ void pk_total(const double* dpart, double (&t)[NV]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = 0.0;
    for (unsigned c = lane; c < gridDim.x; c += 32) a += __ldcg(&dpart[(size_t)c * 4 + k]);
    t[k] = warp_sum(a);
  }
}

template <typename T>
__global__ void __launch_bounds__(PK_THREADS, 1) cgnr_persistent_kernel(PkArgs a) {
  extern __shared__ __align__(16) unsigned char pk_smem[];
  __shared__ DevState Sl;
  __shared__ int s_abort;
  T* ys = reinterpret_cast<T*>(pk_smem);          // [m]
  T* ps = ys + a.m;                                // [CN] each: p, x, r, v of the own columns
  T* xs = ps + a.CN;
  T* rs = xs + a.CN;
  T* vs = rs + a.CN;
  T* As = vs + a.CN;                               // [ncache][m]: the first ncache own columns of A, resident for the whole solve
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x, P = gridDim.x;
  const int j0 = c * a.CN;
  const int nown = max(0, min(a.CN, a.n - j0));
  const T* A = reinterpret_cast<const T*>(a.A);
  T* ypart = reinterpret_cast<T*>(a.ypart);
  T* yg = reinterpret_cast<T*>(a.y);
  unsigned target = 0;
  constexpr bool cplx = Elem<T>::is_complex;

  if (tid == 0) { Sl = *a.S; s_abort = 0; }
  for (int j = tid; j < nown; j += PK_THREADS) {
    ps[j] = reinterpret_cast<const T*>(a.p)[j0 + j];
    xs[j] = reinterpret_cast<const T*>(a.x)[j0 + j];
    rs[j] = reinterpret_cast<const T*>(a.r)[j0 + j];
    vs[j] = reinterpret_cast<const T*>(a.v)[j0 + j];
  }
  const int ncache = min(a.ncache, nown);
  for (int idx = tid; idx < ncache * a.m; idx += PK_THREADS) {
    const int j = idx / a.m, i = idx - j * a.m;
    As[idx] = __ldg(A + (int64_t)(j0 + j) * a.ld + i);
  }
  __syncthreads();
  const int RM = (a.m + P - 1) / P;                // rows of y this CTA reduces in phase B

  for (int it = 0; it < a.cap; ++it) {
    if (Sl.done || s_abort) break;                 // identical in every CTA: same state, same arithmetic
    // ---- A: partial y over the own columns -----------------------------------------------------------------------
    {
      T acc[PK_MAXROWS];
#pragma unroll
      for (int q = 0; q < PK_MAXROWS; ++q) acc[q] = Elem<T>::zero();
#pragma unroll 4
      for (int j = 0; j < ncache; ++j) {           // columns resident in shared memory
        const T pj = ps[j];
        const T* col = As + j * a.m;
#pragma unroll
        for (int q = 0; q < PK_MAXROWS; ++q) {
          const int i = tid + q * PK_THREADS;
          if (i < a.m) pk_fma(acc[q], col[i], pj);
        }
      }
#pragma unroll 4
      for (int j = ncache; j < nown; ++j) {        // the rest streams from L2
        const T pj = ps[j];
        const T* col = A + (int64_t)(j0 + j) * a.ld;
#pragma unroll
        for (int q = 0; q < PK_MAXROWS; ++q) {
          const int i = tid + q * PK_THREADS;
          if (i < a.m) pk_fma(acc[q], __ldg(col + i), pj);
        }
      }
#pragma unroll
      for (int q = 0; q < PK_MAXROWS; ++q) {
        const int i = tid + q * PK_THREADS;
        if (i < a.m) ypart[(size_t)c * a.m + i] = acc[q];
      }
    }
    pk_barrier(a.bar, target, a.abort_flag, &s_abort);
    // ---- B: y rows [c*RM, (c+1)*RM): one warp per row, partials added in CTA order ----------------------------------
    for (int rr_ = warp; rr_ < RM; rr_ += PK_THREADS / 32) {
      const int i = c * RM + rr_;
      if (i < a.m) {
        T sacc = Elem<T>::zero();
        for (int cc = lane; cc < P; cc += 32) sacc = pk_plain_add(sacc, pk_ldcg(ypart + (size_t)cc * a.m + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc = pk_plain_add(sacc, pk_shfl(sacc, o));
        if (lane == 0) yg[i] = sacc;
      }
    }
    pk_barrier(a.bar, target, a.abort_flag, &s_abort);
    // ---- C: v_j = A[:,j]' y for the own columns; partials of p.v and |p|^2 -----------------------------------------
    for (int i = tid; i < a.m; i += PK_THREADS) ys[i] = pk_ldcg(yg + i);
    __syncthreads();
    for (int j = warp; j < nown; j += PK_THREADS / 32) {
      T sacc = Elem<T>::zero();
      if (j < ncache) {
        const T* col = As + j * a.m;
#pragma unroll 8
        for (int i = lane; i < a.m; i += 32) pk_fmac(sacc, col[i], ys[i]);
      } else {
        const T* col = A + (int64_t)(j0 + j) * a.ld;
#pragma unroll 8
        for (int i = lane; i < a.m; i += 32) pk_fmac(sacc, __ldg(col + i), ys[i]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc = pk_plain_add(sacc, pk_shfl(sacc, o));
      if (lane == 0) vs[j] = sacc;
    }
    __syncthreads();
    double* dp = a.dpart + (size_t)(it & 1) * P * 4;
    if (warp == 0) {
      double d0 = 0.0, d1 = 0.0, d2 = 0.0;
      for (int j = lane; j < nown; j += 32) {
        Elem<T>::dotc(ps[j], vs[j], d0, d1);                               // dot(pl, vl)   CGNR.jl:154
        d2 += Elem<T>::abs2(ps[j]);
      }
      d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
      if (lane == 0) { dp[c * 4 + 0] = d0; dp[c * 4 + 1] = d1; dp[c * 4 + 2] = d2; }
    }
    pk_barrier(a.bar, target, a.abort_flag, &s_abort);
    // ---- D: alpha; x, r update on the own columns; partial |r|^2 ---------------------------------------------------
    if (warp == 0) {
      double t[3];
      pk_total<3>(dp, t);
      if (lane == 0) {
        Sl.pp = t[2];
        scalar_step(&Sl, STEP_CGNR_ALPHA, cplx ? 1 : 0, t);
      }
    }
    __syncthreads();
    double* dq = a.dpart + (size_t)(2 + (it & 1)) * P * 4;
    {
      const T al = scal<T>(Sl.cg_alpha);
      const T nal = negs(al);
      const bool lam_pos = Sl.lam_is_f64[0] ? (Sl.lam64[0] > 0.0) : (Sl.lam[0] > 0.f);
      const float lam = Sl.lam_is_f64[0] ? (float)Sl.lam64[0] : Sl.lam[0];
      double racc = 0.0;
      if (warp == 0) {
        for (int j = lane; j < nown; j += 32) {
          const T pv = ps[j];
          xs[j] = Elem<T>::add(xs[j], cmul(pv, al));                       // x += p*α          :163
          T rv = Elem<T>::add(rs[j], cmul(vs[j], nal));                    // x₀ += v*(-α)      :165
          if (lam_pos) rv = Elem<T>::add(rv, cmul(Elem<T>::scale(pv, -lam), al));  // x₀ += (p*-λ)*α :168
          rs[j] = rv;
          racc += Elem<T>::abs2(rv);
        }
        racc = warp_sum(racc);
        if (lane == 0) dq[c * 4 + 0] = racc;
      }
    }
    pk_barrier(a.bar, target, a.abort_flag, &s_abort);
    // ---- E: beta, iteration count, done(); p update ------------------------------------------------------------------
    if (warp == 0) {
      double t[1];
      pk_total<1>(dq, t);
      if (lane == 0) {
        scalar_step(&Sl, STEP_CGNR_BETA, cplx ? 1 : 0, t);
        double keep = Sl.pp;                                               // |p|^2 of the new p is summed with the next p.v
        scalar_step(&Sl, STEP_CGNR_POST, 0, &keep);
      }
    }
    __syncthreads();
    {
      const T be = scal<T>(Sl.cg_beta);
      for (int j = tid; j < nown; j += PK_THREADS) ps[j] = Elem<T>::add(cmul(ps[j], be), rs[j]);   // rmul!(pl, β); pl += x₀ :173-174
    }
    __syncthreads();
  }
  // ---- write back: vectors of the own columns, |p|^2 of the final p, the state --------------------------------------
  for (int j = tid; j < nown; j += PK_THREADS) {
    reinterpret_cast<T*>(a.p)[j0 + j] = ps[j];
    reinterpret_cast<T*>(a.x)[j0 + j] = xs[j];
    reinterpret_cast<T*>(a.r)[j0 + j] = rs[j];
    reinterpret_cast<T*>(a.v)[j0 + j] = vs[j];
  }
  double* df = a.dpart + (size_t)4 * P * 4;
  if (warp == 0) {
    double d2 = 0.0;
    for (int j = lane; j < nown; j += 32) d2 += Elem<T>::abs2(ps[j]);
    d2 = warp_sum(d2);
    if (lane == 0) df[c * 4 + 0] = d2;
  }
  pk_barrier(a.bar, target, a.abort_flag, &s_abort);
  if (c == 0 && warp == 0) {
    double t[1];
    pk_total<1>(df, t);
    if (lane == 0) { Sl.pp = t[0]; *a.S = Sl; }
  }
}

// ================================ ADMM ===============================================
// β = (first ? β_y : β); β = ρ z + β; β = (-ρ) u + β        (identity regTrafo)  ADMM.jl:236-241
template <typename T>
__global__ void __launch_bounds__(EB) admm_beta_identity_kernel(T* __restrict__ beta, const T* __restrict__ beta_y,
                                                                 const T* __restrict__ z, const T* __restrict__ u, int64_t n,
                                                                 const DevState* __restrict__ S, int term, int first) {
  if (S->done) return;
  const float rho = S->a_rho[term];
  EW_LOOP(i, n) {
    T b = first ? beta_y[i] : beta[i];
    b = Elem<T>::add(Elem<T>::scale(z[i], rho), b);
    b = Elem<T>::add(Elem<T>::scale(u[i], -rho), b);
    beta[i] = b;
  }
}

// r = β - c ; |r|²  -> cg_iterator! initial residual
template <typename T>
__global__ void __launch_bounds__(EB) admm_cg_init_kernel(T* __restrict__ r, const T* __restrict__ beta, const T* __restrict__ c,
                                                           int64_t n, DevState* S, double* partials, unsigned* ticket) {
  if (S->done) return;
  double acc[1] = {0.0};
  EW_LOOP(i, n) {
    T rv = Elem<T>::sub(beta[i], c[i]);
    r[i] = rv;
    acc[0] += Elem<T>::abs2(rv);
  }
  grid_reduce_finalize<1, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, STEP_ADMM_CG_INIT, 0, t); });
}

// u = r + β u
template <typename T>
__global__ void __launch_bounds__(EB) admm_cg_u_kernel(T* __restrict__ u, const T* __restrict__ r, int64_t n,
                                                        const DevState* __restrict__ S) {
  if (S->cgi_gate) return;
  const float be = S->cgi_beta;
  EW_LOOP(i, n) u[i] = Elem<T>::add(r[i], Elem<T>::scale(u[i], be));
}

// c = ρ_i v + c  for the identity terms (all_identity: every term), then u·c
template <typename T>
__global__ void __launch_bounds__(EB) admm_cg_dot_kernel(const T* __restrict__ u, T* __restrict__ c, int64_t n, DevState* S,
                                                          int n_identity, double* partials, unsigned* ticket) {
  if (S->cgi_gate) return;
  double acc[2] = {0.0, 0.0};
  EW_LOOP(i, n) {
    const T uv = u[i];
    T cv = c[i];
    for (int k = 0; k < n_identity; ++k) cv = Elem<T>::add(Elem<T>::scale(uv, S->a_rho[k]), cv);
    if (n_identity) c[i] = cv;
    Elem<T>::dotc(uv, cv, acc[0], acc[1]);
  }
  grid_reduce_finalize<2, EB>(acc, partials, ticket,
                              [=](double* t) { scalar_step(S, STEP_ADMM_CG_ALPHA, Elem<T>::is_complex ? 1 : 0, t); });
}

// c = ρ v + c (single identity term, used before the first residual and for mixed trafos)
template <typename T>
__global__ void __launch_bounds__(EB) admm_axpy_rho_kernel(T* __restrict__ c, const T* __restrict__ v, int64_t n,
                                                            const DevState* __restrict__ S, int term, const int* gate) {
  if (gate && *gate) return;
  const float rho = S->a_rho[term];
  EW_LOOP(i, n) c[i] = Elem<T>::add(Elem<T>::scale(v[i], rho), c[i]);
}

// x += α u ; r -= α c ; |r|²
template <typename T>
__global__ void __launch_bounds__(EB) admm_cg_xr_kernel(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ u,
                                                         const T* __restrict__ c, int64_t n, DevState* S, double* partials,
                                                         unsigned* ticket) {
  if (S->cgi_gate) return;
  const T al = scal<T>(S->cgi_alpha);
  double acc[1] = {0.0};
  EW_LOOP(i, n) {
    x[i] = Elem<T>::add(x[i], cmul(al, u[i]));
    T rv = Elem<T>::sub(r[i], cmul(al, c[i]));
    r[i] = rv;
    acc[0] += Elem<T>::abs2(rv);
  }
  grid_reduce_finalize<1, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, STEP_ADMM_CG_POST, 0, t); });
}

// z / u update and residual sums for an identity regTrafo term (ADMM.jl:251-299).
// PART 0 fused; PART 1: z = x + u only (non-elementwise prox follows); PART 2: the rest.
template <typename T, int PART>
__global__ void __launch_bounds__(EB) admm_term_identity_kernel(const T* __restrict__ x, T* __restrict__ xold, T* __restrict__ z,
                                                                 T* __restrict__ zold, T* __restrict__ u, T* __restrict__ uold,
                                                                 int64_t n, DevState* S, int term, int reg_kind,
                                                                 double* partials, unsigned* ticket) {
  if (S->done) return;
  const float thr = S->a_thr[term];
  const bool do_prox = S->a_rho[term] != 0.f;
  const bool sb = S->sb != 0;                       // SplitBregman: ‖ρΦ'(z-zᵒˡᵈ)‖, ‖ρΦ'u‖ with ρ inside (SplitBregman.jl:244,247)
  const float rho_in = sb ? S->a_rho[term] : 1.f;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  EW_LOOP(i, n) {
    const T xv = x[i], uv = u[i];
    T zn;
    if (PART != 2) {
      zn = Elem<T>::add(xv, uv);                                         // z = Φx; z += u       :258-259
      if (PART == 1) { z[i] = zn; continue; }
      if (do_prox) zn = prox_elementwise(zn, reg_kind, thr);             // :260-262
    } else {
      zn = z[i];
    }
    const T un = Elem<T>::sub(Elem<T>::add(xv, uv), zn);                 // u = Φx + u; u -= z    :265-267
    const T dx = Elem<T>::sub(xv, xold[i]);                              // :282-284
    const T dz = Elem<T>::sub(zn, zold[i]);
    const T du = Elem<T>::sub(un, uv);
    const T xz = Elem<T>::sub(xv, zn);
    acc[0] += Elem<T>::abs2(dx); acc[1] += Elem<T>::abs2(dz); acc[2] += Elem<T>::abs2(du);
    acc[3] += Elem<T>::abs2(sb ? Elem<T>::scale(dz, rho_in) : dz);       // Φ'(z - zᵒˡᵈ)          :289-290
    acc[4] += Elem<T>::abs2(xv); acc[5] += Elem<T>::abs2(zn);            // :292-293
    acc[6] += Elem<T>::abs2(xz);                                         // :295-296
    acc[7] += Elem<T>::abs2(sb ? Elem<T>::scale(un, rho_in) : un);       // :298-299
    z[i] = zn; u[i] = un; uold[i] = du; zold[i] = xz; xold[i] = un;
  }
  if (PART == 1) return;
  grid_reduce_finalize<8, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, STEP_ADMM_TERM, term, t); });
}

// u *= uscale after a ρ adaptation (ADMM.jl:302-309)
template <typename T>
__global__ void __launch_bounds__(EB) admm_scale_u_kernel(T* __restrict__ u, int64_t n, const DevState* __restrict__ S, int term) {
  if (S->done) return;
  const float s = S->a_uscale[term];
  if (s == 1.f) return;
  EW_LOOP(i, n) u[i] = Elem<T>::scale(u[i], s);
}

}  // namespace

// ------------------------------------------------------------------------------------
// ADMM term with regTrafo = GradientOp (ADMM.jl:74): one pass over the dual (gradient)
// domain for z/u and their residual sums, one pass over the pixels for the Φ' residuals.
// ------------------------------------------------------------------------------------
namespace {
template <typename T>
__global__ void __launch_bounds__(EB) admm_term_grad_dual_kernel(const T* __restrict__ x, T* __restrict__ z, T* __restrict__ zold,
                                                                  T* __restrict__ u, T* __restrict__ uold, GradGeom G,
                                                                  DevState* S, int term, int reg_kind, double* partials,
                                                                  unsigned* ticket) {
  if (S->done) return;
  const float thr = S->a_thr[term];
  const bool do_prox = S->a_rho[term] != 0.f;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  EW_LOOP(e, G.rows) {
    const T gx = grad_fwd_elem(x, G, e);                                 // Φx
    const T uv = u[e];
    T zn = Elem<T>::add(gx, uv);
    if (do_prox) zn = prox_elementwise(zn, reg_kind, thr);
    const T un = Elem<T>::sub(Elem<T>::add(gx, uv), zn);
    const T dz = Elem<T>::sub(zn, zold[e]);
    const T du = Elem<T>::sub(un, uv);
    acc[1] += Elem<T>::abs2(dz); acc[2] += Elem<T>::abs2(du);
    acc[4] += Elem<T>::abs2(gx); acc[5] += Elem<T>::abs2(zn);
    acc[6] += Elem<T>::abs2(Elem<T>::sub(gx, zn));
    z[e] = zn; u[e] = un; uold[e] = du; zold[e] = dz;
  }
  grid_reduce_finalize<8, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, STEP_ADMM_TERM_SAVE, term, t); });
}

template <typename T>
__global__ void __launch_bounds__(EB) admm_term_grad_pix_kernel(const T* __restrict__ x, T* __restrict__ xold,
                                                                 const T* __restrict__ dz, const T* __restrict__ u, GradGeom G,
                                                                 DevState* S, int term, double* partials, unsigned* ticket) {
  if (S->done) return;
  const bool sb = S->sb != 0;          // SplitBregman: ‖ρ Φ'(z - zᵒˡᵈ)‖ and ‖ρ Φ'u‖ with ρ inside the norm
  const float rho_in = sb ? S->a_rho[term] : 1.f;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  EW_LOOP(p, G.npix) {
    const T dx = Elem<T>::sub(x[p], xold[p]);
    T t1 = Elem<T>::zero(), t2 = Elem<T>::zero();
    for (int k = 0; k < G.ndirs; ++k) {
      t1 = Elem<T>::add(grad_t_block(dz, G, k, p), t1);                  // Φ'(z - zᵒˡᵈ)
      t2 = Elem<T>::add(grad_t_block(u, G, k, p), t2);                   // Φ'u
    }
    acc[0] += Elem<T>::abs2(dx);
    acc[3] += Elem<T>::abs2(sb ? Elem<T>::scale(t1, rho_in) : t1);
    acc[7] += Elem<T>::abs2(sb ? Elem<T>::scale(t2, rho_in) : t2);
    xold[p] = t2;
  }
  grid_reduce_finalize<8, EB>(acc, partials, ticket, [=](double* t) { scalar_step(S, STEP_ADMM_TERM, term + 16, t); });
}
}  // namespace

static int32_t rls_admm_term_gradient(rls_ctx_s* c, int32_t dtype, const void* x, void* xold, void* z, void* zold, void* u,
                                      void* uold, const GradGeom& G, DevState* S, int term, int reg_kind) {
  const int gd = ew_grid(c, G.rows), gp = ew_grid(c, G.npix);
  if (dtype == RLS_C32) {
    admm_term_grad_dual_kernel<float2><<<gd, EB, 0, c->stream>>>((const float2*)x, (float2*)z, (float2*)zold, (float2*)u, (float2*)uold, G, S, term, reg_kind, c->red_partials, c->red_ticket);
    admm_term_grad_pix_kernel<float2><<<gp, EB, 0, c->stream>>>((const float2*)x, (float2*)xold, (const float2*)zold, (const float2*)u, G, S, term, c->red_partials, c->red_ticket);
  } else {
    admm_term_grad_dual_kernel<float><<<gd, EB, 0, c->stream>>>((const float*)x, (float*)z, (float*)zold, (float*)u, (float*)uold, G, S, term, reg_kind, c->red_partials, c->red_ticket);
    admm_term_grad_pix_kernel<float><<<gp, EB, 0, c->stream>>>((const float*)x, (float*)xold, (const float*)zold, (const float*)u, G, S, term, c->red_partials, c->red_ticket);
  }
  c->launches += 2;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// ------------------------------------------------------------------------------------
// host-side solver object
// ------------------------------------------------------------------------------------
enum VecId { V_X = 0, V_X0, V_XOLD, V_RES, V_Y, V_Z, V_ZOLD, V_W, V_P, V_V, V_BETA, V_BETAY, V_CGU, V_CGR, V_CGC,
             V_TZ0, V_TZOLD0 = V_TZ0 + 4, V_TU0 = V_TZOLD0 + 4, V_TUOLD0 = V_TU0 + 4, V_TTMP0 = V_TUOLD0 + 4, V_COUNT = V_TTMP0 + 4 };

static const char* kVecNames[V_COUNT] = {"x", "x0", "xold", "res", "y", "z", "zold", "w", "p", "v", "beta", "beta_y",
                                         "cg_u", "cg_r", "cg_c",
                                         "z0", "z1", "z2", "z3", "zold0", "zold1", "zold2", "zold3",
                                         "u0", "u1", "u2", "u3", "uold0", "uold1", "uold2", "uold3",
                                         "tmp0", "tmp1", "tmp2", "tmp3"};

struct Lane {
  rls_vec_s* v[V_COUNT] = {};
  DevState* dS = nullptr;
  DevState* hS = nullptr;  // pinned mirror
  int enq_swaps = 0;       // pointer swaps enqueued since the roles were last reconciled
  int base_iter = 0;       // device iteration count at that point
};

struct rls_solver_s {
  rls_ctx_s* ctx = nullptr;
  rls_mat_s* A = nullptr;          // borrowed, may be NULL (AHA only)
  rls_normal_s* AHA = nullptr;
  bool own_AHA = false;
  rls_solver_desc desc{};
  int64_t n = 0, m = 0;            // m = length(b)
  int32_t dtype = 0;
  GradGeom geom[4];
  int64_t rows[4] = {0, 0, 0, 0};  // rows of regTrafo[i]
  TvWork tv;
  std::vector<Lane> lanes;
  // multi-RHS: operands of the K normal-operator applies of one batched iteration (tensor-core GEMM path)
  std::vector<const void*> batch_x;
  std::vector<void*> batch_res;
  std::vector<const int*> batch_gate;
  rls_vec_s* b_dev = nullptr;      // staging for host b
  // L2-resident systems are launch-latency bound (C1: five kernels per CGNR iteration, 2.7 us per boundary): the fixed
  // launch sequence of a callback-free solve! — every kernel is gated on the device-side done() flag — is recorded into a
  // CUDA graph on the second solve and replayed by ONE launch afterwards
  cudaGraphExec_t graph = nullptr;
  int64_t graph_launches = 0;      // kernels inside the graph (added to ctx->launches at every replay)
  const void* graph_scratch = nullptr;  // context scratch pointer baked into the recorded gemv launches
  bool graph_off = false;          // capture failed once: stay on the plain path
  int plain_solves = 0;            // solves enqueued launch by launch so far (the first one allocates scratch: not capturable)
  // one cooperative kernel per CGNR solve (cgnr_persistent_kernel, RLS_CGNR_PERSISTENT=1): exchange buffers
  void* pk_mem = nullptr;
  size_t pk_bytes = 0;
  bool pk_off = false, pk_used = false;
  void* pin_b = nullptr;
  void* pin_x = nullptr;
  size_t pin_b_bytes = 0, pin_x_bytes = 0;
};

namespace {

// upper bound on the iterations a callback-free solve! has to enqueue (done() gates the rest on the device)
static inline int iteration_cap(const rls_solver_desc& d, int64_t n) {
  if (d.kind == RLS_CGNR) return (int)std::min<int64_t>(d.iterations, n);
  if (d.kind == RLS_SPLITBREGMAN) return (int)std::min<int64_t>((int64_t)d.iterations * d.iterations_inner, 1 << 20);
  return d.iterations;
}

// SplitBregman shares ADMM's state layout, x update (β build + warm-started cg!) and term update kernels
static inline bool admm_like(int kind) { return kind == RLS_ADMM || kind == RLS_SPLITBREGMAN; }

static void free_lane(Lane& L) {
  for (int k = 0; k < V_COUNT; ++k)
    if (L.v[k]) { rls_vec_destroy(L.v[k]); L.v[k] = nullptr; }
  if (L.dS) cudaFree(L.dS);
  if (L.hS) cudaFreeHost(L.hS);
  L.dS = nullptr;
  L.hS = nullptr;
}

static int32_t need_vec(rls_solver_s* s, Lane& L, int id, int64_t len) {
  if (L.v[id] && L.v[id]->len == len) return RLS_OK;
  if (L.v[id]) rls_vec_destroy(L.v[id]);
  L.v[id] = nullptr;
  return rls_vec_create_internal(s->ctx, s->dtype, len, &L.v[id]);
}

static int32_t alloc_lane(rls_solver_s* s, Lane& L) {
  const int64_t n = s->n;
  const int kind = s->desc.kind;
  RLS_TRY(need_vec(s, L, V_X, n));
  RLS_TRY(need_vec(s, L, V_X0, n));
  if (kind == RLS_FISTA || kind == RLS_POGM || kind == RLS_OPTISTA) RLS_TRY(need_vec(s, L, V_RES, n));
  if (kind == RLS_FISTA || kind == RLS_POGM || admm_like(kind)) RLS_TRY(need_vec(s, L, V_XOLD, n));
  if (kind == RLS_POGM) { RLS_TRY(need_vec(s, L, V_Y, n)); RLS_TRY(need_vec(s, L, V_Z, n)); RLS_TRY(need_vec(s, L, V_W, n)); }
  if (kind == RLS_OPTISTA) { RLS_TRY(need_vec(s, L, V_Y, n)); RLS_TRY(need_vec(s, L, V_Z, n)); RLS_TRY(need_vec(s, L, V_ZOLD, n)); }
  if (kind == RLS_CGNR) { RLS_TRY(need_vec(s, L, V_P, n)); RLS_TRY(need_vec(s, L, V_V, n)); }
  if (kind == RLS_SPLITBREGMAN) RLS_TRY(need_vec(s, L, V_Y, n));     // y = A'b, kept for the Bregman updates
  if (admm_like(kind)) {
    RLS_TRY(need_vec(s, L, V_BETA, n)); RLS_TRY(need_vec(s, L, V_BETAY, n));
    RLS_TRY(need_vec(s, L, V_CGU, n)); RLS_TRY(need_vec(s, L, V_CGR, n)); RLS_TRY(need_vec(s, L, V_CGC, n));
    for (int i = 0; i < s->desc.n_reg; ++i) {
      RLS_TRY(need_vec(s, L, V_TZ0 + i, s->rows[i])); RLS_TRY(need_vec(s, L, V_TZOLD0 + i, s->rows[i]));
      RLS_TRY(need_vec(s, L, V_TU0 + i, s->rows[i])); RLS_TRY(need_vec(s, L, V_TUOLD0 + i, s->rows[i]));
      if (s->desc.reg[i].trafo == RLS_TRAFO_GRADIENT) RLS_TRY(need_vec(s, L, V_TTMP0 + i, s->rows[i]));
    }
  }
  if (!L.dS) {
    RLS_CUDA(cudaMalloc(&L.dS, sizeof(DevState)));
    RLS_CUDA(cudaMallocHost(&L.hS, sizeof(DevState)));
    memset(L.hS, 0, sizeof(DevState));
    L.hS->gamma = 1.f; L.hS->gamma_old = 1.f; L.hS->beta = 1.f; L.hS->sigma = 1.f;  // POGM.jl:110-111 ctor values
  }
  return RLS_OK;
}

// write the configuration part of the device state (keeps POGM's γ across init!, quirk §9.1)
static int32_t push_config(rls_solver_s* s, Lane& L) {
  DevState* h = L.hS;
  const rls_solver_desc& d = s->desc;
  h->kind = d.kind; h->iterations = d.iterations; h->restart = d.restart;
  h->n_cap = (int)std::min<int64_t>(d.iterations, s->n);
  h->iterations_cg = d.iterations_cg; h->vary_rho = d.vary_rho; h->n_reg = d.n_reg; h->proj_mask = d.proj_mask;
  h->rho = d.rho; h->theta0 = d.theta; h->sigma_fac = d.sigma_fac; h->rel_tol = d.rel_tol; h->abs_tol = d.abs_tol;
  h->tol_inner = d.tol_inner;
  h->sb = d.kind == RLS_SPLITBREGMAN ? 1 : 0;
  h->iterations_inner = d.iterations_inner;
  for (int i = 0; i < 4; ++i) {
    h->lam[i] = (float)d.reg[i].lambda; h->lam64[i] = d.reg[i].lambda; h->lam_is_f64[i] = d.reg[i].lambda_is_f64;
    h->rho0[i] = d.reg[i].rho;
  }
  h->b_len = s->m;
  h->done = 1;
  h->cgi_gate = 1;
  RLS_CUDA(cudaMemcpyAsync(L.dS, h, sizeof(DevState), cudaMemcpyHostToDevice, s->ctx->stream));
  return RLS_OK;
}

static int32_t pull_state(rls_solver_s* s, Lane& L) {
  RLS_CUDA(cudaMemcpyAsync(L.hS, L.dS, sizeof(DevState), cudaMemcpyDeviceToHost, s->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(s->ctx->stream));
  RLS_CUDA(cudaGetLastError());
  if (s->pk_used) {                 // the cooperative whole-solve kernel ran: a timed-out grid barrier sets its abort flag
    s->pk_used = false;
    int flag = 0;
    RLS_CUDA(cudaMemcpy(&flag, (char*)s->pk_mem + s->pk_bytes - 8, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) { s->pk_off = true; rls_set_error("CGNR whole-solve kernel timed out on a grid barrier (abort flag set)"); return RLS_ERR_CUDA; }
  }
  return rls_normal_check_abort(s->AHA);
}

static void swap_roles(rls_solver_s* s, Lane& L) {
  switch (s->desc.kind) {
    case RLS_FISTA: std::swap(L.v[V_X], L.v[V_XOLD]); break;
    case RLS_POGM: std::swap(L.v[V_X], L.v[V_Y]); break;
    case RLS_SPLITBREGMAN:
    case RLS_ADMM:
      for (int i = 0; i < s->desc.n_reg; ++i) std::swap(L.v[V_TZ0 + i], L.v[V_TZOLD0 + i]);
      break;
    default: break;
  }
}

// after a synchronisation: iterations that were enqueued but gated off must not count as swaps
static void reconcile_roles(rls_solver_s* s, Lane& L) {
  // SplitBregman's `iteration` restarts at every Bregman update: it counts executed iterations in sb_total
  const int count = s->desc.kind == RLS_SPLITBREGMAN ? L.hS->sb_total : L.hS->iteration;
  const int executed = count - L.base_iter;
  if (((L.enq_swaps - executed) & 1) != 0) swap_roles(s, L);
  L.enq_swaps = 0;
  L.base_iter = count;
}

template <typename T> static T* P(rls_vec_s* v) { return v ? (T*)v->d : nullptr; }

// ---------------------------------- init! ---------------------------------------------
template <typename T>
static int32_t init_lane_t(rls_solver_s* s, Lane& L, const void* b_dev, int64_t b_len, const void* x0_dev, bool have_atb = false) {
  rls_ctx_s* c = s->ctx;
  cudaStream_t st = c->stream;
  const int64_t n = s->n;
  const int kind = s->desc.kind;
  const int g = ew_grid(c, n);
  const size_t nb = (size_t)n * sizeof(T);
  s->m = b_len;
  if (c->nranks > 1 && s->A) {
    // row-sharded: length(b) in σ_abs = sqrt(length(b))·absTol (ADMM.jl:212) is the GLOBAL row count, so that all
    // ranks take the same stopping decision
    double len = (double)b_len;
    RLS_TRY(rls_allreduce_f64_host(c, &len, 1));
    s->m = (int64_t)(len + 0.5);
  }
  RLS_TRY(push_config(s, L));
  L.enq_swaps = 0;
  L.base_iter = 0;
  // x₀ / β_y = A' b   (or b itself when only AHA was given)
  T* x0v = P<T>(admm_like(kind) ? L.v[V_BETAY] : L.v[V_X0]);
  if (s->A && have_atb) {
    // the multi-RHS driver already formed A'b for all columns with one GEMM (and all-reduced it)
  } else if (s->A) {
    RLS_TRY(rls_gemv_c_raw(s->A, b_dev, x0v, nullptr));
    if (c->nranks > 1) RLS_TRY(rls_allreduce_raw(c, x0v, n * (Elem<T>::is_complex ? 2 : 1)));
  } else {
    RLS_CUDA(cudaMemcpyAsync(x0v, b_dev, nb, cudaMemcpyDeviceToDevice, st));
  }
  // x = x0 (argument) or 0
  if (x0_dev) RLS_CUDA(cudaMemcpyAsync(L.v[V_X]->d, x0_dev, nb, cudaMemcpyDeviceToDevice, st));
  else RLS_CUDA(cudaMemsetAsync(L.v[V_X]->d, 0, nb, st));
  const T inf = []() { if constexpr (Elem<T>::is_complex) return make_float2(INFINITY, 0.f); else return INFINITY; }();
  switch (kind) {
    case RLS_FISTA:
      RLS_CUDA(cudaMemsetAsync(L.v[V_XOLD]->d, 0, nb, st));
      fill_gated_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_RES]), inf, n, nullptr);
      break;
    case RLS_POGM:
      RLS_CUDA(cudaMemsetAsync(L.v[V_XOLD]->d, 0, nb, st));
      RLS_CUDA(cudaMemsetAsync(L.v[V_Y]->d, 0, nb, st));
      RLS_CUDA(cudaMemsetAsync(L.v[V_Z]->d, 0, nb, st));
      if (s->desc.restart) RLS_CUDA(cudaMemsetAsync(L.v[V_W]->d, 0, nb, st));
      fill_gated_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_RES]), inf, n, nullptr);
      break;
    case RLS_OPTISTA:
      RLS_CUDA(cudaMemcpyAsync(L.v[V_Y]->d, L.v[V_X]->d, nb, cudaMemcpyDeviceToDevice, st));
      RLS_CUDA(cudaMemcpyAsync(L.v[V_Z]->d, L.v[V_X]->d, nb, cudaMemcpyDeviceToDevice, st));
      RLS_CUDA(cudaMemcpyAsync(L.v[V_ZOLD]->d, L.v[V_X]->d, nb, cudaMemcpyDeviceToDevice, st));
      fill_gated_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_RES]), inf, n, nullptr);
      break;
    case RLS_CGNR:
      RLS_CHECK_ARG(!x0_dev, "CGNR: x0 != 0 is not supported (the reference path is broken upstream, CGNR.jl:119)");
      RLS_CUDA(cudaMemsetAsync(L.v[V_V]->d, 0, nb, st));
      RLS_CUDA(cudaMemcpyAsync(L.v[V_P]->d, L.v[V_X0]->d, nb, cudaMemcpyDeviceToDevice, st));
      break;
    case RLS_SPLITBREGMAN:
      RLS_CUDA(cudaMemcpyAsync(L.v[V_Y]->d, L.v[V_BETAY]->d, nb, cudaMemcpyDeviceToDevice, st));   // y .= β_y  :179
      // fall through
    case RLS_ADMM:
      RLS_CUDA(cudaMemsetAsync(L.v[V_XOLD]->d, 0, nb, st));
      for (int i = 0; i < s->desc.n_reg; ++i) {
        const size_t rb = (size_t)s->rows[i] * sizeof(T);
        if (s->desc.reg[i].trafo == RLS_TRAFO_GRADIENT) RLS_TRY(rls_grad_fwd_launch(c, s->dtype, L.v[V_X]->d, L.v[V_TZ0 + i]->d, s->geom[i], nullptr));
        else RLS_CUDA(cudaMemcpyAsync(L.v[V_TZ0 + i]->d, L.v[V_X]->d, rb, cudaMemcpyDeviceToDevice, st));
        RLS_CUDA(cudaMemsetAsync(L.v[V_TU0 + i]->d, 0, rb, st));
        RLS_CUDA(cudaMemsetAsync(L.v[V_TZOLD0 + i]->d, 0, rb, st));
        RLS_CUDA(cudaMemsetAsync(L.v[V_TUOLD0 + i]->d, 0, rb, st));
      }
      break;
  }
  c->launches++;
  if (admm_like(kind)) {
    scalar_kernel<<<1, 32, 0, st>>>(L.dS, STEP_ADMM_INIT, 0, nullptr);
  } else {
    norm_step_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_X0]), n, L.dS, STEP_INIT, 0, c->red_partials, c->red_ticket, nullptr);
  }
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// ---------------------------------- iterate -------------------------------------------
// ADMM / SplitBregman: one iteration is 1 + iterationsCG (+1 for the Bregman update) operator applies with elementwise
// work between them.  admm_segment enqueues the kernels of segment `seg` and returns the apply that must run before
// segment seg + 1 (none after the last segment).  The single-RHS path executes that apply at once; the multi-RHS driver
// (MultiThreading.jl:45-78 batches any solver) collects the K pending applies of a segment and runs them as ONE batched
// apply — two tensor-core GEMMs that read A once each — with the per-column device gates (done / cgi_gate / sb_outer_gate).
struct PendingApply { const void* x; void* out; const int* gate; };
static int admm_segment_count(const rls_solver_desc& d) { return d.iterations_cg + 2 + (d.kind == RLS_SPLITBREGMAN ? 1 : 0); }

// out += Σ ρ_i (Φ_i'Φ_i v): the part of the composite operator AHA + Σ ρ Φ'Φ (ADMM.jl:141-159) that is not AHA
template <typename T>
static int32_t composite_extras(rls_solver_s* s, Lane& L, const T* v, T* out, const int* gate, bool fuse_identity_later) {
  rls_ctx_s* c = s->ctx;
  if (fuse_identity_later) return RLS_OK;
  const int g = ew_grid(c, s->n);
  for (int i = 0; i < s->desc.n_reg; ++i) {
    if (s->desc.reg[i].trafo == RLS_TRAFO_GRADIENT) {
      RLS_TRY(rls_grad_fwd_launch(c, s->dtype, v, L.v[V_TTMP0 + i]->d, s->geom[i], gate));
      RLS_TRY(rls_grad_t_axpy_launch(c, s->dtype, L.v[V_TTMP0 + i]->d, out, out, 0.f, &L.dS->a_rho[i], 1.f, s->geom[i], gate));
    } else {
      admm_axpy_rho_kernel<T><<<g, EB, 0, c->stream>>>(out, v, s->n, L.dS, i, gate);
      c->launches++;
    }
  }
  return RLS_OK;
}

template <typename T>
static int32_t admm_segment(rls_solver_s* s, Lane& L, int seg, PendingApply* pa) {
  rls_ctx_s* c = s->ctx;
  cudaStream_t st = c->stream;
  const int64_t n = s->n;
  const int g = ew_grid(c, n);
  DevState* S = L.dS;
  const int* gate = &S->done;
  const int* cgate = &S->cgi_gate;
  double* part = c->red_partials;
  unsigned* tick = c->red_ticket;
  const int k = s->desc.n_reg;
  const int ncg = s->desc.iterations_cg;
  bool all_identity = true;
  for (int i = 0; i < k; ++i) all_identity &= (s->desc.reg[i].trafo == RLS_TRAFO_IDENTITY);
  T* beta = P<T>(L.v[V_BETA]);
  // after the apply of segment 0 (c = AHA x): the rest of the composite operator, then r = β - c, ‖r‖ -> CG scalars
  auto after_first_apply = [&]() -> int32_t {
    RLS_TRY(composite_extras<T>(s, L, P<T>(L.v[V_X]), P<T>(L.v[V_CGC]), gate, false));
    admm_cg_init_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_CGR]), beta, P<T>(L.v[V_CGC]), n, S, part, tick);
    c->launches++;
    return RLS_OK;
  };
  // after the apply of a CG step (c = AHA u): α = res² / (u·c), x += α u, r -= α c, ‖r‖
  auto after_cg_apply = [&]() -> int32_t {
    RLS_TRY(composite_extras<T>(s, L, P<T>(L.v[V_CGU]), P<T>(L.v[V_CGC]), cgate, all_identity));
    admm_cg_dot_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_CGU]), P<T>(L.v[V_CGC]), n, S, all_identity ? k : 0, part, tick);
    admm_cg_xr_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_X]), P<T>(L.v[V_CGR]), P<T>(L.v[V_CGU]), P<T>(L.v[V_CGC]), n, S, part, tick);
    c->launches += 2;
    return RLS_OK;
  };
  if (seg == 0) {
    scalar_kernel<<<1, 32, 0, st>>>(S, STEP_ADMM_ITER_BEGIN, 0, gate);
    c->launches++;
    // 1. β = A'b + Σ ρ Φ'(z - u)
    for (int i = 0; i < k; ++i) {
      if (s->desc.reg[i].trafo == RLS_TRAFO_IDENTITY) {
        admm_beta_identity_kernel<T><<<g, EB, 0, st>>>(beta, P<T>(L.v[V_BETAY]), P<T>(L.v[V_TZ0 + i]), P<T>(L.v[V_TU0 + i]), n, S, i, i == 0);
        c->launches++;
      } else {
        if (i == 0) { copy_kernel<T><<<g, EB, 0, st>>>(beta, P<T>(L.v[V_BETAY]), n, gate); c->launches++; }
        RLS_TRY(rls_grad_t_axpy_launch(c, s->dtype, L.v[V_TZ0 + i]->d, beta, beta, 0.f, &S->a_rho[i], 1.f, s->geom[i], gate));
        RLS_TRY(rls_grad_t_axpy_launch(c, s->dtype, L.v[V_TU0 + i]->d, beta, beta, 0.f, &S->a_rho[i], -1.f, s->geom[i], gate));
      }
    }
    if (k == 0) { copy_kernel<T><<<g, EB, 0, st>>>(beta, P<T>(L.v[V_BETAY]), n, gate); c->launches++; }
    copy_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_XOLD]), P<T>(L.v[V_X]), n, gate);                    // xᵒˡᵈ = x   :243
    // cg!(x, AHA + Σ ρ Φ'Φ, β)   :244  — warm start: r = β - (AHA + ...) x
    fill_gated_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_CGU]), T{}, n, gate);
    c->launches += 2;
    *pa = PendingApply{L.v[V_X]->d, L.v[V_CGC]->d, gate};
  } else if (seg <= ncg) {
    if (seg == 1) RLS_TRY(after_first_apply());
    else RLS_TRY(after_cg_apply());
    admm_cg_u_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_CGU]), P<T>(L.v[V_CGR]), n, S);               // u = r + β u
    c->launches++;
    *pa = PendingApply{L.v[V_CGU]->d, L.v[V_CGC]->d, cgate};
  } else if (seg == ncg + 1) {
    if (ncg == 0) RLS_TRY(after_first_apply());
    else RLS_TRY(after_cg_apply());
    RLS_TRY(rls_proj_launch(c, s->dtype, L.v[V_X]->d, n, s->desc.proj_mask, gate));                 // :246-248
    // 2./3. z, u, residuals per term
    swap_roles(s, L); L.enq_swaps++;                                                                 // z <-> zᵒˡᵈ  :253-255
    for (int i = 0; i < k; ++i) {
      const rls_reg_desc& ri = s->desc.reg[i];
      T* z = P<T>(L.v[V_TZ0 + i]); T* zo = P<T>(L.v[V_TZOLD0 + i]); T* u = P<T>(L.v[V_TU0 + i]); T* uo = P<T>(L.v[V_TUOLD0 + i]);
      if (ri.trafo == RLS_TRAFO_IDENTITY) {
        if (rls_reg_is_elementwise(ri.kind)) {
          admm_term_identity_kernel<T, 0><<<g, EB, 0, st>>>(P<T>(L.v[V_X]), P<T>(L.v[V_XOLD]), z, zo, u, uo, n, S, i, ri.kind, part, tick);
          c->launches++;
        } else {
          admm_term_identity_kernel<T, 1><<<g, EB, 0, st>>>(P<T>(L.v[V_X]), P<T>(L.v[V_XOLD]), z, zo, u, uo, n, S, i, ri.kind, part, tick);
          RLS_TRY(rls_prox_launch(c, s->dtype, z, n, &ri, 0.f, &S->a_thr[i], gate, &s->tv));
          admm_term_identity_kernel<T, 2><<<g, EB, 0, st>>>(P<T>(L.v[V_X]), P<T>(L.v[V_XOLD]), z, zo, u, uo, n, S, i, ri.kind, part, tick);
          c->launches += 2;
        }
      } else {
        RLS_TRY(rls_admm_term_gradient(c, s->dtype, L.v[V_X]->d, L.v[V_XOLD]->d, z, zo, u, uo, s->geom[i], S, i, ri.kind));
      }
      if (s->desc.vary_rho != RLS_VARY_RHO_NONE) {
        admm_scale_u_kernel<T><<<ew_grid(c, s->rows[i]), EB, 0, st>>>(u, s->rows[i], S, i);
        c->launches++;
      }
    }
    scalar_kernel<<<1, 32, 0, st>>>(S, STEP_ADMM_ITER_END, 0, gate);
    c->launches++;
    // SplitBregman: Bregman update, gated on the device flag set by ITER_END (converged || iteration >= iterationsInner)
    if (s->desc.kind == RLS_SPLITBREGMAN) *pa = PendingApply{L.v[V_X]->d, L.v[V_CGC]->d, &S->sb_outer_gate};
  } else {
    const int* ogate = &S->sb_outer_gate;
    sb_outer_kernel<T><<<g, EB, 0, st>>>(P<T>(L.v[V_BETAY]), P<T>(L.v[V_Y]), P<T>(L.v[V_CGC]), n, S);
    for (int i = 0; i < k; ++i) {
      if (s->desc.reg[i].trafo == RLS_TRAFO_GRADIENT) {                 // z = Φx ; u = 0   :261-262
        RLS_TRY(rls_grad_fwd_launch(c, s->dtype, L.v[V_X]->d, L.v[V_TZ0 + i]->d, s->geom[i], ogate));
        fill_gated_kernel<T><<<ew_grid(c, s->rows[i]), EB, 0, st>>>(P<T>(L.v[V_TU0 + i]), T{}, s->rows[i], ogate);
      } else {
        sb_reset_term_kernel<T><<<ew_grid(c, s->rows[i]), EB, 0, st>>>(P<T>(L.v[V_TZ0 + i]), P<T>(L.v[V_TU0 + i]), P<T>(L.v[V_X]), s->rows[i], S);
      }
    }
    scalar_kernel<<<1, 32, 0, st>>>(S, STEP_SB_FINISH, 0, gate);
    c->launches += 2 + k;
  }
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// phase 0: the whole iteration.  The multi-RHS driver splits it around the normal-operator apply so that the K
// applies of one batched iteration run as two tensor-core GEMMs: phase 1 = everything before the apply (its
// operands are recorded in s->batch_*), phase 2 = everything after it.
enum { IT_ALL = 0, IT_PRE = 1, IT_POST = 2 };

template <typename T>
static int32_t enqueue_iteration_t(rls_solver_s* s, Lane& L, int phase = IT_ALL) {
  rls_ctx_s* c = s->ctx;
  cudaStream_t st = c->stream;
  const int64_t n = s->n;
  const int g = ew_grid(c, n);
  DevState* S = L.dS;
  const int* gate = &S->done;
  const rls_reg_desc& reg = s->desc.reg[0];
  const bool ew = rls_reg_is_elementwise(reg.kind);
  double* part = c->red_partials;
  unsigned* tick = c->red_ticket;
  auto record = [&](const void* x, void* res) {
    s->batch_x.push_back(x); s->batch_res.push_back(res); s->batch_gate.push_back(gate);
  };
  switch (s->desc.kind) {
    case RLS_FISTA: {
      NormalPartials np{nullptr, 0, 0};
      int fuse = 0;
      if (phase != IT_POST) { swap_roles(s, L); L.enq_swaps++; }
      if (phase == IT_ALL) {
        // single GPU + row-major one-pass operator: momentum fused into the operator's x load and into the
        // epilogue, finish fused into the epilogue — two kernels per iteration instead of four
        RLS_TRY(rls_normal_apply_deferred_raw(s->AHA, L.v[V_X]->d, (const float*)L.v[V_XOLD]->d, &S->theta_old, &S->theta, gate, &np));
        fuse = np.ncl > 0;
      }
      if (!fuse) {
        if (phase != IT_POST) {
          RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), fista_momentum_kernel<T>, P<T>(L.v[V_X]), P<T>(L.v[V_XOLD]), n, S));
          c->launches++;
        }
        if (phase == IT_PRE) { record(L.v[V_X]->d, L.v[V_RES]->d); break; }
        if (phase == IT_ALL) RLS_TRY(rls_normal_apply_raw(s->AHA, L.v[V_X]->d, L.v[V_RES]->d, gate));
      }
      if (ew) {
        rls_trace_begin(st, "fista_main");
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), fista_main_kernel<T, 0>, P<T>(L.v[V_X]), P<T>(L.v[V_RES]), P<T>(L.v[V_X0]), P<T>(L.v[V_XOLD]), n, S, reg.kind, part, tick, np, fuse));
        rls_trace_end(st);
        c->launches += 1;
      } else {
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), fista_main_kernel<T, 1>, P<T>(L.v[V_X]), P<T>(L.v[V_RES]), P<T>(L.v[V_X0]), P<T>(L.v[V_XOLD]), n, S, reg.kind, part, tick, np, fuse));
        RLS_TRY(rls_prox_launch(c, s->dtype, L.v[V_X]->d, n, &reg, 0.f, &S->thr, gate, &s->tv));
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), fista_main_kernel<T, 2>, P<T>(L.v[V_X]), P<T>(L.v[V_RES]), P<T>(L.v[V_X0]), P<T>(L.v[V_XOLD]), n, S, reg.kind, part, tick, NormalPartials{nullptr, 0, 0}, 0));
        c->launches += 2;
      }
      break;
    }
    case RLS_POGM: {
      if (phase != IT_POST) RLS_CUDA(rls_launch_pdl(st, dim3(1), dim3(32), scalar_kernel, S, STEP_POGM_PRE, 0, gate));
      if (phase == IT_PRE) { record(L.v[V_X]->d, L.v[V_RES]->d); c->launches++; break; }
      if (phase == IT_ALL) RLS_TRY(rls_normal_apply_raw(s->AHA, L.v[V_X]->d, L.v[V_RES]->d, gate));
      T* bx = P<T>(L.v[V_X]); T* by = P<T>(L.v[V_Y]);
      if (ew) {
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), pogm_main_kernel<T, 0>, bx, by, P<T>(L.v[V_X0]), P<T>(L.v[V_XOLD]), P<T>(L.v[V_Z]), P<T>(L.v[V_W]), P<T>(L.v[V_RES]), n, S, reg.kind, part, tick));
        c->launches += 2;
      } else {
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), pogm_main_kernel<T, 1>, bx, by, P<T>(L.v[V_X0]), P<T>(L.v[V_XOLD]), P<T>(L.v[V_Z]), P<T>(L.v[V_W]), P<T>(L.v[V_RES]), n, S, reg.kind, part, tick));
        RLS_TRY(rls_prox_launch(c, s->dtype, by, n, &reg, 0.f, &S->thr, gate, &s->tv));
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), pogm_main_kernel<T, 2>, bx, by, P<T>(L.v[V_X0]), P<T>(L.v[V_XOLD]), P<T>(L.v[V_Z]), P<T>(L.v[V_W]), P<T>(L.v[V_RES]), n, S, reg.kind, part, tick));
        c->launches += 3;
      }
      swap_roles(s, L); L.enq_swaps++;   // x <-> y  (POGM.jl:206-208)
      break;
    }
    case RLS_OPTISTA: {
      if (phase != IT_POST) RLS_CUDA(rls_launch_pdl(st, dim3(1), dim3(32), scalar_kernel, S, STEP_OPTISTA_PRE, 0, gate));
      if (phase == IT_PRE) { record(L.v[V_X]->d, L.v[V_RES]->d); c->launches++; break; }
      if (phase == IT_ALL) RLS_TRY(rls_normal_apply_raw(s->AHA, L.v[V_X]->d, L.v[V_RES]->d, gate));
      if (ew) {
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), optista_main_kernel<T, 0>, P<T>(L.v[V_X]), P<T>(L.v[V_Y]), P<T>(L.v[V_Z]), P<T>(L.v[V_ZOLD]), P<T>(L.v[V_X0]), P<T>(L.v[V_RES]), n, S, reg.kind, part, tick));
        c->launches += 2;
      } else {
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), optista_main_kernel<T, 1>, P<T>(L.v[V_X]), P<T>(L.v[V_Y]), P<T>(L.v[V_Z]), P<T>(L.v[V_ZOLD]), P<T>(L.v[V_X0]), P<T>(L.v[V_RES]), n, S, reg.kind, part, tick));
        RLS_TRY(rls_prox_launch(c, s->dtype, L.v[V_Y]->d, n, &reg, 0.f, &S->thr, gate, &s->tv));
        RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), optista_main_kernel<T, 2>, P<T>(L.v[V_X]), P<T>(L.v[V_Y]), P<T>(L.v[V_Z]), P<T>(L.v[V_ZOLD]), P<T>(L.v[V_X0]), P<T>(L.v[V_RES]), n, S, reg.kind, part, tick));
        c->launches += 3;
      }
      break;
    }
    case RLS_CGNR: {
      if (phase == IT_PRE) { record(L.v[V_P]->d, L.v[V_V]->d); break; }
      if (phase == IT_ALL) RLS_TRY(rls_normal_apply_raw(s->AHA, L.v[V_P]->d, L.v[V_V]->d, gate));
      RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), cgnr_dot_kernel<T>, P<T>(L.v[V_P]), P<T>(L.v[V_V]), n, S, part, tick));
      RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), cgnr_update_kernel<T>, P<T>(L.v[V_X]), P<T>(L.v[V_X0]), P<T>(L.v[V_P]), P<T>(L.v[V_V]), n, S, part, tick));
      RLS_CUDA(rls_launch_pdl(st, dim3(g), dim3(EB), cgnr_p_kernel<T>, P<T>(L.v[V_P]), P<T>(L.v[V_X0]), n, S, part, tick));
      c->launches += 3;
      break;
    }
    case RLS_SPLITBREGMAN:
    case RLS_ADMM: {
      // the iteration as a chain of segments, each ending in one operator apply (admm_segment below)
      const int nseg = admm_segment_count(s->desc);
      for (int seg = 0; seg < nseg; ++seg) {
        PendingApply pa{nullptr, nullptr, nullptr};
        RLS_TRY(admm_segment<T>(s, L, seg, &pa));
        if (pa.x) RLS_TRY(rls_normal_apply_raw(s->AHA, pa.x, pa.out, pa.gate));
      }
      break;
    }
    default:
      rls_set_error("unknown solver kind %d", s->desc.kind);
      return RLS_ERR_INVALID;
  }
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

static int32_t init_lane(rls_solver_s* s, Lane& L, const void* b, int64_t blen, const void* x0, bool have_atb = false) {
  if (s->dtype == RLS_C32) return init_lane_t<float2>(s, L, b, blen, x0, have_atb);
  return init_lane_t<float>(s, L, b, blen, x0, have_atb);
}
static int32_t enqueue_iteration(rls_solver_s* s, Lane& L, int phase = IT_ALL) {
  if (s->dtype == RLS_C32) return enqueue_iteration_t<float2>(s, L, phase);
  return enqueue_iteration_t<float>(s, L, phase);
}

static void fill_scalars(const DevState* h, rls_solver_scalars* o) {
  memset(o, 0, sizeof(*o));
  o->iteration = h->iteration; o->done = h->done;
  o->rho = h->rho; o->theta = h->theta; o->theta_old = h->theta_old; o->theta_n = h->theta_n;
  o->alpha = h->alpha; o->beta = h->beta; o->gamma = h->gamma; o->gamma_old = h->gamma_old; o->sigma = h->sigma;
  o->norm_x0 = h->norm_x0; o->rel_res_norm = h->rel_res_norm; o->res_norm = h->res_norm;
  o->cg_alpha[0] = h->cg_alpha.x; o->cg_alpha[1] = h->cg_alpha.y;
  o->cg_beta[0] = h->cg_beta.x; o->cg_beta[1] = h->cg_beta.y;
  o->cg_zeta[0] = h->cg_zeta.x; o->cg_zeta[1] = h->cg_zeta.y;
  for (int i = 0; i < 4; ++i) {
    o->admm_rk[i] = h->a_rk[i]; o->admm_sk[i] = h->a_sk[i]; o->admm_eps_pri[i] = h->a_eps_pri[i];
    o->admm_eps_dua[i] = h->a_eps_dua[i]; o->admm_delta[i] = h->a_delta[i]; o->admm_rho[i] = h->a_rho[i];
  }
  o->admm_sigma_abs = h->sigma_abs;
  o->cg_iterations_last = h->cgi_last;
  o->cg_iterations_total = h->cgi_total;
  o->outer_iteration = h->sb_iter_cnt;
}

static int32_t validate_desc(const rls_solver_desc* d) {
  RLS_CHECK_ARG(d->kind >= RLS_FISTA && d->kind <= RLS_SPLITBREGMAN, "unknown solver kind %d", d->kind);
  RLS_CHECK_ARG(d->iterations >= 0, "iterations must be >= 0");
  if (admm_like(d->kind)) {
    RLS_CHECK_ARG(d->n_reg >= 1 && d->n_reg <= 4, "ADMM / SplitBregman support 1..4 regularization terms, got %d", d->n_reg);
    RLS_CHECK_ARG(d->iterations_cg >= 0, "iterationsCG must be >= 0");
    if (d->kind == RLS_SPLITBREGMAN) {
      RLS_CHECK_ARG(d->iterations_inner >= 1, "iterationsInner must be >= 1");
      RLS_CHECK_ARG(d->vary_rho == RLS_VARY_RHO_NONE, "SplitBregman has no vary_rho keyword");
    }
  } else if (d->kind == RLS_CGNR) {
    RLS_CHECK_ARG(d->n_reg <= 1, "CGNR does not allow for more additional regularization terms, found %d", d->n_reg);
    RLS_CHECK_ARG(d->n_reg == 0 || d->reg[0].kind == RLS_REG_L2 || d->reg[0].kind == RLS_REG_NONE,
                  "CGNR only accepts L2Regularization (CGNR.jl:57-63)");
  } else {
    // FISTA.jl:83-85, POGM.jl:104-106, OptISTA.jl:95-97
    RLS_CHECK_ARG(d->n_reg == 1, "%s does not allow for more additional regularization terms, found %d",
                  d->kind == RLS_FISTA ? "FISTA" : (d->kind == RLS_POGM ? "POGM" : "OptISTA"), d->n_reg);
  }
  for (int i = 0; i < d->n_reg; ++i) {
    const rls_reg_desc& r = d->reg[i];
    RLS_CHECK_ARG(r.kind >= RLS_REG_NONE && r.kind <= RLS_REG_LLR, "unknown regularization kind %d", r.kind);
    RLS_CHECK_ARG(r.trafo == RLS_TRAFO_IDENTITY || r.trafo == RLS_TRAFO_GRADIENT, "unknown regTrafo %d", r.trafo);
    if (r.trafo == RLS_TRAFO_GRADIENT) {
      RLS_CHECK_ARG(admm_like(d->kind), "regTrafo is an ADMM / SplitBregman keyword");
      if (!rls_reg_is_elementwise(r.kind)) {
        rls_set_error("ADMM / SplitBregman with a GradientOp regTrafo support elementwise prox (L1/L2) only");
        return RLS_ERR_UNSUPPORTED;
      }
    }
  }
  return RLS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" int32_t rls_solver_create(rls_mat_t A, rls_normal_t AHA, const rls_solver_desc* desc, rls_solver_t* out) {
  RLS_CHECK_ARG(desc && out, "NULL argument");
  RLS_CHECK_ARG(A || AHA, "either A or AHA must be given");
  RLS_TRY(validate_desc(desc));
  rls_ctx_s* ctx = A ? A->ctx : rls_normal_ctx(AHA);
  RlsDeviceGuard g(ctx->device);
  rls_solver_s* s = new rls_solver_s();
  s->ctx = ctx;
  s->A = A;
  s->desc = *desc;
  if (AHA) {
    s->AHA = AHA;
  } else {
    int32_t st = rls_normal_create(A, RLS_NORMAL_AUTO, &s->AHA);
    if (st != RLS_OK) { delete s; return st; }
    s->own_AHA = true;
  }
  // the solver keeps what it points to alive: handles are freed by garbage collectors in arbitrary order
  rls_ctx_retain(ctx);
  rls_mat_retain(A);
  if (!s->own_AHA) rls_normal_retain(s->AHA);
  rls_normal_shape(s->AHA, &s->n, &s->dtype);
  if (A && (A->n != s->n || A->dtype != s->dtype)) {
    rls_solver_destroy(s);
    rls_set_error("A and AHA disagree in shape or dtype");
    return RLS_ERR_INVALID;
  }
  s->m = A ? A->m : s->n;
  for (int i = 0; i < desc->n_reg; ++i) {
    const rls_reg_desc& r = desc->reg[i];
    s->rows[i] = s->n;
    if (r.trafo == RLS_TRAFO_GRADIENT || r.kind == RLS_REG_TV) {
      int32_t st = rls_make_grad_geom(r.tv_ndims, r.tv_shape, r.tv_ndirs, r.tv_dims, &s->geom[i]);
      if (st == RLS_OK && s->geom[i].npix != s->n) {
        rls_set_error("regularization %d: image shape has %lld pixels but the solution has %lld", i, (long long)s->geom[i].npix, (long long)s->n);
        st = RLS_ERR_INVALID;
      }
      if (st != RLS_OK) { rls_solver_destroy(s); return st; }
      if (r.trafo == RLS_TRAFO_GRADIENT) s->rows[i] = s->geom[i].rows;
    }
  }
  s->lanes.resize(1);
  int32_t st = alloc_lane(s, s->lanes[0]);
  if (st == RLS_OK) st = push_config(s, s->lanes[0]);
  if (st != RLS_OK) { rls_solver_destroy(s); return st; }
  *out = s;
  return RLS_OK;
}

extern "C" int32_t rls_solver_destroy(rls_solver_t s) {
  if (!s) return RLS_OK;
  RlsDeviceGuard g(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  for (Lane& L : s->lanes) free_lane(L);
  if (s->graph) cudaGraphExecDestroy(s->graph);
  if (s->pk_mem) cudaFree(s->pk_mem);
  rls_tv_work_free(&s->tv);
  if (s->b_dev) rls_vec_destroy(s->b_dev);
  if (s->pin_b) cudaFreeHost(s->pin_b);
  if (s->pin_x) cudaFreeHost(s->pin_x);
  rls_ctx_s* c = s->ctx;
  rls_mat_s* A = s->A;
  rls_normal_release(s->AHA);   // own or borrowed: one reference either way
  delete s;
  rls_mat_release(A);
  rls_ctx_release(c);
  return RLS_OK;
}

extern "C" int32_t rls_solver_set_reg(rls_solver_t s, int32_t idx, const rls_reg_desc* reg) {
  RLS_CHECK_ARG(s && reg, "NULL argument");
  RLS_CHECK_ARG(idx >= 0 && idx < (s->desc.n_reg > 0 ? s->desc.n_reg : 1), "regularization index %d out of range", idx);
  RLS_CHECK_ARG(reg->kind == s->desc.reg[idx].kind && reg->trafo == s->desc.reg[idx].trafo,
                "set_reg may change λ / ρ only, not the kind of term");
  RlsDeviceGuard g(s->ctx->device);
  s->desc.reg[idx].lambda = reg->lambda;
  s->desc.reg[idx].lambda_is_f64 = reg->lambda_is_f64;
  s->desc.reg[idx].rho = reg->rho;
  // λ is re-normalised inside init! upstream (FISTA.jl:128): refresh the device copy of an initialised lane
  for (Lane& L : s->lanes) {
    if (!L.dS) continue;
    L.hS->lam[idx] = (float)reg->lambda; L.hS->lam64[idx] = reg->lambda; L.hS->lam_is_f64[idx] = reg->lambda_is_f64;
    L.hS->rho0[idx] = reg->rho;
    RLS_CUDA(cudaMemcpyAsync(&L.dS->lam[idx], &L.hS->lam[idx], sizeof(float), cudaMemcpyHostToDevice, s->ctx->stream));
    RLS_CUDA(cudaMemcpyAsync(&L.dS->lam64[idx], &L.hS->lam64[idx], sizeof(double), cudaMemcpyHostToDevice, s->ctx->stream));
    RLS_CUDA(cudaMemcpyAsync(&L.dS->lam_is_f64[idx], &L.hS->lam_is_f64[idx], sizeof(int), cudaMemcpyHostToDevice, s->ctx->stream));
    RLS_CUDA(cudaMemcpyAsync(&L.dS->rho0[idx], &L.hS->rho0[idx], sizeof(float), cudaMemcpyHostToDevice, s->ctx->stream));
  }
  return RLS_OK;
}

extern "C" int32_t rls_solver_init(rls_solver_t s, rls_vec_t b, rls_vec_t x0) {
  RLS_CHECK_ARG(s && b, "NULL argument");
  RLS_CHECK_ARG(b->dtype == s->dtype, "init!: b has the wrong element type");
  if (s->A) RLS_CHECK_ARG(b->len == s->A->m, "init!: b has %lld elements, A has %lld rows", (long long)b->len, (long long)s->A->m);
  else RLS_CHECK_ARG(b->len == s->n, "init!: with AHA only, b must be A'b of length %lld", (long long)s->n);
  if (x0) RLS_CHECK_ARG(x0->len == s->n && x0->dtype == s->dtype, "init!: x0 shape/dtype mismatch");
  RlsNvtxRange nvtx("rls: init!");
  RlsDeviceGuard g(s->ctx->device);
  if (s->lanes.size() != 1) {
    for (size_t k = 1; k < s->lanes.size(); ++k) free_lane(s->lanes[k]);
    s->lanes.resize(1);
  }
  RLS_TRY(init_lane(s, s->lanes[0], b->d, b->len, x0 ? x0->d : nullptr));
  RLS_TRY(pull_state(s, s->lanes[0]));
  return RLS_OK;
}

extern "C" int32_t rls_solver_iterate(rls_solver_t s, int32_t* advanced, rls_solver_scalars* scalars) {
  RLS_CHECK_ARG(s && advanced, "NULL argument");
  RlsNvtxRange nvtx("rls: iterate");
  RlsDeviceGuard g(s->ctx->device);
  Lane& L = s->lanes[0];
  if (L.hS->done) {
    // Julia's `iterate` returns nothing; CGNR applies its constraints at this point (CGNR.jl:144-149)
    if (s->desc.kind == RLS_CGNR && s->desc.proj_mask) {
      RLS_TRY(rls_proj_launch(s->ctx, s->dtype, L.v[V_X]->d, s->n, s->desc.proj_mask, nullptr));
      RLS_TRY(rls_ctx_sync(s->ctx));
    }
    *advanced = 0;
    if (scalars) fill_scalars(L.hS, scalars);
    return RLS_OK;
  }
  RLS_TRY(enqueue_iteration(s, L));
  RLS_TRY(pull_state(s, L));
  reconcile_roles(s, L);
  *advanced = 1;
  if (scalars) fill_scalars(L.hS, scalars);
  return RLS_OK;
}

// CUDA-graph replay of the iteration loop.  Eligible: CGNR (no host-side buffer rotation between iterations) on one rank
// with an operator whose applies are plain launches, a system small enough to be latency-bound, RLS_SOLVE_GRAPH != 0.
static bool graph_eligible(rls_solver_s* s) {
  if (s->graph_off || s->desc.kind != RLS_CGNR || s->ctx->nranks > 1 || s->lanes.size() != 1) return false;
  if (!rls_env_flag("RLS_SOLVE_GRAPH", true) || rls_trace_enabled() || !rls_normal_graph_safe(s->AHA)) return false;
  const double bytes = (double)(s->A ? s->A->m : s->n) * (double)s->n * (double)rls_elem_size(s->dtype);
  return bytes <= 512.0 * 1024 * 1024;
}

static void graph_drop(rls_solver_s* s) {
  if (s->graph) cudaGraphExecDestroy(s->graph);
  s->graph = nullptr;
}

static int32_t enqueue_all_iterations(rls_solver_s* s, Lane& L, int from, int cap) {
  for (int it = from; it < cap; ++it) RLS_TRY(enqueue_iteration(s, L));
  if (s->desc.kind == RLS_CGNR && s->desc.proj_mask)
    RLS_TRY(rls_proj_launch(s->ctx, s->dtype, L.v[V_X]->d, s->n, s->desc.proj_mask, nullptr));
  return RLS_OK;
}

static int32_t capture_iterations(rls_solver_s* s, Lane& L, int cap) {
  rls_ctx_s* c = s->ctx;
  const int64_t l0 = c->launches;
  if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return RLS_ERR_UNSUPPORTED; }
  rls_pdl_suppress(true);
  const int32_t status = enqueue_all_iterations(s, L, 0, cap);
  rls_pdl_suppress(false);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c->stream, &g);       // always: a stream must not be left capturing
  const int64_t recorded = c->launches - l0;
  c->launches = l0;
  if (status != RLS_OK || e != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return RLS_ERR_UNSUPPORTED;
  }
  const cudaError_t ei = cudaGraphInstantiate(&s->graph, g, 0);
  cudaGraphDestroy(g);
  if (ei != cudaSuccess) { cudaGetLastError(); s->graph = nullptr; return RLS_ERR_UNSUPPORTED; }
  s->graph_launches = recorded;
  s->graph_scratch = c->gemv_scratch;
  return RLS_OK;
}

// ONE cooperative kernel for the whole CGNR solve of an L2-resident column-major system (cgnr_persistent_kernel above).
// Opt-in (RLS_CGNR_PERSISTENT=1): its sums run in another order than the chained kernels', so iterates agree with them to
// rounding, not bit for bit.
static bool persistent_eligible(rls_solver_s* s) {
  if (s->pk_off || s->desc.kind != RLS_CGNR || s->ctx->nranks > 1 || s->lanes.size() != 1) return false;
  if (!rls_env_flag("RLS_CGNR_PERSISTENT", false) || rls_trace_enabled()) return false;
  rls_mat_s* A = s->A;
  if (!A || A->layout != RLS_LAYOUT_COLMAJOR || rls_normal_matrix(s->AHA) != A) return false;
  int32_t form = -1;
  rls_normal_form(s->AHA, &form);
  if (form != RLS_NORMAL_TWOPASS && form != RLS_NORMAL_ONEPASS) return false;      // the lazy A'(A x) on this very matrix
  if (A->m < 1 || A->n < 1 || A->m > PK_THREADS * PK_MAXROWS) return false;
  const int64_t cn = (A->n + s->ctx->sm_count - 1) / s->ctx->sm_count;
  if (cn > PK_MAXCN) return false;
  return (double)A->m * (double)A->n * (double)rls_elem_size(A->dtype) <= 96.0 * 1024 * 1024;   // stays in the 126 MB L2
}

static int32_t run_persistent(rls_solver_s* s, Lane& L, int cap, bool* launched) {
  *launched = false;
  rls_ctx_s* c = s->ctx;
  rls_mat_s* A = s->A;
  const int P = c->sm_count;
  const size_t es = rls_elem_size(A->dtype);
  const int CN = (int)((A->n + P - 1) / P);
  const size_t fixed = ((size_t)A->m + 4 * (size_t)CN) * es;
  const void* fn = A->dtype == RLS_C32 ? (const void*)cgnr_persistent_kernel<float2> : (const void*)cgnr_persistent_kernel<float>;
  // as many own columns of A as fit stay in shared memory for the whole solve (C1: 27 of the 28 columns of a CTA)
  int optin = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) { cudaGetLastError(); s->pk_off = true; return RLS_OK; }
  const size_t budget = (size_t)optin > fa.sharedSizeBytes + fixed ? (size_t)optin - fa.sharedSizeBytes - fixed : 0;
  const int ncache = (int)std::min<size_t>((size_t)CN, budget / ((size_t)A->m * es));
  const size_t smem = fixed + (size_t)ncache * (size_t)A->m * es;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); s->pk_off = true; return RLS_OK; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PK_THREADS, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    s->pk_off = true;
    return RLS_OK;
  }
  // [ypart P*m][y m][dpart 5*P*4 doubles][barrier counter][abort flag]
  const size_t o_y = (size_t)P * A->m * es, o_d = ((o_y + (size_t)A->m * es + 15) / 16) * 16, o_b = o_d + (size_t)5 * P * 4 * 8;
  const size_t need = o_b + 16;
  if (s->pk_bytes < need) {
    if (s->pk_mem) RLS_CUDA(cudaFree(s->pk_mem));
    s->pk_mem = nullptr; s->pk_bytes = 0;
    if (cudaMalloc(&s->pk_mem, need) != cudaSuccess) { cudaGetLastError(); s->pk_off = true; return RLS_OK; }
    s->pk_bytes = need;
  }
  char* base = (char*)s->pk_mem;
  RLS_CUDA(cudaMemsetAsync(base + o_b, 0, 16, c->stream));
  PkArgs a{};
  a.A = A->d; a.ld = A->ld; a.m = (int)A->m; a.n = (int)A->n; a.CN = CN; a.ncache = ncache;
  a.x = L.v[V_X]->d; a.r = L.v[V_X0]->d; a.p = L.v[V_P]->d; a.v = L.v[V_V]->d;
  a.S = L.dS;
  a.ypart = base; a.y = base + o_y; a.dpart = (double*)(base + o_d);
  a.bar = (unsigned*)(base + o_b); a.abort_flag = (int*)(base + o_b + 8);
  a.cap = cap;
  void* args[] = {&a};
  RLS_CUDA(cudaLaunchCooperativeKernel(fn, dim3(P), dim3(PK_THREADS), args, smem, c->stream));
  c->launches++;
  s->pk_used = true;
  *launched = true;
  if (s->desc.proj_mask) RLS_TRY(rls_proj_launch(c, s->dtype, L.v[V_X]->d, s->n, s->desc.proj_mask, nullptr));   // CGNR.jl:146-149
  return RLS_OK;
}

static int32_t run_lane_async(rls_solver_s* s, Lane& L, int already_done) {
  const int cap = iteration_cap(s->desc, s->n);
  if (already_done == 0 && cap > 0 && persistent_eligible(s)) {
    bool launched = false;
    RLS_TRY(run_persistent(s, L, cap, &launched));
    if (launched) return RLS_OK;
  }
  if (already_done == 0 && cap > 0 && s->plain_solves >= 1 && graph_eligible(s)) {
    if (s->graph && s->graph_scratch != s->ctx->gemv_scratch) graph_drop(s);   // another operator grew the shared scratch
    if (!s->graph && capture_iterations(s, L, cap) != RLS_OK) { s->graph_off = true; s->graph = nullptr; }
    if (s->graph) {
      RLS_CUDA(cudaGraphLaunch(s->graph, s->ctx->stream));
      s->ctx->launches += s->graph_launches;
      return RLS_OK;
    }
  }
  RLS_TRY(enqueue_all_iterations(s, L, already_done, cap));
  if (already_done == 0) s->plain_solves++;
  return RLS_OK;
}

static int32_t solve_lane_async(rls_solver_s* s, Lane& L, const void* b, int64_t blen, const void* x0) {
  RLS_TRY(init_lane(s, L, b, blen, x0));
  return run_lane_async(s, L, 0);
}

// run the remaining iterations of an initialised solver without a host round trip per
// iteration (the `for _ in enumerate(solver)` loop of solve!, RegularizedLeastSquares.jl:112-114)
extern "C" int32_t rls_solver_run(rls_solver_t s, int32_t* iterations_done, rls_solver_scalars* scalars) {
  RLS_CHECK_ARG(s, "NULL argument");
  RlsNvtxRange nvtx("rls: solve! (iterate until done)");
  RlsDeviceGuard g(s->ctx->device);
  Lane& L = s->lanes[0];
  if (!L.hS->done) RLS_TRY(run_lane_async(s, L, s->desc.kind == RLS_SPLITBREGMAN ? L.hS->sb_total : L.hS->iteration));
  else if (s->desc.kind == RLS_CGNR && s->desc.proj_mask)
    RLS_TRY(rls_proj_launch(s->ctx, s->dtype, L.v[V_X]->d, s->n, s->desc.proj_mask, nullptr));
  RLS_TRY(pull_state(s, L));
  reconcile_roles(s, L);
  if (iterations_done) *iterations_done = L.hS->iteration;
  if (scalars) fill_scalars(L.hS, scalars);
  return RLS_OK;
}

extern "C" int32_t rls_solver_solve(rls_solver_t s, rls_vec_t b, rls_vec_t x0, int32_t* iterations_done,
                                    rls_solver_scalars* scalars) {
  RLS_CHECK_ARG(s && b, "NULL argument");
  RLS_CHECK_ARG(b->dtype == s->dtype, "solve!: b has the wrong element type");
  if (s->A) RLS_CHECK_ARG(b->len == s->A->m, "solve!: b has %lld elements, A has %lld rows", (long long)b->len, (long long)s->A->m);
  else RLS_CHECK_ARG(b->len == s->n, "solve!: with AHA only, b must be A'b of length %lld", (long long)s->n);
  if (x0) RLS_CHECK_ARG(x0->len == s->n && x0->dtype == s->dtype, "solve!: x0 shape/dtype mismatch");
  RlsNvtxRange nvtx("rls: solve!");
  RlsDeviceGuard g(s->ctx->device);
  if (s->lanes.size() != 1) {
    for (size_t k = 1; k < s->lanes.size(); ++k) free_lane(s->lanes[k]);
    s->lanes.resize(1);
  }
  Lane& L = s->lanes[0];
  RLS_TRY(solve_lane_async(s, L, b->d, b->len, x0 ? x0->d : nullptr));
  RLS_TRY(pull_state(s, L));
  rls_trace_dump();
  reconcile_roles(s, L);
  if (iterations_done) *iterations_done = L.hS->iteration;
  if (scalars) fill_scalars(L.hS, scalars);
  return RLS_OK;
}

static int32_t ensure_pinned(void** p, size_t* have, size_t need) {
  if (*have >= need) return RLS_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *have = 0;
  RLS_CUDA(cudaMallocHost(p, need));
  *have = need;
  return RLS_OK;
}

extern "C" int32_t rls_solver_solve_host(rls_solver_t s, const void* b_host, int64_t b_len, void* x_host, int64_t x_len,
                                         int32_t* iterations_done, rls_solver_scalars* scalars) {
  RLS_CHECK_ARG(s && b_host && x_host, "NULL argument");
  RLS_CHECK_ARG(x_len == s->n, "solve!: x buffer has %lld elements, expected %lld", (long long)x_len, (long long)s->n);
  RlsNvtxRange nvtx("rls: solve! (host buffers)");
  RlsDeviceGuard g(s->ctx->device);
  const size_t es = rls_elem_size(s->dtype);
  if (!s->b_dev || s->b_dev->len != b_len) {
    if (s->b_dev) rls_vec_destroy(s->b_dev);
    s->b_dev = nullptr;
    RLS_TRY(rls_vec_create_internal(s->ctx, s->dtype, b_len, &s->b_dev));
  }
  RLS_TRY(ensure_pinned(&s->pin_b, &s->pin_b_bytes, (size_t)b_len * es));
  RLS_TRY(ensure_pinned(&s->pin_x, &s->pin_x_bytes, (size_t)x_len * es));
  memcpy(s->pin_b, b_host, (size_t)b_len * es);
  RLS_CUDA(cudaMemcpyAsync(s->b_dev->d, s->pin_b, (size_t)b_len * es, cudaMemcpyHostToDevice, s->ctx->stream));
  RLS_TRY(rls_solver_solve(s, s->b_dev, nullptr, iterations_done, scalars));
  RLS_CUDA(cudaMemcpyAsync(s->pin_x, s->lanes[0].v[V_X]->d, (size_t)x_len * es, cudaMemcpyDeviceToHost, s->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(s->ctx->stream));
  memcpy(x_host, s->pin_x, (size_t)x_len * es);
  return RLS_OK;
}

extern "C" int32_t rls_solver_scalars_get(rls_solver_t s, rls_solver_scalars* scalars) {
  RLS_CHECK_ARG(s && scalars, "NULL argument");
  fill_scalars(s->lanes[0].hS, scalars);
  return RLS_OK;
}

extern "C" int32_t rls_solver_vec(rls_solver_t s, const char* name, rls_vec_t* out) {
  RLS_CHECK_ARG(s && name && out, "NULL argument");
  for (int k = 0; k < V_COUNT; ++k) {
    if (strcmp(name, kVecNames[k]) == 0) {
      RLS_CHECK_ARG(s->lanes[0].v[k], "solver has no state vector '%s'", name);
      *out = s->lanes[0].v[k];
      return RLS_OK;
    }
  }
  rls_set_error("unknown state vector '%s'", name);
  return RLS_ERR_INVALID;
}

// FISTA with an elementwise prox: the K momentum kernels and the K epilogues of a batched iteration are one launch
// each (fista_*_batch_kernel); everything else keeps the per-column kernels
static bool fista_batched_kernels(rls_solver_s* s, int K) {
  if (!rls_env_flag("RLS_BATCH_LANE_KERNELS", true)) return false;
  if (s->desc.kind != RLS_FISTA || !rls_reg_is_elementwise(s->desc.reg[0].kind) || K > BATCH_MAXK || K > 4096) return false;
  const int g = ew_grid(s->ctx, s->n);
  return (int64_t)K * g * 2 <= (int64_t)RLS_MAX_RED_BLOCKS * RLS_MAX_ACC;
}

template <typename T>
static int32_t fista_batch_iteration(rls_solver_s* s, int K) {
  rls_ctx_s* c = s->ctx;
  cudaStream_t st = c->stream;
  const int g = ew_grid(c, s->n);
  FistaLanes<T> tab{};
  for (int k = 0; k < K; ++k) {
    Lane& L = s->lanes[k];
    swap_roles(s, L); L.enq_swaps++;
    tab.x[k] = P<T>(L.v[V_X]); tab.res[k] = P<T>(L.v[V_RES]); tab.x0[k] = P<T>(L.v[V_X0]); tab.xold[k] = P<T>(L.v[V_XOLD]);
    tab.S[k] = L.dS;
    s->batch_x.push_back(L.v[V_X]->d); s->batch_res.push_back(L.v[V_RES]->d); s->batch_gate.push_back(&L.dS->done);
  }
  fista_momentum_batch_kernel<T><<<dim3(g, K), EB, 0, st>>>(tab, s->n);
  c->launches++;
  RLS_TRY(rls_normal_apply_batch_raw(s->AHA, K, s->batch_x.data(), s->batch_res.data(), s->batch_gate.data()));
  fista_main_batch_kernel<T><<<dim3(g, K), EB, 0, st>>>(tab, s->n, s->desc.reg[0].kind, c->red_partials, (int64_t)g * 2, c->gemv_tickets);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// multi-RHS: K lanes sharing A / AHA / reg, per-column device-side done() masks
// (MultiThreading.jl:30-80).  Iterations are interleaved lane by lane on the stream.
extern "C" int32_t rls_solver_solve_batch_host(rls_solver_t s, const void* B_host, int64_t ldb, int32_t K, void* X_host,
                                               int64_t ldx, int32_t* iterations_done) {
  RLS_CHECK_ARG(s && B_host && X_host, "NULL argument");
  RLS_CHECK_ARG(K >= 1, "K must be >= 1");
  const int64_t blen = s->A ? s->A->m : s->n;
  RLS_CHECK_ARG(ldb >= blen && ldx >= s->n, "leading dimensions too small");
  RlsNvtxRange nvtx("rls: solve! (multi-RHS)");
  RlsDeviceGuard g(s->ctx->device);
  const size_t es = rls_elem_size(s->dtype);
  // RLS_TRACE_BATCH=1: wall time of the phases (with a stream sync at every phase boundary)
  const bool tr = getenv("RLS_TRACE_BATCH") != nullptr;
  auto now = [&]() { if (tr) cudaStreamSynchronize(s->ctx->stream); return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  const size_t old = s->lanes.size();
  if ((size_t)K < old) for (size_t k = K; k < old; ++k) free_lane(s->lanes[k]);
  s->lanes.resize(K);
  for (int k = 0; k < K; ++k) RLS_TRY(alloc_lane(s, s->lanes[k]));
  rls_vec_s* Bd = nullptr;
  // device copy of B with every column on a 16-byte boundary (the kernels use 128-bit loads)
  const int64_t bstride = (blen + 3) & ~(int64_t)3;
  RLS_TRY(rls_vec_create_internal(s->ctx, s->dtype, bstride * K, &Bd));
  int32_t status = RLS_OK;
  do {
    if ((status = (cudaMemcpy2DAsync(Bd->d, bstride * es, B_host, ldb * es, blen * es, K, cudaMemcpyHostToDevice, s->ctx->stream) == cudaSuccess) ? RLS_OK : RLS_ERR_CUDA) != RLS_OK) break;
    const auto t1 = now();
    // back-projections A'b_k of all columns as one tensor-core GEMM (row-major A, K >= 8), else per column in init!
    bool have_atb = false;
    if (s->A && K > 1) {
      std::vector<const void*> bp(K);
      std::vector<void*> xp(K);
      for (int k = 0; k < K; ++k) {
        bp[k] = (const char*)Bd->d + (size_t)k * bstride * es;
        xp[k] = s->lanes[k].v[admm_like(s->desc.kind) ? V_BETAY : V_X0]->d;
      }
      status = rls_normal_adjoint_batch_raw(s->AHA, K, bp.data(), xp.data(), &have_atb);
      if (status != RLS_OK) break;
    }
    for (int k = 0; k < K && status == RLS_OK; ++k)
      status = init_lane(s, s->lanes[k], (const char*)Bd->d + (size_t)k * bstride * es, blen, nullptr, have_atb);
    if (status != RLS_OK) break;
    const auto t2 = now();
    const int cap = iteration_cap(s->desc, s->n);
    // the K applies of a batched iteration go through rls_normal_apply_batch_raw — two tensor-core GEMMs reading A once
    // each when A is row-major — between the per-column pre and post kernels: one apply per iteration for FISTA / POGM /
    // OptISTA / CGNR, 1 + n_cg (+1) applies with per-column device gates for ADMM / SplitBregman.
    const bool split = K > 1;
    for (int it = 0; it < cap && status == RLS_OK; ++it) {
      if (!split) {
        for (int k = 0; k < K && status == RLS_OK; ++k) status = enqueue_iteration(s, s->lanes[k]);
        continue;
      }
      if (admm_like(s->desc.kind)) {
        // segment by segment: the K pending applies of a segment (AHA x, or AHA u of one inner CG step, each with its
        // column's device gate) run as one batched apply
        const int nseg = admm_segment_count(s->desc);
        for (int seg = 0; seg < nseg && status == RLS_OK; ++seg) {
          s->batch_x.clear(); s->batch_res.clear(); s->batch_gate.clear();
          for (int k = 0; k < K && status == RLS_OK; ++k) {
            PendingApply pa{nullptr, nullptr, nullptr};
            status = s->dtype == RLS_C32 ? admm_segment<float2>(s, s->lanes[k], seg, &pa) : admm_segment<float>(s, s->lanes[k], seg, &pa);
            if (pa.x) { s->batch_x.push_back(pa.x); s->batch_res.push_back(pa.out); s->batch_gate.push_back(pa.gate); }
          }
          if (status == RLS_OK && !s->batch_x.empty())
            status = rls_normal_apply_batch_raw(s->AHA, K, s->batch_x.data(), s->batch_res.data(), s->batch_gate.data());
        }
        continue;
      }
      s->batch_x.clear(); s->batch_res.clear(); s->batch_gate.clear();
      if (fista_batched_kernels(s, K)) {
        status = s->dtype == RLS_C32 ? fista_batch_iteration<float2>(s, K) : fista_batch_iteration<float>(s, K);
        continue;
      }
      for (int k = 0; k < K && status == RLS_OK; ++k) status = enqueue_iteration(s, s->lanes[k], IT_PRE);
      if (status == RLS_OK) status = rls_normal_apply_batch_raw(s->AHA, K, s->batch_x.data(), s->batch_res.data(), s->batch_gate.data());
      for (int k = 0; k < K && status == RLS_OK; ++k) status = enqueue_iteration(s, s->lanes[k], IT_POST);
    }
    if (status != RLS_OK) break;
    const auto t3 = now();
    if (tr) fprintf(stderr, "[batch] K=%d: alloc + H2D %.2f ms, init (A'b per column) %.2f ms, iterations %.2f ms\n", K, ms_since(t0, t1), ms_since(t1, t2), ms_since(t2, t3));
    for (int k = 0; k < K && status == RLS_OK; ++k) {
      Lane& L = s->lanes[k];
      if (s->desc.kind == RLS_CGNR && s->desc.proj_mask)
        status = rls_proj_launch(s->ctx, s->dtype, L.v[V_X]->d, s->n, s->desc.proj_mask, nullptr);
      if (status == RLS_OK) status = pull_state(s, L);
      if (status != RLS_OK) break;
      reconcile_roles(s, L);
      if (iterations_done) iterations_done[k] = L.hS->iteration;
      if (cudaMemcpyAsync((char*)X_host + (size_t)k * ldx * es, L.v[V_X]->d, s->n * es, cudaMemcpyDeviceToHost, s->ctx->stream) != cudaSuccess) status = RLS_ERR_CUDA;
    }
    if (status == RLS_OK && cudaStreamSynchronize(s->ctx->stream) != cudaSuccess) status = RLS_ERR_CUDA;
    if (tr) fprintf(stderr, "[batch] K=%d: state download + D2H of X %.2f ms\n", K, ms_since(t3, now()));
  } while (0);
  if (status == RLS_ERR_CUDA) rls_set_error("CUDA error in batch solve: %s", cudaGetErrorString(cudaGetLastError()));
  const auto t5 = now();
  rls_vec_destroy(Bd);
  if (tr) fprintf(stderr, "[batch] K=%d: free of the device copy of B %.2f ms\n", K, ms_since(t5, now()));
  return status;
}
