// rls_kaczmarz.cu — the row loop of Kaczmarz (src/Kaczmarz.jl:264-283, row step :305-310) on a row-major system matrix.
//
// The reference visits the rows one after the other:
//     tau   = dot_with_matrix_row(A, x, row)                    (Utils.jl:59-105, unconjugated)
//     alpha = denom[i] * (u[row] - tau - eps_w * vl[row])       (Kaczmarz.jl:307)
//     x    += alpha * conj(A[row, :])                           (Kaczmarz.jl:432-436)
//     vl[row] += alpha * eps_w                                  (Kaczmarz.jl:309)
// which on a GPU is one latency-bound dot + axpy per row.  The same recurrence in block form: for R consecutive
// rows a_1..a_R (a block), with t = A_blk x taken BEFORE the block and the block Gram matrix G = A_blk A_blk^H,
//     tau_j   = t_j + sum_{k<j} alpha_k G[j,k]
//     alpha_j = denom_j * (u_j - tau_j - eps_w vl_j)
//     x      += A_blk^H alpha
// gives, in exact arithmetic, the very same iterates (G does not depend on x, so it is built once per row order).
// Per block this is a whole-GPU streaming pass (t), a tiny triangular recurrence in one CTA, and a second pass over the
// block that finds it in L2 (the block is sized for that): one HBM sweep over A per Kaczmarz iteration.
//
//   kz_dot_kernel     t partials: CTA = (column chunk, 8 rows), x chunk in registers, rows streamed with 128-bit loads
//   kz_solve_kernel   one CTA of R threads: 32x32 diagonal blocks by warp shuffles, panels after one __syncthreads
//   kz_update_kernel  x += sum_j alpha_j conj(a_j): thread = 4 floats of x, rows of the block from L2
//   kz_gram_kernel    G = A_blk A_blk^H, 64x64 tiles on the FP32 pipes (lower triangle of tiles), once per row order
//   kz_rownorm2_kernel  rownorm²(A, i) (Utils.jl:16-23) for denom / probabilities, which the host forms as the
//                       reference does (initkaczmarz Kaczmarz.jl:365-376, rowProbabilities :326-334)
#include "rls_common.cuh"

#include <algorithm>
#include <stdlib.h>

struct rls_kaczmarz_s {
  rls_ctx_s* ctx = nullptr;
  rls_mat_s* A = nullptr;
  int fpe = 1;
  int R = 128;            // rows per block
  int S = 1;              // column chunks of the dot kernel
  bool vec4 = false;      // 128-bit loads possible
  int64_t count = 0;      // rows in the current order
  int64_t nblk = 0;
  int32_t* d_rows = nullptr;   // [nblk * R], -1 = padding
  float* d_denom = nullptr;    // [nblk * R]
  float* d_G = nullptr;        // [nblk][R * R * fpe], element (k, j) at (j * R + k) * fpe
  int64_t cap_blk = 0;
  float* d_tpart = nullptr;    // [S][R][fpe]
  float* d_alpha = nullptr;    // [R][fpe]
  float* d_s2 = nullptr;       // [m]
  rls_vec_s *x = nullptr, *vl = nullptr, *u = nullptr;
  float eps_w = 0.f;
  bool initialised = false;
  // persistent sweep kernel (one cooperative launch per iteration)
  float* d_Dinv = nullptr;     // [nblk][R/32][32 jj][32 lane][fpe]: inverses of the 32x32 diagonal blocks of D^-1 + strictlower(G)
  float* d_tpart2 = nullptr;   // [NC][R][fpe] tagged pairs {value, tag}
  float* d_alpha2 = nullptr;   // [R][fpe] tagged pairs
  float* d_tsum = nullptr;     // [R][fpe] tagged pairs: t summed over the CTAs, one reducer CTA per row
  size_t tagged_bytes = 0, alpha2_bytes = 0;
  int64_t dinv_cap = 0;        // blocks d_Dinv has room for
  int* d_abort = nullptr;
  unsigned long long epoch = 0;           // blocks completed since the counters were reset
  long long* d_trace = nullptr;
  int pgrid = 0, pP = 0, pIT = 0;
  size_t psmem = 0;
  bool persistent = false;
};

namespace {

constexpr int KZ_DOT_THREADS = 256;
constexpr int KZ_DOT_LPT = 2;   // loads per thread and row
constexpr int KZ_DOT_RG = 4;    // rows per CTA, all in flight at once
constexpr int KZ_UPD_THREADS = 512;   // 32 packs x 16 row groups
constexpr int KZ_MAX_R = 256;

template <int NF> struct PackT;
template <> struct PackT<1> { using type = float; };
template <> struct PackT<2> { using type = float2; };
template <> struct PackT<4> { using type = float4; };

template <int NF>
__device__ __forceinline__ void ld_pack(const float* __restrict__ p, float (&v)[NF]) {
  using P = typename PackT<NF>::type;
  P q = __ldg(reinterpret_cast<const P*>(p));
  const float* f = reinterpret_cast<const float*>(&q);
#pragma unroll
  for (int i = 0; i < NF; ++i) v[i] = f[i];
}
// L2-coherent load (data written by an earlier kernel of the same chain: x, partials, alpha)
template <int NF>
__device__ __forceinline__ void ld_pack_cg(const float* p, float (&v)[NF]) {
  using P = typename PackT<NF>::type;
  P q = __ldcg(reinterpret_cast<const P*>(p));
  const float* f = reinterpret_cast<const float*>(&q);
#pragma unroll
  for (int i = 0; i < NF; ++i) v[i] = f[i];
}
template <int NF>
__device__ __forceinline__ void ld_pack_stream(const float* __restrict__ p, float (&v)[NF]) {
  using P = typename PackT<NF>::type;
  P q = __ldcs(reinterpret_cast<const P*>(p));
  const float* f = reinterpret_cast<const float*>(&q);
#pragma unroll
  for (int i = 0; i < NF; ++i) v[i] = f[i];
}

__device__ __forceinline__ void st_pack(float* p, const float (&v)[1]) { *p = v[0]; }
__device__ __forceinline__ void st_pack(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
__device__ __forceinline__ void st_pack(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// acc += a . x (unconjugated) over one pack
template <int FPE, int NF>
__device__ __forceinline__ void dot_acc(const float (&a)[NF], const float (&x)[NF], float (&acc)[FPE]) {
  if constexpr (FPE == 1) {
#pragma unroll
    for (int i = 0; i < NF; ++i) acc[0] = fmaf(a[i], x[i], acc[0]);
  } else {
#pragma unroll
    for (int i = 0; i < NF; i += 2) {
      acc[0] = fmaf(a[i], x[i], acc[0]);
      acc[0] = fmaf(-a[i + 1], x[i + 1], acc[0]);
      acc[1] = fmaf(a[i], x[i + 1], acc[1]);
      acc[1] = fmaf(a[i + 1], x[i], acc[1]);
    }
  }
}

// acc += al * conj(a) over one pack
template <int FPE, int NF>
__device__ __forceinline__ void upd_acc(const float (&al)[FPE], const float (&a)[NF], float (&acc)[NF]) {
  if constexpr (FPE == 1) {
#pragma unroll
    for (int i = 0; i < NF; ++i) acc[i] = fmaf(al[0], a[i], acc[i]);
  } else {
#pragma unroll
    for (int i = 0; i < NF; i += 2) {
      acc[i] = fmaf(al[0], a[i], acc[i]);
      acc[i] = fmaf(al[1], a[i + 1], acc[i]);
      acc[i + 1] = fmaf(al[1], a[i], acc[i + 1]);
      acc[i + 1] = fmaf(-al[0], a[i + 1], acc[i + 1]);
    }
  }
}

// c += al * g (complex or real product)
template <int FPE>
__device__ __forceinline__ void mul_acc(const float (&al)[FPE], const float (&g)[FPE], float (&c)[FPE]) {
  if constexpr (FPE == 1) {
    c[0] = fmaf(al[0], g[0], c[0]);
  } else {
    c[0] = fmaf(al[0], g[0], c[0]);
    c[0] = fmaf(-al[1], g[1], c[0]);
    c[1] = fmaf(al[0], g[1], c[1]);
    c[1] = fmaf(al[1], g[0], c[1]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// rownorm²(A, i): one CTA per row
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kz_rownorm2_kernel(const float* __restrict__ A, int64_t ldf, int64_t nfl, int64_t m,
                                                          float* __restrict__ s2) {
  __shared__ float s_w[8];
  for (int64_t row = blockIdx.x; row < m; row += gridDim.x) {
    const float* a = A + row * ldf;
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < nfl; i += 256) { float v = __ldg(a + i); acc = fmaf(v, v, acc); }
    acc = warp_sum(acc);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_w[w];
      s2[row] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// t = A_blk x, split over column chunks: tpart[(s * R + j) * FPE + c]
// ---------------------------------------------------------------------------------------------------------------
template <int FPE, int NF>
__global__ void __launch_bounds__(KZ_DOT_THREADS) kz_dot_kernel(const float* __restrict__ A, int64_t ldf, int64_t nfl,
                                                                const int32_t* __restrict__ rows, int R,
                                                                const float* x, float* __restrict__ tpart) {
  pdl_prologue();
  constexpr int LPT = KZ_DOT_LPT, RG = KZ_DOT_RG, NW = KZ_DOT_THREADS / 32;
  __shared__ float s_red[NW][RG * FPE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t chunk0 = (int64_t)blockIdx.x * (KZ_DOT_THREADS * LPT * NF);
  // out-of-range packs read pack 0 against x = 0, padding rows read row 0 and are zeroed afterwards: all loads are
  // unconditional, so the compiler can put the RG x LPT loads of a CTA in flight at once
  int64_t idx[LPT];
  float xr[LPT][NF];
#pragma unroll
  for (int l = 0; l < LPT; ++l) {
    const int64_t i0 = chunk0 + ((int64_t)l * KZ_DOT_THREADS + tid) * NF;
    idx[l] = i0 < nfl ? i0 : 0;
    ld_pack_cg<NF>(x + idx[l], xr[l]);
    if (i0 >= nfl) {
#pragma unroll
      for (int i = 0; i < NF; ++i) xr[l][i] = 0.f;
    }
  }
  const int j0 = blockIdx.y * RG;
  float acc[RG][FPE];
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int c = 0; c < FPE; ++c) acc[r][c] = 0.f;
  {
    float a[RG][LPT][NF];
    int rowv[RG];
#pragma unroll
    for (int q = 0; q < RG; ++q) {
      const int j = j0 + q;
      rowv[q] = j < R ? rows[j] : -1;
    }
#pragma unroll
    for (int q = 0; q < RG; ++q) {
      const float* ap = A + (int64_t)(rowv[q] < 0 ? 0 : rowv[q]) * ldf;
#pragma unroll
      for (int l = 0; l < LPT; ++l) ld_pack<NF>(ap + idx[l], a[q][l]);   // stays in L2 for the update pass
    }
#pragma unroll
    for (int q = 0; q < RG; ++q) {
#pragma unroll
      for (int l = 0; l < LPT; ++l) dot_acc<FPE, NF>(a[q][l], xr[l], acc[q]);
      if (rowv[q] < 0) {
#pragma unroll
        for (int c = 0; c < FPE; ++c) acc[q][c] = 0.f;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int c = 0; c < FPE; ++c) {
      float w = warp_sum(acc[r][c]);
      if (lane == 0) s_red[warp][r * FPE + c] = w;
    }
  __syncthreads();
  if (tid < RG * FPE) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_red[w][tid];
    const int r = tid / FPE, c = tid % FPE;
    const int j = j0 + r;
    if (j < R) tpart[((int64_t)blockIdx.x * R + j) * FPE + c] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the triangular recurrence of one block; blockDim.x == R (multiple of 32, <= 256).
// STAGE: the block's Gram matrix (immutable) is copied to shared memory BEFORE griddepcontrol.wait, i.e. while the
// dot kernel is still running, so that the serial part only touches shared memory and registers.
// ---------------------------------------------------------------------------------------------------------------
template <int FPE, bool STAGE>
__global__ void __launch_bounds__(KZ_MAX_R) kz_solve_kernel(const int32_t* __restrict__ rows, const float* __restrict__ denom,
                                                            const float* __restrict__ G, int R,
                                                            const float* tpart, int S,
                                                            const float* u, float* vl, float ew,
                                                            float* __restrict__ alpha_out) {
  extern __shared__ float4 s_dyn4[];
  __shared__ float s_alpha[KZ_MAX_R * FPE];
  float* sG = reinterpret_cast<float*>(s_dyn4);
  const int k = threadIdx.x, lane = k & 31, warp = k >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (STAGE) {
    const float4* g4 = reinterpret_cast<const float4*>(G);
    const int n4 = R * R * FPE / 4;
#pragma unroll 8
    for (int i = k; i < n4; i += R) s_dyn4[i] = __ldg(g4 + i);
  }
  const int row = rows[k];
  const float d = row >= 0 ? denom[k] : 0.f;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float t[FPE], uu[FPE], vv[FPE], c[FPE], mine[FPE];
#pragma unroll
  for (int q = 0; q < FPE; ++q) { t[q] = 0.f; c[q] = 0.f; mine[q] = 0.f; uu[q] = 0.f; vv[q] = 0.f; }
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int q = 0; q < FPE; ++q) t[q] += __ldcg(tpart + ((int64_t)s * R + k) * FPE + q);
  if (row >= 0) {
#pragma unroll
    for (int q = 0; q < FPE; ++q) { uu[q] = __ldcg(u + (int64_t)row * FPE + q); vv[q] = __ldcg(vl + (int64_t)row * FPE + q); }
  }
  if (STAGE) __syncthreads();
  auto ldG = [&](int j, int q) -> float {
    const int64_t o = ((int64_t)j * R + k) * FPE + q;
    return STAGE ? sG[o] : __ldg(G + o);
  };
  const int nb = R >> 5;
  for (int jb = 0; jb < nb; ++jb) {
    const int j0 = jb << 5;
    if (warp == jb) {
      float g[32][FPE];
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
#pragma unroll
        for (int q = 0; q < FPE; ++q) g[jj][q] = ldG(j0 + jj, q);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        float al[FPE];
#pragma unroll
        for (int q = 0; q < FPE; ++q) {
          // alpha = denom * ((u - tau) - eps_w * vl), tau = t + c   (Kaczmarz.jl:307)
          float tau = fadd(t[q], c[q]);
          float a_ = fmul(d, fsub(fsub(uu[q], tau), fmul(ew, vv[q])));
          al[q] = __shfl_sync(0xffffffffu, a_, jj);
        }
        if (lane > jj) mul_acc<FPE>(al, g[jj], c);
        if (lane == jj) {
#pragma unroll
          for (int q = 0; q < FPE; ++q) mine[q] = al[q];
        }
      }
#pragma unroll
      for (int q = 0; q < FPE; ++q) {
        s_alpha[k * FPE + q] = mine[q];
        alpha_out[k * FPE + q] = mine[q];
        if (row >= 0) vl[(int64_t)row * FPE + q] = fadd(vv[q], fmul(mine[q], ew));   // Kaczmarz.jl:309
      }
    }
    __syncthreads();
    if (k >= j0 + 32) {
#pragma unroll 8
      for (int jj = 0; jj < 32; ++jj) {
        float al[FPE], g[FPE];
#pragma unroll
        for (int q = 0; q < FPE; ++q) {
          al[q] = s_alpha[(j0 + jj) * FPE + q];
          g[q] = ldG(j0 + jj, q);
        }
        mul_acc<FPE>(al, g, c);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// x += sum_j alpha_j conj(a_j).  CTA = 32 packs of x (one warp-wide, contiguous) x 16 row groups (one per warp): group g
// accumulates rows [g R/16, (g+1) R/16) of the block, the 16 partial sums meet in shared memory and warp 0 adds them in
// a fixed order.
// ---------------------------------------------------------------------------------------------------------------
template <int FPE, int NF>
__global__ void __launch_bounds__(KZ_UPD_THREADS) kz_update_kernel(const float* __restrict__ A, int64_t ldf, int64_t nfl,
                                                                   const int32_t* __restrict__ rows, int R,
                                                                   const float* alpha, float* x) {
  constexpr int NG = KZ_UPD_THREADS / 32;
  __shared__ float s_al[KZ_MAX_R * FPE];
  __shared__ int s_row[KZ_MAX_R];
  __shared__ float s_part[NG][32][NF];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = threadIdx.x; i < R; i += KZ_UPD_THREADS) s_row[i] = rows[i];   // immutable: before the dependency wait
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int i = threadIdx.x; i < R * FPE; i += KZ_UPD_THREADS) s_al[i] = __ldcg(alpha + i);
  __syncthreads();
  const int64_t idx0 = ((int64_t)blockIdx.x * 32 + lane) * NF;
  const bool live = idx0 < nfl;
  const int64_t idx = live ? idx0 : 0;   // dead lanes read pack 0 and drop the result: loads stay unconditional
  float acc[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = 0.f;
  const int rpg = R / NG;   // R is a multiple of 64, NG = 16
  constexpr int U = 4;
  for (int j0 = grp * rpg; j0 < (grp + 1) * rpg; j0 += U) {
    float a[U][NF];
#pragma unroll
    for (int r = 0; r < U; ++r) {
      const int row = s_row[j0 + r];     // padding rows (-1) have alpha = 0: read row 0
      ld_pack_stream<NF>(A + (int64_t)(row < 0 ? 0 : row) * ldf + idx, a[r]);   // last use of the block
    }
#pragma unroll
    for (int r = 0; r < U; ++r) {
      float al[FPE];
#pragma unroll
      for (int q = 0; q < FPE; ++q) al[q] = s_al[(j0 + r) * FPE + q];
      upd_acc<FPE, NF>(al, a[r], acc);
    }
  }
#pragma unroll
  for (int i = 0; i < NF; ++i) s_part[grp][lane][i] = acc[i];
  __syncthreads();
  if (grp == 0 && live) {
    float xv[NF];
    ld_pack_cg<NF>(x + idx, xv);
    float tot[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) tot[i] = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int i = 0; i < NF; ++i) tot[i] += s_part[g][lane][i];
#pragma unroll
    for (int i = 0; i < NF; ++i) xv[i] += tot[i];
    st_pack(x + idx, xv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// G = A_blk A_blk^H per block: 64x64 tiles (lower triangle of tiles), 256 threads, 4x4 outputs per thread
// ---------------------------------------------------------------------------------------------------------------
template <int FPE>
__global__ void __launch_bounds__(256) kz_gram_kernel(const float* __restrict__ A, int64_t ldf, int64_t n,
                                                      const int32_t* __restrict__ rows_all, int R, float* __restrict__ G_all,
                                                      int64_t blk0) {
  constexpr int TS = 64, KC = 16, LD = TS * FPE + 4;
  __shared__ float As[KC][LD];
  __shared__ float Bs[KC][LD];
  const int64_t b = blk0 + blockIdx.y;
  const int32_t* rows = rows_all + b * R;
  float* G = G_all + b * (int64_t)R * R * FPE;
  int t = blockIdx.x, ti = 0;
  while (t >= ti + 1) { t -= ti + 1; ++ti; }
  const int tj = t;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int rowA = rows[ti * TS + lr], rowB = rows[tj * TS + lr];
  const float* pa = A + (int64_t)(rowA < 0 ? 0 : rowA) * ldf;
  const float* pb = A + (int64_t)(rowB < 0 ? 0 : rowB) * ldf;
  float acc[4][4][FPE];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < FPE; ++q) acc[i][j][q] = 0.f;
  for (int64_t k0 = 0; k0 < n; k0 += KC) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const int64_t kk = k0 + lk + q4;
#pragma unroll
      for (int q = 0; q < FPE; ++q) {
        As[lk + q4][lr * FPE + q] = (rowA >= 0 && kk < n) ? __ldg(pa + kk * FPE + q) : 0.f;
        Bs[lk + q4][lr * FPE + q] = (rowB >= 0 && kk < n) ? __ldg(pb + kk * FPE + q) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float a[4][FPE], bb[4][FPE];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < FPE; ++q) {
          a[i][q] = As[kk][(ty * 4 + i) * FPE + q];
          bb[i][q] = Bs[kk][(tx * 4 + i) * FPE + q];
        }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (FPE == 1) {
            acc[i][j][0] = fmaf(a[i][0], bb[j][0], acc[i][j][0]);
          } else {   // a * conj(b)
            acc[i][j][0] = fmaf(a[i][0], bb[j][0], acc[i][j][0]);
            acc[i][j][0] = fmaf(a[i][1], bb[j][1], acc[i][j][0]);
            acc[i][j][1] = fmaf(a[i][1], bb[j][0], acc[i][j][1]);
            acc[i][j][1] = fmaf(-a[i][0], bb[j][1], acc[i][j][1]);
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kr = ti * TS + ty * 4 + i, jc = tj * TS + tx * 4 + j;
#pragma unroll
      for (int q = 0; q < FPE; ++q) G[((int64_t)jc * R + kr) * FPE + q] = acc[i][j][q];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Inverses of the 32x32 diagonal blocks of M = D^-1 + strictlower(G) (the matrix of the block recurrence
// M alpha = r): with them a diagonal block of the forward substitution is 32 independent shuffle+FMA per lane
// instead of a 32-step dependent chain.  One warp per diagonal block, inversion in double, lane = column.
// Padding rows (denom 0, alpha must stay 0) get a zero row and a zero column.
// ---------------------------------------------------------------------------------------------------------------
template <int FPE>
__global__ void __launch_bounds__(32) kz_dinv_kernel(const float* __restrict__ G_all, const float* __restrict__ denom_all, int R,
                                                     float* __restrict__ Dinv_all) {
  const int64_t b = blockIdx.y;
  const int jb = blockIdx.x, lane = threadIdx.x, j0 = jb * 32;
  const float* G = G_all + b * (int64_t)R * R * FPE;
  const float* den = denom_all + b * R;
  float* out = Dinv_all + ((b * (R / 32) + jb) * 32 * 32) * FPE;
  __shared__ double sM[32][32][FPE];   // M[i][k], k <= i
  for (int i = 0; i < 32; ++i) {
    const int k = lane;
    double v[FPE];
#pragma unroll
    for (int q = 0; q < FPE; ++q) v[q] = 0.0;
    if (k < i) {
#pragma unroll
      for (int q = 0; q < FPE; ++q) v[q] = (double)G[((int64_t)(j0 + k) * R + (j0 + i)) * FPE + q];   // G[row i, col k]
    } else if (k == i) {
      const float d = den[j0 + i];
      v[0] = d > 0.f ? 1.0 / (double)d : 1.0;
    }
#pragma unroll
    for (int q = 0; q < FPE; ++q) sM[i][k][q] = v[q];
  }
  __syncwarp();
  // column `lane` of X = M^-1 by forward substitution
  double X[32][FPE];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    double acc[FPE];
#pragma unroll
    for (int q = 0; q < FPE; ++q) acc[q] = 0.0;
    acc[0] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) {
      if constexpr (FPE == 1) acc[0] -= sM[i][k][0] * X[k][0];
      else {
        acc[0] -= sM[i][k][0] * X[k][0] - sM[i][k][1] * X[k][1];
        acc[1] -= sM[i][k][0] * X[k][1] + sM[i][k][1] * X[k][0];
      }
    }
    const double dii = sM[i][i][0];   // real diagonal
#pragma unroll
    for (int q = 0; q < FPE; ++q) X[i][q] = acc[q] / dii;
  }
  const bool dead_col = !(den[j0 + lane] > 0.f);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const bool dead = dead_col || !(den[j0 + i] > 0.f);
    // X[i][lane] = element (row i, column lane); stored so that lane = ROW reads coalesced: [jj = column][row]
#pragma unroll
    for (int q = 0; q < FPE; ++q) out[((int64_t)lane * 32 + i) * FPE + q] = dead ? 0.f : (float)X[i][q];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One Kaczmarz iteration in ONE cooperative launch.  CTA c owns the columns [c P, (c+1) P) (P packs of 4 floats) of x
// for the whole sweep, in shared memory; a block of R rows needs only ONE grid-wide exchange:
//     every CTA:  partial t_j over its columns                      -> tagged partials
//     CTA j:      polls the 148 partials of row j, sums them        -> tagged t_j          (one reducer CTA per row)
//     CTA 0:      polls t, runs the block recurrence                -> tagged alpha
//     every CTA:  polls alpha, x_slice += sum_j alpha_j conj(a_j[slice])              (rows again, from L2)
// and no second barrier, because nobody else touches a CTA's columns.  While waiting for alpha every CTA prefetches
// its part of the next block into L2 (cp.async.bulk.prefetch), so the HBM stream overlaps the serial recurrence.
// Every wait is bounded; a time-out raises the abort flag and all CTAs leave.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KZ_PT = 512;            // threads of the persistent kernel
constexpr int KZ_PW = KZ_PT / 32;
constexpr int KZ_PMAX_R = 128;
constexpr unsigned KZ_SPIN_LIMIT = 1u << 20;   // polls of ~1 us each before a wait gives up

struct KzSweep {
  const float* A; int64_t ldf; int64_t npacks;
  const int32_t* rows; const float* denom; const float* G; const float* Dinv;
  int R; int nblk; int P;
  const float* u; float* vl; float ew; float* x;
  float* tpart; float* tsum; float* alpha;
  int* abort_flag;
  unsigned long long epoch;
  long long* trace; int trace_block;   // RLS_KACZMARZ_TRACE: clock64 stamps of CTA 0 and CTA 5 in one block
};

// Both exchanges carry their flag IN the data: every float travels as an 8-byte {value, tag} pair written with one
// store (tag = number of the block since init), and the reader polls the pair itself.  No fence, no counter: a value is
// usable the moment its own tag matches.
__device__ __forceinline__ void st_tagged(float* slot, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" :: "l"(slot), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_tagged(const float* slot) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(slot) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_tagged2(const float* slot) {   // two consecutive pairs
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slot) : "memory");
  return v;
}

template <int FPE, int IT>
__global__ void __launch_bounds__(KZ_PT, 1) kz_sweep_kernel(const KzSweep p) {
  constexpr int HB = 16 / IT;   // rows per batch: HB x IT 128-bit loads in flight per thread
  extern __shared__ float4 kz_smem4[];
  __shared__ float s_alpha[KZ_PMAX_R * FPE];
  __shared__ int s_rows[KZ_PMAX_R];
  __shared__ __align__(16) float s_t[KZ_PMAX_R * FPE];
  __shared__ float s_red[KZ_PW][FPE];
  const int R = p.R, P = p.P;
  float* sG = reinterpret_cast<float*>(kz_smem4);           // [R*R*FPE]   (CTA 0)
  float4* xs4 = kz_smem4 + (size_t)R * R * FPE / 4;          // [P]
  float4* red4 = xs4 + P;                                    // [KZ_PW][P]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, NC = gridDim.x;
  const int64_t pk0 = (int64_t)cta * P;
  const int np = (int)(p.npacks - pk0 < (int64_t)P ? (p.npacks - pk0 > 0 ? p.npacks - pk0 : 0) : P);
  const float4* A4 = reinterpret_cast<const float4*>(p.A) + (np > 0 ? pk0 : 0);   // this CTA's first column pack
  const int64_t ld4 = p.ldf / 4;
  float4* x4 = reinterpret_cast<float4*>(p.x);
  for (int i = tid; i < np; i += KZ_PT) xs4[i] = __ldcg(x4 + pk0 + i);
  bool alive = true;
#define KZ_STAMP(i) do { if (p.trace && b == p.trace_block && tid == 0 && (cta == 0 || cta == 5)) p.trace[(cta ? 16 : 0) + (i)] = clock64(); } while (0)
  for (int b = 0; b < p.nblk && alive; ++b) {
    const int32_t* rows_b = p.rows + (int64_t)b * R;
    const unsigned tag = (unsigned)(p.epoch + (unsigned long long)b + 1ull);
    KZ_STAMP(0);
    __syncthreads();                       // s_rows / s_alpha / red4 of the previous block are free; xs4 is complete
    for (int i = tid; i < R; i += KZ_PT) s_rows[i] = rows_b[i];
    if (cta == 0) {                        // the block's Gram matrix -> shared memory, asynchronously
      // (its 32x32 diagonal blocks are replaced by their inverted counterparts from Dinv: same [column][row] layout)
      const float* gsrc = p.G + (int64_t)b * R * R * FPE;
      const float* dsrc = p.Dinv + (int64_t)b * R * 32 * FPE;
      const int n4 = R * R * FPE / 4;
      for (int i = tid; i < n4; i += KZ_PT) {
        const int f = i * 4, jc = f / (R * FPE), kr = (f - jc * R * FPE) / FPE;
        const float* src = (jc >> 5) == (kr >> 5) ? dsrc + ((int64_t)jc * 32 + (kr & 31)) * FPE : gsrc + f;
        unsigned dst = (unsigned)__cvta_generic_to_shared(kz_smem4 + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __syncthreads();
    KZ_STAMP(1);
    // ---- partial t_j over this CTA's columns: warp w takes rows w, w+16, ... HB at a time
    for (int j = warp; j < R; j += HB * KZ_PW) {
      float4 a[HB][IT];
      int rowv[HB];
#pragma unroll
      for (int h = 0; h < HB; ++h) {
        rowv[h] = s_rows[j + h * KZ_PW];
        const float4* ap = A4 + (int64_t)(rowv[h] < 0 ? 0 : rowv[h]) * ld4;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const int pk = it * 32 + lane;
          a[h][it] = __ldg(ap + (pk < np ? pk : 0));
        }
      }
      __syncwarp();                        // keeps the HB x IT loads ahead of their consumers
#pragma unroll
      for (int h = 0; h < HB; ++h) {
        float acc[FPE];
#pragma unroll
        for (int q = 0; q < FPE; ++q) acc[q] = 0.f;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const int pk = it * 32 + lane;
          if (pk < np) {
            const float4 xv = xs4[pk];
            const float av[4] = {a[h][it].x, a[h][it].y, a[h][it].z, a[h][it].w};
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
            dot_acc<FPE, 4>(av, xa, acc);
          }
        }
#pragma unroll
        for (int q = 0; q < FPE; ++q) acc[q] = warp_sum(acc[q]);
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < FPE; ++q)
            st_tagged(p.tpart + 2 * (((int64_t)cta * R + j + h * KZ_PW) * FPE + q), rowv[h] < 0 ? 0.f : acc[q], tag);
        }
      }
    }
    KZ_STAMP(2);
    // ---- the next block's rows (this CTA's columns) -> L2 while the recurrence runs
    if (b + 1 < p.nblk && np > 0) {
      const int32_t* rows_n = p.rows + (int64_t)(b + 1) * R;
      for (int j = tid; j < R; j += KZ_PT) {
        const int row = rows_n[j];
        if (row >= 0) {
          const float4* src = A4 + (int64_t)row * ld4;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(np * 16) : "memory");
        }
      }
    }
    KZ_STAMP(4);
    // ---- reducers: CTA c sums row c (c + NC, ...) of the partials over all CTAs — thread i polls CTA i's tagged value,
    //      a shuffle tree and the warps in order make the sum — and publishes t_row, again tagged
    for (int rr = cta; rr < R && alive; rr += NC) {
      float v[FPE];
#pragma unroll
      for (int q = 0; q < FPE; ++q) v[q] = 0.f;
      int bad = 0;
      if (tid < NC) {
        const float* slot = p.tpart + 2 * (((int64_t)tid * R + rr) * FPE);
        unsigned spins = 0;
        for (;;) {
          bool ready;
          if constexpr (FPE == 1) { const uint2 e = ld_tagged(slot); ready = e.y == tag; v[0] = __uint_as_float(e.x); }
          else { const uint4 e = ld_tagged2(slot); ready = e.y == tag && e.w == tag; v[0] = __uint_as_float(e.x); v[FPE - 1] = __uint_as_float(e.z); }
          if (ready) break;
          if (++spins > KZ_SPIN_LIMIT || ((spins & 255u) == 0u && *((volatile int*)p.abort_flag))) { *p.abort_flag = 1; bad = 1; break; }
        }
      }
#pragma unroll
      for (int q = 0; q < FPE; ++q) {
        v[q] = warp_sum(v[q]);
        if (lane == 0) s_red[warp][q] = v[q];
      }
      alive = !__syncthreads_or(bad);
      if (alive && tid < FPE) {
        float tot = 0.f;
        for (int w = 0; w < (NC + 31) / 32; ++w) tot += s_red[w][tid];
        st_tagged(p.tsum + 2 * (rr * FPE + tid), tot, tag);
      }
      __syncthreads();
    }
    if (!alive) break;
    // ---- CTA 0: the block recurrence
    if (cta == 0) {
      const int k = tid;
      const bool mine_row = k < R;
      const int row = mine_row ? s_rows[k] : -1;
      float uu[FPE], vv[FPE];
#pragma unroll
      for (int q = 0; q < FPE; ++q) { uu[q] = 0.f; vv[q] = 0.f; }
      if (mine_row) {
        if (row >= 0) {
#pragma unroll
          for (int q = 0; q < FPE; ++q) { uu[q] = __ldcg(p.u + (int64_t)row * FPE + q); vv[q] = __ldcg(p.vl + (int64_t)row * FPE + q); }
        }
      }
      {
        int bad = 0;
        if (tid < R * FPE) {             // one tagged t value per thread, published by the row's reducer CTA
          unsigned spins = 0;
          uint2 e = ld_tagged(p.tsum + 2 * tid);
          while (e.y != tag) {
            if (++spins > KZ_SPIN_LIMIT || ((spins & 255u) == 0u && *((volatile int*)p.abort_flag))) { *p.abort_flag = 1; bad = 1; break; }
            e = ld_tagged(p.tsum + 2 * tid);
          }
          s_t[tid] = __uint_as_float(e.x);
        }
        alive = !__syncthreads_or(bad);
      }
      KZ_STAMP(5);
      if (alive) {
        KZ_STAMP(6);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        KZ_STAMP(7);
        float r[FPE], c[FPE];
#pragma unroll
        for (int q = 0; q < FPE; ++q) { c[q] = 0.f; r[q] = 0.f; }
        if (mine_row) {
#pragma unroll
          for (int q = 0; q < FPE; ++q) r[q] = fsub(fsub(uu[q], s_t[k * FPE + q]), fmul(p.ew, vv[q]));   // u - t - eps_w vl
        }
        const int nb = R >> 5;
        for (int jb = 0; jb < nb; ++jb) {
          const int j0 = jb << 5;
          if (mine_row && warp == jb) {
            float rr[FPE], al[FPE];
#pragma unroll
            for (int q = 0; q < FPE; ++q) { rr[q] = fsub(r[q], c[q]); al[q] = 0.f; }
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              float rj[FPE];
#pragma unroll
              for (int q = 0; q < FPE; ++q) rj[q] = __shfl_sync(0xffffffffu, rr[q], jj);
              float dv[FPE];                     // Minv[k][jj], staged in place of G's diagonal block
#pragma unroll
              for (int q = 0; q < FPE; ++q) dv[q] = sG[((int64_t)(j0 + jj) * R + k) * FPE + q];
              mul_acc<FPE>(dv, rj, al);          // alpha_k += Minv[k][jj] r_jj   (zero above the diagonal)
            }
#pragma unroll
            for (int q = 0; q < FPE; ++q) {
              s_alpha[k * FPE + q] = al[q];
              st_tagged(p.alpha + 2 * (k * FPE + q), al[q], tag);
              if (row >= 0) p.vl[(int64_t)row * FPE + q] = fadd(vv[q], fmul(al[q], p.ew));   // Kaczmarz.jl:309
            }
          }
          __syncthreads();
          if (mine_row && k >= j0 + 32) {
#pragma unroll 8
            for (int jj = 0; jj < 32; ++jj) {
              float al[FPE], g[FPE];
#pragma unroll
              for (int q = 0; q < FPE; ++q) {
                al[q] = s_alpha[(j0 + jj) * FPE + q];
                g[q] = sG[((int64_t)(j0 + jj) * R + k) * FPE + q];
              }
              mul_acc<FPE>(al, g, c);
            }
          }
        }
        KZ_STAMP(8);
      }
    } else {
      int bad = 0;
      if (tid < R * FPE) {               // R * FPE <= 256 threads poll one tagged alpha each
        unsigned spins = 0;
        uint2 e = ld_tagged(p.alpha + 2 * tid);
        while (e.y != tag) {
          if (++spins > KZ_SPIN_LIMIT || ((spins & 255u) == 0u && *((volatile int*)p.abort_flag))) { *p.abort_flag = 1; bad = 1; break; }
          if (spins > 64) __nanosleep(32);
          e = ld_tagged(p.alpha + 2 * tid);
        }
        s_alpha[tid] = __uint_as_float(e.x);
      }
      alive = !__syncthreads_or(bad);
    }
    if (!alive) break;
    __syncthreads();
    KZ_STAMP(10);
    // ---- x_slice += sum_j alpha_j conj(a_j[slice]); rows again (L2), warp w takes rows w, w+16, ...
    float4 acc4[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) acc4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = warp; j < R; j += HB * KZ_PW) {
      float4 a[HB][IT];
      int rowv[HB];
#pragma unroll
      for (int h = 0; h < HB; ++h) {
        rowv[h] = s_rows[j + h * KZ_PW];
        const float4* ap = A4 + (int64_t)(rowv[h] < 0 ? 0 : rowv[h]) * ld4;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const int pk = it * 32 + lane;
          a[h][it] = __ldcs(ap + (pk < np ? pk : 0));     // last use of the block
        }
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < HB; ++h) {
        float al[FPE];
#pragma unroll
        for (int q = 0; q < FPE; ++q) al[q] = s_alpha[(j + h * KZ_PW) * FPE + q];   // 0 for padding rows
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const float av[4] = {a[h][it].x, a[h][it].y, a[h][it].z, a[h][it].w};
          float ac[4] = {acc4[it].x, acc4[it].y, acc4[it].z, acc4[it].w};
          upd_acc<FPE, 4>(al, av, ac);
          acc4[it] = make_float4(ac[0], ac[1], ac[2], ac[3]);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int pk = it * 32 + lane;
      if (pk < np) red4[(size_t)warp * P + pk] = acc4[it];
    }
    __syncthreads();
    KZ_STAMP(11);
    for (int i = tid; i < np; i += KZ_PT) {
      float4 v = xs4[i];
#pragma unroll
      for (int w = 0; w < KZ_PW; ++w) {
        const float4 t = red4[(size_t)w * P + i];
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
      xs4[i] = v;
    }
    KZ_STAMP(12);
  }
#undef KZ_STAMP
  __syncthreads();
  if (alive)
    for (int i = tid; i < np; i += KZ_PT) x4[pk0 + i] = xs4[i];
}

void kz_free(rls_kaczmarz_s* K) {
  if (!K) return;
  cudaFree(K->d_rows); cudaFree(K->d_denom); cudaFree(K->d_G); cudaFree(K->d_tpart); cudaFree(K->d_alpha); cudaFree(K->d_s2);
  cudaFree(K->d_Dinv); cudaFree(K->d_tpart2); cudaFree(K->d_alpha2); cudaFree(K->d_tsum); cudaFree(K->d_abort); cudaFree(K->d_trace);
  if (K->x) rls_vec_destroy(K->x);
  if (K->vl) rls_vec_destroy(K->vl);
  if (K->u) rls_vec_destroy(K->u);
  rls_ctx_s* c = K->ctx;
  rls_mat_s* A = K->A;
  delete K;
  rls_mat_release(A);   // taken in rls_kaczmarz_create: handles are freed by garbage collectors in arbitrary order
  rls_ctx_release(c);
}

}  // namespace

// persistent-kernel plan for the current row order: inverted diagonal blocks, exchange buffers, launch geometry
static int32_t kz_plan_persistent(rls_kaczmarz_s* K) {
  rls_ctx_s* c = K->ctx;
  const int R = K->R, fpe = K->fpe;
  K->persistent = false;
  if (!rls_env_flag("RLS_KACZMARZ_PERSISTENT", true)) return RLS_OK;
  if (!K->vec4 || !(R == 64 || R == 128) || K->nblk == 0 || K->nblk > 0x7fffffff) return RLS_OK;
  const int64_t npacks = K->A->n * fpe / 4;
  int NC = (int)std::min<int64_t>(c->sm_count, npacks);
  const int P = (int)((npacks + NC - 1) / NC);
  if (P > 256) return RLS_OK;
  NC = (int)((npacks + P - 1) / P);
  const int IT = P <= 128 ? 4 : 8;
  const size_t smem = ((size_t)R * R * fpe + (size_t)(1 + KZ_PW) * P * 4) * 4;
  const void* fn = fpe == 2 ? (IT == 4 ? (const void*)kz_sweep_kernel<2, 4> : (const void*)kz_sweep_kernel<2, 8>)
                            : (IT == 4 ? (const void*)kz_sweep_kernel<1, 4> : (const void*)kz_sweep_kernel<1, 8>);
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return RLS_OK; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, KZ_PT, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return RLS_OK; }
  if (NC > per_sm * c->sm_count) return RLS_OK;
  // the exchange buffers depend only on the geometry (allocated once); the inverted diagonal blocks grow with the order
  if (!K->d_tpart2) {
    K->tagged_bytes = (size_t)NC * R * fpe * 8;
    K->alpha2_bytes = (size_t)R * fpe * 8;
    if (cudaMalloc(&K->d_tpart2, K->tagged_bytes) != cudaSuccess || cudaMalloc(&K->d_alpha2, K->alpha2_bytes) != cudaSuccess ||
        cudaMalloc(&K->d_tsum, K->alpha2_bytes) != cudaSuccess || cudaMalloc(&K->d_abort, sizeof(int)) != cudaSuccess) {
      rls_set_error("Kaczmarz: cudaMalloc of the sweep-kernel buffers failed: %s", cudaGetErrorString(cudaGetLastError()));
      return RLS_ERR_NOMEM;
    }
    // tags count blocks from 1: zeroed buffers never match; later orders keep counting (tags only ever grow)
    RLS_CUDA(cudaMemsetAsync(K->d_tpart2, 0, K->tagged_bytes, c->stream));
    RLS_CUDA(cudaMemsetAsync(K->d_alpha2, 0, K->alpha2_bytes, c->stream));
    RLS_CUDA(cudaMemsetAsync(K->d_tsum, 0, K->alpha2_bytes, c->stream));
    RLS_CUDA(cudaMemsetAsync(K->d_abort, 0, sizeof(int), c->stream));
    K->epoch = 0;
  }
  if (K->nblk > K->dinv_cap) {
    cudaFree(K->d_Dinv);
    K->d_Dinv = nullptr; K->dinv_cap = 0;
    if (cudaMalloc(&K->d_Dinv, (size_t)K->nblk * R * 32 * fpe * 4) != cudaSuccess) {
      rls_set_error("Kaczmarz: cudaMalloc of the inverted diagonal blocks failed: %s", cudaGetErrorString(cudaGetLastError()));
      return RLS_ERR_NOMEM;
    }
    K->dinv_cap = K->nblk;
  }
  for (int64_t b0 = 0; b0 < K->nblk; b0 += 32768) {
    const int64_t nb = std::min<int64_t>(32768, K->nblk - b0);
    dim3 grid((unsigned)(R / 32), (unsigned)nb);
    const float* Gp = K->d_G + b0 * (int64_t)R * R * fpe;
    const float* dp = K->d_denom + b0 * R;
    float* op = K->d_Dinv + b0 * (int64_t)R * 32 * fpe;
    if (fpe == 2) kz_dinv_kernel<2><<<grid, 32, 0, c->stream>>>(Gp, dp, R, op);
    else kz_dinv_kernel<1><<<grid, 32, 0, c->stream>>>(Gp, dp, R, op);
    c->launches++;
  }
  RLS_CUDA(cudaGetLastError());
  K->pgrid = NC; K->pP = P; K->pIT = IT; K->psmem = smem;
  K->persistent = true;
  return RLS_OK;
}

extern "C" int32_t rls_kaczmarz_create(rls_mat_t A, int32_t block_rows, rls_kaczmarz_t* out) {
  RLS_CHECK_ARG(A && out, "NULL argument");
  if (A->layout != RLS_LAYOUT_ROWMAJOR) {
    rls_set_error("Kaczmarz needs the row-major device layout (create the matrix with RLS_LAYOUT_ROWMAJOR)");
    return RLS_ERR_UNSUPPORTED;
  }
  RLS_CHECK_ARG(A->m > 0 && A->n > 0, "Kaczmarz: empty matrix");
  RLS_CHECK_ARG(A->m < (int64_t)1 << 31, "Kaczmarz: more than 2^31 rows");
  rls_ctx_s* c = A->ctx;
  RlsDeviceGuard g(c->device);
  const int fpe = A->dtype == RLS_C32 ? 2 : 1;
  int R = block_rows;
  if (const char* e = getenv("RLS_KACZMARZ_BLOCK")) { if (R <= 0 && atoi(e) > 0) R = atoi(e); }
  if (R <= 0) {
    // 128 rows per block: the per-block exchange (~7 us) amortises over more rows, the Gram tile (64 / 128 KB) still fits
    // the solver CTA's shared memory, and two blocks (current + prefetched) stay L2-resident up to 512 KB rows
    R = A->m > 64 ? 128 : 64;
  }
  RLS_CHECK_ARG(R == 64 || R == 128 || R == 192 || R == 256, "Kaczmarz: block_rows must be 64, 128, 192 or 256 (got %d)", R);
  rls_kaczmarz_s* K = new rls_kaczmarz_s();
  K->ctx = c; K->A = A; K->fpe = fpe; K->R = R;
  rls_ctx_retain(c);
  rls_mat_retain(A);
  const int64_t nfl = A->n * fpe, ldf = A->ld * fpe;
  K->vec4 = (nfl % 4 == 0) && (ldf % 4 == 0) && (((uintptr_t)A->d) % 16 == 0);
  const int nf = K->vec4 ? 4 : fpe;
  K->S = (int)((nfl + (int64_t)KZ_DOT_THREADS * KZ_DOT_LPT * nf - 1) / ((int64_t)KZ_DOT_THREADS * KZ_DOT_LPT * nf));
  int32_t st = RLS_OK;
  do {
    if (cudaMalloc(&K->d_tpart, (size_t)K->S * R * fpe * 4) != cudaSuccess || cudaMalloc(&K->d_alpha, (size_t)R * fpe * 4) != cudaSuccess ||
        cudaMalloc(&K->d_s2, (size_t)A->m * 4) != cudaSuccess) {
      rls_set_error("Kaczmarz: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
      st = RLS_ERR_NOMEM; break;
    }
    if ((st = rls_vec_create_internal(c, A->dtype, A->n, &K->x)) != RLS_OK) break;
    if ((st = rls_vec_create_internal(c, A->dtype, A->m, &K->vl)) != RLS_OK) break;
    if ((st = rls_vec_create_internal(c, A->dtype, A->m, &K->u)) != RLS_OK) break;
    int64_t grid = A->m < (int64_t)c->sm_count * 8 ? A->m : (int64_t)c->sm_count * 8;
    kz_rownorm2_kernel<<<(unsigned)grid, 256, 0, c->stream>>>((const float*)A->d, ldf, nfl, A->m, K->d_s2);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) { rls_set_error("Kaczmarz: rownorm2 launch failed"); st = RLS_ERR_CUDA; break; }
  } while (0);
  if (st != RLS_OK) { kz_free(K); return st; }
  *out = K;
  return RLS_OK;
}

extern "C" int32_t rls_kaczmarz_destroy(rls_kaczmarz_t K) {
  if (!K) return RLS_OK;
  RlsDeviceGuard g(K->ctx->device);
  cudaStreamSynchronize(K->ctx->stream);
  kz_free(K);
  return RLS_OK;
}

extern "C" int32_t rls_kaczmarz_block_rows(rls_kaczmarz_t K, int32_t* block_rows) {
  RLS_CHECK_ARG(K && block_rows, "NULL argument");
  *block_rows = K->R;
  return RLS_OK;
}

extern "C" int32_t rls_kaczmarz_rownorm2(rls_kaczmarz_t K, float* host, int64_t len) {
  RLS_CHECK_ARG(K && host, "NULL argument");
  RLS_CHECK_ARG(len == K->A->m, "rownorm2: expected %lld values", (long long)K->A->m);
  RlsDeviceGuard g(K->ctx->device);
  RLS_CUDA(cudaMemcpyAsync(host, K->d_s2, (size_t)len * 4, cudaMemcpyDeviceToHost, K->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(K->ctx->stream));
  return RLS_OK;
}

extern "C" int32_t rls_kaczmarz_set_rows(rls_kaczmarz_t K, const int64_t* rows, const float* denom, int64_t count) {
  RLS_CHECK_ARG(K && (count == 0 || (rows && denom)), "NULL argument");
  RLS_CHECK_ARG(count >= 0, "negative count");
  rls_ctx_s* c = K->ctx;
  RlsDeviceGuard g(c->device);
  const int R = K->R, fpe = K->fpe;
  const int64_t m = K->A->m;
  const int64_t nblk = (count + R - 1) / R;
  std::vector<int32_t> hr((size_t)(nblk * R), -1);
  std::vector<float> hd((size_t)(nblk * R), 0.f);
  std::vector<uint8_t> seen((size_t)m, 0);
  for (int64_t i = 0; i < count; ++i) {
    RLS_CHECK_ARG(rows[i] >= 0 && rows[i] < m, "set_rows: row %lld out of range", (long long)rows[i]);
    RLS_CHECK_ARG(!seen[(size_t)rows[i]], "set_rows: row %lld listed twice", (long long)rows[i]);
    seen[(size_t)rows[i]] = 1;
    hr[(size_t)i] = (int32_t)rows[i];
    hd[(size_t)i] = denom[i];
  }
  RLS_CUDA(cudaStreamSynchronize(c->stream));   // a sweep may still be reading the previous order
  if (nblk > K->cap_blk) {
    cudaFree(K->d_rows); cudaFree(K->d_denom); cudaFree(K->d_G);
    K->d_rows = nullptr; K->d_denom = nullptr; K->d_G = nullptr; K->cap_blk = 0;
    if (cudaMalloc(&K->d_rows, (size_t)nblk * R * 4) != cudaSuccess || cudaMalloc(&K->d_denom, (size_t)nblk * R * 4) != cudaSuccess ||
        cudaMalloc(&K->d_G, (size_t)nblk * R * R * fpe * 4) != cudaSuccess) {
      rls_set_error("Kaczmarz: cudaMalloc of the block Gram matrices failed: %s", cudaGetErrorString(cudaGetLastError()));
      return RLS_ERR_NOMEM;
    }
    K->cap_blk = nblk;
  }
  K->count = count; K->nblk = nblk;
  if (nblk == 0) return RLS_OK;
  RLS_CUDA(cudaMemcpyAsync(K->d_rows, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice, c->stream));
  RLS_CUDA(cudaMemcpyAsync(K->d_denom, hd.data(), hd.size() * 4, cudaMemcpyHostToDevice, c->stream));
  RLS_CUDA(cudaStreamSynchronize(c->stream));   // hr / hd are pageable and die with this call
  RLS_CUDA(cudaMemsetAsync(K->d_G, 0, (size_t)nblk * R * R * fpe * 4, c->stream));
  const int T = R / 64;
  const int64_t ldf = K->A->ld * fpe;
  for (int64_t b0 = 0; b0 < nblk; b0 += 32768) {
    const int64_t nb = std::min<int64_t>(32768, nblk - b0);
    dim3 grid((unsigned)(T * (T + 1) / 2), (unsigned)nb);
    if (fpe == 2) kz_gram_kernel<2><<<grid, 256, 0, c->stream>>>((const float*)K->A->d, ldf, K->A->n, K->d_rows, R, K->d_G, b0);
    else kz_gram_kernel<1><<<grid, 256, 0, c->stream>>>((const float*)K->A->d, ldf, K->A->n, K->d_rows, R, K->d_G, b0);
    c->launches++;
  }
  RLS_CUDA(cudaGetLastError());
  RLS_TRY(kz_plan_persistent(K));
  return RLS_OK;
}

extern "C" int32_t rls_kaczmarz_init(rls_kaczmarz_t K, rls_vec_t b, rls_vec_t x0, float eps_w) {
  RLS_CHECK_ARG(K && b, "NULL argument");
  RLS_CHECK_ARG(b->dtype == K->A->dtype && b->len == K->A->m, "Kaczmarz init: b must have %lld elements of the matrix type", (long long)K->A->m);
  RLS_CHECK_ARG(!x0 || (x0->dtype == K->A->dtype && x0->len == K->A->n), "Kaczmarz init: x0 must have %lld elements", (long long)K->A->n);
  rls_ctx_s* c = K->ctx;
  RlsDeviceGuard g(c->device);
  const size_t es = rls_elem_size(K->A->dtype);
  if (x0) RLS_CUDA(cudaMemcpyAsync(K->x->d, x0->d, (size_t)K->A->n * es, cudaMemcpyDeviceToDevice, c->stream));
  else RLS_CUDA(cudaMemsetAsync(K->x->d, 0, (size_t)K->A->n * es, c->stream));
  RLS_CUDA(cudaMemsetAsync(K->vl->d, 0, (size_t)K->A->m * es, c->stream));
  RLS_CUDA(cudaMemcpyAsync(K->u->d, b->d, (size_t)K->A->m * es, cudaMemcpyDeviceToDevice, c->stream));
  if (K->d_tpart2) {   // tags restart at 1 with zeroed exchange buffers
    RLS_CUDA(cudaMemsetAsync(K->d_tpart2, 0, K->tagged_bytes, c->stream));
    RLS_CUDA(cudaMemsetAsync(K->d_alpha2, 0, K->alpha2_bytes, c->stream));
    RLS_CUDA(cudaMemsetAsync(K->d_tsum, 0, K->alpha2_bytes, c->stream));
    RLS_CUDA(cudaMemsetAsync(K->d_abort, 0, sizeof(int), c->stream));
    K->epoch = 0;
  }
  K->eps_w = eps_w;
  K->initialised = true;
  return RLS_OK;
}

template <int FPE, int NF>
static int32_t kz_sweep_impl(rls_kaczmarz_s* K) {
  rls_ctx_s* c = K->ctx;
  const int R = K->R;
  const int64_t nfl = K->A->n * FPE, ldf = K->A->ld * FPE;
  const float* A = (const float*)K->A->d;
  float* x = (float*)K->x->d;
  const dim3 gdot((unsigned)K->S, (unsigned)((R + KZ_DOT_RG - 1) / KZ_DOT_RG));
  const int64_t packs = (nfl + NF - 1) / NF;
  const dim3 gupd((unsigned)((packs + 31) / 32));
  const size_t gbytes = (size_t)R * R * FPE * 4;
  const bool stage = gbytes <= 200 * 1024;
  if (stage) RLS_CUDA(cudaFuncSetAttribute(kz_solve_kernel<FPE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gbytes));
  int skip = 0;   // RLS_KACZMARZ_SKIP: timing experiments only (1 = dot, 2 = solve, 4 = update); results are wrong
  if (const char* e = getenv("RLS_KACZMARZ_SKIP")) skip = atoi(e);
  for (int64_t b = 0; b < K->nblk; ++b) {
    const int32_t* rows = K->d_rows + b * R;
    if (!(skip & 1)) RLS_CUDA(rls_launch_pdl(c->stream, gdot, dim3(KZ_DOT_THREADS), kz_dot_kernel<FPE, NF>, A, ldf, nfl, rows, R, (const float*)x, K->d_tpart));
    if (!(skip & 2)) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1); cfg.blockDim = dim3(R); cfg.dynamicSmemBytes = stage ? gbytes : 0; cfg.stream = c->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = rls_pdl_enabled() ? 1 : 0;
      const float* den = K->d_denom + b * R;
      const float* Gb = K->d_G + b * (int64_t)R * R * FPE;
      const float* tp = K->d_tpart;
      const float* uu = (const float*)K->u->d;
      float* vl = (float*)K->vl->d;
      RLS_CUDA(cudaLaunchKernelEx(&cfg, stage ? kz_solve_kernel<FPE, true> : kz_solve_kernel<FPE, false>, rows, den, Gb, R, tp, K->S, uu, vl,
                                  K->eps_w, K->d_alpha));
    }
    if (!(skip & 4)) RLS_CUDA(rls_launch_pdl(c->stream, gupd, dim3(KZ_UPD_THREADS), kz_update_kernel<FPE, NF>, A, ldf, nfl, rows, R, (const float*)K->d_alpha, x));
    c->launches += 3;
  }
  return RLS_OK;
}

extern "C" int32_t rls_kaczmarz_sweep(rls_kaczmarz_t K) {
  RLS_CHECK_ARG(K, "NULL argument");
  RLS_CHECK_ARG(K->initialised, "Kaczmarz sweep before init");
  RlsNvtxRange nvtx("rls: Kaczmarz iterate (one sweep over the rows)");
  RlsDeviceGuard g(K->ctx->device);
  if (K->persistent && K->nblk > 0) {
    rls_ctx_s* c = K->ctx;
    if (K->epoch + (unsigned long long)K->nblk + 2ull > 0xffffffffull) {
      // the exchange tags are the low 32 bits of epoch + block + 1: before they wrap (a tag of 0 would match a zeroed
      // or stale slot) restart the epoch with clean exchange buffers, as rls_kaczmarz_init does
      RLS_CUDA(cudaMemsetAsync(K->d_tpart2, 0, K->tagged_bytes, c->stream));
      RLS_CUDA(cudaMemsetAsync(K->d_alpha2, 0, K->alpha2_bytes, c->stream));
      RLS_CUDA(cudaMemsetAsync(K->d_tsum, 0, K->alpha2_bytes, c->stream));
      K->epoch = 0;
    }
    KzSweep sp;
    sp.A = (const float*)K->A->d; sp.ldf = K->A->ld * K->fpe; sp.npacks = K->A->n * K->fpe / 4;
    sp.rows = K->d_rows; sp.denom = K->d_denom; sp.G = K->d_G; sp.Dinv = K->d_Dinv;
    sp.R = K->R; sp.nblk = (int)K->nblk; sp.P = K->pP;
    sp.u = (const float*)K->u->d; sp.vl = (float*)K->vl->d; sp.ew = K->eps_w; sp.x = (float*)K->x->d;
    sp.tpart = K->d_tpart2; sp.tsum = K->d_tsum; sp.alpha = K->d_alpha2; sp.abort_flag = K->d_abort; sp.epoch = K->epoch;
    sp.trace = nullptr; sp.trace_block = 0;
    if (rls_env_flag("RLS_KACZMARZ_TRACE", false)) {
      if (!K->d_trace) { RLS_CUDA(cudaMalloc(&K->d_trace, 32 * sizeof(long long))); RLS_CUDA(cudaMemsetAsync(K->d_trace, 0, 32 * sizeof(long long), c->stream)); }
      sp.trace = K->d_trace; sp.trace_block = (int)(K->nblk / 2);
    }
    const void* fn = K->fpe == 2 ? (K->pIT == 4 ? (const void*)kz_sweep_kernel<2, 4> : (const void*)kz_sweep_kernel<2, 8>)
                                 : (K->pIT == 4 ? (const void*)kz_sweep_kernel<1, 4> : (const void*)kz_sweep_kernel<1, 8>);
    void* args[] = {&sp};
    RLS_CUDA(cudaLaunchCooperativeKernel(fn, dim3(K->pgrid), dim3(KZ_PT), args, K->psmem, c->stream));
    c->launches++;
    K->epoch += (unsigned long long)K->nblk;
    return RLS_OK;
  }
  if (K->fpe == 2) return K->vec4 ? kz_sweep_impl<2, 4>(K) : kz_sweep_impl<2, 2>(K);
  return K->vec4 ? kz_sweep_impl<1, 4>(K) : kz_sweep_impl<1, 1>(K);
}

extern "C" int32_t rls_kaczmarz_vec(rls_kaczmarz_t K, const char* name, rls_vec_t* out) {
  RLS_CHECK_ARG(K && name && out, "NULL argument");
  if (!strcmp(name, "x")) *out = K->x;
  else if (!strcmp(name, "vl")) *out = K->vl;
  else if (!strcmp(name, "u")) *out = K->u;
  else { rls_set_error("Kaczmarz state has no vector '%s' (x, vl, u)", name); return RLS_ERR_INVALID; }
  return RLS_OK;
}

// diagnostics: 0 = block Gram matrices [nblk][R*R*fpe], 1 = dot partials of the last block [S][R][fpe],
// 2 = alpha of the last block [R][fpe], 3 = padded denominators [nblk*R], 4 = RLS_KACZMARZ_TRACE stamps (32 x int64)
extern "C" int32_t rls_kaczmarz_debug(rls_kaczmarz_t K, int32_t which, float* host, int64_t nfloats) {
  RLS_CHECK_ARG(K && host, "NULL argument");
  RlsDeviceGuard g(K->ctx->device);
  const float* src = nullptr;
  int64_t have = 0;
  switch (which) {
    case 0: src = K->d_G; have = K->nblk * (int64_t)K->R * K->R * K->fpe; break;
    case 1: src = K->d_tpart; have = (int64_t)K->S * K->R * K->fpe; break;
    case 2: src = K->d_alpha; have = (int64_t)K->R * K->fpe; break;
    case 3: src = K->d_denom; have = K->nblk * (int64_t)K->R; break;
    case 4: src = (const float*)K->d_trace; have = K->d_trace ? 64 : 0; break;   // 32 clock64 stamps
    default: rls_set_error("kaczmarz_debug: which = %d", which); return RLS_ERR_INVALID;
  }
  RLS_CHECK_ARG(nfloats <= have, "kaczmarz_debug: %lld floats requested, %lld available", (long long)nfloats, (long long)have);
  RLS_CUDA(cudaMemcpyAsync(host, src, (size_t)nfloats * 4, cudaMemcpyDeviceToHost, K->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(K->ctx->stream));
  return RLS_OK;
}

// synchronise and report a timed-out exchange of the persistent sweep kernel (bounded waits raise an abort flag)
extern "C" int32_t rls_kaczmarz_check(rls_kaczmarz_t K) {
  RLS_CHECK_ARG(K, "NULL argument");
  RlsDeviceGuard g(K->ctx->device);
  int flag = 0;
  if (K->d_abort) RLS_CUDA(cudaMemcpyAsync(&flag, K->d_abort, sizeof(int), cudaMemcpyDeviceToHost, K->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(K->ctx->stream));
  if (flag) {
    rls_set_error("Kaczmarz sweep kernel timed out waiting for the block exchange (abort flag set)");
    return RLS_ERR_CUDA;
  }
  return RLS_OK;
}

/* "sweep: one cooperative kernel ..." or "sweep: 3 kernels per block ..." */
extern "C" int32_t rls_kaczmarz_describe(rls_kaczmarz_t K, char* buf, int32_t len) {
  RLS_CHECK_ARG(K && buf && len > 0, "bad argument");
  if (K->persistent)
    snprintf(buf, (size_t)len, "persistent: grid=%d x %d threads, %d packs of x per CTA, smem=%zu, block_rows=%d, blocks=%lld", K->pgrid, KZ_PT,
             K->pP, K->psmem, K->R, (long long)K->nblk);
  else
    snprintf(buf, (size_t)len, "chained: 3 kernels per block, block_rows=%d, blocks=%lld, vec4=%d", K->R, (long long)K->nblk, (int)K->vec4);
  return RLS_OK;
}
