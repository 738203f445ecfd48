// rls_gemv.cu — HBM-streaming dense matrix-vector kernels for column-major A
// (Float32 / interleaved ComplexF32):
//   gemv_n : y = A x        (mul!(y, A, x))
//   gemv_c : g = A' y       (mul!(g, adjoint(A), y); FISTA.jl:114, CGNR.jl:132, ADMM.jl:198)
// Both stream A exactly once with 128-bit coalesced loads; x is broadcast from shared
// memory, y is re-read through L1, reductions are warp shuffles + fixed-order partial
// sums (deterministic — no floating-point atomics).
//
// Roofline: HBM.  Algorithmic bytes per launch = m*n*sizeof(T) (vectors are noise).
#include "rls_common.cuh"

namespace {

constexpr int GEMV_THREADS = 256;

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ----------------------------------------------------------------------------------
// gemv_n: thread owns VEC consecutive rows (one 16-byte slice of each column) and
// walks a chunk of columns; the CTA covers GEMV_THREADS*VEC rows.  grid = (row blocks,
// column chunks).  Column-chunk partials go to scratch and the last CTA of each row
// block (ticket) adds them in chunk order.
// ----------------------------------------------------------------------------------
template <typename T> struct AccN;
template <> struct AccN<float> {
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  __device__ __forceinline__ void fma(float4 v, float x) {
    a[0] = fmaf(v.x, x, a[0]); a[1] = fmaf(v.y, x, a[1]); a[2] = fmaf(v.z, x, a[2]); a[3] = fmaf(v.w, x, a[3]);
  }
  __device__ __forceinline__ float4 get() const { return make_float4(a[0], a[1], a[2], a[3]); }
};
template <> struct AccN<float2> {
  float a[4] = {0.f, 0.f, 0.f, 0.f};  // (re0, im0, re1, im1)
  __device__ __forceinline__ void fma(float4 v, float2 x) {
    a[0] = fmaf(v.x, x.x, a[0]); a[0] = fmaf(-v.y, x.y, a[0]);
    a[1] = fmaf(v.x, x.y, a[1]); a[1] = fmaf(v.y, x.x, a[1]);
    a[2] = fmaf(v.z, x.x, a[2]); a[2] = fmaf(-v.w, x.y, a[2]);
    a[3] = fmaf(v.z, x.y, a[3]); a[3] = fmaf(v.w, x.x, a[3]);
  }
  __device__ __forceinline__ float4 get() const { return make_float4(a[0], a[1], a[2], a[3]); }
};

constexpr int XTILE = 1024;  // columns of x staged in shared memory at a time

template <typename T>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_n_kernel(const T* __restrict__ A, int64_t ld, int64_t m, int64_t n,
                                                               const T* __restrict__ x, T* __restrict__ y,
                                                               float4* __restrict__ scratch, unsigned* __restrict__ tickets,
                                                               int64_t cols_per_chunk, const int* __restrict__ gate) {
  if (gate && *gate) return;
  constexpr int VEC = Elem<T>::vec;
  __shared__ T xs[XTILE];
  __shared__ bool s_last;
  const int64_t mv = (m + VEC - 1) / VEC;            // 16-byte row slices
  const int64_t rv = (int64_t)blockIdx.x * GEMV_THREADS + threadIdx.x;
  const bool active = rv < mv;
  const int64_t j0 = (int64_t)blockIdx.y * cols_per_chunk;
  const int64_t j1 = min(n, j0 + cols_per_chunk);
  const float4* __restrict__ Ap = reinterpret_cast<const float4*>(A) + (active ? rv : 0);
  const int64_t ldv = ld / VEC;
  AccN<T> acc;
  for (int64_t jt = j0; jt < j1; jt += XTILE) {
    const int cnt = (int)min((int64_t)XTILE, j1 - jt);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += GEMV_THREADS) xs[k] = x[jt + k];
    __syncthreads();
    if (active) {
      const float4* __restrict__ col = Ap + jt * ldv;
      int k = 0;
      for (; k + 8 <= cnt; k += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ldg_stream(col + (int64_t)(k + u) * ldv);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc.fma(v[u], xs[k + u]);
      }
      for (; k < cnt; ++k) acc.fma(ldg_stream(col + (int64_t)k * ldv), xs[k]);
    }
  }
  float4 r = acc.get();
  const int nchunks = gridDim.y;
  float4* yv = reinterpret_cast<float4*>(y);
  auto store_masked = [&](float4 val) {
    // rows beyond m in the last slice are padding: store only valid scalars
    const int64_t row0 = rv * VEC;
    if (row0 + VEC <= m) { yv[rv] = val; return; }
    float tmp[4] = {val.x, val.y, val.z, val.w};
    float* ys = reinterpret_cast<float*>(y);
    constexpr int FPE = 4 / VEC;  // floats per element
    for (int e = 0; e < VEC; ++e)
      if (row0 + e < m)
        for (int f = 0; f < FPE; ++f) ys[(row0 + e) * FPE + f] = tmp[e * FPE + f];
  };
  if (nchunks == 1) {
    if (active) store_masked(r);
    return;
  }
  if (active) scratch[(int64_t)blockIdx.y * mv + rv] = r;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(&tickets[blockIdx.x], 1u);
    s_last = (prev == (unsigned)nchunks - 1);
    if (s_last) tickets[blockIdx.x] = 0u;
  }
  __syncthreads();
  if (!s_last || !active) return;
  __threadfence();
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nchunks; ++c) {
    float4 p = __ldcg(&scratch[(int64_t)c * mv + rv]);
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  store_masked(s);
}

// ----------------------------------------------------------------------------------
// gemv_c: a group of GT threads (a warp, or the whole CTA for tall matrices) owns CW
// columns at a time and strides down the rows with 128-bit loads; y slices are
// re-read through L1 (identical for every column).  Fixed-tree shuffle reduction.
// ----------------------------------------------------------------------------------
template <typename T> struct AccC;
template <> struct AccC<float> {
  float s = 0.f;
  __device__ __forceinline__ void fma(float4 a, float4 y) {
    s = fmaf(a.x, y.x, s); s = fmaf(a.y, y.y, s); s = fmaf(a.z, y.z, s); s = fmaf(a.w, y.w, s);
  }
};
template <> struct AccC<float2> {
  float re = 0.f, im = 0.f;  // conj(a) * y
  __device__ __forceinline__ void fma(float4 a, float4 y) {
    re = fmaf(a.x, y.x, re); re = fmaf(a.y, y.y, re); im = fmaf(a.x, y.y, im); im = fmaf(-a.y, y.x, im);
    re = fmaf(a.z, y.z, re); re = fmaf(a.w, y.w, re); im = fmaf(a.z, y.w, im); im = fmaf(-a.w, y.z, im);
  }
};

template <typename T>
__device__ __forceinline__ float4 mask_rows(float4 v, int64_t row0, int64_t m) {
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

template <typename T, int GT, int CW>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_c_kernel(const T* __restrict__ A, int64_t ld, int64_t m, int64_t n,
                                                               const T* __restrict__ y, T* __restrict__ g,
                                                               const int* __restrict__ gate) {
  if (gate && *gate) return;
  constexpr int VEC = Elem<T>::vec;
  constexpr int GROUPS = GEMV_THREADS / GT;
  constexpr int NW = GT / 32;
  constexpr int NF = Elem<T>::is_complex ? 2 : 1;
  __shared__ float s_red[GROUPS][NW][CW * NF];
  const int grp = threadIdx.x / GT, tig = threadIdx.x % GT;
  const int lane = threadIdx.x & 31, wig = tig >> 5;
  const int64_t mv = (m + VEC - 1) / VEC;
  const int64_t ldv = ld / VEC;
  const int64_t units = (n + CW - 1) / CW;
  const float4* __restrict__ Av = reinterpret_cast<const float4*>(A);
  const float4* __restrict__ yv = reinterpret_cast<const float4*>(y);
  for (int64_t u = (int64_t)blockIdx.x * GROUPS + grp; u < units; u += (int64_t)gridDim.x * GROUPS) {
    const int64_t jb = u * CW;
    AccC<T> acc[CW];
    const float4* __restrict__ cp[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) cp[c] = Av + min(jb + c, n - 1) * ldv;  // clamp: duplicates are discarded at the store
    int64_t r = tig;
    // main loop: two row slices in flight per column
    for (; r + GT < mv; r += 2 * GT) {
      float4 a0[CW], a1[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) { a0[c] = ldg_stream(cp[c] + r); a1[c] = ldg_stream(cp[c] + r + GT); }
      float4 y0 = __ldg(yv + r), y1 = __ldg(yv + r + GT);
      if ((r + GT + 1) * VEC > m) {  // only the very last slice can contain padding rows
        y1 = mask_rows<T>(y1, (r + GT) * VEC, m);
#pragma unroll
        for (int c = 0; c < CW; ++c) a1[c] = mask_rows<T>(a1[c], (r + GT) * VEC, m);
      }
#pragma unroll
      for (int c = 0; c < CW; ++c) { acc[c].fma(a0[c], y0); acc[c].fma(a1[c], y1); }
    }
    for (; r < mv; r += GT) {
      float4 y0 = mask_rows<T>(__ldg(yv + r), r * VEC, m);
#pragma unroll
      for (int c = 0; c < CW; ++c) acc[c].fma(mask_rows<T>(ldg_stream(cp[c] + r), r * VEC, m), y0);
    }
    // reduce across the group
    float red[CW * NF];
#pragma unroll
    for (int c = 0; c < CW; ++c) {
      if constexpr (Elem<T>::is_complex) { red[2 * c] = warp_sum(acc[c].re); red[2 * c + 1] = warp_sum(acc[c].im); }
      else red[c] = warp_sum(acc[c].s);
    }
    if constexpr (NW > 1) {
      __syncthreads();  // GT == CTA here, so a CTA barrier is legal
      if (lane == 0)
#pragma unroll
        for (int k = 0; k < CW * NF; ++k) s_red[grp][wig][k] = red[k];
      __syncthreads();
      if (tig < CW * NF) {
        float t = 0.f;
        for (int w = 0; w < NW; ++w) t += s_red[grp][w][tig];
        const int c = tig / NF;
        if (jb + c < n) reinterpret_cast<float*>(g)[(jb + c) * NF + (tig % NF)] = t;
      }
    } else {
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < CW; ++c)
          if (jb + c < n) {
            if constexpr (Elem<T>::is_complex) g[jb + c] = make_float2(red[2 * c], red[2 * c + 1]);
            else g[jb + c] = red[c];
          }
      }
    }
  }
}

static bool aligned_for_vec(const rls_mat_s* A) {
  int64_t vec = A->dtype == RLS_C32 ? 2 : 4;
  return ((uintptr_t)A->d % 16 == 0) && (A->ld % vec == 0);
}

}  // namespace

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------
int32_t rls_gemv_n_raw(rls_mat_s* A, const void* x, void* y, const int* gate) {
  rls_ctx_s* c = A->ctx;
  if (A->m == 0) return RLS_OK;
  if (A->layout == RLS_LAYOUT_ROWMAJOR) {
    RowPlan* rp = rls_mat_rowplan(A);
    if (!rp) return RLS_ERR_UNSUPPORTED;
    if (A->n == 0) { RLS_CUDA(cudaMemsetAsync(y, 0, A->m * rls_elem_size(A->dtype), c->stream)); return RLS_OK; }
    return rls_rowpass_gemv_n(rp, x, y, gate);
  }
  RLS_CHECK_ARG(aligned_for_vec(A), "matrix storage must be 16-byte aligned with ld a multiple of %d elements",
                A->dtype == RLS_C32 ? 2 : 4);
  const int64_t vec = A->dtype == RLS_C32 ? 2 : 4;
  const int64_t mv = (A->m + vec - 1) / vec;
  const int64_t row_blocks = (mv + GEMV_THREADS - 1) / GEMV_THREADS;
  RLS_CHECK_ARG(row_blocks <= 4096, "m too large for gemv_n row-block tickets");
  // aim for ~4 co-resident CTAs per SM, all with equal work
  int64_t target = (int64_t)c->sm_count * 4;
  int64_t chunks = (target + row_blocks - 1) / row_blocks;
  if (chunks < 1) chunks = 1;
  int64_t max_chunks = (A->n + 63) / 64;  // at least 64 columns per chunk
  if (max_chunks < 1) max_chunks = 1;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 65535) chunks = 65535;
  int64_t cpc = A->n > 0 ? (A->n + chunks - 1) / chunks : 1;
  chunks = A->n > 0 ? (A->n + cpc - 1) / cpc : 1;
  if (chunks > 1) RLS_TRY(rls_ensure_gemv_scratch(c, (size_t)chunks * (size_t)mv * 16));
  dim3 grid((unsigned)row_blocks, (unsigned)chunks);
  if (A->dtype == RLS_C32)
    gemv_n_kernel<float2><<<grid, GEMV_THREADS, 0, c->stream>>>((const float2*)A->d, A->ld, A->m, A->n, (const float2*)x, (float2*)y, (float4*)c->gemv_scratch, c->gemv_tickets, cpc, gate);
  else
    gemv_n_kernel<float><<<grid, GEMV_THREADS, 0, c->stream>>>((const float*)A->d, A->ld, A->m, A->n, (const float*)x, (float*)y, (float4*)c->gemv_scratch, c->gemv_tickets, cpc, gate);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

template <typename T>
static void launch_gemv_c(rls_mat_s* A, const void* y, void* g, const int* gate) {
  rls_ctx_s* c = A->ctx;
  constexpr int VEC = Elem<T>::vec;
  constexpr int CW = 4;
  const int64_t mv = (A->m + VEC - 1) / VEC;
  const int64_t units = (A->n + CW - 1) / CW;
  const int64_t cap = (int64_t)c->sm_count * 8;
  if (mv >= 8 * GEMV_THREADS) {
    // tall: the whole CTA walks each column group (fine-grained units => no tail imbalance)
    int64_t grid = units < cap ? units : cap;
    gemv_c_kernel<T, GEMV_THREADS, CW><<<(unsigned)grid, GEMV_THREADS, 0, c->stream>>>((const T*)A->d, A->ld, A->m, A->n, (const T*)y, (T*)g, gate);
  } else {
    constexpr int GROUPS = GEMV_THREADS / 32;
    int64_t grid = (units + GROUPS - 1) / GROUPS;
    if (grid > cap) grid = cap;
    gemv_c_kernel<T, 32, CW><<<(unsigned)grid, GEMV_THREADS, 0, c->stream>>>((const T*)A->d, A->ld, A->m, A->n, (const T*)y, (T*)g, gate);
  }
}

int32_t rls_gemv_c_raw(rls_mat_s* A, const void* y, void* g, const int* gate) {
  rls_ctx_s* c = A->ctx;
  if (A->n == 0) return RLS_OK;
  if (A->layout == RLS_LAYOUT_ROWMAJOR) {
    RowPlan* rp = rls_mat_rowplan(A);
    if (!rp) return RLS_ERR_UNSUPPORTED;
    if (A->m == 0) { RLS_CUDA(cudaMemsetAsync(g, 0, A->n * rls_elem_size(A->dtype), c->stream)); return RLS_OK; }
    return rls_rowpass_gemv_c(rp, y, g, gate);
  }
  RLS_CHECK_ARG(aligned_for_vec(A), "matrix storage must be 16-byte aligned with ld a multiple of %d elements",
                A->dtype == RLS_C32 ? 2 : 4);
  if (A->dtype == RLS_C32) launch_gemv_c<float2>(A, y, g, gate);
  else launch_gemv_c<float>(A, y, g, gate);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

extern "C" int32_t rls_gemv_n(rls_mat_t A, rls_vec_t x, rls_vec_t y) {
  RLS_CHECK_ARG(A && x && y, "NULL argument");
  RLS_CHECK_ARG(x->dtype == A->dtype && y->dtype == A->dtype, "gemv_n: dtype mismatch");
  RLS_CHECK_ARG(x->len == A->n && y->len == A->m, "gemv_n: A is %lldx%lld, x has %lld, y has %lld", (long long)A->m,
                (long long)A->n, (long long)x->len, (long long)y->len);
  RlsDeviceGuard g(A->ctx->device);
  return rls_gemv_n_raw(A, x->d, y->d, nullptr);
}

extern "C" int32_t rls_gemv_c(rls_mat_t A, rls_vec_t y, rls_vec_t g) {
  RLS_CHECK_ARG(A && y && g, "NULL argument");
  RLS_CHECK_ARG(y->dtype == A->dtype && g->dtype == A->dtype, "gemv_c: dtype mismatch");
  RLS_CHECK_ARG(y->len == A->m && g->len == A->n, "gemv_c: A is %lldx%lld, y has %lld, g has %lld", (long long)A->m,
                (long long)A->n, (long long)y->len, (long long)g->len);
  RlsDeviceGuard gd(A->ctx->device);
  return rls_gemv_c_raw(A, y->d, g->d, nullptr);
}
