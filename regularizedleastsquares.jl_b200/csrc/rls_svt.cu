// rls_svt.cu — singular-value soft-thresholding proximal maps:
//   NuclearRegularization   prox!(reg, x, λ):  U,S,V = svd(reshape(x, svtShape)); S = soft(S, λ); x = U S V'
//                           (src/proximalMaps/ProxNuclear.jl:27-32)
//   LLRRegularization       the same on every blockSize patch of an image series, patches x frames
//                           (src/proximalMaps/ProxLLR.jl:44-90 non-overlapping, :163-199 fully overlapping)
// The reference calls LAPACK's dense SVD on the host ("Computation is always performed on the CPU", ProxLLR.jl:7).
// Here the thresholding never forms U: with the q x q Gram matrix of the SHORT side, G = X'X = V S² V',
//     SVT_λ(X) = U max(S-λ,0) V' = X · W,      W = V diag(max(s_i-λ,0)/s_i) V'   (q x q),
// so the work is two streaming passes over X (Gram, then X·W) around a small Hermitian eigenproblem:
//   svt_gram_kernel   partial Gram matrices over 64-row chunks (a CTA walks several), accumulated in Float64   (HBM / L2 bound)
//   svt_eig_kernel    one warp per problem: sums the partials, two-sided Jacobi in Float64 in shared memory with the round-robin
//                     ordering (q/2 disjoint pairs rotated per round), builds W                     (latency bound, q <= 64)
//   svt_apply_kernel  out = X·W chunk by chunk, in place (or accumulated for the overlapping LLR)  (HBM / L2 bound)
// Forming G squares the condition number, which is why G, V and W live in Float64: the inputs are Float32, so G is
// exact to 2^-53 relative and singular values down to 1e-7 s_max keep full Float32 accuracy.
// When the short side is the ROW side (M <= 64 < N; LLR: fewer pixels per patch than frames) the same kernels run on the
// view Y = X' (conjugate on load and store).
#include <algorithm>

#include "rls_prox.cuh"

namespace {

constexpr int SVT_CH = 64;       // long-side rows per chunk
constexpr int SVT_THREADS = 256;
constexpr int SVT_MAXQ = 64;
constexpr int SVT_MAXD = 4;

struct SvtGeom {
  int mode;        // 0: matrix, short side = columns, element (l,i) at l + M i;  1: short side = rows, element (l,i) = conj(x[i + M l]);
                   // 2: LLR patches, element (l,i) = pixel l of the patch in frame i (short side = frames);
                   // 3: LLR patches seen transposed, element (l,i) = conj(pixel i in frame l) (short side = pixels)
  int q;           // short side
  int64_t L;       // long side
  int64_t M;       // leading dimension of the stored matrix (modes 0/1)
  int ndims;       // LLR
  int64_t shape[SVT_MAXD], stride[SVT_MAXD], block[SVT_MAXD], nblk[SVT_MAXD], shift[SVT_MAXD];
  int64_t npix;
};

// offset of element (l, i) of problem `prob`, or -1 when the pixel lies outside the image (LLR boundary patches:
// "remove out-of bounds idx and fill the corresponding entries with 0", ProxLLR.jl:62-65)
__device__ __forceinline__ int64_t svt_offset(const SvtGeom& g, int64_t prob, int64_t l, int i) {
  if (g.mode == 0) return l + g.M * (int64_t)i;
  if (g.mode == 1) return (int64_t)i + g.M * l;
  const int64_t frame = g.mode == 2 ? (int64_t)i : l;
  int64_t off = frame * g.npix, pr = prob, lr = g.mode == 2 ? l : (int64_t)i;
#pragma unroll
  for (int d = 0; d < SVT_MAXD; ++d) {
    if (d < g.ndims) {
      const int64_t o = (pr % g.nblk[d]) * g.block[d];   // patch origin along d (patches enumerate first dim fastest)
      pr /= g.nblk[d];
      const int64_t t = lr % g.block[d];                 // pixel inside the patch, first dim fastest
      lr /= g.block[d];
      int64_t c = o + t;                                 // coordinate in the circularly shifted image xs
      if (c >= g.shape[d]) return -1;
      c -= g.shift[d];                                   // xs[c] = x[c - shift] (circshift)
      if (c < 0) c += g.shape[d];
      off += c * g.stride[d];
    }
  }
  return off;
}

template <typename T> struct SvtElem;
template <> struct SvtElem<float> {
  __device__ static __forceinline__ float2 load(const float* x, int64_t off, int conj) { (void)conj; return make_float2(x[off], 0.f); }
  __device__ static __forceinline__ float make(double2 v, int conj) { (void)conj; return (float)v.x; }
  __device__ static __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct SvtElem<float2> {
  __device__ static __forceinline__ float2 load(const float2* x, int64_t off, int conj) {
    const float2 v = x[off];
    return make_float2(v.x, conj ? -v.y : v.y);
  }
  __device__ static __forceinline__ float2 make(double2 v, int conj) { return make_float2((float)v.x, (float)(conj ? -v.y : v.y)); }
  __device__ static __forceinline__ float2 add(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// stage the chunk [l0, l0 + SVT_CH) x [0, q) of problem `prob` into shared memory (zeros outside); the data are Float32
template <typename T>
__device__ __forceinline__ void svt_stage(const T* __restrict__ x, const SvtGeom& g, int64_t prob, int64_t l0, float2* __restrict__ xs /*[SVT_CH][q]*/) {
  const int q = g.q;
  const int total = SVT_CH * q;
  for (int idx = threadIdx.x; idx < total; idx += SVT_THREADS) {
    int l, i;
    if (g.mode & 1) { i = idx % q; l = idx / q; }             // x[i + M l] / pixels of a frame: i is the contiguous index
    else { l = idx % SVT_CH; i = idx / SVT_CH; }               // x[l + M i] / pixels of a patch: l is the contiguous index
    float2 v = make_float2(0.f, 0.f);
    if (l0 + l < g.L) {
      const int64_t off = svt_offset(g, prob, l0 + l, i);
      if (off >= 0) v = SvtElem<T>::load(x, off, g.mode & 1);
    }
    xs[l * q + i] = v;
  }
}

// Gpart[(prob - p0) * gridDim.y + blockIdx.y][a][b] = sum over the chunks blockIdx.y, blockIdx.y + gridDim.y, ... of
// sum_l conj(X[l][a]) X[l][b]: a CTA walks its chunks with the q*q/256 pair sums of every thread in registers, so a long
// matrix (NuclearRegularization with 10^6 rows) leaves a few hundred partial Gram matrices, not one per 64 rows
constexpr int SVT_NP = SVT_MAXQ * SVT_MAXQ / SVT_THREADS;   // pairs per thread at q = 64
template <typename T>
__global__ void __launch_bounds__(SVT_THREADS) svt_gram_kernel(const T* __restrict__ x, SvtGeom g, int64_t p0, int nchunk, double2* __restrict__ Gpart,
                                                               double* __restrict__ rowmax, const int* __restrict__ gate) {
  if (gate && *gate) return;
  __shared__ float2 xs[SVT_CH * SVT_MAXQ];
  __shared__ double wmax[2];
  const int64_t prob = p0 + blockIdx.x;
  const int q = g.q;
  double2 acc[SVT_NP];
#pragma unroll
  for (int k = 0; k < SVT_NP; ++k) acc[k] = make_double2(0.0, 0.0);
  double rmax = 0.0;
  for (int chunk = blockIdx.y; chunk < nchunk; chunk += gridDim.y) {
    __syncthreads();                       // the previous chunk has been consumed
    svt_stage<T>(x, g, prob, (int64_t)chunk * SVT_CH, xs);
    __syncthreads();
    if (rowmax && threadIdx.x < SVT_CH) {
      // largest squared norm of a long-side row: in the transposed LLR view these are the FRAMES, i.e. the diagonal of
      // the frames x frames Gram matrix whose largest |entry| the reference's shortcut needs (a Gram matrix has it on the diagonal)
      double r2 = 0.0;
      for (int i = 0; i < q; ++i) {
        const float2 v = xs[threadIdx.x * q + i];
        r2 = fma((double)v.x, (double)v.x, fma((double)v.y, (double)v.y, r2));
      }
      rmax = fmax(rmax, r2);
    }
#pragma unroll
    for (int k = 0; k < SVT_NP; ++k) {
      const int idx = threadIdx.x + k * SVT_THREADS;
      if (idx < q * q) {
        const int a = idx / q, b = idx - a * q;
        double2 s = acc[k];
#pragma unroll 4
        for (int l = 0; l < SVT_CH; ++l) {
          const float2 uf = xs[l * q + a], vf = xs[l * q + b];
          const double ux = uf.x, uy = uf.y, vx = vf.x, vy = vf.y;
          s.x = fma(ux, vx, fma(uy, vy, s.x));
          s.y = fma(ux, vy, fma(-uy, vx, s.y));
        }
        acc[k] = s;
      }
    }
  }
  if (rowmax) {
    if (threadIdx.x < SVT_CH) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = rmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) rowmax[(int64_t)blockIdx.x * gridDim.y + blockIdx.y] = fmax(wmax[0], wmax[1]);
  }
  double2* out = Gpart + ((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * (int64_t)(q * q);
#pragma unroll
  for (int k = 0; k < SVT_NP; ++k) {
    const int idx = threadIdx.x + k * SVT_THREADS;
    if (idx < q * q) out[idx] = acc[k];
  }
}

// One warp per problem.  Shared memory per warp: G and V, q x (q+1) double2 each (odd pitch: a lane per ROW is
// conflict-free).  W goes to Wout[(prob - p0)][a][b].
//   llr != 0: the reference's shortcut (ProxLLR.jl:67-71): ub = sqrt(norm(X'X, Inf)) — Julia's `norm` of a matrix is the
//   largest |entry| — and the whole patch is set to zero when λ >= ub.
__global__ void svt_eig_kernel(const double2* __restrict__ Gpart, const double* __restrict__ rowmax, int nchunk, int q, int64_t nprob, float lam,
                               const float* __restrict__ lam_dev, int llr, double2* __restrict__ Wout, const int* __restrict__ gate) {
  if (gate && *gate) return;
  extern __shared__ double2 svt_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t pl = (int64_t)blockIdx.x * wpb + warp;   // problem index inside this launch
  if (pl >= nprob) return;
  const int pitch = q + 1;
  double2* G = svt_smem + (size_t)warp * (2 * q * pitch + 64);     // + 64 double2: the rotation table of a round
  double2* V = G + q * pitch;
  const double thr = (double)(lam_dev ? *lam_dev : lam);

  // G = sum of the chunk partials, V = I
  double maxabs2 = 0.0, tr = 0.0;
  const double2* gp = Gpart + pl * nchunk * (int64_t)(q * q);
  for (int idx = lane; idx < q * q; idx += 32) {
    double2 s = make_double2(0.0, 0.0);
    for (int c = 0; c < nchunk; ++c) {
      const double2 v = gp[(int64_t)c * (q * q) + idx];
      s.x += v.x; s.y += v.y;
    }
    const int a = idx / q, b = idx - a * q;
    if (a == b) { s.y = 0.0; tr += s.x; }
    G[a * pitch + b] = s;
    V[a * pitch + b] = make_double2(a == b ? 1.0 : 0.0, 0.0);
    maxabs2 = fmax(maxabs2, s.x * s.x + s.y * s.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    maxabs2 = fmax(maxabs2, __shfl_xor_sync(0xffffffffu, maxabs2, o));
    tr += __shfl_xor_sync(0xffffffffu, tr, o);
  }
  __syncwarp();
  double2* W = Wout + pl * (int64_t)(q * q);
  bool zero = maxabs2 == 0.0;                               // all-zero patch: nothing to do, W = 0 leaves it zero
  if (llr && !zero) {
    double g2 = sqrt(maxabs2);                              // largest |entry| of the Gram matrix at hand
    if (rowmax) {                                           // transposed view: largest entry of the frames x frames Gram matrix
      g2 = 0.0;
      for (int c = 0; c < nchunk; ++c) g2 = fmax(g2, rowmax[pl * nchunk + c]);
    }
    const float ub = sqrtf((float)g2);
    if ((float)thr >= ub) zero = true;
  }
  if (zero) {
    for (int idx = lane; idx < q * q; idx += 32) W[idx] = make_double2(0.0, 0.0);
    return;
  }

  // Jacobi with the round-robin ("circle") ordering: a sweep is q-1 rounds (q rounded up to even), every round rotates q/2
  // DISJOINT pairs at once.  Lane k computes the parameters of pair k — for the pair (p, r) the unitary
  // [[c, s e^{iφ}], [-s e^{-iφ}, c]] with G[p][r] = |b| e^{iφ} annihilates G[p][r] — then all lanes apply the round's
  // rotations: G <- G J (lane = row), V <- V J, G <- J' G (lane = column).  One parameter chain and three warp
  // synchronisations per ROUND instead of per pair (q = 64: 63 instead of 2016 per sweep).
  double2* rot_s = V + q * pitch;                            // [32] s e^{iφ} of the round's pairs
  double* rot_c = reinterpret_cast<double*>(rot_s + 32);     // [32] c
  int* rot_p = reinterpret_cast<int*>(rot_c + 32);           // [32] p (-1: nothing to rotate)
  int* rot_r = rot_p + 32;                                   // [32] r
  const double tol2 = (1e-15 * tr) * (1e-15 * tr);
  const int qe = q + (q & 1), n1 = qe - 1, npairs = qe >> 1;
  for (int sweep = 0; sweep < 40; ++sweep) {
    unsigned rotated = 0;
    for (int rr = 0; rr < n1; ++rr) {
      int pa = -1, pb = -1;
      double c = 1.0;
      double2 sph = make_double2(0.0, 0.0);
      if (lane < npairs) {
        const int a = lane == 0 ? rr : (rr + lane) % n1;
        const int b2 = lane == 0 ? n1 : (rr - lane + n1) % n1;
        if (a < q && b2 < q) {                               // (the padding index of an odd q sits out)
          const int lo = min(a, b2), hi = max(a, b2);
          const double2 b = G[lo * pitch + hi];
          const double ab2 = b.x * b.x + b.y * b.y;
          if (ab2 > tol2) {
            // with d = G[r][r] - G[p][p], τ = d / (2|b|), t = sgn(τ) / (|τ| + sqrt(1 + τ²)), c = 1/sqrt(1 + t²), s = t c.  Written
            // without |b| and τ — h = sqrt(d² + 4|b|²), w = 1/(|d| + h): t² = 4|b|² w², s e^{iφ} = b · 2 sgn(d) w c — the dependent
            // chain is one sqrt, one division and one rsqrt
            const double d = G[hi * pitch + hi].x - G[lo * pitch + lo].x;
            const double h = sqrt(fma(d, d, 4.0 * ab2));
            const double w = 1.0 / (fabs(d) + h);
            c = rsqrt(fma(4.0 * ab2, w * w, 1.0));
            const double k = (d >= 0.0 ? 2.0 : -2.0) * w * c;
            sph = make_double2(k * b.x, k * b.y);
            pa = lo; pb = hi;
          }
        }
      }
      rot_p[lane] = pa; rot_r[lane] = pb; rot_c[lane] = c; rot_s[lane] = sph;
      const unsigned active = __ballot_sync(0xffffffffu, pa >= 0);
      __syncwarp();
      if (active == 0) continue;                             // warp-uniform
      rotated |= active;
      for (int row = lane; row < q; row += 32) {             // columns p, r of G and V:  col_p' = c col_p - s e^{-iφ} col_r,  col_r' = s e^{iφ} col_p + c col_r
        for (int k = 0; k < npairs; ++k) {
          const int p = rot_p[k];
          if (p < 0) continue;
          const int r = rot_r[k];
          const double ck = rot_c[k];
          const double2 sk = rot_s[k], skc = make_double2(sk.x, -sk.y);
          const double2 gpv = G[row * pitch + p], grv = G[row * pitch + r];
          const double2 a1 = cmul(skc, grv), a2 = cmul(sk, gpv);
          G[row * pitch + p] = make_double2(ck * gpv.x - a1.x, ck * gpv.y - a1.y);
          G[row * pitch + r] = make_double2(a2.x + ck * grv.x, a2.y + ck * grv.y);
          const double2 vpv = V[row * pitch + p], vrv = V[row * pitch + r];
          const double2 b1 = cmul(skc, vrv), b3 = cmul(sk, vpv);
          V[row * pitch + p] = make_double2(ck * vpv.x - b1.x, ck * vpv.y - b1.y);
          V[row * pitch + r] = make_double2(b3.x + ck * vrv.x, b3.y + ck * vrv.y);
        }
      }
      __syncwarp();
      for (int col = lane; col < q; col += 32) {             // rows p, r of G:  row_p' = c row_p - s e^{iφ} row_r,  row_r' = s e^{-iφ} row_p + c row_r
        for (int k = 0; k < npairs; ++k) {
          const int p = rot_p[k];
          if (p < 0) continue;
          const int r = rot_r[k];
          const double ck = rot_c[k];
          const double2 sk = rot_s[k], skc = make_double2(sk.x, -sk.y);
          const double2 rp = G[p * pitch + col], rw = G[r * pitch + col];
          const double2 a1 = cmul(sk, rw), a2 = cmul(skc, rp);
          G[p * pitch + col] = make_double2(ck * rp.x - a1.x, ck * rp.y - a1.y);
          G[r * pitch + col] = make_double2(a2.x + ck * rw.x, a2.y + ck * rw.y);
        }
      }
      __syncwarp();
      if (pa >= 0) {
        G[pa * pitch + pb] = make_double2(0.0, 0.0);
        G[pb * pitch + pa] = make_double2(0.0, 0.0);
        G[pa * pitch + pa].y = 0.0;
        G[pb * pitch + pb].y = 0.0;
      }
      __syncwarp();
    }
    if (!rotated) break;
  }
  __syncwarp();
  // D_k = max(s_k - λ, 0) / s_k  (prox!(L1Regularization, S, λ) on S >= 0 is exactly max(S - λ, 0)), stored in G's diagonal
  for (int k = lane; k < q; k += 32) {
    const double ev = G[k * pitch + k].x;
    const double sv = ev > 0.0 ? sqrt(ev) : 0.0;
    G[k * pitch + k].y = sv > 0.0 ? fmax(sv - thr, 0.0) / sv : 0.0;
  }
  __syncwarp();
  // W[a][b] = sum_k V[a][k] D_k conj(V[b][k])
  for (int idx = lane; idx < q * q; idx += 32) {
    const int a = idx / q, b = idx - a * q;
    double2 acc = make_double2(0.0, 0.0);
    for (int k = 0; k < q; ++k) {
      const double d = G[k * pitch + k].y;
      const double2 va = V[a * pitch + k], vb = V[b * pitch + k];
      // va * conj(vb) * d
      acc.x = fma(d, va.x * vb.x + va.y * vb.y, acc.x);
      acc.y = fma(d, va.y * vb.x - va.x * vb.y, acc.y);
    }
    W[idx] = acc;
  }
}

// out(l, j) = sum_i X(l, i) W[i][j] for the chunk.  acc == NULL: written in place over x; else acc[off] += value
// (fully overlapping LLR: every shift's result is summed, ProxLLR.jl:188-191).
template <typename T>
__global__ void __launch_bounds__(SVT_THREADS) svt_apply_kernel(T* __restrict__ x, SvtGeom g, int64_t p0, int nchunk, const double2* __restrict__ Wmat,
                                                                T* __restrict__ acc, const int* __restrict__ gate) {
  if (gate && *gate) return;
  extern __shared__ double2 svt_apply_smem[];
  const int q = g.q;
  double2* Ws = svt_apply_smem;                                     // [q][q]
  float2* xs = reinterpret_cast<float2*>(svt_apply_smem + q * q);   // [SVT_CH][q]
  const int64_t prob = p0 + blockIdx.x;
  const double2* W = Wmat + (int64_t)blockIdx.x * (q * q);
  for (int idx = threadIdx.x; idx < q * q; idx += SVT_THREADS) Ws[idx] = W[idx];
  for (int chunk = blockIdx.y; chunk < nchunk; chunk += gridDim.y) {     // W is loaded once per CTA, chunks are walked
    const int64_t l0 = (int64_t)chunk * SVT_CH;
    __syncthreads();     // the previous chunk has been written out
    svt_stage<T>(x, g, prob, l0, xs);
    __syncthreads();     // the whole chunk is on chip: the in-place stores below cannot disturb a later read
    const int total = SVT_CH * q;
    for (int idx = threadIdx.x; idx < total; idx += SVT_THREADS) {
      int l, j;
      if (g.mode & 1) { j = idx % q; l = idx / q; }
      else { l = idx % SVT_CH; j = idx / SVT_CH; }
      if (l0 + l >= g.L) continue;
      const int64_t off = svt_offset(g, prob, l0 + l, j);
      if (off < 0) continue;
      double2 s = make_double2(0.0, 0.0);
      for (int i = 0; i < q; ++i) {
        const float2 xf = xs[l * q + i];
        s = cfma(make_double2((double)xf.x, (double)xf.y), Ws[i * q + j], s);
      }
      const T v = SvtElem<T>::make(s, g.mode & 1);
      if (acc) acc[off] = SvtElem<T>::add(acc[off], v);
      else x[off] = v;
    }
  }
}

// x = acc / count (ProxLLR.jl:197), acc in place
template <typename T>
__global__ void __launch_bounds__(SVT_THREADS) svt_scale_kernel(T* __restrict__ x, const T* __restrict__ acc, int64_t n, float count,
                                                                const int* __restrict__ gate) {
  if (gate && *gate) return;
  for (int64_t i = (int64_t)blockIdx.x * SVT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * SVT_THREADS) x[i] = Elem<T>::divr(acc[i], count);
}

int32_t ensure_scratch(rls_ctx_s* c, size_t bytes) {
  if (c->svt_scratch_bytes >= bytes) return RLS_OK;
  if (c->svt_scratch) RLS_CUDA(cudaFree(c->svt_scratch));
  c->svt_scratch = nullptr;
  c->svt_scratch_bytes = 0;
  RLS_CUDA(cudaMalloc(&c->svt_scratch, bytes));
  c->svt_scratch_bytes = bytes;
  return RLS_OK;
}

// all problems of one geometry: Gram -> eigen -> apply, in batches that bound the scratch memory
template <typename T>
int32_t svt_run(rls_ctx_s* c, T* x, const SvtGeom& g, int64_t nprob, float lam, const float* lam_dev, int llr, T* acc, const int* gate) {
  const int q = g.q;
  const int64_t nchunk64 = (g.L + SVT_CH - 1) / SVT_CH;
  RLS_CHECK_ARG(nchunk64 <= 0x7fffffff, "singular-value thresholding: long side %lld too large", (long long)g.L);
  const int nchunk = (int)nchunk64;
  // CTAs per problem along the long side: one per chunk, but no more than ~4 waves of the device in total
  const int gy = (int)std::max<int64_t>(1, std::min<int64_t>(nchunk, (4 * (int64_t)c->sm_count + nprob - 1) / nprob));
  const size_t per_prob = (size_t)(gy + 1) * q * q * sizeof(double2) + (size_t)gy * sizeof(double);
  int64_t batch = (int64_t)std::max<size_t>(1, ((size_t)128 << 20) / per_prob);
  if (batch > nprob) batch = nprob;
  if (batch > 0x3fffffff) batch = 0x3fffffff;
  RLS_TRY(ensure_scratch(c, (size_t)batch * per_prob));
  double2* Gpart = (double2*)c->svt_scratch;
  double2* Wm = Gpart + (size_t)batch * gy * q * q;
  double* rowmax = g.mode == 3 ? (double*)(Wm + (size_t)batch * q * q) : nullptr;
  const int wpb = q > 16 ? 1 : 4;
  const size_t eig_smem = (size_t)wpb * (2 * (size_t)q * (q + 1) + 64) * sizeof(double2);
  const size_t apply_smem = (size_t)q * q * sizeof(double2) + (size_t)SVT_CH * q * sizeof(float2);
  if (eig_smem > 48 * 1024) RLS_CUDA(cudaFuncSetAttribute((const void*)svt_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_smem));
  if (apply_smem > 48 * 1024)
    RLS_CUDA(cudaFuncSetAttribute((const void*)svt_apply_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem));
  for (int64_t p0 = 0; p0 < nprob; p0 += batch) {
    const int64_t np = std::min(batch, nprob - p0);
    svt_gram_kernel<T><<<dim3((unsigned)np, (unsigned)gy), SVT_THREADS, 0, c->stream>>>(x, g, p0, nchunk, Gpart, rowmax, gate);
    svt_eig_kernel<<<(unsigned)((np + wpb - 1) / wpb), 32 * wpb, eig_smem, c->stream>>>(Gpart, rowmax, gy, q, np, lam, lam_dev, llr, Wm, gate);
    svt_apply_kernel<T><<<dim3((unsigned)np, (unsigned)gy), SVT_THREADS, apply_smem, c->stream>>>(x, g, p0, nchunk, Wm, acc, gate);
    c->launches += 3;
  }
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

}  // namespace

// NuclearRegularization: x is the column-major rows x cols matrix (ProxNuclear.jl:27-32)
int32_t rls_prox_nuclear_launch(rls_ctx_s* c, int32_t dtype, void* x, int64_t n, int64_t rows, int64_t cols, float lam, const float* lam_dev,
                                const int* gate) {
  RLS_CHECK_ARG(rows >= 1 && cols >= 1 && rows * cols == n, "NuclearRegularization: svtShape %lldx%lld does not match %lld elements",
                (long long)rows, (long long)cols, (long long)n);
  if (std::min(rows, cols) > SVT_MAXQ) {
    rls_set_error("NuclearRegularization: the shorter side of svtShape must be <= %d on the accelerated path (got %lldx%lld)", SVT_MAXQ,
                  (long long)rows, (long long)cols);
    return RLS_ERR_UNSUPPORTED;
  }
  SvtGeom g{};
  g.M = rows;
  if (cols <= rows) { g.mode = 0; g.q = (int)cols; g.L = rows; }          // the Gram matrix of the SHORTER side
  else { g.mode = 1; g.q = (int)rows; g.L = cols; }
  if (dtype == RLS_C32) return svt_run<float2>(c, (float2*)x, g, 1, lam, lam_dev, 0, nullptr, gate);
  return svt_run<float>(c, (float*)x, g, 1, lam, lam_dev, 0, nullptr, gate);
}

// LLRRegularization: x = image series shape x K (K = n / prod(shape) frames); every patch of `block` pixels is a
// (pixels x K) matrix that is singular-value thresholded.  shift: the circular shift of the patch grid
// (randshift, ProxLLR.jl:55), 0 <= shift[d] < shape[d].  acc != NULL: results are added into acc instead of written.
static int32_t llr_pass(rls_ctx_s* c, int32_t dtype, void* x, int64_t n, int32_t ndims, const int64_t* shape, const int64_t* block,
                        const int64_t* shift, float lam, const float* lam_dev, void* acc, const int* gate) {
  SvtGeom g{};
  g.mode = 2; g.ndims = ndims;
  int64_t npix = 1, nprob = 1, ppix = 1;
  for (int d = 0; d < ndims; ++d) {
    g.shape[d] = shape[d]; g.block[d] = block[d]; g.stride[d] = npix;
    g.nblk[d] = (shape[d] + block[d] - 1) / block[d];
    g.shift[d] = ((shift ? shift[d] : 0) % shape[d] + shape[d]) % shape[d];
    npix *= shape[d]; nprob *= g.nblk[d]; ppix *= block[d];
  }
  g.npix = npix;
  const int64_t K = n / npix;
  if (K <= ppix) { g.mode = 2; g.q = (int)K; g.L = ppix; }                     // short side = frames
  else { g.mode = 3; g.q = (int)ppix; g.L = K; }                               // short side = pixels of a patch
  if (dtype == RLS_C32) return svt_run<float2>(c, (float2*)x, g, nprob, lam, lam_dev, 1, (float2*)acc, gate);
  return svt_run<float>(c, (float*)x, g, nprob, lam, lam_dev, 1, (float*)acc, gate);
}

int32_t rls_prox_llr_launch(rls_ctx_s* c, int32_t dtype, void* x, int64_t n, int32_t ndims, const int64_t* shape, const int64_t* block,
                            const int64_t* shift, int fully_overlapping, float lam, const float* lam_dev, const int* gate) {
  RLS_CHECK_ARG(ndims >= 1 && ndims <= SVT_MAXD && shape && block, "LLRRegularization: 1..%d image dimensions", SVT_MAXD);
  int64_t npix = 1;
  for (int d = 0; d < ndims; ++d) {
    RLS_CHECK_ARG(shape[d] >= 1 && block[d] >= 1, "LLRRegularization: shape and blockSize must be positive");
    npix *= shape[d];
  }
  RLS_CHECK_ARG(n % npix == 0 && n >= npix, "LLRRegularization: %lld elements are not a whole number of %lld-pixel frames", (long long)n, (long long)npix);
  const int64_t K = n / npix;
  int64_t ppix_all = 1;
  for (int d = 0; d < ndims; ++d) ppix_all *= block[d];
  if (std::min(K, ppix_all) > SVT_MAXQ) {
    rls_set_error("LLRRegularization: min(frames, pixels per patch) must be <= %d on the accelerated path (got %lld frames, %lld pixels)",
                  SVT_MAXQ, (long long)K, (long long)ppix_all);
    return RLS_ERR_UNSUPPORTED;
  }
  if (!fully_overlapping) return llr_pass(c, dtype, x, n, ndims, shape, block, shift, lam, lam_dev, nullptr, gate);
  // fully overlapping (ProxLLR.jl:163-199): every shift (1..blockSize per dimension) of the patch grid, averaged.  The
  // reference pads the image to a multiple of blockSize and then reshapes the padded array with the UNPADDED shape
  // (:170-176, :50), which only works when no padding is needed.
  int64_t nshift = 1;
  for (int d = 0; d < ndims; ++d) {
    if (shape[d] % block[d] != 0) {
      rls_set_error("LLRRegularization(fullyOverlapping=true): shape must be a multiple of blockSize (the reference fails on the padded reshape)");
      return RLS_ERR_UNSUPPORTED;
    }
    nshift *= block[d];
  }
  const size_t es = rls_elem_size(dtype);
  void* acc = nullptr;
  RLS_CUDA(cudaMalloc(&acc, (size_t)n * es));
  int32_t st = (cudaMemsetAsync(acc, 0, (size_t)n * es, c->stream) == cudaSuccess) ? RLS_OK : RLS_ERR_CUDA;
  for (int64_t sidx = 0; sidx < nshift && st == RLS_OK; ++sidx) {
    int64_t sh[SVT_MAXD] = {0, 0, 0, 0}, r = sidx;
    for (int d = 0; d < ndims; ++d) {            // block_idx enumerates 1..blockSize, first dimension fastest
      sh[d] = 1 + r % block[d] + (shift ? shift[d] : 0);
      r /= block[d];
    }
    st = llr_pass(c, dtype, x, n, ndims, shape, block, sh, lam, lam_dev, acc, gate);
  }
  if (st == RLS_OK) {
    const int grid = (int)std::min<int64_t>((n + SVT_THREADS - 1) / SVT_THREADS, (int64_t)c->sm_count * 8);
    if (dtype == RLS_C32) svt_scale_kernel<float2><<<grid, SVT_THREADS, 0, c->stream>>>((float2*)x, (const float2*)acc, n, (float)nshift, gate);
    else svt_scale_kernel<float><<<grid, SVT_THREADS, 0, c->stream>>>((float*)x, (const float*)acc, n, (float)nshift, gate);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) st = RLS_ERR_CUDA;
  }
  cudaStreamSynchronize(c->stream);
  cudaFree(acc);
  return st;
}

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" int32_t rls_prox_nuclear(rls_vec_t x, float lambda, int64_t rows, int64_t cols) {
  RLS_CHECK_ARG(x, "vec is NULL");
  RlsDeviceGuard g(x->ctx->device);
  return rls_prox_nuclear_launch(x->ctx, x->dtype, x->d, x->len, rows, cols, lambda, nullptr, nullptr);
}

extern "C" int32_t rls_prox_llr(rls_vec_t x, float lambda, int32_t ndims, const int64_t* shape, const int64_t* block_size, const int64_t* shift,
                                int32_t fully_overlapping) {
  RLS_CHECK_ARG(x && shape && block_size, "NULL argument");
  RlsDeviceGuard g(x->ctx->device);
  return rls_prox_llr_launch(x->ctx, x->dtype, x->d, x->len, ndims, shape, block_size, shift, fully_overlapping, lambda, nullptr, nullptr);
}
