// rls_rowpass.cu — one-HBM-pass normal operator g = A'(A x) (and y = A x, g = A' y) for a
// ROW-MAJOR device copy of the system matrix.
//
// Replaces mul!(res, AHA, x) (src/FISTA.jl:152, src/POGM.jl:181, src/OptISTA.jl:182, src/CGNR.jl:151,
// cg! inside src/ADMM.jl:244) and mul!(x0, adjoint(A), b) (src/FISTA.jl:114, src/CGNR.jl:132).
//
// Why row-major.  A'(A x) = sum_i conj(a_i) (a_i . x) over the ROWS a_i of A.  With the rows contiguous
// in HBM a thread-block cluster can take a row, keep it on chip, form y_i = a_i . x and immediately
// add conj(a_i) y_i into its accumulators: A crosses HBM -> SM exactly once, and the only exchange is
// the handful of partial dot products between the CTAs of ONE cluster (distributed shared memory,
// ~200 cycles) instead of a grid-wide all-reduce per row panel.
//
// Decomposition.  A cluster of G CTAs (G = 1,2,4,8,16) owns whole rows; CTA `rank` owns the column
// slice [rank*W, rank*W+W) of every row (W <= V*2048 floats).  Clusters take rounds of R consecutive
// rows, round-robin.  Per CTA:
//   producer warp : one lane streams the CTA's slice of each row with 1-D bulk copies
//                   (cp.async.bulk global->shared, mbarrier complete_tx) into an NS-stage ring.
//   16 compute warps: thread t owns the float4 groups t, t+512, ... of the slice for the whole
//                   launch: x values and g accumulators live in registers.  A row slice is moved
//                   ring -> registers once (the stage is released at once, so the ring is purely
//                   the landing zone of the HBM stream and keeps filling while the exchange runs),
//                   partial dot products are reduced warp -> CTA -> cluster (fixed order, so all
//                   CTAs hold bit-identical y and the result is deterministic), and the R rows held
//                   in registers are folded into g.
// At the end every cluster writes its partial g; a second tiny kernel sums the partials over clusters
// in a fixed order.
#include "rls_common.cuh"

namespace {

constexpr int RP_CW = 16;              // compute warps
constexpr int RP_CT = RP_CW * 32;      // compute threads
constexpr int RP_THREADS = RP_CT;      // no producer warp: 16 warps -> 128 registers per thread
constexpr int RP_MAXG = 16;            // largest cluster
constexpr int RP_MAXR = 4;             // rows per round (template parameter R <= RP_MAXR)
enum { RP_NORMAL = 0, RP_GEMV_N = 1, RP_GEMV_C = 2 };

struct RowpassArgs {
  const float* A;     // row-major, row stride ldf floats
  int64_t ldf;
  int64_t m;
  int nf;             // floats per row (n * FPE)
  int W;              // slice width in floats (multiple of 4)
  int NS;             // ring stages
  int R;              // rows per round (host-side copy of the template parameter)
  const float* x;     // n-vector            (NORMAL, GEMV_N)
  const float* xold;  // fused FISTA momentum (FISTA.jl:144-148): the operator is applied to x*c1 + xold*c2 with
  const float* th_old;//   c1 = (1-θold)/θ, c2 = (θold-1)/θ + 1 formed from the device-resident θ's; NULL = plain x
  const float* th;
  const float* yin;   // m-vector            (GEMV_C)
  float* yout;        // m-vector            (GEMV_N)
  float* gpart;       // [nclusters][gstride] (NORMAL, GEMV_C)
  int64_t gstride;
  const int* gate;
  int* abort_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_count() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
template <bool CLUSTER_SCOPE>
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  if (CLUSTER_SCOPE)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a wait that does not complete within ~1 s raises the abort flags (shared + global);
// every later wait returns at once, so a protocol error ends the launch instead of hanging the GPU.
template <bool CLUSTER_SCOPE>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, volatile int* s_abort, int* g_abort) {
  if (mbar_try<CLUSTER_SCOPE>(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try<CLUSTER_SCOPE>(bar, parity)) {
    if (*s_abort) return;
    if (clock64() - t0 > 2000000000ll) {
      *s_abort = 1;
      atomicExch(g_abort, 1);
      return;
    }
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// Deposit one float in CTA `peer`'s copy of a shared slot and signal its mbarrier in the SAME message
// (st.async + complete_tx): no release fence on the sender (a release.cluster arrive compiles to
// MEMBAR.ALL.GPU), no L1 invalidate on the receiver.
__device__ __forceinline__ void dsmem_post(float* local_slot, uint64_t* local_bar, unsigned peer, float v) {
  uint32_t rs, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rs) : "r"(smem_u32(local_slot)), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(peer));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rs), "r"(__float_as_uint(v)), "r"(rb) : "memory");
}

// partial dot product a . x over one float4 group
template <int FPE> __device__ __forceinline__ void dot_acc(float (&acc)[FPE], float4 a, float4 x);
template <> __device__ __forceinline__ void dot_acc<1>(float (&acc)[1], float4 a, float4 x) {
  acc[0] = fmaf(a.x, x.x, acc[0]); acc[0] = fmaf(a.y, x.y, acc[0]); acc[0] = fmaf(a.z, x.z, acc[0]); acc[0] = fmaf(a.w, x.w, acc[0]);
}
template <> __device__ __forceinline__ void dot_acc<2>(float (&acc)[2], float4 a, float4 x) {
  acc[0] = fmaf(a.x, x.x, acc[0]); acc[0] = fmaf(-a.y, x.y, acc[0]); acc[1] = fmaf(a.x, x.y, acc[1]); acc[1] = fmaf(a.y, x.x, acc[1]);
  acc[0] = fmaf(a.z, x.z, acc[0]); acc[0] = fmaf(-a.w, x.w, acc[0]); acc[1] = fmaf(a.z, x.w, acc[1]); acc[1] = fmaf(a.w, x.z, acc[1]);
}
// g += conj(a) * y
template <int FPE> __device__ __forceinline__ void axpy_conj(float4& g, float4 a, float yr, float yi);
template <> __device__ __forceinline__ void axpy_conj<1>(float4& g, float4 a, float yr, float) {
  g.x = fmaf(a.x, yr, g.x); g.y = fmaf(a.y, yr, g.y); g.z = fmaf(a.z, yr, g.z); g.w = fmaf(a.w, yr, g.w);
}
template <> __device__ __forceinline__ void axpy_conj<2>(float4& g, float4 a, float yr, float yi) {
  g.x = fmaf(a.x, yr, g.x); g.x = fmaf(a.y, yi, g.x); g.y = fmaf(a.x, yi, g.y); g.y = fmaf(-a.y, yr, g.y);
  g.z = fmaf(a.z, yr, g.z); g.z = fmaf(a.w, yi, g.z); g.w = fmaf(a.z, yi, g.w); g.w = fmaf(-a.w, yr, g.w);
}

template <int FPE, int V, int R, int MODE, bool FULL>
__global__ void __launch_bounds__(RP_THREADS, 1) rowpass_kernel(RowpassArgs p) {
  pdl_prologue();
  if (p.gate && *p.gate) return;  // device-side done() gate: uniform over the grid
  constexpr int RF = R * FPE;     // floats exchanged per round
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned G = cluster_size(), rank = cluster_rank(), cid = cluster_id(), ncl = cluster_count();
  const int W = p.W, NS = p.NS;

  float* ring = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * W * 4);
  uint64_t* ybar = full + NS;                                    // [2]
  float* wpart = reinterpret_cast<float*>(ybar + 2);             // [RP_CW][RF]
  float* ybuf = wpart + RP_CW * RF;                              // [2][RP_MAXG][RF]
  float* ysm = ybuf + 2 * RP_MAXG * RF;                          // [RF]
  volatile int* s_abort = reinterpret_cast<volatile int*>(ysm + RF);

  const int nf_pad = (p.nf + 3) & ~3;
  const int col0 = (int)rank * W;
  const int slice = max(0, min(W, nf_pad - col0));  // floats of this CTA's slice (multiple of 4)
  const int64_t nrounds = (p.m + R - 1) / R;
  const int nr = (int)((nrounds - cid + ncl - 1) / ncl);  // rounds of this cluster (>= 1)
  const int rv_last = (int)min((int64_t)R, p.m - ((int64_t)cid + (int64_t)(nr - 1) * ncl) * R);
  const int Q = (nr - 1) * R + rv_last;                   // rows this cluster streams, in sequence q = 0..Q-1

  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  // bulk copy of row q of this cluster's sequence into ring stage st (one elected thread)
  auto issue = [&](int q, int st) {
    if (slice > 0) {
      const int64_t row = ((int64_t)cid + (int64_t)(q / R) * ncl) * R + (q % R);
      mbar_expect_tx(&full[st], (unsigned)slice * 4u);
      bulk_load(ring + (size_t)st * W, p.A + row * p.ldf + col0, (unsigned)slice * 4u, &full[st], pol);
    } else {
      mbar_arrive(&full[st]);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    mbar_init(&ybar[0], 1);  // one local arrive.expect_tx per round; the peers' st.async complete the bytes
    mbar_init(&ybar[1], 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int q = 0; q < min(NS, Q); ++q) issue(q, q);  // prologue: fill the ring
  }
  __syncthreads();
  cluster_sync_all();  // peers' barriers are initialised before anyone posts to them

  bool valid[V];
  float4 xr[V], g[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int q4 = 4 * (tid + v * RP_CT);
    valid[v] = FULL || q4 < slice;
    g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    xr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE != RP_GEMV_C && valid[v]) {
      const int c = col0 + q4;
      if (c + 3 < p.nf) xr[v] = *reinterpret_cast<const float4*>(p.x + c);
      else {
        xr[v].x = c < p.nf ? p.x[c] : 0.f;
        xr[v].y = c + 1 < p.nf ? p.x[c + 1] : 0.f;
        xr[v].z = c + 2 < p.nf ? p.x[c + 2] : 0.f;
      }
      if (MODE == RP_NORMAL && p.xold) {
        // same individually rounded operations as fista_momentum_kernel, so the operand is bit-identical to the
        // vector the epilogue kernel forms for itself
        float4 xo = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c + 3 < p.nf) xo = *reinterpret_cast<const float4*>(p.xold + c);
        else {
          xo.x = c < p.nf ? p.xold[c] : 0.f;
          xo.y = c + 1 < p.nf ? p.xold[c + 1] : 0.f;
          xo.z = c + 2 < p.nf ? p.xold[c + 2] : 0.f;
        }
        const float tho = *p.th_old, thn = *p.th;
        const float c1 = fdiv(fsub(1.f, tho), thn), c2 = fadd(fdiv(fsub(tho, 1.f), thn), 1.f);
        xr[v].x = fadd(fmul(xr[v].x, c1), fmul(xo.x, c2));
        xr[v].y = fadd(fmul(xr[v].y, c1), fmul(xo.y, c2));
        xr[v].z = fadd(fmul(xr[v].z, c1), fmul(xo.z, c2));
        xr[v].w = fadd(fmul(xr[v].w, c1), fmul(xo.w, c2));
      }
    }
  }
  int s = 0, q0 = 0;
  unsigned phase = 0, par = 0, ypar = 0;
  for (int i = 0; i < nr; ++i) {
    const int rv = (i == nr - 1) ? rv_last : R;
    const int64_t row0 = ((int64_t)cid + (int64_t)i * ncl) * R;
    const int s0 = s;
    float4 a[R][V];
    float acc[R][FPE];
    float y[RF];
    if (MODE == RP_GEMV_C) {
#pragma unroll
      for (int k = 0; k < RF; ++k) y[k] = (k / FPE) < rv ? __ldg(p.yin + row0 * FPE + k) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int e = 0; e < FPE; ++e) acc[r][e] = 0.f;
      if (r < rv) {
        mbar_wait<false>(&full[s], phase, s_abort, p.abort_flag);
        const float4* st = reinterpret_cast<const float4*>(ring + (size_t)s * W);
#pragma unroll
        for (int v = 0; v < V; ++v) a[r][v] = valid[v] ? st[tid + v * RP_CT] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (++s == NS) { s = 0; phase ^= 1u; }
        if (MODE != RP_GEMV_C) {
#pragma unroll
          for (int v = 0; v < V; ++v) dot_acc<FPE>(acc[r], a[r][v], xr[v]);
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) a[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (MODE != RP_GEMV_C) {
      // warp -> CTA -> cluster reduction of the RF partial dot products, fixed order throughout
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int e = 0; e < FPE; ++e) {
          const float w = warp_sum(acc[r][e]);
          if (lane == 0) wpart[warp * RF + r * FPE + e] = w;
        }
    }
    // every warp has moved this round's rows ring -> registers (the dot products / the CSE-proof
    // asm below consumed them): the stages are free, one thread re-arms them with the rows NS ahead
    if (MODE == RP_GEMV_C) {
      // no dot product here: make the barrier wait for the loads with a (never taken) dependent store
      unsigned u = 0u;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int v = 0; v < V; ++v) u |= __float_as_uint(a[r][v].w) ^ __float_as_uint(a[r][v].x);
      if (u == 0x7fc12345u) ysm[0] = 1.f;
    }
    __syncthreads();
    if (tid == 0) {
      int st = s0;
      for (int r = 0; r < rv; ++r) {
        if (q0 + r + NS < Q) issue(q0 + r + NS, st);
        if (++st == NS) st = 0;
      }
    }
    q0 += rv;
    if (MODE != RP_GEMV_C) {
      if (warp == 0) {
        float pv = 0.f;
        if (lane < RF) {
#pragma unroll
          for (int w = 0; w < RP_CW; ++w) pv += wpart[w * RF + lane];
        }
        float tot = pv;
        if (G > 1) {
          constexpr int PER = 32 / RF;  // peers served per pass over the warp
          const int k = lane % RF;
          const float pk = __shfl_sync(0xffffffffu, pv, k);
          if (lane == 0) mbar_expect_tx(&ybar[par], G * RF * 4u);
          if (lane < PER * RF)
            for (unsigned peer = lane / RF; peer < G; peer += PER)
              dsmem_post(&ybuf[(par * RP_MAXG + rank) * RF + k], &ybar[par], peer, pk);
          mbar_wait<false>(&ybar[par], ypar, s_abort, p.abort_flag);
          if (lane < RF) {
            // four interleaved partial sums (independent loads), combined in a fixed order
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
            const float* yb = ybuf + (par * RP_MAXG) * RF + lane;
            for (unsigned c = 0; c < G; c += 4) {
              t0 += yb[c * RF];
              if (c + 1 < G) t1 += yb[(c + 1) * RF];
              if (c + 2 < G) t2 += yb[(c + 2) * RF];
              if (c + 3 < G) t3 += yb[(c + 3) * RF];
            }
            tot = (t0 + t1) + (t2 + t3);
          }
        }
        if (lane < RF) {
          ysm[lane] = tot;
          if (MODE == RP_GEMV_N && rank == 0 && (lane / FPE) < rv) p.yout[row0 * FPE + lane] = tot;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < RF; ++k) y[k] = ysm[k];
      par ^= 1u;
      if (par == 0) ypar ^= 1u;
    }
    if (MODE != RP_GEMV_N) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int v = 0; v < V; ++v) axpy_conj<FPE>(g[v], a[r][v], y[r * FPE], y[r * FPE + FPE - 1]);
    }
  }
  if (MODE != RP_GEMV_N) {
    float* out = p.gpart + (size_t)cid * p.gstride + col0;
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (valid[v]) *reinterpret_cast<float4*>(out + 4 * (tid + v * RP_CT)) = g[v];
  }
  __syncwarp();
  cluster_sync_all();  // nobody leaves while a peer may still post into its shared memory
}

// res[j] = sum over clusters of gpart[k][j], fixed order
__global__ void __launch_bounds__(256) rowpass_finish_kernel(const float* __restrict__ gpart, int64_t gstride, int ncl, int nf,
                                                            float* __restrict__ res, const int* gate) {
  pdl_prologue();
  if (gate && *gate) return;
  const int nf4 = nf >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf4; i += gridDim.x * blockDim.x) {
    float4 s = __ldcg(reinterpret_cast<const float4*>(gpart) + i);
    for (int k = 1; k < ncl; ++k) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(gpart + (size_t)k * gstride) + i);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4*>(res)[i] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < (nf & 3)) {
    const int j = (nf4 << 2) + threadIdx.x;
    float s = 0.f;
    for (int k = 0; k < ncl; ++k) s += __ldcg(gpart + (size_t)k * gstride + j);
    res[j] = s;
  }
}

typedef void (*rowpass_fn)(RowpassArgs);
template <int FPE, int V, int R, bool FULL>
rowpass_fn pick_mode(int mode) {
  switch (mode) {
    case RP_NORMAL: return rowpass_kernel<FPE, V, R, RP_NORMAL, FULL>;
    case RP_GEMV_N: return rowpass_kernel<FPE, V, R, RP_GEMV_N, FULL>;
    default: return rowpass_kernel<FPE, V, R, RP_GEMV_C, FULL>;
  }
}
template <int FPE, int V>
rowpass_fn pick_r(int R, int mode, bool full) {
  if (full) {
    switch (R) {
      case 2: return pick_mode<FPE, V, 2, true>(mode);
      case 3: return pick_mode<FPE, V, 3, true>(mode);
      default: return pick_mode<FPE, V, 4, true>(mode);
    }
  }
  switch (R) {
    case 2: return pick_mode<FPE, V, 2, false>(mode);
    case 3: return pick_mode<FPE, V, 3, false>(mode);
    default: return pick_mode<FPE, V, 4, false>(mode);
  }
}
template <int FPE>
rowpass_fn pick_v(int V, int R, int mode, bool full) {
  switch (V) {
    case 1: return pick_r<FPE, 1>(R, mode, full);
    case 2: return pick_r<FPE, 2>(R, mode, full);
    case 3: return pick_r<FPE, 3>(R, mode, full);
    case 4: return pick_r<FPE, 4>(R, mode, full);
    // wide slices (cluster sizes that are not powers of two, e.g. 6 CTAs x 10924 floats): two rows per round
    case 5: return pick_mode<FPE, 5, 2, false>(mode);
    case 6: return pick_mode<FPE, 6, 2, false>(mode);
    // 16384-float slices: one row per round (a, x and g at 32 registers each), half the cluster size
    case 7: return pick_mode<FPE, 7, 1, false>(mode);
    default: return full ? pick_mode<FPE, 8, 1, true>(mode) : pick_mode<FPE, 8, 1, false>(mode);
  }
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct RowPlan {
  rls_ctx_s* ctx = nullptr;
  rls_mat_s* A = nullptr;
  int fpe = 1, G = 1, V = 4, W = 0, NS = 0, R = 4, ncl = 0;
  size_t smem = 0;
  rowpass_fn fn[3] = {nullptr, nullptr, nullptr};
  float* gpart = nullptr;
  int64_t gstride = 0;
  int* abort_flag = nullptr;
};

void rls_rowpass_plan_destroy(RowPlan* p) {
  if (!p) return;
  if (p->gpart) cudaFree(p->gpart);
  if (p->abort_flag) cudaFree(p->abort_flag);
  delete p;
}

static size_t rowpass_smem(int NS, int W, int fpe) {
  const int RF = RP_MAXR * fpe;
  return (size_t)NS * W * 4 + (size_t)(NS + 2) * 8 + (size_t)(RP_CW * RF + 2 * RP_MAXG * RF + RF) * 4 + 16;
}

int32_t rls_rowpass_plan_create(rls_ctx_s* c, rls_mat_s* A, RowPlan** out) {
  *out = nullptr;
  if (A->layout != RLS_LAYOUT_ROWMAJOR) { rls_set_error("rowpass needs a row-major matrix"); return RLS_ERR_INVALID; }
  const int fpe = A->dtype == RLS_C32 ? 2 : 1;
  const int64_t nf64 = A->n * fpe;
  const int64_t nf_pad = (nf64 + 3) & ~(int64_t)3;
  if (((uintptr_t)A->d % 16) != 0 || (A->ld * fpe) % 4 != 0) { rls_set_error("rowpass needs 16-byte aligned rows"); return RLS_ERR_UNSUPPORTED; }
  if (nf_pad > (int64_t)RP_MAXG * 4 * 2048) {
    rls_set_error("rowpass supports rows of at most %d floats, got %lld", RP_MAXG * 4 * 2048, (long long)nf64);
    return RLS_ERR_UNSUPPORTED;
  }
  RowPlan* p = new RowPlan();
  p->ctx = c; p->A = A; p->fpe = fpe;
  int G = 1;
  while ((int64_t)G * 4 * 2048 < nf_pad) G *= 2;
  {
    // cluster sizes need not be powers of two: what counts is how many SMs the co-resident clusters cover
    // (GPCs of 20/18/14 SMs: 8 -> 120 SMs, 6 -> 138, 4 -> 132, 2 and 1 -> 148) against the slice width
    const int g_env = env_int("RLS_ROWPASS_G", 0);
    if (g_env > 0 && (int64_t)g_env * 8 * 2048 >= nf_pad) G = g_env;
  }
  if (G > RP_MAXG) G = RP_MAXG;
  int W = (int)(((nf_pad + G - 1) / G + 3) & ~(int64_t)3);
  if (W < 4) W = 4;
  int V = (W + 2047) / 2048;
  if (V > 8) { rls_set_error("rowpass: slice of %d floats too wide", W); rls_rowpass_plan_destroy(p); return RLS_ERR_UNSUPPORTED; }
  int R = env_int("RLS_ROWPASS_R", fpe == 2 ? 3 : 4);  // measured on B200 (profiles/r01_rowpass_sweep.txt)
  if (R < 1) R = 1;
  if (R > RP_MAXR) R = RP_MAXR;
  if (V > 4) R = 2;
  if (V > 6) R = 1;
  p->G = G; p->W = W; p->V = V; p->R = R;
  // every CTA's slice is exactly V*2048 floats: no column predicates in the kernel
  const bool full = (W == V * 4 * RP_CT) && ((int64_t)G * W == nf_pad);
  for (int mode = 0; mode < 3; ++mode) p->fn[mode] = fpe == 2 ? pick_v<2>(V, R, mode, full) : pick_v<1>(V, R, mode, full);
  // ring: as many stages as fit in the 227 KB of shared memory (at least 2)
  int dev_max = 0;
  cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
  int NS = env_int("RLS_ROWPASS_NS", 0);
  if (NS <= 0) {
    NS = 2;
    while (NS < 32 && rowpass_smem(NS + 1, W, fpe) <= (size_t)dev_max) ++NS;
    // small slices: no point in more than ~160 KB in flight
    while (NS > 8 && (size_t)(NS - 1) * W * 4 >= (size_t)200 * 1024) --NS;
  }
  if (NS < R) NS = R;  // a round's rows are resident together
  p->NS = NS;
  p->smem = rowpass_smem(NS, W, fpe);
  if (p->smem > (size_t)dev_max) { rls_set_error("rowpass: %zu B shared memory needed, device allows %d", p->smem, dev_max); rls_rowpass_plan_destroy(p); return RLS_ERR_UNSUPPORTED; }
  for (int mode = 0; mode < 3; ++mode) {
    cudaError_t e = cudaFuncSetAttribute((const void*)p->fn[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    if (e == cudaSuccess && G > 8) e = cudaFuncSetAttribute((const void*)p->fn[mode], cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { rls_set_error("rowpass: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); rls_rowpass_plan_destroy(p); return RLS_ERR_CUDA; }
  }
  // how many clusters are co-resident (persistent grid = exactly that many)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G * 1024);
  cfg.blockDim = dim3(RP_THREADS);
  cfg.dynamicSmemBytes = p->smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, (const void*)p->fn[0], &cfg);
  if (e != cudaSuccess || ncl < 1) {
    rls_set_error("rowpass: no co-resident cluster of %d CTAs (%s)", G, cudaGetErrorString(e));
    cudaGetLastError();
    rls_rowpass_plan_destroy(p);
    return RLS_ERR_UNSUPPORTED;
  }
  const int cap = env_int("RLS_ROWPASS_CLUSTERS", 0);
  if (cap > 0 && cap < ncl) ncl = cap;
  const int ncl_max = ncl;
  const int64_t nrounds = (A->m + R - 1) / R;
  if (nrounds < ncl) ncl = (int)std::max<int64_t>(nrounds, 1);
  p->ncl = ncl;
  const int ncl_alloc = std::max(ncl, 1);
  (void)ncl_max;
  p->gstride = ((int64_t)G * W + 63) & ~(int64_t)63;
  if (cudaMalloc(&p->gpart, (size_t)ncl_alloc * p->gstride * 4) != cudaSuccess || cudaMalloc(&p->abort_flag, 4) != cudaSuccess) {
    rls_set_error("rowpass: out of device memory for the cluster partials");
    cudaGetLastError();
    rls_rowpass_plan_destroy(p);
    return RLS_ERR_NOMEM;
  }
  cudaMemsetAsync(p->abort_flag, 0, 4, c->stream);
  *out = p;
  return RLS_OK;
}

static int32_t rowpass_launch(RowPlan* p, int mode, const void* x, const void* yin, void* yout, void* res, const int* gate,
                              const float* xold = nullptr, const float* th_old = nullptr, const float* th = nullptr, bool defer_finish = false) {
  rls_ctx_s* c = p->ctx;
  rls_mat_s* A = p->A;
  RowpassArgs a;
  a.xold = xold; a.th_old = th_old; a.th = th;
  a.A = (const float*)A->d; a.ldf = A->ld * p->fpe; a.m = A->m; a.nf = (int)(A->n * p->fpe);
  a.W = p->W; a.NS = p->NS; a.R = p->R;
  a.x = (const float*)x; a.yin = (const float*)yin; a.yout = (float*)yout;
  a.gpart = p->gpart; a.gstride = p->gstride; a.gate = gate; a.abort_flag = p->abort_flag;
  const int ncl = p->ncl;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncl * p->G);
  cfg.blockDim = dim3(RP_THREADS);
  cfg.dynamicSmemBytes = p->smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p->G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = rls_pdl_enabled() ? 2 : 1;
  rls_trace_begin(c->stream, mode == RP_NORMAL ? "rowpass normal" : mode == RP_GEMV_N ? "rowpass gemv_n" : "rowpass gemv_c");
  RLS_CUDA(cudaLaunchKernelEx(&cfg, p->fn[mode], a));
  rls_trace_end(c->stream);
  c->launches++;
  if (mode != RP_GEMV_N && !defer_finish) {
    const int nf = a.nf;
    int grid = std::min(c->sm_count * 2, std::max(1, (nf / 4 + 255) / 256));
    RLS_CUDA(rls_launch_pdl(c->stream, dim3(grid), dim3(256), rowpass_finish_kernel, (const float*)p->gpart, p->gstride, ncl, nf, (float*)res, gate));
    c->launches++;
  }
  return RLS_OK;
}

int32_t rls_rowpass_normal(RowPlan* p, const void* x, void* res, const int* gate) { return rowpass_launch(p, RP_NORMAL, x, nullptr, nullptr, res, gate); }
// the cluster kernel only: the caller's epilogue kernel sums the per-cluster partials itself (fixed order, the same
// arithmetic as rowpass_finish_kernel) — one kernel boundary less per iteration.  Optional fused FISTA momentum.
int32_t rls_rowpass_normal_deferred(RowPlan* p, const void* x, const float* xold, const float* th_old, const float* th, const int* gate,
                                    const float** gpart, int64_t* gstride, int* ncl) {
  RLS_TRY(rowpass_launch(p, RP_NORMAL, x, nullptr, nullptr, nullptr, gate, xold, th_old, th, true));
  *gpart = p->gpart; *gstride = p->gstride; *ncl = p->ncl;
  return RLS_OK;
}
int32_t rls_rowpass_gemv_n(RowPlan* p, const void* x, void* y, const int* gate) { return rowpass_launch(p, RP_GEMV_N, x, nullptr, y, nullptr, gate); }
int32_t rls_rowpass_gemv_c(RowPlan* p, const void* y, void* g, const int* gate) { return rowpass_launch(p, RP_GEMV_C, nullptr, y, nullptr, g, gate); }

int32_t rls_rowpass_check_abort(RowPlan* p) {
  int flag = 0;
  RLS_CUDA(cudaMemcpyAsync(&flag, p->abort_flag, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (flag) {
    rls_set_error("row-major one-pass kernel timed out on a barrier (abort flag set)");
    return RLS_ERR_CUDA;
  }
  return RLS_OK;
}

void rls_rowpass_describe(RowPlan* p, char* buf, int len) {
  snprintf(buf, len, "onepass/rowmajor: clusters=%d x %d CTAs (%d SMs) slice=%d floats (V=%d) rows/round=%d ring=%d x %d B smem=%zu B",
           p->ncl, p->G, p->ncl * p->G, p->W, p->V, p->R, p->NS, p->W * 4, p->smem);
}
