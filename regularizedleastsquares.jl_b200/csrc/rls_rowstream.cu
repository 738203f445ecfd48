// rls_rowstream.cu — one-HBM-pass normal operator g = A'(A x) (and y = A x, g = A' y) on a ROW-MAJOR
// device copy of A: warp-specialised, per-row software pipeline.
//
// Replaces mul!(res, AHA, x) (src/FISTA.jl:152, src/POGM.jl:181, src/OptISTA.jl:182, src/CGNR.jl:151,
// cg! inside src/ADMM.jl:244) and mul!(x0, adjoint(A), b) (src/FISTA.jl:114, src/CGNR.jl:132).
//
// A'(A x) = sum_i conj(a_i) (a_i . x) over the ROWS a_i.  A group of G CTAs (= one thread-block cluster,
// G = ceil(row / 8192 floats) <= 16) owns whole rows; CTA `rank` owns the column slice [rank*W, rank*W+W)
// of every row of its group, its x values and its g accumulators stay in registers for the whole launch.
// A crosses HBM exactly once; the only exchange is the partial dot product of each row between the CTAs
// of ONE cluster through distributed shared memory.
//
// Round 1's kernel (rls_rowpass.cu, removed; numbers in profiles/r01_*) made all 16 warps stand at a CTA barrier while warp 0 ran that
// exchange once per round of rows (37 % of the stall samples for ComplexF32 rows of 65536 elements).
// Here nobody waits for an exchange that was started in the same iteration:
//   16 compute warps, iteration i:  row i ring -> registers, partial dot product, warp sum -> shared memory;
//                                   then fold row i-1 (still in registers) g += conj(a) y_{i-1}.
//   service warp A (per row):       waits until all 16 warps hold the row in registers, adds their partials
//                                   (fixed tree), posts the CTA's sum into the shared memory of ALL CTAs of the
//                                   cluster (one st.async per lane = (peer, component): data and mbarrier
//                                   signal in one message) and re-arms the ring stage with the row NS ahead
//                                   (bulk copy global -> shared, evict-first).
//   service warp B (per row):       waits for the posted numbers, adds them in a fixed order (every CTA of the
//                                   cluster obtains the bit-identical y) and publishes y for the fold.
// The chain partials -> A -> DSMEM -> B -> y (~0.4 us) runs beside the compute warps' next row.  Measured dead ends
// (profiles/r02_rowstream_design.txt): every compute warp posting its own partial to the peers (16 x G x components
// messages per row saturate the receiving mbarrier for G = 16); every compute warp adding the posted numbers itself
// (no service warp B, but six dependent shuffles in each of their iterations); deeper rings (4 x 32 KB per SM is
// the optimum: more outstanding bulk copies lower the DRAM efficiency).
// 20 warps (5 per scheduler) leave 96 registers per thread: 2 rows x 4 float4 of A + x + g = 64.
// Every sum has a fixed order: results are deterministic run to run and identical in all CTAs.
// Accuracy: a thread folds up to tens of thousands of rows into its g accumulators (262144 x 65536 on one GPU: 37449 rows
// per cluster); a plain Float32 chain of that length loses ~sqrt(rows) ulps.  Every RS_FLUSH rows the Float32 accumulators
// are therefore flushed into a Float64 copy of the slice in shared memory (owner-thread only: no synchronisation), and
// the per-cluster partials are added in Float64 by the finish / epilogue kernels, so g carries the rounding of a 64-term
// Float32 sum, not of a 37449-term one.
//
// What did not work (profiles/r02_rowstream_design.txt): exchanging through L2 with {tag,value} words so that
// groups need not be clusters and 144-148 SMs stream — an L2 round trip under a saturated HBM stream costs
// ~1 us per hop, more than the three rows a CTA can hold back in registers; and letting compute warps 0/1
// double as service warps — their longer iteration became everybody's iteration (1540 cycles per row).
#include "rls_common.cuh"
#include "rls_async.cuh"

#include <algorithm>

using namespace rls_async;

namespace {

constexpr int RS_CW = 16;                    // compute warps
constexpr int RS_SW = 4;                     // service warps (two used; four keep the schedulers balanced)
constexpr int RS_CT = RS_CW * 32;            // compute threads
constexpr int RS_THREADS = (RS_CW + RS_SW) * 32;
constexpr int RS_MAXG = 16;                  // CTAs per group (G * FPE <= 32: one lane per posted word)
constexpr int RS_NSLOT = 4;                  // exchange slots (rows in flight between the CTAs of a cluster: <= 3)
constexpr int RS_XROW = RS_MAXG * 2;         // floats per exchange slot: [component][rank]
constexpr int RS_FLUSH = 64;                 // rows folded into the Float32 accumulators between two flushes into Float64
enum { RS_NORMAL = 0, RS_GEMV_N = 1, RS_GEMV_C = 2 };

struct RowstreamArgs {
  const float* A;     // row-major, row stride ldf floats
  int64_t ldf;
  int64_t m;
  int nf;             // floats per row (n * FPE)
  int W;              // slice width in floats (multiple of 4)
  int NS;             // ring stages
  int G;              // CTAs per group (= cluster size)
  int NG;             // groups
  const float* x;     // n-vector (NORMAL, GEMV_N)
  const float* xold;  // fused FISTA momentum (FISTA.jl:144-148): the operator is applied to x*c1 + xold*c2 with
  const float* th_old;//   c1 = (1-θold)/θ, c2 = (θold-1)/θ + 1 formed from the device-resident θ's; NULL = plain x
  const float* th;
  const float* yin;   // m-vector (GEMV_C)
  float* yout;        // m-vector (GEMV_N)
  float* gpart;       // [NG][gstride] (NORMAL, GEMV_C)
  int64_t gstride;
  const int* gate;
  int* abort_flag;
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Deposit one float in CTA `peer`'s copy of a shared slot and signal its mbarrier in the SAME message (st.async +
// complete_tx): no release fence on the sender, no L1 invalidate on the receiver.
__device__ __forceinline__ void dsmem_post(float* local_slot, uint64_t* local_bar, unsigned peer, float v) {
  uint32_t rs, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rs) : "r"(smem_u32(local_slot)), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(peer));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rs), "r"(__float_as_uint(v)), "r"(rb) : "memory");
}

template <int FPE> __device__ __forceinline__ void dot_acc(float (&acc)[FPE], float4 a, float4 x);
template <> __device__ __forceinline__ void dot_acc<1>(float (&acc)[1], float4 a, float4 x) {
  acc[0] = fmaf(a.x, x.x, acc[0]); acc[0] = fmaf(a.y, x.y, acc[0]); acc[0] = fmaf(a.z, x.z, acc[0]); acc[0] = fmaf(a.w, x.w, acc[0]);
}
template <> __device__ __forceinline__ void dot_acc<2>(float (&acc)[2], float4 a, float4 x) {
  acc[0] = fmaf(a.x, x.x, acc[0]); acc[0] = fmaf(-a.y, x.y, acc[0]); acc[1] = fmaf(a.x, x.y, acc[1]); acc[1] = fmaf(a.y, x.x, acc[1]);
  acc[0] = fmaf(a.z, x.z, acc[0]); acc[0] = fmaf(-a.w, x.w, acc[0]); acc[1] = fmaf(a.z, x.w, acc[1]); acc[1] = fmaf(a.w, x.z, acc[1]);
}
template <int FPE> __device__ __forceinline__ void axpy_conj(float4& g, float4 a, float yr, float yi);
template <> __device__ __forceinline__ void axpy_conj<1>(float4& g, float4 a, float yr, float) {
  g.x = fmaf(a.x, yr, g.x); g.y = fmaf(a.y, yr, g.y); g.z = fmaf(a.z, yr, g.z); g.w = fmaf(a.w, yr, g.w);
}
template <> __device__ __forceinline__ void axpy_conj<2>(float4& g, float4 a, float yr, float yi) {
  g.x = fmaf(a.x, yr, g.x); g.x = fmaf(a.y, yi, g.x); g.y = fmaf(a.x, yi, g.y); g.y = fmaf(-a.y, yr, g.y);
  g.z = fmaf(a.z, yr, g.z); g.z = fmaf(a.w, yi, g.z); g.w = fmaf(a.z, yi, g.w); g.w = fmaf(-a.w, yr, g.w);
}

template <int FPE, int V, int MODE, bool FULL>
__global__ void __launch_bounds__(RS_THREADS, 1) rowstream_kernel(RowstreamArgs p) {
  pdl_prologue();
  if (p.gate && *p.gate) return;  // device-side done() gate: uniform over the grid
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = p.G, NG = p.NG, W = p.W, NS = p.NS;
  const int grp = blockIdx.x / G, rank = blockIdx.x - grp * G;  // 1-D clusters are G consecutive blocks

  float* ring = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * W * 4);  // [NS] bulk copy landed
  uint64_t* ebar = full + NS;                                               // [NS] stage read by all compute warps (GEMV_C)
  uint64_t* wbar = ebar + NS;                                               // [2]  row held in registers by all compute warps
  uint64_t* ybar = wbar + 2;                                                // [2]  y of a row published
  uint64_t* xbar = ybar + 2;                                                // [RS_NSLOT] the cluster's partials of a row arrived
  float* xbuf = reinterpret_cast<float*>(xbar + RS_NSLOT);                  // [RS_NSLOT][RS_XROW]  (16-byte aligned)
  float* wpart = xbuf + RS_NSLOT * RS_XROW;                                 // [2][component][16 warps] (two-stage exchange)
  float* ysm = wpart + 2 * 32;                                              // [2][2]
  double* g2 = reinterpret_cast<double*>(ysm + 2 * 2 + 4);                   // [W] Float64 second-level accumulators (8-byte aligned)
  volatile int* s_abort = reinterpret_cast<volatile int*>(g2 + W);

  const int nf_pad = (p.nf + 3) & ~3;
  const int col0 = rank * W;
  const int slice = max(0, min(W, nf_pad - col0));  // floats of this CTA's slice (multiple of 4)
  const int Q = (int)((p.m - grp + NG - 1) / NG);   // rows of this group: grp, grp + NG, ...

  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  auto issue = [&](int q, int st) {  // bulk copy of this CTA's slice of the group's q-th row into ring stage st
    if (slice > 0) {
      const int64_t row = (int64_t)grp + (int64_t)q * NG;
      mbar_expect_tx(&full[st], (unsigned)slice * 4u);
      bulk_load(ring + (size_t)st * W, p.A + row * p.ldf + col0, (unsigned)slice * 4u, &full[st], pol);
    } else {
      mbar_arrive(&full[st]);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&ebar[s], RS_CW); }
    for (int d = 0; d < 2; ++d) { mbar_init(&wbar[d], RS_CW); mbar_init(&ybar[d], 1); }
    for (int k = 0; k < RS_NSLOT; ++k) mbar_init(&xbar[k], 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int q = 0; q < min(NS, Q); ++q) issue(q, q);  // prologue: fill the ring
  }
  for (int k = tid; k < RS_NSLOT * RS_XROW; k += RS_THREADS) xbuf[k] = 0.f;  // entries nobody posts to must read as zero
  __syncthreads();
  cluster_sync_all();  // the peers' barriers are initialised before anyone posts to them

  if (warp < RS_CW) {
    // =========================== compute warps ===========================
    bool valid[V];
    float4 xr[V], g[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int q4 = 4 * (tid + v * RS_CT);
      valid[v] = FULL || q4 < slice;
      g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE != RS_GEMV_N && valid[v]) {
        reinterpret_cast<double2*>(g2 + q4)[0] = make_double2(0.0, 0.0);
        reinterpret_cast<double2*>(g2 + q4)[1] = make_double2(0.0, 0.0);
      }
      xr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE != RS_GEMV_C && valid[v]) {
        const int c = col0 + q4;
        if (c + 3 < p.nf) xr[v] = *reinterpret_cast<const float4*>(p.x + c);
        else {
          xr[v].x = c < p.nf ? p.x[c] : 0.f;
          xr[v].y = c + 1 < p.nf ? p.x[c + 1] : 0.f;
          xr[v].z = c + 2 < p.nf ? p.x[c + 2] : 0.f;
        }
        if (MODE == RS_NORMAL && p.xold) {
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
    // second-level accumulation: g (Float32, at most RS_FLUSH rows) -> g2 (Float64, shared memory, this thread's entries)
    auto flush = [&]() {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (valid[v]) {
          double2* q = reinterpret_cast<double2*>(g2 + 4 * (tid + v * RS_CT));
          double2 d0 = q[0], d1 = q[1];
          d0.x += (double)g[v].x; d0.y += (double)g[v].y; d1.x += (double)g[v].z; d1.y += (double)g[v].w;
          q[0] = d0; q[1] = d1;
          g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    int nfold = 0;
    int s = 0;
    unsigned phase = 0;
    if (MODE == RS_GEMV_C) {
      // y is given: no exchange.  Stage -> registers -> fold; service warp A re-arms a stage once all warps have read it.
      for (int i = 0; i < Q; ++i) {
        const int64_t row = (int64_t)grp + (int64_t)i * NG;
        const float yr = __ldg(p.yin + row * FPE), yi = __ldg(p.yin + row * FPE + FPE - 1);
        mbar_wait(&full[s], phase, s_abort, p.abort_flag);
        const float4* st = reinterpret_cast<const float4*>(ring + (size_t)s * W);
        float4 av[V];
#pragma unroll
        for (int v = 0; v < V; ++v) av[v] = valid[v] ? st[tid + v * RS_CT] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int v = 0; v < V; ++v) axpy_conj<FPE>(g[v], av[v], yr, yi);  // consumes the loads before the arrive below
        __syncwarp();
        if (lane == 0) mbar_arrive(&ebar[s]);
        if (++s == NS) { s = 0; phase ^= 1u; }
        if (++nfold == RS_FLUSH) { flush(); nfold = 0; }
      }
    } else {
      float4 a[2][V];
      unsigned par = 0;  // parity of the use (i / 2) of the per-row barriers
      for (int i0 = 0; i0 < Q + 1; i0 += 2, par ^= 1u) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = i0 + u;
          if (i < Q) {
            // ---- row i: ring -> registers, partial dot product, hand over ----
            mbar_wait(&full[s], phase, s_abort, p.abort_flag);
            const float4* st = reinterpret_cast<const float4*>(ring + (size_t)s * W);
#pragma unroll
            for (int v = 0; v < V; ++v) a[u][v] = valid[v] ? st[tid + v * RS_CT] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (++s == NS) { s = 0; phase ^= 1u; }
            float accv[V][FPE];  // V independent chains, combined pairwise in a fixed order
#pragma unroll
            for (int v = 0; v < V; ++v) {
#pragma unroll
              for (int e = 0; e < FPE; ++e) accv[v][e] = 0.f;
              dot_acc<FPE>(accv[v], a[u][v], xr[v]);
            }
            float wsum[FPE];
#pragma unroll
            for (int e = 0; e < FPE; ++e) {
              float acc = accv[0][e];
              if (V == 2) acc = accv[0][e] + accv[1][e];
              if (V == 3) acc = (accv[0][e] + accv[1][e]) + accv[2][e];
              if (V == 4) acc = (accv[0][e] + accv[1][e]) + (accv[2][e] + accv[3][e]);
              wsum[e] = warp_sum(acc);  // butterfly: every lane holds the warp's partial
            }
            if (lane < FPE) wpart[u * 32 + lane * RS_CW + warp] = lane ? wsum[FPE - 1] : wsum[0];
            __syncwarp();
            if (lane == 0) mbar_arrive(&wbar[u]);  // partial written, row in registers: its ring stage may be re-armed
          }
          // ---- fold row i-1, still in registers ----
          const int j = i - 1;
          if (j >= 0 && j < Q) {
            const int uj = u ^ 1;
            const unsigned parj = (u == 1) ? par : (par ^ 1u);
            mbar_wait(&ybar[uj], parj, s_abort, p.abort_flag);
            if (MODE == RS_NORMAL) {
              const float yr = ysm[uj * 2], yi = ysm[uj * 2 + FPE - 1];
#pragma unroll
              for (int v = 0; v < V; ++v) axpy_conj<FPE>(g[v], a[uj][v], yr, yi);
              if (++nfold == RS_FLUSH) { flush(); nfold = 0; }
            }
          }
        }
      }
    }
    if (MODE != RS_GEMV_N) {
      float* out = p.gpart + (size_t)grp * p.gstride + col0;
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (valid[v]) {
          const double2* q = reinterpret_cast<const double2*>(g2 + 4 * (tid + v * RS_CT));
          const double2 d0 = q[0], d1 = q[1];
          *reinterpret_cast<float4*>(out + 4 * (tid + v * RS_CT)) =
              make_float4((float)(d0.x + (double)g[v].x), (float)(d0.y + (double)g[v].y), (float)(d1.x + (double)g[v].z), (float)(d1.y + (double)g[v].w));
        }
    }
  } else if (warp == RS_CW) {
    // =========================== service warp A ===========================
    int s = 0;
    unsigned phase = 0;
    if (MODE == RS_GEMV_C) {
      for (int i = 0; i < Q; ++i) {
        mbar_wait(&ebar[s], phase, s_abort, p.abort_flag);
        if (lane == 0 && i + NS < Q) issue(i + NS, s);
        if (++s == NS) { s = 0; phase ^= 1u; }
      }
    } else {
      for (int i = 0; i < Q; ++i) {
        const int u = i & 1;
        mbar_wait(&wbar[u], (unsigned)(i >> 1) & 1u, s_abort, p.abort_flag);  // all 16 warps hold row i in registers
        {
          // lane e adds the 16 warp partials of component e in a fixed tree, then lane (peer, e) posts the CTA's sum
          // into slot [e][rank] of that peer
          float t = 0.f;
          if (lane < FPE) {
            const float4* w4 = reinterpret_cast<const float4*>(wpart + u * 32 + lane * RS_CW);
            const float4 q0 = w4[0], q1 = w4[1], q2 = w4[2], q3 = w4[3];
            t = (((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w))) +
                (((q2.x + q2.y) + (q2.z + q2.w)) + ((q3.x + q3.y) + (q3.z + q3.w)));
          }
          const int slot = i & (RS_NSLOT - 1);
          if (G > 1) {
            t = __shfl_sync(0xffffffffu, t, lane % FPE);
            if (lane < G * FPE) dsmem_post(&xbuf[slot * RS_XROW + (lane % FPE) * RS_MAXG + rank], &xbar[slot], (unsigned)(lane / FPE), t);
          } else {
            // a cluster of one: plain shared-memory hand-over (st.async needs a peer CTA)
            if (lane < FPE) xbuf[slot * RS_XROW + lane * RS_MAXG] = t;
            __syncwarp();
            if (lane == 0) mbar_arrive(&xbar[slot]);
          }
        }
        if (lane == 0 && i + NS < Q) issue(i + NS, s);
        if (++s == NS) s = 0;
      }
    }
  } else if (warp == RS_CW + 1) {
    // =========================== service warp B ===========================
    if (MODE != RS_GEMV_C) {
      for (int r = 0; r < Q; ++r) {
        const int slot = r & (RS_NSLOT - 1), u = r & 1;
        if (G > 1 && lane == 0) mbar_expect_tx(&xbar[slot], (unsigned)(G * FPE) * 4u);   // G == 1: warp A arrives
        mbar_wait(&xbar[slot], (unsigned)(r / RS_NSLOT) & 1u, s_abort, p.abort_flag);
        float yr = 0.f;
        if (lane < FPE) {
          // [component][rank], 16 floats per component (ranks >= G stay zero): lane e adds its row in a fixed tree —
          // every CTA of the cluster forms the bit-identical y from the same numbers
          const float4* x4 = reinterpret_cast<const float4*>(xbuf + slot * RS_XROW + lane * RS_MAXG);
          const float4 q0 = x4[0], q1 = x4[1], q2 = x4[2], q3 = x4[3];
          yr = (((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w))) +
               (((q2.x + q2.y) + (q2.z + q2.w)) + ((q3.x + q3.y) + (q3.z + q3.w)));
        }
        if (lane < FPE) {
          ysm[u * 2 + lane] = yr;
          if (MODE == RS_GEMV_N && rank == 0) p.yout[((int64_t)grp + (int64_t)r * NG) * FPE + lane] = yr;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ybar[u]);
      }
    }
  }
  __syncwarp();
  cluster_sync_all();  // nobody leaves while a peer may still post into its shared memory
}

// res[j] = sum over the groups of gpart[k][j], fixed order, accumulated in Float64
__global__ void __launch_bounds__(256) rowpass_finish_kernel(const float* __restrict__ gpart, int64_t gstride, int ncl, int nf,
                                                            float* __restrict__ res, const int* gate) {
  pdl_prologue();
  if (gate && *gate) return;
  const int nf4 = nf >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf4; i += gridDim.x * blockDim.x) {
    double sx = 0.0, sy = 0.0, sz = 0.0, sw = 0.0;   // Float64: the sum of the partials adds no rounding of its own
    for (int k = 0; k < ncl; ++k) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(gpart + (size_t)k * gstride) + i);
      sx += (double)t.x; sy += (double)t.y; sz += (double)t.z; sw += (double)t.w;
    }
    reinterpret_cast<float4*>(res)[i] = make_float4((float)sx, (float)sy, (float)sz, (float)sw);
  }
  if (blockIdx.x == 0 && threadIdx.x < (nf & 3)) {
    const int j = (nf4 << 2) + threadIdx.x;
    double s = 0.0;
    for (int k = 0; k < ncl; ++k) s += (double)__ldcg(gpart + (size_t)k * gstride + j);
    res[j] = (float)s;
  }
}

typedef void (*rowstream_fn)(RowstreamArgs);
template <int FPE, int V, bool FULL>
rowstream_fn pick_mode(int mode) {
  switch (mode) {
    case RS_NORMAL: return rowstream_kernel<FPE, V, RS_NORMAL, FULL>;
    case RS_GEMV_N: return rowstream_kernel<FPE, V, RS_GEMV_N, FULL>;
    default: return rowstream_kernel<FPE, V, RS_GEMV_C, FULL>;
  }
}
template <int FPE>
rowstream_fn pick(int V, int mode, bool full) {
  if (V == 4 && full) return pick_mode<FPE, 4, true>(mode);
  switch (V) {
    case 1: return pick_mode<FPE, 1, false>(mode);
    case 2: return pick_mode<FPE, 2, false>(mode);
    case 3: return pick_mode<FPE, 3, false>(mode);
    default: return pick_mode<FPE, 4, false>(mode);
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
  int fpe = 1, G = 1, NG = 1, V = 4, W = 0, NS = 0;
  size_t smem = 0;
  rowstream_fn fn[3] = {nullptr, nullptr, nullptr};
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

static size_t rowstream_smem(int NS, int W) {
  const size_t b = (size_t)NS * W * 4 + (size_t)(2 * NS + 4 + RS_NSLOT) * 8;  // ring, barriers (even count: 16-byte aligned)
  return b + (size_t)(RS_NSLOT * RS_XROW + 2 * 32 + 2 * 2 + 4) * 4 + (size_t)W * 8 + 16;
}

int32_t rls_rowpass_plan_create(rls_ctx_s* c, rls_mat_s* A, RowPlan** out) {
  *out = nullptr;
  if (A->layout != RLS_LAYOUT_ROWMAJOR) { rls_set_error("rowpass needs a row-major matrix"); return RLS_ERR_INVALID; }
  const int fpe = A->dtype == RLS_C32 ? 2 : 1;
  const int64_t nf64 = A->n * fpe;
  const int64_t nf_pad = (nf64 + 3) & ~(int64_t)3;
  if (((uintptr_t)A->d % 16) != 0 || (A->ld * fpe) % 4 != 0) { rls_set_error("rowpass needs 16-byte aligned rows"); return RLS_ERR_UNSUPPORTED; }
  if (nf_pad > (int64_t)RS_MAXG * 8192) {
    rls_set_error("rowpass supports rows of at most %d floats, got %lld", RS_MAXG * 8192, (long long)nf64);
    return RLS_ERR_UNSUPPORTED;
  }
  RowPlan* p = new RowPlan();
  p->ctx = c; p->A = A; p->fpe = fpe;
  int G = (int)((nf_pad + 8191) / 8192);
  if (G < 1) G = 1;
  {
    const int g_env = env_int("RLS_ROWSTREAM_G", 0);
    if (g_env >= G && g_env <= RS_MAXG) G = g_env;
  }
  int W = (int)(((nf_pad + G - 1) / G + 3) & ~(int64_t)3);
  if (W < 4) W = 4;
  const int V = (W + 4 * RS_CT - 1) / (4 * RS_CT);
  const bool full = (W == V * 4 * RS_CT) && ((int64_t)G * W == nf_pad);
  p->G = G; p->W = W; p->V = V;
  for (int mode = 0; mode < 3; ++mode) p->fn[mode] = fpe == 2 ? pick<2>(V, mode, full) : pick<1>(V, mode, full);
  int dev_max = 0;
  cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
  int NS = env_int("RLS_ROWSTREAM_NS", 0);
  if (NS <= 0) {
    // measured on B200 (profiles/r02_rowstream_sweep.txt): ~128 KB of bulk copies in flight per SM is the optimum —
    // 4 stages of 32 KB; deeper rings lower the DRAM efficiency (6.4 instead of 7.1 TB/s at 7 stages on 148 SMs)
    NS = (int)((128 * 1024 + (size_t)W * 4 - 1) / ((size_t)W * 4));
    if (NS < 4) NS = 4;
    if (NS > 32) NS = 32;
  }
  while (NS > 2 && rowstream_smem(NS, W) > (size_t)dev_max) --NS;
  if (NS < 2) NS = 2;
  p->NS = NS;
  p->smem = rowstream_smem(NS, W);
  if (p->smem > (size_t)dev_max) { rls_set_error("rowpass: %zu B shared memory needed, device allows %d", p->smem, dev_max); rls_rowpass_plan_destroy(p); return RLS_ERR_UNSUPPORTED; }
  // persistent grid: the CTAs of a group wait for each other, so a whole group must be resident — the cluster launch
  // guarantees exactly that; as many groups as clusters are co-resident (GPCs of 20/18/14 SMs: 8-CTA clusters cover
  // 120 SMs, 16-CTA clusters 112, 4 -> 132, 2 and 1 -> 148)
  int sms = c->sm_count;
  const int cap = env_int("RLS_ROWSTREAM_SMS", 0);
  if (cap > 0 && cap < sms) sms = cap;
  int NG = sms / G;
  cudaError_t e = cudaSuccess;
  for (int mode = 0; mode < 3 && e == cudaSuccess; ++mode) {
    e = cudaFuncSetAttribute((const void*)p->fn[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    if (e == cudaSuccess && G > 8) e = cudaFuncSetAttribute((const void*)p->fn[mode], cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int k = 0;
    if (e == cudaSuccess && G > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(G * 1024);
      cfg.blockDim = dim3(RS_THREADS);
      cfg.dynamicSmemBytes = p->smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      e = cudaOccupancyMaxActiveClusters(&k, (const void*)p->fn[mode], &cfg);
    } else if (e == cudaSuccess) {
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, (const void*)p->fn[mode], RS_THREADS, p->smem);
      k *= sms;
    }
    if (e == cudaSuccess && k < 1) e = cudaErrorLaunchOutOfResources;
    if (e == cudaSuccess) NG = std::min(NG, k);
  }
  if (e != cudaSuccess || NG < 1) {
    rls_set_error("rowpass: no co-resident group of %d CTAs (%s)", G, cudaGetErrorString(e));
    cudaGetLastError();
    rls_rowpass_plan_destroy(p);
    return RLS_ERR_UNSUPPORTED;
  }
  if ((int64_t)NG > A->m) NG = (int)std::max<int64_t>(A->m, 1);
  p->NG = NG;
  p->gstride = ((int64_t)G * W + 63) & ~(int64_t)63;
  if (cudaMalloc(&p->gpart, (size_t)NG * p->gstride * 4) != cudaSuccess || cudaMalloc(&p->abort_flag, 4) != cudaSuccess) {
    rls_set_error("rowpass: out of device memory for the group partials");
    cudaGetLastError();
    rls_rowpass_plan_destroy(p);
    return RLS_ERR_NOMEM;
  }
  cudaMemsetAsync(p->abort_flag, 0, 4, c->stream);
  *out = p;
  return RLS_OK;
}

static int32_t rowstream_launch(RowPlan* p, int mode, const void* x, const void* yin, void* yout, void* res, const int* gate,
                                const float* xold = nullptr, const float* th_old = nullptr, const float* th = nullptr,
                                bool defer_finish = false) {
  rls_ctx_s* c = p->ctx;
  rls_mat_s* A = p->A;
  RowstreamArgs a;
  a.A = (const float*)A->d; a.ldf = A->ld * p->fpe; a.m = A->m; a.nf = (int)(A->n * p->fpe);
  a.W = p->W; a.NS = p->NS; a.G = p->G; a.NG = p->NG;
  a.x = (const float*)x; a.xold = xold; a.th_old = th_old; a.th = th;
  a.yin = (const float*)yin; a.yout = (float*)yout;
  a.gpart = p->gpart; a.gstride = p->gstride;
  a.gate = gate; a.abort_flag = p->abort_flag;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p->NG * p->G);
  cfg.blockDim = dim3(RS_THREADS);
  cfg.dynamicSmemBytes = p->smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  at[na].id = cudaLaunchAttributeClusterDimension;   // G = 1 is a cluster of one: the same code path posts to itself
  at[na].val.clusterDim.x = p->G; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
  ++na;
  if (rls_pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  rls_trace_begin(c->stream, mode == RS_NORMAL ? "rowstream normal" : mode == RS_GEMV_N ? "rowstream gemv_n" : "rowstream gemv_c");
  RLS_CUDA(cudaLaunchKernelEx(&cfg, p->fn[mode], a));
  rls_trace_end(c->stream);
  c->launches++;
  if (mode != RS_GEMV_N && !defer_finish) {
    const int nf = a.nf;
    int grid = std::min(c->sm_count * 2, std::max(1, (nf / 4 + 255) / 256));
    RLS_CUDA(rls_launch_pdl(c->stream, dim3(grid), dim3(256), rowpass_finish_kernel, (const float*)p->gpart, p->gstride, p->NG, nf,
                            (float*)res, gate));
    c->launches++;
  }
  return RLS_OK;
}

int32_t rls_rowpass_normal(RowPlan* p, const void* x, void* res, const int* gate) { return rowstream_launch(p, RS_NORMAL, x, nullptr, nullptr, res, gate); }
// the streaming kernel only: the caller's epilogue kernel sums the per-group partials itself (fixed order, the same
// arithmetic as rowpass_finish_kernel) — one kernel boundary less per iteration.  Optional fused FISTA momentum.
int32_t rls_rowpass_normal_deferred(RowPlan* p, const void* x, const float* xold, const float* th_old, const float* th, const int* gate,
                                    const float** gpart, int64_t* gstride, int* ncl) {
  RLS_TRY(rowstream_launch(p, RS_NORMAL, x, nullptr, nullptr, nullptr, gate, xold, th_old, th, true));
  *gpart = p->gpart; *gstride = p->gstride; *ncl = p->NG;
  return RLS_OK;
}
int32_t rls_rowpass_gemv_n(RowPlan* p, const void* x, void* y, const int* gate) { return rowstream_launch(p, RS_GEMV_N, x, nullptr, y, nullptr, gate); }
int32_t rls_rowpass_gemv_c(RowPlan* p, const void* y, void* g, const int* gate) { return rowstream_launch(p, RS_GEMV_C, nullptr, y, nullptr, g, gate); }

int32_t rls_rowpass_check_abort(RowPlan* p) {
  int flag = 0;
  RLS_CUDA(cudaMemcpyAsync(&flag, p->abort_flag, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
  RLS_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (flag) {
    rls_set_error("row-major streaming kernel timed out on a barrier or a cluster exchange (abort flag set)");
    return RLS_ERR_CUDA;
  }
  return RLS_OK;
}

void rls_rowpass_describe(RowPlan* p, char* buf, int len) {
  snprintf(buf, len, "onepass/rowstream: groups=%d x %d CTAs (%d SMs%s) slice=%d floats (V=%d) 16 compute + 2 service warps, 2-row pipeline, ring=%d x %d B smem=%zu B",
           p->NG, p->G, p->NG * p->G, p->G == 1 ? "" : ", cluster DSMEM exchange", p->W, p->V, p->NS, p->W * 4, p->smem);
}
