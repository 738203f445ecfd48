// rls_tc.cu — tensor-core path of the multi-right-hand-side normal operator and of the Gram build.
//
// Replaces the K x gemv loop of MultiThreading.jl:45-78 (K independent states sharing one A: A is read
// K times per iteration) by two GEMMs per batched iteration, Y = A X and G = A' Y, that read A once
// each, and the `A'*A` constructor default (FISTA.jl:58, CGNR.jl:49, ADMM.jl:81) by one GEMM.
//
// FP32-accurate split precision on tcgen05 (kind::tf32, FP32 accumulators in TMEM).  The tensor core
// uses the upper 19 bits of every FP32 operand it is handed.  Four converter warps split every tile TMA delivers,
//     hi(a) = rn_tf32(a)          (written back over the tile)
//     lo(a) = rn_tf32(a - hi(a))  (|lo| <= 2^-11 |a|; written next to it)
// and the product is accumulated as  hi(A) hi(B) + hi(A) lo(B) + lo(A) hi(B)  — three MMAs per k-step; the
// dropped lo*lo term and the rounding of lo are O(2^-22) relative.
//
// Complex data stay interleaved.  With A~ the m x 2n real view of A (row-major: rows contiguous):
//   mode N:  Y~ (m x 2K)  = A~ (m x 2n) . B,  B rows (2j,2j+1) x cols (2k,2k+1) = [xr xi; -xi xr]
//            i.e. B^T row 2k = conj(x_k), row 2k+1 = i conj(x_k) as float vectors (pack kernel)
//   mode T:  P (2n x 2K) = A~^T (2n x m) . Y~,  then g_jk = (P[2j,2k] + P[2j+1,2k+1]) + i (P[2j,2k+1] - P[2j+1,2k])
//            (unpack kernel).  The Gram matrix is mode T with Y~ replaced by A~ itself.
// In mode N both operands are K-major (the reduction index is the contiguous one); in mode T both are
// MN-major — the same row-major tiles, described to the tensor core as transposed (a_major/b_major = 1).
//
// CTA = one 128 x N output tile, whole reduction:  warp 0 TMA producer, warp 1 MMA issuer (one lane),
// warps 2-5 lo-converters during the main loop and TMEM -> global epilogue afterwards.
#include <cuda.h>

#include <algorithm>

#include "rls_common.cuh"

namespace {

constexpr int TC_BM = 128;        // UMMA M
constexpr int TC_BK = 32;         // floats per k-block: one 128-byte swizzle row
constexpr int TC_THREADS = 320;    // warp 0 TMA, warp 1 MMA, warps 2-5 lo converters, warps 6-9 drain + epilogue
constexpr int TC_CONV_THREADS = 128;
constexpr int TC_BOX_BYTES = 32 * 32 * 4;  // one TMA box: 32 rows x 128 bytes

struct TcArgs {
  int transposed;     // 0: mode N (K-major operands), 1: mode T (MN-major operands)
  int Npad;           // UMMA N (multiple of 32, <= 256)
  int nkb;            // k-blocks
  int stages;
  int tmem_cols;      // power of two >= Npad
  int b_col0_from_y;  // mode T: B tile column origin = blockIdx.y * Npad (Gram), else 0
  int b_presplit;     // 1: the B operand arrives already split (hi via mapB, lo via mapBlo); the converters touch A only
  float* Dlo;         // != NULL: write the result split, D = rn_tf32(acc), Dlo = rn_tf32(acc - D) (it is the next GEMM's B)
  float* D;           // output, row-major [Mtot][ldd]
  int64_t ldd;
  int64_t Mtot;       // valid output rows
  int Nvalid;         // valid output columns (per blockIdx.y tile: min(Npad, Ntot - y*Npad))
  int64_t Ntot;
  int* abort_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait (see rls_async.cuh): a protocol error ends the launch instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, volatile int* s_abort, int* g_abort) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (*s_abort) return;
    if (clock64() - t0 > 4000000000ll) {
      *s_abort = 1;
      atomicExch(g_abort, 1);
      return;
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte
// offsets in 16-byte units, version 1 (sm_100), 128-byte swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;             // version
  d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Split a = hi + lo with BOTH parts rounded to TF32 to nearest (cvt.rna): |lo| <= 2^-11 |a| and the representation
// error |a - hi - lo| <= 2^-22 |a|, half of what the tensor core's own truncation of a raw FP32 operand would leave
// (hi = trunc(a), |lo| < 2^-10 |a|).  The price is that hi has to be written back over the tile TMA delivered.
__device__ __forceinline__ float rna_tf32(float a) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
  hi = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
  lo = make_float4(rna_tf32(v.x - hi.x), rna_tf32(v.y - hi.y), rna_tf32(v.z - hi.z), rna_tf32(v.w - hi.w));
}

// TMEM -> registers: 32 consecutive FP32 columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Accumulation.  The tensor core adds into its FP32 accumulator with round-toward-zero, a bias that grows
// linearly with the length of the accumulation chain (measured: 4e-5 relative at a reduction length of 4096).
// The dominant hi*hi products are therefore accumulated in TMEM over ONE k-block only (4 MMAs) into one of two
// alternating accumulators; four drain warps pull each finished partial into FP32 registers and add it there
// with round-to-nearest while the tensor core fills the other accumulator.  The two cross terms are 2^-11 of
// the result, so their chain may run over the whole reduction in a third accumulator (its bias is 2^-11 smaller).
template <int NPAD>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo,
               TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int Npad = NPAD;
  const int NS = p.stages;
  constexpr uint32_t a_bytes = TC_BM * TC_BK * 4;             // 16 KB
  constexpr uint32_t b_bytes = (uint32_t)Npad * TC_BK * 4;
  constexpr uint32_t hi_bytes = a_bytes + b_bytes;
  constexpr uint32_t stage_bytes = 2 * hi_bytes;              // [A hi | B hi | A lo | B lo]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * stage_bytes);
  uint64_t* conv = full + NS;
  uint64_t* empty = conv + NS;
  uint64_t* accf = empty + NS;   // [2] partial accumulator b complete (tcgen05.commit)
  uint64_t* acce = accf + 2;     // [2] partial accumulator b drained (128 drain threads)
  uint64_t* done = acce + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(tmem_slot + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], TC_CONV_THREADS); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&accf[b], 1); mbar_init(&acce[b], TC_CONV_THREADS); }
    mbar_init(done, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_small = tmem_base + 2u * Npad;  // columns [0,N) and [N,2N): alternating partials; [2N,3N): cross terms

  const int m0 = blockIdx.x * TC_BM;       // output rows of this tile
  const int n0 = p.b_col0_from_y ? blockIdx.y * Npad : 0;
  constexpr int ngb = Npad / 32;           // 32-wide groups of the B tile

  if (warp == 0) {
    // ------------------------------ TMA producer --------------------------------------
    if (lane == 0) {
      int s = 0;
      unsigned ph = 0;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1u, s_abort, p.abort_flag);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], hi_bytes + (p.b_presplit ? b_bytes : 0u));
        if (!p.transposed) {
          // K-major tiles: rows = output index, 32 reduction floats per row
          for (int g = 0; g < TC_BM / 32; ++g) tma_load_2d(st + g * TC_BOX_BYTES, &mapA, kb * TC_BK, m0 + 32 * g, &full[s]);
          for (int g = 0; g < ngb; ++g) tma_load_2d(st + a_bytes + g * TC_BOX_BYTES, &mapB, kb * TC_BK, n0 + 32 * g, &full[s]);
          if (p.b_presplit)
            for (int g = 0; g < ngb; ++g) tma_load_2d(st + hi_bytes + a_bytes + g * TC_BOX_BYTES, &mapBlo, kb * TC_BK, n0 + 32 * g, &full[s]);
        } else {
          // MN-major tiles: rows = reduction index (32 of them), 32 output floats per row, one box per 32-wide group
          for (int g = 0; g < TC_BM / 32; ++g) tma_load_2d(st + g * TC_BOX_BYTES, &mapA, m0 + 32 * g, kb * TC_BK, &full[s]);
          for (int g = 0; g < ngb; ++g) tma_load_2d(st + a_bytes + g * TC_BOX_BYTES, &mapB, n0 + 32 * g, kb * TC_BK, &full[s]);
          if (p.b_presplit)
            for (int g = 0; g < ngb; ++g) tma_load_2d(st + hi_bytes + a_bytes + g * TC_BOX_BYTES, &mapBlo, n0 + 32 * g, kb * TC_BK, &full[s]);
        }
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ----------------------------------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.transposed << 15) | ((uint32_t)p.transposed << 16) |
                             ((uint32_t)(Npad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      // K-major (SWIZZLE_128B): 8-row groups 1024 B apart, k-step = 32 B inside the swizzle row.
      // MN-major: 32-bit operands only exist in the 128B swizzle with 32-byte atoms (SWIZZLE_128B_BASE32B, the TMA
      // mode 128B_ATOM_32B): atom = 32 floats (MN) x 4 k-rows; 32-wide MN groups one box (4096 B) apart (LBO),
      // 4-k-row groups 512 B apart (SBO), k-step (8 k-rows) = 1024 B.
      const uint32_t lbo = p.transposed ? (uint32_t)TC_BOX_BYTES : 16u;
      const uint32_t sbo = p.transposed ? 512u : 1024u;
      const uint32_t kstep = p.transposed ? 1024u : 32u;
      const uint32_t lt = p.transposed ? 1u : 2u;
      int s = 0;
      unsigned ph = 0;
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int b = kb & 1;
        mbar_wait(&acce[b], (((unsigned)kb >> 1) & 1u) ^ 1u, s_abort, p.abort_flag);  // partial accumulator b has been drained
        mbar_wait(&full[s], ph, s_abort, p.abort_flag);   // hi tiles (TMA)
        mbar_wait(&conv[s], ph, s_abort, p.abort_flag);   // lo tiles (converter warps)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_hi = a_hi + a_bytes;
        const uint32_t a_lo = a_hi + hi_bytes;
        const uint32_t b_lo = b_hi + hi_bytes;
        const uint32_t tmem_main = tmem_base + (uint32_t)b * Npad;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t dah = make_desc(a_hi + k * kstep, lbo, sbo, lt), dbh = make_desc(b_hi + k * kstep, lbo, sbo, lt);
          umma_tf32(tmem_main, dah, dbh, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(&accf[b]);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t dah = make_desc(a_hi + k * kstep, lbo, sbo, lt), dbh = make_desc(b_hi + k * kstep, lbo, sbo, lt);
          const uint64_t dal = make_desc(a_lo + k * kstep, lbo, sbo, lt), dbl = make_desc(b_lo + k * kstep, lbo, sbo, lt);
          umma_tf32(tmem_small, dal, dbh, idesc, (kb | k) > 0 ? 1u : 0u);
          umma_tf32(tmem_small, dah, dbl, idesc, 1u);
        }
        umma_commit(&empty[s]);  // implies tcgen05.fence::before_thread_sync
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
      umma_commit(done);
    }
  } else if (warp < 6) {
    // ------------------------------ lo converters -------------------------------------
    const int ct = threadIdx.x - 64;  // 0..127
    int s = 0;
    unsigned ph = 0;
    const int n16 = (int)((p.b_presplit ? a_bytes : hi_bytes) >> 4);  // [A | B] are contiguous: A only when B is pre-split
    for (int kb = 0; kb < p.nkb; ++kb) {
      mbar_wait(&full[s], ph, s_abort, p.abort_flag);
      float4* hi = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
      float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + hi_bytes);
#pragma unroll 4
      for (int i = ct; i < n16; i += TC_CONV_THREADS) {
        float4 h, l;
        split4(hi[i], h, l);
        hi[i] = h;
        lo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&conv[s]);
      if (++s == NS) { s = 0; ph ^= 1u; }
    }
  } else {
    // ------------------------------ drain warps + epilogue ----------------------------
    // warp q = warp % 4 owns TMEM lanes [32q, 32q+32); thread = one output row, NPAD running sums in registers
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    float acc[NPAD];
#pragma unroll
    for (int j = 0; j < NPAD; ++j) acc[j] = 0.f;
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int b = kb & 1;
      mbar_wait(&accf[b], ((unsigned)kb >> 1) & 1u, s_abort, p.abort_flag);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < NPAD; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)(b * NPAD + c0), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acce[b]);
    }
    mbar_wait(done, 0u, s_abort, p.abort_flag);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t row = (int64_t)m0 + 32 * q + lane;
    float* drow = p.D + row * p.ldd + n0;
#pragma unroll
    for (int c0 = 0; c0 < NPAD; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_small + lane_base + (uint32_t)c0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      if (row < p.Mtot) {
        const int nv = min(32, p.Nvalid - c0);
        if (p.Dlo) {  // Npad-wide, aligned: the result is the next GEMM's B operand, stored split
          float* lrow = p.Dlo + row * p.ldd + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 h, l;
            split4(make_float4(acc[c0 + j], acc[c0 + j + 1], acc[c0 + j + 2], acc[c0 + j + 3]), h, l);
            *reinterpret_cast<float4*>(drow + c0 + j) = h;
            *reinterpret_cast<float4*>(lrow + c0 + j) = l;
          }
        } else if (nv == 32 && ((p.ldd | n0) & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(drow + c0 + j) = make_float4(acc[c0 + j], acc[c0 + j + 1], acc[c0 + j + 2], acc[c0 + j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv) drow[c0 + j] = acc[c0 + j];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- pack / unpack between K lane vectors and the GEMM operands -------------------------
// The K device pointers travel as kernel parameters (<= 128 columns per GEMM pass): no per-iteration H2D copy.
constexpr int TC_MAXK = 128;
struct PtrTable { const void* p[TC_MAXK]; };

// B^T of mode N, [Npad][ldb] floats.  complex: row 2k = conj(x_k), row 2k+1 = i conj(x_k); real: row k = x_k.
__global__ void tc_pack_x_kernel(const __grid_constant__ PtrTable xs, int K, int fpe, int64_t n, float* __restrict__ BT, float* __restrict__ BTlo,
                                 int64_t ldb, int Npad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int rowp = blockIdx.y;  // B^T row
  if (j >= n) return;
  if (fpe == 1) {
    const float v = rowp < K ? reinterpret_cast<const float*>(xs.p[rowp])[j] : 0.f;
    const float h = rna_tf32(v);
    BT[(int64_t)rowp * ldb + j] = h;
    BTlo[(int64_t)rowp * ldb + j] = rna_tf32(v - h);
  } else {
    const int k = rowp >> 1;
    float2 v = make_float2(0.f, 0.f);
    if (k < K) v = reinterpret_cast<const float2*>(xs.p[k])[j];
    const float2 o = (rowp & 1) ? make_float2(v.y, v.x) : make_float2(v.x, -v.y);
    const float2 h = make_float2(rna_tf32(o.x), rna_tf32(o.y));
    reinterpret_cast<float2*>(BT + (int64_t)rowp * ldb)[j] = h;
    reinterpret_cast<float2*>(BTlo + (int64_t)rowp * ldb)[j] = make_float2(rna_tf32(o.x - h.x), rna_tf32(o.y - h.y));
  }
}
// res_k[j] from P [nf][ldp]: complex (P[2j,2k] + P[2j+1,2k+1]) + i (P[2j,2k+1] - P[2j+1,2k]); real P[j,k].  Lane k is
// skipped when its done() gate is set.
__global__ void tc_unpack_g_kernel(const float* __restrict__ P, int64_t ldp, int K, int fpe, int64_t n, const __grid_constant__ PtrTable outs,
                                   const __grid_constant__ PtrTable gates, int have_gates, int conj_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  const int k = threadIdx.x + blockIdx.y * blockDim.x;
  if (j >= n || k >= K) return;
  const int* gate = have_gates ? reinterpret_cast<const int*>(gates.p[k]) : nullptr;
  if (gate && *gate) return;
  if (fpe == 1) {
    reinterpret_cast<float*>(const_cast<void*>(outs.p[k]))[j] = P[j * ldp + k];
  } else {
    const float2 a = *reinterpret_cast<const float2*>(P + (2 * j) * ldp + 2 * k);
    const float2 b = *reinterpret_cast<const float2*>(P + (2 * j + 1) * ldp + 2 * k);
    const float im = a.y - b.x;
    reinterpret_cast<float2*>(const_cast<void*>(outs.p[k]))[j] = make_float2(a.x + b.y, conj_out ? -im : im);
  }
}
// Y~ of mode T from K m-vectors (the batched back-projection A'B of init!): Y~[i][fpe*k + c] = b_k[i*fpe + c]
// (conj_in: the imaginary parts change sign — the Gram form below feeds conj(x_k))
__global__ void tc_pack_y_kernel(const __grid_constant__ PtrTable bs, int K, int fpe, int64_t m, float* __restrict__ Y, float* __restrict__ Ylo, int Npad,
                                 int conj_in) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  const int col = threadIdx.x + blockIdx.y * blockDim.x;  // float column of Y~
  if (i >= m || col >= Npad) return;
  const int k = col / fpe, c = col - k * fpe;
  float v = k < K ? reinterpret_cast<const float*>(bs.p[k])[i * fpe + c] : 0.f;
  if (conj_in && c == 1) v = -v;
  const float h = rna_tf32(v);
  Y[i * Npad + col] = h;
  Ylo[i * Npad + col] = rna_tf32(v - h);
}
// Gram epilogue: G (n x n, column-major) from P = A~^T A~ (nf x ldp).  complex: G[j,j'] = conj-combined 2x2 block.
__global__ void tc_gram_finish_kernel(const float* __restrict__ P, int64_t ldp, int fpe, int64_t n, float* __restrict__ G, int64_t ldg) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // row of G
  const int64_t jp = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;  // column of G
  if (j >= n || jp >= n) return;
  if (fpe == 1) {
    G[j + jp * ldg] = P[j * ldp + jp];
  } else {
    const float2 a = *reinterpret_cast<const float2*>(P + (2 * j) * ldp + 2 * jp);
    const float2 b = *reinterpret_cast<const float2*>(P + (2 * j + 1) * ldp + 2 * jp);
    reinterpret_cast<float2*>(G)[j + jp * ldg] = make_float2(a.x + b.y, a.y - b.x);
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  fn = (encode_tiled_fn)p;
  return fn;
}
// row-major float matrix [rows][ld], box = 32 floats x 32 rows, zero fill out of bounds.  128-byte swizzle with
// 16-byte atoms for K-major operands, with 32-byte atoms for MN-major (transposed) ones.
int32_t make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, bool mn_major) {
  encode_tiled_fn enc = get_encode();
  if (!enc) { rls_set_error("cuTensorMapEncodeTiled is not available from this driver"); return RLS_ERR_UNSUPPORTED; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rls_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return RLS_ERR_UNSUPPORTED; }
  return RLS_OK;
}

int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

size_t tc_smem(int stages, int Npad) {
  const size_t stage = 2 * ((size_t)TC_BM * TC_BK * 4 + (size_t)Npad * TC_BK * 4);
  return (size_t)stages * stage + (size_t)(3 * stages + 5) * 8 + 16 + 1024;
}

int32_t tc_launch(rls_ctx_s* c, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapBlo, TcArgs a, int grid_x, int grid_y) {
  int dev_max = 0;
  cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
  int stages = 6;
  while (stages > 2 && tc_smem(stages, a.Npad) > (size_t)dev_max) --stages;
  a.stages = stages;
  const size_t smem = tc_smem(stages, a.Npad);
  a.tmem_cols = pow2_cols(3 * a.Npad);
  void (*fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, TcArgs) =
      a.Npad == 32 ? tc_gemm_kernel<32> : a.Npad == 64 ? tc_gemm_kernel<64> : a.Npad == 96 ? tc_gemm_kernel<96> : tc_gemm_kernel<128>;
  RLS_CHECK_ARG(a.Npad == 32 || a.Npad == 64 || a.Npad == 96 || a.Npad == 128, "tensor-core GEMM: N tile %d not in {32,64,96,128}", a.Npad);
  RLS_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fn<<<dim3(grid_x, grid_y), TC_THREADS, smem, c->stream>>>(mapA, mapB, mapBlo, a);
  c->launches++;
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------
// batched normal operator: res_k = A'(A x_k), k = 0..K-1
// ------------------------------------------------------------------------------------
struct TcBatchPlan {
  rls_ctx_s* ctx = nullptr;
  rls_mat_s* A = nullptr;
  rls_mat_s view;         // Gram form: the column-major G seen as the row-major matrix M = G^T on the same memory
  bool gram = false;
  int K = 0, fpe = 1, Npad = 0;
  int64_t nf = 0, ldx = 0;
  float* XT = nullptr;    // [Npad][ldx]   B^T of mode N, TF32 hi part
  float* XTlo = nullptr;  //               ... lo part
  float* Y = nullptr;     // [m][Npad]     Y~, hi part (written split by mode N's epilogue / tc_pack_y_kernel)
  float* Ylo = nullptr;   //               ... lo part
  float* P = nullptr;     // [nf][Npad]    A~^T Y~
  int* abort_flag = nullptr;
  CUtensorMap mapA, mapAt, mapXT, mapXTlo, mapY, mapYlo;  // mapAt: A described for the transposed (MN-major) use
};

void rls_tc_batch_destroy(TcBatchPlan* p) {
  if (!p) return;
  cudaFree(p->XT); cudaFree(p->XTlo); cudaFree(p->Y); cudaFree(p->Ylo); cudaFree(p->P);
  cudaFree(p->abort_flag);
  delete p;
}

bool rls_tc_batch_supported(const rls_mat_s* A, int K) {
  // below ~8 columns K one-pass sweeps (K x m n s bytes at HBM speed) are cheaper than two GEMMs whose time is
  // nearly independent of N (measured at the C4 shape: 3.4 ms for N = 32 against 0.33 ms per one-pass apply)
  const char* mk = getenv("RLS_BATCH_MIN_K");
  const int min_k = mk ? atoi(mk) : 8;
  if (!A || A->layout != RLS_LAYOUT_ROWMAJOR || K < 2 || K < min_k) return false;
  const int fpe = A->dtype == RLS_C32 ? 2 : 1;
  if (K * fpe > 128) return false;  // one 128 x N tile per CTA, N <= 128 (three N-wide accumulators in TMEM)
  if (((uintptr_t)A->d & 15) != 0 || (A->ld * fpe) % 4 != 0) return false;
  if (A->m * (int64_t)fpe > 0x7fffffff || A->n * (int64_t)fpe > 0x7fffffff) return false;
  return A->ctx->cc_major == 10;
}

int32_t rls_tc_batch_create(rls_mat_s* A, int K, TcBatchPlan** out) {
  *out = nullptr;
  if (!rls_tc_batch_supported(A, K)) { rls_set_error("tensor-core batch path: unsupported matrix / K"); return RLS_ERR_UNSUPPORTED; }
  TcBatchPlan* p = new TcBatchPlan();
  p->ctx = A->ctx; p->A = A; p->K = K;
  p->fpe = A->dtype == RLS_C32 ? 2 : 1;
  p->nf = A->n * p->fpe;
  p->Npad = ((K * p->fpe + 31) / 32) * 32;
  p->ldx = (p->nf + 3) & ~(int64_t)3;
  bool ok = cudaMalloc(&p->XT, (size_t)p->Npad * p->ldx * 4) == cudaSuccess && cudaMalloc(&p->XTlo, (size_t)p->Npad * p->ldx * 4) == cudaSuccess &&
            cudaMalloc(&p->Y, (size_t)std::max<int64_t>(A->m, 1) * p->Npad * 4) == cudaSuccess &&
            cudaMalloc(&p->Ylo, (size_t)std::max<int64_t>(A->m, 1) * p->Npad * 4) == cudaSuccess &&
            cudaMalloc(&p->P, (size_t)p->nf * p->Npad * 4) == cudaSuccess &&
            cudaMalloc(&p->abort_flag, 4) == cudaSuccess;
  if (!ok) { cudaGetLastError(); rls_tc_batch_destroy(p); rls_set_error("tensor-core batch path: out of device memory"); return RLS_ERR_NOMEM; }
  cudaMemsetAsync(p->abort_flag, 0, 4, A->ctx->stream);
  int32_t s = make_map(&p->mapA, (const float*)A->d, A->m, p->nf, A->ld * p->fpe, false);
  if (s == RLS_OK) s = make_map(&p->mapAt, (const float*)A->d, A->m, p->nf, A->ld * p->fpe, true);
  if (s == RLS_OK) s = make_map(&p->mapXT, p->XT, p->Npad, p->nf, p->ldx, false);
  if (s == RLS_OK) s = make_map(&p->mapXTlo, p->XTlo, p->Npad, p->nf, p->ldx, false);
  if (s == RLS_OK) s = make_map(&p->mapY, p->Y, A->m, p->Npad, p->Npad, true);
  if (s == RLS_OK) s = make_map(&p->mapYlo, p->Ylo, A->m, p->Npad, p->Npad, true);
  if (s != RLS_OK) { rls_tc_batch_destroy(p); return s; }
  *out = p;
  return RLS_OK;
}

// xs / outs / gates: host arrays of K device pointers (gates may be NULL, entries may be NULL)
int32_t rls_tc_batch_apply(TcBatchPlan* p, const void* const* xs, void* const* outs, const int* const* gates) {
  rls_ctx_s* c = p->ctx;
  rls_mat_s* A = p->A;
  const int K = p->K;
  if (A->m == 0 || A->n == 0) {  // empty reduction: A'(A x) = 0
    for (int k = 0; k < K; ++k) RLS_CUDA(cudaMemsetAsync(outs[k], 0, A->n * rls_elem_size(A->dtype), c->stream));
    return RLS_OK;
  }
  PtrTable tx{}, to{}, tg{};
  for (int k = 0; k < K; ++k) { tx.p[k] = xs[k]; to.p[k] = outs[k]; tg.p[k] = gates ? gates[k] : nullptr; }
  {
    dim3 grid((unsigned)((A->n + 255) / 256), (unsigned)p->Npad);
    tc_pack_x_kernel<<<grid, 256, 0, c->stream>>>(tx, K, p->fpe, A->n, p->XT, p->XTlo, p->ldx, p->Npad);
    c->launches++;
  }
  TcArgs a{};
  a.Npad = p->Npad; a.b_col0_from_y = 0; a.abort_flag = p->abort_flag;
  // mode N: Y~ = A~ . B
  a.transposed = 0; a.nkb = (int)((p->nf + TC_BK - 1) / TC_BK);
  a.b_presplit = 1;
  a.D = p->Y; a.Dlo = p->Ylo; a.ldd = p->Npad; a.Mtot = A->m; a.Nvalid = p->Npad; a.Ntot = p->Npad;
  RLS_TRY(tc_launch(c, p->mapA, p->mapXT, p->mapXTlo, a, (int)((A->m + TC_BM - 1) / TC_BM), 1));
  // mode T: P = A~^T . Y~
  a.transposed = 1; a.nkb = (int)((A->m + TC_BK - 1) / TC_BK);
  a.D = p->P; a.Dlo = nullptr; a.ldd = p->Npad; a.Mtot = p->nf;
  RLS_TRY(tc_launch(c, p->mapAt, p->mapY, p->mapYlo, a, (int)((p->nf + TC_BM - 1) / TC_BM), 1));
  if (c->nranks > 1) RLS_TRY(rls_allreduce_raw(c, p->P, p->nf * p->Npad));
  {
    dim3 block(32, 8);
    dim3 grid((unsigned)((A->n + 7) / 8), (unsigned)((K + 31) / 32));
    tc_unpack_g_kernel<<<grid, block, 0, c->stream>>>(p->P, p->Npad, K, p->fpe, A->n, to, tg, gates ? 1 : 0, 0);
    c->launches++;
  }
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}

// outs_k = A' b_k for K m-vectors (init!: mul!(x0, adjoint(A), b), FISTA.jl:114, CGNR.jl:132, ADMM.jl:198) as ONE mode-T GEMM.
// conj != 0: outs_k = conj(A' conj(b_k)) = A^T b_k (the Gram form); gates as in rls_tc_batch_apply.
static int32_t tc_adjoint(TcBatchPlan* p, const void* const* bs, void* const* outs, const int* const* gates, int conj) {
  rls_ctx_s* c = p->ctx;
  rls_mat_s* A = p->A;
  const int K = p->K;
  if (A->m == 0 || A->n == 0) {
    for (int k = 0; k < K; ++k) RLS_CUDA(cudaMemsetAsync(outs[k], 0, A->n * rls_elem_size(A->dtype), c->stream));
    return RLS_OK;
  }
  PtrTable tb{}, to{}, tg{};
  for (int k = 0; k < K; ++k) { tb.p[k] = bs[k]; to.p[k] = outs[k]; tg.p[k] = gates ? gates[k] : nullptr; }
  {
    dim3 block(32, 8);
    dim3 grid((unsigned)((A->m + 7) / 8), (unsigned)((p->Npad + 31) / 32));
    tc_pack_y_kernel<<<grid, block, 0, c->stream>>>(tb, K, p->fpe, A->m, p->Y, p->Ylo, p->Npad, conj);
    c->launches++;
  }
  TcArgs a{};
  a.Npad = p->Npad; a.b_col0_from_y = 0; a.abort_flag = p->abort_flag;
  a.transposed = 1; a.nkb = (int)((A->m + TC_BK - 1) / TC_BK);
  a.b_presplit = 1;
  a.D = p->P; a.ldd = p->Npad; a.Mtot = p->nf; a.Nvalid = p->Npad; a.Ntot = p->Npad;
  RLS_TRY(tc_launch(c, p->mapAt, p->mapY, p->mapYlo, a, (int)((p->nf + TC_BM - 1) / TC_BM), 1));
  if (c->nranks > 1 && !p->gram) RLS_TRY(rls_allreduce_raw(c, p->P, p->nf * p->Npad));  // G is already the sum over the ranks
  {
    dim3 block(32, 8);
    dim3 grid((unsigned)((A->n + 7) / 8), (unsigned)((K + 31) / 32));
    tc_unpack_g_kernel<<<grid, block, 0, c->stream>>>(p->P, p->Npad, K, p->fpe, A->n, to, tg, gates ? 1 : 0, conj);
    c->launches++;
  }
  RLS_CUDA(cudaGetLastError());
  return RLS_OK;
}
int32_t rls_tc_batch_adjoint(TcBatchPlan* p, const void* const* bs, void* const* outs) {
  if (p->gram) { rls_set_error("tensor-core batch plan of a Gram operator has no A"); return RLS_ERR_INVALID; }
  return tc_adjoint(p, bs, outs, nullptr, 0);
}

// ------------------------------------------------------------------------------------
// Gram form of the batched normal operator: res_k = G x_k, G = A'A dense n x n (the reference's DEFAULT AHA,
// FISTA.jl:58 / CGNR.jl:49 / ADMM.jl:81, under MultiThreading.jl:45-78) as ONE GEMM that reads G once:
// n^2 s bytes and 8 n^2 K flop (complex) per batched iteration against 2 m n s and 16 m n K for the A-form pair.
// The column-major G is, on the same memory, the row-major matrix M = G^T, and
//     (G x)_j = sum_i M[i][j] x_i = conj( sum_i conj(M[i][j]) conj(x_i) ) = conj( (M' conj(x))_j ),
// i.e. the mode-T GEMM of the back-projection with the imaginary parts flipped on the way in and out — no
// symmetry of G is assumed (a user-supplied AHA, FISTA.jl:55, is applied as given).
// ------------------------------------------------------------------------------------
bool rls_tc_gram_batch_supported(const rls_mat_s* G, int K) {
  const char* mk = getenv("RLS_BATCH_MIN_K");
  const int min_k = mk ? atoi(mk) : 8;
  if (!G || G->layout != RLS_LAYOUT_COLMAJOR || G->m != G->n || K < 2 || K < min_k) return false;
  const int fpe = G->dtype == RLS_C32 ? 2 : 1;
  if (K * fpe > 128) return false;
  if (((uintptr_t)G->d & 15) != 0 || (G->ld * fpe) % 4 != 0) return false;
  if (G->n * (int64_t)fpe > 0x7fffffff) return false;
  return G->ctx->cc_major == 10;
}

int32_t rls_tc_gram_batch_create(rls_mat_s* G, int K, TcBatchPlan** out) {
  *out = nullptr;
  if (!rls_tc_gram_batch_supported(G, K)) { rls_set_error("tensor-core Gram batch path: unsupported matrix / K"); return RLS_ERR_UNSUPPORTED; }
  TcBatchPlan* p = new TcBatchPlan();
  p->gram = true;
  p->view.ctx = G->ctx; p->view.dtype = G->dtype; p->view.m = G->n; p->view.n = G->n; p->view.ld = G->ld; p->view.d = G->d;
  p->view.owned = false; p->view.layout = RLS_LAYOUT_ROWMAJOR;
  p->ctx = G->ctx; p->A = &p->view; p->K = K;
  p->fpe = G->dtype == RLS_C32 ? 2 : 1;
  p->nf = G->n * p->fpe;
  p->Npad = ((K * p->fpe + 31) / 32) * 32;
  p->ldx = 0;
  bool ok = cudaMalloc(&p->Y, (size_t)std::max<int64_t>(G->n, 1) * p->Npad * 4) == cudaSuccess &&
            cudaMalloc(&p->Ylo, (size_t)std::max<int64_t>(G->n, 1) * p->Npad * 4) == cudaSuccess &&
            cudaMalloc(&p->P, (size_t)std::max<int64_t>(p->nf, 1) * p->Npad * 4) == cudaSuccess &&
            cudaMalloc(&p->abort_flag, 4) == cudaSuccess;
  if (!ok) { cudaGetLastError(); rls_tc_batch_destroy(p); rls_set_error("tensor-core Gram batch path: out of device memory"); return RLS_ERR_NOMEM; }
  cudaMemsetAsync(p->abort_flag, 0, 4, G->ctx->stream);
  int32_t s = RLS_OK;
  if (G->n > 0) {
    s = make_map(&p->mapAt, (const float*)G->d, G->n, p->nf, G->ld * p->fpe, true);
    if (s == RLS_OK) s = make_map(&p->mapY, p->Y, G->n, p->Npad, p->Npad, true);
    if (s == RLS_OK) s = make_map(&p->mapYlo, p->Ylo, G->n, p->Npad, p->Npad, true);
  }
  if (s != RLS_OK) { rls_tc_batch_destroy(p); return s; }
  *out = p;
  return RLS_OK;
}

int32_t rls_tc_gram_batch_apply(TcBatchPlan* p, const void* const* xs, void* const* outs, const int* const* gates) {
  if (!p->gram) { rls_set_error("not a Gram batch plan"); return RLS_ERR_INVALID; }
  return tc_adjoint(p, xs, outs, gates, p->fpe == 2 ? 1 : 0);
}

int32_t rls_tc_check_abort(rls_ctx_s* c, int* abort_flag) {
  int flag = 0;
  RLS_CUDA(cudaMemcpyAsync(&flag, abort_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  if (flag) { rls_set_error("tensor-core GEMM timed out on a barrier (abort flag set)"); return RLS_ERR_CUDA; }
  return RLS_OK;
}
// diagnostics: copy an internal operand (0 = B^T of mode N, 1 = Y~, 2 = P) to the host
int32_t rls_tc_batch_debug(TcBatchPlan* p, int which, float* host, int64_t nfloats) {
  const float* src = which == 0 ? p->XT : which == 1 ? p->Y : p->P;
  if (!src) { rls_set_error("operand %d does not exist in this plan", which); return RLS_ERR_INVALID; }
  const int64_t have = which == 0 ? (int64_t)p->Npad * p->ldx : which == 1 ? p->A->m * (int64_t)p->Npad : p->nf * (int64_t)p->Npad;
  RLS_CUDA(cudaStreamSynchronize(p->ctx->stream));
  RLS_CUDA(cudaMemcpy(host, src, sizeof(float) * std::min(nfloats, have), cudaMemcpyDeviceToHost));
  return RLS_OK;
}
int32_t rls_tc_batch_check_abort(TcBatchPlan* p) { return rls_tc_check_abort(p->ctx, p->abort_flag); }

// ------------------------------------------------------------------------------------
// Gram matrix G = A'A (n x n, column-major, dense) on the tensor cores
// ------------------------------------------------------------------------------------
int32_t rls_tc_gram(rls_mat_s* A, rls_mat_s* G) {
  rls_ctx_s* c = A->ctx;
  const int fpe = A->dtype == RLS_C32 ? 2 : 1;
  const int64_t nf = A->n * fpe;
  if (A->layout != RLS_LAYOUT_ROWMAJOR || G->layout != RLS_LAYOUT_COLMAJOR || c->cc_major != 10 || ((uintptr_t)A->d & 15) != 0 ||
      (A->ld * fpe) % 4 != 0 || nf > 0x7fffffff || A->m > 0x7fffffff) {
    rls_set_error("tensor-core Gram: unsupported matrix");
    return RLS_ERR_UNSUPPORTED;
  }
  if (A->m == 0 || A->n == 0) {
    RLS_CUDA(cudaMemsetAsync(G->d, 0, (size_t)G->ld * (size_t)std::max<int64_t>(G->n, 1) * rls_elem_size(G->dtype), c->stream));
    return RLS_OK;
  }
  const int Npad = 128;
  const int64_t ldp = ((nf + Npad - 1) / Npad) * Npad;
  float* P = nullptr;
  int* abort_flag = nullptr;
  if (cudaMalloc(&P, (size_t)nf * ldp * 4) != cudaSuccess || cudaMalloc(&abort_flag, 4) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(P);
    rls_set_error("tensor-core Gram: out of device memory for the %lld x %lld real product", (long long)nf, (long long)ldp);
    return RLS_ERR_NOMEM;
  }
  cudaMemsetAsync(abort_flag, 0, 4, c->stream);
  CUtensorMap mapA;
  int32_t s = make_map(&mapA, (const float*)A->d, A->m, nf, A->ld * fpe, true);
  if (s == RLS_OK) {
    TcArgs a{};
    a.transposed = 1; a.Npad = Npad; a.b_col0_from_y = 1; a.abort_flag = abort_flag;
    a.nkb = (int)((A->m + TC_BK - 1) / TC_BK);
    a.D = P; a.ldd = ldp; a.Mtot = nf; a.Nvalid = Npad; a.Ntot = ldp;
    s = tc_launch(c, mapA, mapA, mapA, a, (int)((nf + TC_BM - 1) / TC_BM), (int)(ldp / Npad));
  }
  if (s == RLS_OK) {
    dim3 block(32, 8);
    dim3 grid((unsigned)((A->n + 31) / 32), (unsigned)((A->n + 7) / 8));
    tc_gram_finish_kernel<<<grid, block, 0, c->stream>>>(P, ldp, fpe, A->n, (float*)G->d, G->ld);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) s = RLS_ERR_CUDA;
  }
  if (s == RLS_OK) s = rls_tc_check_abort(c, abort_flag);
  cudaStreamSynchronize(c->stream);
  cudaFree(P);
  cudaFree(abort_flag);
  return s;
}
