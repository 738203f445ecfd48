// rls_p2p.cu — one-shot all-reduce of the n-vector A_i'(A_i x) over NVLink peer memory, fused with the sum of
// the per-cluster partials of the one-pass kernel.
//
// Row-sharded solves need one sum over ranks of an n-vector per normal-operator apply.  With NCCL
// that is: finish kernel -> ncclAllReduce -> gated copy.  Here it is ONE kernel per rank:
//   1. add the per-cluster partials (fixed order) and write the rank's vector into its own exchange slot;
//   2. the last block publishes "epoch e is complete" into every peer's flag word (system-scope release store
//      through the peer mapping);
//   3. every block waits for the flags of all ranks (acquire, system scope, bounded), then adds the ranks'
//      vectors in rank order straight out of the peers' slots (128-bit loads over NVLink) into res.
// All ranks add in the same order, so replicas stay bit-identical, exactly as with the NCCL path.
// The exchange buffers are plain cudaMalloc allocations shared through CUDA IPC handles that the host ships with its
// own transport (torch.distributed here, MPI.jl in Julia), like the NCCL unique id.
// Two slots alternate by a DEVICE-side epoch counter (gated-off launches must not advance it): a rank can start
// epoch e+1 while a slow peer still reads its epoch-e slot, but not epoch e+2 (it needs that peer's e+1 flag first).
#include "rls_common.cuh"

namespace {

constexpr int P2P_THREADS = 256;
constexpr int P2P_FLAG_STRIDE = 32;  // uint32 words between flags (128 bytes)

struct PeerTable { float* base[RLS_MAX_PEERS]; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {  // never satisfied from a stale local line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// exchange buffer of one rank:  [2][cap] floats | flags[RLS_MAX_PEERS] (128 B apart) | epoch | ticket0 | ticket1
__device__ __forceinline__ unsigned* flags_of(float* base, int64_t cap) { return reinterpret_cast<unsigned*>(base + 2 * cap); }

__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_kernel(PeerTable peers, int rank, int nranks, int64_t cap, const float* __restrict__ src,
                                                                    int64_t sstride, int nsrc, int nf, float* __restrict__ res, const int* gate,
                                                                    int* abort_flag) {
  pdl_prologue();
  if (gate && *gate) return;
  __shared__ unsigned s_epoch;
  __shared__ int s_last;
  float* own = peers.base[rank];
  unsigned* oflags = flags_of(own, cap);
  unsigned* ctrl = oflags + RLS_MAX_PEERS * P2P_FLAG_STRIDE;  // [0] epoch, [1] ticket (publish), [2] ticket (retire)
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(ctrl) + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  float* slot = own + (size_t)(epoch & 1u) * cap;
  const int nf4 = (nf + 3) >> 2;  // cap is a multiple of 4 and the partial buffers are padded: whole float4s
  // 1. this rank's vector
  for (int i = blockIdx.x * P2P_THREADS + threadIdx.x; i < nf4; i += gridDim.x * P2P_THREADS) {
    double sx = 0.0, sy = 0.0, sz = 0.0, sw = 0.0;   // Float64, index order: the same sum as rowpass_finish_kernel
    for (int k = 0; k < nsrc; ++k) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * sstride) + i);
      sx += (double)t.x; sy += (double)t.y; sz += (double)t.z; sw += (double)t.w;
    }
    reinterpret_cast<float4*>(slot)[i] = make_float4((float)sx, (float)sy, (float)sz, (float)sw);
  }
  // 2. publish (the barrier orders the block's stores before thread 0's system-scope fence: fences are cumulative)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(&ctrl[1], 1u) == gridDim.x - 1;
    if (s_last) {
      ctrl[1] = 0u;
      __threadfence_system();
      for (int r = 0; r < nranks; ++r) st_release_sys(flags_of(peers.base[r], cap) + rank * P2P_FLAG_STRIDE, epoch);
    }
  }
  // 3. wait for every rank's epoch, then reduce in rank order
  if (threadIdx.x < nranks) {
    const unsigned* f = oflags + threadIdx.x * P2P_FLAG_STRIDE;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > 4000000000ll) { atomicExch(abort_flag, 1); break; }
    }
  }
  __syncthreads();
  for (int i = blockIdx.x * P2P_THREADS + threadIdx.x; i < nf4; i += gridDim.x * P2P_THREADS) {
    // all ranks' loads in flight together (one NVLink round trip), then the sum in rank order
    float4 v[RLS_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < RLS_MAX_PEERS; ++r)
      if (r < nranks) v[r] = ld_peer(reinterpret_cast<const float4*>(peers.base[r] + (size_t)(epoch & 1u) * cap) + i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < RLS_MAX_PEERS; ++r)
      if (r < nranks) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    const int j = 4 * i;
    if (j + 3 < nf) reinterpret_cast<float4*>(res)[i] = s;
    else {
      if (j < nf) res[j] = s.x;
      if (j + 1 < nf) res[j + 1] = s.y;
      if (j + 2 < nf) res[j + 2] = s.z;
    }
  }
  // retire: the last block to finish advances the device-side epoch
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&ctrl[2], 1u) == gridDim.x - 1) {
    ctrl[2] = 0u;
    __threadfence();
    *reinterpret_cast<volatile unsigned*>(ctrl) = epoch;
  }
}

}  // namespace

static size_t p2p_bytes(int64_t cap) { return (size_t)2 * cap * 4 + (size_t)(RLS_MAX_PEERS * P2P_FLAG_STRIDE + 32) * 4; }

extern "C" int32_t rls_ctx_peer_export(rls_ctx_t c, int64_t max_floats, void* handle64) {
  RLS_CHECK_ARG(c && handle64 && max_floats > 0, "bad argument");
  RLS_CHECK_ARG(c->nranks > 1 && c->nranks <= RLS_MAX_PEERS, "peer exchange needs 2..%d ranks (call rls_ctx_comm_init first)", RLS_MAX_PEERS);
  RLS_CHECK_ARG(!c->peer_own, "peer exchange buffer already exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  RlsDeviceGuard g(c->device);
  c->peer_cap = (max_floats + 3) & ~(int64_t)3;
  RLS_CUDA(cudaMalloc((void**)&c->peer_own, p2p_bytes(c->peer_cap)));
  RLS_CUDA(cudaMemset(c->peer_own, 0, p2p_bytes(c->peer_cap)));
  cudaIpcMemHandle_t h;
  RLS_CUDA(cudaIpcGetMemHandle(&h, c->peer_own));
  memcpy(handle64, &h, 64);
  return RLS_OK;
}

extern "C" int32_t rls_ctx_peer_import(rls_ctx_t c, const void* handles, int32_t nranks) {
  RLS_CHECK_ARG(c && handles && nranks == c->nranks, "bad argument");
  RLS_CHECK_ARG(c->peer_own, "call rls_ctx_peer_export first");
  RlsDeviceGuard g(c->device);
  for (int r = 0; r < nranks; ++r) {
    if (r == c->rank) { c->peer_base[r] = c->peer_own; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      rls_set_error("cudaIpcOpenMemHandle for rank %d failed: %s (peer access between the GPUs is required)", r, cudaGetErrorString(e));
      cudaGetLastError();
      return RLS_ERR_COMM;
    }
    c->peer_base[r] = (float*)p;
  }
  if (!c->peer_abort) { RLS_CUDA(cudaMalloc((void**)&c->peer_abort, 4)); RLS_CUDA(cudaMemset(c->peer_abort, 0, 4)); }
  c->peer_ready = true;
  return RLS_OK;
}

void rls_ctx_peer_release(rls_ctx_s* c) {
  for (int r = 0; r < RLS_MAX_PEERS; ++r)
    if (c->peer_base[r] && c->peer_base[r] != c->peer_own) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->peer_own) cudaFree(c->peer_own);
  if (c->peer_abort) cudaFree(c->peer_abort);
  c->peer_own = nullptr; c->peer_abort = nullptr; c->peer_ready = false;
}

// Opt-in (RLS_P2P=1).  Measured on 8 B200s (profiles/r01_allreduce_p2p_vs_nccl_n8.txt): NCCL's all-reduce of the
// 256 KB vector (NVLS, in-switch reduction) is ~1 % faster per iteration than this one-shot kernel, so NCCL stays
// the default exchange.
bool rls_p2p_available(const rls_ctx_s* c, int64_t nfloats) {
  return c->peer_ready && c->nranks > 1 && nfloats <= c->peer_cap && rls_env_flag("RLS_P2P", false);
}

// res = sum over ranks of (sum over the nsrc partial vectors src + k*sstride); float counts, nf <= peer_cap
int32_t rls_p2p_allreduce(rls_ctx_s* c, const float* src, int64_t sstride, int nsrc, int64_t nf, float* res, const int* gate) {
  RLS_CHECK_ARG(rls_p2p_available(c, nf), "peer exchange not set up / vector too long");
  PeerTable t{};
  for (int r = 0; r < c->nranks; ++r) t.base[r] = c->peer_base[r];
  const int nf4 = (int)((nf + 3) / 4);
  int grid = std::min(c->sm_count, std::max(1, (nf4 + P2P_THREADS - 1) / P2P_THREADS));  // all blocks co-resident (they wait on each other)
  RLS_CUDA(rls_launch_pdl(c->stream, dim3(grid), dim3(P2P_THREADS), p2p_allreduce_kernel, t, c->rank, c->nranks, c->peer_cap, src, sstride, nsrc,
                          (int)nf, res, gate, c->peer_abort));
  c->launches++;
  return RLS_OK;
}

int32_t rls_p2p_check_abort(rls_ctx_s* c) {
  if (!c->peer_ready) return RLS_OK;
  int flag = 0;
  RLS_CUDA(cudaMemcpyAsync(&flag, c->peer_abort, 4, cudaMemcpyDeviceToHost, c->stream));
  RLS_CUDA(cudaStreamSynchronize(c->stream));
  if (flag) { rls_set_error("peer all-reduce timed out waiting for a rank (abort flag set)"); return RLS_ERR_COMM; }
  return RLS_OK;
}
