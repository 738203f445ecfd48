/*
 * rls_b200.h — C ABI of librls_b200.so, the sm_100a (B200) implementation of the
 * RegularizedLeastSquares.jl iterative-solver inner loop.
 *
 * This is the drop-in boundary: a Julia shim (julia/RLSB200.jl)
 * binds these symbols with `ccall`; the Python host mirror binds them with ctypes.
 * Plain pointers and sizes only — no torch / CUDA types in any signature.
 *
 * Conventions
 *   - every function returns an int32 status (RLS_OK == 0); on failure
 *     rls_last_error() returns a NUL-terminated message (thread-local).  Nothing
 *     aborts or throws across the ABI.  There is NO CPU fallback: without a usable
 *     sm_100 device rls_ctx_create fails with RLS_ERR_CUDA.
 *   - matrices are column-major, element (i,j) at i + j*ld (Julia `Matrix`);
 *     ComplexF32 is interleaved (re,im) float pairs; vectors are contiguous.
 *   - handles are opaque and own device memory; host pointers are borrowed only
 *     for the duration of a call.
 *   - all work of a context is issued on the context's own CUDA stream; calls are
 *     asynchronous unless they return a host scalar / copy to host memory.
 *
 * Each entry point cites the reference interface (file:line under /root/reference)
 * it replaces.
 */
#ifndef RLS_B200_H
#define RLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLS_B200_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------ */
enum {
  RLS_OK = 0,
  RLS_ERR_INVALID = 1,      /* bad argument / shape / dtype mismatch              */
  RLS_ERR_CUDA = 2,         /* CUDA runtime error (message has the CUDA string)   */
  RLS_ERR_COMM = 3,         /* NCCL error or NCCL not loadable                    */
  RLS_ERR_UNSUPPORTED = 4,  /* outside the accelerated path (no CPU fallback)     */
  RLS_ERR_NOMEM = 5
};

/* ---- element types: eltype(A) on the hot path (SURVEY 8b "Types") ------------ */
enum { RLS_F32 = 0, RLS_C32 = 1 };

/* ---- solver kinds: src/{FISTA,POGM,OptISTA,CGNR,ADMM,SplitBregman}.jl (Kaczmarz: rls_kaczmarz_*) */
enum { RLS_FISTA = 0, RLS_POGM = 1, RLS_OPTISTA = 2, RLS_CGNR = 3, RLS_ADMM = 4, RLS_SPLITBREGMAN = 5 };

/* ---- regularisation sinks: src/proximalMaps/Prox{L1,L2,L21,TV,Nuclear,LLR}.jl -- */
enum { RLS_REG_NONE = 0, RLS_REG_L1 = 1, RLS_REG_L2 = 2, RLS_REG_L21 = 3, RLS_REG_TV = 4,
       RLS_REG_NUCLEAR = 5,  /* rls_reg_desc: svtShape in tv_shape[0..1], tv_ndims = 2                            */
       RLS_REG_LLR = 6       /* rls_reg_desc: shape in tv_shape, blockSize in tv_dims, RLS_LLR_* flags in
                                tv_iterations, seed of the randshift stream in slices                              */ };
/* LLRRegularization keywords randshift / fullyOverlapping (ProxLLR.jl:15-16) */
enum { RLS_LLR_RANDSHIFT = 1, RLS_LLR_OVERLAPPING = 2 };

/* ---- projections: src/proximalMaps/Prox{Real,Positive}.jl (bit mask) --------- */
enum { RLS_PROJ_REAL = 1, RLS_PROJ_POSITIVE = 2 };

/* ---- normal-operator forms (mul!(res, AHA, x): FISTA.jl:152, CGNR.jl:151, ...) */
enum {
  RLS_NORMAL_TWOPASS = 0,   /* y = A x ; g = A' y      (2 sweeps over A)          */
  RLS_NORMAL_ONEPASS = 1,   /* fused panel kernel      (1 HBM sweep over A)       */
  RLS_NORMAL_GRAM = 2,      /* g = G x with G = A'A    (reference default form)   */
  RLS_NORMAL_AUTO = 3,
  RLS_NORMAL_MATRIXFREE = 4 /* reported by rls_normal_form for rls_normal_from_linop / rls_normal_from_callback */
};

/* ---- device storage layout of a system matrix ---------------------------------
 * The HOST side is always column-major (Julia `Matrix`).  On the device the library may
 * keep the rows contiguous instead: the normal operator A'(A x) = sum_i conj(a_i)(a_i.x)
 * then needs ONE sweep over HBM (csrc/rls_rowpass.cu).  Upload / download / Philox fills
 * translate; results do not depend on the layout. */
enum { RLS_LAYOUT_COLMAJOR = 0, RLS_LAYOUT_ROWMAJOR = 1,
       RLS_LAYOUT_AUTO = 2 /* row-major whenever the one-pass cluster kernel supports the shape */ };

/* ---- ADMM regTrafo (ADMM.jl:63,74) -------------------------------------------- */
enum { RLS_TRAFO_IDENTITY = 0, RLS_TRAFO_GRADIENT = 1 };
enum { RLS_VARY_RHO_NONE = 0, RLS_VARY_RHO_BALANCE = 1, RLS_VARY_RHO_PNP = 2 };

/* ---- synthetic data distributions for rls_*_fill_philox ---------------------- */
enum {
  RLS_DIST_UNIFORM01 = 0,   /* U[0,1) 24-bit                                       */
  RLS_DIST_IH4 = 1          /* Irwin–Hall(4) centred, unit variance (≈N(0,1)),
                               integer-exact so CPU and GPU agree bit for bit      */
};

typedef struct rls_ctx_s* rls_ctx_t;
typedef struct rls_mat_s* rls_mat_t;
typedef struct rls_vec_s* rls_vec_t;
typedef struct rls_normal_s* rls_normal_t;
typedef struct rls_linop_s* rls_linop_t;   /* matrix-free system operator (SamplingOp, FFTOp, products) */
/* matrix-free AHA as a callback: enqueue res = AHA x on `cuda_stream` (x, res: DEVICE pointers to n elements); 0 = success */
typedef int32_t (*rls_apply_fn)(void* user, const void* x_dev, void* res_dev, void* cuda_stream);
typedef struct rls_solver_s* rls_solver_t;
typedef struct rls_kaczmarz_s* rls_kaczmarz_t;

/* ============================ context ========================================= */
int32_t rls_abi_version(void);
const char* rls_last_error(void);
int32_t rls_device_count(int32_t* count);
/* one context = one device + one stream (+ optionally one NCCL rank) */
int32_t rls_ctx_create(int32_t device, rls_ctx_t* out);
int32_t rls_ctx_destroy(rls_ctx_t ctx);
int32_t rls_ctx_sync(rls_ctx_t ctx);
int32_t rls_ctx_device_info(rls_ctx_t ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                            int64_t* l2_bytes, int64_t* hbm_bytes);
/* CUDA-event timer on the context stream (bench.py; torch events cannot see this stream) */
int32_t rls_timer_start(rls_ctx_t ctx);
int32_t rls_timer_stop(rls_ctx_t ctx, float* elapsed_ms);
/* number of kernels this library launched on ctx since creation (bench "gpu_launches") */
int32_t rls_ctx_launch_count(rls_ctx_t ctx, int64_t* launches);
/* write > L2 bytes to a scratch buffer (timing hygiene for L2-resident configs) */
int32_t rls_ctx_flush_l2(rls_ctx_t ctx);

/* ---- row-sharded multi-GPU: one process per GPU, NCCL over NVLink -- */
/* rank 0 creates the 128-byte id and ships it to the peers with the host's own
 * transport (torch.distributed / MPI.jl); every rank then joins. */
int32_t rls_comm_unique_id(void* id128);
int32_t rls_ctx_comm_init(rls_ctx_t ctx, int32_t rank, int32_t nranks, const void* id128);
int32_t rls_ctx_comm_info(rls_ctx_t ctx, int32_t* rank, int32_t* nranks);
/* optional: one-shot all-reduce over NVLink peer memory, fused with the final sum of the one-pass kernel (replaces
 * finish kernel + ncclAllReduce + copy by one kernel per apply).  Every rank exports a 64-byte CUDA IPC handle of its
 * exchange buffer (vectors of up to max_floats floats), the host gathers the nranks handles with its own transport
 * and every rank imports the table.  Needs peer access between the GPUs (one NVSwitch box). */
int32_t rls_ctx_peer_export(rls_ctx_t ctx, int64_t max_floats, void* handle64);
int32_t rls_ctx_peer_import(rls_ctx_t ctx, const void* handles /* nranks x 64 bytes, rank order */, int32_t nranks);
/* sum-allreduce of a device vector across ranks (the n-vector A_i' r_i) */
int32_t rls_vec_allreduce(rls_vec_t v);
/* sum of n <= 8 host doubles over the ranks, in place: the global scalars of a row-sharded solve — ‖A‖_F²
 * (SystemMatrixBasedNormalization, src/Regularization/NormalizedRegularization.jl:47-58), ‖b‖₁ and length(b)
 * (MeasurementBasedNormalization, :40-43).  No-op on a single rank. */
int32_t rls_ctx_allreduce_f64(rls_ctx_t ctx, double* vals, int32_t n);

/* ============================ vectors ========================================= */
/* `similar(b, n)` on the device: FISTA.jl:94-103, CGNR.jl:91-100, ADMM.jl:166-184 */
int32_t rls_vec_create(rls_ctx_t ctx, int32_t dtype, int64_t len, rls_vec_t* out);
int32_t rls_vec_destroy(rls_vec_t v);
int32_t rls_vec_len(rls_vec_t v, int64_t* len, int32_t* dtype);
int32_t rls_vec_upload(rls_vec_t v, const void* host, int64_t len);       /* async on ctx stream if host is pinned */
int32_t rls_vec_download(rls_vec_t v, void* host, int64_t len);           /* synchronises */
int32_t rls_vec_copy(rls_vec_t dst, rls_vec_t src);
int32_t rls_vec_fill(rls_vec_t v, float re, float im);
int32_t rls_vec_fill_philox(rls_vec_t v, uint64_t seed, uint64_t stream, int32_t dist, float scale,
                            int64_t offset);
/* device pointer escape hatch for zero-copy interop (CuArray / torch tensors) */
int32_t rls_vec_device_ptr(rls_vec_t v, void** ptr);
/* norm(x), dot(a,b)=conj(a).b : Utils / LinearAlgebra call sites FISTA.jl:118,156,172; CGNR.jl:125,153-171 */
int32_t rls_vec_nrm2(rls_vec_t v, double* out);
int32_t rls_vec_asum(rls_vec_t v, double* out);                           /* norm(b,1): NormalizedRegularization.jl:41 */
int32_t rls_vec_dot(rls_vec_t a, rls_vec_t b, double out_re_im[2]);

/* ============================ matrices ======================================== */
/* Dense system matrix (or row shard of it); the host array is column-major.  host==NULL allocates
 * uninitialised device storage (fill with rls_mat_fill_philox or rls_mat_upload).  The device
 * layout is the library's choice (RLS_LAYOUT_AUTO); use rls_mat_create_layout to force one. */
int32_t rls_mat_create(rls_ctx_t ctx, int32_t dtype, int64_t m, int64_t n, const void* host, int64_t ld,
                       rls_mat_t* out);
/* same, choosing the device layout (RLS_LAYOUT_*); host data stays column-major with leading dimension ld */
int32_t rls_mat_create_layout(rls_ctx_t ctx, int32_t dtype, int64_t m, int64_t n, const void* host, int64_t ld,
                              int32_t layout, rls_mat_t* out);
int32_t rls_mat_layout(rls_mat_t A, int32_t* layout);
/* adopt an existing device allocation (CuArray) without copying; not freed by destroy */
int32_t rls_mat_wrap_device(rls_ctx_t ctx, int32_t dtype, int64_t m, int64_t n, void* dev, int64_t ld,
                            rls_mat_t* out);
int32_t rls_mat_destroy(rls_mat_t A);
/* a copy of A in the other device layout, made on the device (one tiled-transpose launch, ~2 x bytes(A) of HBM traffic):
 * how an adopted column-major CuArray (rls_mat_wrap_device — Julia's Matrix layout, src/FISTA.jl:57 takes A as it is
 * stored) reaches the one-pass row kernels: wrap, re-layout, drop the wrapper */
int32_t rls_mat_relayout(rls_mat_t A, int32_t layout, rls_mat_t* out);
int32_t rls_mat_shape(rls_mat_t A, int64_t* m, int64_t* n, int32_t* dtype);
int32_t rls_mat_upload(rls_mat_t A, const void* host, int64_t ld);
int32_t rls_mat_download(rls_mat_t A, void* host, int64_t ld);
/* A[i,j] = scale * dist(philox(seed; counter = (row_offset+i) + j*m_global [, component]))
 * so that any row shard regenerates exactly its rows of the global matrix. */
int32_t rls_mat_fill_philox(rls_mat_t A, uint64_t seed, int32_t dist, float scale, int64_t row_offset,
                            int64_t m_global);
/* sum |A_ij|^2 : SystemMatrixBasedNormalization, NormalizedRegularization.jl:47-58 + Utils.jl:6-20 */
int32_t rls_mat_frob2(rls_mat_t A, double* out);
/* mul!(y, A, x) and mul!(g, adjoint(A), y): FISTA.jl:114, CGNR.jl:132, ADMM.jl:198 */
int32_t rls_gemv_n(rls_mat_t A, rls_vec_t x, rls_vec_t y);
int32_t rls_gemv_c(rls_mat_t A, rls_vec_t y, rls_vec_t g);

/* ============================ normal operator ================================= */
/* AHA: normalOperator(A) (lazy) or A'*A (Gram, FISTA.jl:58).  With a communicator
 * on ctx, A is this rank's row shard and apply() allreduces the n-vector. */
int32_t rls_normal_create(rls_mat_t A, int32_t form, rls_normal_t* out);
/* AHA supplied directly as a dense n×n matrix: createLinearSolver(S; AHA=G) (FISTA.jl:55, CGNR.jl:46) */
int32_t rls_normal_from_gram(rls_mat_t G, rls_normal_t* out);
int32_t rls_normal_destroy(rls_normal_t op);
int32_t rls_normal_form(rls_normal_t op, int32_t* form);
/* kernel plan behind the operator, for diagnostics (e.g. "onepass/tma: grid=148 ...") */
int32_t rls_normal_describe(rls_normal_t op, char* buf, int32_t len);
/* mul!(res, AHA, x): FISTA.jl:152, POGM.jl:181, OptISTA.jl:182, CGNR.jl:151, cg! in ADMM.jl:244 */
int32_t rls_normal_apply(rls_normal_t op, rls_vec_t x, rls_vec_t res);
/* res_k = AHA x_k for K right-hand sides sharing one operator (MultiThreading.jl:45-78 applies it K times).
 * Lazy forms on a row-major A run as two tensor-core GEMMs (Y = A X, G = A' Y; FP32-accurate split-precision
 * tcgen05) that read A once each; other cases fall back to K single applies. */
int32_t rls_normal_apply_batch(rls_normal_t op, int32_t K, const rls_vec_t* xs, const rls_vec_t* outs);
/* diagnostics: internal operands of the last batched apply (0 = packed X, 1 = Y = A X, 2 = A' Y before unpacking) */
int32_t rls_normal_batch_debug(rls_normal_t op, int32_t which, float* host, int64_t nfloats);
/* ---- matrix-free operators ---------------------------------
 * The documentation's examples use operators that are never stored: SamplingOp (examples/compressed_sensing.jl:23), FFTOp
 * and products of them (LinearOperatorCollection).  A solver built on such an operator takes the AHA-only interface the
 * reference already has (`createLinearSolver(S, A; AHA=...)`, solve!(solver, A'b)): the host computes A'b with
 * rls_linop_mul_adjoint and creates the solver with A = NULL and AHA = rls_normal_from_linop(L).
 * pattern_1based: Julia linear indices, no duplicates.  FFTOp: ComplexF32, y = fftshift(fft(ifftshift(x))) / sqrt(N) for
 * shift != 0, unitary != 0 (cuFFT, loaded with dlopen on first use); its adjoint is the inverse transform.
 * rls_normal_from_callback: AHA is the caller's function (a Julia @cfunction around any LinearOperator). */
int32_t rls_linop_sampling_create(rls_ctx_t ctx, int32_t dtype, int64_t n, int64_t npattern, const int64_t* pattern_1based, rls_linop_t* out);
int32_t rls_linop_fft_create(rls_ctx_t ctx, int32_t ndims, const int64_t* shape, int32_t shift, int32_t unitary, rls_linop_t* out);
int32_t rls_linop_compose(rls_linop_t outer, rls_linop_t inner, rls_linop_t* out);   /* y = outer (inner x) */
int32_t rls_linop_destroy(rls_linop_t L);
int32_t rls_linop_shape(rls_linop_t L, int64_t* m, int64_t* n, int32_t* dtype);
int32_t rls_linop_mul(rls_linop_t L, rls_vec_t x, rls_vec_t y);                      /* mul!(y, A, x)           */
int32_t rls_linop_mul_adjoint(rls_linop_t L, rls_vec_t y, rls_vec_t x);              /* mul!(x, adjoint(A), y)  */
int32_t rls_normal_from_linop(rls_linop_t L, rls_normal_t* out);                     /* normalOperator(A)       */
int32_t rls_normal_from_callback(rls_ctx_t ctx, int32_t dtype, int64_t n, rls_apply_fn fn, void* user, rls_normal_t* out);
/* power_iterations(AHA, b; rtol, maxiter) Utils.jl:262-287; b0 replaces the randn start vector */
int32_t rls_power_iterations(rls_normal_t op, rls_vec_t b0, double rtol, int32_t maxiter, double* lambda_max);

/* ============================ proximal maps =================================== */
/* prox!(reg, x, λ) on a device vector; λ already converted to the real type of x
 * (Regularization.jl:31).  ProxL1.jl:18-22, ProxL2.jl:18-21, ProxL21.jl:30-35,
 * ProxTV.jl:89-125 (FGP), ProxPositive.jl:16-20, ProxReal.jl:16-19. */
int32_t rls_prox_l1(rls_vec_t x, float lambda);
int32_t rls_prox_l2(rls_vec_t x, float lambda);
int32_t rls_prox_l21(rls_vec_t x, float lambda, int64_t slices);
int32_t rls_prox_tv(rls_vec_t x, float lambda, int32_t ndims, const int64_t* shape, int32_t ndirs,
                    const int32_t* dims_1based, int32_t iterations_tv);
int32_t rls_prox_positive(rls_vec_t x);
int32_t rls_prox_real(rls_vec_t x);
/* Singular-value soft-thresholding, computed on the device as X·W with W from the Float64
 * eigen-decomposition of the short side's Gram matrix (csrc/rls_svt.cu); the shorter side must be <= 32.
 * prox!(::NuclearRegularization, x, λ), ProxNuclear.jl:27-32: x is the column-major rows x cols matrix (svtShape).
 * prox!(::LLRRegularization, x, λ), ProxLLR.jl:36-90,163-199: x = image series shape x K; every blockSize patch is a
 * (pixels x K) matrix; `shift` (NULL = 0) is the circular shift of the patch grid that `randshift` draws (:55), injected
 * by the caller so that a result can be reproduced; fully_overlapping averages all blockSize shifts (needs
 * shape % blockSize == 0: the reference's padded reshape fails otherwise, :170-176 with :50).  The reference's shortcut
 * "λ >= sqrt(norm(X'X, Inf)) -> patch = 0" (:67-71, `norm` = largest |entry|) is kept. */
int32_t rls_prox_nuclear(rls_vec_t x, float lambda, int64_t rows, int64_t cols);
int32_t rls_prox_llr(rls_vec_t x, float lambda, int32_t ndims, const int64_t* shape, const int64_t* block_size,
                     const int64_t* shift, int32_t fully_overlapping);
/* GradientOp(T; shape, dims) forward / transpose (LinearOperatorCollection; ProxTV.jl:46, ADMM.jl:74) */
int32_t rls_grad_rows(int32_t ndims, const int64_t* shape, int32_t ndirs, const int32_t* dims_1based, int64_t* rows);
int32_t rls_grad_apply(rls_vec_t img, rls_vec_t out, int32_t ndims, const int64_t* shape, int32_t ndirs,
                       const int32_t* dims_1based);
int32_t rls_grad_apply_t(rls_vec_t g, rls_vec_t out, int32_t ndims, const int64_t* shape, int32_t ndirs,
                         const int32_t* dims_1based);

/* ============================ solvers ========================================= */
#define RLS_MAX_TV_DIMS 4

/* One regularisation term after the Julia-side decorator plumbing has been resolved
 * (λ(reg) incl. normalisation factor; Regularization/*.jl stays in the host). */
typedef struct {
  int32_t kind;              /* RLS_REG_*                                          */
  int32_t lambda_is_f64;     /* λ typed Float64 upstream: thresholds formed in double then convert(T,·) */
  double lambda;             /* λ(reg)                                             */
  int64_t slices;            /* L21                                                */
  int32_t tv_ndims;          /* TV: image shape / directions / FGP iterations      */
  int32_t tv_ndirs;
  int64_t tv_shape[RLS_MAX_TV_DIMS];
  int32_t tv_dims[RLS_MAX_TV_DIMS];
  int32_t tv_iterations;
  int32_t trafo;             /* ADMM regTrafo: RLS_TRAFO_*; gradient uses tv_shape/tv_dims */
  float rho;                 /* ADMM penalty for this term                         */
  int32_t _pad;
} rls_reg_desc;

/* Constructor keywords of FISTA.jl:57-67, POGM.jl:75-86, OptISTA.jl:62-71,
 * CGNR.jl:48-53, ADMM.jl:80-94 after `rT(...)` conversion. */
typedef struct {
  int32_t kind;              /* RLS_FISTA ...                                      */
  int32_t iterations;
  int32_t restart;           /* 0 = :none, 1 = :gradient (FISTA, POGM)             */
  int32_t proj_mask;         /* RLS_PROJ_* applied after prox (CGNR: at termination) */
  float rho;                 /* step size (FISTA/POGM/OptISTA)                     */
  float theta;
  float sigma_fac;           /* POGM                                               */
  float rel_tol;
  float abs_tol;             /* ADMM                                               */
  float tol_inner;           /* ADMM cg! reltol                                    */
  int32_t iterations_cg;     /* ADMM                                               */
  int32_t vary_rho;          /* ADMM RLS_VARY_RHO_*                                */
  int32_t n_reg;             /* 1 (ADMM / SplitBregman: 1..4)                      */
  int32_t iterations_inner;  /* SplitBregman: inner iterations per Bregman update; `iterations` counts the outer ones */
  rls_reg_desc reg[4];
} rls_solver_desc;

/* Scalars of the solver state structs (FISTAState FISTA.jl:15-27, POGMState POGM.jl:15-36,
 * OptISTAState OptISTA.jl:15-33, CGNRState CGNR.jl:13-24, ADMMState ADMM.jl:19-46) */
typedef struct {
  int32_t iteration;
  int32_t done;              /* done(solver,state) evaluated for the NEXT iterate   */
  float rho, theta, theta_old, theta_n, alpha, beta, gamma, gamma_old, sigma;
  float norm_x0;             /* FISTA norm_x₀ / CGNR z0                             */
  float rel_res_norm;        /* FISTA-family; CGNR: ‖r‖/z0                          */
  float res_norm;            /* ‖res‖ (solverconvergence)                           */
  float cg_alpha[2], cg_beta[2], cg_zeta[2];      /* CGNR αl, βl, ζl (Tc)           */
  float admm_rk[4], admm_sk[4], admm_eps_pri[4], admm_eps_dua[4], admm_delta[4], admm_rho[4];
  float admm_sigma_abs;
  int32_t cg_iterations_last;                     /* ADMM: inner CG steps of the last outer iteration */
  int32_t cg_iterations_total;
  int32_t outer_iteration;                        /* SplitBregman: iter_cnt (SplitBregman.jl:35) */
} rls_solver_scalars;

/* createLinearSolver(S, A; AHA=op, kwargs...) RegularizedLeastSquares.jl:288-294.
 * A may be NULL when only AHA is supplied (FISTA.jl:55); then b is A'b (FISTA.jl:112). */
int32_t rls_solver_create(rls_mat_t A, rls_normal_t AHA, const rls_solver_desc* desc, rls_solver_t* out);
int32_t rls_solver_destroy(rls_solver_t s);
/* update λ / ρ after host-side normalisation in init! (FISTA.jl:128) without re-creating */
int32_t rls_solver_set_reg(rls_solver_t s, int32_t idx, const rls_reg_desc* reg);
/* init!(solver, state, b; x0): FISTA.jl:110-129, POGM.jl:138-164, OptISTA.jl:128-155,
 * CGNR.jl:107-130, ADMM.jl:191-220.  b and x0 are device vectors (x0 may be NULL = 0). */
int32_t rls_solver_init(rls_solver_t s, rls_vec_t b, rls_vec_t x0);
/* iterate(solver, state): FISTA.jl:139-185, POGM.jl:173-237, OptISTA.jl:164-204,
 * CGNR.jl:143-178, ADMM.jl:230-322, SplitBregman.jl:203-272.  *advanced = 0 when done() was already true
 * (Julia `nothing`).  Synchronises to return the scalars. */
int32_t rls_solver_iterate(rls_solver_t s, int32_t* advanced, rls_solver_scalars* scalars);
/* the `for _ in enumerate(solver)` loop of solve! for an initialised solver: enqueue the
 * remaining iterations with device-side done() gating, synchronise once. */
int32_t rls_solver_run(rls_solver_t s, int32_t* iterations_done, rls_solver_scalars* scalars);
/* callback-free solve!: enqueue up to `iterations` iterations with device-side
 * done() gating (no host round trip per iteration), then synchronise once.
 * RegularizedLeastSquares.jl:103-117 with the default no-op callback. */
int32_t rls_solver_solve(rls_solver_t s, rls_vec_t b, rls_vec_t x0, int32_t* iterations_done,
                         rls_solver_scalars* scalars);
/* same, host buffers in and out (b: length m or n host elements, x: length n) */
int32_t rls_solver_solve_host(rls_solver_t s, const void* b_host, int64_t b_len, void* x_host, int64_t x_len,
                              int32_t* iterations_done, rls_solver_scalars* scalars);
int32_t rls_solver_scalars_get(rls_solver_t s, rls_solver_scalars* scalars);
/* state vectors by name: "x","x0","xold","res","y","z","zold","w","p","v","beta","beta_y","z0".."u0".. (borrowed handle) */
int32_t rls_solver_vec(rls_solver_t s, const char* name, rls_vec_t* out);

/* ---- multi right-hand-side solve (MultiThreading.jl:30-80) ---------------------- */
/* K independent states sharing A; B is m×K column-major on the host, X n×K out.
 * Per-column done() masks are kept on the device; iterations run until no column
 * is active.  iterations_done[k] returns each column's count. */
int32_t rls_solver_solve_batch_host(rls_solver_t s, const void* B_host, int64_t ldb, int32_t K, void* X_host,
                                    int64_t ldx, int32_t* iterations_done);

/* ---- Kaczmarz row-action solver (src/Kaczmarz.jl) ----------------- */
/* The row loop of iterate(::Kaczmarz) (Kaczmarz.jl:270-273, row step :305-310) on a ROW-MAJOR device matrix, evaluated
 * block-wise: for `block_rows` consecutive rows of the visiting order the projections are resolved exactly through the
 * block Gram matrix A_blk A_blk^H (built once per order), so that one iteration is one HBM sweep over A instead of m
 * dependent dot/axpy pairs.  The constructor logic (L2 / denom / rowindex / probabilities / row order, Kaczmarz.jl:73-159,
 * :326-392) and the prox! calls after the sweep (:275-277, rls_prox_*) stay with the host, as in the reference.
 * block_rows: 64, 128, 192, 256, or 0 = the library's choice (128; 64 for systems of at most 64 rows).  With 64 or 128 and rows
 * of a multiple of 16 bytes a whole iteration is ONE cooperative kernel (columns of x pinned to CTAs, one grid-wide
 * exchange per block); otherwise three kernels per block are chained on the stream. */
int32_t rls_kaczmarz_create(rls_mat_t A, int32_t block_rows, rls_kaczmarz_t* out);
int32_t rls_kaczmarz_destroy(rls_kaczmarz_t K);
int32_t rls_kaczmarz_block_rows(rls_kaczmarz_t K, int32_t* block_rows);
/* rownorm²(A, i) for all rows (Utils.jl:16-23; initkaczmarz Kaczmarz.jl:365-376, rowProbabilities :326-334) */
int32_t rls_kaczmarz_rownorm2(rls_kaczmarz_t K, float* host, int64_t len);
/* the visiting order of one iteration: rows[i] = rowindex[usedIndices[i]] (0-based, distinct), denom[i] as in
 * Kaczmarz.jl:372.  Re-call after shuffle! (:201) or sample! (:268); rebuilds the block Gram matrices. */
int32_t rls_kaczmarz_set_rows(rls_kaczmarz_t K, const int64_t* rows, const float* denom, int64_t count);
/* init!(solver, state, b; x0): x = x0 (NULL = 0), vl = 0, u = b, eps_w = sqrt(λ) (Kaczmarz.jl:205-215) */
int32_t rls_kaczmarz_init(rls_kaczmarz_t K, rls_vec_t b, rls_vec_t x0, float eps_w);
/* `for i in usedIndices; iterate_row_index(...)` (Kaczmarz.jl:270-273); asynchronous on the context stream */
int32_t rls_kaczmarz_sweep(rls_kaczmarz_t K);
/* state vectors "x", "vl", "u" (borrowed handles) */
int32_t rls_kaczmarz_vec(rls_kaczmarz_t K, const char* name, rls_vec_t* out);
/* synchronise; reports a timed-out block exchange of the one-kernel sweep (every device-side wait is bounded) */
int32_t rls_kaczmarz_check(rls_kaczmarz_t K);
/* kernel plan of the sweep, for diagnostics ("persistent: grid=148 ..." / "chained: 3 kernels per block ...") */
int32_t rls_kaczmarz_describe(rls_kaczmarz_t K, char* buf, int32_t len);
/* diagnostics: 0 = block Gram matrices, 1 = dot partials of the last block, 2 = alpha of the last block, 3 = denominators */
int32_t rls_kaczmarz_debug(rls_kaczmarz_t K, int32_t which, float* host, int64_t nfloats);

/* ---- single-process multi-device ("multi-GPU handled inside one call") --------------------------------
 * What createLinearSolver(FISTA, A; ...) + solve!(solver, b) (src/RegularizedLeastSquares.jl:288-294, :103-117) bind
 * when the host is ONE process and the system is larger than one GPU: a group of devices with one NCCL communicator
 * (ncclCommInitAll), a matrix row-partitioned over them, and a solve! that takes the whole b and returns x.  Inside a
 * call one host thread per device runs the single-context entry points above (NCCL thread-per-rank). */
typedef struct rls_group_s* rls_group_t;
typedef struct rls_gmat_s* rls_gmat_t;
typedef struct rls_gsolver_s* rls_gsolver_t;
int32_t rls_group_create(int32_t ndev, const int32_t* dev_ids /* NULL: devices 0..ndev-1 */, rls_group_t* out);
int32_t rls_group_destroy(rls_group_t g);
int32_t rls_group_size(rls_group_t g, int32_t* ndev);
int32_t rls_group_ctx(rls_group_t g, int32_t i, rls_ctx_t* ctx);                       /* borrowed */
/* A (m x n): column-major host matrix (Julia Matrix), or host = NULL and rls_group_mat_fill_philox; contiguous row
 * blocks (multiples of 4 rows), device layout chosen per block like rls_mat_create */
int32_t rls_group_mat_create(rls_group_t g, int32_t dtype, int64_t m, int64_t n, const void* host, int64_t ld, rls_gmat_t* out);
int32_t rls_group_mat_fill_philox(rls_gmat_t A, uint64_t seed, int32_t dist, float scale);
int32_t rls_group_mat_part(rls_gmat_t A, int32_t i, rls_mat_t* part /* borrowed */, int64_t* row_lo, int64_t* row_hi);
int32_t rls_group_mat_destroy(rls_gmat_t A);
int32_t rls_group_solver_create(rls_gmat_t A, int32_t normal_form, const rls_solver_desc* desc, rls_gsolver_t* out);
int32_t rls_group_solver_destroy(rls_gsolver_t s);
/* solve!(solver, b): host b (m elements) in, host x (n elements) out; fails if the replicas are not bit-identical */
int32_t rls_group_solver_solve_host(rls_gsolver_t s, const void* b_host, int64_t b_len, void* x_host, int64_t x_len,
                                    int32_t* iterations_done, rls_solver_scalars* scalars);

#ifdef __cplusplus
}
#endif
#endif /* RLS_B200_H */
