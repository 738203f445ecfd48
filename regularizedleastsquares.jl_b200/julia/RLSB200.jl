# RLSB200.jl — the reference-side binding of librls_b200.so.
#
# A package extension of RegularizedLeastSquares.jl that adds *methods*, not new API
# (the extension point documented in docs/src/solvers.md:85-98 and used by
# ext/RegularizedLeastSquaresGPUArraysExt): device array types `B200Matrix` / `B200Vector`,
# `similar` so that `init!` re-allocates solver state on the device (src/FISTA.jl:94-103),
# and `init!` / `iterate` / `solve!` / `prox!` methods that `ccall` the C ABI of
# include/rls_b200.h.  NOT EXECUTED in this repository's CI: Julia is not installed in the
# build image; the Python mirror (regularizedleastsquares.jl_b200/*.py) makes exactly the
# same call sequence and is what the tests drive.
module RLSB200

using RegularizedLeastSquares
using LinearAlgebra
using Random, StatsBase                 # Kaczmarz: sample! of the randomised row order stays in Julia
import RegularizedLeastSquares: init!, iterate, solve!, prox!, solversolution, solverconvergence,
       L1Regularization, L2Regularization, L21Regularization, TVRegularization, PositiveRegularization,
       RealRegularization, FISTA, POGM, OptISTA, CGNR, ADMM, SplitBregman, Kaczmarz, λ, sink

const LIB = get(ENV, "RLS_B200_LIB", "librls_b200.so")

# ---- status handling: never let a C error pass silently --------------------------------------
struct RlsError <: Exception
  status::Int32
  msg::String
end
function check(status::Int32)
  status == 0 && return nothing
  throw(RlsError(status, unsafe_string(ccall((:rls_last_error, LIB), Cstring, ()))))
end

# ---- enums of include/rls_b200.h ----------------------------------------------------------------
const RLS_F32, RLS_C32 = Int32(0), Int32(1)
const RLS_FISTA, RLS_POGM, RLS_OPTISTA, RLS_CGNR, RLS_ADMM, RLS_SPLITBREGMAN = Int32.(0:5)
const RLS_REG_NONE, RLS_REG_L1, RLS_REG_L2, RLS_REG_L21, RLS_REG_TV, RLS_REG_NUCLEAR, RLS_REG_LLR = Int32.(0:6)
const RLS_LLR_RANDSHIFT, RLS_LLR_OVERLAPPING = Int32(1), Int32(2)
const RLS_PROJ_REAL, RLS_PROJ_POSITIVE = Int32(1), Int32(2)
const RLS_NORMAL_AUTO = Int32(3)
dtypecode(::Type{Float32}) = RLS_F32
dtypecode(::Type{ComplexF32}) = RLS_C32
dtypecode(T) = error("librls_b200 accelerates Float32 / ComplexF32 only, got $T (no CPU fallback)")

# ---- POD structs (layout checked against the header by tests/test_abi.py) ----------------------
struct RegDesc
  kind::Int32; lambda_is_f64::Int32; lambda::Float64; slices::Int64
  tv_ndims::Int32; tv_ndirs::Int32; tv_shape::NTuple{4,Int64}; tv_dims::NTuple{4,Int32}
  tv_iterations::Int32; trafo::Int32; rho::Float32; _pad::Int32
end
struct SolverDesc
  kind::Int32; iterations::Int32; restart::Int32; proj_mask::Int32
  rho::Float32; theta::Float32; sigma_fac::Float32; rel_tol::Float32; abs_tol::Float32; tol_inner::Float32
  iterations_cg::Int32; vary_rho::Int32; n_reg::Int32; iterations_inner::Int32   # iterations_inner: SplitBregman
  reg::NTuple{4,RegDesc}
end
struct SolverScalars
  iteration::Int32; done::Int32
  rho::Float32; theta::Float32; theta_old::Float32; theta_n::Float32; alpha::Float32; beta::Float32
  gamma::Float32; gamma_old::Float32; sigma::Float32; norm_x0::Float32; rel_res_norm::Float32; res_norm::Float32
  cg_alpha::NTuple{2,Float32}; cg_beta::NTuple{2,Float32}; cg_zeta::NTuple{2,Float32}
  admm_rk::NTuple{4,Float32}; admm_sk::NTuple{4,Float32}; admm_eps_pri::NTuple{4,Float32}
  admm_eps_dua::NTuple{4,Float32}; admm_delta::NTuple{4,Float32}; admm_rho::NTuple{4,Float32}
  admm_sigma_abs::Float32; cg_iterations_last::Int32; cg_iterations_total::Int32; outer_iteration::Int32   # SplitBregman iter_cnt
end

# ---- context (one per device) -------------------------------------------------------------------
mutable struct Context
  handle::Ptr{Cvoid}
  function Context(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rls_ctx_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h))
    finalizer(c -> ccall((:rls_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), new(h[]))
  end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
ctx() = something(DEFAULT_CTX[], (DEFAULT_CTX[] = Context(0)))

# ---- device arrays -------------------------------------------------------------------------------
mutable struct B200Vector{T} <: AbstractVector{T}
  handle::Ptr{Cvoid}; len::Int; owned::Bool
end
function B200Vector{T}(::UndefInitializer, n::Integer) where T
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:rls_vec_create, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ref{Ptr{Cvoid}}), ctx().handle, dtypecode(T), n, h))
  finalizer(v -> v.owned && ccall((:rls_vec_destroy, LIB), Int32, (Ptr{Cvoid},), v.handle), B200Vector{T}(h[], n, true))
end
Base.size(v::B200Vector) = (v.len,)
Base.similar(v::B200Vector{T}, n::Integer...) where T = B200Vector{T}(undef, prod(n))   # FISTA.jl:95-98
Base.similar(v::B200Vector, ::Type{T}, dims::Dims) where T = B200Vector{T}(undef, prod(dims))
function B200Vector(x::Vector{T}) where T
  v = B200Vector{T}(undef, length(x))
  GC.@preserve x check(ccall((:rls_vec_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), v.handle, pointer(x), length(x)))
  v
end
function Base.Array(v::B200Vector{T}) where T
  x = Vector{T}(undef, v.len)
  GC.@preserve x check(ccall((:rls_vec_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), v.handle, pointer(x), v.len))
  x
end
Base.getindex(v::B200Vector, i::Int) = error("scalar indexing of a B200Vector is disabled; use Array(v)")
function LinearAlgebra.norm(v::B200Vector)
  r = Ref{Float64}(0); check(ccall((:rls_vec_nrm2, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}), v.handle, r)); real(eltype(v))(r[])
end

mutable struct B200Matrix{T} <: AbstractMatrix{T}
  handle::Ptr{Cvoid}; m::Int; n::Int
end
# The Julia side hands over its column-major `Matrix`; the DEVICE layout is the library's choice
# (layout = :auto -> rows contiguous whenever the one-pass cluster kernel supports the shape, so that
# A'(A x) sweeps HBM once; :col mirrors the host storage; :row forces it).  Upload / download translate.
const LAYOUTS = Dict(:col => Int32(0), :row => Int32(1), :auto => Int32(2))
function B200Matrix(A::Matrix{T}; layout::Symbol = :auto) where T
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve A check(ccall((:rls_mat_create_layout, LIB), Int32,
                             (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Cvoid}, Int64, Int32, Ref{Ptr{Cvoid}}),
                             ctx().handle, dtypecode(T), size(A, 1), size(A, 2), pointer(A), stride(A, 2), LAYOUTS[layout], h))
  finalizer(M -> ccall((:rls_mat_destroy, LIB), Int32, (Ptr{Cvoid},), M.handle), B200Matrix{T}(h[], size(A)...))
end
# adopt a CuArray without copying (column-major, CUDA.jl's layout): two-sweep / TMA panel kernels
function B200Matrix(dA::Ptr{Cvoid}, ::Type{T}, m::Int, n::Int, ld::Int) where T
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:rls_mat_wrap_device, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}),
              ctx().handle, dtypecode(T), m, n, dA, ld, h))
  finalizer(M -> ccall((:rls_mat_destroy, LIB), Int32, (Ptr{Cvoid},), M.handle), B200Matrix{T}(h[], m, n))
end
Base.size(A::B200Matrix) = (A.m, A.n)

"AHA for a device matrix: `normalOperator(A)` / `A'*A` both resolve to the library's normal operator"
mutable struct B200NormalOp{T}
  handle::Ptr{Cvoid}; n::Int; keep::Any      # keep: what the operator was built from (a matrix, a matrix-free operator, a callback)
end
function B200NormalOp(A::B200Matrix{T}; form = RLS_NORMAL_AUTO) where T
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:rls_normal_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), A.handle, form, h))
  finalizer(o -> ccall((:rls_normal_destroy, LIB), Int32, (Ptr{Cvoid},), o.handle), B200NormalOp{T}(h[], A.n, A))
end
Base.:*(At::Adjoint{T,B200Matrix{T}}, A::B200Matrix{T}) where T = B200NormalOp(parent(At))   # FISTA.jl:58 `AHA = A'*A`
Base.eltype(::B200NormalOp{T}) where T = T
Base.size(op::B200NormalOp, d...) = op.n
function LinearAlgebra.mul!(res::B200Vector, op::B200NormalOp, x::B200Vector)                # FISTA.jl:152
  check(ccall((:rls_normal_apply, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), op.handle, x.handle, res.handle)); res
end
# K right-hand sides at once (MultiThreading.jl:45-78 applies AHA per column): two tensor-core GEMMs
function mul_batch!(res::Vector{<:B200Vector}, op::B200NormalOp, xs::Vector{<:B200Vector})
  xh = Ptr{Cvoid}[x.handle for x in xs]; rh = Ptr{Cvoid}[r.handle for r in res]
  GC.@preserve xh rh check(ccall((:rls_normal_apply_batch, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                                 op.handle, length(xs), pointer(xh), pointer(rh)))
  res
end
function LinearAlgebra.mul!(g::B200Vector, At::Adjoint{T,B200Matrix{T}}, y::B200Vector) where T  # FISTA.jl:114
  check(ccall((:rls_gemv_c, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), parent(At).handle, y.handle, g.handle)); g
end

# ---- prox! on device vectors (mirrors ext/RegularizedLeastSquaresGPUArraysExt) --------------------
prox!(::L1Regularization, x::B200Vector, λ::Float32) = (check(ccall((:rls_prox_l1, LIB), Int32, (Ptr{Cvoid}, Float32), x.handle, λ)); x)
prox!(::L2Regularization, x::B200Vector, λ::Float32) = (check(ccall((:rls_prox_l2, LIB), Int32, (Ptr{Cvoid}, Float32), x.handle, λ)); x)
prox!(r::L21Regularization, x::B200Vector, λ::Float32) = (check(ccall((:rls_prox_l21, LIB), Int32, (Ptr{Cvoid}, Float32, Int64), x.handle, λ, r.slices)); x)
function prox!(r::TVRegularization, x::B200Vector, λ::Float32)
  shape = collect(Int64, r.shape); dims = collect(Int32, r.dims)
  check(ccall((:rls_prox_tv, LIB), Int32, (Ptr{Cvoid}, Float32, Int32, Ptr{Int64}, Int32, Ptr{Int32}, Int32),
              x.handle, λ, length(shape), shape, length(dims), dims, r.iterationsTV)); x
end
# singular-value thresholding on the device (the reference moves GPU arrays to the CPU for these: ProxLLR.jl:7)
function prox!(r::NuclearRegularization, x::B200Vector, λ::Float32)                 # ProxNuclear.jl:27-32
  check(ccall((:rls_prox_nuclear, LIB), Int32, (Ptr{Cvoid}, Float32, Int64, Int64), x.handle, λ, r.svtShape[1], r.svtShape[2])); x
end
function prox!(r::LLRRegularization, x::B200Vector, λ::Float32)                     # ProxLLR.jl:36-38
  shape = collect(Int64, r.shape); block = collect(Int64, r.blockSize)
  shift = r.randshift ? Int64[rand(1:b) for b in block] : zeros(Int64, length(block))   # rand(block_idx), ProxLLR.jl:55
  check(ccall((:rls_prox_llr, LIB), Int32, (Ptr{Cvoid}, Float32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32),
              x.handle, λ, length(shape), shape, block, shift, r.fullyOverlapping ? 1 : 0)); x
end
prox!(::PositiveRegularization, x::B200Vector) = (check(ccall((:rls_prox_positive, LIB), Int32, (Ptr{Cvoid},), x.handle)); x)
prox!(::RealRegularization, x::B200Vector) = (check(ccall((:rls_prox_real, LIB), Int32, (Ptr{Cvoid},), x.handle)); x)

# ---- matrix-free system operators (SamplingOp, FFTOp, products; compressed_sensing.jl:23) -------------
mutable struct B200LinOp{T}
  handle::Ptr{Cvoid}; m::Int; n::Int; keep::Any
end
function B200SamplingOp(::Type{T}; pattern::AbstractVector{<:Integer}, shape) where T
  h = Ref{Ptr{Cvoid}}(); pat = collect(Int64, pattern)
  check(ccall((:rls_linop_sampling_create, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Ptr{Ptr{Cvoid}}),
              ctx().handle, T <: Complex ? RLS_C32 : RLS_F32, prod(shape), length(pat), pat, h))
  L = B200LinOp{T}(h[], length(pat), prod(shape), nothing)
  finalizer(l -> ccall((:rls_linop_destroy, LIB), Int32, (Ptr{Cvoid},), l.handle), L)
end
function B200FFTOp(::Type{ComplexF32}; shape, shift::Bool = true, unitary::Bool = true)
  h = Ref{Ptr{Cvoid}}(); shp = collect(Int64, shape)
  check(ccall((:rls_linop_fft_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, Int32, Int32, Ptr{Ptr{Cvoid}}),
              ctx().handle, length(shp), shp, shift, unitary, h))
  L = B200LinOp{ComplexF32}(h[], prod(shape), prod(shape), nothing)
  finalizer(l -> ccall((:rls_linop_destroy, LIB), Int32, (Ptr{Cvoid},), l.handle), L)
end
function Base.:*(A::B200LinOp{T}, B::B200LinOp{T}) where T                                 # ProdOp
  h = Ref{Ptr{Cvoid}}()
  check(ccall((:rls_linop_compose, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), A.handle, B.handle, h))
  L = B200LinOp{T}(h[], A.m, B.n, (A, B))
  finalizer(l -> ccall((:rls_linop_destroy, LIB), Int32, (Ptr{Cvoid},), l.handle), L)
end
LinearAlgebra.mul!(y::B200Vector, A::B200LinOp, x::B200Vector) =
  (check(ccall((:rls_linop_mul, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), A.handle, x.handle, y.handle)); y)
LinearAlgebra.mul!(x::B200Vector, At::Adjoint{T,B200LinOp{T}}, y::B200Vector) where T =
  (check(ccall((:rls_linop_mul_adjoint, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), parent(At).handle, y.handle, x.handle)); x)
# normalOperator(A) of a matrix-free A; createLinearSolver(S, A::B200LinOp; AHA = normalOperator(A)) then takes the AHA-only
# path (rls_solver_create with A = C_NULL) and solve!(solver, A' * b)
function LinearOperatorCollection.normalOperator(A::B200LinOp{T}) where T
  h = Ref{Ptr{Cvoid}}()
  check(ccall((:rls_normal_from_linop, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), A.handle, h))
  finalizer(o -> ccall((:rls_normal_destroy, LIB), Int32, (Ptr{Cvoid},), o.handle), B200NormalOp{T}(h[], A.n, A))
end
# any other LinearOperator (NFFT, Radon, wavelets on the GPU, ...): AHA as a @cfunction on device pointers + stream
#   apply(user, x::Ptr{Cvoid}, res::Ptr{Cvoid}, stream::Ptr{Cvoid})::Int32 = (my_gpu_normal_op!(res, x, stream); Int32(0))
#   fn = @cfunction(apply, Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}))
function B200NormalOp(fn::Ptr{Cvoid}, ::Type{T}, n::Integer; user::Ptr{Cvoid} = C_NULL) where T
  h = Ref{Ptr{Cvoid}}()
  check(ccall((:rls_normal_from_callback, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
              ctx().handle, T <: Complex ? RLS_C32 : RLS_F32, n, fn, user, h))
  finalizer(o -> ccall((:rls_normal_destroy, LIB), Int32, (Ptr{Cvoid},), o.handle), B200NormalOp{T}(h[], n, (fn, user)))
end

# ---- solvers: one library-side solver object per Julia solver, keyed by objectid ---------------------
regkind(::L1Regularization) = RLS_REG_L1; regkind(::L2Regularization) = RLS_REG_L2
regkind(::L21Regularization) = RLS_REG_L21; regkind(::TVRegularization) = RLS_REG_TV
regkind(::NuclearRegularization) = RLS_REG_NUCLEAR; regkind(::LLRRegularization) = RLS_REG_LLR
pad4(T, v) = ntuple(i -> i <= length(v) ? T(collect(v)[i]) : T(0), 4)
function regdesc(reg; rho = 0f0)
  s = sink(reg); l = λ(reg)
  if s isa NuclearRegularization      # svtShape in tv_shape[1:2] (include/rls_b200.h, RLS_REG_NUCLEAR)
    return RegDesc(RLS_REG_NUCLEAR, l isa Float64 ? 1 : 0, Float64(l), 1, 2, 0, pad4(Int64, s.svtShape), pad4(Int32, ()), 0, 0, Float32(rho), 0)
  elseif s isa LLRRegularization      # shape / blockSize / flags / seed of the shift stream (RLS_REG_LLR)
    flags = (s.randshift ? RLS_LLR_RANDSHIFT : Int32(0)) | (s.fullyOverlapping ? RLS_LLR_OVERLAPPING : Int32(0))
    return RegDesc(RLS_REG_LLR, l isa Float64 ? 1 : 0, Float64(l), rand(Int64(1):Int64(2)^31), length(s.shape), length(s.blockSize),
                   pad4(Int64, s.shape), pad4(Int32, s.blockSize), flags, 0, Float32(rho), 0)
  end
  shape = s isa TVRegularization ? ntuple(i -> i <= length(s.shape) ? Int64(s.shape[i]) : Int64(0), 4) : ntuple(_ -> Int64(0), 4)
  dims = s isa TVRegularization ? ntuple(i -> i <= length(s.dims) ? Int32(collect(s.dims)[i]) : Int32(0), 4) : ntuple(_ -> Int32(0), 4)
  RegDesc(regkind(s), l isa Float64 ? 1 : 0, Float64(l), s isa L21Regularization ? s.slices : 1,
          s isa TVRegularization ? length(s.shape) : 0, s isa TVRegularization ? length(s.dims) : 0, shape, dims,
          s isa TVRegularization ? s.iterationsTV : 0, 0, Float32(rho), 0)
end
projmask(proj) = reduce(|, (p isa PositiveRegularization ? RLS_PROJ_POSITIVE : RLS_PROJ_REAL for p in proj); init = Int32(0))
const EMPTYREG = RegDesc(0, 0, 0.0, 1, 0, 0, ntuple(_ -> Int64(0), 4), ntuple(_ -> Int32(0), 4), 0, 0, 0f0, 0)

const HANDLES = IdDict{Any,Ptr{Cvoid}}()
function handle!(solver::FISTA, state)        # POGM / OptISTA / CGNR / ADMM / SplitBregman are built the same way
  get!(HANDLES, solver) do
    d = SolverDesc(RLS_FISTA, solver.iterations, solver.restart == :gradient ? 1 : 0, projmask(solver.proj),
                   state.ρ, state.theta, 1f0, state.relTol, 0f0, 0f0, 0, 0, 1, 0,
                   (regdesc(solver.reg), EMPTYREG, EMPTYREG, EMPTYREG))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rls_solver_create, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{SolverDesc}, Ref{Ptr{Cvoid}}),
                solver.A.handle, solver.AHA.handle, Ref(d), h))
    h[]
  end
end

# state-type dispatch: FISTAState{rT, <:B200Vector} is what `init!` builds once `b isa B200Vector`
function init!(solver::FISTA, state::RegularizedLeastSquares.FISTAState{rT,<:B200Vector}, b::B200Vector; x0 = 0, theta = 1) where rT
  check(ccall((:rls_solver_init, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), handle!(solver, state), b.handle, C_NULL))
  sync_scalars!(solver, state)
end
function iterate(solver::FISTA, state::RegularizedLeastSquares.FISTAState{rT,<:B200Vector}) where rT
  adv = Ref{Int32}(0); sc = Ref{SolverScalars}()
  check(ccall((:rls_solver_iterate, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}, Ref{SolverScalars}), handle!(solver, state), adv, sc))
  adv[] == 0 && return nothing
  state.theta = sc[].theta; state.thetaᵒˡᵈ = sc[].theta_old; state.rel_res_norm = sc[].rel_res_norm; state.iteration = sc[].iteration
  return state.x, state
end
# callback-free solve!: one C call for the whole loop, host buffers in and out
function solve!(solver::FISTA, b::Vector{T}; callbacks = nothing, kwargs...) where T <: Union{Float32,ComplexF32}
  callbacks === nothing || return invoke(solve!, Tuple{RegularizedLeastSquares.AbstractLinearSolver,Any}, solver, B200Vector(b); callbacks, kwargs...)
  x = Vector{T}(undef, size(solver.AHA, 2)); it = Ref{Int32}(0); sc = Ref{SolverScalars}()
  GC.@preserve b x check(ccall((:rls_solver_solve_host, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ref{Int32}, Ref{SolverScalars}),
      handle!(solver, solver.state), pointer(b), length(b), pointer(x), length(x), it, sc))
  x
end

function sync_scalars!(solver, state)
  sc = Ref{SolverScalars}()
  check(ccall((:rls_solver_scalars_get, LIB), Int32, (Ptr{Cvoid}, Ref{SolverScalars}), HANDLES[solver], sc))
  state.norm_x₀ = sc[].norm_x0; state.iteration = sc[].iteration
  state
end

# ---- Kaczmarz (src/Kaczmarz.jl): the row loop of iterate runs in librls_b200, everything else stays as upstream -------
# The solver is constructed on the HOST matrix as usual (L2 / Tikhonov handling, denom, rowindex, probabilities:
# Kaczmarz.jl:73-159); `B200Kaczmarz(solver)` uploads solver.A with rows contiguous and owns the device-side plan.
mutable struct B200Kaczmarz{T}
  handle::Ptr{Cvoid}
  A::B200Matrix{T}
  order::Vector{Int64}        # the visiting order last sent to the device (usedIndices mapped through rowindex)
end
function B200Kaczmarz(solver::Kaczmarz; block_rows::Integer = 0)
  T = eltype(solver.A)
  A = B200Matrix(Matrix(solver.A); layout = :row)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:rls_kaczmarz_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), A.handle, Int32(block_rows), h))
  k = B200Kaczmarz{T}(h[], A, Int64[])
  finalizer(x -> ccall((:rls_kaczmarz_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), k)
end
function set_rows!(k::B200Kaczmarz, solver::Kaczmarz, usedIndices)
  rows = Int64[solver.rowindex[i] - 1 for i in usedIndices]        # 0-based, distinct
  rows == k.order && return k
  denom = Float32[solver.denom[i] for i in usedIndices]
  GC.@preserve rows denom check(ccall((:rls_kaczmarz_set_rows, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float32}, Int64),
                                      k.handle, rows, denom, length(rows)))
  k.order = rows
  k
end
# init!(solver, state, b; x0): Kaczmarz.jl:178-216 runs first (normalisation, denom, shuffle); then the device state
function init!(k::B200Kaczmarz, solver::Kaczmarz, b::B200Vector; x0 = nothing)
  solver.randomized || set_rows!(k, solver, solver.state.usedIndices)
  check(ccall((:rls_kaczmarz_init, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float32), k.handle, b.handle,
              x0 === nothing ? C_NULL : x0.handle, Float32(real(solver.state.ɛw))))
end
# iterate(solver, state): Kaczmarz.jl:264-283 with the row loop :270-273 replaced by one library call
function iterate(k::B200Kaczmarz, solver::Kaczmarz, state = solver.state)
  RegularizedLeastSquares.done(solver, state) && return nothing
  if solver.randomized
    StatsBase.sample!(Random.GLOBAL_RNG, solver.rowIndexCycle, weights(solver.probabilities), state.usedIndices, replace = false)
    set_rows!(k, solver, state.usedIndices)
  end
  check(ccall((:rls_kaczmarz_sweep, LIB), Int32, (Ptr{Cvoid},), k.handle))
  xh = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:rls_kaczmarz_vec, LIB), Int32, (Ptr{Cvoid}, Cstring, Ref{Ptr{Cvoid}}), k.handle, "x", xh))
  x = B200Vector{eltype(k.A)}(xh[], size(k.A, 2), false)            # borrowed handle
  for r in solver.reg
    prox!(r, x)                                                      # rls_prox_* (Kaczmarz.jl:275-277)
  end
  state.iteration += 1
  return x, state
end
kaczmarz_check(k::B200Kaczmarz) = check(ccall((:rls_kaczmarz_check, LIB), Int32, (Ptr{Cvoid},), k.handle))

end # module
