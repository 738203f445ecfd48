"""Host-side mirror of the reference's solver API on top of the C ABI.

  createLinearSolver(S, A; kwargs...)        src/RegularizedLeastSquares.jl:288-294
  solve!(solver, b; x0, callbacks)      -> solve_(solver, b, ...)     :103-117
  init!(solver, b; x0)                  -> init_(solver, b, ...)      :190
  iterate(solver)                       -> iterate(solver)            :191
  solversolution / solverconvergence / solverstate                    :157-183
  FISTA, POGM, OptISTA, CGNR, ADMM constructors with the reference's keyword names
  (FISTA.jl:57-67, POGM.jl:75-86, OptISTA.jl:62-71, CGNR.jl:48-53, ADMM.jl:80-94).

What differs from the reference, deliberately:
  * `AHA` defaults to the LAZY normal operator A'(A x) on the device instead of
    materialising A'*A (FISTA.jl:58): for the named shapes (n = 65536) the Gram matrix is
    4x larger than A.  Pass normal="gram" for the reference's default form.
  * the default `rho = 0.95/power_iterations(AHA)` uses a Philox start vector instead of
    `randn` (Utils.jl:265) so that it is reproducible; pass `rho` for parity runs.
Everything that touches data runs in librls_b200.so; this module only resolves keyword
arguments, normalisation factors and callbacks — the part that stays in Julia.
"""
from __future__ import annotations

import ctypes as C
import inspect
import warnings

import numpy as np

from . import _capi as capi
from .arrays import B200Context, B200Group, B200GroupMatrix, B200Matrix, B200NormalOp, B200Vector, dtype_code, _FORMS
from .regularization import (AbstractProjectionRegularization, GradientOp, L1Regularization, L2Regularization,
                             MeasurementBasedNormalization, NoNormalization, NormalizedRegularization,
                             PositiveRegularization, RealRegularization, SystemMatrixBasedNormalization, findsink,
                             findsinks, lam, normalize_reg, reg_desc, sink)

_KIND = {"FISTA": capi.RLS_FISTA, "POGM": capi.RLS_POGM, "OptISTA": capi.RLS_OPTISTA, "CGNR": capi.RLS_CGNR,
         "ADMM": capi.RLS_ADMM}
_VARY = {"none": capi.RLS_VARY_RHO_NONE, "balance": capi.RLS_VARY_RHO_BALANCE, "PnP": capi.RLS_VARY_RHO_PNP}


def _real_type(dt):
    return np.float32  # Float32 / ComplexF32 only on this path


def _as_matrix(A, ctx):
    if A is None or isinstance(A, B200Matrix):
        return A
    A = np.asarray(A)
    if A.dtype not in (np.float32, np.complex64):
        raise TypeError(f"librls_b200 accelerates Float32 / ComplexF32 systems only, got {A.dtype} (no CPU fallback)")
    return B200Matrix.from_numpy(A, ctx)


class SolverState:
    """View of the device-resident solver state (FISTAState etc.): vectors are fetched on
    access, scalars mirror the last synchronisation."""
    def __init__(self, solver):
        self._s = solver

    def __getattr__(self, name):
        s = object.__getattribute__(self, "_s")
        sc = s._scalars
        if hasattr(sc, name):
            v = getattr(sc, name)
            return v if isinstance(v, (int, float)) else list(v)
        return s._vec(name).to_numpy()


class AbstractLinearSolver:
    name = ""
    _group = False          # True: A is a B200GroupMatrix (single-process multi-device)

    # ---------------- construction helpers ----------------
    _linop = None           # A is a matrix-free LinearOperator (operators.py): b is back-projected with it first

    def _setup(self, A, AHA, normal, ctx):
        from .operators import LinearOperator
        if isinstance(A, LinearOperator):
            # matrix-free A: the solver runs on the lazy normal operator A'A through the AHA-only interface (FISTA.jl:55)
            self._linop = A
            AHA = A.normal() if AHA is None else AHA
            ctx = A.ctx
            A = None
        self._group = isinstance(A, B200GroupMatrix)
        if self._group:
            # single-process multi-device: the C library keeps one operator + solver per device behind ONE handle
            if AHA is not None:
                raise NotImplementedError("a device group builds its own (lazy) normal operator; AHA cannot be supplied")
            self.ctx, self.A, self.AHA = None, A, None
            self._normal_form = _FORMS[normal]
            self.dtype, self.n = A.dtype, A.n
            self._handle = None
            self._scalars = capi.SolverScalars()
            self._vec_cache = {}
            self.state = SolverState(self)
            return
        self.ctx = ctx if ctx is not None else (A.ctx if isinstance(A, B200Matrix) else B200Context.default())
        self.A = _as_matrix(A, self.ctx)
        if AHA is None:
            if self.A is None:
                raise ValueError("either A or AHA must be given")
            self.AHA = B200NormalOp(self.A, form=normal)
        elif isinstance(AHA, B200NormalOp):
            self.AHA = AHA
        else:
            G = AHA if isinstance(AHA, B200Matrix) else _as_matrix(AHA, self.ctx)
            self.AHA = B200NormalOp(G=G)
        self.dtype = self.AHA.dtype
        self.n = self.AHA.n
        self._handle = None
        self._scalars = capi.SolverScalars()
        self._vec_cache = {}
        self.state = SolverState(self)

    def _normalize_ctor(self, regs):
        """normalize(S, normalizeReg, reg, A, nothing) in the constructors (FISTA.jl:87 ...)."""
        nz = self.normalizeReg
        if isinstance(nz, NoNormalization):
            return regs
        if self._group:
            raise NotImplementedError("normalizeReg on a device group: normalise λ on the host or use one process per GPU")
        if isinstance(nz, MeasurementBasedNormalization):
            f = np.float32(1)
        elif isinstance(nz, SystemMatrixBasedNormalization):
            if self.A is None:
                raise ValueError("SystemMatrixBasedNormalization requires supplying A to the constructor of the solver")
            fro2 = self.A.frob2()           # Σ_m ‖A[m,:]‖² (NormalizedRegularization.jl:47-58)
            if self.ctx.nranks > 1:     # row shards: Σ over all rows of the global matrix
                fro2 = self.ctx.allreduce_f64(fro2)[0]
            e = np.float32(np.sqrt(fro2))
            f = np.float32(e * e / np.float32(self.A.n))
        else:
            raise TypeError(nz)
        return [normalize_reg(r, f) for r in regs]

    def _default_rho(self):
        if self._group:
            raise ValueError("a solver on a device group needs an explicit rho (the default runs power iterations on one context)")
        b0 = B200Vector(self.ctx, self.dtype, self.n).fill_philox(seed=0x5EED, stream=7, dist=capi.RLS_DIST_IH4)
        return np.float32(0.95 / self.AHA.power_iterations(b0))

    def _create(self, desc):
        h = C.c_void_p()
        if self._group:
            import weakref
            capi.call("rls_group_solver_create", self.A.handle, self._normal_form, C.byref(desc), C.byref(h))
            self._handle, self._desc = h, desc
            self._fin = weakref.finalize(self, capi.load().rls_group_solver_destroy, h)
            return
        capi.call("rls_solver_create", self.A.handle if self.A is not None else None, self.AHA.handle, C.byref(desc), C.byref(h))
        import weakref
        self._handle = h
        self._desc = desc
        self._fin = weakref.finalize(self, capi.load().rls_solver_destroy, h)

    # ---------------- device state access ----------------
    def _vec(self, name):
        h = C.c_void_p()
        capi.call("rls_solver_vec", self._handle, name.encode(), C.byref(h))
        ln, dt = C.c_int64(), C.c_int32()
        capi.call("rls_vec_len", h, C.byref(ln), C.byref(dt))
        return B200Vector(self.ctx, self.dtype, ln.value, _handle=h, _owned=False)

    @property
    def x(self):
        return self._vec("x").to_numpy()

    @property
    def iteration(self):
        return self._scalars.iteration

    # ---------------- λ normalisation inside init! ----------------
    def _renormalize(self, b_host, b_dev):
        """solver.reg = normalize(solver, normalizeReg, reg, A, ·) at the end of init!."""
        if not isinstance(self.normalizeReg, MeasurementBasedNormalization):
            return
        src = self._norm_source(b_host, b_dev)          # FISTA family: x₀ = A'b ; CGNR/ADMM: b
        if isinstance(src, B200Vector):
            asum, length = src.asum(), src.length
        else:
            asum, length = float(np.sum(np.abs(src), dtype=np.float32)), src.size
        if self.ctx.nranks > 1 and self._norm_source_is_sharded:
            # b is row-sharded: ‖b‖₁ and length(b) are sums over the ranks (x₀ = A'b is replicated and is not)
            asum, length = self.ctx.allreduce_f64(asum, length)
        f = np.float32(np.float32(asum) / np.float32(length))
        self._apply_factor(f)

    def _apply_factor(self, f):
        regs = self.reg if isinstance(self.reg, list) else [self.reg]
        regs = [normalize_reg(r, f) for r in regs]
        if isinstance(self.reg, list):
            self.reg = regs
        else:
            self.reg = regs[0]
        for i, r in enumerate(regs):
            d = self._reg_desc(i, r)
            capi.call("rls_solver_set_reg", self._handle, i, C.byref(d))

    def _reg_desc(self, i, r):
        return reg_desc(r)

    def _norm_source(self, b_host, b_dev):
        return self._vec("x0")

    _norm_source_is_sharded = False

    # ---------------- the iterator protocol ----------------
    def _b_to_device(self, b):
        if self._linop is not None and not getattr(self, "_in_linop_solve", False):
            bd = b if isinstance(b, B200Vector) else B200Vector.from_numpy(np.ascontiguousarray(b, dtype=self.dtype), self.ctx)
            return self._linop.tmul(bd)                     # init!(solver, b) on a matrix-free A: x0 = A'b
        if isinstance(b, B200Vector):
            return b
        b = np.ascontiguousarray(b, dtype=self.dtype)
        return B200Vector.from_numpy(b, self.ctx)

    def init_(self, b, x0=0):
        """init!(solver, b; x0)"""
        bd = self._b_to_device(b)
        x0d = None
        if not (np.isscalar(x0) and x0 == 0):
            x0d = x0 if isinstance(x0, B200Vector) else B200Vector.from_numpy(np.asarray(x0, dtype=self.dtype), self.ctx)
        if isinstance(self.normalizeReg, MeasurementBasedNormalization) and not self._norm_after_init:
            self._renormalize(b, bd)
        capi.call("rls_solver_init", self._handle, bd.handle, x0d.handle if x0d is not None else None)
        if isinstance(self.normalizeReg, MeasurementBasedNormalization) and self._norm_after_init:
            self._renormalize(b, bd)
        capi.call("rls_solver_scalars_get", self._handle, C.byref(self._scalars))
        self._b_keepalive = (bd, x0d)
        return self

    _norm_after_init = True

    def iterate(self):
        """iterate(solver): returns False when done (Julia `nothing`), True otherwise."""
        adv = C.c_int32()
        capi.call("rls_solver_iterate", self._handle, C.byref(adv), C.byref(self._scalars))
        return bool(adv.value)

    def solve_(self, b, x0=0, callbacks=None, scheduler=None):
        """solve!(solver, b; x0, callbacks): RegularizedLeastSquares.jl:103-117; a matrix b runs the
        multi-right-hand-side path of MultiThreading.jl:30-80."""
        host_in = not isinstance(b, B200Vector)
        if self._linop is not None and not getattr(self, "_in_linop_solve", False):
            if isinstance(self.normalizeReg, MeasurementBasedNormalization):
                raise NotImplementedError("MeasurementBasedNormalization with a matrix-free operator")
            if host_in and np.ndim(b) != 1:
                raise NotImplementedError("a matrix-free operator takes one right-hand side at a time")
            bd = b if not host_in else B200Vector.from_numpy(np.ascontiguousarray(b, dtype=self.dtype), self.ctx)
            self._in_linop_solve = True
            try:
                xv = self.solve_(self._linop.tmul(bd), x0=x0, callbacks=callbacks)     # solve!(solver, A'b)
            finally:
                self._in_linop_solve = False
            return xv.to_numpy() if host_in else xv
        if self._group:
            if not host_in or np.ndim(b) != 1 or callbacks or not (np.isscalar(x0) and x0 == 0):
                raise NotImplementedError("a solver on a device group takes one host vector b (no callbacks, x0 = 0)")
            bh = np.ascontiguousarray(b, dtype=self.dtype).ravel()
            xh = np.empty(self.n, self.dtype)
            it = C.c_int32()
            capi.call("rls_group_solver_solve_host", self._handle, bh.ctypes.data_as(C.c_void_p), bh.size,
                      xh.ctypes.data_as(C.c_void_p), xh.size, C.byref(it), C.byref(self._scalars))
            return xh
        if host_in and np.ndim(b) == 2:
            return self._solve_batch(np.asarray(b))
        cbs = [] if callbacks is None else (list(callbacks) if isinstance(callbacks, (list, tuple)) else [callbacks])
        needs_host_step = isinstance(self.normalizeReg, MeasurementBasedNormalization) and self._norm_after_init
        if not cbs and host_in and not needs_host_step and np.isscalar(x0) and x0 == 0:
            # callback-free fast path: one C call, host buffers in and out
            if isinstance(self.normalizeReg, MeasurementBasedNormalization):
                self._renormalize(np.asarray(b), None)
            bh = np.ascontiguousarray(b, dtype=self.dtype).ravel()
            xh = np.empty(self.n, self.dtype)
            it = C.c_int32()
            capi.call("rls_solver_solve_host", self._handle, bh.ctypes.data_as(C.c_void_p), bh.size,
                      xh.ctypes.data_as(C.c_void_p), xh.size, C.byref(it), C.byref(self._scalars))
            return xh
        self.init_(b, x0=x0)
        for cb in cbs:
            cb(self, 0)
        if not cbs:
            it = C.c_int32()
            capi.call("rls_solver_run", self._handle, C.byref(it), C.byref(self._scalars))
        else:
            k = 0
            while self.iterate():
                k += 1
                for cb in cbs:
                    cb(self, k)
        xv = self._vec("x")
        return xv.to_numpy() if host_in else xv

    def _solve_batch(self, B):
        B = np.asfortranarray(B, dtype=self.dtype)
        K = B.shape[1]
        X = np.empty((self.n, K), self.dtype, order="F")
        its = (C.c_int32 * K)()
        if isinstance(self.normalizeReg, MeasurementBasedNormalization):
            raise NotImplementedError("MeasurementBasedNormalization with a matrix b: the reference lets the last "
                                      "column's factor win for all (SURVEY quirk 11); normalise per column instead")
        capi.call("rls_solver_solve_batch_host", self._handle, B.ctypes.data_as(C.c_void_p), B.shape[0], K,
                  X.ctypes.data_as(C.c_void_p), self.n, its)
        self.batch_iterations = list(its)
        capi.call("rls_solver_scalars_get", self._handle, C.byref(self._scalars))
        return X

    def convergence(self):
        """solverconvergence(solver)"""
        return {"residual": self._scalars.res_norm}


class _ProxGradSolver(AbstractLinearSolver):
    """FISTA / POGM / OptISTA share constructor and init! structure."""
    def __init__(self, A, *, AHA=None, reg=None, normalizeReg=None, iterations=50, verbose=False, rho=None, theta=1,
                 relTol=None, restart="none", sigma_fac=1, normal="auto", ctx=None):
        self._setup(A, AHA, normal, ctx)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        if reg is None:
            reg = L1Regularization(np.float32(0))
        regs = list(reg) if isinstance(reg, (list, tuple)) else [reg]
        idx = findsinks(AbstractProjectionRegularization, regs)
        self.proj = [regs[i] for i in idx]
        regs = [r for i, r in enumerate(regs) if i not in idx]
        if len(regs) != 1:
            raise ValueError(f"{self.name} does not allow for more additional regularization terms, found {len(regs)}")
        self.reg = self._normalize_ctor(regs)[0]
        self.iterations = int(iterations)
        self.verbose = verbose
        self.restart = restart
        if restart not in ("none", "gradient"):
            raise ValueError("restart must be :none or :gradient")
        self.rho = np.float32(self._default_rho() if rho is None else rho)
        self.relTol = np.float32(np.finfo(np.float32).eps if relTol is None else relTol)
        d = capi.SolverDesc()
        d.kind = _KIND[self.name]
        d.iterations = self.iterations
        d.restart = 1 if (restart == "gradient" and self.name != "OptISTA") else 0
        d.proj_mask = 0
        if self.name != "OptISTA":          # OptISTA stores but never applies them (quirk 10)
            for p in self.proj:
                d.proj_mask |= sink(p).mask
        d.rho = self.rho
        d.theta = float(theta)
        d.sigma_fac = float(sigma_fac)
        d.rel_tol = self.relTol
        d.n_reg = 1
        d.reg[0] = reg_desc(self.reg)
        self._create(d)


class FISTA(_ProxGradSolver):
    """FISTA(A; AHA, reg, normalizeReg, iterations, verbose, rho, theta, relTol, restart)  FISTA.jl:57-92"""
    name = "FISTA"


class POGM(_ProxGradSolver):
    """POGM(A; ..., sigma_fac, restart)  POGM.jl:75-114"""
    name = "POGM"


class OptISTA(_ProxGradSolver):
    """OptISTA(A; AHA, reg, normalizeReg, iterations, verbose, rho, theta, relTol)  OptISTA.jl:62-105"""
    name = "OptISTA"

    def __init__(self, A, **kw):
        kw.pop("restart", None)
        kw.pop("sigma_fac", None)
        super().__init__(A, **kw)


class CGNR(AbstractLinearSolver):
    """CGNR(A; AHA, reg, normalizeReg, iterations, relTol)  CGNR.jl:48-89"""
    name = "CGNR"
    _norm_after_init = False   # init! normalises with b itself (CGNR.jl:129)

    def __init__(self, A, *, AHA=None, reg=None, normalizeReg=None, iterations=10, relTol=None, normal="auto", ctx=None):
        self._setup(A, AHA, normal, ctx)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        regs = [] if reg is None else (list(reg) if isinstance(reg, (list, tuple)) else [reg])
        regs = self._normalize_ctor(regs)
        i2 = findsink(L2Regularization, regs)
        if i2 is None:
            self.L2 = L2Regularization(np.float32(0))
        else:
            self.L2 = regs.pop(i2)
        idx = findsinks(RealRegularization, regs) + findsinks(PositiveRegularization, regs)
        self.constr = [regs[i] for i in idx]
        regs = [r for i, r in enumerate(regs) if i not in idx]
        if regs:
            raise ValueError(f"CGNR does not allow for more additional regularization terms, found {len(regs)}")
        self.reg = self.L2
        self.iterations = int(iterations)
        self.relTol = np.float32(np.finfo(np.float32).eps if relTol is None else relTol)
        d = capi.SolverDesc()
        d.kind = capi.RLS_CGNR
        d.iterations = self.iterations
        d.rel_tol = self.relTol
        for p in self.constr:
            d.proj_mask |= sink(p).mask
        d.n_reg = 1
        d.reg[0] = reg_desc(self.L2)
        self._create(d)

    def _norm_source(self, b_host, b_dev):
        return b_dev if b_dev is not None else np.asarray(b_host)

    _norm_source_is_sharded = True

    def _apply_factor(self, f):
        super()._apply_factor(f)
        self.L2 = self.reg


class ADMM(AbstractLinearSolver):
    """ADMM(A; AHA, precon, reg, regTrafo, normalizeReg, rho, vary_rho, iterations, iterationsCG,
    absTol, relTol, tolInner, verbose)  ADMM.jl:80-164 (precon = Identity only)"""
    name = "ADMM"
    _norm_after_init = False   # init! normalises with b (ADMM.jl:219)

    def __init__(self, A, *, AHA=None, precon=None, reg=None, regTrafo=None, normalizeReg=None, rho=1e-1,
                 vary_rho="none", iterations=10, iterationsCG=10, absTol=None, relTol=None, tolInner=1e-5,
                 verbose=False, normal="auto", ctx=None):
        if precon is not None:
            raise NotImplementedError("ADMM: only the Identity() preconditioner is on the accelerated path")
        self._setup(A, AHA, normal, ctx)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        if reg is None:
            reg = L1Regularization(np.float32(0))
        regs = list(reg) if isinstance(reg, (list, tuple)) else [reg]
        idx = findsinks(AbstractProjectionRegularization, regs)
        self.proj = [regs[i] for i in idx]
        regs = [r for i, r in enumerate(regs) if i not in idx]
        if regTrafo is None:
            trafo = [None] * len(regs)
        else:
            trafo = list(regTrafo) if isinstance(regTrafo, (list, tuple)) else [regTrafo]
        assert len(regs) == len(trafo), "reg and regTrafo must have the same length"
        if not 1 <= len(regs) <= 4:
            raise ValueError("ADMM on this path takes 1..4 regularization terms")
        self.regTrafo = trafo
        self.rho = [np.float32(rho)] * len(regs) if np.isscalar(rho) else [np.float32(r) for r in rho]
        self.reg = self._normalize_ctor(regs)
        eps = np.finfo(np.float32).eps
        self.iterations, self.iterationsCG = int(iterations), int(iterationsCG)
        d = capi.SolverDesc()
        d.kind = capi.RLS_ADMM
        d.iterations = self.iterations
        d.iterations_cg = self.iterationsCG
        d.abs_tol = np.float32(eps if absTol is None else absTol)
        d.rel_tol = np.float32(eps if relTol is None else relTol)
        d.tol_inner = np.float32(tolInner)
        d.vary_rho = _VARY[vary_rho]
        for p in self.proj:
            d.proj_mask |= sink(p).mask
        d.n_reg = len(regs)
        for i, r in enumerate(self.reg):
            d.reg[i] = self._reg_desc(i, r)
        self._create(d)

    def _reg_desc(self, i, r):
        return reg_desc(r, rho=self.rho[i], trafo=self.regTrafo[i])

    def _norm_source(self, b_host, b_dev):
        return b_dev if b_dev is not None else np.asarray(b_host)

    _norm_source_is_sharded = True

    def convergence(self):
        k = len(self.reg)
        return {"primal": list(self._scalars.admm_rk)[:k], "dual": list(self._scalars.admm_sk)[:k]}


class SplitBregman(ADMM):
    """SplitBregman(A; AHA, precon, reg, regTrafo, normalizeReg, rho, iterations, iterationsInner, iterationsCG,
    absTol, relTol, tolInner, verbose)  SplitBregman.jl:80-146 (precon = Identity only; regTrafo = identity or
    GradientOp, as for ADMM).  `iterations` counts the outer (Bregman) iterations, `iterationsInner` the ADMM-like inner
    ones; `state.iteration` restarts at every Bregman update, `iter_cnt` is `outer_iteration`."""
    name = "SplitBregman"
    _norm_after_init = False

    def __init__(self, A, *, AHA=None, precon=None, reg=None, regTrafo=None, normalizeReg=None, rho=1e-1,
                 iterations=10, iterationsInner=10, iterationsCG=10, absTol=None, relTol=None, tolInner=1e-5,
                 verbose=False, normal="auto", ctx=None):
        if precon is not None:
            raise NotImplementedError("SplitBregman: only the Identity() preconditioner is on the accelerated path")
        self._setup(A, AHA, normal, ctx)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        if reg is None:
            reg = L1Regularization(np.float32(0))
        regs = list(reg) if isinstance(reg, (list, tuple)) else [reg]
        idx = findsinks(AbstractProjectionRegularization, regs)
        self.proj = [regs[i] for i in idx]
        regs = [r for i, r in enumerate(regs) if i not in idx]
        if regTrafo is None:
            trafo = [None] * len(regs)
        else:
            trafo = list(regTrafo) if isinstance(regTrafo, (list, tuple)) else [regTrafo]
        assert len(regs) == len(trafo), "reg and regTrafo must have the same length"      # SplitBregman.jl:108
        if not 1 <= len(regs) <= 4:
            raise ValueError("SplitBregman on this path takes 1..4 regularization terms")
        self.regTrafo = trafo
        self.rho = [np.float32(rho)] * len(regs) if np.isscalar(rho) else [np.float32(r) for r in rho]
        self.reg = self._normalize_ctor(regs)
        eps = np.finfo(np.float32).eps
        self.iterations, self.iterationsInner, self.iterationsCG = int(iterations), int(iterationsInner), int(iterationsCG)
        d = capi.SolverDesc()
        d.kind = capi.RLS_SPLITBREGMAN
        d.iterations = self.iterations
        d.iterations_inner = self.iterationsInner
        d.iterations_cg = self.iterationsCG
        d.abs_tol = np.float32(eps if absTol is None else absTol)
        d.rel_tol = np.float32(eps if relTol is None else relTol)
        d.tol_inner = np.float32(tolInner)
        for p in self.proj:
            d.proj_mask |= sink(p).mask
        d.n_reg = len(regs)
        for i, r in enumerate(self.reg):
            d.reg[i] = self._reg_desc(i, r)
        self._create(d)

    @property
    def iter_cnt(self):
        return self._scalars.outer_iteration


class Kaczmarz(AbstractLinearSolver):
    """Kaczmarz(A; reg, normalizeReg, randomized, subMatrixFraction, shuffleRows, seed, iterations)  Kaczmarz.jl:73-159

    The constructor / init! bookkeeping of the reference stays here (L2 term and Tikhonov matrix :83-93, denom and
    rowindex :365-392, probabilities :326-334, row order :195-203, prox! after each sweep :275-277); the row loop
    (:270-273) runs in librls_b200 (csrc/rls_kaczmarz.cu) on a row-major device matrix.  Not accelerated:
    greedy_randomized (the reference excludes it on GPU arrays too, test/testKaczmarz.jl:114) and a communicator
    (the row loop is sequential; it does not shard).

    Row order: Julia's RNG stream cannot be reproduced; shuffleRows / randomized draw from
    numpy.random.default_rng(seed) — a permutation at init!, a weighted sample without replacement per iteration.
    Every new order rebuilds the block Gram matrices, so `randomized=True` pays that build per iteration."""
    name = "Kaczmarz"

    def __init__(self, A, *, AHA=None, reg=None, normalizeReg=None, randomized=False, subMatrixFraction=0.15,
                 shuffleRows=False, seed=1234, iterations=10, greedy_randomized=False, block_rows=0, ctx=None):
        if greedy_randomized:
            raise NotImplementedError("greedy randomised Kaczmarz is not on the accelerated path (no CPU fallback)")
        if A is None:
            raise ValueError("Kaczmarz needs the system matrix A")
        self.ctx = ctx if ctx is not None else (A.ctx if isinstance(A, B200Matrix) else B200Context.default())
        if self.ctx.nranks > 1:
            raise NotImplementedError("Kaczmarz does not shard over ranks (sequential row loop)")
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        dtype = A.dtype if isinstance(A, B200Matrix) else np.asarray(A).dtype
        if dtype not in (np.float32, np.complex64):
            raise TypeError(f"librls_b200 accelerates Float32 / ComplexF32 systems only, got {dtype} (no CPU fallback)")
        self.dtype = np.dtype(dtype)
        regs = [L2Regularization(np.float32(0))] if reg is None else (list(reg) if isinstance(reg, (list, tuple)) else [reg])
        regs = list(regs)
        # --- L2 term, Tikhonov matrix (:83-93, :377-392): the scaled matrix A·diag(1/sqrt(λ)) is what goes to the device
        i2 = findsink(L2Regularization, regs)
        self._tikhonov = None
        if i2 is not None and np.ndim(lam(regs[i2])) > 0:
            if not isinstance(self.normalizeReg, (NoNormalization, SystemMatrixBasedNormalization)):
                raise ValueError("Tikhonov matrix for Kaczmarz is only valid with no or system matrix based normalization")
            if not isinstance(self.normalizeReg, NoNormalization):
                raise NotImplementedError("Tikhonov matrix together with SystemMatrixBasedNormalization")
            if isinstance(A, B200Matrix):
                raise NotImplementedError("a Tikhonov matrix needs the host array A (it rescales the columns before upload)")
        if isinstance(A, B200Matrix):
            if A.layout != "row":
                raise ValueError("Kaczmarz needs a row-major device matrix: B200Matrix.from_numpy(A, layout='row')")
            self.A = A
        else:
            self._A_host = np.asarray(A)
            self.A = None                    # uploaded below, once a Tikhonov scaling is known
        self.m, self.n = (self.A.m, self.A.n) if self.A is not None else self._A_host.shape
        self._upload_and_norms(regs, i2, block_rows)
        regs = self._normalize_ctor(regs)                                     # :82 (needs self.A for SystemMatrixBased)
        if i2 is None:
            self.L2 = L2Regularization(np.float32(0))
        else:
            self.L2 = regs.pop(i2)
        idx = findsinks(AbstractProjectionRegularization, regs)               # :95-97
        other = [regs[i] for i in idx]
        regs = [r for i, r in enumerate(regs) if i not in idx]
        if len(regs) == 1:
            other.append(regs[0])
        elif len(regs) > 1:
            raise ValueError(f"Kaczmarz does not allow for more than one additional regularization term, found {len(regs)}")
        self.reg = other
        self._initkaczmarz(self._lam_scalar())
        self.randomized = bool(randomized)
        self.shuffleRows = bool(shuffleRows)
        self.seed = int(seed)
        self.subMatrixSize = int(np.round(subMatrixFraction * self.m))        # :121
        self.rowIndexCycle = np.arange(len(self.rowindex))
        self.iterations = int(iterations)
        self._iteration = 0
        self._order = None
        self._scalars = capi.SolverScalars()
        self.state = SolverState(self)

    # ---------------- setup ----------------
    def _upload_and_norms(self, regs, i2, block_rows):
        if self.A is None:
            Ah = self._A_host
            if i2 is not None and np.ndim(lam(regs[i2])) > 0:                 # initikhonov :392
                lv = np.asarray(lam(regs[i2])).astype(np.float32)
                if lv.shape != (self.n,):
                    raise ValueError("the Tikhonov matrix needs one λ per column of A")
                self._tikhonov = lv
                Ah = (Ah * (np.float32(1) / np.sqrt(lv))[None, :]).astype(self.dtype)
            self.A = B200Matrix.from_numpy(np.asarray(Ah, dtype=self.dtype), self.ctx, layout="row")
            del self._A_host
        h = C.c_void_p()
        capi.call("rls_kaczmarz_create", self.A.handle, int(block_rows), C.byref(h))
        import weakref
        self._handle = h
        self._fin = weakref.finalize(self, capi.load().rls_kaczmarz_destroy, h)
        s2 = np.empty(self.m, np.float32)
        capi.call("rls_kaczmarz_rownorm2", h, s2.ctypes.data_as(C.c_void_p), s2.size)
        self._s2 = s2                                                         # rownorm²(A, i), Utils.jl:16-23
        br = C.c_int32()
        capi.call("rls_kaczmarz_block_rows", h, C.byref(br))
        self.block_rows = br.value

    def _lam_scalar(self):
        """λ seen by initkaczmarz: one(T) for a Tikhonov matrix (:390), λ(L2) otherwise"""
        return np.float32(1) if self._tikhonov is not None else lam(self.L2)

    def _initkaczmarz(self, lam_):
        """:365-376: rows with s² > 0, denom = T(1.0 / (s² + λ))"""
        keep = np.nonzero(self._s2 > 0)[0]
        tot = self._s2[keep] + lam_                                           # Float32 (+ Float64 λ promotes)
        self.denom = (1.0 / np.asarray(tot, dtype=np.float64)).astype(np.float32)
        self.rowindex = keep.astype(np.int64)
        self._lam_used = lam_

    def _row_probabilities(self):
        """:326-334, converted to T at :125"""
        tot = np.float32(np.sum(self._s2, dtype=np.float32))
        return (self._s2[self.rowindex].astype(np.float64) / np.float64(tot)).astype(np.float32)

    def _set_order(self, used):
        used = np.asarray(used, dtype=np.int64)
        if self._order is not None and self._order[0] is self.denom and np.array_equal(self._order[1], used):
            return
        rows = np.ascontiguousarray(self.rowindex[used], dtype=np.int64)
        den = np.ascontiguousarray(self.denom[used], dtype=np.float32)
        capi.call("rls_kaczmarz_set_rows", self._handle, rows.ctypes.data_as(C.c_void_p), den.ctypes.data_as(C.c_void_p),
                  rows.size)
        self._order = (self.denom, used.copy())

    # ---------------- state access ----------------
    def _vec(self, name):
        h = C.c_void_p()
        capi.call("rls_kaczmarz_vec", self._handle, name.encode(), C.byref(h))
        ln, dt = C.c_int64(), C.c_int32()
        capi.call("rls_vec_len", h, C.byref(ln), C.byref(dt))
        return B200Vector(self.ctx, self.dtype, ln.value, _handle=h, _owned=False)

    def describe(self):
        """kernel plan of the sweep ("persistent: ..." = one cooperative kernel per iteration, "chained: ..." otherwise)"""
        buf = C.create_string_buffer(256)
        capi.call("rls_kaczmarz_describe", self._handle, buf, 256)
        return buf.value.decode()

    @property
    def x(self):
        """solversolution(solver): the Tikhonov-matrix form returns x ./ sqrt.(λ) (:253-256)"""
        capi.call("rls_kaczmarz_check", self._handle)
        x = self._vec("x").to_numpy()
        if self._tikhonov is not None:
            x = (x * (np.float32(1) / np.sqrt(self._tikhonov))).astype(self.dtype)
        return x

    @property
    def iteration(self):
        return self._iteration

    # ---------------- init! / iterate ----------------
    def init_(self, b, x0=0):
        """init!(solver, state, b; x0)  :178-216"""
        bd = b if isinstance(b, B200Vector) else B200Vector.from_numpy(np.ascontiguousarray(b, dtype=self.dtype), self.ctx)
        lam_prev = self._lam_scalar()
        if isinstance(self.normalizeReg, MeasurementBasedNormalization):      # :179-181 (SystemMatrixBased: unchanged)
            f = np.float32(np.float32(bd.asum()) / np.float32(bd.length))
            self.L2 = normalize_reg(self.L2, f)
            self.reg = [normalize_reg(r, f) for r in self.reg]
        lam_ = self._lam_scalar()
        if lam_ != lam_prev:                                                  # :186-193
            self._initkaczmarz(lam_)
            self.rowIndexCycle = np.arange(len(self.rowindex))
        self._rng = np.random.default_rng(self.seed)                          # :195-197
        if self.randomized:
            self.probabilities = self._row_probabilities()
        elif self.shuffleRows:
            self.rowIndexCycle = self._rng.permutation(self.rowIndexCycle)    # :201
        if not self.randomized:
            self._set_order(self.rowIndexCycle)
        x0d = None
        if not (np.isscalar(x0) and x0 == 0):
            x0d = x0 if isinstance(x0, B200Vector) else B200Vector.from_numpy(np.asarray(x0, dtype=self.dtype), self.ctx)
        eps_w = np.float32(1) if self._tikhonov is not None else np.float32(np.sqrt(lam_))   # :210-214
        capi.call("rls_kaczmarz_init", self._handle, bd.handle, x0d.handle if x0d is not None else None, eps_w)
        self._iteration = 0
        self._b_keepalive = (bd, x0d)
        return self

    def iterate(self):
        """iterate(solver, state)  :264-283"""
        if self._iteration >= self.iterations:                                # done() :315
            return False
        if self.randomized:                                                   # sample! :267-269
            p = self.probabilities.astype(np.float64)
            used = self._rng.choice(len(self.rowIndexCycle), size=self.subMatrixSize, replace=False, p=p / p.sum())
            self._set_order(used)
        capi.call("rls_kaczmarz_sweep", self._handle)
        if self.reg:
            from .prox import prox_
            xv = self._vec("x")
            for r in self.reg:                                                # :275-277
                prox_(r, xv)
        self._iteration += 1
        return True

    def solve_(self, b, x0=0, callbacks=None, scheduler=None):
        host_in = not isinstance(b, B200Vector)
        if self._linop is not None and not getattr(self, "_in_linop_solve", False):
            if isinstance(self.normalizeReg, MeasurementBasedNormalization):
                raise NotImplementedError("MeasurementBasedNormalization with a matrix-free operator")
            if host_in and np.ndim(b) != 1:
                raise NotImplementedError("a matrix-free operator takes one right-hand side at a time")
            bd = b if not host_in else B200Vector.from_numpy(np.ascontiguousarray(b, dtype=self.dtype), self.ctx)
            self._in_linop_solve = True
            try:
                xv = self.solve_(self._linop.tmul(bd), x0=x0, callbacks=callbacks)     # solve!(solver, A'b)
            finally:
                self._in_linop_solve = False
            return xv.to_numpy() if host_in else xv
        if self._group:
            if not host_in or np.ndim(b) != 1 or callbacks or not (np.isscalar(x0) and x0 == 0):
                raise NotImplementedError("a solver on a device group takes one host vector b (no callbacks, x0 = 0)")
            bh = np.ascontiguousarray(b, dtype=self.dtype).ravel()
            xh = np.empty(self.n, self.dtype)
            it = C.c_int32()
            capi.call("rls_group_solver_solve_host", self._handle, bh.ctypes.data_as(C.c_void_p), bh.size,
                      xh.ctypes.data_as(C.c_void_p), xh.size, C.byref(it), C.byref(self._scalars))
            return xh
        if host_in and np.ndim(b) == 2:
            return np.stack([self.solve_(np.asarray(b)[:, k], x0=x0, callbacks=callbacks) for k in range(np.shape(b)[1])], axis=1)
        cbs = [] if callbacks is None else (list(callbacks) if isinstance(callbacks, (list, tuple)) else [callbacks])
        self.init_(b, x0=x0)
        for cb in cbs:
            cb(self, 0)
        k = 0
        while self.iterate():
            k += 1
            for cb in cbs:
                cb(self, k)
        if host_in or self._tikhonov is not None:
            return self.x
        capi.call("rls_kaczmarz_check", self._handle)
        return self._vec("x")

    def convergence(self):
        """solverconvergence :258: ‖A x − u‖"""
        r = self.A.mul(self._vec("x"))
        return {"residual": float(np.linalg.norm(r.to_numpy() - self._vec("u").to_numpy()))}


def linearSolverList():
    """the solvers on the accelerated path (RegularizedLeastSquares.jl:213-220 lists all of upstream's)"""
    return [CGNR, Kaczmarz, FISTA, OptISTA, POGM, ADMM, SplitBregman]


def createLinearSolver(solver, A=None, *, AHA=None, kwargWarning=True, **kwargs):
    """createLinearSolver(S, A; kwargs...): unknown keywords are dropped with a warning
    (filterKwargs, RegularizedLeastSquares.jl:267-278)."""
    names = set()
    for klass in solver.__mro__:
        if klass is object:
            continue
        names |= set(inspect.signature(klass.__init__).parameters)
    kept = {k: v for k, v in kwargs.items() if k in names}
    dropped = [k for k in kwargs if k not in names]
    if dropped and kwargWarning:
        warnings.warn("The following arguments were passed but filtered out: " + ", ".join(dropped) +
                      ". Please watch closely if this introduces unexpexted behaviour in your code.")
    return solver(A, AHA=AHA, **kept)


def solve_(solver, b, **kw):
    """solve!(solver, b; kwargs...)"""
    return solver.solve_(b, **kw)


def init_(solver, b, **kw):
    """init!(solver, b; kwargs...)"""
    return solver.init_(b, **kw)


def iterate(solver):
    return solver.iterate()


def solversolution(solver):
    return solver.x


def solverconvergence(solver):
    return solver.convergence()


def solverstate(solver):
    return solver.state
