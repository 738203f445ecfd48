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
from .arrays import B200Context, B200Matrix, B200NormalOp, B200Vector, dtype_code
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

    # ---------------- construction helpers ----------------
    def _setup(self, A, AHA, normal, ctx):
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
        if isinstance(nz, MeasurementBasedNormalization):
            f = np.float32(1)
        elif isinstance(nz, SystemMatrixBasedNormalization):
            if self.A is None:
                raise ValueError("SystemMatrixBasedNormalization requires supplying A to the constructor of the solver")
            fro2 = self.A.frob2()           # Σ_m ‖A[m,:]‖² (NormalizedRegularization.jl:47-58)
            if self.ctx.nranks > 1:
                import torch
                import torch.distributed as dist
                t = torch.tensor([fro2], dtype=torch.float64)
                dist.all_reduce(t)
                fro2 = float(t[0])
            e = np.float32(np.sqrt(fro2))
            f = np.float32(e * e / np.float32(self.A.n))
        else:
            raise TypeError(nz)
        return [normalize_reg(r, f) for r in regs]

    def _default_rho(self):
        b0 = B200Vector(self.ctx, self.dtype, self.n).fill_philox(seed=0x5EED, stream=7, dist=capi.RLS_DIST_IH4)
        return np.float32(0.95 / self.AHA.power_iterations(b0))

    def _create(self, desc):
        h = C.c_void_p()
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
        f = np.float32(np.float32(src.asum()) / np.float32(src.length)) if isinstance(src, B200Vector) \
            else np.float32(np.sum(np.abs(src), dtype=np.float32) / np.float32(src.size))
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

    # ---------------- the iterator protocol ----------------
    def _b_to_device(self, b):
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


def linearSolverList():
    """the solvers on the accelerated path (RegularizedLeastSquares.jl:213-220 lists all of upstream's)"""
    return [CGNR, FISTA, OptISTA, POGM, ADMM, SplitBregman]


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
