"""Host-side mirror of src/Regularization/*.jl: the decorator plumbing that turns a user's
regularization term + normalisation scheme into the scalar λ handed to the library
(SURVEY 8a row a21 — this part stays on the host in the Julia shim as well).

Names, argument meaning and error behaviour follow the reference:
  L1Regularization(λ), L2Regularization(λ), L21Regularization(λ; slices),
  TVRegularization(λ; shape, dims, iterationsTV), NuclearRegularization(λ; svtShape),
  LLRRegularization(λ; shape, blockSize, randshift, fullyOverlapping), PositiveRegularization(), RealRegularization(),
  NormalizedRegularization(reg, factor), NoNormalization / MeasurementBasedNormalization /
  SystemMatrixBasedNormalization, λ(reg) -> lam(reg), sink(reg), findsink(s).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi


def _julia_scalar(v):
    # a Python literal plays the role of a Julia Float64 literal; NumPy scalars keep their type
    if isinstance(v, (np.floating,)):
        return v
    return np.float64(v)


class AbstractRegularization:
    pass


class AbstractParameterizedRegularization(AbstractRegularization):
    def __init__(self, lam):
        self.lam = _julia_scalar(lam)   # reg.λ


class AbstractProjectionRegularization(AbstractRegularization):
    pass


class L1Regularization(AbstractParameterizedRegularization):
    """src/proximalMaps/ProxL1.jl:8-11"""
    kind = capi.RLS_REG_L1


class L2Regularization(AbstractParameterizedRegularization):
    """src/proximalMaps/ProxL2.jl:8-11"""
    kind = capi.RLS_REG_L2


class L21Regularization(AbstractParameterizedRegularization):
    """src/proximalMaps/ProxL21.jl:14-18"""
    kind = capi.RLS_REG_L21

    def __init__(self, lam, slices=1):
        super().__init__(lam)
        self.slices = int(slices)


class TVRegularization(AbstractParameterizedRegularization):
    """src/proximalMaps/ProxTV.jl:32-39; all TV goes through FGP (ProxTV.jl:82-125)"""
    kind = capi.RLS_REG_TV

    def __init__(self, lam, shape=(0,), dims=None, iterationsTV=10):
        super().__init__(lam)
        self.shape = tuple(int(s) for s in shape)
        if dims is None:
            dims = range(1, len(self.shape) + 1)
        self.dims = (int(dims),) if np.isscalar(dims) else tuple(int(d) for d in dims)
        self.iterationsTV = int(iterationsTV)
        if not 1 <= len(self.shape) <= capi.RLS_MAX_TV_DIMS:
            raise ValueError(f"TVRegularization: 1..{capi.RLS_MAX_TV_DIMS} dimensions supported")


class NuclearRegularization(AbstractParameterizedRegularization):
    """src/proximalMaps/ProxNuclear.jl:15-19: singular-value soft-thresholding of reshape(x, svtShape)"""
    kind = capi.RLS_REG_NUCLEAR

    def __init__(self, lam, svtShape=()):
        super().__init__(lam)
        self.svtShape = tuple(int(s) for s in svtShape)
        if len(self.svtShape) != 2:
            raise ValueError("NuclearRegularization: svtShape must be (rows, cols)")


class LLRRegularization(AbstractParameterizedRegularization):
    """src/proximalMaps/ProxLLR.jl:20-29: locally low rank — singular-value thresholding of every blockSize patch of an
    image series.  `randshift` draws a new circular shift of the patch grid at every prox! call (:55); `seed` makes that
    stream reproducible here (the reference uses the global RNG)."""
    kind = capi.RLS_REG_LLR

    def __init__(self, lam, shape=(0,), blockSize=None, randshift=True, fullyOverlapping=False, L=1, seed=0):
        super().__init__(lam)
        self.shape = tuple(int(s) for s in shape)
        self.blockSize = tuple(int(b) for b in blockSize) if blockSize is not None else (2,) * len(self.shape)
        if len(self.blockSize) != len(self.shape) or not 1 <= len(self.shape) <= capi.RLS_MAX_TV_DIMS:
            raise ValueError(f"LLRRegularization: shape and blockSize need the same 1..{capi.RLS_MAX_TV_DIMS} dimensions")
        self.randshift, self.fullyOverlapping, self.L, self.seed = bool(randshift), bool(fullyOverlapping), int(L), int(seed)
        self._calls = 0

    def next_shift(self):
        """the shift of this prox! call: rand(CartesianIndices(blockSize)) (:55), zero when randshift is off"""
        if not self.randshift:
            return (0,) * len(self.shape)
        self._calls += 1
        rng = np.random.default_rng([self.seed, self._calls])
        return tuple(int(rng.integers(1, b + 1)) for b in self.blockSize)


class PositiveRegularization(AbstractProjectionRegularization):
    """src/proximalMaps/ProxPositive.jl:8-9"""
    mask = capi.RLS_PROJ_POSITIVE


class RealRegularization(AbstractProjectionRegularization):
    """src/proximalMaps/ProxReal.jl:8-9"""
    mask = capi.RLS_PROJ_REAL


class NormalizedRegularization(AbstractRegularization):
    """src/Regularization/NormalizedRegularization.jl:30-38"""
    def __init__(self, reg, factor):
        self.reg = reg
        self.factor = factor


class NoNormalization:
    pass


class MeasurementBasedNormalization:
    pass


class SystemMatrixBasedNormalization:
    pass


def sink(reg):
    """NestedRegularization.jl:15"""
    while isinstance(reg, NormalizedRegularization):
        reg = reg.reg
    return reg


def lam(reg):
    """λ(reg): Regularization.jl:29, ScaledRegularization.jl:23"""
    if isinstance(reg, NormalizedRegularization):
        return lam(reg.reg) * reg.factor
    if isinstance(reg, AbstractProjectionRegularization):
        return None
    return reg.lam


def findsinks(cls, regs):
    """Regularization.jl:87"""
    return [i for i, r in enumerate(regs) if isinstance(sink(r), cls)]


def findsink(cls, regs):
    """Regularization.jl:76-85"""
    idx = findsinks(cls, regs)
    if not idx:
        return None
    if len(idx) == 1:
        return idx[0]
    raise ValueError(f"Cannot unambigiously retrieve reg term of type {cls.__name__}, found {len(idx)} instances")


def normalize_reg(reg, factor):
    """normalize(reg, factor): NormalizedRegularization.jl:69-78"""
    if factor is None or isinstance(reg, AbstractProjectionRegularization):
        return reg
    if isinstance(reg, NormalizedRegularization):
        return NormalizedRegularization(reg.reg, factor)
    return NormalizedRegularization(reg, factor)


def reg_desc(reg, rho=0.0, trafo=None):
    """Resolve a (possibly nested) term into the POD struct of the C ABI."""
    d = capi.RegDesc()
    s = sink(reg)
    d.kind = s.kind
    l = lam(reg)
    d.lambda_is_f64 = 0 if isinstance(l, np.float32) else 1
    d.lambda_ = float(l)
    d.slices = getattr(s, "slices", 1)
    d.rho = float(rho)
    d.trafo = capi.RLS_TRAFO_IDENTITY
    geom = None
    if isinstance(s, TVRegularization):
        geom = (s.shape, s.dims)
        d.tv_iterations = s.iterationsTV
    if isinstance(s, NuclearRegularization):           # svtShape in tv_shape[0..1] (include/rls_b200.h)
        d.tv_ndims = 2
        d.tv_shape[0], d.tv_shape[1] = s.svtShape
    if isinstance(s, LLRRegularization):               # shape / blockSize / flags / seed of the shift stream
        geom = (s.shape, s.blockSize)
        d.tv_iterations = (capi.RLS_LLR_RANDSHIFT if s.randshift else 0) | (capi.RLS_LLR_OVERLAPPING if s.fullyOverlapping else 0)
        d.slices = s.seed
    if trafo is not None:
        d.trafo = capi.RLS_TRAFO_GRADIENT
        geom = (trafo.shape, trafo.dims)
    if geom is not None:
        shape, dims = geom
        d.tv_ndims = len(shape)
        d.tv_ndirs = len(dims)
        for i, v in enumerate(shape):
            d.tv_shape[i] = v
        for i, v in enumerate(dims):
            d.tv_dims[i] = v
    return d


class GradientOp:
    """GradientOp(T; shape, dims) as ADMM's regTrafo (ADMM.jl:74): forward differences
    without boundary rows, one block per direction."""
    def __init__(self, dtype=np.float32, shape=(0,), dims=None):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        if dims is None:
            dims = range(1, len(self.shape) + 1)
        self.dims = (int(dims),) if np.isscalar(dims) else tuple(int(d) for d in dims)
        tot = int(np.prod(self.shape))
        self.rows = sum((self.shape[d - 1] - 1) * tot // self.shape[d - 1] for d in self.dims)
        self.cols = tot

    def _geom(self):
        shape = (C.c_int64 * len(self.shape))(*self.shape)
        dims = (C.c_int32 * max(1, len(self.dims)))(*self.dims)
        return len(self.shape), shape, len(self.dims), dims

    def mul(self, x):
        from .arrays import B200Vector
        out = B200Vector(x.ctx, x.dtype, self.rows)
        capi.call("rls_grad_apply", x.handle, out.handle, *self._geom())
        return out

    def tmul(self, g):
        from .arrays import B200Vector
        out = B200Vector(g.ctx, g.dtype, self.cols)
        capi.call("rls_grad_apply_t", g.handle, out.handle, *self._geom())
        return out
