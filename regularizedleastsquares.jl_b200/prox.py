"""prox!(reg, x[, λ]) and norm(reg, x[, λ]) on the device.

Mirrors src/Regularization/Regularization.jl:13-61 (dispatch, λ conversion to the real
type of x) and the GPU methods the reference adds for `AbstractGPUArray`
(ext/RegularizedLeastSquaresGPUArraysExt/{ProxL21,ProxTV,Utils}.jl) — here as
hand-written sm_100a kernels behind the C ABI.  A NumPy argument is staged through
HBM (upload → kernel → download in place); there is no CPU implementation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .arrays import B200Vector
from .regularization import (AbstractProjectionRegularization, L1Regularization, L2Regularization, L21Regularization,
                             LLRRegularization, NormalizedRegularization, NuclearRegularization, PositiveRegularization,
                             RealRegularization, TVRegularization, lam as lam_of, sink)


def _device_prox(reg, v, lam, shift=None):
    s = sink(reg)
    if isinstance(s, PositiveRegularization):
        capi.call("rls_prox_positive", v.handle)
        return
    if isinstance(s, RealRegularization):
        capi.call("rls_prox_real", v.handle)
        return
    lam = np.float32(lam)          # convert(T, λ)   Regularization.jl:31
    if isinstance(s, L1Regularization):
        capi.call("rls_prox_l1", v.handle, lam)
    elif isinstance(s, L2Regularization):
        capi.call("rls_prox_l2", v.handle, lam)
    elif isinstance(s, L21Regularization):
        capi.call("rls_prox_l21", v.handle, lam, s.slices)
    elif isinstance(s, TVRegularization):
        shape = (C.c_int64 * len(s.shape))(*s.shape)
        dims = (C.c_int32 * max(1, len(s.dims)))(*s.dims)
        capi.call("rls_prox_tv", v.handle, lam, len(s.shape), shape, len(s.dims), dims, s.iterationsTV)
    elif isinstance(s, NuclearRegularization):
        capi.call("rls_prox_nuclear", v.handle, lam, s.svtShape[0], s.svtShape[1])
    elif isinstance(s, LLRRegularization):
        nd = len(s.shape)
        capi.call("rls_prox_llr", v.handle, lam, nd, (C.c_int64 * nd)(*s.shape), (C.c_int64 * nd)(*s.blockSize),
                  (C.c_int64 * nd)(*(shift if shift is not None else s.next_shift())), 1 if s.fullyOverlapping else 0)
    else:
        raise TypeError(f"prox! is not accelerated for {type(s).__name__} (no CPU fallback)")


def prox_(reg, x, lam=None, shift=None, **kwargs):
    """prox!(reg, x[, λ]; kwargs...) — in place, returns x.

    `reg` may be an instance, or a type (then `reg(λ; kwargs...)` is constructed first,
    Regularization.jl:39,55).  `shift` (LLRRegularization only) fixes the patch-grid shift that `randshift` would draw."""
    if isinstance(reg, type):
        if issubclass(reg, AbstractProjectionRegularization):
            reg = reg()
        else:
            reg = reg(lam, **kwargs)
    if lam is None and not isinstance(sink(reg), AbstractProjectionRegularization):
        lam = lam_of(reg)
    if isinstance(x, B200Vector):
        _device_prox(reg, x, lam, shift)
        return x
    arr = np.asarray(x)
    if arr.dtype not in (np.float32, np.complex64):
        raise TypeError(f"prox!: Float32 / ComplexF32 only on this path, got {arr.dtype}")
    v = B200Vector.from_numpy(arr.ravel(order="F"))
    _device_prox(reg, v, lam, shift)
    out = v.to_numpy().reshape(arr.shape, order="F")
    if isinstance(x, np.ndarray):
        x[...] = out
        return x
    return out
