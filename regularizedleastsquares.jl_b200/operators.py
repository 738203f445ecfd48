"""Matrix-free system operators of the documentation's examples, device-resident:

  SamplingOp(T; pattern, shape)          docs/src/literate/examples/compressed_sensing.jl:23
  FFTOp(T; shape, shift=true, unitary=true)   LinearOperatorCollection.jl (cuFFT behind rls_linop_fft_create)
  outer * inner                          product of operators, e.g. SamplingOp(...) * FFTOp(...)

`createLinearSolver(FISTA, op; reg=..., ...)` builds the solver on the lazy normal operator op'op (rls_normal_from_linop)
and `solve_(solver, b)` back-projects b with rls_linop_mul_adjoint first — the AHA-only interface of the reference
(FISTA.jl:55, test/testSolvers.jl:44-65).  Vectors are B200Vector or NumPy (staged through HBM; no CPU implementation).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _capi as capi
from .arrays import _DT, _NP, B200Context, B200NormalOp, B200Vector


class LinearOperator:
    def __init__(self, handle, ctx, keep=()):
        self.handle, self.ctx, self._keep = handle, ctx, keep
        m, n, dt = C.c_int64(), C.c_int64(), C.c_int32()
        capi.call("rls_linop_shape", handle, C.byref(m), C.byref(n), C.byref(dt))
        self.shape = (m.value, n.value)
        self.dtype = _NP[dt.value]
        self._fin = weakref.finalize(self, capi.load().rls_linop_destroy, handle)

    def _dev(self, v, length):
        if isinstance(v, B200Vector):
            return v, False
        a = np.ascontiguousarray(np.asarray(v).ravel(order="F"), dtype=self.dtype)
        if a.size != length:
            raise ValueError(f"operator is {self.shape[0]}x{self.shape[1]}, vector has {a.size} elements")
        return B200Vector.from_numpy(a, self.ctx), True

    def mul(self, x):
        """A * x   (mul!(y, A, x))"""
        xd, host = self._dev(x, self.shape[1])
        y = B200Vector(self.ctx, self.dtype, self.shape[0])
        capi.call("rls_linop_mul", self.handle, xd.handle, y.handle)
        return y.to_numpy() if host else y

    def tmul(self, y):
        """adjoint(A) * y   (mul!(x, adjoint(A), y))"""
        yd, host = self._dev(y, self.shape[0])
        x = B200Vector(self.ctx, self.dtype, self.shape[1])
        capi.call("rls_linop_mul_adjoint", self.handle, yd.handle, x.handle)
        return x.to_numpy() if host else x

    def normal(self):
        """normalOperator(A): the lazy A'A as a B200NormalOp"""
        return B200NormalOp(linop=self)

    def __mul__(self, other):
        if isinstance(other, LinearOperator):
            h = C.c_void_p()
            capi.call("rls_linop_compose", self.handle, other.handle, C.byref(h))
            return LinearOperator(h, self.ctx, keep=(self, other))
        return self.mul(other)

    __matmul__ = __mul__


class SamplingOp(LinearOperator):
    """SamplingOp(T; pattern, shape): y = vec(x)[pattern]; `pattern` holds Julia (1-based) linear indices."""
    def __init__(self, dtype, pattern, shape, ctx=None):
        ctx = ctx if ctx is not None else B200Context.default()
        pat = np.ascontiguousarray(np.asarray(pattern).ravel(), dtype=np.int64)
        n = int(np.prod(shape))
        h = C.c_void_p()
        capi.call("rls_linop_sampling_create", ctx.handle, _DT[np.dtype(dtype)], n, pat.size,
                  pat.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(h))
        super().__init__(h, ctx)


class FFTOp(LinearOperator):
    """FFTOp(ComplexF32; shape, shift=true, unitary=true): the centred, unitary n-dimensional DFT."""
    def __init__(self, dtype=np.complex64, shape=(1,), shift=True, unitary=True, ctx=None):
        if np.dtype(dtype) != np.complex64:
            raise TypeError("FFTOp: ComplexF32 only on this path")
        ctx = ctx if ctx is not None else B200Context.default()
        shape = tuple(int(s) for s in shape)
        h = C.c_void_p()
        capi.call("rls_linop_fft_create", ctx.handle, len(shape), (C.c_int64 * len(shape))(*shape), 1 if shift else 0,
                  1 if unitary else 0, C.byref(h))
        super().__init__(h, ctx)
