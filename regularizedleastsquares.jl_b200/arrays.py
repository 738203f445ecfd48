"""Device arrays behind opaque library handles.

Python twins of the Julia shim's `B200Vector{T} <: AbstractVector{T}` and
`B200Matrix{T} <: AbstractMatrix{T}`: `similar`, `copyto!`, `Array(x)`,
`norm`, `dot` map to the C ABI.  Element types are Float32 and ComplexF32 only.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _capi as capi

_DT = {np.dtype(np.float32): capi.RLS_F32, np.dtype(np.complex64): capi.RLS_C32}
_NP = {capi.RLS_F32: np.dtype(np.float32), capi.RLS_C32: np.dtype(np.complex64)}


def dtype_code(dt):
    dt = np.dtype(dt)
    if dt not in _DT:
        raise TypeError(f"librls_b200 accelerates Float32 / ComplexF32 only, got {dt} (no CPU fallback)")
    return _DT[dt]


class B200Context:
    """One device + one stream (+ optionally one NCCL rank)."""
    _default = {}

    def __init__(self, device=0):
        h = C.c_void_p()
        capi.call("rls_ctx_create", int(device), C.byref(h))
        self.handle = h
        self.device = int(device)
        self._fin = weakref.finalize(self, capi.load().rls_ctx_destroy, h)
        self.rank, self.nranks = 0, 1

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def sync(self):
        capi.call("rls_ctx_sync", self.handle)

    def device_info(self):
        sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32()
        l2, hbm = C.c_int64(), C.c_int64()
        capi.call("rls_ctx_device_info", self.handle, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2), C.byref(hbm))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "l2_bytes": l2.value, "hbm_bytes": hbm.value}

    def timer_start(self):
        capi.call("rls_timer_start", self.handle)

    def timer_stop(self):
        ms = C.c_float()
        capi.call("rls_timer_stop", self.handle, C.byref(ms))
        return ms.value

    def launch_count(self):
        n = C.c_int64()
        capi.call("rls_ctx_launch_count", self.handle, C.byref(n))
        return n.value

    def flush_l2(self):
        capi.call("rls_ctx_flush_l2", self.handle)

    # ---- row-sharded multi-GPU: one process per GPU ----
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        capi.call("rls_comm_unique_id", buf)
        return buf.raw

    def comm_init(self, rank, nranks, uid):
        buf = C.create_string_buffer(bytes(uid), 128)
        capi.call("rls_ctx_comm_init", self.handle, int(rank), int(nranks), buf)
        self.rank, self.nranks = int(rank), int(nranks)

    def allreduce_f64(self, *vals):
        """Sum of up to 8 host scalars over the ranks (the global ‖A‖_F², ‖b‖₁, length(b) of a row-sharded solve)."""
        buf = (C.c_double * len(vals))(*[float(v) for v in vals])
        capi.call("rls_ctx_allreduce_f64", self.handle, buf, len(vals))
        return [float(v) for v in buf]


def _peer_export(self, max_floats):
    buf = C.create_string_buffer(64)
    capi.call("rls_ctx_peer_export", self.handle, int(max_floats), buf)
    return buf.raw


def _peer_import(self, handles):
    blob = b"".join(bytes(h) for h in handles)
    buf = C.create_string_buffer(blob, len(blob))
    capi.call("rls_ctx_peer_import", self.handle, buf, len(handles))


B200Context.peer_export = _peer_export
B200Context.peer_import = _peer_import


def _ctx(ctx):
    return ctx if ctx is not None else B200Context.default()


class B200Vector:
    def __init__(self, ctx, dtype, length, _handle=None, _owned=True):
        self.ctx = _ctx(ctx)
        self.dtype = np.dtype(dtype)
        self.length = int(length)
        if _handle is None:
            h = C.c_void_p()
            capi.call("rls_vec_create", self.ctx.handle, dtype_code(dtype), self.length, C.byref(h))
            self.handle = h
            self._fin = weakref.finalize(self, capi.load().rls_vec_destroy, h)
        else:
            self.handle = _handle     # borrowed from a solver; lifetime tied to it

    def __len__(self):
        return self.length

    @property
    def shape(self):
        return (self.length,)

    @classmethod
    def from_numpy(cls, x, ctx=None):
        x = np.ascontiguousarray(x)
        v = cls(ctx, x.dtype, x.size)
        v.upload(x)
        return v

    def upload(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype).ravel()
        capi.call("rls_vec_upload", self.handle, x.ctypes.data_as(C.c_void_p), x.size)
        self.ctx.sync()
        return self

    def to_numpy(self):
        out = np.empty(self.length, self.dtype)
        capi.call("rls_vec_download", self.handle, out.ctypes.data_as(C.c_void_p), out.size)
        return out

    __array__ = lambda self, dtype=None, copy=None: self.to_numpy() if dtype is None else self.to_numpy().astype(dtype)

    def similar(self, length=None):
        return B200Vector(self.ctx, self.dtype, self.length if length is None else length)

    def copy(self):
        v = self.similar()
        capi.call("rls_vec_copy", v.handle, self.handle)
        return v

    def copy_from(self, other):
        capi.call("rls_vec_copy", self.handle, other.handle)
        return self

    def fill(self, value):
        value = complex(value)
        capi.call("rls_vec_fill", self.handle, value.real, value.imag)
        return self

    def fill_philox(self, seed, stream=0, dist=capi.RLS_DIST_UNIFORM01, scale=1.0, offset=0):
        capi.call("rls_vec_fill_philox", self.handle, int(seed), int(stream), int(dist), float(scale), int(offset))
        return self

    def norm(self):
        out = C.c_double()
        capi.call("rls_vec_nrm2", self.handle, C.byref(out))
        return out.value

    def asum(self):
        out = C.c_double()
        capi.call("rls_vec_asum", self.handle, C.byref(out))
        return out.value

    def dot(self, other):
        out = (C.c_double * 2)()
        capi.call("rls_vec_dot", self.handle, other.handle, out)
        return complex(out[0], out[1]) if self.dtype.kind == "c" else out[0]

    def allreduce(self):
        capi.call("rls_vec_allreduce", self.handle)
        return self

    def device_ptr(self):
        p = C.c_void_p()
        capi.call("rls_vec_device_ptr", self.handle, C.byref(p))
        return p.value


_LAYOUTS = {"col": capi.RLS_LAYOUT_COLMAJOR, "row": capi.RLS_LAYOUT_ROWMAJOR, "auto": capi.RLS_LAYOUT_AUTO}
_LAYOUT_NAMES = {capi.RLS_LAYOUT_COLMAJOR: "col", capi.RLS_LAYOUT_ROWMAJOR: "row"}


class B200Matrix:
    """Dense system matrix (or this rank's row shard of it) in HBM.

    The host side is always column-major (Julia `Matrix`).  `layout="row"` keeps the ROWS contiguous on
    the device, which lets the normal operator A'(A x) sweep HBM once (csrc/rls_rowstream.cu); `layout="col"`
    mirrors the host storage (what `rls_mat_wrap_device` adopts from a CuArray); `"auto"` (default) lets
    the library choose (row-major whenever the one-pass kernel supports the shape)."""
    def __init__(self, ctx, dtype, m, n, host=None, layout="auto"):
        self.ctx = _ctx(ctx)
        self.dtype = np.dtype(dtype)
        self.m, self.n = int(m), int(n)
        h = C.c_void_p()
        ptr, ld = None, self.m
        if host is not None:
            host = np.asfortranarray(host, dtype=self.dtype)
            assert host.shape == (self.m, self.n)
            ptr = host.ctypes.data_as(C.c_void_p)
        capi.call("rls_mat_create_layout", self.ctx.handle, dtype_code(dtype), self.m, self.n, ptr, ld, _LAYOUTS[layout],
                  C.byref(h))
        self.handle = h
        self._fin = weakref.finalize(self, capi.load().rls_mat_destroy, h)
        lay = C.c_int32()
        capi.call("rls_mat_layout", h, C.byref(lay))
        self.layout = _LAYOUT_NAMES[lay.value]

    @property
    def shape(self):
        return (self.m, self.n)

    @classmethod
    def from_numpy(cls, A, ctx=None, layout="auto"):
        A = np.asarray(A)
        return cls(ctx, A.dtype, A.shape[0], A.shape[1], host=A, layout=layout)

    @classmethod
    def philox(cls, dtype, m, n, seed, dist=capi.RLS_DIST_IH4, scale=1.0, row_offset=0, m_global=None, ctx=None,
               layout="auto"):
        """Generate A (or rows [row_offset, row_offset+m) of a global m_global x n matrix) on the device."""
        A = cls(ctx, dtype, m, n, layout=layout)
        capi.call("rls_mat_fill_philox", A.handle, int(seed), int(dist), float(scale), int(row_offset),
                  int(m if m_global is None else m_global))
        return A

    def to_numpy(self):
        out = np.empty((self.m, self.n), self.dtype, order="F")
        capi.call("rls_mat_download", self.handle, out.ctypes.data_as(C.c_void_p), self.m)
        return out

    def relayout(self, layout):
        """A copy in the other device layout, made on the device (tiled transpose at HBM speed)."""
        B = object.__new__(B200Matrix)
        B.ctx, B.dtype, B.m, B.n = self.ctx, self.dtype, self.m, self.n
        h = C.c_void_p()
        capi.call("rls_mat_relayout", self.handle, _LAYOUTS[layout], C.byref(h))
        B.handle = h
        B._fin = weakref.finalize(B, capi.load().rls_mat_destroy, h)
        lay = C.c_int32()
        capi.call("rls_mat_layout", h, C.byref(lay))
        B.layout = _LAYOUT_NAMES[lay.value]
        return B

    def frob2(self):
        out = C.c_double()
        capi.call("rls_mat_frob2", self.handle, C.byref(out))
        return out.value

    def mul(self, x, out=None):
        """mul!(y, A, x)"""
        out = B200Vector(self.ctx, self.dtype, self.m) if out is None else out
        capi.call("rls_gemv_n", self.handle, x.handle, out.handle)
        return out

    def adjoint_mul(self, y, out=None):
        """mul!(g, adjoint(A), y)"""
        out = B200Vector(self.ctx, self.dtype, self.n) if out is None else out
        capi.call("rls_gemv_c", self.handle, y.handle, out.handle)
        return out


_FORMS = {"twopass": capi.RLS_NORMAL_TWOPASS, "onepass": capi.RLS_NORMAL_ONEPASS, "gram": capi.RLS_NORMAL_GRAM,
          "auto": capi.RLS_NORMAL_AUTO, "lazy": capi.RLS_NORMAL_AUTO}
_FORM_NAMES = {capi.RLS_NORMAL_TWOPASS: "twopass", capi.RLS_NORMAL_ONEPASS: "onepass", capi.RLS_NORMAL_GRAM: "gram",
               capi.RLS_NORMAL_MATRIXFREE: "matrixfree"}


class B200NormalOp:
    """AHA: `normalOperator(A)` (lazy forms), `A'*A` (Gram), the normal operator of a matrix-free operator (`linop=`,
    operators.py) or a callback working on device pointers (`from_callback`)."""
    def __init__(self, A=None, form="auto", G=None, linop=None):
        h = C.c_void_p()
        if linop is not None:
            self.A, self.G, self.linop = None, None, linop
            capi.call("rls_normal_from_linop", linop.handle, C.byref(h))
            self.ctx, self.dtype, self.n = linop.ctx, linop.dtype, linop.shape[1]
        elif G is not None:
            self.A, self.G = None, G
            capi.call("rls_normal_from_gram", G.handle, C.byref(h))
            self.ctx, self.dtype, self.n = G.ctx, G.dtype, G.n
        else:
            self.A, self.G = A, None
            capi.call("rls_normal_create", A.handle, _FORMS[form], C.byref(h))
            self.ctx, self.dtype, self.n = A.ctx, A.dtype, A.n
        self.handle = h
        self._fin = weakref.finalize(self, capi.load().rls_normal_destroy, h)

    @classmethod
    def from_callback(cls, fn, dtype, n, ctx=None):
        """AHA as a function fn(x_ptr, res_ptr, stream) -> status that enqueues res = AHA x on the CUDA stream `stream`
        (device pointers as integers) — the hook a Julia `@cfunction` around any LinearOperator uses
        (rls_normal_from_callback)."""
        ctx = ctx if ctx is not None else B200Context.default()
        self = cls.__new__(cls)

        def tramp(_user, x, res, stream):
            try:
                r = fn(x, res, stream)
                return int(r) if r is not None else 0
            except Exception:          # an exception cannot cross the C frames
                import traceback
                traceback.print_exc()
                return 1          # RLS_ERR_INVALID
        self._cb = capi.APPLY_FN(tramp)
        h = C.c_void_p()
        capi.call("rls_normal_from_callback", ctx.handle, _DT[np.dtype(dtype)], int(n), self._cb, None, C.byref(h))
        self.A, self.G = None, None
        self.ctx, self.dtype, self.n = ctx, np.dtype(dtype), int(n)
        self.handle = h
        self._fin = weakref.finalize(self, capi.load().rls_normal_destroy, h)
        return self

    @property
    def form(self):
        f = C.c_int32()
        capi.call("rls_normal_form", self.handle, C.byref(f))
        return _FORM_NAMES[f.value]

    def describe(self):
        buf = C.create_string_buffer(256)
        capi.call("rls_normal_describe", self.handle, buf, 256)
        return buf.value.decode()

    def apply(self, x, out=None):
        """mul!(res, AHA, x)"""
        out = B200Vector(self.ctx, self.dtype, self.n) if out is None else out
        capi.call("rls_normal_apply", self.handle, x.handle, out.handle)
        return out

    def apply_batch(self, xs, outs=None):
        """res_k = AHA x_k for a list of vectors (tensor-core GEMM path when A is row-major)"""
        K = len(xs)
        outs = [B200Vector(self.ctx, self.dtype, self.n) for _ in range(K)] if outs is None else outs
        xa = (C.c_void_p * K)(*[x.handle for x in xs])
        oa = (C.c_void_p * K)(*[o.handle for o in outs])
        capi.call("rls_normal_apply_batch", self.handle, K, xa, oa)
        return outs

    def power_iterations(self, b0, rtol=1e-3, maxiter=30):
        lam = C.c_double()
        capi.call("rls_power_iterations", self.handle, b0.handle, float(rtol), int(maxiter), C.byref(lam))
        return lam.value


# ------------------------------------------------------------------------------------------------------------------
# single-process multi-device: one Python / Julia process, the system row-partitioned over a group of GPUs
# ------------------------------------------------------------------------------------------------------------------
class B200Group:
    """A group of devices with one NCCL communicator (ncclCommInitAll).  `B200GroupMatrix` row-partitions A over it and
    createLinearSolver(FISTA, A_group; ...) / solve_(solver, b) then run the row-sharded solve inside ONE call."""
    def __init__(self, ndev=None, devices=None):
        if devices is None:
            if ndev is None:
                n = C.c_int32()
                capi.call("rls_device_count", C.byref(n))
                ndev = n.value
            devices = list(range(int(ndev)))
        self.devices = [int(d) for d in devices]
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        h = C.c_void_p()
        capi.call("rls_group_create", len(self.devices), arr, C.byref(h))
        self.handle = h
        self._fin = weakref.finalize(self, capi.load().rls_group_destroy, h)

    def __len__(self):
        return len(self.devices)


class B200GroupMatrix:
    """Dense m x n system matrix in contiguous row blocks on the devices of a group (host side: column-major, as ever)."""
    def __init__(self, group, dtype, m, n, host=None):
        self.group = group
        self.dtype = np.dtype(dtype)
        self.m, self.n = int(m), int(n)
        ptr = None
        if host is not None:
            host = np.asfortranarray(host, dtype=self.dtype)
            assert host.shape == (self.m, self.n)
            ptr = host.ctypes.data_as(C.c_void_p)
        h = C.c_void_p()
        capi.call("rls_group_mat_create", group.handle, dtype_code(dtype), self.m, self.n, ptr, self.m, C.byref(h))
        self.handle = h
        self._fin = weakref.finalize(self, capi.load().rls_group_mat_destroy, h)

    @classmethod
    def from_numpy(cls, A, group):
        A = np.asarray(A)
        return cls(group, A.dtype, A.shape[0], A.shape[1], host=A)

    @classmethod
    def philox(cls, group, dtype, m, n, seed, dist=capi.RLS_DIST_IH4, scale=1.0):
        A = cls(group, dtype, m, n)
        capi.call("rls_group_mat_fill_philox", A.handle, int(seed), int(dist), float(scale))
        return A

    @property
    def shape(self):
        return (self.m, self.n)

    def row_blocks(self):
        out = []
        for i in range(len(self.group)):
            lo, hi = C.c_int64(), C.c_int64()
            capi.call("rls_group_mat_part", self.handle, i, None, C.byref(lo), C.byref(hi))
            out.append((lo.value, hi.value))
        return out
