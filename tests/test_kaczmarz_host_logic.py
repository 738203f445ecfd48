"""Host-side logic of rls_b200.Kaczmarz (L2 / Tikhonov handling, denom and rowindex, row order, normalisation, prox
after the sweep, solversolution) on the CPU: the C ABI is replaced by an in-test stand-in that stores vectors in NumPy
and performs the row loop exactly as Kaczmarz.jl:305-310 states it, so that whatever differs from the oracle is a bug
of the Python host code.  (The CUDA path itself is tested in test_gpu_kaczmarz.py.)"""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from util import rel


def _host(ptr, dtype, count):
    addr = ptr.value if hasattr(ptr, "value") else int(ptr)
    nbytes = int(count) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_char * nbytes).from_address(addr), dtype=dtype, count=int(count))


def _handle(h):
    return h.value if hasattr(h, "value") else h


class FakeABI:
    """just enough of include/rls_b200.h for the Kaczmarz flow"""
    DT = {0: np.float32, 1: np.complex64}

    def __init__(self):
        self.obj = {}
        self.next = 100
        self.calls = []

    def new(self, o):
        self.next += 1
        self.obj[self.next] = o
        return self.next

    def call(self, name, *a):
        self.calls.append(name)
        return getattr(self, name)(*a)

    # ---- vectors
    def rls_vec_create(self, ctx, dt, n, out):
        out._obj.value = self.new(np.zeros(n, self.DT[dt]))

    def rls_vec_upload(self, h, ptr, n):
        v = self.obj[_handle(h)]
        v[...] = _host(ptr, v.dtype, n)

    def rls_vec_download(self, h, ptr, n):
        v = self.obj[_handle(h)]
        _host(ptr, v.dtype, n)[...] = v

    def rls_vec_len(self, h, ln, dt):
        v = self.obj[_handle(h)]
        ln._obj.value = v.size
        dt._obj.value = 1 if v.dtype == np.complex64 else 0

    def rls_vec_asum(self, h, out):
        v = self.obj[_handle(h)]
        out._obj.value = float(np.sum(np.abs(v.astype(np.complex128))))

    def rls_ctx_sync(self, ctx):
        pass

    # ---- matrices
    def rls_mat_create_layout(self, ctx, dt, m, n, ptr, ld, layout, out):
        A = _host(ptr, self.DT[dt], ld * n).reshape((n, ld)).T[:m].copy()       # host side is column-major
        out._obj.value = self.new({"A": A, "layout": layout})

    def rls_mat_layout(self, h, out):
        out._obj.value = self.obj[_handle(h)]["layout"]

    def rls_mat_frob2(self, h, out):
        A = self.obj[_handle(h)]["A"]
        out._obj.value = float(np.sum(np.abs(A.astype(np.complex128)) ** 2))

    # ---- prox (in place)
    def rls_prox_l1(self, h, lam):
        O.prox_(O.L1Regularization(np.float32(lam)), self.obj[_handle(h)])

    def rls_prox_l2(self, h, lam):
        O.prox_(O.L2Regularization(np.float32(lam)), self.obj[_handle(h)])

    def rls_prox_positive(self, h):
        O.prox_(O.PositiveRegularization(), self.obj[_handle(h)])

    def rls_prox_real(self, h):
        O.prox_(O.RealRegularization(), self.obj[_handle(h)])

    # ---- Kaczmarz
    def rls_kaczmarz_create(self, Ah, block_rows, out):
        M = self.obj[_handle(Ah)]
        assert M["layout"] == 1, "the host code must ask for the row-major device layout"
        A = M["A"]
        K = {"A": A, "x": np.zeros(A.shape[1], A.dtype), "vl": np.zeros(A.shape[0], A.dtype), "u": np.zeros(A.shape[0], A.dtype),
             "rows": None, "denom": None, "eps": None, "sets": 0, "block_rows": block_rows or 128}
        K["hx"], K["hvl"], K["hu"] = self.new(K["x"]), self.new(K["vl"]), self.new(K["u"])
        out._obj.value = self.new(K)

    def rls_kaczmarz_block_rows(self, h, out):
        out._obj.value = self.obj[_handle(h)]["block_rows"]

    def rls_kaczmarz_rownorm2(self, h, ptr, n):
        A = self.obj[_handle(h)]["A"]
        a2 = (A.real.astype(np.float32) ** 2 + A.imag.astype(np.float32) ** 2) if np.iscomplexobj(A) else A.astype(np.float32) ** 2
        _host(ptr, np.float32, n)[...] = np.sum(a2, axis=1, dtype=np.float32)

    def rls_kaczmarz_set_rows(self, h, rows, denom, count):
        K = self.obj[_handle(h)]
        K["rows"] = _host(rows, np.int64, count).copy()
        K["denom"] = _host(denom, np.float32, count).copy()
        assert len(set(K["rows"].tolist())) == count
        K["sets"] += 1

    def rls_kaczmarz_init(self, h, b, x0, eps_w):
        K = self.obj[_handle(h)]
        K["x"][...] = 0 if x0 is None else self.obj[_handle(x0)]
        K["vl"][...] = 0
        K["u"][...] = self.obj[_handle(b)]
        K["eps"] = np.float32(eps_w)

    def rls_kaczmarz_sweep(self, h):
        K = self.obj[_handle(h)]
        A, x, vl, u, ew = K["A"], K["x"], K["vl"], K["u"], K["eps"]
        T = A.dtype.type
        for i, row in enumerate(K["rows"]):                                  # Kaczmarz.jl:305-310
            tau = T(np.dot(A[row], x))
            alpha = T(K["denom"][i] * (u[row] - tau - ew * vl[row]))
            x += alpha * np.conj(A[row])
            vl[row] += alpha * ew

    def rls_kaczmarz_vec(self, h, name, out):
        out._obj.value = self.obj[_handle(h)]["h" + name.decode()]

    def rls_kaczmarz_check(self, h):
        pass

    def rls_kaczmarz_describe(self, h, buf, n):
        buf.value = b"persistent: in-test stand-in"


class _NoopLib:
    def __getattr__(self, name):
        return lambda *a: 0


@pytest.fixture
def fake(rls, monkeypatch):
    f = FakeABI()
    monkeypatch.setattr(rls._capi, "call", f.call)
    monkeypatch.setattr(rls._capi, "load", lambda: _NoopLib())               # finalizers must not reach the real library
    ctx = rls.B200Context.__new__(rls.B200Context)
    ctx.handle, ctx.device, ctx.rank, ctx.nranks = C.c_void_p(1), 0, 0, 1
    f.ctx = ctx
    return f


def _system(dtype, m, n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n)).astype(np.float32) / np.float32(np.sqrt(m))
    if dtype == np.complex64:
        A = (A + 1j * rng.standard_normal((m, n)).astype(np.float32) / np.float32(np.sqrt(m)))
    A = A.astype(dtype)
    x = rng.standard_normal(n).astype(dtype)
    return A, x, (A @ x).astype(dtype)


def _both(rls, fake, A, kw_host, kw_oracle, b, iters, x0=None):
    S = rls.createLinearSolver(rls.Kaczmarz, A, ctx=fake.ctx, iterations=iters, **kw_host)
    R = O.createLinearSolver(O.Kaczmarz, A, iterations=iters, **kw_oracle)
    kw = {} if x0 is None else {"x0": x0}
    S.init_(b, **kw); R.init(b, **kw)
    for k in range(iters + 1):
        a1, a2 = S.iterate(), R.iterate()
        assert a1 == a2
        if not a1:
            break
        assert rel(S.x, R.solution()) < 2e-6, f"iteration {k + 1}"
    assert S.iteration == R.iteration == iters
    return S, R


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_plain_shuffled_randomized_orders(rls, fake, dtype):
    A, x, b = _system(dtype, 60, 24, 1)
    for kw in ({}, {"shuffleRows": True, "seed": 7}, {"randomized": True, "subMatrixFraction": 0.3, "seed": 7}):
        l2h, l2o = rls.L2Regularization(np.float32(1e-2)), O.L2Regularization(np.float32(1e-2))
        S, R = _both(rls, fake, A, dict(reg=l2h, **kw), dict(reg=l2o, **kw), b, 5)
        assert np.array_equal(S.rowindex, R.rowindex) and np.allclose(S.denom, R.denom, rtol=1e-6)
    K = [o for o in fake.obj.values() if isinstance(o, dict) and "sets" in o]
    assert [k["sets"] for k in K] == [1, 1, 5], "one row order per solve, except randomized: one per iteration"


def test_zero_rows_x0_and_float64_lambda(rls, fake):
    A, x, b = _system(np.complex64, 40, 16, 2)
    A[[0, 13, 39]] = 0
    b = (A @ x).astype(np.complex64)
    x0 = np.random.default_rng(3).standard_normal(16).astype(np.complex64)
    S, R = _both(rls, fake, A, dict(reg=rls.L2Regularization(0.05)), dict(reg=O.L2Regularization(0.05)), b, 4, x0=x0)
    assert len(S.rowindex) == 37


@pytest.mark.parametrize("strategy", ["NoNormalization", "MeasurementBasedNormalization", "SystemMatrixBasedNormalization"])
def test_normalization_and_extra_terms(rls, fake, strategy):
    A, x, b = _system(np.complex64, 48, 20, 4)
    regs = lambda M: [M.L2Regularization(0.1), M.L1Regularization(np.float32(1e-3)), M.RealRegularization()]
    S, R = _both(rls, fake, A, dict(reg=regs(rls), normalizeReg=getattr(rls, strategy)()),
                 dict(reg=regs(O), normalizeReg=getattr(O, strategy)()), b, 4)
    assert abs(float(rls.lam(S.L2)) - float(O.lam_of(R.L2))) <= 1e-6 * abs(float(O.lam_of(R.L2)))
    assert "rls_prox_l1" in fake.calls and "rls_prox_real" in fake.calls


def test_tikhonov_matrix_and_solution_scaling(rls, fake):
    A, x, b = _system(np.complex64, 36, 12, 5)
    lamv = (np.random.default_rng(6).random(12) + 0.2).astype(np.float32)
    S = rls.createLinearSolver(rls.Kaczmarz, A, ctx=fake.ctx, iterations=6, reg=[rls.L2Regularization(lamv)])
    R = O.createLinearSolver(O.Kaczmarz, A, iterations=6, reg=[O.L2Regularization(lamv)])
    xs, xr = rls.solve_(S, b), R.solve(b)
    assert rel(xs, xr) < 2e-6
    # the device matrix is A·diag(1/sqrt(λ)) and eps_w = 1 (Kaczmarz.jl:377-392, :210-211)
    K = [o for o in fake.obj.values() if isinstance(o, dict) and "sets" in o][-1]
    assert rel(K["A"], A * (1 / np.sqrt(lamv))[None, :]) < 1e-6 and K["eps"] == np.float32(1)
    with pytest.raises(ValueError):
        rls.createLinearSolver(rls.Kaczmarz, A, ctx=fake.ctx, reg=[rls.L2Regularization(lamv)],
                               normalizeReg=rls.MeasurementBasedNormalization())


def test_constructor_errors(rls, fake):
    A, x, b = _system(np.float32, 20, 8, 7)
    with pytest.raises(ValueError):
        rls.Kaczmarz(A, ctx=fake.ctx, reg=[rls.L1Regularization(np.float32(1e-3)), rls.L21Regularization(np.float32(1e-3))])
    with pytest.raises(NotImplementedError):
        rls.Kaczmarz(A, ctx=fake.ctx, greedy_randomized=True)
    with pytest.raises(TypeError):
        rls.Kaczmarz(A.astype(np.float64), ctx=fake.ctx)
    with pytest.raises(ValueError):
        rls.Kaczmarz(None, ctx=fake.ctx)
