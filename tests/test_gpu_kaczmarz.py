"""Kaczmarz (src/Kaczmarz.jl) on the device against the oracle's sequential row loop: rel-L2 <= 1e-5 after every
iteration (the block-Gram evaluation is algebraically the same recurrence), plus the properties the reference's own
tests pin (test/testKaczmarz.jl:37-125)."""
import numpy as np
import pytest

import oracle as O
from util import rel, rand_matrix, rand_vector

pytestmark = pytest.mark.gpu

TOL = 1e-5
DTYPES = [np.float32, np.complex64]


@pytest.fixture(autouse=True, params=["persistent", "chained"])
def sweep_path(request, monkeypatch):
    """every test runs on both sweep implementations: one cooperative kernel per iteration, and three chained kernels
    per block (the path for block sizes / shapes the persistent kernel does not take)"""
    monkeypatch.setenv("RLS_KACZMARZ_PERSISTENT", "1" if request.param == "persistent" else "0")
    return request.param


def system(dtype, m, n, seed=300):
    A, _ = rand_matrix(dtype, m, n, seed)
    xt = rand_vector(dtype, n, seed + 1)
    b = (A @ xt).astype(dtype)
    return A, xt, b


def stepwise(S, R, b, iters, tol=TOL, **kw):
    S.init_(b, **kw); R.init(b, **kw)
    for k in range(iters + 1):
        a1, a2 = S.iterate(), R.iterate()
        assert a1 == a2, f"stopping decision differs at iteration {k}"
        if not a1:
            break
        e = rel(S.x, R.solution())
        assert e < tol, f"iterate {k + 1}: rel-L2 {e:.3e}"
        ev = rel(S._vec("vl").to_numpy(), R.vl) if np.linalg.norm(R.vl) > 0 else 0.0
        assert ev < 10 * tol, f"iterate {k + 1}: vl rel-L2 {ev:.3e}"
    assert S.iteration == R.iteration


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(300, 200), (130, 67), (64, 512), (700, 96)])
@pytest.mark.parametrize("lam", [0.0, 5e-2])
@pytest.mark.parametrize("block_rows", [0, 64])
def test_per_iterate(rls, ctx, dtype, shape, lam, block_rows):
    A, xt, b = system(dtype, *shape)
    lam = np.float32(lam)
    S = rls.createLinearSolver(rls.Kaczmarz, A, reg=rls.L2Regularization(lam), iterations=6, block_rows=block_rows)
    R = O.createLinearSolver(O.Kaczmarz, A, reg=O.L2Regularization(lam), iterations=6)
    stepwise(S, R, b, 6)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("block_rows", [128, 192, 256])
def test_block_sizes_and_float64_lambda(rls, ctx, dtype, block_rows):
    A, xt, b = system(dtype, 900, 256, seed=310)
    S = rls.createLinearSolver(rls.Kaczmarz, A, reg=rls.L2Regularization(0.01), iterations=4, block_rows=block_rows)
    assert S.block_rows == block_rows
    R = O.createLinearSolver(O.Kaczmarz, A, reg=O.L2Regularization(0.01), iterations=4)
    stepwise(S, R, b, 4)


@pytest.mark.parametrize("dtype", DTYPES)
def test_larger_system(rls, ctx, dtype, sweep_path):
    """rows of 16384 elements: several column chunks in the dot kernel / more than one pack per lane in the sweep kernel"""
    A, xt, b = system(dtype, 1024, 16384, seed=320)
    S = rls.createLinearSolver(rls.Kaczmarz, A, reg=rls.L2Regularization(np.float32(1e-2)), iterations=3)
    assert S.init_(b).describe().startswith(sweep_path)
    R = O.createLinearSolver(O.Kaczmarz, A, reg=O.L2Regularization(np.float32(1e-2)), iterations=3)
    stepwise(S, R, b, 3)


@pytest.mark.parametrize("dtype", DTYPES)
def test_zero_rows_are_skipped_and_x0(rls, ctx, dtype):
    A, xt, b = system(dtype, 200, 120, seed=330)
    A = A.copy(); A[[0, 17, 63, 64, 199], :] = 0
    b = (A @ xt).astype(dtype)
    S = rls.createLinearSolver(rls.Kaczmarz, A, iterations=5)
    R = O.createLinearSolver(O.Kaczmarz, A, iterations=5)
    assert len(S.rowindex) == 195 and np.array_equal(S.rowindex, R.rowindex)
    x0 = rand_vector(dtype, 120, 77)
    stepwise(S, R, b, 5, x0=x0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_shuffle_and_randomized_match_oracle_order(rls, ctx, dtype):
    A, xt, b = system(dtype, 256, 96, seed=340)
    for kw in (dict(shuffleRows=True), dict(randomized=True, subMatrixFraction=0.4)):
        S = rls.createLinearSolver(rls.Kaczmarz, A, reg=rls.L2Regularization(np.float32(1e-3)), iterations=5, seed=99, **kw)
        R = O.createLinearSolver(O.Kaczmarz, A, reg=O.L2Regularization(np.float32(1e-3)), iterations=5, seed=99, **kw)
        stepwise(S, R, b, 5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_additional_regularization_terms(rls, ctx, dtype):
    """prox! of the projections and of one more term after every sweep (Kaczmarz.jl:275-277)"""
    A, xt, b = system(dtype, 240, 128, seed=350)
    regs = lambda M: [M.L2Regularization(np.float32(1e-2)), M.L1Regularization(np.float32(1e-3)), M.PositiveRegularization()]
    S = rls.createLinearSolver(rls.Kaczmarz, A, reg=regs(rls), iterations=5)
    R = O.createLinearSolver(O.Kaczmarz, A, reg=regs(O), iterations=5)
    stepwise(S, R, b, 5)
    with pytest.raises(ValueError):
        rls.createLinearSolver(rls.Kaczmarz, A, reg=[rls.L1Regularization(np.float32(1e-3)), rls.L21Regularization(np.float32(1e-3))])


@pytest.mark.parametrize("strategy", ["SystemMatrixBasedNormalization", "MeasurementBasedNormalization"])
def test_normalization(rls, ctx, strategy):
    A, xt, b = system(np.complex64, 180, 64, seed=360)
    S = rls.createLinearSolver(rls.Kaczmarz, A, reg=rls.L2Regularization(0.1), normalizeReg=getattr(rls, strategy)(), iterations=5)
    R = O.createLinearSolver(O.Kaczmarz, A, reg=O.L2Regularization(0.1), normalizeReg=getattr(O, strategy)(), iterations=5)
    stepwise(S, R, b, 5)


def test_reference_properties(rls, ctx):
    """test/testKaczmarz.jl:94-125 (parameters) and :37-70 (Tikhonov matrix), in single precision"""
    rng = np.random.default_rng(12345)
    M, N = 12, 8
    A = (rng.random((M, N)) + 1j * rng.random((M, N))).astype(np.complex64)
    x = (rng.random(N) + 1j * rng.random(N)).astype(np.complex64)
    b = (A @ x).astype(np.complex64)
    for kw in (dict(iterations=200), dict(iterations=200, shuffleRows=True), dict(iterations=400, randomized=True)):
        xa = rls.solve_(rls.createLinearSolver(rls.Kaczmarz, A, **kw), b)
        assert np.linalg.norm(x - xa) / np.linalg.norm(x) < 0.1, kw
    for strategy in (rls.SystemMatrixBasedNormalization(), rls.MeasurementBasedNormalization()):
        S = rls.createLinearSolver(rls.Kaczmarz, A, iterations=200, randomized=True, reg=rls.L2Regularization(0.1), normalizeReg=strategy)
        xa = rls.solve_(S, b)
        assert np.linalg.norm(x - xa) / np.linalg.norm(x) < 0.3
    # Tikhonov matrix == column-scaled system with λ = 1
    lamv = rng.random(N).astype(np.float32) + np.float32(0.1)
    xm = rls.solve_(rls.createLinearSolver(rls.Kaczmarz, A, iterations=100, reg=[rls.L2Regularization(lamv)]), b)
    As = (A * (1 / np.sqrt(lamv))[None, :]).astype(np.complex64)
    xs = rls.solve_(rls.createLinearSolver(rls.Kaczmarz, As, iterations=100, reg=[rls.L2Regularization(np.float32(1))]), b) / np.sqrt(lamv)
    assert np.linalg.norm(xs - xm) / np.linalg.norm(xs) < 1e-5
    # a constant Tikhonov matrix == the scalar λ
    lam = np.float32(0.37)
    x1 = rls.solve_(rls.createLinearSolver(rls.Kaczmarz, A, iterations=100, reg=[rls.L2Regularization(lam)]), b)
    x2 = rls.solve_(rls.createLinearSolver(rls.Kaczmarz, A, iterations=100, reg=[rls.L2Regularization(np.full(N, lam))]), b)
    assert np.allclose(x1, x2, rtol=1e-3, atol=1e-5)
    xo = O.Kaczmarz(A, iterations=100, reg=[O.L2Regularization(lamv)]).solve(b)
    assert rel(xm, xo) < 1e-4


def test_update_matches_reference_kat(rls, ctx):
    """test/testKaczmarz.jl:6-33 pins kaczmarz_update!: b += β conj(A[k, :]).  Through the C ABI: from x = 0 with
    u = e_k, denom = 1 and eps_w = 0, a sweep over the single row k has α = 1 and leaves x = conj(A[k, :])."""
    import ctypes as C
    for dtype in DTYPES:
        A, _, _ = system(dtype, 16, 127, seed=370)
        S = rls.createLinearSolver(rls.Kaczmarz, A, iterations=1)
        k = 5
        rows = np.array([k], np.int64); den = np.array([1.0], np.float32)
        rls._capi.call("rls_kaczmarz_set_rows", S._handle, rows.ctypes.data_as(C.c_void_p), den.ctypes.data_as(C.c_void_p), 1)
        e = np.zeros(16, dtype); e[k] = 1
        bd = rls.B200Vector.from_numpy(e, ctx)
        rls._capi.call("rls_kaczmarz_init", S._handle, bd.handle, None, np.float32(0))
        rls._capi.call("rls_kaczmarz_sweep", S._handle)
        assert rel(S._vec("x").to_numpy(), np.conj(A[k, :])) < 1e-6


def test_errors(rls, ctx):
    A, _, b = system(np.float32, 64, 32, seed=380)
    Ac = rls.B200Matrix.from_numpy(A, ctx, layout="col")
    with pytest.raises(ValueError):
        rls.Kaczmarz(Ac)
    with pytest.raises(NotImplementedError):
        rls.Kaczmarz(A, greedy_randomized=True)
    with pytest.raises(TypeError):
        rls.Kaczmarz(A.astype(np.float64))
    import ctypes as C
    h = C.c_void_p()
    with pytest.raises(rls.RlsError):
        rls._capi.call("rls_kaczmarz_create", Ac.handle, 0, C.byref(h))
    Ar = rls.B200Matrix.from_numpy(A, ctx, layout="row")
    with pytest.raises(rls.RlsError):
        rls._capi.call("rls_kaczmarz_create", Ar.handle, 100, C.byref(h))
    S = rls.Kaczmarz(Ar)
    rows = np.array([3, 3], np.int64); den = np.ones(2, np.float32)
    with pytest.raises(rls.RlsError):
        rls._capi.call("rls_kaczmarz_set_rows", S._handle, rows.ctypes.data_as(C.c_void_p), den.ctypes.data_as(C.c_void_p), 2)
    with pytest.raises(rls.RlsError):
        rls._capi.call("rls_kaczmarz_sweep", rls.Kaczmarz(Ar)._handle)     # sweep before init
