"""Pins the CPU oracle against every known answer / acceptance property the reference
holds for the hot path: the solve! docstring KAT, the getting-started CGNR
example, the 256-point DFT compressed-sensing problem of test/testSolvers.jl, the prox
property tests of test/testProxMaps.jl and the multi-RHS test.  Runs on CPU."""
import numpy as np
import pytest

import oracle as O
from oracle.philox import philox4x32_10


def test_admm_docstring_kat():
    """src/RegularizedLeastSquares.jl:44-61 — the one exact known answer upstream."""
    A = np.array([[0.831658, 0.96717], [0.383056, 0.39043], [0.820692, 0.08118]])
    x = np.array([0.5932234523399985, 0.2697534345340015])
    b = A @ x
    for mode in ("gram", "lazy"):
        S = O.ADMM(A, reg=O.L1Regularization(0.0001), normal=mode)
        xa = O.solve_(S, b)
        assert np.allclose(xa, [0.5932171509222105, 0.26971370566079866], rtol=0, atol=1e-12)
        assert S.iteration == 10


def test_philox_known_answers():
    """Random123 KAT vectors for philox4x32_10."""
    z = np.zeros(1, np.uint32); f = np.full(1, 0xFFFFFFFF, np.uint32)
    assert [int(v[0]) for v in philox4x32_10(z, z, z, z, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(v[0]) for v in philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_cgnr_getting_started():
    """docs/src/literate/examples/getting_started.jl:29-38,49-51: 32x16 Float64, 32 iterations, rtol 1e-3."""
    rng = np.random.default_rng(0)
    A = rng.random((32, 16)); x = rng.random(16); b = A @ x
    xa = O.solve_(O.createLinearSolver(O.CGNR, A, iterations=32), b)
    assert np.allclose(xa, x, rtol=1e-3)
    xa = O.solve_(O.createLinearSolver(O.CGNR, A, iterations=32, reg=O.L2Regularization(0.0001)), b)
    assert np.allclose(xa, x, rtol=1e-2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_all_solvers_3x2(dtype):
    """test/testSolvers.jl:3-43: 3x2 systems, rtol 0.1."""
    rng = np.random.default_rng(12345)
    A = rng.random((3, 2)).astype(dtype); x = rng.random(2).astype(dtype)
    if np.dtype(dtype).kind == "c":
        A = (A + 1j * rng.random((3, 2))).astype(dtype); x = (x + 1j * rng.random(2)).astype(dtype)
    b = A @ x
    for solver in (O.CGNR, O.FISTA, O.POGM, O.OptISTA, O.ADMM):
        S = O.createLinearSolver(solver, A, iterations=200, normal="gram", bogus_keyword=1)
        xa = O.solve_(S, b)
        assert np.linalg.norm(xa - x) <= 0.1 * np.linalg.norm(x), solver.__name__
        # AHA-only interface (test/testSolvers.jl:45-65)
        S = O.createLinearSolver(solver, None, AHA=A.conj().T @ A, iterations=200)
        assert np.linalg.norm(O.solve_(S, A.conj().T @ b) - x) <= 0.1 * np.linalg.norm(x)


@pytest.mark.parametrize("elType", [np.float32, np.float64])
def test_convex_dft_compressed_sensing(elType):
    """test/testSolvers.jl:67-171.  (Julia's Random.seed!(12345) stream cannot be reproduced
    here; the instance is drawn from NumPy's generator instead.)"""
    rng = np.random.default_rng(1)
    N = 256
    F = (np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(N)) / N) / np.sqrt(N))
    x = np.zeros(N)
    for _ in range(3):
        x[rng.integers(0, N)] = rng.random()
    b = F @ x
    idx = np.sort(np.unique(rng.integers(0, N, N // 2)))
    ctype = np.complex128      # upstream F and b are ComplexF64 for every elType; elType only types λ
    b = b[idx].astype(ctype); F = F[idx, :].astype(ctype)
    for solver in (O.POGM, O.OptISTA, O.FISTA, O.ADMM):
        reg = O.L1Regularization(elType(1e-3))
        S = O.createLinearSolver(solver, F, reg=reg, iterations=200, normalizeReg=O.NoNormalization())
        xa = O.solve_(S, b)
        assert np.linalg.norm(x - xa) <= 0.1 * np.linalg.norm(x), solver.__name__
        if solver in (O.POGM, O.FISTA):
            S = O.createLinearSolver(solver, F, reg=reg, iterations=200, restart="gradient")
            assert np.linalg.norm(x - O.solve_(S, b)) <= 0.1 * np.linalg.norm(x)
        # invariance to the scale of F with MeasurementBasedNormalization (:108-124)
        reg2 = O.L1Regularization(elType(reg.lam * len(b) / np.sum(np.abs(b))))
        S = O.createLinearSolver(solver, (F * 1e3).astype(ctype), reg=reg2, iterations=200,
                                 normalizeReg=O.MeasurementBasedNormalization())
        xa = O.solve_(S, b) * 1e3
        assert np.linalg.norm(x - xa) <= 0.1 * np.linalg.norm(x), solver.__name__ + " rescaled"
    # :127-171 — rho = 1e6 and 1e-6 with :balance, rho = 1e-6 with :PnP (it only ever increases rho)
    for rho, vary in ((1e6, "balance"), (1e-6, "balance"), (1e-6, "PnP")):
        S = O.createLinearSolver(O.ADMM, F, reg=O.L1Regularization(elType(1e-3)), iterations=200, rho=rho, vary_rho=vary)
        assert np.linalg.norm(x - O.solve_(S, b)) <= 0.1 * np.linalg.norm(x), (rho, vary)


def test_prox_l2_l1():
    """test/testProxMaps.jl:2-39."""
    rng = np.random.default_rng(1234)
    N = 256
    x = np.zeros(N)
    x[rng.integers(0, N, 5)] = rng.random(5)
    lam = 0.01
    x_l2 = O.prox_(O.L2Regularization(lam), x.copy())
    assert np.linalg.norm(x_l2 - x / (1 + 2 * lam)) / np.linalg.norm(x / (1 + 2 * lam)) < 1e-3
    assert 0.5 * np.linalg.norm(x - x_l2) ** 2 + O.reg_norm(O.L2Regularization(lam), x_l2) <= O.reg_norm(O.L2Regularization(lam), x)
    sig = 0.03
    x = np.zeros(N)
    x[rng.integers(0, N, 5)] = (1 - 2 * sig) * rng.random(5) + 2 * sig
    s = np.sum(np.abs(x)) / N * sig
    noisy = x + s / np.sqrt(2.0) * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    x_l1 = O.prox_(O.L1Regularization(2 * s), noisy.copy())
    assert np.linalg.norm(x - x_l1) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - x_l1) / np.linalg.norm(x) < 0.1
    r = O.L1Regularization(2 * s)
    assert 0.5 * np.linalg.norm(noisy - x_l1) ** 2 + O.reg_norm(r, x_l1) <= O.reg_norm(r, noisy)


def test_prox_l21_ragged_groups_by_hand():
    """ProxL21.jl:30-35 with length(x) = 7, slices = 3: L = 2, groups x[1:2:end] = {1,3,5,7} and x[2:2:end] = {2,4,6}
    (1-based) — every element is rescaled by its group's factor, including the seventh."""
    x = np.array([3.0, 1.0, 4.0, 2.0, 12.0, 2.0, 0.0])
    lam = 1.0
    g0, g1 = np.linalg.norm(x[0::2]), np.linalg.norm(x[1::2])      # 13, 3
    want = x.copy()
    want[0::2] *= (g0 - lam) / g0
    want[1::2] *= (g1 - lam) / g1
    got = O.prox_(O.L21Regularization(lam, slices=3), x.copy())
    assert np.allclose(got, want, rtol=1e-15)
    assert np.isclose(O.reg_norm(O.L21Regularization(lam, slices=3), x), lam * (g0 + g1))
    with pytest.raises(ValueError):
        O.prox_(O.L21Regularization(lam, slices=8), x.copy())


def test_prox_l21():
    """test/testProxMaps.jl:44-72 (N=256, 8 slices, last 2 noisy)."""
    rng = np.random.default_rng(1234)
    N, S, noisyS, sig = 256, 8, 2, 0.05
    X = np.zeros((N, S), np.complex128)
    for _ in range(5):
        X[rng.integers(0, N), :] = (1 - 2 * sig) * rng.random(S) + 2 * sig
    x = X.ravel(order="F")
    s = np.sum(np.abs(x)) / x.size * sig
    noisy = x.copy()
    noisy[(S - noisyS) * N:] += s / np.sqrt(2.0) * (rng.standard_normal(N * noisyS) + 1j * rng.standard_normal(N * noisyS))
    x_l1 = O.prox_(O.L1Regularization(2 * s), noisy.copy())
    r = O.L21Regularization(2 * s, slices=S)
    x_l21 = O.prox_(r, noisy.copy())
    assert np.linalg.norm(x - x_l21) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - x_l21) <= np.linalg.norm(x - x_l1)
    assert np.linalg.norm(x - x_l21) / np.linalg.norm(x) < 0.05
    obj = lambda v: 0.5 * np.linalg.norm(noisy - v) ** 2 + O.reg_norm(r, v)
    assert obj(x_l21) <= O.reg_norm(r, noisy)
    assert obj(x_l21) <= obj(x_l1)


def test_prox_tv_2d_and_directional():
    """test/testProxMaps.jl:75-136."""
    rng = np.random.default_rng(1234)
    N, sig = 64, 0.05
    X = np.zeros((N, N), np.complex128)
    for _ in range(5):
        i, j = rng.integers(0, N, 2)
        X[i:, j:] += rng.standard_normal()
    x = X.ravel(order="F")
    s = np.sum(np.abs(x)) / x.size * sig
    noisy = x + s / np.sqrt(2.0) * (rng.standard_normal(N * N) + 1j * rng.standard_normal(N * N))
    x_l1 = O.prox_(O.L1Regularization(2 * s), noisy.copy())
    r = O.TVRegularization(2 * s, shape=(N, N))
    x_tv = O.prox_(r, noisy.copy())
    assert np.linalg.norm(x - x_tv) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - x_tv) <= np.linalg.norm(x - x_l1)
    obj = lambda v: 0.5 * np.linalg.norm(noisy - v) ** 2 + O.reg_norm(r, v)
    assert obj(x_tv) <= O.reg_norm(r, noisy)
    assert obj(x_tv) <= obj(x_l1)
    # directional TV == column-wise 1-D TV (FGP both ways) to 1e-8
    a = O.prox_(O.TVRegularization(2 * s, shape=(N, N), dims=1), noisy.copy()).reshape(N, N, order="F")
    cols = noisy.reshape(N, N, order="F").copy()
    for j in range(N):
        cols[:, j] = O.prox_(O.TVRegularization(2 * s, shape=(N,), dims=1), cols[:, j].copy())
    assert np.linalg.norm(a - cols) / np.linalg.norm(x) < 1e-8


def test_prox_positive_real():
    """test/testProxMaps.jl:139-151."""
    rng = np.random.default_rng(1234)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    xp = O.prox_(O.PositiveRegularization(), x.copy())
    assert np.array_equal(xp, np.maximum(x.real, 0) + 0j)
    xr = O.prox_(O.RealRegularization(), x.copy())
    assert np.array_equal(xr, x.real + 0j)
    xf = rng.standard_normal(16).astype(np.float32)
    assert np.array_equal(O.prox_(O.PositiveRegularization(), xf.copy()), np.maximum(xf, 0))


def test_multi_rhs_sequential_equals_columns():
    """test/testMultiThreading.jl:1-20."""
    rng = np.random.default_rng(3)
    A = (rng.random((3, 2)) + 1j * rng.random((3, 2))).astype(np.complex64)
    X = (rng.random((2, 4)) + 1j * rng.random((2, 4))).astype(np.complex64)
    B = A @ X
    for solver in (O.CGNR, O.FISTA, O.POGM, O.OptISTA, O.ADMM):
        S = O.createLinearSolver(solver, A, iterations=100, normal="gram")
        Xb = O.solve_(S, B)
        assert np.linalg.norm(Xb - X) <= 0.1 * np.linalg.norm(X)
        xv = O.solve_(S, B[:, 0].copy())
        assert np.allclose(xv, Xb[:, 0])


def test_callbacks_fire_iterations_plus_one():
    """test/testCallbacks.jl:1-57 idiom: relTol = 0, iterations+1 invocations, last == solution."""
    rng = np.random.default_rng(5)
    A = rng.random((16, 8)).astype(np.float32); b = A @ rng.random(8).astype(np.float32)
    S = O.FISTA(A, iterations=12, relTol=0.0, rho=np.float32(0.05))
    seen = []
    x = S.solve(b, callbacks=lambda s, it: seen.append((it, s.x.copy())))
    assert [k for k, _ in seen] == list(range(13))
    assert np.array_equal(seen[-1][1], x)


def test_gram_and_lazy_forms_agree():
    rng = np.random.default_rng(7)
    A = (rng.standard_normal((40, 24)) + 1j * rng.standard_normal((40, 24))).astype(np.complex64)
    b = (A @ rng.standard_normal(24)).astype(np.complex64)
    for solver in (O.FISTA, O.CGNR):
        kw = dict(iterations=15)
        if solver is O.FISTA:
            kw.update(rho=np.float32(0.005), reg=O.L1Regularization(np.float32(1e-3)), relTol=0.0)
        else:
            kw.update(iterations=6)      # Float32 CG steering scalars amplify rounding once converged
        a = solver(A, normal="gram", **kw).solve(b)
        c = solver(A, normal="lazy", **kw).solve(b)
        assert np.linalg.norm(a - c) <= 2e-4 * np.linalg.norm(c)


def test_float32_scalar_recurrences_stay_float32():
    A = np.eye(4, dtype=np.float32)
    S = O.FISTA(A, iterations=5, rho=0.5, relTol=0.0)
    S.solve(np.ones(4, np.float32))
    assert type(S.theta) is np.float32 and type(S.rel_res_norm) is np.float32
    S = O.OptISTA(A, iterations=5, rho=0.5, relTol=0.0)
    S.solve(np.ones(4, np.float32))
    assert type(S.gamma) is np.float32 and type(S.thetan) is np.float32
    S = O.CGNR(A.astype(np.complex64), iterations=3)
    S.solve(np.ones(4, np.complex64))
    assert type(S.alpha) is np.complex64


# ---------------------------------------------------------------- SplitBregman
def _dft_cs_problem(seed=1):
    rng = np.random.default_rng(seed)
    N = 256
    F = (np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(N)) / N) / np.sqrt(N))
    x = np.zeros(N)
    for _ in range(3):
        x[rng.integers(0, N)] = rng.random()
    b = F @ x
    idx = np.sort(np.unique(rng.integers(0, N, N // 2)))
    return F[idx, :].astype(np.complex128), b[idx].astype(np.complex128), x


@pytest.mark.parametrize("elType", [np.float32, np.float64])
def test_split_bregman_dft_compressed_sensing(elType):
    """test/testSolvers.jl:174-201: SplitBregman, L1Regularization(2e-3), iterations=5, iterationsInner=40, rho=1.0,
    plain and with MeasurementBasedNormalization on a rescaled system, rtol 0.1."""
    F, b, x = _dft_cs_problem()
    reg = O.L1Regularization(elType(2e-3))
    S = O.createLinearSolver(O.SplitBregman, F, reg=reg, iterations=5, iterationsInner=40, rho=1.0,
                             normalizeReg=O.NoNormalization())
    xa = O.solve_(S, b)
    assert np.linalg.norm(x - xa) <= 0.1 * np.linalg.norm(x)
    assert S.iter_cnt <= 6 and S.total_iterations <= 5 * 40
    reg2 = O.L1Regularization(elType(reg.lam * len(b) / np.sum(np.abs(b))))
    S = O.createLinearSolver(O.SplitBregman, F, reg=reg2, iterations=5, iterationsInner=40, rho=1.0,
                             normalizeReg=O.MeasurementBasedNormalization())
    xa = O.solve_(S, b)
    assert np.linalg.norm(x - xa) <= 0.1 * np.linalg.norm(x)


def test_split_bregman_one_outer_iteration_is_admm():
    """SplitBregman.jl docstring: `iterations = 1` is the unconstrained problem ADMM solves; the inner iteration
    is ADMM's with the threshold λ/ρ instead of ADMM's λ/(2ρ) (ADMM.jl:262), so SplitBregman(λ) ≡ ADMM(2λ)
    iterate by iterate as long as neither has stopped."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((40, 24)); xt = np.zeros(24); xt[[3, 11, 17]] = [1.0, -0.5, 0.7]
    b = A @ xt
    lam, rho, inner = 1e-2, 0.5, 15
    S = O.SplitBregman(A, reg=O.L1Regularization(lam), rho=rho, iterations=1, iterationsInner=inner, absTol=0.0, relTol=0.0)
    R = O.ADMM(A, reg=O.L1Regularization(2 * lam), rho=rho, iterations=inner, absTol=0.0, relTol=0.0)
    S.init(b); R.init(b)
    for k in range(inner):
        assert S.iterate() and R.iterate()
        if k < inner - 1:        # after the last inner iteration SplitBregman performs its Bregman update of β_y only
            assert np.allclose(S.x, R.x, rtol=1e-12, atol=1e-14), k
            assert np.allclose(S.z[0], R.z[0], rtol=1e-12, atol=1e-14)
    assert np.allclose(S.x, R.x, rtol=1e-12, atol=1e-14)
    assert not S.iterate()       # iteration == 1 and iter_cnt (2) > iterations (1)   (:289)
    assert S.iter_cnt == 2 and S.total_iterations == inner


def test_split_bregman_constraint_is_enforced():
    """More outer (Bregman) iterations drive ||A x - b|| down at fixed λ (Goldstein & Osher eq. 4.7)."""
    rng = np.random.default_rng(6)
    A = rng.standard_normal((60, 30)); xt = np.zeros(30); xt[[2, 9, 21]] = [1.0, 0.8, -0.6]
    b = A @ xt
    res = []
    for outer in (1, 4):
        S = O.SplitBregman(A, reg=O.L1Regularization(0.5), rho=1.0, iterations=outer, iterationsInner=20, absTol=0.0, relTol=0.0)
        xa = O.solve_(S, b)
        res.append(np.linalg.norm(A @ xa - b))
    assert res[1] < 0.5 * res[0]


# ---- Kaczmarz (src/Kaczmarz.jl): the properties test/testKaczmarz.jl pins ---------------------------------------
def _kaczmarz_system(rng, M=12, N=8):
    A = rng.random((M, N)) + 1j * rng.random((M, N))
    x = rng.random(N) + 1j * rng.random(N)
    return A, x, A @ x


def test_kaczmarz_parameters():
    """test/testKaczmarz.jl:94-125: plain, shuffled, randomized, normalised"""
    rng = np.random.default_rng(12345)
    A, x, b = _kaczmarz_system(rng)
    for kw in (dict(iterations=200), dict(iterations=200, shuffleRows=True), dict(iterations=2000, randomized=True)):
        xa = O.Kaczmarz(A, **kw).solve(b)
        assert np.linalg.norm(x - xa) / np.linalg.norm(x) < 0.1
    for strategy in (O.SystemMatrixBasedNormalization(), O.MeasurementBasedNormalization()):
        xa = O.Kaczmarz(A, iterations=200, randomized=True, reg=O.L2Regularization(0.1), normalizeReg=strategy).solve(b)
        assert np.linalg.norm(x - xa) / np.linalg.norm(x) < 0.3
    with pytest.raises(ValueError):
        O.Kaczmarz(A, reg=[O.L1Regularization(1e-3), O.L21Regularization(1e-3)])


def test_kaczmarz_tikhonov_matrix():
    """test/testKaczmarz.jl:37-70"""
    rng = np.random.default_rng(12345)
    A, x, b = _kaczmarz_system(rng)
    N = A.shape[1]
    lamv = rng.random(N)
    xm = O.Kaczmarz(A, iterations=100, reg=[O.L2Regularization(lamv)]).solve(b)
    xs = O.Kaczmarz(A * (1 / np.sqrt(lamv))[None, :], iterations=100, reg=[O.L2Regularization(1.0)]).solve(b) / np.sqrt(lamv)
    assert np.linalg.norm(xs - xm) / np.linalg.norm(xs) < 0.1
    lam = float(rng.random())
    x1 = O.Kaczmarz(A, iterations=100, reg=[O.L2Regularization(lam)]).solve(b)
    x2 = O.Kaczmarz(A, iterations=100, reg=[O.L2Regularization(np.full(N, lam))]).solve(b)
    assert np.allclose(x1, x2)


def test_kaczmarz_fixed_point_is_the_tikhonov_solution():
    """the iteration on the extended system [A  sqrt(λ) I] converges to argmin ‖Ax − b‖² + λ‖x‖² (Kaczmarz.jl:305-310)"""
    rng = np.random.default_rng(7)
    A = rng.standard_normal((30, 10)); b = rng.standard_normal(30); lam = 0.5
    xa = O.Kaczmarz(A, iterations=3000, reg=O.L2Regularization(lam)).solve(b)
    xr = np.linalg.solve(A.T @ A + lam * np.eye(10), A.T @ b)
    assert np.linalg.norm(xa - xr) / np.linalg.norm(xr) < 1e-8


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_kaczmarz_block_gram_form_is_the_same_recurrence(dtype):
    """the evaluation csrc/rls_kaczmarz.cu uses — per block of R rows: t = A_blk x, tau_j = t_j + sum_{k<j} alpha_k G[j,k]
    with G = A_blk A_blk^H, x += A_blk^H alpha — reproduces the row-by-row iterates to rounding in single precision"""
    rng = np.random.default_rng(3)
    m, n, R = 200, 96, 64
    A = (rng.standard_normal((m, n)) / np.sqrt(m)).astype(np.float32)
    if dtype == np.complex64:
        A = (A + 1j * rng.standard_normal((m, n)).astype(np.float32) / np.float32(np.sqrt(m))).astype(dtype)
    A = A.astype(dtype)
    b = (A @ rng.standard_normal(n).astype(dtype)).astype(dtype)
    lam = np.float32(1e-2)
    S = O.Kaczmarz(A, reg=O.L2Regularization(lam), iterations=4); S.init(b)
    x = np.zeros(n, dtype); vl = np.zeros(m, dtype); ew = np.float32(np.sqrt(lam))
    for _ in range(4):
        S.iterate()
        for r0 in range(0, m, R):
            Ab = A[r0:r0 + R]; G = (Ab @ Ab.conj().T).astype(dtype); t = (Ab @ x).astype(dtype)
            c = np.zeros(len(Ab), dtype); al = np.zeros(len(Ab), dtype)
            for j in range(len(Ab)):
                al[j] = S.denom[r0 + j] * (b[r0 + j] - (t[j] + c[j]) - ew * vl[r0 + j])
                c[j + 1:] += al[j] * G[j + 1:, j]
                vl[r0 + j] += al[j] * ew
            x += (Ab.conj().T @ al).astype(dtype)
        assert np.linalg.norm(x - S.x) / np.linalg.norm(S.x) < 1e-5


@pytest.mark.parametrize("corr", [0.0, 0.99, 0.9999])
def test_kaczmarz_inverted_diagonal_blocks_stay_accurate_on_correlated_rows(corr):
    """the sweep kernel solves the block recurrence (D^-1 + strictlower(G)) alpha = r by blocked forward substitution with
    the 32x32 diagonal blocks inverted once (in double, stored in single).  Nearly parallel rows make G ill-conditioned
    (cond ~ 2e5 at row correlation 0.9999); the iterates must still follow the row-by-row loop."""
    rng = np.random.default_rng(11)
    m, n, R, lam = 256, 512, 128, np.float32(1e-2)
    base = rng.standard_normal(n).astype(np.float32)
    A = ((np.sqrt(1 - corr ** 2) * rng.standard_normal((m, n)).astype(np.float32) + corr * base[None, :]) / np.float32(np.sqrt(m))).astype(np.float32)
    b = (A @ rng.standard_normal(n).astype(np.float32)).astype(np.float32)
    S = O.Kaczmarz(A, reg=O.L2Regularization(lam), iterations=3); S.init(b)
    ew = np.float32(np.sqrt(lam))
    x = np.zeros(n, np.float32); vl = np.zeros(m, np.float32)
    blocks = []
    for r0 in range(0, m, R):
        G = (A[r0:r0 + R] @ A[r0:r0 + R].T).astype(np.float32)
        inv = [np.linalg.inv(np.tril(G[j:j + 32, j:j + 32].astype(np.float64), -1) +
                             np.diag(1.0 / S.denom[r0 + j:r0 + j + 32].astype(np.float64))).astype(np.float32) for j in range(0, R, 32)]
        blocks.append((G, inv))
    for _ in range(3):
        S.iterate()
        for bi, r0 in enumerate(range(0, m, R)):
            G, inv = blocks[bi]
            Ab = A[r0:r0 + R]
            r = (b[r0:r0 + R] - (Ab @ x).astype(np.float32) - ew * vl[r0:r0 + R]).astype(np.float32)
            c = np.zeros(R, np.float32); al = np.zeros(R, np.float32)
            for jb, j0 in enumerate(range(0, R, 32)):
                al[j0:j0 + 32] = inv[jb] @ (r[j0:j0 + 32] - c[j0:j0 + 32])
                c[j0 + 32:] += G[j0 + 32:, j0:j0 + 32] @ al[j0:j0 + 32]
            vl[r0:r0 + R] += al * ew
            x += Ab.T @ al
        assert np.linalg.norm(x - S.x) / np.linalg.norm(S.x) < 1e-5
