"""Per-iterate parity of the CUDA solvers against the oracle (rel-L2 <= 1e-5, Float32 /
ComplexF32 — the tolerance BASELINE.json's north_star states), identical iteration counts
and stopping decisions, through the C ABI.  No tolerance above 1e-5 is used against the Float32 oracle: where an iterate
can exceed it because of the oracle's own Float32 BLAS rounding (CG-steered solvers, long accelerated runs), the test
carries a Float64 run of the oracle and the iterate must be as close to it as the Float32 oracle is
(util.stepwise_vs_fp64 / assert_close_or_fp64)."""
import numpy as np
import pytest

import oracle as O
from util import rel, rand_matrix, rand_vector, sparse_truth, up64, to64, stepwise_vs_fp64, assert_close_or_fp64

pytestmark = pytest.mark.gpu

TOL = 1e-5
DTYPES = [np.float32, np.complex64]


def problem(dtype, m, n, seed=100, noise=1e-3):
    A, _ = rand_matrix(dtype, m, n, seed)
    xt = sparse_truth(dtype, n, seed + 1)
    b = (A @ xt + noise * rand_vector(dtype, m, seed + 2)).astype(dtype)
    return A, xt, b


def rho_for(A):
    return np.float32(0.95 / np.linalg.norm(A.astype(np.complex128), 2) ** 2)


def stepwise(S, R, b, iters, what="x", tol=TOL):
    """drive both through init!/iterate and compare after every iteration"""
    S.init_(b); R.init(b)
    worst = 0.0
    for k in range(iters + 2):
        a1, a2 = S.iterate(), R.iterate()
        assert a1 == a2, f"stopping decision differs at iteration {k}: gpu={a1} oracle={a2}"
        if not a1:
            break
        e = rel(getattr(S, what), getattr(R, what))
        worst = max(worst, e)
        assert e < tol, f"iterate {k + 1}: rel-L2 {e:.3e}"
    assert S.iteration == R.iteration
    return worst


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("solver", ["FISTA", "POGM", "OptISTA"])
@pytest.mark.parametrize("restart", ["none", "gradient"])
def test_proxgrad_l1_per_iterate(rls, ctx, dtype, solver, restart):
    if solver == "OptISTA" and restart == "gradient":
        pytest.skip("OptISTA has no restart keyword")
    A, xt, b = problem(dtype, 384, 1024)
    rho = rho_for(A)
    lam = np.float32(2e-2)
    kw = dict(iterations=40, rho=rho, relTol=0.0)
    if solver != "OptISTA":
        kw["restart"] = restart
    S = rls.createLinearSolver(getattr(rls, solver), A, reg=rls.L1Regularization(lam), normal="twopass", **kw)
    R = O.createLinearSolver(getattr(O, solver), A, reg=O.L1Regularization(lam), **kw)
    stepwise(S, R, b, 40)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("form", ["twopass", "onepass", "gram"])
def test_fista_forms_whole_solve(rls, ctx, dtype, form):
    A, xt, b = problem(dtype, 1024, 4096)
    rho = rho_for(A)
    lam = np.float32(1e-2)
    S = rls.FISTA(A, reg=rls.L1Regularization(lam), iterations=100, rho=rho, relTol=0.0, normal=form)
    R = O.FISTA(A, reg=O.L1Regularization(lam), iterations=100, rho=rho, relTol=0.0)
    x = rls.solve_(S, b)
    xr = R.solve(b)
    assert S.iteration == R.iteration == 100
    # 100 thresholded iterations amplify the summation-order rounding of x0 = A'b (1e-7) about a hundredfold
    assert_close_or_fp64(x, xr, lambda: O.FISTA(up64(A), reg=O.L1Regularization(float(lam)), iterations=100, rho=float(rho),
                                                  relTol=0.0).solve(up64(b)), what="100 iterations")
    assert abs(S.state.rel_res_norm - R.rel_res_norm) <= 1e-4 * abs(R.rel_res_norm)
    # whole-solve fast path == init!/iterate loop with a callback
    trace = []
    x2 = rls.solve_(S, b, callbacks=lambda s, it: trace.append(it))
    assert trace == list(range(101))          # iterations+1 calls (test/testCallbacks.jl:13)
    assert np.array_equal(x, x2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("lam", [np.float32(0), np.float32(1e-3), 1e-3])
def test_cgnr_per_iterate(rls, ctx, dtype, lam):
    A, xt, b = problem(dtype, 512, 256)
    regs = lambda M: M.L2Regularization(lam)
    S = rls.CGNR(A, reg=regs(rls), iterations=30, relTol=0.0, normal="twopass")
    R = O.CGNR(A, reg=regs(O), iterations=30, relTol=0.0)
    R64 = O.CGNR(up64(A), reg=to64(regs(O)), iterations=30, relTol=0.0)
    stepwise_vs_fp64(S, R, R64, b, 30)       # CG steering scalars amplify rounding (SURVEY 7 hard part 2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_cgnr_c1_shape_and_stop(rls, ctx, dtype):
    """BASELINE configs[0] shape family: CGNR + L2 on a wide system with uniform entries
    (test/testSolvers.jl:25-27).  In Float32 the raw U[0,1) system is numerically chaotic — the
    oracle itself departs from its own Float64 run by 8 % at iteration 3 and by 400 % at
    iteration 30 — so per-iterate parity is asserted on the centred (well-conditioned) system,
    and on the raw system only for the first iterates plus the data residual."""
    from oracle.philox import philox_matrix, philox_vector, UNIFORM01
    m, n = 256, 1024
    A = philox_matrix(dtype, m, n, 12345, UNIFORM01, 1.0)
    xt = philox_vector(dtype, n, 12345, 5, UNIFORM01)
    lam = np.float32(1e-3)
    # (a) raw U[0,1): first two iterates and the final data residual
    b = (A @ xt).astype(dtype)
    S = rls.CGNR(A, reg=rls.L2Regularization(lam), iterations=50, relTol=0.0)
    R = O.CGNR(A, reg=O.L2Regularization(lam), iterations=50, relTol=0.0)
    R64 = O.CGNR(up64(A), reg=O.L2Regularization(float(lam)), iterations=50, relTol=0.0)
    # in this chaotic regime the distance of EITHER Float32 run to the Float64 recurrence grows by an order of magnitude
    # every few iterations: a factor 4 between the two is less than one iteration's growth
    stepwise_vs_fp64(S, R, R64, b, 50, slack=4.0)
    x = rls.solve_(S, b)
    assert S.iteration == R.iteration
    assert rel(A @ x, b) <= max(1e-3, 2 * rel(A @ R.x, b))
    # (b) centred entries: stopping decision and solution
    Ac = (A - (0.5 + 0.5j if np.dtype(dtype).kind == "c" else 0.5)).astype(dtype)
    b = (Ac @ xt).astype(dtype)
    S = rls.CGNR(Ac, reg=rls.L2Regularization(lam), iterations=50, relTol=1e-3)
    R = O.CGNR(Ac, reg=O.L2Regularization(lam), iterations=50, relTol=1e-3)
    x = rls.solve_(S, b); xr = R.solve(b)
    assert 0 < R.iteration < 50
    assert S.iteration == R.iteration, "identical iteration counts / stopping decisions"
    assert_close_or_fp64(x, xr, lambda: O.CGNR(up64(Ac), reg=O.L2Regularization(float(lam)), iterations=R.iteration,
                                               relTol=0.0).solve(up64(b)), what="centred system")
    S = rls.CGNR(Ac, reg=rls.L2Regularization(lam), iterations=8, relTol=0.0)
    R = O.CGNR(Ac, reg=O.L2Regularization(lam), iterations=8, relTol=0.0)
    R64 = O.CGNR(up64(Ac), reg=O.L2Regularization(float(lam)), iterations=8, relTol=0.0)
    stepwise_vs_fp64(S, R, R64, b, 8)
    # (c) iteration cap min(iterations, n) (CGNR.jl:185) and projections at termination only
    S = rls.CGNR(A[:, :8].copy(), reg=[rls.L2Regularization(lam), rls.PositiveRegularization()], iterations=50, relTol=0.0)
    R = O.CGNR(A[:, :8].copy(), reg=[O.L2Regularization(lam), O.PositiveRegularization()], iterations=50, relTol=0.0)
    x = rls.solve_(S, b); xr = R.solve(b)
    assert S.iteration == R.iteration == 8
    if np.dtype(dtype).kind == "c":
        assert np.all(x.imag == 0)
    assert np.all(x.real >= 0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_fista_reltol_stop_and_projection(rls, ctx, dtype):
    A, _ = rand_matrix(dtype, 300, 200, 100)
    xt = np.abs(sparse_truth(dtype, 200, 101)).astype(dtype)     # real, non-negative truth: the projections can fit it
    b = (A @ xt).astype(dtype)
    rho = rho_for(A)
    lam = np.float32(1e-4)
    for proj in ("Positive", "Real"):
        S = rls.FISTA(A, reg=[rls.L1Regularization(lam), getattr(rls, proj + "Regularization")()], iterations=300,
                      rho=rho, relTol=1e-3)
        R = O.FISTA(A, reg=[O.L1Regularization(lam), getattr(O, proj + "Regularization")()], iterations=300,
                    rho=rho, relTol=1e-3)
        x = rls.solve_(S, b); xr = R.solve(b)
        assert 0 < R.iteration < 300, "the tolerance must trigger before the cap for this test to mean anything"
        assert S.iteration == R.iteration
        assert rel(x, xr) < TOL


@pytest.mark.parametrize("solver", ["FISTA", "POGM", "OptISTA"])
@pytest.mark.parametrize("regname", ["L2", "L21", "TV"])
def test_proxgrad_other_regs(rls, ctx, solver, regname):
    dtype = np.complex64
    A, xt, b = problem(dtype, 256, 32 * 24)
    rho = rho_for(A)
    mk = {"L2": lambda M: M.L2Regularization(np.float32(5e-2)),
          "L21": lambda M: M.L21Regularization(np.float32(5e-3), slices=8),
          "TV": lambda M: M.TVRegularization(np.float32(5e-3), shape=(32, 24))}[regname]
    S = getattr(rls, solver)(A, reg=mk(rls), iterations=25, rho=rho, relTol=0.0, normal="twopass")
    R = getattr(O, solver)(A, reg=mk(O), iterations=25, rho=rho, relTol=0.0)
    R64 = getattr(O, solver)(up64(A), reg=to64(mk(O)), iterations=25, rho=float(rho), relTol=0.0)
    stepwise_vs_fp64(S, R, R64, b, 25)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("variant", ["l1", "l1_vary_balance", "l1_vary_pnp", "tv_identity", "l1_gradient", "two_terms"])
def test_admm(rls, ctx, dtype, variant):
    shape = (24, 20)
    n = shape[0] * shape[1]
    A, xt, b = problem(dtype, 320, n)
    kw = dict(iterations=15, iterationsCG=10, rho=0.1, absTol=0.0, relTol=0.0)
    def mk(M):
        if variant == "l1":
            return dict(reg=M.L1Regularization(np.float32(1e-2)))
        if variant == "l1_vary_balance":
            return dict(reg=M.L1Regularization(np.float32(1e-2)), vary_rho="balance")
        if variant == "l1_vary_pnp":
            return dict(reg=M.L1Regularization(np.float32(1e-2)), vary_rho="PnP")
        if variant == "tv_identity":
            return dict(reg=M.TVRegularization(np.float32(1e-2), shape=shape))
        if variant == "l1_gradient":
            return dict(reg=M.L1Regularization(np.float32(1e-2)), regTrafo=M.GradientOp(dtype, shape))
        return dict(reg=[M.L1Regularization(np.float32(1e-2)), M.L1Regularization(1e-3)],
                    regTrafo=[None, M.GradientOp(dtype, shape)], rho=[0.1, 0.2])
    k1, k2 = dict(kw), dict(kw)
    k1.update(mk(rls)); k2.update(mk(O))
    S = rls.ADMM(A, normal="twopass", **k1)
    R = O.ADMM(A, **k2)
    R64 = O.ADMM(up64(A), **to64(k2))

    def each(k):
        assert S._scalars.cg_iterations_last == R.cg_iters[-1], f"inner CG count differs at outer {k}"
    stepwise_vs_fp64(S, R, R64, b, 15, each=each)
    assert S.iteration == 15
    conv = S.convergence()
    assert np.allclose(conv["primal"], R.rk, rtol=2e-3)
    assert np.allclose(conv["dual"], R.sk, rtol=2e-3)


def test_admm_docstring_kat_float32(rls, ctx):
    """src/RegularizedLeastSquares.jl:44-61, Float32 edition of the one true KAT."""
    A = np.array([[0.831658, 0.96717], [0.383056, 0.39043], [0.820692, 0.08118]])
    x = np.array([0.5932234523399985, 0.2697534345340015])
    b = A @ x
    S = rls.ADMM(A.astype(np.float32), reg=rls.L1Regularization(0.0001))
    xg = rls.solve_(S, b.astype(np.float32))
    assert np.allclose(xg, [0.5932171509222105, 0.26971370566079866], rtol=2e-4)
    R = O.ADMM(A.astype(np.float32), reg=O.L1Regularization(0.0001))
    assert rel(xg, R.solve(b.astype(np.float32))) < 1e-5


@pytest.mark.parametrize("solver", ["FISTA", "CGNR", "ADMM", "POGM", "OptISTA", "SplitBregman"])
@pytest.mark.parametrize("tensor_cores,layout", [(True, "row"), (True, "col"), (False, "row")])
def test_multi_rhs(rls, ctx, solver, tensor_cores, layout, monkeypatch):
    """test/testMultiThreading.jl: batched == sequential, and a vector solve still works afterwards.
    On a row-major A the K applies of a batched iteration (for ADMM / SplitBregman: of every segment — AHA x and each
    inner CG step, with the per-column device gates) are two tcgen05 GEMMs; the columns then agree with the sequential
    solves to the per-iterate parity bound.  With RLS_BATCH_TENSOR_CORES=0 or a column-major A (K single applies) they are
    bit-identical."""
    monkeypatch.setenv("RLS_BATCH_TENSOR_CORES", "1" if tensor_cores else "0")
    monkeypatch.setenv("RLS_BATCH_MIN_K", "2")      # the GEMM path normally starts at 8 columns
    dtype = np.complex64
    A, _, _ = problem(dtype, 201, 96)      # odd m: the columns of B are not 16-byte multiples
    X = np.stack([sparse_truth(dtype, 96, 300 + k, every=7) for k in range(5)], axis=1)
    B = (A @ X).astype(dtype)
    kw = dict(iterations=30)
    okw = dict(iterations=30)
    if solver in ("FISTA", "POGM", "OptISTA"):
        kw.update(rho=rho_for(A), reg=rls.L1Regularization(np.float32(1e-4)))
        okw.update(rho=rho_for(A), reg=O.L1Regularization(np.float32(1e-4)))
    if solver == "SplitBregman":
        kw.update(iterations=3, iterationsInner=5)
        okw.update(iterations=3, iterationsInner=5)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout=layout)
    S = rls.createLinearSolver(getattr(rls, solver), Ad, **kw)
    Xb = rls.solve_(S, B)
    Xs = np.stack([rls.solve_(S, B[:, k].copy()) for k in range(5)], axis=1)
    if tensor_cores and layout == "row":
        assert not np.array_equal(Xb, Xs), "the batched solve was meant to take the tensor-core GEMM path"
        for k in range(5):
            e = rel(Xb[:, k], Xs[:, k])
            if e < TOL:
                continue
            # CG-steered solvers amplify the GEMM's summation order: the batched column must then be at least as close
            # to the Float64 recurrence as the reference's own Float32 arithmetic is
            assert solver in ("CGNR", "ADMM", "SplitBregman"), (solver, k, e)
            x32 = getattr(O, solver)(A, **okw).solve(B[:, k].copy())
            okw64 = {kk: (float(v) if isinstance(v, np.floating) else v) for kk, v in okw.items()}
            x64 = getattr(O, solver)(up64(A), **okw64).solve(up64(B[:, k]))
            assert rel(Xb[:, k], x64) <= rel(x32, x64), (solver, k, e, rel(Xb[:, k], x64), rel(x32, x64))
        if solver in ("FISTA", "POGM", "OptISTA"):
            assert S.batch_iterations == [S.iteration] * 5
    else:
        assert np.array_equal(Xb, Xs)
    xv = rls.solve_(S, B[:, 0].copy())
    assert np.array_equal(xv, Xs[:, 0])


def test_measurement_based_normalization(rls, ctx):
    dtype = np.complex64
    A, xt, b = problem(dtype, 256, 512)
    rho = rho_for(A)
    for name in ("FISTA", "CGNR", "ADMM"):
        if name == "CGNR":
            kw1 = dict(reg=rls.L2Regularization(np.float32(1e-2))); kw2 = dict(reg=O.L2Regularization(np.float32(1e-2)))
        else:
            kw1 = dict(reg=rls.L1Regularization(np.float32(1e-2))); kw2 = dict(reg=O.L1Regularization(np.float32(1e-2)))
        if name == "FISTA":
            kw1["rho"] = kw2["rho"] = rho
        S = getattr(rls, name)(A, iterations=20, normalizeReg=rls.MeasurementBasedNormalization(), **kw1)
        R = getattr(O, name)(A, iterations=20, normalizeReg=O.MeasurementBasedNormalization(), **kw2)
        x = rls.solve_(S, b); xr = R.solve(b)
        assert_close_or_fp64(x, xr, lambda: getattr(O, name)(up64(A), iterations=20, normalizeReg=O.MeasurementBasedNormalization(),
                                                             **to64(kw2)).solve(up64(b)), what=name)
    S = rls.FISTA(A, iterations=20, rho=rho, reg=rls.L1Regularization(np.float32(1e-2)),
                  normalizeReg=rls.SystemMatrixBasedNormalization())
    R = O.FISTA(A, iterations=20, rho=rho, reg=O.L1Regularization(np.float32(1e-2)),
                normalizeReg=O.SystemMatrixBasedNormalization())
    assert rel(rls.solve_(S, b), R.solve(b)) < TOL


@pytest.mark.parametrize("solver", ["FISTA", "POGM", "OptISTA"])
@pytest.mark.parametrize("regname", ["TV", "L21"])
def test_measurement_based_normalization_non_elementwise_prox(rls, ctx, solver, regname):
    """init! re-normalises λ AFTER the solver state exists (FISTA.jl:128): the threshold ρλ of the TV / L21 prox must
    follow it — on the first solve and again when the next b has another scale."""
    dtype = np.complex64
    A, xt, b = problem(dtype, 192, 256)
    rho = rho_for(A)
    if regname == "TV":
        mk = lambda M: M.TVRegularization(np.float32(5e-2), shape=(16, 16))
    else:
        mk = lambda M: M.L21Regularization(np.float32(5e-2), slices=4)
    S = getattr(rls, solver)(A, iterations=15, rho=rho, relTol=0.0, reg=mk(rls), normalizeReg=rls.MeasurementBasedNormalization())
    R = getattr(O, solver)(A, iterations=15, rho=rho, relTol=0.0, reg=mk(O), normalizeReg=O.MeasurementBasedNormalization())
    for scale in (1.0, 37.0, 0.02):            # the factor ‖A'b‖₁/n changes from solve to solve
        bs = (b * np.float32(scale)).astype(dtype)
        x = rls.solve_(S, bs); xr = R.solve(bs)
        assert_close_or_fp64(x, xr, lambda: getattr(O, solver)(up64(A), iterations=15, rho=float(rho), relTol=0.0, reg=to64(mk(O)),
                                                               normalizeReg=O.MeasurementBasedNormalization()).solve(up64(bs)),
                             what=f"{solver} {regname} scale {scale}")


def test_power_iterations_and_default_rho(rls, ctx):
    dtype = np.complex64
    A, xt, b = problem(dtype, 256, 512)
    Ad = rls.B200Matrix.from_numpy(A, ctx)
    op = rls.B200NormalOp(Ad, form="twopass")
    b0 = rand_vector(dtype, 512, 77)
    lam_gpu = op.power_iterations(rls.B200Vector.from_numpy(b0, ctx))
    lam_ref = O.power_iterations(O.NormalOp(A), b0)
    assert abs(lam_gpu - lam_ref) < 1e-4 * lam_ref
    S = rls.FISTA(Ad, reg=rls.L1Regularization(np.float32(1e-3)), iterations=50)     # default rho
    x = rls.solve_(S, b)
    assert np.all(np.isfinite(x)) and rel(A @ x, b) < 0.5


def test_errors_are_reference_errors(rls, ctx):
    A = np.ones((8, 8), np.float32)
    with pytest.raises(ValueError, match="does not allow for more additional regularization terms"):
        rls.FISTA(A, reg=[rls.L1Regularization(1.0), rls.L2Regularization(1.0)], rho=0.1)
    with pytest.raises(TypeError, match="Float32 / ComplexF32"):
        rls.FISTA(np.ones((8, 8), np.float64), rho=0.1)
    S = rls.FISTA(A, rho=0.1)
    with pytest.raises(rls.RlsError):
        rls.solve_(S, np.ones(5, np.float32))     # wrong length b: status code, no abort
    with pytest.warns(UserWarning, match="filtered out"):
        rls.createLinearSolver(rls.CGNR, A, iterations=3, rho=0.1)


@pytest.mark.parametrize("dtype", DTYPES)
def test_multi_rhs_per_column_stopping(rls, ctx, dtype, monkeypatch):
    """MultiThreading.jl:45-78 keeps a per-column convergence mask: columns of very different difficulty stop at
    different iterations; the batched (tensor-core) driver must reproduce every column's count and iterate."""
    monkeypatch.setenv("RLS_BATCH_MIN_K", "2")
    m, n, K = 384, 160, 6
    A, _ = rand_matrix(dtype, m, n, 900)
    rho = rho_for(A)
    X = np.stack([sparse_truth(dtype, n, 910 + k, every=5 + 3 * k) for k in range(K)], axis=1)
    X[:, 1] *= 0                     # a zero right-hand side converges at once
    X[:, 4] *= np.float32(1e3)
    B = (A @ X).astype(dtype)
    S = rls.FISTA(A, reg=rls.L1Regularization(np.float32(1e-3)), iterations=60, rho=rho, relTol=np.float32(2e-3))
    Xb = rls.solve_(S, B)
    counts = list(S.batch_iterations)
    seq = []
    for k in range(K):
        xk = rls.solve_(S, B[:, k].copy())
        seq.append(S.iteration)
        assert rel(Xb[:, k], xk) < TOL or np.linalg.norm(xk) == 0, k
    assert counts == seq, (counts, seq)
    assert len(set(counts)) > 1, "the columns were meant to stop at different iterations"


@pytest.mark.parametrize("solver", ["FISTA", "CGNR", "ADMM"])
def test_multi_rhs_gram_form_on_tensor_cores(rls, ctx, solver, monkeypatch):
    """The reference's DEFAULT normal operator is the materialised Gram matrix (FISTA.jl:58, CGNR.jl:49, ADMM.jl:81); under
    the multi-RHS driver (MultiThreading.jl:45-78) its K applies G x_k per iteration run as ONE tcgen05 GEMM over G.
    Columns agree with the sequential Gram-form solves to the parity bound (CG-steered solvers: Float64 criterion) and
    keep their own iteration counts."""
    monkeypatch.setenv("RLS_BATCH_MIN_K", "2")
    dtype = np.complex64
    m, n, K = 201, 96, 6
    A, _, _ = problem(dtype, m, n)
    X = np.stack([sparse_truth(dtype, n, 400 + k, every=5 + 3 * k) for k in range(K)], axis=1)
    X[:, 4] *= np.float32(1e2)
    B = (A @ X).astype(dtype)
    kw, okw = dict(iterations=30), dict(iterations=30)
    if solver == "FISTA":
        kw.update(rho=rho_for(A), reg=rls.L1Regularization(np.float32(1e-4)), relTol=np.float32(1e-3))
        okw.update(rho=rho_for(A), reg=O.L1Regularization(np.float32(1e-4)), relTol=np.float32(1e-3))
    S = rls.createLinearSolver(getattr(rls, solver), rls.B200Matrix.from_numpy(A, ctx, layout="row"), normal="gram", **kw)
    assert S.AHA.form == "gram"
    Xb = rls.solve_(S, B)
    counts = list(S.batch_iterations)
    monkeypatch.setenv("RLS_GRAM_BATCH_TENSOR_CORES", "0")
    Xg = rls.solve_(S, B)                                   # K gemvs over G per iteration
    assert list(S.batch_iterations) == counts
    assert not np.array_equal(Xb, Xg), "the batched solve was meant to take the tensor-core GEMM over G"
    for k in range(K):
        e = rel(Xb[:, k], Xg[:, k])
        if e < TOL:
            continue
        assert solver in ("CGNR", "ADMM"), (solver, k, e)
        x32 = getattr(O, solver)(A, **okw).solve(B[:, k].copy())
        okw64 = {kk: (float(v) if isinstance(v, np.floating) else v) for kk, v in okw.items()}
        x64 = getattr(O, solver)(up64(A), **okw64).solve(up64(B[:, k]))
        assert rel(Xb[:, k], x64) <= rel(x32, x64), (solver, k, e, rel(Xb[:, k], x64), rel(x32, x64))


def test_multi_rhs_more_columns_than_one_gemm_tile(rls, ctx):
    """K * 2 > 128 complex columns do not fit one 128-wide GEMM tile: the driver falls back to per-column applies."""
    dtype = np.complex64
    m, n, K = 96, 64, 70
    A, _ = rand_matrix(dtype, m, n, 950)
    B = np.stack([rand_vector(dtype, m, 960 + k) for k in range(K)], axis=1)
    S = rls.CGNR(A, iterations=8, relTol=0.0)
    Xb = rls.solve_(S, B)
    for k in (0, 33, 69):
        assert np.array_equal(Xb[:, k], rls.solve_(S, B[:, k].copy()))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("form", ["twopass", "gram"])
def test_cgnr_whole_solve_graph_replay(rls, ctx, dtype, form, monkeypatch):
    """L2-resident systems are launch-latency bound (C1): from the second callback-free solve! on, the fixed launch
    sequence of a CGNR solve (every kernel gated on the device-side done() flag) is replayed from a CUDA graph.  Replays
    must be bit-identical to the launch-by-launch path for every right-hand side, including solves that stop early."""
    m, n = 96, 256
    A, _, _ = problem(dtype, m, n)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout="col")
    bs = [(A @ sparse_truth(dtype, n, 700 + k, every=5 + k)).astype(dtype) for k in range(4)]
    bs.append(np.zeros(m, dtype) + dtype(1e-3))
    kw = dict(reg=rls.L2Regularization(np.float32(1e-3)), iterations=25, relTol=np.float32(1e-4), normal=form)
    monkeypatch.setenv("RLS_SOLVE_GRAPH", "0")
    S0 = rls.CGNR(Ad, **kw)
    ref = [(rls.solve_(S0, b).copy(), S0.iteration) for b in bs]
    monkeypatch.setenv("RLS_SOLVE_GRAPH", "1")
    S1 = rls.CGNR(Ad, **kw)
    l0 = ctx.launch_count()
    x_first = rls.solve_(S1, bs[0])
    per_solve = ctx.launch_count() - l0                       # launch by launch (also sizes the scratch buffers)
    for rep in range(2):
        for b, (xr, itr) in zip(bs, ref):
            l0 = ctx.launch_count()
            x = rls.solve_(S1, b)                              # recorded on the first pass of this loop, replayed afterwards
            assert np.array_equal(x, xr) and S1.iteration == itr, (rep, itr, S1.iteration)
            assert ctx.launch_count() - l0 == per_solve        # the kernels inside the graph are counted
    assert np.array_equal(x_first, ref[0][0])
    assert len({itr for _, itr in ref}) > 1, "the right-hand sides were meant to stop at different iterations"
    # stepping through init! / iterate afterwards still works (no graph involved) and lands on the same iterates
    S1.init_(bs[1])
    while S1.iterate():
        pass
    assert np.array_equal(S1.x, ref[1][0])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,lam,reltol", [((96, 256), 1e-3, 1e-4), ((96, 256), 0.0, 0.0), ((300, 1000), 1e-2, 0.0),
                                               ((1024, 4096), 1e-3, 0.0), ((1500, 700), 1e-3, 1e-5), ((33, 2), 1e-3, 0.0)])
def test_cgnr_persistent_whole_solve_kernel(rls, ctx, dtype, shape, lam, reltol, monkeypatch):
    """RLS_CGNR_PERSISTENT=1: the whole CGNR solve of an L2-resident column-major system as ONE cooperative kernel (columns of
    A pinned to CTAs, four grid-wide exchanges per iteration, the chained path's scalar code on a private copy of the
    state).  Its sums run in another order than the chained kernels': same iteration count and stopping decision, iterates
    within the parity bound of the oracle (or as close to the Float64 recurrence as the Float32 oracle is)."""
    m, n = shape
    A, _, b = problem(dtype, m, n)
    its = min(30, n)
    reg = rls.L2Regularization(np.float32(lam)) if lam > 0 else None
    oreg = O.L2Regularization(np.float32(lam)) if lam > 0 else None
    kw = dict(iterations=its, relTol=np.float32(reltol))
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout="col")
    monkeypatch.setenv("RLS_CGNR_PERSISTENT", "0")
    Sc = rls.CGNR(Ad, reg=reg, **kw)
    xc = rls.solve_(Sc, b)
    monkeypatch.setenv("RLS_CGNR_PERSISTENT", "1")
    Sp = rls.CGNR(Ad, reg=reg, **kw)
    l0 = ctx.launch_count()
    xp = rls.solve_(Sp, b)
    launches = ctx.launch_count() - l0
    assert launches < 12, f"{launches} launches: the whole-solve kernel did not run"
    assert Sp.iteration == Sc.iteration, (Sp.iteration, Sc.iteration)
    R = O.CGNR(A, reg=oreg, **kw)
    x32 = R.solve(b)
    x64 = O.CGNR(up64(A), reg=(O.L2Regularization(float(np.float32(lam))) if lam > 0 else None), iterations=its,
                 relTol=float(np.float32(reltol))).solve(up64(b))
    assert Sp.iteration == R.iteration
    e = rel(xp, x32)
    assert e < TOL or rel(xp, x64) <= 1.5 * rel(x32, x64), (e, rel(xp, x64), rel(x32, x64), rel(xc, x64))
    # against the chained kernels: equal to rounding — or, where CG has run into the Float32 floor and every rounding is
    # amplified (the Float32 oracle itself is then that far from its Float64 run), both within that noise
    noise = 1.5 * rel(x32, x64)
    assert rel(xp, xc) < 5e-5 or (rel(xp, x64) <= noise and rel(xc, x64) <= noise), (rel(xp, xc), rel(xp, x64), rel(xc, x64), noise)
    assert abs(Sp._scalars.rel_res_norm - Sc._scalars.rel_res_norm) <= 1e-4 * Sc._scalars.rel_res_norm + 2e-6   # Float32 floor of ‖r‖/‖A'b‖
    xp2 = rls.solve_(Sp, b)                                    # deterministic, and the state it leaves behind is reusable
    assert np.array_equal(xp2, xp)


# ---------------------------------------------------------------- row-major device layout (one-pass cluster kernel)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("solver,kw", [("FISTA", dict(restart="none")), ("FISTA", dict(restart="gradient")),
                                       ("POGM", dict(restart="gradient")), ("OptISTA", {}), ("CGNR", {})])
def test_row_major_per_iterate(rls, ctx, dtype, solver, kw):
    """The layout the library chooses for large systems, forced here on a small one: every iterate against the
    oracle.  FISTA runs in its fused two-kernel form (momentum inside the operator's x load, finish inside the
    epilogue), the others through the one-pass operator + finish kernel."""
    A, xt, b = problem(dtype, 384, 1024)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout="row")
    assert Ad.layout == "row"
    if solver == "CGNR":
        S = rls.CGNR(Ad, reg=rls.L2Regularization(np.float32(1e-3)), iterations=20, relTol=0.0)
        R = O.CGNR(A, reg=O.L2Regularization(np.float32(1e-3)), iterations=20, relTol=0.0)
        R64 = O.CGNR(up64(A), reg=O.L2Regularization(float(np.float32(1e-3))), iterations=20, relTol=0.0)
        stepwise_vs_fp64(S, R, R64, b, 20)
        return
    rho = rho_for(A)
    lam = np.float32(2e-2)
    S = getattr(rls, solver)(Ad, reg=rls.L1Regularization(lam), iterations=40, rho=rho, relTol=0.0, **kw)
    assert S.AHA.form == "onepass" and "rowstream" in S.AHA.describe()
    R = getattr(O, solver)(A, reg=O.L1Regularization(lam), iterations=40, rho=rho, relTol=0.0, **kw)
    stepwise(S, R, b, 40)


@pytest.mark.parametrize("regname", ["TV", "L21", "L1+Positive"])
def test_row_major_fista_split_epilogue_and_projections(rls, ctx, regname, monkeypatch):
    """non-elementwise prox (the fused FISTA epilogue splits into PART 1 / prox kernels / PART 2) and projections on the
    fused path; RLS_FUSE_ITERATION=0 (four-kernel chain) must give bit-identical iterates."""
    dtype = np.complex64
    A, xt, b = problem(dtype, 256, 32 * 24)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout="row")
    rho = rho_for(A)
    mk = {"TV": lambda M: M.TVRegularization(np.float32(5e-3), shape=(32, 24)),
          "L21": lambda M: M.L21Regularization(np.float32(5e-3), slices=8),
          "L1+Positive": lambda M: [M.L1Regularization(np.float32(1e-2)), M.PositiveRegularization()]}[regname]
    S = rls.FISTA(Ad, reg=mk(rls), iterations=25, rho=rho, relTol=0.0, restart="gradient")
    R = O.FISTA(A, reg=mk(O), iterations=25, rho=rho, relTol=0.0, restart="gradient")
    R64 = O.FISTA(up64(A), reg=to64(mk(O)), iterations=25, rho=float(rho), relTol=0.0, restart="gradient")
    stepwise_vs_fp64(S, R, R64, b, 25)
    x_fused = rls.solve_(S, b)
    monkeypatch.setenv("RLS_FUSE_ITERATION", "0")
    x_chain = rls.solve_(S, b)
    assert np.array_equal(x_fused, x_chain)


# ---------------------------------------------------------------- SplitBregman
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", ["l1", "l1_stops_early", "two_terms", "l21_split_prox", "l1_gradient"])
def test_split_bregman(rls, ctx, dtype, case):
    """SplitBregman.jl:203-289 against the oracle, iterate by iterate: inner ADMM-like iterations, the Bregman update
    of β_y every `iterationsInner` inner iterations (or at convergence), `iter_cnt`, stopping decisions and the
    inner-CG counts."""
    A, xt, b = problem(dtype, 160, 96, seed=700)
    kw = dict(rho=0.5, iterations=3, iterationsInner=6, iterationsCG=8, absTol=0.0, relTol=0.0)
    if case == "l1":
        mk = lambda M: dict(reg=M.L1Regularization(np.float32(2e-2)))
    elif case == "l1_stops_early":
        mk = lambda M: dict(reg=M.L1Regularization(np.float32(2e-2)))
        kw.update(relTol=np.float32(3e-2), iterations=4, iterationsInner=25)
    elif case == "two_terms":
        mk = lambda M: dict(reg=[M.L1Regularization(np.float32(1e-2)), M.L2Regularization(np.float32(5e-2)), M.RealRegularization()],
                            regTrafo=[None, None])
        kw.update(rho=[0.5, 0.25])
    elif case == "l1_gradient":
        # the formulation the docstring recommends for TV (SplitBregman.jl:74): L1 on the finite differences
        mk = lambda M: dict(reg=M.L1Regularization(np.float32(1e-2)), regTrafo=M.GradientOp(dtype, shape=(12, 8)))
    else:
        mk = lambda M: dict(reg=M.L21Regularization(np.float32(5e-3), slices=8))
    S = rls.SplitBregman(A, **mk(rls), **kw)
    R = O.SplitBregman(A, **mk(O), **kw)
    R64 = O.SplitBregman(up64(A), **to64(mk(O)), **to64(kw))
    S.init_(b); R.init(b); R64.init(up64(b))
    steps = 0
    while True:
        a1, a2 = S.iterate(), R.iterate()
        a3 = R64.iterate()
        assert a1 == a2, f"stopping decision differs after {steps} iterations: gpu={a1} oracle={a2}"
        if not a1:
            break
        steps += 1
        assert S.iteration == R.iteration and S.iter_cnt == R.iter_cnt, (steps, S.iteration, R.iteration, S.iter_cnt, R.iter_cnt)
        # the Float64 run is only a yardstick while it walks the same path (same inner / outer counters)
        if a3 and R64.iteration == R.iteration and R64.iter_cnt == R.iter_cnt:
            assert_close_or_fp64(S.x, R.x, lambda: R64.x, what=f"iteration {steps}")
        else:
            assert rel(S.x, R.x) < TOL, f"iteration {steps}: rel-L2 {rel(S.x, R.x):.3e}"
        assert S._scalars.cg_iterations_last == R.cg_iters[-1]
    assert steps == R.total_iterations and steps > 0
    if case == "l1_stops_early":
        assert steps < 4 * 25
    # whole-solve fast path (iterations enqueued back to back, device-side gating of the Bregman updates): the same
    # kernels in the same order as the init!/iterate loop above
    x_step = S.x.copy()
    x = rls.solve_(S, b)
    assert np.array_equal(x, x_step)
    assert S.iter_cnt == R.iter_cnt


def test_split_bregman_reference_acceptance(rls, ctx):
    """test/testSolvers.jl:174-201 on the accelerated path (ComplexF32 instance of the 256-point DFT problem)."""
    rng = np.random.default_rng(1)
    N = 256
    F = (np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(N)) / N) / np.sqrt(N))
    x = np.zeros(N)
    for _ in range(3):
        x[rng.integers(0, N)] = rng.random()
    idx = np.sort(np.unique(rng.integers(0, N, N // 2)))
    F = F[idx, :].astype(np.complex64); b = (F @ x).astype(np.complex64)
    S = rls.createLinearSolver(rls.SplitBregman, F, reg=rls.L1Regularization(np.float32(2e-3)), iterations=5, iterationsInner=40,
                               rho=1.0, normalizeReg=rls.NoNormalization())
    xa = rls.solve_(S, b)
    assert np.linalg.norm(x - xa) <= 0.1 * np.linalg.norm(x)


def test_callbacks_like_the_reference(rls, ctx):
    """test/testCallbacks.jl:1-57 on the accelerated path: every callback fires iterations + 1 times, the last stored
    solution is the returned one, the comparison improves, several callbacks can be combined."""
    rng = np.random.default_rng(3)
    A = rng.random((32, 32)).astype(np.float32); x = rng.random(32).astype(np.float32); b = A @ x
    its = 10
    S = rls.createLinearSolver(rls.CGNR, A, iterations=its, relTol=0.0)
    cbk = rls.StoreSolutionCallback()
    xa = rls.solve_(S, b, callbacks=cbk)
    assert len(cbk.solutions) == its + 1 and np.array_equal(cbk.solutions[-1], xa)
    cmp = rls.CompareSolutionCallback(x)
    rls.solve_(S, b, callbacks=cmp)
    assert len(cmp.results) == its + 1 and cmp.results[0] > cmp.results[-1]
    conv = rls.StoreConvergenceCallback()
    rls.solve_(S, b, callbacks=conv)
    key = next(iter(conv.convMeas))
    assert len(conv.convMeas[key]) == its + 1 and conv.convMeas[key][-1] == rls.solverconvergence(S)[key]
    counter = []
    rls.solve_(S, b, callbacks=lambda s, it: counter.append(it))
    assert len(counter) == its + 1
    both = [rls.StoreSolutionCallback(), rls.StoreConvergenceCallback()]
    xa = rls.solve_(S, b, callbacks=both)
    assert len(both[0].solutions) == its + 1 and np.array_equal(both[0].solutions[-1], xa)
    assert len(both[1].convMeas[key]) == its + 1
