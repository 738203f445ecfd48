"""GPU parity of the building blocks against NumPy / the oracle, through the C ABI."""
import numpy as np
import pytest

import oracle as O
from oracle.philox import philox_matrix, philox_vector, IH4, UNIFORM01
from util import rel, rand_matrix, rand_vector

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.complex64]
# (m, n): aligned, ragged rows (m % 4 != 0), single row / column, tall, wide
SHAPES = [(256, 512), (1023, 771), (1, 17), (19, 1), (4100, 300), (130, 6000), (8192 + 8, 1536)]


LAYOUTS = ["row", "col"]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("layout", LAYOUTS)
def test_philox_bit_exact(rls, ctx, dtype, layout):
    A = rls.B200Matrix.philox(dtype, 300, 70, seed=99, scale=0.25, ctx=ctx, layout=layout).to_numpy()
    assert np.array_equal(A, philox_matrix(dtype, 300, 70, 99, IH4, 0.25))
    # a row shard regenerates exactly its rows of the global matrix
    S = rls.B200Matrix.philox(dtype, 100, 70, seed=99, scale=0.25, row_offset=120, m_global=300, ctx=ctx,
                              layout=layout).to_numpy()
    assert np.array_equal(S, A[120:220])
    v = rls.B200Vector(ctx, dtype, 1000).fill_philox(5, stream=3, dist=0, scale=2.0).to_numpy()
    assert np.array_equal(v, philox_vector(dtype, 1000, 5, 3, UNIFORM01, 2.0))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES + [(37, 20000), (9, 70001)])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_gemv_n_c(rls, ctx, dtype, shape, layout):
    m, n = shape
    if layout == "row" and n * (2 if np.dtype(dtype).kind == "c" else 1) > 131072:
        pytest.skip("row-major layout supports rows of at most 131072 floats")
    A, _ = rand_matrix(dtype, m, n, 11)
    x = rand_vector(dtype, n, 12)
    y = rand_vector(dtype, m, 13)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout=layout)
    assert Ad.layout == layout
    assert np.array_equal(Ad.to_numpy(), A)
    yd = Ad.mul(rls.B200Vector.from_numpy(x, ctx)).to_numpy()
    gd = Ad.adjoint_mul(rls.B200Vector.from_numpy(y, ctx)).to_numpy()
    A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    assert rel(yd, A64 @ x) < 2e-6
    assert rel(gd, A64.conj().T @ y) < 2e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("form", ["twopass", "onepass", "gram"])
@pytest.mark.parametrize("shape", [(256, 512), (1023, 772), (515, 2048), (64, 4096), (3000, 1000), (50, 20000), (23, 40000)])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_normal_operator_forms(rls, ctx, dtype, form, shape, layout):
    m, n = shape
    if form == "gram" and n > 4096:
        pytest.skip("Gram matrix of a wide system is not interesting here")
    A, _ = rand_matrix(dtype, m, n, 21)
    x = rand_vector(dtype, n, 22)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout=layout)
    op = rls.B200NormalOp(Ad, form=form)
    assert op.form == form
    xd = rls.B200Vector.from_numpy(x, ctx)
    g1 = op.apply(xd).to_numpy()
    g2 = op.apply(xd).to_numpy()           # second launch: tickets / flags / tags must have reset
    A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    ref = A64.conj().T @ (A64 @ x)
    assert rel(g1, ref) < 3e-6
    assert np.array_equal(g1, g2), "normal operator is not deterministic run to run"


@pytest.mark.parametrize("lpc", [8, 16])
@pytest.mark.parametrize("stage_kb,lag,hint", [(32, 2, 1), (8, 1, 0), (64, 6, 1), (16, 3, 0)])
def test_onepass_tma_variants(rls, ctx, lpc, stage_kb, lag, hint, monkeypatch):
    """the TMA two-phase streaming kernel under different segment widths, stage sizes, lags and L2 hints"""
    monkeypatch.setenv("RLS_TMA_LPC", str(lpc))
    monkeypatch.setenv("RLS_TMA_STAGE_KB", str(stage_kb))
    monkeypatch.setenv("RLS_TMA_LAG", str(lag))
    monkeypatch.setenv("RLS_TMA_HINT", str(hint))
    for dtype in DTYPES:
        for (m, n) in [(2052, 9000), (70, 33), (1, 5000), (4099, 20000), (300, 70000)]:
            A, _ = rand_matrix(dtype, m, n, 31)
            x = rand_vector(dtype, n, 32)
            op = rls.B200NormalOp(rls.B200Matrix.from_numpy(A, ctx, layout="col"), form="onepass")
            xd = rls.B200Vector.from_numpy(x, ctx)
            g = op.apply(xd).to_numpy()
            A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
            assert rel(g, A64.conj().T @ (A64 @ x)) < 3e-6, (dtype, m, n)
            assert np.array_equal(g, op.apply(xd).to_numpy())


@pytest.mark.parametrize("lpc", [8, 16, 32])
@pytest.mark.parametrize("lag", [1, 3])
def test_onepass_variants(rls, ctx, lpc, lag, monkeypatch):
    """the L2-lag fallback kernel (used when the TMA kernel cannot describe the matrix)"""
    monkeypatch.setenv("RLS_ONEPASS_IMPL", "l2")
    monkeypatch.setenv("RLS_ONEPASS_LPC", str(lpc))
    monkeypatch.setenv("RLS_ONEPASS_LAG", str(lag))
    for dtype in DTYPES:
        m, n = 2052, 9000
        A, _ = rand_matrix(dtype, m, n, 31)
        x = rand_vector(dtype, n, 32)
        op = rls.B200NormalOp(rls.B200Matrix.from_numpy(A, ctx, layout="col"), form="onepass")
        g = op.apply(rls.B200Vector.from_numpy(x, ctx)).to_numpy()
        A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
        assert rel(g, A64.conj().T @ (A64 @ x)) < 3e-6


@pytest.mark.parametrize("dtype", DTYPES)
def test_reductions(rls, ctx, dtype):
    a = rand_vector(dtype, 100003, 41)
    b = rand_vector(dtype, 100003, 42)
    ad, bd = rls.B200Vector.from_numpy(a, ctx), rls.B200Vector.from_numpy(b, ctx)
    assert abs(ad.norm() - np.linalg.norm(a.astype(np.complex128))) < 1e-9 * 400
    assert abs(ad.asum() - np.sum(np.abs(a.astype(np.complex128)))) < 1e-3
    assert abs(ad.dot(bd) - np.vdot(a.astype(np.complex128), b.astype(np.complex128))) < 1e-6 * 400
    e = rls.B200Vector(ctx, dtype, 0)
    assert e.norm() == 0.0
    A, _ = rand_matrix(dtype, 301, 77, 43)
    assert abs(rls.B200Matrix.from_numpy(A, ctx).frob2() - np.sum(np.abs(A.astype(np.complex128)) ** 2)) < 1e-6 * 77


# ---------------------------------------------------------------- proximal maps
@pytest.mark.parametrize("dtype", DTYPES)
def test_prox_elementwise(rls, ctx, dtype):
    x = rand_vector(dtype, 5000, 51)
    x[::7] = 0
    for lam in (np.float32(0.3), np.float32(0.0), 0.7):
        a = x.copy(); b = x.copy()
        rls.prox_(rls.L1Regularization(lam), a)
        O.prox_(O.L1Regularization(lam), b)
        assert rel(a, b) < 1e-6, "L1"
        a = x.copy(); b = x.copy()
        rls.prox_(rls.L2Regularization(lam), a)
        O.prox_(O.L2Regularization(lam), b)
        assert np.array_equal(a, b), "L2 is bit exact"
    a = x.copy(); b = x.copy()
    rls.prox_(rls.PositiveRegularization, a); O.prox_(O.PositiveRegularization(), b)
    assert np.array_equal(a, b)
    a = x.copy(); b = x.copy()
    rls.prox_(rls.RealRegularization, a); O.prox_(O.RealRegularization(), b)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype", DTYPES)
def test_prox_l1_bit_exact_vs_oracle(rls, ctx, dtype):
    x = rand_vector(dtype, 4096, 52)
    a = x.copy(); b = x.copy()
    rls.prox_(rls.L1Regularization(np.float32(0.4)), a)
    O.prox_(O.L1Regularization(np.float32(0.4)), b)
    # every op individually rounded on both sides; complex modulus via double hypot
    assert np.max(np.abs(a - b)) <= 2 * np.finfo(np.float32).eps * np.max(np.abs(b))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("slices", [1, 8, 5])
def test_prox_l21(rls, ctx, dtype, slices):
    x = rand_vector(dtype, 256 * slices, 53)
    a = x.copy(); b = x.copy()
    rls.prox_(rls.L21Regularization(np.float32(1.5), slices=slices), a)
    O.prox_(O.L21Regularization(np.float32(1.5), slices=slices), b)
    assert rel(a, b) < 1e-6
    # all-zero group with λ = 0 -> 0/0 = NaN is preserved (reference quirk)
    z = np.zeros(16, dtype)
    rls.prox_(rls.L21Regularization(np.float32(0.0), slices=4), z)
    assert np.all(np.isnan(z.real))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,slices", [(1003, 8), (37, 5), (9, 4), (13, 13)])
def test_prox_l21_length_not_a_multiple_of_slices(rls, ctx, dtype, n, slices):
    """ProxL21.jl:30-35: L = n ÷ slices and group j is x[j:L:end] — the trailing n - L*slices elements belong to the
    first groups (they enter the group norms and are rescaled), nothing is left untouched."""
    x = rand_vector(dtype, n, 59)
    a = x.copy(); b = x.copy()
    rls.prox_(rls.L21Regularization(np.float32(0.7), slices=slices), a)
    O.prox_(O.L21Regularization(np.float32(0.7), slices=slices), b)
    assert rel(a, b) < 1e-6
    L = n // slices
    if n % slices:
        assert not np.array_equal(a[L * slices:], x[L * slices:])     # the tail was thresholded too
    with pytest.raises(Exception):                                    # more slices than elements: x[i:0:end] throws upstream
        rls.prox_(rls.L21Regularization(np.float32(0.7), slices=n + 1), x.copy())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,dims", [((64, 48), None), ((64, 48), (1,)), ((64, 48), (2,)), ((300,), None),
                                        ((12, 10, 9), None), ((12, 10, 9), (1, 3)), ((1, 40), None)])
def test_gradient_op_and_tv(rls, ctx, dtype, shape, dims):
    n = int(np.prod(shape))
    x = rand_vector(dtype, n, 54)
    dd = tuple(range(1, len(shape) + 1)) if dims is None else dims
    G = rls.GradientOp(dtype, shape, dims)
    Go = O.GradientOp(dtype, shape, dims)
    assert G.rows == Go.rows
    xd = rls.B200Vector.from_numpy(x, ctx)
    gx = G.mul(xd)
    assert np.array_equal(gx.to_numpy(), Go.mul(x))
    g = rand_vector(dtype, G.rows, 55)
    assert rel(G.tmul(rls.B200Vector.from_numpy(g, ctx)).to_numpy(), Go.tmul(g)) < 1e-6
    for iters in (1, 10):
        a = x.copy(); b = x.copy()
        rls.prox_(rls.TVRegularization(np.float32(0.2), shape=shape, dims=dims, iterationsTV=iters), a)
        O.prox_(O.TVRegularization(np.float32(0.2), shape=shape, dims=dd, iterationsTV=iters), b)
        assert rel(a, b) < 2e-6, f"TV {shape} dims={dims} iters={iters}"


def test_tv_denoises_piecewise_constant(rls, ctx):
    """test/testProxMaps.jl:75-103 properties on the device path (Float32 edition)."""
    rng = np.random.default_rng(1234)
    N = 128
    x = np.zeros((N, N), np.complex64, order="F")
    for _ in range(5):
        i, j = rng.integers(0, N, 2)
        x[i:, j:] += np.float32(rng.standard_normal())
    x = x.ravel(order="F")
    sigma = np.float32(np.sum(np.abs(x)) / x.size * 0.05)
    noisy = (x + sigma / np.sqrt(2.0) * (rng.standard_normal(N * N) + 1j * rng.standard_normal(N * N))).astype(np.complex64)
    x_tv = noisy.copy()
    rls.prox_(rls.TVRegularization(np.float32(2 * sigma), shape=(N, N)), x_tv)
    x_l1 = noisy.copy()
    rls.prox_(rls.L1Regularization(np.float32(2 * sigma)), x_l1)
    assert np.linalg.norm(x - x_tv) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - x_tv) <= np.linalg.norm(x - x_l1)
    tv = lambda v: 2 * sigma * np.sum(np.abs(O.grad_op(v, (N, N), (1, 2))))
    assert 0.5 * np.linalg.norm(noisy - x_tv) ** 2 + tv(x_tv) <= tv(noisy)


# ---------------------------------------------------------------- tensor-core paths (csrc/rls_tc.cu)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape_k", [(128, 128, 32), (300, 200, 5), (1000, 516, 64), (77, 1030, 3), (2050, 4100, 17)])
def test_normal_apply_batch_tensor_cores(rls, ctx, dtype, shape_k, monkeypatch):
    """K right-hand sides through two tcgen05 GEMMs (kind::tf32, three-term split, FP32 accumulation outside the
    tensor core) against NumPy float64 and against K single CUDA-core applies."""
    monkeypatch.setenv("RLS_BATCH_MIN_K", "2")
    m, n, K = shape_k
    A, _ = rand_matrix(dtype, m, n, 51)
    X = np.stack([rand_vector(dtype, n, 60 + k) for k in range(K)], axis=1)
    op = rls.B200NormalOp(rls.B200Matrix.from_numpy(A, ctx, layout="row"), form="onepass")
    xs = [rls.B200Vector.from_numpy(np.ascontiguousarray(X[:, k]), ctx) for k in range(K)]
    G = np.stack([o.to_numpy() for o in op.apply_batch(xs)], axis=1)
    A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    ref = A64.conj().T @ (A64 @ X)
    for k in range(K):
        assert rel(G[:, k], ref[:, k]) < 1.5e-6, (k, rel(G[:, k], ref[:, k]))
    # column-major storage: no tensor-core plan, falls back to K single applies with identical results
    opc = rls.B200NormalOp(rls.B200Matrix.from_numpy(A, ctx, layout="col"), form="twopass")
    Gc = np.stack([o.to_numpy() for o in opc.apply_batch(xs)], axis=1)
    assert np.array_equal(Gc[:, 0], opc.apply(xs[0]).to_numpy())
    assert rel(Gc, ref) < 3e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(256, 128), (300, 200), (1000, 516)])
def test_gram_on_tensor_cores(rls, ctx, dtype, shape, monkeypatch):
    m, n = shape
    A, _ = rand_matrix(dtype, m, n, 71)
    x = rand_vector(dtype, n, 72)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout="row")
    op = rls.B200NormalOp(Ad, form="gram")
    assert "tensor cores" in op.describe()
    A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    ref = (A64.conj().T @ A64) @ x
    g = op.apply(rls.B200Vector.from_numpy(x, ctx)).to_numpy()
    assert rel(g, ref) < 1.5e-6
    monkeypatch.setenv("RLS_GRAM_CUDA_CORES", "1")
    op2 = rls.B200NormalOp(Ad, form="gram")
    assert "tensor cores" not in op2.describe()
    assert rel(op2.apply(rls.B200Vector.from_numpy(x, ctx)).to_numpy(), g) < 2e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape_k", [(256, 128, 32), (300, 200, 5), (700, 516, 64), (90, 1032, 17)])   # Gram form: n % 4 == 0
def test_gram_apply_batch_tensor_cores(rls, ctx, dtype, shape_k, monkeypatch):
    """Gram form (the reference's default AHA = A'*A, FISTA.jl:58) under the multi-RHS driver: K columns G x_k as ONE
    tcgen05 GEMM over G, against NumPy float64 and against the K single gemvs; a user-supplied non-Hermitian AHA
    (FISTA.jl:55) is applied as given, not as its adjoint."""
    monkeypatch.setenv("RLS_BATCH_MIN_K", "2")
    m, n, K = shape_k
    A, _ = rand_matrix(dtype, m, n, 81)
    X = np.stack([rand_vector(dtype, n, 90 + k) for k in range(K)], axis=1)
    xs = [rls.B200Vector.from_numpy(np.ascontiguousarray(X[:, k]), ctx) for k in range(K)]
    op = rls.B200NormalOp(rls.B200Matrix.from_numpy(A, ctx, layout="row"), form="gram")
    R = np.stack([o.to_numpy() for o in op.apply_batch(xs)], axis=1)
    A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    ref = (A64.conj().T @ A64) @ X
    for k in range(K):
        assert rel(R[:, k], ref[:, k]) < 1.5e-6, (k, rel(R[:, k], ref[:, k]))
    monkeypatch.setenv("RLS_GRAM_BATCH_TENSOR_CORES", "0")
    R1 = np.stack([o.to_numpy() for o in op.apply_batch(xs)], axis=1)        # K gemvs over G
    assert np.array_equal(R1[:, 0], op.apply(xs[0]).to_numpy())
    assert rel(R, R1) < 2e-6 and not np.array_equal(R, R1)                    # the tensor-core path really ran
    monkeypatch.setenv("RLS_GRAM_BATCH_TENSOR_CORES", "1")
    # AHA handed over as a matrix that is NOT Hermitian
    H, _ = rand_matrix(dtype, n, n, 83)
    oph = rls.B200NormalOp(G=rls.B200Matrix.from_numpy(H, ctx, layout="col"))
    Rh = np.stack([o.to_numpy() for o in oph.apply_batch(xs)], axis=1)
    refh = H.astype(A64.dtype) @ X
    assert rel(Rh, refh) < 1.5e-6, rel(Rh, refh)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("shape", [(0, 5), (5, 0), (0, 0), (1, 1), (2, 3)])
def test_degenerate_shapes(rls, ctx, dtype, layout, shape):
    """empty and tiny systems: A x of an m = 0 system is empty, A'(A x) is the zero vector, nothing hangs"""
    m, n = shape
    A = np.zeros((m, n), dtype)
    if m and n:
        A, _ = rand_matrix(dtype, m, n, 5)
    Ad = rls.B200Matrix.from_numpy(A, ctx, layout=layout)
    assert Ad.to_numpy().shape == (m, n)
    x = rand_vector(dtype, n, 6) if n else np.zeros(0, dtype)
    y = rand_vector(dtype, m, 7) if m else np.zeros(0, dtype)
    xd, yd = rls.B200Vector.from_numpy(x, ctx), rls.B200Vector.from_numpy(y, ctx)
    A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    assert np.allclose(Ad.mul(xd).to_numpy(), A64 @ x, atol=1e-5)
    assert np.allclose(Ad.adjoint_mul(yd).to_numpy(), A64.conj().T @ y, atol=1e-5)
    for form in ("twopass", "onepass"):
        if n == 0:
            continue
        g = rls.B200NormalOp(Ad, form=form).apply(xd).to_numpy()
        assert np.allclose(g, A64.conj().T @ (A64 @ x), atol=1e-5), form


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,n", [(1, 1), (33, 70), (257, 1024), (100, 4099)])
def test_device_relayout(rls, ctx, dtype, m, n):
    """rls_mat_relayout: an adopted column-major matrix (Julia's layout) becomes the row-major layout of the one-pass
    kernels on the device, and back — the very same entries, the very same operator results."""
    A, _ = rand_matrix(dtype, m, n, 611)
    Ac = rls.B200Matrix.from_numpy(A, ctx, layout="col")
    Ar = Ac.relayout("row")
    assert Ar.layout == "row" and np.array_equal(Ar.to_numpy(), A)
    assert np.array_equal(Ar.relayout("col").to_numpy(), A)
    assert np.array_equal(Ac.relayout("col").to_numpy(), A)          # same layout: a plain copy
    x = rls.B200Vector.from_numpy(rand_vector(dtype, n, 612), ctx)
    ref = rls.B200NormalOp(rls.B200Matrix.from_numpy(A, ctx, layout="row"), form="onepass").apply(x).to_numpy()
    assert np.array_equal(rls.B200NormalOp(Ar, form="onepass").apply(x).to_numpy(), ref)


def test_handles_may_be_destroyed_in_any_order(rls):
    """Garbage collectors (Julia finalizers, Python weakrefs) free handles in arbitrary order: the library counts
    references, so destroying the context, the matrix and the operator BEFORE the solver that uses them is harmless —
    the solver keeps working and the memory goes with the last holder."""
    import oracle as O
    c = rls.B200Context(0)
    A, _ = rand_matrix(np.float32, 96, 160, 71)
    b = rand_vector(np.float32, 96, 72)
    Ad = rls.B200Matrix.from_numpy(A, c, layout="row")
    op = rls.B200NormalOp(Ad, form="onepass")
    S = rls.FISTA(Ad, AHA=op, reg=rls.L1Regularization(np.float32(1e-2)), iterations=10, rho=np.float32(0.05), relTol=0.0, ctx=c)
    K = rls.Kaczmarz(Ad, reg=rls.L2Regularization(np.float32(1e-2)), iterations=2)
    x_before = rls.solve_(S, b)
    for obj in (c, Ad, op):          # rls_ctx_destroy, rls_mat_destroy, rls_normal_destroy — now, in the worst order
        obj._fin()
    x_after = rls.solve_(S, b)
    assert np.array_equal(x_before, x_after)
    xr = O.FISTA(A, reg=O.L1Regularization(np.float32(1e-2)), iterations=10, rho=np.float32(0.05), relTol=0.0).solve(b)
    assert rel(x_after, xr) < 1e-5
    assert np.all(np.isfinite(rls.solve_(K, b)))
    S._fin(); K._fin()
