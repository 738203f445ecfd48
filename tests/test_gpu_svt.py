"""SURVEY §8(f) rank 4 on the GPU: singular-value thresholding prox maps (csrc/rls_svt.cu) through the C ABI against the
oracle's LAPACK-SVD restatement of ProxNuclear.jl:27-32 and ProxLLR.jl:44-90,163-199, at the shapes of the reference's own
tests (test/testProxMaps.jl:167-277), and inside FISTA / ADMM per iterate."""
import numpy as np
import pytest

import oracle as O
from util import rel, rand_matrix, rand_vector, sparse_truth, stepwise_vs_fp64, up64

pytestmark = pytest.mark.gpu
TOL = 1e-5
DTYPES = [np.float32, np.complex64]


def _rnd(rng, n, dt):
    return (rng.standard_normal(n) + (1j * rng.standard_normal(n) if np.dtype(dt).kind == "c" else 0)).astype(dt)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shp", [(40, 6), (6, 40), (5, 5), (200, 32), (64, 64), (3, 100), (100, 64), (5000, 16), (17, 3001), (1, 9), (9, 1)])
def test_prox_nuclear_vs_oracle(rls, ctx, dtype, shp):
    rng = np.random.default_rng(17)
    x = _rnd(rng, shp[0] * shp[1], dtype)
    smax = np.linalg.svd(x.reshape(shp, order="F"), compute_uv=False)[0]
    for frac in (0.0, 0.3, 0.9, 1.5):
        lam = np.float32(frac * smax)
        ref = O.prox_(O.NuclearRegularization(lam, svtShape=shp), x.copy())
        got = rls.prox_(rls.NuclearRegularization(lam, svtShape=shp), x.copy())
        if frac >= 1.5:
            assert not np.any(got) and np.linalg.norm(ref) <= 1e-5 * np.linalg.norm(x)
        else:
            assert rel(got, ref) < TOL, (frac, rel(got, ref))
        if frac == 0.0:
            assert rel(got, x) < 1e-6                      # λ = 0: W is the identity


def test_prox_nuclear_reference_test_problem(rls, ctx):
    """testNuclear (test/testProxMaps.jl:167-192): rank-2 32x32 matrix + noise, λ = 5σ"""
    from test_svt_host_logic import _nuclear_problem, _nuc_norm
    x, noisy, sigma = _nuclear_problem()
    lam = np.float32(5 * sigma)
    xl = rls.prox_(rls.NuclearRegularization, noisy.astype(np.complex64), lam, svtShape=(32, 32))
    assert np.linalg.norm(x - xl) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - xl) / np.linalg.norm(x) < 0.05
    assert 0.5 * np.linalg.norm(noisy - xl) ** 2 + _nuc_norm(xl, lam, (32, 32)) <= _nuc_norm(noisy, lam, (32, 32)) * (1 + 1e-6)
    assert rel(xl, O.prox_(O.NuclearRegularization(lam, svtShape=(32, 32)), noisy.astype(np.complex64))) < TOL


def test_prox_nuclear_graded_spectrum(rls, ctx):
    """singular values spread over six decades, two of them zero: the Float64 Gram + Jacobi path keeps Float32 accuracy"""
    rng = np.random.default_rng(8)
    U, _ = np.linalg.qr(rng.standard_normal((50, 8)) + 1j * rng.standard_normal((50, 8)))
    V, _ = np.linalg.qr(rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8)))
    S = np.array([1.0, 1.0, 0.5, 1e-2, 1e-4, 1e-6, 0.0, 0.0])
    x = ((U * S) @ V.conj().T).astype(np.complex64).reshape(-1, order="F")
    for lam in (np.float32(0.0), np.float32(1e-5), np.float32(0.25)):
        ref = ((U * np.maximum(S - lam, 0)) @ V.conj().T).reshape(-1, order="F")
        assert rel(rls.prox_(rls.NuclearRegularization(lam, svtShape=(50, 8)), x.copy()), ref) < 1e-6


def test_prox_svt_rejects_what_it_cannot_do(rls, ctx):
    x = np.zeros(100 * 70, np.float32)
    with pytest.raises(rls.RlsError, match="shorter side"):
        rls.prox_(rls.NuclearRegularization(np.float32(1), svtShape=(100, 70)), x)
    with pytest.raises(rls.RlsError, match="does not match"):
        rls.prox_(rls.NuclearRegularization(np.float32(1), svtShape=(10, 70)), x)
    with pytest.raises(rls.RlsError, match="multiple of blockSize"):
        rls.prox_(rls.LLRRegularization(np.float32(1), shape=(7, 10), blockSize=(2, 2), randshift=False, fullyOverlapping=True),
                  np.zeros(70 * 5, np.float32))
    with pytest.raises(rls.RlsError, match="min\\(frames, pixels per patch\\)"):
        rls.prox_(rls.LLRRegularization(np.float32(1), shape=(10, 10), blockSize=(10, 10), randshift=False), np.zeros(100 * 70, np.float32))


LLR_CASES = [((8, 8), (2, 2), 5, None, False), ((7, 9), (4, 4), 3, (1, 3), False), ((6,), (3,), 4, None, False),
             ((8, 8), (4, 4), 40, None, False), ((8, 8), (2, 2), 6, None, True), ((8, 4), (2, 2), 3, (1, 2), True),
             ((4, 4, 4), (2, 2, 2), 5, None, False), ((9,), (4,), 70, (2,), False), ((64, 64), (4, 4), 24, (3, 1), False),
             ((30, 20), (8, 8), 12, None, False)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", LLR_CASES)
def test_prox_llr_vs_oracle(rls, ctx, dtype, case):
    shape, block, K, shift, overlapping = case
    rng = np.random.default_rng(9)
    x = _rnd(rng, int(np.prod(shape)) * K, dtype)
    X = x.reshape(tuple(shape) + (K,), order="F")          # a view: the weak patches fall under the λ >= ub shortcut
    X[: max(block[0], shape[0] // 2)] *= np.float32(0.02)
    lam = np.float32(0.6)
    ref = O.prox_llr(x.copy(), lam, shape, block, shift, overlapping)
    reg = rls.LLRRegularization(lam, shape=shape, blockSize=block, randshift=False, fullyOverlapping=overlapping)
    got = rls.prox_(reg, x.copy(), shift=shift)
    assert np.any(ref != 0)
    assert rel(got, ref) < TOL, rel(got, ref)
    if shift is None and not overlapping:
        assert np.array_equal(got == 0, ref == 0)          # the shortcut zeroes the same patches


@pytest.mark.parametrize("overlapping", [False, True])
def test_prox_llr_reference_test_problem(rls, ctx, overlapping):
    """testLLR / testLLROverlapping (test/testProxMaps.jl:194-248): 32x32x80 series, 4x4 blocks (16 pixels x 80 frames per
    patch: the device works on the transposed view), λ = 10σ"""
    from test_svt_host_logic import _llr_problem, _llr_norm
    shape, block = (32, 32, 80), (4, 4)
    x, noisy, sigma = _llr_problem(shape, block)
    lam = np.float32(10 * sigma)
    reg = rls.LLRRegularization(lam, shape=shape[:2], blockSize=block, randshift=False, fullyOverlapping=overlapping)
    xl = rls.prox_(reg, noisy.astype(np.complex64))
    assert np.linalg.norm(x - xl) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - xl) / np.linalg.norm(x) < 0.05
    if not overlapping:
        assert 0.5 * np.linalg.norm(noisy - xl) ** 2 + _llr_norm(xl, lam, shape[:2], block) <= _llr_norm(noisy, lam, shape[:2], block) * (1 + 1e-6)
    ref = O.prox_(O.LLRRegularization(lam, shape=shape[:2], blockSize=block, fullyOverlapping=overlapping), noisy.astype(np.complex64))
    assert rel(xl, ref) < TOL


def test_prox_llr_3d_reference_test_problem(rls, ctx):
    """testLLR_3D (test/testProxMaps.jl:250-278) on 16^3 x 80: 4x4x4 blocks = 64 pixels x 80 frames per patch (q = 64)"""
    from test_svt_host_logic import _llr_problem
    shape, block = (16, 16, 16, 80), (4, 4, 4)
    x, noisy, sigma = _llr_problem(shape, block)
    lam = np.float32(10 * sigma)
    xl = rls.prox_(rls.LLRRegularization(lam, shape=shape[:3], blockSize=block, randshift=False), noisy.astype(np.complex64))
    assert np.linalg.norm(x - xl) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - xl) / np.linalg.norm(x) < 0.05
    assert rel(xl, O.prox_(O.LLRRegularization(lam, shape=shape[:3], blockSize=block), noisy.astype(np.complex64))) < TOL


def test_prox_llr_randshift_is_reproducible_and_a_valid_shift(rls, ctx):
    rng = np.random.default_rng(5)
    shape, block, K = (12, 10), (4, 2), 6
    x = _rnd(rng, 120 * K, np.complex64)
    lam = np.float32(0.8)
    a = rls.prox_(rls.LLRRegularization(lam, shape=shape, blockSize=block, randshift=True, seed=3), x.copy())
    b = rls.prox_(rls.LLRRegularization(lam, shape=shape, blockSize=block, randshift=True, seed=3), x.copy())
    assert np.array_equal(a, b)
    cands = [O.prox_llr(x.copy(), lam, shape, block, (s0, s1)) for s0 in range(1, 5) for s1 in range(1, 3)]
    assert min(rel(a, c) for c in cands) < TOL              # rand(CartesianIndices(blockSize)), ProxLLR.jl:55


@pytest.mark.parametrize("dtype", DTYPES)
def test_fista_with_nuclear_regularization_per_iterate(rls, ctx, dtype):
    """dynamic-imaging style: the unknown is a 64-pixel x 12-frame matrix of rank 2"""
    rng = np.random.default_rng(21)
    npix, K, m = 64, 12, 500
    n = npix * K
    L = (_rnd(rng, npix * 2, dtype).reshape(npix, 2) @ _rnd(rng, 2 * K, dtype).reshape(2, K)).astype(dtype)
    xt = L.reshape(-1, order="F")
    A, _ = rand_matrix(dtype, m, n, 31)
    b = (A @ xt).astype(dtype)
    rho = np.float32(0.9 / np.linalg.norm(A.astype(np.complex128), 2) ** 2)
    lam = np.float32(0.05)
    kw = dict(iterations=25, rho=rho, relTol=0.0)
    S = rls.FISTA(A, reg=rls.NuclearRegularization(lam, svtShape=(npix, K)), **kw)
    R = O.FISTA(A, reg=O.NuclearRegularization(lam, svtShape=(npix, K)), **kw)
    R64 = O.FISTA(up64(A), reg=O.NuclearRegularization(float(lam), svtShape=(npix, K)), iterations=25, rho=float(rho), relTol=0.0)
    w = stepwise_vs_fp64(S, R, R64, b, 25)
    assert S.iteration == 25
    print(f"FISTA + Nuclear {np.dtype(dtype).name}: worst gpu-o32 {w[0]:.2e}, gpu-o64 {w[1]:.2e}, o32-o64 {w[2]:.2e}")


def test_admm_with_llr_regularization_matches_the_oracle(rls, ctx):
    rng = np.random.default_rng(22)
    shape, block, K, m = (8, 8), (4, 4), 10, 400
    n = 64 * K
    dtype = np.complex64
    xt = np.zeros(shape + (K,), dtype, order="F")
    for i in range(2):
        for j in range(2):
            xt[4 * i:4 * i + 4, 4 * j:4 * j + 4, :] = (rng.random() + 0.5) * np.exp(-rng.random() * np.arange(K))
    xt = xt.reshape(-1, order="F")
    A, _ = rand_matrix(dtype, m, n, 33)
    b = (A @ xt).astype(dtype)
    lam = np.float32(0.02)
    kw = dict(iterations=8, iterationsCG=6, rho=0.5)
    S = rls.ADMM(A, reg=rls.LLRRegularization(lam, shape=shape, blockSize=block, randshift=False), **kw)
    R = O.ADMM(A, reg=O.LLRRegularization(lam, shape=shape, blockSize=block), **kw)
    R64 = O.ADMM(up64(A), reg=O.LLRRegularization(float(lam), shape=shape, blockSize=block), **kw)
    stepwise_vs_fp64(S, R, R64, b, 8)
