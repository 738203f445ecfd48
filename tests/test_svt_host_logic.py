"""SURVEY §8(f) rank 4, CPU side: the oracle's restatement of ProxNuclear.jl / ProxLLR.jl is pinned to the reference's own
acceptance tests (test/testProxMaps.jl:167-277: denoising error bounds and objective decrease), and the algorithm the
device uses (Gram matrix of the short side + Jacobi in Float64, tests/svt_mirror.py mirrors csrc/rls_svt.cu) is checked
against it.  The kernels themselves are checked in tests/test_gpu_svt.py."""
import numpy as np
import pytest

import oracle as O
import svt_mirror as M
from util import rel


def _nuclear_problem(N=32, rank=2, sigma=0.05, seed=1234):
    """test/testProxMaps.jl:167-184"""
    rng = np.random.default_rng(seed)
    x = np.zeros((N, N), np.complex128)
    for i in range(rank):
        x[:, i] = (0.3 + 0.7 * rng.standard_normal()) * np.cos(2 * np.pi / N * rng.integers(1, N // 4 + 1) * np.arange(1, N + 1))
    for i in range(rank, N):
        for j in range(rank):
            x[:, i] += rng.random() * x[:, j]
    x = x.reshape(-1, order="F")
    sigma = np.sum(np.abs(x)) / x.size * sigma
    noisy = x + sigma / np.sqrt(2.0) * (rng.standard_normal(N * N) + 1j * rng.standard_normal(N * N))
    return x, noisy, sigma


def _llr_problem(shape, block, sigma=0.05, seed=1234):
    """test/testProxMaps.jl:194-211 (and :250-269 in 3-D): one decaying exponential per block"""
    rng = np.random.default_rng(seed)
    x = np.zeros(shape, np.complex128)
    nb = [shape[d] // block[d] for d in range(len(block))]
    t = np.arange(1, shape[-1] + 1)
    for idx in np.ndindex(*nb):
        ampl, r = rng.random(), rng.random()
        sl = tuple(slice(idx[d] * block[d], (idx[d] + 1) * block[d]) for d in range(len(block)))
        x[sl] = ampl * np.exp(-r * t)
    x = x.reshape(-1, order="F")
    sigma = np.sum(np.abs(x)) / x.size * sigma
    noisy = x + sigma / np.sqrt(2.0) * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))
    return x, noisy, sigma


def _nuc_norm(x, lam, shape):
    return lam * np.sum(np.linalg.svd(x.reshape(shape, order="F"), compute_uv=False))


def _llr_norm(x, lam, shape, block):
    """norm(::LLRRegularization) for non-overlapping 2-D blocks without shift (ProxLLR.jl:97-161)"""
    X = x.reshape(tuple(shape) + (-1,), order="F")
    tot = 0.0
    for i in range(0, shape[0], block[0]):
        for j in range(0, shape[1], block[1]):
            tot += np.sum(np.linalg.svd(X[i:i + block[0], j:j + block[1]].reshape((-1, X.shape[-1]), order="F"), compute_uv=False))
    return lam * tot


def test_oracle_nuclear_like_the_reference_test():
    """testNuclear (test/testProxMaps.jl:167-192), ComplexF64 as there and ComplexF32 as on the device path"""
    x, noisy, sigma = _nuclear_problem()
    for dt in (np.complex128, np.complex64):
        lam = 5 * sigma
        xl = O.prox_(O.NuclearRegularization(lam, svtShape=(32, 32)), noisy.astype(dt).copy())
        assert np.linalg.norm(x - xl) <= np.linalg.norm(x - noisy)
        assert np.linalg.norm(x - xl) / np.linalg.norm(x) < 0.05
        assert 0.5 * np.linalg.norm(noisy - xl) ** 2 + _nuc_norm(xl, lam, (32, 32)) <= _nuc_norm(noisy, lam, (32, 32)) * (1 + 1e-6)
        assert O.reg_norm(O.NuclearRegularization(lam, svtShape=(32, 32)), noisy.astype(dt)) == pytest.approx(_nuc_norm(noisy, lam, (32, 32)), rel=1e-5)


@pytest.mark.parametrize("overlapping", [False, True])
def test_oracle_llr_like_the_reference_test(overlapping):
    """testLLR / testLLROverlapping (test/testProxMaps.jl:194-248) at a quarter of the frames (20 instead of 80)"""
    shape, block = (32, 32, 20), (4, 4)
    x, noisy, sigma = _llr_problem(shape, block)
    lam = 10 * sigma
    xl = O.prox_(O.LLRRegularization(lam, shape=shape[:2], blockSize=block, randshift=False, fullyOverlapping=overlapping),
                 noisy.astype(np.complex64).copy())
    assert np.linalg.norm(x - xl) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - xl) / np.linalg.norm(x) < 0.05
    if not overlapping:
        assert 0.5 * np.linalg.norm(noisy - xl) ** 2 + _llr_norm(xl, lam, shape[:2], block) <= _llr_norm(noisy, lam, shape[:2], block) * (1 + 1e-6)


def test_oracle_llr_3d_like_the_reference_test():
    """testLLR_3D (test/testProxMaps.jl:250-278) on a 16^3 x 20 series"""
    shape, block = (16, 16, 16, 20), (4, 4, 4)
    x, noisy, sigma = _llr_problem(shape, block)
    xl = O.prox_(O.LLRRegularization(10 * sigma, shape=shape[:3], blockSize=block, randshift=False), noisy.astype(np.complex64).copy())
    assert np.linalg.norm(x - xl) <= np.linalg.norm(x - noisy)
    assert np.linalg.norm(x - xl) / np.linalg.norm(x) < 0.05


def test_oracle_llr_shortcut_uses_the_largest_entry_of_the_gram_matrix():
    """ProxLLR.jl:67-71: `norm(x2, Inf)` of a MATRIX is its largest |entry| in Julia (not the operator norm), so a patch is
    zeroed as soon as λ reaches the largest frame norm even though its largest singular value is above λ."""
    X = np.ones((4, 3), np.float32)                   # one 2x2 patch, 3 identical frames: s_max = sqrt(12), frame norm = 2
    x = X.reshape(-1, order="F").copy()
    out = O.prox_llr(x.copy(), np.float32(2.5), (2, 2), (2, 2))
    assert np.all(out == 0)
    out = O.prox_llr(x.copy(), np.float32(1.5), (2, 2), (2, 2))
    assert np.allclose(out, (np.sqrt(12) - 1.5) / np.sqrt(12), rtol=1e-6)


def test_round_robin_ordering_visits_every_pair_once_per_sweep():
    for q in (1, 2, 3, 5, 8, 16, 31, 32, 33, 63, 64):
        qe = q + (q & 1)
        seen = set()
        for rr in range(qe - 1):
            ps = M.pairs_of_round(q, rr)
            flat = [i for pr in ps for i in pr]
            assert len(flat) == len(set(flat)) and len(ps) <= 32          # disjoint within a round, one lane per pair
            assert not (seen & set(ps))
            seen |= set(ps)
        assert len(seen) == q * (q - 1) // 2


def _rnd(rng, n, dt):
    return (rng.standard_normal(n) + (1j * rng.standard_normal(n) if np.dtype(dt).kind == "c" else 0)).astype(dt)


@pytest.mark.parametrize("dt", [np.float32, np.complex64])
@pytest.mark.parametrize("shp", [(40, 6), (6, 40), (5, 5), (200, 32), (64, 64), (3, 100), (100, 64)])
def test_device_algorithm_nuclear_equals_lapack_svd(dt, shp):
    rng = np.random.default_rng(7)
    x = _rnd(rng, shp[0] * shp[1], dt)
    for frac in (0.0, 0.3, 0.9, 1.5):
        lam = np.float32(frac * np.linalg.svd(x.reshape(shp, order="F"), compute_uv=False)[0])
        assert rel(M.prox_nuclear(x, lam, *shp), O.prox_nuclear(x.copy(), lam, shp)) < 1e-6 or frac >= 1.5
        if frac >= 1.5:
            assert not np.any(M.prox_nuclear(x, lam, *shp))


def test_device_algorithm_nuclear_rank_deficient_and_graded():
    """repeated / zero singular values and a 1e-6 spread: the Float64 Gram + Jacobi path keeps Float32 accuracy"""
    rng = np.random.default_rng(8)
    U, _ = np.linalg.qr(rng.standard_normal((50, 8)) + 1j * rng.standard_normal((50, 8)))
    V, _ = np.linalg.qr(rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8)))
    S = np.array([1.0, 1.0, 0.5, 1e-2, 1e-4, 1e-6, 0.0, 0.0])
    X = ((U * S) @ V.conj().T).astype(np.complex64)
    x = X.reshape(-1, order="F")
    for lam in (np.float32(0.0), np.float32(1e-5), np.float32(0.25)):
        ref = ((U * np.maximum(S - lam, 0)) @ V.conj().T).reshape(-1, order="F")
        assert rel(M.prox_nuclear(x, lam, 50, 8), ref) < 3e-7


@pytest.mark.parametrize("dt", [np.float32, np.complex64])
@pytest.mark.parametrize("case", [((8, 8), (2, 2), 5, None, False), ((7, 9), (4, 4), 3, (1, 3), False), ((6,), (3,), 4, None, False),
                                  ((8, 8), (4, 4), 40, None, False), ((8, 8), (2, 2), 6, None, True), ((8, 4), (2, 2), 3, (1, 2), True),
                                  ((4, 4, 4), (2, 2, 2), 5, None, False), ((9,), (4,), 70, (2,), False)])
def test_device_algorithm_llr_equals_the_reference_loop(dt, case):
    shape, block, K, shift, overlapping = case
    rng = np.random.default_rng(9)
    x = _rnd(rng, int(np.prod(shape)) * K, dt)
    X = x.reshape(tuple(shape) + (K,), order="F")      # a view: weak patches fall under the λ >= ub shortcut
    X[: max(block[0], shape[0] // 2)] *= np.float32(0.02)
    lam = np.float32(0.6)
    ref = O.prox_llr(x.copy(), lam, shape, block, shift, overlapping)
    assert np.any(ref != 0) and (shift is not None or overlapping or np.any(ref == 0))
    assert rel(M.prox_llr(x, lam, shape, block, shift, overlapping), ref) < 1e-6
