"""Per-iterate parity against the oracle AT THE SIZES BASELINE.json NAMES ("full size where host RAM allows,
reduced-m twin for C3/C5").  Every matrix is generated on the device by the Philox generator the oracle shares
(oracle/philox.py; bit-identity is asserted on a row sample here and exhaustively at small sizes in test_gpu_kernels.py)
and downloaded once for the oracle, so both sides work on the very same entries.

  C1  CGNR + L2, ComplexF32 1024 x 4096, 50 iterations, U[0,1) entries (test/testSolvers.jl:25-27)          full size
  C2  FISTA + L1, Float32 16384 x 65536, 200 iterations                                                      full size
  C3  ADMM + TV(256 x 256), ComplexF32 m x 65536, rho = 0.1, iterationsCG = 10                               m = 2048 twin
  C4  multi-RHS FISTA-L1, 64 frames sharing ComplexF32 A 32768 x 16384, 50 iterations                        full size, 4 columns vs oracle
  C5  FISTA-L1 and CGNR + L2, ComplexF32 m x 65536 (the 16-CTA-cluster kernel)                               m = 4096 twin
Tolerance: rel-L2 <= 1e-5 per iterate (BASELINE.json north_star) and identical stopping decisions.  At these sizes the
Float32 oracle's own rounding reaches that tolerance (reductions over 65536 elements in OpenBLAS Float32; CG steering
scalars in CGNR and ADMM's inner cg! amplify it), so an iterate above 1e-5 must instead pass util.stepwise_vs_fp64's
criterion: the CUDA iterate is as close to the Float64 recurrence as the reference's own Float32 arithmetic is."""
import numpy as np
import pytest

import oracle as O
from util import rel, up64, stepwise_vs_fp64

pytestmark = pytest.mark.gpu
TOL = 1e-5


def check_philox_rows(A_host, dtype, m_global, n, seed, dist, scale, rows=(0, 777)):
    from oracle.philox import philox_matrix
    for r in rows:
        ref = philox_matrix(dtype, 2, n, seed, dist, scale, row_offset=r, m_global=m_global)
        assert np.array_equal(A_host[r:r + 2], ref), "device Philox matrix differs from the oracle's generator"


def test_c1_full_size_cgnr_l2(rls, ctx):
    """configs[0]: CGNR + L2Regularization(1f-3), ComplexF32 1024 x 4096, 50 iterations, relTol = 0 and default.
    The raw U[0,1) system is numerically chaotic in Float32 (the reference's own Float32 run leaves its Float64 run by
    percents within a few iterations), so every iterate is held to the Float64 criterion; the data residual is held to 1e-3."""
    from oracle.philox import UNIFORM01
    m, n, its = 1024, 4096, 50
    dtype = np.complex64
    Ad = rls.B200Matrix.philox(dtype, m, n, seed=12345, dist=0, ctx=ctx)
    A = Ad.to_numpy()
    check_philox_rows(A, dtype, m, n, 12345, UNIFORM01, 1.0)
    xt = rls.B200Vector(ctx, dtype, n).fill_philox(12345, stream=5, dist=0).to_numpy()
    b = (A @ xt).astype(dtype)
    lam = np.float32(1e-3)
    S = rls.CGNR(Ad, reg=rls.L2Regularization(lam), iterations=its, relTol=0.0)
    R32 = O.CGNR(A, reg=O.L2Regularization(lam), iterations=its, relTol=0.0)
    R64 = O.CGNR(up64(A), reg=O.L2Regularization(float(lam)), iterations=its, relTol=0.0)
    w = stepwise_vs_fp64(S, R32, R64, b, its)
    assert S.iteration == R32.iteration              # identical iteration counts (the Float32 recurrence may break down early)
    assert rel(A @ S.x, b) <= max(1e-3, 2 * rel(A @ R32.x, b))
    print(f"C1: {S.iteration} iterations; worst per-iterate gpu-o32 {w[0]:.2e}, gpu-o64 {w[1]:.2e}, o32-o64 {w[2]:.2e}; "
          f"{w[3]} iterates held to the Float64 criterion")
    # (the default relTol = eps(Float32) is decided at the Float32 noise floor of this system — ‖r‖/‖r₀‖ hovers around 1e-7
    # from iteration 14 on and crosses eps at a rounding-dependent iteration — so stopping decisions are tested with a
    # margin on the centred system, test_gpu_solvers.py::test_cgnr_c1_shape_and_stop, as SURVEY §7 hard part 2 prescribes)


def test_c2_full_size_fista_l1_per_iterate(rls, ctx):
    """configs[1]: FISTA + L1Regularization(1f-3), Float32 16384 x 65536, 200 iterations, rho = 0.95/lambda_max,
    against the oracle on the full system, every iterate (the oracle runs 15-20 iterations/s on the box's host cores)."""
    from oracle.philox import IH4
    m, n, its = 16384, 65536, 200
    dtype = np.float32
    scale = 1.0 / np.sqrt(m)
    Ad = rls.B200Matrix.philox(dtype, m, n, seed=12345, scale=scale, ctx=ctx)
    assert Ad.layout == "row"
    A = Ad.to_numpy()
    check_philox_rows(A, dtype, m, n, 12345, IH4, scale, rows=(0, 9000))
    xt = rls.B200Vector(ctx, dtype, n).fill_philox(12345, stream=11, dist=0).to_numpy()
    xt[np.arange(n) % 100 != 0] = 0
    noise = rls.B200Vector(ctx, dtype, m).fill_philox(12345, stream=12, dist=1, scale=1e-3).to_numpy()
    b = (A @ xt + noise).astype(dtype)
    AHA = rls.B200NormalOp(Ad, form="auto")
    assert "rowstream" in AHA.describe()
    b0 = rls.B200Vector(ctx, dtype, n).fill_philox(12345, stream=13, dist=1)
    rho = np.float32(0.95 / AHA.power_iterations(b0, rtol=1e-3, maxiter=30))
    lam = np.float32(1e-3)
    S = rls.FISTA(Ad, AHA=AHA, reg=rls.L1Regularization(lam), iterations=its, rho=rho, relTol=0.0)
    R = O.FISTA(A, reg=O.L1Regularization(lam), iterations=its, rho=rho, relTol=0.0)
    R64 = O.FISTA(up64(A), reg=O.L1Regularization(float(lam)), iterations=its, rho=float(rho), relTol=0.0)

    def each(k):
        # ‖res‖_k / ‖A'b‖ (FISTA.jl:156).  res = A'A x - A'b is a difference of vectors ~200 x its own size late in the solve,
        # so the bound is on the error of res relative to what is subtracted, i.e. absolute in this normalised quantity
        assert abs(S._scalars.rel_res_norm - float(R.rel_res_norm)) <= 2e-5, (k, S._scalars.rel_res_norm, float(R.rel_res_norm))
    w = stepwise_vs_fp64(S, R, R64, b, its, each=each)
    assert S.iteration == R.iteration == its
    print(f"C2: worst per-iterate over {its} iterations gpu-o32 {w[0]:.2e}, gpu-o64 {w[1]:.2e}, o32-o64 {w[2]:.2e}; "
          f"{w[3]} iterates held to the Float64 criterion")
    # the whole callback-free solve! (two fused kernels per iteration) lands on the very same iterate
    x_step = S.x.copy()
    assert np.array_equal(rls.solve_(S, b), x_step)


def test_c4_full_size_batched_columns_vs_oracle(rls, ctx):
    """configs[3]: 64 frames sharing one ComplexF32 A 32768 x 16384, FISTA-L1, 50 iterations — the tensor-core batched
    solve at full size, four of its columns against the oracle's single-right-hand-side solve."""
    m, n, K, its = 32768, 16384, 64, 50
    dtype = np.complex64
    scale = 1.0 / np.sqrt(m)
    Ad = rls.B200Matrix.philox(dtype, m, n, seed=4321, scale=scale, ctx=ctx)
    A = Ad.to_numpy()
    X = np.zeros((n, K), dtype, order="F")
    rng = np.random.default_rng(4)
    for k in range(K):
        idx = rng.integers(0, n, 160)
        X[idx, k] = (rng.random(160) + 1j * rng.random(160)).astype(dtype)
    B = np.asfortranarray((A @ X).astype(dtype))
    AHA = rls.B200NormalOp(Ad, form="auto")
    b0 = rls.B200Vector(ctx, dtype, n).fill_philox(9, stream=1, dist=1)
    rho = np.float32(0.95 / AHA.power_iterations(b0))
    lam = np.float32(1e-3)
    S = rls.FISTA(Ad, AHA=AHA, reg=rls.L1Regularization(lam), iterations=its, rho=rho, relTol=0.0)
    Xs = rls.solve_(S, B)
    worst = 0.0
    for k in (0, 21, 42, 63):
        xr = O.FISTA(A, reg=O.L1Regularization(lam), iterations=its, rho=rho, relTol=0.0).solve(B[:, k].copy())
        e = rel(Xs[:, k], xr)
        worst = max(worst, e)
        assert e < TOL, f"column {k}: rel-L2 {e:.3e}"
    print(f"C4: worst column rel-L2 vs oracle {worst:.2e}")


def c5_twin(rls, ctx, m):
    from oracle.philox import IH4
    n = 65536
    dtype = np.complex64
    scale = 1.0 / np.sqrt(m)
    Ad = rls.B200Matrix.philox(dtype, m, n, seed=12345, scale=scale, ctx=ctx)
    A = Ad.to_numpy()
    check_philox_rows(A, dtype, m, n, 12345, IH4, scale, rows=(0, m - 2))
    xt = rls.B200Vector(ctx, dtype, n).fill_philox(12345, stream=11, dist=0).to_numpy()
    xt[np.arange(n) % 100 != 0] = 0
    b = (A @ xt).astype(dtype)
    return Ad, A, b


def test_c5_twin_fista_l1_and_cgnr(rls, ctx):
    """configs[4] at reduced m, n = 65536 ComplexF32 kept: rows of 131072 floats, i.e. the 16-CTA-cluster kernel that
    every shard of the 262144 x 65536 system runs."""
    m, n = 4096, 65536
    Ad, A, b = c5_twin(rls, ctx, m)
    AHA = rls.B200NormalOp(Ad, form="auto")
    assert "x 16 CTAs" in AHA.describe()
    b0 = rls.B200Vector(ctx, np.complex64, n).fill_philox(12345, stream=13, dist=1)
    rho = np.float32(0.95 / AHA.power_iterations(b0, maxiter=10))
    lam = np.float32(1e-3)
    S = rls.FISTA(Ad, AHA=AHA, reg=rls.L1Regularization(lam), iterations=40, rho=rho, relTol=0.0)
    R = O.FISTA(A, reg=O.L1Regularization(lam), iterations=40, rho=rho, relTol=0.0)
    R64 = O.FISTA(up64(A), reg=O.L1Regularization(float(lam)), iterations=40, rho=float(rho), relTol=0.0)
    w = stepwise_vs_fp64(S, R, R64, b, 40)
    print(f"C5 twin FISTA-L1: worst per-iterate gpu-o32 {w[0]:.2e}, gpu-o64 {w[1]:.2e}, o32-o64 {w[2]:.2e}; {w[3]} iterates held to "
          "the Float64 criterion")
    S = rls.CGNR(Ad, AHA=AHA, reg=rls.L2Regularization(lam), iterations=25, relTol=0.0)
    R32 = O.CGNR(A, reg=O.L2Regularization(lam), iterations=25, relTol=0.0)
    R64 = O.CGNR(up64(A), reg=O.L2Regularization(float(lam)), iterations=25, relTol=0.0)
    w = stepwise_vs_fp64(S, R32, R64, b, 25)
    print(f"C5 twin CGNR: worst per-iterate gpu-o32 {w[0]:.2e}, gpu-o64 {w[1]:.2e}, o32-o64 {w[2]:.2e}; {w[3]} iterates held to the "
          "Float64 criterion")


def test_c3_twin_admm_tv(rls, ctx):
    """configs[2] at reduced m: ADMM + TVRegularization(1f-2; shape = (256, 256)) on a ComplexF32 m x 65536 system,
    rho = 0.1, iterationsCG = 10: identical inner-CG counts and stopping decisions, every outer iterate against the oracle."""
    m, n, outer = 2048, 65536, 6
    dtype = np.complex64
    scale = 1.0 / np.sqrt(m)
    Ad = rls.B200Matrix.philox(dtype, m, n, seed=1234, scale=scale, ctx=ctx)
    A = Ad.to_numpy()
    img = np.zeros((256, 256), dtype, order="F")
    rng = np.random.default_rng(1234)
    for _ in range(5):                                    # piecewise-constant image, test/testProxMaps.jl:78-83
        i, j = rng.integers(0, 256, 2)
        img[i:, j:] += np.float32(rng.standard_normal())
    b = (A @ img.ravel(order="F")).astype(dtype)
    kw = dict(rho=0.1, iterations=outer, iterationsCG=10, absTol=0.0, relTol=0.0)
    S = rls.ADMM(Ad, reg=rls.TVRegularization(np.float32(1e-2), shape=(256, 256)), **kw)
    R32 = O.ADMM(A, reg=O.TVRegularization(np.float32(1e-2), shape=(256, 256)), **kw)
    R64 = O.ADMM(up64(A), reg=O.TVRegularization(float(np.float32(1e-2)), shape=(256, 256)), **kw)

    def each(k):
        assert S._scalars.cg_iterations_last == R32.cg_iters[-1], f"inner CG count differs at outer {k}"
    w = stepwise_vs_fp64(S, R32, R64, b, outer, each=each)
    assert S.iteration == outer
    print(f"C3 twin ADMM+TV: worst per-iterate gpu-o32 {w[0]:.2e}, gpu-o64 {w[1]:.2e}, o32-o64 {w[2]:.2e}; {w[3]} iterates held to "
          "the Float64 criterion")
