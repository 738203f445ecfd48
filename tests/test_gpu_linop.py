"""SURVEY §8(f) rank 4 on the GPU: matrix-free system operators (csrc/rls_linop.cu) — SamplingOp, FFTOp (cuFFT), their
product, and a caller-supplied AHA callback — through the C ABI against NumPy, and the solvers running on them against
the oracle running on the same operator written out as a dense matrix (the way test/testSolvers.jl:67-82 builds it)."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from util import rel

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rnd(rng, n, dt=np.complex64):
    return (rng.standard_normal(n) + (1j * rng.standard_normal(n) if np.dtype(dt).kind == "c" else 0)).astype(dt)


def _fft_ref(x, shape, shift=True, unitary=True, adjoint=False):
    X = x.reshape(shape, order="F").astype(np.complex128)
    if shift:
        X = np.fft.ifftshift(X)
    X = np.fft.ifftn(X) * X.size if adjoint else np.fft.fftn(X)
    if shift:
        X = np.fft.fftshift(X)
    if unitary:
        X = X / np.sqrt(X.size)
    return X.reshape(-1, order="F")


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_sampling_op(rls, ctx, dtype):
    rng = np.random.default_rng(1)
    shape = (13, 9)
    n = 117
    pattern = np.sort(rng.permutation(n)[: n // 3]) + 1          # compressed_sensing.jl:19-20
    A = rls.SamplingOp(dtype, pattern=pattern, shape=shape, ctx=ctx)
    assert A.shape == (n // 3, n)
    x, y = _rnd(rng, n, dtype), _rnd(rng, n // 3, dtype)
    assert np.array_equal(A * x, x[pattern - 1])
    back = np.zeros(n, dtype); back[pattern - 1] = y
    assert np.array_equal(A.tmul(y), back)
    mask = np.zeros(n, bool); mask[pattern - 1] = True
    AHA = A.normal()
    assert AHA.form == "matrixfree" and "SamplingOp" in AHA.describe()
    assert np.array_equal(AHA.apply(rls.B200Vector.from_numpy(x, ctx)).to_numpy(), x * mask)
    with pytest.raises(rls.RlsError, match="twice"):
        rls.SamplingOp(dtype, pattern=[1, 2, 2], shape=shape, ctx=ctx)
    with pytest.raises(rls.RlsError, match="outside"):
        rls.SamplingOp(dtype, pattern=[0, 1], shape=shape, ctx=ctx)


@pytest.mark.parametrize("shape", [(256,), (15,), (16, 12), (9, 7), (8, 6, 5), (1, 32), (64, 1, 3)])
@pytest.mark.parametrize("shift,unitary", [(True, True), (False, True), (True, False)])
def test_fft_op(rls, ctx, shape, shift, unitary):
    rng = np.random.default_rng(2)
    n = int(np.prod(shape))
    F = rls.FFTOp(np.complex64, shape=shape, shift=shift, unitary=unitary, ctx=ctx)
    x, y = _rnd(rng, n), _rnd(rng, n)
    assert rel(F * x, _fft_ref(x, shape, shift, unitary)) < 2e-6
    assert rel(F.tmul(y), _fft_ref(y, shape, shift, unitary, adjoint=True)) < 2e-6
    lhs, rhs = np.vdot(y, (F * x).astype(np.complex128)), np.vdot(F.tmul(y).astype(np.complex128), x)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)                      # <y, A x> = <A' y, x>
    if unitary:
        assert rel(F.normal().apply(rls.B200Vector.from_numpy(x, ctx)).to_numpy(), x) < 2e-6


def _cs_problem(N=256, seed=1):
    """test/testSolvers.jl:67-82: sparse x, unitary DFT, a random half of the rows"""
    rng = np.random.default_rng(seed)
    Fd = np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(N)) / N) / np.sqrt(N)
    x = np.zeros(N)
    for _ in range(3):
        x[rng.integers(0, N)] = rng.random()
    idx = np.sort(np.unique(rng.integers(0, N, N // 2)))
    return x, idx, Fd[idx, :].astype(np.complex64), (Fd @ x)[idx].astype(np.complex64)


@pytest.mark.parametrize("solver", ["FISTA", "POGM", "OptISTA", "ADMM", "CGNR"])
def test_solvers_on_undersampled_fourier_operator(rls, ctx, solver):
    """The compressed-sensing problem of test/testSolvers.jl:67-125 with A = SamplingOp * FFTOp never stored, per iterate
    against the oracle on the dense 128x256 matrix of the same operator."""
    N = 256
    x, idx, Fd, b = _cs_problem(N)
    A = rls.SamplingOp(np.complex64, pattern=idx + 1, shape=(N,), ctx=ctx) * rls.FFTOp(np.complex64, shape=(N,), shift=False, ctx=ctx)
    assert A.shape == Fd.shape
    rng = np.random.default_rng(3)
    v = _rnd(rng, N)
    assert rel(A * v, Fd.astype(np.complex128) @ v) < 2e-6
    lam = np.float32(1e-3)
    its = 60
    if solver == "CGNR":
        # the rows of a unitary DFT are orthonormal (A A' = I): CG converges in one step and the residual then decays to an
        # exact Float32 zero (iteration 9 in the oracle), where `‖r‖/z0 <= relTol` stops at a rounding-dependent iteration
        kw, okw = dict(iterations=4, relTol=0.0), dict(iterations=4, relTol=0.0)
        reg, oreg = rls.L2Regularization(lam), O.L2Regularization(lam)
        its = 4
    elif solver == "ADMM":
        kw, okw = dict(iterations=30, rho=0.1), dict(iterations=30, rho=0.1)
        reg, oreg = rls.L1Regularization(lam), O.L1Regularization(lam)
        its = 30
    else:
        kw = okw = dict(iterations=its, rho=np.float32(0.95), relTol=0.0)
        reg, oreg = rls.L1Regularization(lam), O.L1Regularization(lam)
    S = rls.createLinearSolver(getattr(rls, solver), A, reg=reg, **kw)
    R = O.createLinearSolver(getattr(O, solver), Fd, reg=oreg, **okw)
    R64 = O.createLinearSolver(getattr(O, solver), Fd.astype(np.complex128), reg=type(oreg)(float(lam)),
                               **{k: (float(w) if isinstance(w, np.floating) else w) for k, w in okw.items()})
    S.init_(b); R.init(b); R64.init(b.astype(np.complex128))
    for k in range(its + 2):
        a, r1 = S.iterate(), R.iterate()
        R64.iterate()
        assert a == r1, (k, a, r1)
        if not a:
            break
        e = rel(S.x, R.x)
        assert e < TOL or rel(S.x, R64.x) <= 1.5 * rel(R.x, R64.x), (k, e, rel(S.x, R64.x), rel(R.x, R64.x))
    assert S.iteration == R.iteration
    xs = rls.solve_(S, b)                                         # whole solve, host buffers
    assert rel(xs, R.x) < 2e-5 or rel(xs, R64.x) <= 1.5 * rel(R.x, R64.x)
    if solver != "CGNR":
        long = rls.createLinearSolver(getattr(rls, solver), A, reg=reg, **dict(kw, iterations=200))
        assert np.linalg.norm(x - rls.solve_(long, b)) <= 0.1 * np.linalg.norm(x)     # the reference's acceptance bound (:91)


def test_compressed_sensing_example_tv_fista(rls, ctx):
    """docs/src/literate/examples/compressed_sensing.jl: a third of the pixels of a piecewise-constant image, TV-regularised
    FISTA, 20 iterations (on a 48x48 phantom so that the oracle can hold the sampling matrix)."""
    N = 48
    rng = np.random.default_rng(4)
    img = np.zeros((N, N), np.float32)
    for _ in range(5):
        i, j = rng.integers(0, N, 2)
        img[i:, j:] += np.float32(rng.random())
    pattern = np.sort(rng.permutation(N * N)[: N * N // 3]) + 1
    A = rls.SamplingOp(np.float32, pattern=pattern, shape=(N, N), ctx=ctx)
    b = A * img.reshape(-1, order="F")
    Ad = np.zeros((pattern.size, N * N), np.float32)
    Ad[np.arange(pattern.size), pattern - 1] = 1
    lam = np.float32(0.01)
    kw = dict(iterations=20, rho=np.float32(0.95), relTol=0.0)
    S = rls.createLinearSolver(rls.FISTA, A, reg=rls.TVRegularization(lam, shape=(N, N)), **kw)
    R = O.createLinearSolver(O.FISTA, Ad, reg=O.TVRegularization(lam, shape=(N, N)), **kw)
    xs, xr = rls.solve_(S, b), O.solve_(R, b)
    assert S.iteration == R.iteration == 20
    assert rel(xs, xr) < 2e-5, rel(xs, xr)
    assert np.linalg.norm(xs - img.reshape(-1, order="F")) < np.linalg.norm(A.tmul(b) - img.reshape(-1, order="F"))


def test_normal_operator_as_a_callback(rls, ctx):
    """rls_normal_from_callback: AHA is the caller's function on device pointers and the solver's stream.  Here the callback
    is a device-to-device copy (AHA = I) issued with the CUDA runtime; the solve must equal the one on the identity matrix."""
    import ctypes.util
    rt = None
    for name in ("libcudart.so.12", "libcudart.so", ctypes.util.find_library("cudart")):
        try:
            rt = C.CDLL(name)
            break
        except (OSError, TypeError):
            continue
    if rt is None:
        import glob
        cands = glob.glob("/usr/local/cuda*/lib64/libcudart.so*") + glob.glob("/opt/**/libcudart.so*", recursive=True)
        if not cands:
            pytest.skip("no libcudart to issue the device copy from Python")
        rt = C.CDLL(cands[0])
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    n = 3000
    calls = []

    def aha(x_ptr, res_ptr, stream):
        calls.append(stream)
        return rt.cudaMemcpyAsync(res_ptr, x_ptr, n * 8, 3, stream)          # cudaMemcpyDeviceToDevice

    op = rls.B200NormalOp.from_callback(aha, np.complex64, n, ctx=ctx)
    assert op.form == "matrixfree"
    rng = np.random.default_rng(6)
    b = _rnd(rng, n)
    assert np.array_equal(op.apply(rls.B200Vector.from_numpy(b, ctx)).to_numpy(), b)
    lam = np.float32(0.3)
    kw = dict(reg=rls.L1Regularization(lam), iterations=15, rho=np.float32(0.5), relTol=0.0)
    S = rls.FISTA(None, AHA=op, ctx=ctx, **kw)
    S1 = rls.FISTA(np.eye(n, dtype=np.complex64), normal="twopass", ctx=ctx, **kw)
    xs, x1 = rls.solve_(S, b), rls.solve_(S1, b)
    assert S.iteration == S1.iteration == 15 and len(calls) >= 16
    assert rel(xs, x1) < 1e-6
    bad = rls.B200NormalOp.from_callback(lambda x, r, s: 7, np.complex64, n, ctx=ctx)
    with pytest.raises(rls.RlsError, match="callback returned 7"):
        bad.apply(rls.B200Vector.from_numpy(b, ctx))
