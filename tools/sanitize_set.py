"""The smoke set run under compute-sanitizer (tools/sanitize.sh): every kernel family once, at sizes the tool finishes in
minutes — the cluster/DSMEM exchange of the streaming row kernel (clusters of 1, 2, 5, 8 and 16 CTAs), the fused FISTA
iteration, CGNR, ADMM + TV, the tcgen05 multi-RHS GEMMs and Gram build, the Kaczmarz sweep with its flag-in-data exchange."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import rls_b200 as rls
from util import rel, rand_matrix, rand_vector, sparse_truth

ctx = rls.B200Context.default(0)
only = set(sys.argv[1:])


def want(name):
    return not only or name in only


if want("rows"):
    for dtype, m, n in ((np.float32, 40, 65536), (np.complex64, 24, 65536), (np.complex64, 33, 20000), (np.float32, 64, 300),
                        (np.float32, 100, 16384), (np.complex64, 7, 4099)):
        A, _ = rand_matrix(dtype, m, n, 5)
        x = rand_vector(dtype, n, 6)
        y = rand_vector(dtype, m, 7)
        Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
        A64 = A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
        e1 = rel(Ad.mul(rls.B200Vector.from_numpy(x, ctx)).to_numpy(), A64 @ x)
        e2 = rel(Ad.adjoint_mul(rls.B200Vector.from_numpy(y, ctx)).to_numpy(), A64.conj().T @ y)
        op = rls.B200NormalOp(Ad, form="onepass")
        e3 = rel(op.apply(rls.B200Vector.from_numpy(x, ctx)).to_numpy(), A64.conj().T @ (A64 @ x))
        assert max(e1, e2, e3) < 2e-6, (m, n, e1, e2, e3)
        print(f"rows {np.dtype(dtype).name} {m}x{n}: {e1:.1e} {e2:.1e} {e3:.1e} [{op.describe()[:70]}]", flush=True)

if want("solvers"):
    dtype = np.complex64
    A, _ = rand_matrix(dtype, 96, 2048, 11)
    b = (A @ sparse_truth(dtype, 2048, 12)).astype(dtype)
    Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
    x = rls.solve_(rls.FISTA(Ad, reg=rls.L1Regularization(np.float32(1e-2)), iterations=5, rho=np.float32(0.02), relTol=0.0), b)
    x = rls.solve_(rls.CGNR(Ad, reg=rls.L2Regularization(np.float32(1e-3)), iterations=5, relTol=0.0), b)
    x = rls.solve_(rls.POGM(Ad, reg=rls.L1Regularization(np.float32(1e-2)), iterations=4, rho=np.float32(0.02), relTol=0.0, restart="gradient"), b)
    A2, _ = rand_matrix(dtype, 64, 256, 13)
    b2 = (A2 @ sparse_truth(dtype, 256, 14)).astype(dtype)
    x = rls.solve_(rls.ADMM(A2, reg=rls.TVRegularization(np.float32(1e-2), shape=(16, 16)), iterations=2, iterationsCG=3, rho=0.1), b2)
    x = rls.solve_(rls.SplitBregman(A2, reg=rls.L1Regularization(np.float32(1e-2)), iterations=2, iterationsInner=2, iterationsCG=3), b2)
    print("solvers ok", flush=True)

if want("tensor"):
    dtype = np.complex64
    A, _ = rand_matrix(dtype, 256, 128, 21)
    Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
    B = np.stack([rand_vector(dtype, 256, 30 + k) for k in range(8)], axis=1)
    X = rls.solve_(rls.FISTA(Ad, reg=rls.L1Regularization(np.float32(1e-3)), iterations=3, rho=np.float32(0.1), relTol=0.0), B)
    G = rls.B200NormalOp(Ad, form="gram")
    g = G.apply(rls.B200Vector.from_numpy(rand_vector(dtype, 128, 40), ctx)).to_numpy()
    print("tensor-core paths ok", flush=True)

if want("kaczmarz"):
    A, _ = rand_matrix(np.float32, 128, 4096, 31)
    b = rand_vector(np.float32, 128, 32)
    Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
    x = rls.solve_(rls.Kaczmarz(Ad, reg=rls.L2Regularization(np.float32(1e-2)), iterations=2), b)
    print("kaczmarz ok", flush=True)

if want("svt"):
    # singular-value thresholding (rls_svt.cu): both views (short side = columns / rows), q up to 64, LLR patches with a
    # shift, border patches and the fully overlapping variant; Gram-form batched apply on the tensor cores
    for dtype in (np.float32, np.complex64):
        for shp in ((70, 9), (9, 70), (130, 64)):
            x = rand_vector(dtype, shp[0] * shp[1], 50)
            rls.prox_(rls.NuclearRegularization(np.float32(3.0), svtShape=shp), x.copy())
        x = rand_vector(dtype, 9 * 7 * 5, 51)
        rls.prox_(rls.LLRRegularization(np.float32(0.8), shape=(9, 7), blockSize=(4, 4), randshift=False), x.copy(), shift=(1, 3))
        x = rand_vector(dtype, 8 * 8 * 70, 52)
        rls.prox_(rls.LLRRegularization(np.float32(0.8), shape=(8, 8), blockSize=(2, 2), randshift=False, fullyOverlapping=True), x.copy())
    os.environ["RLS_BATCH_MIN_K"] = "2"
    A, _ = rand_matrix(np.complex64, 200, 128, 53)
    G = rls.B200NormalOp(rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row"), form="gram")
    xs = [rls.B200Vector.from_numpy(rand_vector(np.complex64, 128, 60 + k), ctx) for k in range(5)]
    G.apply_batch(xs)
    print("svt + gram batch ok", flush=True)

if want("linop"):
    N = 96
    idx = np.arange(1, N * N + 1)[::3]
    op = rls.SamplingOp(np.complex64, pattern=idx, shape=(N, N), ctx=ctx) * rls.FFTOp(np.complex64, shape=(N, N), ctx=ctx)
    b = rand_vector(np.complex64, idx.size, 70)
    rls.solve_(rls.FISTA(op, reg=rls.TVRegularization(np.float32(1e-2), shape=(N, N)), iterations=3, rho=np.float32(0.9), relTol=0.0), b)
    print("linop ok", flush=True)
