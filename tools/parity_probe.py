"""Where do the per-iterate differences between the CUDA path and the Float32 oracle come from?  For every solver family
run three things on the same inputs — the CUDA path, the oracle in Float32 (the reference's arithmetic), the oracle in
Float64 — and print, per case, the worst per-iterate rel-L2 of  gpu vs oracle32,  gpu vs oracle64,  oracle32 vs oracle64.
If gpu-vs-64 <= oracle32-vs-64 the CUDA iterate is at least as close to the exact recurrence as the reference's own
Float32 run, and the gpu-vs-oracle32 distance is the oracle's BLAS rounding, not an error of the CUDA path.
usage: python tools/parity_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle as O
import rls_b200 as rls
from util import rel, rand_matrix, rand_vector, sparse_truth

ctx = rls.B200Context.default(0)


def up(a):
    return a.astype(np.complex128 if np.asarray(a).dtype.kind == "c" else np.float64)


def problem(dtype, m, n, seed=100, noise=1e-3):
    A, _ = rand_matrix(dtype, m, n, seed)
    xt = sparse_truth(dtype, n, seed + 1)
    b = (A @ xt + noise * rand_vector(dtype, m, seed + 2)).astype(dtype)
    return A, xt, b


def three_way(name, S, R32, R64, b, iters):
    S.init_(b); R32.init(b); R64.init(up(b))
    w = [0.0, 0.0, 0.0]
    worst_ratio = 0.0
    for k in range(iters):
        a = S.iterate(); r1 = R32.iterate(); r2 = R64.iterate()
        if not (a and r1 and r2):
            break
        x, x32, x64 = S.x, R32.x, R64.x
        e = (rel(x, x32), rel(x, x64), rel(x32, x64))
        for q in range(3):
            w[q] = max(w[q], e[q])
        worst_ratio = max(worst_ratio, e[1] / max(e[2], 1e-12))
    print(f"{name:44s} gpu-o32 {w[0]:.2e}  gpu-o64 {w[1]:.2e}  o32-o64 {w[2]:.2e}  worst per-iterate (gpu-o64)/(o32-o64) {worst_ratio:.2f}", flush=True)


for dtype in (np.float32, np.complex64):
    tn = np.dtype(dtype).name
    for lam in (np.float32(0), np.float32(1e-3)):
        A, xt, b = problem(dtype, 512, 256)
        three_way(f"CGNR 512x256 {tn} lam={lam}", rls.CGNR(A, reg=rls.L2Regularization(lam), iterations=30, relTol=0.0, normal="twopass"),
                  O.CGNR(A, reg=O.L2Regularization(lam), iterations=30, relTol=0.0),
                  O.CGNR(up(A), reg=O.L2Regularization(float(lam)), iterations=30, relTol=0.0), b, 30)
    shape = (24, 20)
    A, xt, b = problem(dtype, 320, shape[0] * shape[1])
    kw = dict(iterations=15, iterationsCG=10, rho=0.1, absTol=0.0, relTol=0.0)
    for variant in ("l1", "tv_identity", "l1_gradient"):
        def mk(M, dt):
            if variant == "l1":
                return dict(reg=M.L1Regularization(np.float32(1e-2) if dt == dtype else 1e-2))
            if variant == "tv_identity":
                return dict(reg=M.TVRegularization(np.float32(1e-2) if dt == dtype else 1e-2, shape=shape))
            return dict(reg=M.L1Regularization(np.float32(1e-2) if dt == dtype else 1e-2), regTrafo=M.GradientOp(dt, shape))
        for solver in ("ADMM", "SplitBregman"):
            kws = dict(kw)
            if solver == "SplitBregman":
                kws = dict(iterations=3, iterationsInner=5, iterationsCG=10, rho=0.1, absTol=0.0, relTol=0.0)
            dt64 = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
            try:
                three_way(f"{solver} {variant} {tn}", getattr(rls, solver)(A, normal="twopass", **kws, **mk(rls, dtype)),
                          getattr(O, solver)(A, **kws, **mk(O, dtype)), getattr(O, solver)(up(A), **kws, **mk(O, dt64)), b, 15)
            except Exception as e:  # noqa: BLE001
                print(f"{solver} {variant} {tn}: EXC {e}", flush=True)
    A, xt, b = problem(dtype, 384, 1024)
    rho = np.float32(0.95 / np.linalg.norm(up(A), 2) ** 2)
    for solver in ("FISTA", "POGM", "OptISTA"):
        three_way(f"{solver}-L1 100 its {tn}", getattr(rls, solver)(A, reg=rls.L1Regularization(np.float32(2e-2)), iterations=100, rho=rho, relTol=0.0),
                  getattr(O, solver)(A, reg=O.L1Regularization(np.float32(2e-2)), iterations=100, rho=rho, relTol=0.0),
                  getattr(O, solver)(up(A), reg=O.L1Regularization(2e-2), iterations=100, rho=float(rho), relTol=0.0), b, 100)

# C1 at full size: CGNR + L2 on U[0,1) ComplexF32 1024 x 4096
from oracle.philox import philox_matrix, philox_vector, UNIFORM01
A = philox_matrix(np.complex64, 1024, 4096, 12345, UNIFORM01, 1.0)
xt = philox_vector(np.complex64, 4096, 12345, 5, UNIFORM01)
b = (A @ xt).astype(np.complex64)
lam = np.float32(1e-3)
three_way("C1 CGNR+L2 1024x4096 complex64 U[0,1)", rls.CGNR(A, reg=rls.L2Regularization(lam), iterations=50, relTol=0.0),
          O.CGNR(A, reg=O.L2Regularization(lam), iterations=50, relTol=0.0),
          O.CGNR(up(A), reg=O.L2Regularization(1e-3), iterations=50, relTol=0.0), b, 50)
