"""Timing of the SURVEY 8(f) rank-4 pieces on one B200: singular-value thresholding (device-resident vector, CUDA events)
against the oracle's LAPACK loop on the host cores (what the reference does: ProxLLR.jl:7 'always performed on the CPU'),
and a TV-FISTA iteration on the matrix-free SamplingOp * FFTOp.  One JSON line per case."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
import oracle as O

ctx = rls.B200Context.default(0)
rng = np.random.default_rng(0)


def rnd(n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def time_prox(reg, x, reps=20, **kw):
    v = rls.B200Vector.from_numpy(x, ctx)
    for _ in range(3):
        v.upload(x); rls.prox_(reg, v, **kw)
    ctx.sync()
    ms = []
    for _ in range(reps):
        v.upload(x)
        ctx.sync()
        ctx.timer_start()
        rls.prox_(reg, v, **kw)
        ms.append(ctx.timer_stop())
    return float(np.median(ms)), v.to_numpy()


def case_llr(shape, block, K, overlapping=False, lam=0.6):
    n = int(np.prod(shape)) * K
    x = rnd(n)
    reg = rls.LLRRegularization(np.float32(lam), shape=shape, blockSize=block, randshift=False, fullyOverlapping=overlapping)
    ms, got = time_prox(reg, x)
    t0 = time.perf_counter()
    ref = O.prox_llr(x.copy(), np.float32(lam), shape, block, None, overlapping)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    passes = int(np.prod(block)) if overlapping else 1
    print(json.dumps({"case": f"LLRRegularization {shape}x{K} ComplexF32, blockSize {block}, fullyOverlapping={overlapping}",
                      "patches": int(np.prod([-(-s // b) for s, b in zip(shape, block)])), "short_side": min(K, int(np.prod(block))),
                      "gpu_ms": ms, "cpu_lapack_loop_ms": cpu_ms, "speedup": cpu_ms / ms,
                      "streaming_bytes": 3 * n * 8 * passes, "streaming_gbs": 3 * n * 8 * passes / ms / 1e6,
                      "rel_l2_vs_lapack": float(np.linalg.norm(got - ref) / np.linalg.norm(ref))}), flush=True)


def case_nuclear(rows, cols, lam_frac=0.3):
    x = rnd(rows * cols)
    lam = np.float32(lam_frac * np.linalg.svd(x.reshape((rows, cols), order="F"), compute_uv=False)[0])
    ms, got = time_prox(rls.NuclearRegularization(lam, svtShape=(rows, cols)), x)
    t0 = time.perf_counter()
    ref = O.prox_nuclear(x.copy(), lam, (rows, cols))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"case": f"NuclearRegularization svtShape ({rows}, {cols}) ComplexF32", "gpu_ms": ms, "cpu_lapack_ms": cpu_ms,
                      "speedup": cpu_ms / ms, "streaming_bytes": 3 * rows * cols * 8, "streaming_gbs": 3 * rows * cols * 8 / ms / 1e6,
                      "rel_l2_vs_lapack": float(np.linalg.norm(got - ref) / np.linalg.norm(ref))}), flush=True)


def case_cs(N=256, its=50):
    idx = np.sort(rng.permutation(N * N)[: N * N // 3]) + 1
    op = rls.SamplingOp(np.complex64, pattern=idx, shape=(N, N), ctx=ctx) * rls.FFTOp(np.complex64, shape=(N, N), ctx=ctx)
    img = np.zeros((N, N), np.complex64)
    for _ in range(5):
        i, j = rng.integers(0, N, 2)
        img[i:, j:] += np.float32(rng.random())
    b = op * img.reshape(-1, order="F")
    S = rls.createLinearSolver(rls.FISTA, op, reg=rls.TVRegularization(np.float32(1e-2), shape=(N, N)), iterations=its,
                               rho=np.float32(0.95), relTol=0.0)
    bd = op.tmul(rls.B200Vector.from_numpy(b, ctx))
    import ctypes as C
    it = C.c_int32()
    call = lambda: rls._capi.call("rls_solver_solve", S._handle, bd.handle, None, C.byref(it), C.byref(S._scalars))
    call(); ctx.sync()
    ctx.timer_start()
    for _ in range(5):
        call()
    ms = ctx.timer_stop() / 5
    x = S.x
    print(json.dumps({"case": f"FISTA + TVRegularization on SamplingOp * FFTOp, {N}x{N} image, a third of k-space, {its} iterations",
                      "ms_per_solve": ms, "us_per_iteration": 1e3 * ms / its, "iterations": it.value,
                      "normal_operator": S.AHA.describe(),
                      "rel_err_vs_image": float(np.linalg.norm(x - img.reshape(-1, order="F")) / np.linalg.norm(img))}), flush=True)


case_llr((256, 256), (4, 4), 16)
case_llr((256, 256), (8, 8), 32)
case_llr((128, 128), (4, 4), 80)
case_llr((128, 128), (4, 4), 16, overlapping=True)
case_llr((32, 32, 32), (4, 4, 4), 80)
case_nuclear(65536, 16)
case_nuclear(1 << 20, 32)
case_nuclear(64, 65536)
case_cs()
