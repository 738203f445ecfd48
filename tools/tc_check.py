"""Tensor-core paths (csrc/rls_tc.cu): batched normal operator and Gram build against NumPy float64,
plus timing at the C4 shape.  usage: python tools/tc_check.py [check] [time]"""
import os
import sys

os.environ.setdefault("RLS_BATCH_MIN_K", "2")

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls

ctx = rls.B200Context.default(0)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rand(shape, dtype, rng):
    a = rng.standard_normal(shape).astype(np.float32)
    if np.dtype(dtype).kind == "c":
        a = (a + 1j * rng.standard_normal(shape)).astype(np.complex64)
    return a


def check(m, n, K, dtype):
    rng = np.random.default_rng(m * 7 + n * 3 + K)
    A, X = rand((m, n), dtype, rng), rand((n, K), dtype, rng)
    Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
    op = rls.B200NormalOp(Ad, form="onepass")
    xs = [rls.B200Vector.from_numpy(np.ascontiguousarray(X[:, k]), ctx) for k in range(K)]
    outs = op.apply_batch(xs)
    G = np.stack([o.to_numpy() for o in outs], axis=1)
    A64 = A.astype(np.complex128 if A.dtype.kind == "c" else np.float64)
    ref = A64.conj().T @ (A64 @ X)
    single = np.stack([op.apply(x).to_numpy() for x in xs], axis=1)
    e, e1 = rel(G, ref), rel(single, ref)
    worst = max(rel(G[:, k], ref[:, k]) for k in range(K))
    ok = worst < 3e-6
    print(f"{'ok ' if ok else 'BAD'} batch {np.dtype(dtype).name:9s} {m:6d}x{n:<6d} K={K:3d} rel {e:.2e} (worst column {worst:.2e}; CUDA-core single applies {e1:.2e})", flush=True)
    return ok


def check_gram(m, n, dtype):
    rng = np.random.default_rng(m + n)
    A = rand((m, n), dtype, rng)
    Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
    op = rls.B200NormalOp(Ad, form="gram")
    x = rand((n,), dtype, rng)
    g = op.apply(rls.B200Vector.from_numpy(x, ctx)).to_numpy()
    A64 = A.astype(np.complex128 if A.dtype.kind == "c" else np.float64)
    ref = (A64.conj().T @ A64) @ x
    e = rel(g, ref)
    ok = e < 3e-6
    print(f"{'ok ' if ok else 'BAD'} gram  {np.dtype(dtype).name:9s} {m:6d}x{n:<6d} rel {e:.2e}  [{op.describe()}]", flush=True)
    return ok


def time_batch(m, n, K, dtype, reps=5):
    dtype = np.dtype(dtype)
    A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
    op = rls.B200NormalOp(A, form="onepass")
    xs = [rls.B200Vector(ctx, dtype, n).fill_philox(2 + k, stream=1, dist=1) for k in range(K)]
    outs = [rls.B200Vector(ctx, dtype, n) for _ in range(K)]
    for _ in range(2):
        op.apply_batch(xs, outs)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        op.apply_batch(xs, outs)
    ms = ctx.timer_stop() / reps
    fpe = 2 if dtype.kind == "c" else 1
    flops = 2 * 2.0 * (m * fpe) * (n * fpe) * (K * fpe) / fpe   # two GEMMs on the real views (complex: 8 m n K each)
    by = 2 * m * n * dtype.itemsize
    print(f"batch normal {dtype.name} {m}x{n} K={K}: {ms:.3f} ms per batched apply = {ms / K * 1e3:.1f} us per right-hand side; "
          f"{flops / ms / 1e9:.1f} TFLOP/s useful ({3 * flops / ms / 1e9:.1f} TF/s of tf32 MMAs), A traffic {by / ms / 1e6:.0f} GB/s", flush=True)
    # the same through K single one-pass applies (CUDA cores)
    ctx.timer_start()
    for k in range(K):
        op.apply(xs[k], outs[k])
    ms1 = ctx.timer_stop()
    print(f"   K single one-pass applies: {ms1:.3f} ms  -> tensor-core batch is {ms1 / ms:.1f}x faster", flush=True)


def time_gram(m, n, dtype):
    import time
    dtype = np.dtype(dtype)
    A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
    ctx.sync()
    t0 = time.perf_counter()
    op = rls.B200NormalOp(A, form="gram")
    ctx.sync()
    dt = time.perf_counter() - t0
    fpe = 2 if dtype.kind == "c" else 1
    flops = 2.0 * (n * fpe) ** 2 * m
    print(f"gram {dtype.name} {m}x{n}: {dt * 1e3:.1f} ms  {flops / dt / 1e12:.1f} TFLOP/s useful (real-view flops; x3 MMAs issued)  [{op.describe()}]", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["check", "time"]
    if "check" in what:
        allok = True
        for dt in (np.float32, np.complex64):
            for (m, n, K) in [(128, 128, 32), (256, 64, 8), (300, 200, 5), (1000, 516, 64), (77, 1030, 3), (4096, 2048, 64)]:
                allok &= check(m, n, K, dt)
            for (m, n) in [(256, 128), (300, 200), (1000, 516)]:
                allok &= check_gram(m, n, dt)
        print("ALL OK" if allok else "FAILURES")
    if "time" in what:
        time_batch(32768, 16384, 64, np.complex64)
        time_batch(16384, 65536, 64, np.float32)
        time_gram(32768, 16384, np.complex64)
