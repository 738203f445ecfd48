"""Row-major one-pass kernels: correctness against NumPy (float64) on small / ragged shapes and timing at
the benchmark shapes.  usage: python tools/rowpass_check.py [check] [time]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls

ctx = rls.B200Context.default(0)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def check(m, n, dtype):
    rng = np.random.default_rng(m * 131 + n)
    A = rng.standard_normal((m, n)).astype(np.float32)
    x = rng.standard_normal(n).astype(np.float32)
    y = rng.standard_normal(m).astype(np.float32)
    if np.dtype(dtype).kind == "c":
        A = (A + 1j * rng.standard_normal((m, n))).astype(np.complex64)
        x = (x + 1j * rng.standard_normal(n)).astype(np.complex64)
        y = (y + 1j * rng.standard_normal(m)).astype(np.complex64)
    Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
    back = Ad.to_numpy()
    assert np.array_equal(back, A), "upload/download round trip"
    xd, yd = rls.B200Vector.from_numpy(x, ctx), rls.B200Vector.from_numpy(y, ctx)
    A64 = A.astype(np.complex128 if A.dtype.kind == "c" else np.float64)
    e1 = rel(Ad.mul(xd).to_numpy(), A64 @ x)
    e2 = rel(Ad.adjoint_mul(yd).to_numpy(), A64.conj().T @ y)
    op = rls.B200NormalOp(Ad, form="onepass")
    e3 = rel(op.apply(xd).to_numpy(), A64.conj().T @ (A64 @ x))
    op2 = rls.B200NormalOp(Ad, form="twopass")
    e4 = rel(op2.apply(xd).to_numpy(), A64.conj().T @ (A64 @ x))
    ok = max(e1, e2, e3, e4) < 2e-6
    print(f"{'ok ' if ok else 'BAD'} {np.dtype(dtype).name:9s} {m:6d}x{n:<6d} gemv_n {e1:.1e} gemv_c {e2:.1e} onepass {e3:.1e} twopass {e4:.1e}  [{op.describe()}]", flush=True)
    return ok


def timeit(m, n, dtype, form="onepass", reps=10):
    dtype = np.dtype(dtype)
    A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
    x = rls.B200Vector(ctx, dtype, n).fill_philox(2, stream=1, dist=1)
    op = rls.B200NormalOp(A, form=form)
    g = rls.B200Vector(ctx, dtype, n)
    for _ in range(3):
        op.apply(x, g)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        op.apply(x, g)
    ms = ctx.timer_stop() / reps
    by = m * n * dtype.itemsize
    env = {k: v for k, v in os.environ.items() if k.startswith("RLS_ROWPASS")}
    print(f"{form} {dtype.name} {m}x{n}: {ms:.4f} ms/apply {by / ms / 1e6:.1f} GB/s  [{op.describe()}] {env}", flush=True)


def time_gemv(m, n, dtype, reps=10):
    dtype = np.dtype(dtype)
    A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
    x = rls.B200Vector(ctx, dtype, n).fill_philox(2, stream=1, dist=1)
    y = rls.B200Vector(ctx, dtype, m)
    g = rls.B200Vector(ctx, dtype, n)
    by = m * n * dtype.itemsize
    for name, fn in (("gemv_n", lambda: A.mul(x, y)), ("gemv_c", lambda: A.adjoint_mul(y, g))):
        for _ in range(3):
            fn()
        ctx.sync()
        ctx.timer_start()
        for _ in range(reps):
            fn()
        ms = ctx.timer_stop() / reps
        print(f"{name} {dtype.name} {m}x{n}: {ms:.4f} ms {by / ms / 1e6:.1f} GB/s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["check", "time"]
    if "check" in what:
        allok = True
        for dt in (np.float32, np.complex64):
            for (m, n) in [(1, 1), (3, 2), (5, 7), (64, 300), (257, 2048), (130, 4099), (100, 8192), (67, 20000), (41, 65536),
                           (35, 70001 if dt == np.float32 else 40000)]:
                allok &= check(m, n, dt)
        print("ALL OK" if allok else "FAILURES")
    if "shapes" in what:
        for (m, n) in [(131072, 8192), (65536, 16384), (32768, 32768), (16384, 65536)]:
            timeit(m, n, np.float32)
            time_gemv(m, n, np.float32)
    if "time" in what:
        timeit(16384, 65536, np.float32)
        timeit(8192, 65536, np.complex64)
        timeit(16384, 16384, np.complex64)
        timeit(16384, 65536, np.float32, form="twopass")
