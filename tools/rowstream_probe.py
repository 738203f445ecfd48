"""Streaming row kernel (csrc/rls_rowstream.cu): correctness on small / ragged shapes against NumPy float64, burst
timing at the benchmark shapes, sustained timing under the power cap.
usage: python tools/rowstream_probe.py [check] [time] [sustained]"""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls

ctx = rls.B200Context.default(0)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


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
    xd, yd = rls.B200Vector.from_numpy(x, ctx), rls.B200Vector.from_numpy(y, ctx)
    A64 = A.astype(np.complex128 if A.dtype.kind == "c" else np.float64)
    e1 = rel(Ad.mul(xd).to_numpy(), A64 @ x)
    e2 = rel(Ad.adjoint_mul(yd).to_numpy(), A64.conj().T @ y)
    op = rls.B200NormalOp(Ad, form="onepass")
    r1 = op.apply(xd).to_numpy()
    r2 = op.apply(xd).to_numpy()
    e3 = rel(r1, A64.conj().T @ (A64 @ x))
    det = np.array_equal(r1, r2)
    ok = max(e1, e2, e3) < 2e-6 and det
    print(f"{'ok ' if ok else 'BAD'} {np.dtype(dtype).name:9s} {m:6d}x{n:<6d} gemv_n {e1:.1e} gemv_c {e2:.1e} onepass {e3:.1e} deterministic={det}  [{op.describe()}]", flush=True)
    return ok


def timeit(m, n, dtype, reps=20, label=""):
    dtype = np.dtype(dtype)
    A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
    x = rls.B200Vector(ctx, dtype, n).fill_philox(2, stream=1, dist=1)
    op = rls.B200NormalOp(A, form="onepass")
    g = rls.B200Vector(ctx, dtype, n)
    for _ in range(3):
        op.apply(x, g)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        op.apply(x, g)
    ms = ctx.timer_stop() / reps
    by = m * n * dtype.itemsize
    print(f"{label:28s} {dtype.name:9s} {m}x{n}: {ms:.4f} ms/apply {by / ms / 1e6:7.1f} GB/s  [{op.describe()}]", flush=True)
    return A, x, op, g


def sustained(m, n, dtype, secs, label):
    dtype = np.dtype(dtype)
    A, x, op, g = timeit(m, n, dtype, label=label)
    by = m * n * dtype.itemsize
    t_end = time.time() + secs
    out = []
    while time.time() < t_end:
        ctx.timer_start()
        for _ in range(100):
            op.apply(x, g)
        ms = ctx.timer_stop() / 100
        clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
        out.append((ms, clk))
    for ms, clk in out[:1] + out[len(out) // 2:len(out) // 2 + 1] + out[-2:]:
        print(f"      sustained {ms:.4f} ms/apply {by / ms / 1e6:.0f} GB/s   sm MHz, W: {clk}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["check", "time"]
    if "check" in what:
        allok = True
        for dt in (np.float32, np.complex64):
            for (m, n) in [(1, 1), (3, 2), (5, 7), (64, 300), (257, 2048), (130, 4099), (100, 8192), (67, 20000), (41, 65536), (1000, 65536),
                           (35, 70001 if dt == np.float32 else 40000), (777, 16384), (300, 131072 if dt == np.float32 else 65536)]:
                try:
                    allok &= check(m, n, dt)
                except Exception as e:  # noqa: BLE001
                    allok = False
                    print(f"EXC {np.dtype(dt).name} {m}x{n}: {e}", flush=True)
        print("ALL OK" if allok else "FAILURES", flush=True)
    if "time" in what:
        for (m, n, dt) in [(16384, 65536, np.float32), (8192, 65536, np.complex64), (16384, 16384, np.complex64), (65536, 16384, np.float32),
                           (131072, 8192, np.float32)]:
            timeit(m, n, dt)
    if "relayout" in what:
        # what it costs to bring a matrix into the row-major device layout: from the host (staged 64 MB column blocks +
        # tiled transpose, PCIe-bound) and from an adopted column-major device array (rls_mat_relayout, HBM-bound)
        m, n = 16384, 65536
        Ac = rls.B200Matrix.philox(np.float32, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="col")
        for _ in range(2):
            Ar = Ac.relayout("row"); del Ar
        ctx.sync()
        ctx.timer_start()
        Ar = Ac.relayout("row")
        ms = ctx.timer_stop()
        print(f"device re-layout col -> row, Float32 {m}x{n} (4.3 GB): {ms:.3f} ms = {2 * m * n * 4 / ms / 1e6:.0f} GB/s of read+write "
              f"= {ms / 0.582:.1f} one-pass applies", flush=True)
        host = Ac.to_numpy()
        del Ac
        t0 = time.perf_counter()
        Ah = rls.B200Matrix.from_numpy(host, ctx=ctx, layout="row")
        ctx.sync()
        dt = time.perf_counter() - t0
        print(f"host (pageable, column-major) -> device row-major, same matrix: {dt:.2f} s = {m * n * 4 / dt / 1e9:.1f} GB/s", flush=True)
        assert np.array_equal(Ah.to_numpy()[:64], host[:64]) and np.array_equal(Ar.to_numpy()[:64], host[:64])
        del Ah, Ar, host
    if "sustained" in what:
        sustained(16384, 65536, np.float32, 5.0, "C2 shape")
        sustained(8192, 65536, np.complex64, 5.0, "C5 shard shape")
