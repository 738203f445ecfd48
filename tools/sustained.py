"""Time the one-pass normal operator over a long back-to-back run (clocks settle under the power cap).
usage: python tools/sustained.py m n dtype seconds"""
import os, sys, subprocess, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
m, n, dtype, secs = int(sys.argv[1]), int(sys.argv[2]), np.dtype(sys.argv[3]), float(sys.argv[4])
ctx = rls.B200Context.default(0)
A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
x = rls.B200Vector(ctx, dtype, n).fill_philox(2, stream=1, dist=1)
if os.environ.get("SPARSE_X"):
    xh = x.to_numpy(); xh[np.arange(n) % 100 != 0] = 0; x.upload(xh)
op = rls.B200NormalOp(A, form="onepass")
g = rls.B200Vector(ctx, dtype, n)
by = m * n * dtype.itemsize
for _ in range(3):
    op.apply(x, g)
ctx.sync()
t_end = time.time() + secs
out = []
while time.time() < t_end:
    ctx.timer_start()
    for _ in range(100):
        op.apply(x, g)
    ms = ctx.timer_stop() / 100
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    out.append((ms, clk))
print(op.describe())
print("x:", "sparse (1 % non-zeros)" if os.environ.get("SPARSE_X") else "dense")
for ms, clk in out[:2] + out[len(out)//2:len(out)//2+1] + out[-2:]:
    print(f"  {ms:.4f} ms/apply {by / ms / 1e6:.0f} GB/s   sm MHz, W: {clk}")
