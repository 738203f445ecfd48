"""Run the normal operator alone at the C2 shape (for ncu captures / timing sweeps).
usage: python tools/profile_normal.py FORM [reps] [dtype] [m] [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls

form = sys.argv[1] if len(sys.argv) > 1 else "onepass"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dtype = np.dtype(sys.argv[3]) if len(sys.argv) > 3 else np.dtype(np.float32)
m = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
n = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
ctx = rls.B200Context.default(0)
A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx)
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
print(f"[{op.describe()}] {form} {dtype} {m}x{n} : {ms:.4f} ms/apply  {by / ms / 1e6:.1f} GB/s algorithmic", flush=True)
