"""One shape, few applies of the row-major one-pass normal operator (for ncu captures).
usage: python tools/rowpass_one.py m n dtype reps"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
m, n, dtype, reps = int(sys.argv[1]), int(sys.argv[2]), np.dtype(sys.argv[3]), int(sys.argv[4])
ctx = rls.B200Context.default(0)
A = rls.B200Matrix.philox(dtype, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
x = rls.B200Vector(ctx, dtype, n).fill_philox(2, stream=1, dist=1)
op = rls.B200NormalOp(A, form="onepass")
g = rls.B200Vector(ctx, dtype, n)
for _ in range(reps):
    op.apply(x, g)
ctx.sync()
print(op.describe())
