"""One batched tensor-core apply at the C4 shape (for ncu captures)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
m, n, K = 32768, 16384, 64
ctx = rls.B200Context.default(0)
A = rls.B200Matrix.philox(np.complex64, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
op = rls.B200NormalOp(A, form="onepass")
xs = [rls.B200Vector(ctx, np.complex64, n).fill_philox(2 + k, stream=1, dist=1) for k in range(K)]
outs = [rls.B200Vector(ctx, np.complex64, n) for _ in range(K)]
for _ in range(2):
    op.apply_batch(xs, outs)
ctx.sync()
