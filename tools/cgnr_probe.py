"""FISTA vs CGNR per-iteration time on one GPU at a C5-shard-like shape (ComplexF32, n = 65536)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
capi = rls._capi
m, n = int(os.environ.get("M", "8192")), 65536
dt = np.complex64
ctx = rls.B200Context.default(0)
A = rls.B200Matrix.philox(dt, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
b = rls.B200Vector(ctx, dt, m).fill_philox(3, stream=1, dist=1)
op = rls.B200NormalOp(A, form="onepass")
it = C.c_int32()
def t(S, reps=3):
    f = lambda: capi.call("rls_solver_solve", S._handle, b.handle, None, C.byref(it), C.byref(S._scalars))
    f(); ctx.sync(); ctx.timer_start()
    for _ in range(reps): f()
    return ctx.timer_stop() / reps
x = rls.B200Vector(ctx, dt, n).fill_philox(2, stream=1, dist=1); g = rls.B200Vector(ctx, dt, n)
for _ in range(3): op.apply(x, g)
ctx.sync(); ctx.timer_start()
for _ in range(50): op.apply(x, g)
print("apply only: %.4f ms" % (ctx.timer_stop() / 50))
for its in (50, 100):
    F = rls.FISTA(A, AHA=op, reg=rls.L1Regularization(np.float32(1e-3)), iterations=its, rho=np.float32(0.05), relTol=0.0)
    Cg = rls.CGNR(A, AHA=op, reg=rls.L2Regularization(np.float32(1e-3)), iterations=its, relTol=0.0)
    tf, tc = t(F), t(Cg)
    print(f"{its} iterations: FISTA {tf:.3f} ms ({tf/its:.4f}/it)   CGNR {tc:.3f} ms ({tc/its:.4f}/it)")
