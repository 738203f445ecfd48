"""Where does the per-iteration time outside the normal-operator kernel go?  (launch-gap probe)"""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
capi = rls._capi
m, n = 16384, 65536
ctx = rls.B200Context.default(0)
A = rls.B200Matrix.philox(np.float32, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
x = rls.B200Vector(ctx, np.float32, n).fill_philox(2, stream=1, dist=1)
g = rls.B200Vector(ctx, np.float32, n)
b = rls.B200Vector(ctx, np.float32, m).fill_philox(3, stream=1, dist=1)
op = rls.B200NormalOp(A, form="onepass")
def t(fn, reps=200):
    for _ in range(5): fn()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): fn()
    return ctx.timer_stop() / reps
print("apply only                 : %.4f ms" % t(lambda: op.apply(x, g)))
print("apply + prox_l1            : %.4f ms" % t(lambda: (op.apply(x, g), rls.prox_(rls.L1Regularization(np.float32(1e-3)), g, np.float32(1e-3)))))
print("apply + 3 x prox_l1        : %.4f ms" % t(lambda: (op.apply(x, g), [rls.prox_(rls.L1Regularization(np.float32(1e-3)), g, np.float32(1e-3)) for _ in range(3)])))
print("apply + nrm2 (host sync)   : %.4f ms" % t(lambda: (op.apply(x, g), g.norm())))
for its in (50, 200):
    S = rls.FISTA(A, AHA=op, reg=rls.L1Regularization(np.float32(1e-3)), iterations=its, rho=np.float32(0.05), relTol=0.0)
    it = C.c_int32()
    def solve():
        capi.call("rls_solver_solve", S._handle, b.handle, None, C.byref(it), C.byref(S._scalars))
    ms = t(solve, 5)
    print(f"FISTA solve {its:3d} iterations   : {ms:.3f} ms per solve = {ms / its:.4f} ms per iteration")
S = rls.CGNR(A, AHA=op, reg=rls.L2Regularization(np.float32(1e-3)), iterations=100, relTol=0.0)
ms = t(lambda: capi.call("rls_solver_solve", S._handle, b.handle, None, C.byref(it), C.byref(S._scalars)), 5)
print(f"CGNR solve 100 iterations    : {ms:.3f} ms per solve = {ms / 100:.4f} ms per iteration")
