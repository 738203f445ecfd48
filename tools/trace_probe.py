import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
capi = rls._capi
m, n = int(os.environ.get("M", "16384")), int(os.environ.get("N", "65536"))
ctx = rls.B200Context.default(0)
A = rls.B200Matrix.philox(np.float32, m, n, seed=1, scale=1.0 / np.sqrt(m), ctx=ctx, layout="row")
b = rls.B200Vector(ctx, np.float32, m).fill_philox(3, stream=1, dist=1)
op = rls.B200NormalOp(A, form="onepass")
S = rls.FISTA(A, AHA=op, reg=rls.L1Regularization(np.float32(1e-3)), iterations=int(os.environ.get("ITS", "40")), rho=np.float32(0.05), relTol=0.0)
it = C.c_int32()
for _ in range(int(os.environ.get("SOLVES", "2"))):
    capi.call("rls_solver_solve", S._handle, b.handle, None, C.byref(it), C.byref(S._scalars))
ctx.sync()
import time
for rep in range(3):
    ctx.timer_start()
    capi.call("rls_solver_solve", S._handle, b.handle, None, C.byref(it), C.byref(S._scalars))
    ms = ctx.timer_stop()
    print(f"solve of {it.value} iterations: {ms:.3f} ms GPU time = {ms / it.value:.4f} ms per iteration  (RLS_TRACE_EVENTS={os.environ.get('RLS_TRACE_EVENTS')}, RLS_PDL={os.environ.get('RLS_PDL')}, RLS_FUSE_ITERATION={os.environ.get('RLS_FUSE_ITERATION')})", flush=True)
