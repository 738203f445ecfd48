import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, rls_b200 as rls, oracle as O
ctx = rls.B200Context.default(0)
m, n, its = 1024, 4096, 50
Ad = rls.B200Matrix.philox(np.complex64, m, n, seed=12345, dist=0, ctx=ctx)
A = Ad.to_numpy()
xt = rls.B200Vector(ctx, np.complex64, n).fill_philox(12345, stream=5, dist=0).to_numpy()
b = (A @ xt).astype(np.complex64)
lam = np.float32(1e-3)
for rt in (0.0, None):
    kw = {} if rt is None else dict(relTol=rt)
    S = rls.CGNR(Ad, reg=rls.L2Regularization(lam), iterations=its, **kw)
    R = O.CGNR(A, reg=O.L2Regularization(lam), iterations=its, **kw)
    S.init_(b); R.init(b)
    for k in range(its + 1):
        a, r = S.iterate(), R.iterate()
        print(rt, k, a, r, S._scalars.rel_res_norm, getattr(R, "rel_res_norm", None), S.iteration, R.iteration, flush=True)
        if not a and not r: break
