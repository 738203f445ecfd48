import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rls_b200 as rls, faulthandler
faulthandler.enable()
ctx = rls.B200Context.default(0)
ctx1 = rls.B200Context(0)
dtype, m, n, layout = np.complex64, 3001, 9000, "row"
scale = 1.0 / np.sqrt(m)
A = rls.B200Matrix.philox(dtype, m, n, seed=77, scale=scale, ctx=ctx1, layout=layout)
b_full = rls.B200Vector(ctx, dtype, m).fill_philox(78, stream=2, dist=1).to_numpy()
for name, mk in (("ADMM", lambda A: rls.ADMM(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=5, iterationsCG=5, normal="twopass", ctx=ctx1)),
                 ("CGNR-mbn", lambda A: rls.CGNR(A, reg=rls.L2Regularization(np.float32(1e-2)), iterations=10, relTol=0.0, normal="twopass",
                                                 normalizeReg=rls.MeasurementBasedNormalization(), ctx=ctx1)),
                 ("ADMM-mbn-abstol", lambda A: rls.ADMM(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=40, iterationsCG=5,
                                                        normal="twopass", absTol=3e-3, normalizeReg=rls.MeasurementBasedNormalization(), ctx=ctx1))):
    print("start", name, flush=True)
    S1 = mk(A)
    x1 = rls.solve_(S1, b_full)
    print("solved", name, S1.iteration, flush=True)
    import oracle as O
    A64 = A.to_numpy().astype(np.complex128)
    print("downloaded", flush=True)
    okw = {"ADMM": dict(reg=O.L1Regularization(1e-2), iterations=5, iterationsCG=5),
           "CGNR-mbn": dict(reg=O.L2Regularization(1e-2), iterations=10, relTol=0.0, normalizeReg=O.MeasurementBasedNormalization()),
           "ADMM-mbn-abstol": dict(reg=O.L1Regularization(1e-2), iterations=40, iterationsCG=5, absTol=3e-3, normalizeReg=O.MeasurementBasedNormalization())}[name]
    x64 = getattr(O, name.split("-")[0])(A64, **okw).solve(b_full.astype(A64.dtype))
    print(name, np.linalg.norm(x1 - x64) / np.linalg.norm(x64), flush=True)
