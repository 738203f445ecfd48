import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls
from rls_b200 import _capi as capi
ctx = rls.B200Context.default(0)
m, n, K = 128, 128, 32
rng = np.random.default_rng(0)
A = rng.standard_normal((m, n)).astype(np.float32)
X = rng.standard_normal((n, K)).astype(np.float32)
if len(sys.argv) > 1 and sys.argv[1] == "simple":
    A = np.zeros((m, n), np.float32); A[np.arange(m), np.arange(n)] = 1.0; A[5, 9] = 2.0
    X = np.zeros((n, K), np.float32); X[:K, :K] = np.eye(K) * 3; X[40, 1] = 7
Ad = rls.B200Matrix.from_numpy(A, ctx=ctx, layout="row")
op = rls.B200NormalOp(Ad, form="onepass")
xs = [rls.B200Vector.from_numpy(np.ascontiguousarray(X[:, k]), ctx) for k in range(K)]
outs = op.apply_batch(xs)
Npad = 32
XT = np.zeros((Npad, n), np.float32); capi.call("rls_normal_batch_debug", op.handle, 0, XT.ctypes.data_as(C.c_void_p), XT.size)
Y = np.zeros((m, Npad), np.float32); capi.call("rls_normal_batch_debug", op.handle, 1, Y.ctypes.data_as(C.c_void_p), Y.size)
P = np.zeros((n, Npad), np.float32); capi.call("rls_normal_batch_debug", op.handle, 2, P.ctypes.data_as(C.c_void_p), P.size)
print("XT ok:", np.allclose(XT, X.T))
Yref = A @ X
print("Y  rel:", np.linalg.norm(Y - Yref) / np.linalg.norm(Yref), "absmax", np.abs(Y).max(), "nnz", np.count_nonzero(Y))
Pref = A.T @ Yref
print("P  rel:", np.linalg.norm(P - Pref) / np.linalg.norm(Pref), "absmax", np.abs(P).max())
np.set_printoptions(linewidth=200, precision=3, suppress=True)
print("Y[:8,:8]\n", Y[:8, :8]); print("Yref[:8,:8]\n", Yref[:8, :8])
if np.count_nonzero(Y):
    # which permutation?  correlate rows
    print("Y rows vs ref rows best match:", [int(np.argmax([abs(np.dot(Y[i], Yref[r])) for r in range(m)])) for i in range(8)])
G = np.stack([o.to_numpy() for o in outs], 1)
print("G rel:", np.linalg.norm(G - Pref) / np.linalg.norm(Pref))
