import os, sys, faulthandler
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rls_b200 as rls
def p(*a):
    print(*a, flush=True); sys.stderr.flush()
rng = np.random.default_rng(0)
A = rng.standard_normal((320, 480)).astype(np.float32) / 18
b = rng.standard_normal(320).astype(np.float32)
p("create")
S = rls.ADMM(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=5, iterationsCG=3, normal="twopass")
p("created")
S.init_(b)
p("init ok", S._scalars.done, S._scalars.iteration)
for k in range(3):
    r = S.iterate()
    p("iterate", k, r, S._scalars.iteration, S._scalars.cg_iterations_last, list(S._scalars.admm_rk)[:1])
p("x norm", np.linalg.norm(S.x))
x = rls.solve_(S, b)
p("solve ok", np.linalg.norm(x))
