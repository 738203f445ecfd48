"""Generates the golden fixtures in this directory from the CPU oracle.

The Julia reference cannot be executed in this image (no `julia`), so these vectors are
produced by the oracle at the commit that was pinned against the reference's own known
answers (tests/test_oracle.py).  They freeze the oracle: a later change to oracle/ or to
the CUDA path that moves any iterate shows up against these files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O  # noqa: E402
from util import rand_matrix, rand_vector, sparse_truth  # noqa: E402


def problem(dtype, m, n, seed):
    A, _ = rand_matrix(dtype, m, n, seed)
    xt = sparse_truth(dtype, n, seed + 1)
    b = (A @ xt + 1e-3 * rand_vector(dtype, m, seed + 2)).astype(dtype)
    return A, b


def trace(solver, b, iters):
    solver.init(b)
    xs = []
    while solver.iterate():
        xs.append(solver.x.copy())
    assert len(xs) == iters
    return np.stack(xs)


CASES = {
    # name: (dtype, m, n, seed, solver factory(A), iterations)
    "fista_l1_f32": (np.float32, 96, 160, 500, lambda A: O.FISTA(A, reg=O.L1Regularization(np.float32(2e-2)), iterations=12, rho=np.float32(0.2), relTol=0.0), 12),
    "fista_l1_restart_c64": (np.complex64, 96, 160, 510, lambda A: O.FISTA(A, reg=O.L1Regularization(np.float32(2e-2)), iterations=12, rho=np.float32(0.2), relTol=0.0, restart="gradient"), 12),
    "pogm_l1_c64": (np.complex64, 96, 160, 520, lambda A: O.POGM(A, reg=O.L1Regularization(np.float32(2e-2)), iterations=12, rho=np.float32(0.2), relTol=0.0, restart="gradient"), 12),
    "optista_l1_f32": (np.float32, 96, 160, 530, lambda A: O.OptISTA(A, reg=O.L1Regularization(np.float32(2e-2)), iterations=12, rho=np.float32(0.2), relTol=0.0), 12),
    "cgnr_l2_c64": (np.complex64, 160, 96, 540, lambda A: O.CGNR(A, reg=O.L2Regularization(np.float32(1e-3)), iterations=10, relTol=0.0), 10),
    "admm_l1_c64": (np.complex64, 128, 96, 550, lambda A: O.ADMM(A, reg=O.L1Regularization(np.float32(1e-2)), iterations=8, iterationsCG=10, rho=0.1, absTol=0.0, relTol=0.0), 8),
    "kaczmarz_l2_f32": (np.float32, 160, 96, 570, lambda A: O.Kaczmarz(A, reg=O.L2Regularization(np.float32(1e-2)), iterations=6), 6),
    "kaczmarz_l2_l1_pos_c64": (np.complex64, 96, 160, 580, lambda A: O.Kaczmarz(A, reg=[O.L2Regularization(np.float32(1e-2)), O.L1Regularization(np.float32(1e-3)), O.PositiveRegularization()], iterations=6), 6),
    "admm_tv_f32": (np.float32, 128, 12 * 8, 560, lambda A: O.ADMM(A, reg=O.TVRegularization(np.float32(1e-2), shape=(12, 8)), iterations=8, iterationsCG=10, rho=0.1, absTol=0.0, relTol=0.0), 8),
}


def main():
    out = {}
    for name, (dtype, m, n, seed, make, iters) in CASES.items():
        A, b = problem(dtype, m, n, seed)
        out[name + "_x"] = trace(make(A), b, iters)
    # proximal maps on a fixed vector
    x = rand_vector(np.complex64, 12 * 8 * 4, 600)
    for key, reg in (("l1", O.L1Regularization(np.float32(0.3))), ("l2", O.L2Regularization(np.float32(0.3))),
                     ("l21", O.L21Regularization(np.float32(1.2), slices=4)),
                     ("tv", O.TVRegularization(np.float32(0.2), shape=(12, 8, 4))), ("pos", O.PositiveRegularization())):
        out["prox_" + key] = O.prox_(reg, x.copy())
    np.savez_compressed(os.path.join(HERE, "golden_r01.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_r01.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
