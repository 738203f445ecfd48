"""N>1 path on real GPUs: launches tools/multi_gpu_check.py under torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_row_sharded_solvers_match_single_gpu(exchange):
    """exchange = nccl: ncclAllReduce of the n-vector (default); p2p: the one-shot NVLink peer-memory all-reduce fused
    with the partial sum of the one-pass kernel (csrc/rls_p2p.cu, RLS_P2P=1)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the single-GPU box runs the gloo composition test on CPU instead)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    env = dict(os.environ, RLS_P2P="1" if exchange == "p2p" else "0")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=env)
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [1, 2, 4])
def test_device_group_inside_one_process(ndev):
    """"multi-GPU handled inside one call" — rls_group_*: ONE process, the system row-partitioned over a
    device group, createLinearSolver / solve! unchanged for the caller.  A group of one device must reproduce the plain
    single-GPU solve bit for bit; larger groups (skipped when the box has fewer GPUs) must agree with it to the parity
    bound (the only difference is the summation order of the all-reduced n-vector) with bit-identical replicas (checked
    inside rls_group_solver_solve_host)."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rls_b200 as rls
    import oracle as O
    from util import rel, rand_matrix, rand_vector, sparse_truth, up64, to64, assert_close_or_fp64
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    for dtype, m, n in ((np.complex64, 1030, 9000), (np.float32, 515, 1024)):
        A, _ = rand_matrix(dtype, m, n, 41)
        b = (A @ sparse_truth(dtype, n, 42) + 1e-3 * rand_vector(dtype, m, 43)).astype(dtype)
        rho = np.float32(0.9 / (1.0 + np.sqrt(n / m)) ** 2)
        G = rls.B200Group(ndev)
        Ag = rls.B200GroupMatrix.from_numpy(A, G)
        blocks = Ag.row_blocks()
        assert blocks[0][0] == 0 and blocks[-1][1] == m and all(blocks[i][1] == blocks[i + 1][0] for i in range(ndev - 1))
        A1 = rls.B200Matrix.from_numpy(A)                    # same layout choice (AUTO) as the blocks of the group
        cases = (("FISTA", dict(reg=rls.L1Regularization(np.float32(1e-2)), iterations=30, rho=rho, relTol=0.0),
                  dict(reg=O.L1Regularization(np.float32(1e-2)), iterations=30, rho=rho, relTol=0.0)),
                 ("CGNR", dict(reg=rls.L2Regularization(np.float32(1e-2)), iterations=10, relTol=0.0),
                  dict(reg=O.L2Regularization(np.float32(1e-2)), iterations=10, relTol=0.0)),
                 ("ADMM", dict(reg=rls.L1Regularization(np.float32(1e-2)), iterations=5, iterationsCG=5),
                  dict(reg=O.L1Regularization(np.float32(1e-2)), iterations=5, iterationsCG=5)))
        for name, kw, okw in cases:
            Sg = getattr(rls, name)(Ag, **kw)
            xg = rls.solve_(Sg, b)
            S1 = getattr(rls, name)(A1, **kw)
            x1 = rls.solve_(S1, b)
            assert Sg.iteration == S1.iteration
            if ndev == 1:
                assert np.array_equal(xg, x1), name
            else:
                x32 = getattr(O, name)(A, **okw).solve(b)
                assert_close_or_fp64(xg, x32, lambda: getattr(O, name)(up64(A), **to64(okw)).solve(up64(b)), what=f"{name} on {ndev} devices")
        # Philox generation per row block reproduces the global matrix
        Ap = rls.B200GroupMatrix.philox(G, dtype, m, n, seed=5, scale=0.1)
        Sp = rls.FISTA(Ap, reg=rls.L1Regularization(np.float32(1e-2)), iterations=5, rho=np.float32(0.01), relTol=0.0)
        S1 = rls.FISTA(rls.B200Matrix.philox(dtype, m, n, seed=5, scale=0.1), reg=rls.L1Regularization(np.float32(1e-2)),
                       iterations=5, rho=np.float32(0.01), relTol=0.0)
        xp, x1 = rls.solve_(Sp, b), rls.solve_(S1, b)
        assert np.array_equal(xp, x1) if ndev == 1 else rel(xp, x1) < 1e-5
