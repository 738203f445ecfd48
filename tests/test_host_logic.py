"""CPU tests of the host-side mirror: regularization decorator plumbing, descriptor
resolution, row sharding, and the N>1 composition (world_size-2 gloo) with the oracle as
the per-rank arithmetic."""
import os
import socket

import numpy as np
import pytest

import oracle as O


def test_lambda_plumbing_matches_oracle(rls):
    for lam in (np.float32(1e-3), 1e-3):
        r = rls.L1Regularization(lam); o = O.L1Regularization(lam)
        assert type(rls.lam(r)) is type(O.lam_of(o))
        for f in (np.float32(2.5), np.float64(2.5)):
            rn = rls.NormalizedRegularization(r, f); on = O.NormalizedRegularization(o, f)
            assert rls.lam(rn) == O.lam_of(on) and type(rls.lam(rn)) is type(O.lam_of(on))
    assert rls.lam(rls.PositiveRegularization()) is None
    assert isinstance(rls.sink(rls.NormalizedRegularization(rls.L2Regularization(1.0), 2.0)), rls.L2Regularization)
    regs = [rls.L1Regularization(1.0), rls.PositiveRegularization(), rls.L2Regularization(1.0)]
    assert rls.findsinks(rls.AbstractProjectionRegularization, regs) == [1]
    assert rls.findsink(rls.L2Regularization, regs) == 2
    with pytest.raises(ValueError, match="unambigiously"):
        rls.findsink(rls.L2Regularization, regs + [rls.L2Regularization(2.0)])


def test_reg_desc_resolution(rls):
    from rls_b200.regularization import reg_desc
    d = reg_desc(rls.NormalizedRegularization(rls.L1Regularization(np.float32(0.5)), np.float32(4)))
    assert d.kind == rls._capi.RLS_REG_L1 and d.lambda_ == 2.0 and d.lambda_is_f64 == 0
    d = reg_desc(rls.L21Regularization(0.25, slices=8))
    assert d.kind == rls._capi.RLS_REG_L21 and d.slices == 8 and d.lambda_is_f64 == 1
    d = reg_desc(rls.TVRegularization(np.float32(0.1), shape=(256, 256), iterationsTV=7))
    assert (d.tv_ndims, d.tv_ndirs, list(d.tv_shape)[:2], list(d.tv_dims)[:2], d.tv_iterations) == (2, 2, [256, 256], [1, 2], 7)
    d = reg_desc(rls.L1Regularization(1e-3), rho=0.2, trafo=rls.GradientOp(np.complex64, (16, 8), dims=(2,)))
    assert d.trafo == rls._capi.RLS_TRAFO_GRADIENT and d.tv_ndirs == 1 and list(d.tv_dims)[:1] == [2]
    assert abs(d.rho - 0.2) < 1e-7
    G = rls.GradientOp(np.float32, (16, 8))
    assert G.rows == O.grad_rows((16, 8), (1, 2)) == 15 * 8 + 16 * 7
    # SVD-type terms ride in the TV fields of the POD struct (include/rls_b200.h, RLS_REG_NUCLEAR / RLS_REG_LLR)
    d = reg_desc(rls.NuclearRegularization(np.float32(0.3), svtShape=(4096, 16)))
    assert (d.kind, d.tv_ndims, list(d.tv_shape)[:2]) == (rls._capi.RLS_REG_NUCLEAR, 2, [4096, 16])
    d = reg_desc(rls.LLRRegularization(np.float32(0.3), shape=(64, 48), blockSize=(4, 2), randshift=True, fullyOverlapping=True, seed=11))
    assert (d.kind, d.tv_ndims, list(d.tv_shape)[:2], list(d.tv_dims)[:2]) == (rls._capi.RLS_REG_LLR, 2, [64, 48], [4, 2])
    assert d.tv_iterations == rls._capi.RLS_LLR_RANDSHIFT | rls._capi.RLS_LLR_OVERLAPPING and d.slices == 11
    d = reg_desc(rls.LLRRegularization(np.float32(0.3), shape=(8, 8, 8), randshift=False))
    assert list(d.tv_dims)[:3] == [2, 2, 2] and d.tv_iterations == 0          # blockSize defaults to 2 per dimension (ProxLLR.jl:27)
    with pytest.raises(ValueError):
        rls.NuclearRegularization(np.float32(0.3), svtShape=(4, 4, 4))
    r = rls.LLRRegularization(np.float32(0.3), shape=(8, 8), blockSize=(4, 2), seed=5)
    shifts = [r.next_shift() for _ in range(20)]
    assert all(1 <= a <= 4 and 1 <= b <= 2 for a, b in shifts) and len(set(shifts)) > 1     # rand(CartesianIndices(blockSize))
    assert rls.LLRRegularization(np.float32(0.3), shape=(8, 8), randshift=False).next_shift() == (0, 0)


def test_bench_host_thread_count_respects_affinity():
    """bench.py sizes the BLAS pool of its CPU legs to the threads the process is granted, not to os.cpu_count()"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        n = mod.host_threads()
        assert 1 <= n <= len(os.sched_getaffinity(0))
        assert os.environ["OMP_NUM_THREADS"] == str(n)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_row_ranges_tile_the_matrix(rls):
    for m in (1, 7, 64, 1000, 16384, 262144):
        for w in (1, 2, 3, 4, 8):
            spans = [rls.dist.row_range(m, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % 4 == 0 for lo, hi in spans if lo < m)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rls_b200 as rls
    from oracle.philox import philox_matrix, philox_vector, IH4
    m, n = 301, 64
    lo, hi = rls.dist.row_range(m, rank, world)
    # every rank regenerates exactly its rows of the global Philox matrix (what the GPU ranks do on device)
    A_i = philox_matrix(np.complex64, hi - lo, n, 7, IH4, 0.1, row_offset=lo, m_global=m)
    b = philox_vector(np.complex64, m, 8, 1, IH4)
    x = philox_vector(np.complex64, n, 9, 1, IH4)
    # the N>1 data path: local A_i'(A_i x) and A_i' b_i, one sum-allreduce of the n-vector each
    g = torch.from_numpy(np.ascontiguousarray((A_i.conj().T @ (A_i @ x)).view(np.float32)))
    x0 = torch.from_numpy(np.ascontiguousarray((A_i.conj().T @ b[lo:hi]).view(np.float32)))
    dist.all_reduce(g); dist.all_reduce(x0)
    fro = torch.tensor([float(np.sum(np.abs(A_i.astype(np.complex128)) ** 2))], dtype=torch.float64)
    dist.all_reduce(fro)
    if rank == 0:
        q.put((g.numpy().view(np.complex64), x0.numpy().view(np.complex64), float(fro[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_composition_world2_gloo():
    import torch.multiprocessing as mp
    from oracle.philox import philox_matrix, philox_vector, IH4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g, x0, fro = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, n = 301, 64
    A = philox_matrix(np.complex64, m, n, 7, IH4, 0.1)
    b = philox_vector(np.complex64, m, 8, 1, IH4)
    x = philox_vector(np.complex64, n, 9, 1, IH4)
    assert np.linalg.norm(g - A.conj().T @ (A @ x)) < 1e-5 * np.linalg.norm(g)
    assert np.linalg.norm(x0 - A.conj().T @ b) < 1e-5 * np.linalg.norm(x0)
    assert abs(fro - np.sum(np.abs(A.astype(np.complex128)) ** 2)) < 1e-9 * fro
