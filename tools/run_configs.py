"""Measure the BASELINE.json configs that are not the bench line (C1, C3, C4, C5) and print one
JSON line each.  C5 runs under torchrun (row-sharded); the others on one GPU.
  python tools/run_configs.py c1 c3 c4
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py c5"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls

PEAK = 6537.3
rank, world, local = rls.dist.env_rank()
multi = world > 1
if multi:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = rls.B200Context.default(local)
if multi:
    rls.dist.init_comm(ctx, rank, world)


def timed_solve(S, b_dev, reps):
    it = C.c_int32()
    call = lambda: rls._capi.call("rls_solver_solve", S._handle, b_dev.handle, None, C.byref(it), C.byref(S._scalars))
    call()
    ctx.sync()
    if multi:
        dist.barrier()
    ctx.timer_start()
    for _ in range(reps):
        call()
    ms = ctx.timer_stop() / reps
    if multi:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms, it.value


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)


def c1():
    m, n, its = 1024, 4096, 50
    A = rls.B200Matrix.philox(np.complex64, m, n, seed=12345, dist=0, ctx=ctx)
    xt = rls.B200Vector(ctx, np.complex64, n).fill_philox(12345, stream=5, dist=0)
    b = A.mul(xt)
    for graph, persistent in (("1", "0"), ("0", "0"), ("0", "1")):
        os.environ["RLS_SOLVE_GRAPH"] = graph
        os.environ["RLS_CGNR_PERSISTENT"] = persistent
        S = rls.CGNR(A, reg=rls.L2Regularization(np.float32(1e-3)), iterations=its, relTol=0.0)
        timed_solve(S, b, 2)                                   # first solve launch by launch, second one records the graph
        ms, done = timed_solve(S, b, 20)
        emit({"config": "C1 CGNR + L2, ComplexF32 1024x4096, 50 iterations (L2-resident, latency-bound)", "iterations": done,
              "whole_solve_cuda_graph": graph == "1", "whole_solve_cooperative_kernel": persistent == "1",
              "rel_res_norm": S._scalars.rel_res_norm, "ms_per_solve": ms, "us_per_iteration": 1e3 * ms / its,
              "iterations_per_s": its / ms * 1e3, "normal_operator": S.AHA.describe(),
              "note": "33.6 MB of A stays in L2; per-iteration time is launch/latency, not HBM; ms_per_solve includes init! (A'b)"})
    os.environ.pop("RLS_SOLVE_GRAPH")
    os.environ.pop("RLS_CGNR_PERSISTENT")


def c3():
    m, n = 131072, 65536
    A = rls.B200Matrix.philox(np.complex64, m, n, seed=1234, scale=1.0 / np.sqrt(m), ctx=ctx)
    img = np.zeros((256, 256), np.complex64, order="F")
    rng = np.random.default_rng(1234)
    for _ in range(5):
        i, j = rng.integers(0, 256, 2)
        img[i:, j:] += np.float32(rng.standard_normal())
    xt = rls.B200Vector.from_numpy(img.ravel(order="F"), ctx)
    b = A.mul(xt)
    outer = int(os.environ.get("C3_OUTER", "10"))
    S = rls.ADMM(A, reg=rls.TVRegularization(np.float32(1e-2), shape=(256, 256)), rho=0.1, iterations=outer, iterationsCG=10,
                 absTol=0.0, relTol=0.0)
    ms, done = timed_solve(S, b, 1)
    applies = S._scalars.cg_iterations_total + outer      # 1 residual apply + n_cg per outer iteration
    by = m * n * 8
    emit({"config": "C3 ADMM + TVRegularization(256x256), ComplexF32 131072x65536 (68.7 GB), rho=0.1, iterationsCG=10",
          "outer_iterations": done, "cg_steps_total": S._scalars.cg_iterations_total, "normal_applies": applies,
          "ms_per_outer_iteration": ms / outer, "ms_per_apply": ms / applies, "applies_per_s": applies / ms * 1e3,
          "achieved_gbs": by * applies / ms / 1e6, "frac_of_measured_hbm": by * applies / ms / 1e6 / PEAK,
          "note": f"timed {outer} outer iterations (the config names 100; time per outer iteration is what scales)"})


def c4():
    m, n, K, its = 32768, 16384, 64, 50
    A = rls.B200Matrix.philox(np.complex64, m, n, seed=4321, scale=1.0 / np.sqrt(m), ctx=ctx)
    X = np.zeros((n, K), np.complex64, order="F")
    rng = np.random.default_rng(4)
    for k in range(K):
        idx = rng.integers(0, n, 160)
        X[idx, k] = (rng.random(160) + 1j * rng.random(160)).astype(np.complex64)
    B = np.empty((m, K), np.complex64, order="F")
    for k in range(K):
        B[:, k] = A.mul(rls.B200Vector.from_numpy(X[:, k].copy(), ctx)).to_numpy()
    AHA = rls.B200NormalOp(A, form="auto")
    b0 = rls.B200Vector(ctx, np.complex64, n).fill_philox(9, stream=1, dist=1)
    rho = np.float32(0.95 / AHA.power_iterations(b0))
    S = rls.FISTA(A, AHA=AHA, reg=rls.L1Regularization(np.float32(1e-3)), iterations=its, rho=rho, relTol=0.0)
    import time
    rls.solve_(S, B)          # warm-up at the full width: lane allocation and the GEMM plan are one-time costs
    ctx.sync()
    t0 = time.perf_counter()
    Xs = rls.solve_(S, B)
    ctx.sync()
    dt = time.perf_counter() - t0
    err = np.linalg.norm(Xs - X) / np.linalg.norm(X)
    emit({"config": "C4 multi-RHS FISTA-L1: 64 frames sharing ComplexF32 A 32768x16384, 50 iterations", "s_per_batched_solve": dt,
          "frame_iterations_per_s": K * its / dt, "rel_err_vs_truth": float(err),
          "ms_per_batched_iteration": dt / its * 1e3,
          "note": "per batched iteration: K per-frame pre kernels, two tcgen05 GEMMs (Y = A X, G = A' Y; kind::tf32 x3 split, "
                  "A read once each), K per-frame fused epilogues; init A'b per frame on CUDA cores; includes H2D of B and D2H of X; "
                  f"matrix layout {A.layout}"})
    if not os.environ.get("RLS_C4_NO_GRAM"):
        # the reference's default form: AHA = A'*A built once (tensor cores), then ONE GEMM over G per batched iteration
        ctx.sync()
        t0 = time.perf_counter()
        G = rls.B200NormalOp(A, form="gram")
        ctx.sync()
        t_build = time.perf_counter() - t0
        Sg = rls.FISTA(A, AHA=G, reg=rls.L1Regularization(np.float32(1e-3)), iterations=its, rho=rho, relTol=0.0)
        rls.solve_(Sg, B)
        ctx.sync()
        t0 = time.perf_counter()
        Xg = rls.solve_(Sg, B)
        ctx.sync()
        dtg = time.perf_counter() - t0
        emit({"config": "C4 same, Gram form (A'*A precomputed on tensor cores; one tcgen05 GEMM over G per batched iteration)",
              "gram_build_s": t_build, "gram_build_tflops_fp32_equiv": 8.0 * m * n * n / t_build / 1e12,
              "s_per_batched_solve": dtg, "frame_iterations_per_s": K * its / dtg, "ms_per_batched_iteration": dtg / its * 1e3,
              "rel_err_vs_truth": float(np.linalg.norm(Xg - X) / np.linalg.norm(X)),
              "max_rel_diff_vs_a_form_columns": float(max(np.linalg.norm(Xg[:, k] - Xs[:, k]) / np.linalg.norm(Xs[:, k]) for k in range(K))),
              "normal_operator": G.describe()})
        del Sg, G
    if os.environ.get("RLS_C4_COMPARE"):
        os.environ["RLS_BATCH_TENSOR_CORES"] = "0"
        t0 = time.perf_counter()
        Xc = rls.solve_(S, B)
        ctx.sync()
        dt2 = time.perf_counter() - t0
        emit({"config": "C4 same, K single one-pass applies per iteration (CUDA cores, RLS_BATCH_TENSOR_CORES=0)",
              "s_per_batched_solve": dt2, "frame_iterations_per_s": K * its / dt2,
              "max_rel_diff_vs_tensor_core_columns": float(max(np.linalg.norm(Xc[:, k] - Xs[:, k]) / np.linalg.norm(Xc[:, k]) for k in range(K)))})
        os.environ.pop("RLS_BATCH_TENSOR_CORES")


def c5():
    m, n = 262144, 65536
    lo, hi = rls.dist.row_range(m, rank, world, align=4)
    A = rls.B200Matrix.philox(np.complex64, hi - lo, n, seed=12345, scale=1.0 / np.sqrt(m), row_offset=lo, m_global=m, ctx=ctx)
    xt = rls.B200Vector(ctx, np.complex64, n).fill_philox(12345, stream=11, dist=0)
    xh = xt.to_numpy(); xh[np.arange(n) % 100 != 0] = 0; xt.upload(xh)
    b = A.mul(xt)
    by = m * n * 8
    AHA = rls.B200NormalOp(A, form="auto")
    b0 = rls.B200Vector(ctx, np.complex64, n).fill_philox(12345, stream=13, dist=1)
    rho = np.float32(0.95 / AHA.power_iterations(b0, maxiter=10))
    for name, S, its in (("FISTA-L1", rls.FISTA(A, AHA=AHA, reg=rls.L1Regularization(np.float32(1e-3)), iterations=50, rho=rho, relTol=0.0), 50),
                         ("CGNR-L2", rls.CGNR(A, AHA=AHA, reg=rls.L2Regularization(np.float32(1e-3)), iterations=50, relTol=0.0), 50)):
        ms, done = timed_solve(S, b, 1)
        emit({"config": f"C5 row-sharded {name}, ComplexF32 262144x65536 (137.4 GB) on {world} GPU(s)", "n_gpus": world,
              "iterations": done, "ms_per_iteration": ms / its, "iterations_per_s": its / ms * 1e3,
              "aggregate_gbs": by * its / ms / 1e6, "frac_of_aggregate_measured_hbm": by * its / ms / 1e6 / (PEAK * world),
              "normal_operator": AHA.describe(),
              "note": "one-pass normal operator on the row shard + one NCCL allreduce of the 65536-vector per iteration; 50 iterations timed"})


for name in sys.argv[1:]:
    {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[name]()
if multi:
    dist.barrier()
    dist.destroy_process_group()
