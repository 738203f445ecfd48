"""Row-sharded solve under torchrun: every rank holds a row block of the same Philox system,
rank 0 also solves the full system on its own GPU; the sharded FISTA / CGNR / ADMM results must
match it (rel-L2 <= 1e-5; the only difference is the summation order of the n-vector allreduce).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py"""
import os
import sys

os.environ.setdefault("RLS_BATCH_MIN_K", "2")   # exercise the tensor-core multi-RHS path with 3 columns

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rls_b200 as rls

rank, world, local = rls.dist.env_rank()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = rls.B200Context.default(local)
rls.dist.init_comm(ctx, rank, world)

ok = True
# a rank that raises must not leave its peers waiting in a collective until the job's time limit
import faulthandler, threading
faulthandler.enable()
threading.Timer(240.0, lambda: (print(f"rank {rank}: watchdog expired", flush=True), os._exit(3))).start()
_excepthook = sys.excepthook
def _die(t, v, tb):
    _excepthook(t, v, tb)
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(2)
sys.excepthook = _die
for dtype, m, n, layout in ((np.float32, 4096, 8192, "row"), (np.complex64, 3001, 2048, "col"), (np.complex64, 3001, 9000, "row")):
    lo, hi = rls.dist.row_range(m, rank, world, align=4)
    scale = 1.0 / np.sqrt(m)
    rho = np.float32(min(0.2, 0.9 / (1.0 + np.sqrt(n / m)) ** 2))   # below 1 / lambda_max(A'A) of the random system
    A_i = rls.B200Matrix.philox(dtype, hi - lo, n, seed=77, scale=scale, row_offset=lo, m_global=m, ctx=ctx, layout=layout)
    b_full = rls.B200Vector(ctx, dtype, m).fill_philox(78, stream=2, dist=1).to_numpy()
    b_i = b_full[lo:hi].copy()
    results = {}
    its = {}
    for name, mk in (("FISTA", lambda A: rls.FISTA(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=30, rho=rho, relTol=0.0, normal="twopass")),
                     ("FISTA-onepass", lambda A: rls.FISTA(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=30, rho=rho, relTol=0.0, normal="onepass")),
                     ("CGNR", lambda A: rls.CGNR(A, reg=rls.L2Regularization(np.float32(1e-2)), iterations=10, relTol=0.0, normal="twopass")),
                     ("ADMM", lambda A: rls.ADMM(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=5, iterationsCG=5, normal="twopass")),
                     # global scalars of a sharded solve: MeasurementBasedNormalization needs the global ‖b‖₁ / length(b),
                     # ADMM's σ_abs = sqrt(length(b))·absTol the global row count — with a stopping rule that triggers
                     ("CGNR-mbn", lambda A: rls.CGNR(A, reg=rls.L2Regularization(np.float32(1e-2)), iterations=10, relTol=0.0, normal="twopass",
                                                     normalizeReg=rls.MeasurementBasedNormalization())),
                     ("ADMM-mbn-abstol", lambda A: rls.ADMM(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=40, iterationsCG=5,
                                                            normal="twopass", absTol=3e-3, normalizeReg=rls.MeasurementBasedNormalization()))):
        S = mk(A_i)
        results[name] = rls.solve_(S, b_i)
        its[name] = S.iteration
    # multi-RHS on the shard (tensor-core GEMM path when the shard is row-major; its n x K product is allreduced)
    B_full = np.stack([np.roll(b_full, 17 * k) for k in range(3)], axis=1)
    Sb = rls.FISTA(A_i, reg=rls.L1Regularization(np.float32(1e-2)), iterations=10, rho=rho, relTol=0.0)
    results["FISTA-batch"] = np.ascontiguousarray(rls.solve_(Sb, np.asfortranarray(B_full[lo:hi])).T).ravel()
    # replicas must be bit-identical across ranks
    for name, x in results.items():
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.float32).copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(t, ref))
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0 and not int(flag[0]):
            print(f"FAIL {name} {np.dtype(dtype).name}: replicas differ across ranks"); ok = False
    dist.barrier()
    if rank == 0:
        # single-GPU reference on a separate, communicator-free context
        ctx1 = rls.B200Context(local)
        A = rls.B200Matrix.philox(dtype, m, n, seed=77, scale=scale, ctx=ctx1, layout=layout)
        S1 = rls.FISTA(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=10, rho=rho, relTol=0.0, ctx=ctx1)
        xb1 = np.concatenate([rls.solve_(S1, B_full[:, k].copy()) for k in range(3)])
        e = np.linalg.norm(results["FISTA-batch"] - xb1) / np.linalg.norm(xb1)
        ok &= e < 1e-5
        print(f"{'ok  ' if e < 1e-5 else 'FAIL'} FISTA-batch    {np.dtype(dtype).name:9s} {m}x{n} ({layout}-major) over {world} GPUs: rel-L2 vs 1 GPU sequential = {e:.2e}")
        for name, mk in (("FISTA", lambda A: rls.FISTA(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=30, rho=rho, relTol=0.0, normal="twopass", ctx=ctx1)),
                         ("CGNR", lambda A: rls.CGNR(A, reg=rls.L2Regularization(np.float32(1e-2)), iterations=10, relTol=0.0, normal="twopass", ctx=ctx1)),
                         ("ADMM", lambda A: rls.ADMM(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=5, iterationsCG=5, normal="twopass", ctx=ctx1)),
                         ("CGNR-mbn", lambda A: rls.CGNR(A, reg=rls.L2Regularization(np.float32(1e-2)), iterations=10, relTol=0.0, normal="twopass",
                                                         normalizeReg=rls.MeasurementBasedNormalization(), ctx=ctx1)),
                         ("ADMM-mbn-abstol", lambda A: rls.ADMM(A, reg=rls.L1Regularization(np.float32(1e-2)), iterations=40, iterationsCG=5,
                                                                normal="twopass", absTol=3e-3, normalizeReg=rls.MeasurementBasedNormalization(), ctx=ctx1))):
            S1 = mk(A)
            x1 = rls.solve_(S1, b_full)
            if S1.iteration != its[name]:
                ok = False
                print(f"FAIL {name}: {its[name]} iterations sharded, {S1.iteration} on one GPU")
            for key in ([name, name + "-onepass"] if name == "FISTA" else [name]):
                e = np.linalg.norm(results[key] - x1) / np.linalg.norm(x1)
                good = e < 1e-5
                if not good and name != "FISTA":
                    # CG-steered solvers amplify the summation order of the all-reduce: the sharded result must then be as
                    # close to the Float64 oracle as the single-GPU result is (the criterion of tests/util.py)
                    import oracle as O
                    A64 = A.to_numpy().astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
                    okw = {"CGNR": dict(reg=O.L2Regularization(1e-2), iterations=10, relTol=0.0),
                           "ADMM": dict(reg=O.L1Regularization(1e-2), iterations=5, iterationsCG=5),
                           "CGNR-mbn": dict(reg=O.L2Regularization(1e-2), iterations=10, relTol=0.0, normalizeReg=O.MeasurementBasedNormalization()),
                           "ADMM-mbn-abstol": dict(reg=O.L1Regularization(1e-2), iterations=40, iterationsCG=5, absTol=3e-3,
                                                   normalizeReg=O.MeasurementBasedNormalization())}[name]
                    x64 = getattr(O, name.split("-")[0])(A64, **okw).solve(b_full.astype(A64.dtype))
                    rel64 = lambda v: np.linalg.norm(v - x64) / np.linalg.norm(x64)
                    good = rel64(results[key]) <= 1.5 * rel64(x1)
                    print(f"     {key}: sharded-vs-oracle64 {rel64(results[key]):.2e}, single-vs-oracle64 {rel64(x1):.2e}")
                ok &= good
                print(f"{'ok  ' if good else 'FAIL'} {key:14s} {np.dtype(dtype).name:9s} {m}x{n} ({layout}-major) over {world} GPUs: rel-L2 vs 1 GPU = {e:.2e}")
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
os._exit(0 if ok else 1)
