#!/usr/bin/env python
"""bench.py — FISTA-L1 and CGNR iterations/s on the dense ComplexF32 system 262144 x 65536 (137.4 GB; BASELINE.json
configs[4], the configuration its metric "iterations/s ... at 1/2/4/8" and its north-star target are quoted on) through
librls_b200, with the roofline of the one-pass normal-operator kernel and the restated-reference CPU baseline beside it.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                    (the oracle port on the host cores, bounded row sample)

STRONG scaling: the global problem is fixed; A is row-partitioned over the N ranks (N = 1: the whole 137.4 GB matrix in
the 180 GB of one B200), x and every n-vector are replicated, one NCCL all-reduce of the 65536-vector per iteration.
A "step" is one solve!(solver, b): init! (one back-projection A'b) + 100 FISTA-L1 iterations on one right-hand side;
`value` = true iterations/s of the whole job (not multiplied by N), timed over K steps with b resident in HBM; `e2e`
times the same K steps through the public API with HOST buffers (H2D of the b shard and D2H of x inside the timed
region).  `cgnr` carries the same measurement for CGNR + L2 (50 iterations per step).  At N > 1 the run first solves a
small twin (8192 x 65536) row-sharded AND on rank 0's GPU alone and records rel-L2 between the two (`parity`, bound 2e-5:
two Float32 renderings, each held to 1e-5 against the oracle by the tests).
`secondary_c2` keeps round 1's line: FISTA-L1 on one Float32 16384 x 65536 shard per GPU (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)



def host_threads():
    """CPU threads this process may really use: the affinity mask, cut by a cgroup CPU quota when there is one (a BLAS pool
    sized to os.cpu_count() inside a smaller quota runs SLOWER than one thread per granted core)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        if os.path.exists("/sys/fs/cgroup/cpu.max"):                                   # cgroup v2
            with open("/sys/fs/cgroup/cpu.max") as f:
                quota, period = f.read().split()[:2]
        else:                                                                          # cgroup v1
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
                quota = f.read().strip()
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                period = f.read().strip()
        if quota not in ("max", "-1"):
            n = max(1, min(n, int(-(-int(quota) // int(period)))))
    except Exception:
        pass
    return n


# The CPU legs (reference arm, cpu_baseline) use every host thread they are granted and say how many: torchrun exports
# OMP_NUM_THREADS=1 for N > 1, and an unset variable lets OpenBLAS size its pool to the machine instead of the quota.
for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_k] = str(host_threads())

import numpy as np  # noqa: E402

M_GLOBAL, N_COLS = 262144, 65536
FISTA_ITERS, CGNR_ITERS = 100, 50
LAMBDA = np.float32(1e-3)
SEED = 12345
TWIN_M = 8192
TWIN_TOL = 2e-5
METRIC = "FISTA-L1 iterations/s on dense ComplexF32 A 262144x65536, row-sharded over N GPUs (strong scaling)"
WORKLOAD = ("C5: FISTA + L1Regularization(1f-3) [value] and CGNR + L2Regularization(1f-3) [cgnr] on the dense ComplexF32 "
            "system 262144x65536 (137.4 GB, Philox CN(0,1)/sqrt(m) generated on the devices), rho = 0.95/lambda_max "
            "(10 power iterations), relTol = 0; one step = one solve! = init! + 100 (FISTA) / 50 (CGNR) iterations")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[1]) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 5 + k and s[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's loop (NumPy -> threaded OpenBLAS cgemv) on a bounded ROW SAMPLE of the same Philox matrix
# ------------------------------------------------------------------------------------------------------------------
_CPU = {}


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get("num_threads", 1)) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


def cpu_problem(m_sub):
    """Rows [0, m_sub) of the global C5 matrix (the very entries the devices generate) and the matching part of b."""
    if m_sub not in _CPU:
        from concurrent.futures import ThreadPoolExecutor
        from oracle.philox import philox_values, IH4
        scale = np.float32(1.0 / np.sqrt(M_GLOBAL))
        A = np.empty((m_sub, N_COLS), dtype=np.complex64, order="F")
        rows = np.arange(m_sub, dtype=np.uint64)

        def fill(j0):
            cols = np.arange(j0, min(N_COLS, j0 + 128), dtype=np.uint64)
            idx = rows[:, None] + cols[None, :] * np.uint64(M_GLOBAL)
            A[:, j0:j0 + cols.size] = philox_values(SEED, idx, 0, 0, IH4, scale) + 1j * philox_values(SEED, idx, 0, 1, IH4, scale)

        with ThreadPoolExecutor(max_workers=host_threads()) as ex:   # NumPy releases the GIL inside the big array ops
            list(ex.map(fill, range(0, N_COLS, 128)))
        rng = np.random.default_rng(SEED)
        b = (rng.standard_normal(m_sub) + 1j * rng.standard_normal(m_sub)).astype(np.complex64)
        _CPU[m_sub] = (A, b)
    return _CPU[m_sub]


def cpu_step(kind, m_sub, iters):
    """`iters` iterations of the oracle solver on the row sample; returns (iterations done, seconds)."""
    import oracle as O
    A, b = cpu_problem(m_sub)
    if kind == "fista":
        S = O.FISTA(A, reg=O.L1Regularization(LAMBDA), iterations=iters + 1, rho=np.float32(0.05), relTol=0.0)
    else:
        S = O.CGNR(A, reg=O.L2Regularization(LAMBDA), iterations=iters + 1, relTol=0.0)
    S.init(b)
    S.iterate()                      # first touch of the work arrays
    t0 = time.perf_counter()
    k = 0
    while k < iters and S.iterate():
        k += 1
    return k, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                  # N > 1: rank 0 alone runs the CPU arm
    cores = host_threads()
    m_sub = 4096                                # 2.1 GB of the 137.4 GB: the gemv pair is DRAM-bound, i.e. linear in m
    scale_to_full = m_sub / M_GLOBAL
    t_gen = time.perf_counter()
    cpu_problem(m_sub)
    t_gen = time.perf_counter() - t_gen
    k0, dt0 = cpu_step("fista", m_sub, 4)
    iters = int(max(8, min(60, round(k0 / dt0 * 4.0))))     # ~4 s of CPU work per step
    for _ in range(min(args.warmup, 1)):
        cpu_step("fista", m_sub, iters)
    t0 = time.perf_counter()
    k_tot, dt_tot = 0, 0.0
    for _ in range(args.steps):
        k, dt = cpu_step("fista", m_sub, iters)
        k_tot += k
        dt_tot += dt
    wall = time.perf_counter() - t0
    value = k_tot / dt_tot * scale_to_full
    kc, dtc = cpu_step("cgnr", m_sub, iters)
    threads = blas_threads()
    sample = (f"{iters} FISTA-L1 iterations per step x {args.steps} steps of the oracle loop (NumPy + OpenBLAS cgemv pair, {threads} BLAS "
              f"threads on {cores} host cores) on rows 0..{m_sub - 1} of the same Philox matrix ({m_sub}x{N_COLS} ComplexF32, 2.1 GB); "
              f"iterations/s scaled by {m_sub}/{M_GLOBAL} to the full system (the iteration is DRAM-bound in A: linear in m)")
    line = {"metric": METRIC, "value": value, "unit": "iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt_tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c64 (ComplexF32)", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "sample_rows": m_sub, "iterations_timed_per_step": iters,
                       "iterations_per_s_on_the_sample": k_tot / dt_tot, "scaled_by": scale_to_full,
                       "ms_per_step_is": "the measured time of one step on the SAMPLE (iterations_timed_per_step iterations of the "
                                         f"{m_sub}-row system), not a synthesised full-size figure",
                       "matrix_generation_s": t_gen,
                       "note": "Julia is not installed in this image; the reference arm is the float32-faithful oracle port of "
                               "src/FISTA.jl:139-185 with the lazy two-gemv normal operator on OpenBLAS (the BLAS family Julia ships)"},
            "cgnr": {"value": kc / dtc * scale_to_full, "unit": "iterations/s", "iterations_timed": kc},
            "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import ctypes as C
    import rls_b200 as rls
    capi = rls._capi
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
    dt = np.dtype(np.complex64)

    def barrier():
        ctx.sync()
        if multi:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        barrier()
        if multi:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    def sparse_truth(c, n, stream):
        xt = rls.B200Vector(c, dt, n).fill_philox(SEED, stream=stream, dist=0)
        h = xt.to_numpy()
        h[np.arange(n) % 100 != 0] = 0                         # 1 % non-zeros
        xt.upload(h)
        return xt

    # ---- parity twin (N > 1): the sharded solve against the same system on ONE GPU -------------------------------
    parity = None
    if multi:
        lo, hi = rls.dist.row_range(TWIN_M, rank, world, align=4)
        sc = 1.0 / np.sqrt(TWIN_M)
        At = rls.B200Matrix.philox(dt, hi - lo, N_COLS, seed=SEED + 1, scale=sc, row_offset=lo, m_global=TWIN_M, ctx=ctx)
        xt = sparse_truth(ctx, N_COLS, 21)
        bt = At.mul(xt).to_numpy()
        rho_t = np.float32(0.05)
        xs_f = rls.solve_(rls.FISTA(At, reg=rls.L1Regularization(LAMBDA), iterations=20, rho=rho_t, relTol=0.0), bt)
        xs_c = rls.solve_(rls.CGNR(At, reg=rls.L2Regularization(LAMBDA), iterations=10, relTol=0.0), bt)
        parts = [None] * world
        dist.all_gather_object(parts, bt)
        if rank == 0:
            solo = rls.B200Context(local)                      # second context on the same GPU, no communicator
            Af = rls.B200Matrix.philox(dt, TWIN_M, N_COLS, seed=SEED + 1, scale=sc, ctx=solo)
            bf = np.concatenate(parts)
            x1_f = rls.solve_(rls.FISTA(Af, reg=rls.L1Regularization(LAMBDA), iterations=20, rho=rho_t, relTol=0.0, ctx=solo), bf)
            x1_c = rls.solve_(rls.CGNR(Af, reg=rls.L2Regularization(LAMBDA), iterations=10, relTol=0.0, ctx=solo), bf)
            rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
            # noise floor: the same single-GPU solve on the COLUMN-major copy (two-sweep gemv kernels) — other kernels, another
            # order of the same sums, no sharding involved
            try:
                Ac = rls.B200Matrix.philox(dt, TWIN_M, N_COLS, seed=SEED + 1, scale=sc, ctx=solo, layout="col")
                floor = rel(rls.solve_(rls.FISTA(Ac, reg=rls.L1Regularization(LAMBDA), iterations=20, rho=rho_t, relTol=0.0,
                                                 normal="twopass", ctx=solo), bf), x1_f)
                del Ac
            except Exception as e:                                  # informative only
                floor = f"not measured: {e}"
            f_rel, c_rel = rel(xs_f, x1_f), rel(xs_c, x1_c)
            parity = {"twin": f"{TWIN_M}x{N_COLS} ComplexF32, row-sharded over {world} GPUs vs the same system on one GPU",
                      "fista_l1_20_iterations_rel_l2": f_rel, "cgnr_10_iterations_rel_l2": c_rel,
                      "single_gpu_rowmajor_onepass_vs_colmajor_twosweep_fista_rel_l2": floor, "tolerance": TWIN_TOL,
                      "tolerance_is": "2 x 1e-5: the sharded and the single-GPU solve are two Float32 renderings of the same "
                                      "iteration that differ only in the order of the row sums (N partial vectors added by the "
                                      "all-reduce); each is held to 1e-5 against the oracle by tests/test_gpu_configs.py, so "
                                      "they may be 2e-5 apart",
                      "pass": bool(f_rel <= TWIN_TOL and c_rel <= TWIN_TOL)}
            del Af, solo
            if not parity["pass"]:                                  # reported in the line, never a hang: the other ranks wait in a collective
                print(f"bench.py: PARITY TWIN FAILED: {parity}", file=sys.stderr, flush=True)
        del At

    # ---- the workload: this rank's row block of the global system, generated on the device -------------------------
    lo, hi = rls.dist.row_range(M_GLOBAL, rank, world, align=4)
    m_loc = hi - lo
    scale = 1.0 / np.sqrt(M_GLOBAL)
    A = rls.B200Matrix.philox(dt, m_loc, N_COLS, seed=SEED, scale=scale, row_offset=lo, m_global=M_GLOBAL, ctx=ctx)
    xt = sparse_truth(ctx, N_COLS, 11)
    b_dev = A.mul(xt)
    noise = rls.B200Vector(ctx, dt, m_loc).fill_philox(SEED, stream=12, dist=1, scale=1e-3, offset=lo)
    b_host = (b_dev.to_numpy() + noise.to_numpy()).astype(np.complex64)
    b_dev.upload(b_host)
    AHA = rls.B200NormalOp(A, form=args.normal)
    b0 = rls.B200Vector(ctx, dt, N_COLS).fill_philox(SEED, stream=13, dist=1)
    lam_max = AHA.power_iterations(b0, rtol=1e-3, maxiter=10)
    rho = np.float32(0.95 / lam_max)
    S = rls.FISTA(A, AHA=AHA, reg=rls.L1Regularization(LAMBDA), iterations=FISTA_ITERS, rho=rho, relTol=0.0)
    Sc = rls.CGNR(A, AHA=AHA, reg=rls.L2Regularization(LAMBDA), iterations=CGNR_ITERS, relTol=0.0)
    it = C.c_int32()

    def solve_dev(solver):
        capi.call("rls_solver_solve", solver._handle, b_dev.handle, None, C.byref(it), C.byref(solver._scalars))

    warm = max(args.warmup, 3)
    for _ in range(warm):
        solve_dev(S)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    ms = timed(lambda: solve_dev(S), args.steps)
    launches = ctx.launch_count() - l0
    assert it.value == FISTA_ITERS
    its_per_s = args.steps * FISTA_ITERS / (ms * 1e-3)
    x_fista = S._vec("x").to_numpy()
    assert np.all(np.isfinite(x_fista)) and np.linalg.norm(x_fista) > 0

    # e2e: the user's call — host b (this rank's rows) in, host x out, every step
    x_host = None

    def solve_host():
        nonlocal x_host
        x_host = rls.solve_(S, b_host)
    for _ in range(2):
        solve_host()
    ms_e2e = timed(solve_host, args.steps)
    e2e_its = args.steps * FISTA_ITERS / (ms_e2e * 1e-3)
    assert np.array_equal(x_host, x_fista), "host-buffer solve and device-buffer solve differ"

    # CGNR on the same system
    for _ in range(2):
        solve_dev(Sc)
    ms_c = timed(lambda: solve_dev(Sc), args.steps)
    assert it.value == CGNR_ITERS
    cgnr_its = args.steps * CGNR_ITERS / (ms_c * 1e-3)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    clocks = sampler.summary()

    # roofline of the dominant kernel: the one-pass normal-operator apply on this rank's shard, timed alone on its stream
    xv = S._vec("x")
    res = rls.B200Vector(ctx, dt, N_COLS)
    reps = 20
    for _ in range(3):
        AHA.apply(xv, res)
    ms_k = timed(lambda: AHA.apply(xv, res), reps) / reps
    peak, peak_src = measured_peaks()
    alg_bytes = m_loc * N_COLS * 8
    achieved = alg_bytes / (ms_k * 1e-3) / 1e9
    total_bytes = M_GLOBAL * N_COLS * 8
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                per_byte = json.load(f).get("rowstream_complex_dram_bytes_per_algorithmic_byte")
            traffic = int(per_byte * alg_bytes) if per_byte else None
        except Exception:
            traffic = None

    # secondary: round 1's line — FISTA-L1 on one Float32 16384 x 65536 shard per GPU (weak scaling)
    secondary = None
    if not args.no_secondary:
        try:
            m2 = 16384
            A2 = rls.B200Matrix.philox(np.float32, m2, N_COLS, seed=SEED, scale=1.0 / np.sqrt(m2 * world), row_offset=rank * m2,
                                       m_global=m2 * world, ctx=ctx)
            x2 = rls.B200Vector(ctx, np.float32, N_COLS).fill_philox(SEED, stream=11, dist=0)
            h2 = x2.to_numpy(); h2[np.arange(N_COLS) % 100 != 0] = 0; x2.upload(h2)
            b2 = A2.mul(x2)
            op2 = rls.B200NormalOp(A2, form=args.normal)
            rho2 = np.float32(0.95 / op2.power_iterations(rls.B200Vector(ctx, np.float32, N_COLS).fill_philox(SEED, stream=13, dist=1), maxiter=30))
            S2 = rls.FISTA(A2, AHA=op2, reg=rls.L1Regularization(LAMBDA), iterations=200, rho=rho2, relTol=0.0)
            call2 = lambda: capi.call("rls_solver_solve", S2._handle, b2.handle, None, C.byref(it), C.byref(S2._scalars))
            for _ in range(3):
                call2()
            ms2 = timed(call2, 5)
            secondary = {"workload": "C2 shard per GPU: FISTA-L1, Float32 16384x65536, 200 iterations per solve (weak scaling, "
                                     "round-1 bench line)", "shard_iterations_per_s_per_gpu": 5 * 200 / (ms2 * 1e-3),
                         "ms_per_iteration": ms2 / 5 / 200, "frac_of_measured_hbm": m2 * N_COLS * 4 / (ms2 / 5 / 200 * 1e-3) / 1e9 / peak,
                         "normal_operator": op2.describe()}
            del S2, op2, A2
        except Exception as e:                                      # never at the price of the main line
            if multi:
                raise                                               # ranks must not part ways in front of a collective
            secondary = {"error": f"{type(e).__name__}: {e}"}

    # secondary: BASELINE config C4 — 64 frames sharing one ComplexF32 A 32768 x 16384 (MultiThreading.jl path) on the tensor
    # cores, A-form (two GEMMs per batched iteration) and Gram form (the reference's default AHA = A'*A: one GEMM over G)
    secondary_c4 = None
    if not args.no_secondary and world == 1:
        try:
            m4, n4, K4, its4 = 32768, 16384, 64, 50
            A4 = rls.B200Matrix.philox(dt, m4, n4, seed=4321, scale=1.0 / np.sqrt(m4), ctx=ctx)
            X4 = np.zeros((n4, K4), np.complex64, order="F")
            rng4 = np.random.default_rng(4)
            for k in range(K4):
                i4 = rng4.integers(0, n4, 160)
                X4[i4, k] = (rng4.random(160) + 1j * rng4.random(160)).astype(np.complex64)
            B4 = np.empty((m4, K4), np.complex64, order="F")
            for k in range(K4):
                B4[:, k] = A4.mul(rls.B200Vector.from_numpy(X4[:, k].copy(), ctx)).to_numpy()
            op4 = rls.B200NormalOp(A4, form="auto")
            rho4 = np.float32(0.95 / op4.power_iterations(rls.B200Vector(ctx, dt, n4).fill_philox(9, stream=1, dist=1)))
            secondary_c4 = {"workload": "C4: multi-RHS FISTA-L1, 64 frames sharing ComplexF32 A 32768x16384, 50 iterations per solve; one "
                                        "step = solve!(solver, B) with host B in and host X out (tcgen05 split-precision GEMMs)"}
            for form, opf in (("a_form", op4), ("gram_form", None)):
                ctx.sync()
                t0 = time.perf_counter()
                if opf is None:
                    opf = rls.B200NormalOp(A4, form="gram")
                    ctx.sync()
                    secondary_c4["gram_build_s"] = time.perf_counter() - t0
                S4 = rls.FISTA(A4, AHA=opf, reg=rls.L1Regularization(LAMBDA), iterations=its4, rho=rho4, relTol=0.0)
                Xs4 = None

                def solve4():
                    nonlocal Xs4
                    Xs4 = rls.solve_(S4, B4)
                solve4()                                            # lane allocation and the GEMM plans are one-time costs
                ms4 = timed(solve4, 3) / 3
                secondary_c4[form] = {"frame_iterations_per_s": K4 * its4 / (ms4 * 1e-3), "ms_per_batched_iteration": ms4 / its4,
                                      "rel_err_vs_truth": float(np.linalg.norm(Xs4 - X4) / np.linalg.norm(X4)),
                                      "normal_operator": opf.describe()}
                del S4
            del op4, opf, A4
        except Exception as e:                                      # a secondary line must never cost the bench its main line
            secondary_c4 = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        m_sub = 4096
        cpu_problem(m_sub)
        k0, dt0 = cpu_step("fista", m_sub, 4)
        k, dts = cpu_step("fista", m_sub, int(max(20, min(400, round(k0 / dt0 * 12.0)))))      # ~12 s of CPU work
        threads = blas_threads()
        cpu = {"value": k / dts * m_sub / M_GLOBAL, "unit": "iterations/s", "cores": threads, "kind": "port",
               "sample": f"{k} FISTA-L1 iterations of the oracle loop (NumPy + OpenBLAS cgemv pair, {threads} BLAS threads on "
                         f"{host_threads()} usable host cores of {os.cpu_count()}) in {dts:.1f} s on rows 0..{m_sub - 1} of the same Philox matrix "
                         f"({m_sub}x{N_COLS} ComplexF32); iterations/s scaled by {m_sub}/{M_GLOBAL} to the full system (DRAM-bound, "
                         "linear in m); restated reference — Julia is not installed"}
    if rank == 0:
        ms_it = ms / args.steps / FISTA_ITERS
        ms_it_c = ms_c / args.steps / CGNR_ITERS
        line = {
            "metric": METRIC,
            "value": its_per_s, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c64 (ComplexF32)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "normal_operator": AHA.form, "iterations_per_step": FISTA_ITERS,
                       "parallelism": "single GPU (all 137.4 GB in one B200)" if world == 1 else
                                      f"row-sharded x{world}: {m_loc} rows per GPU, one NCCL all-reduce of the 65536-vector per "
                                      "iteration, epilogues replicated",
                       "l2": f"inputs ({alg_bytes / 1e9:.1f} GB per GPU) are far larger than L2; no flush needed",
                       "ms_per_iteration": ms_it,
                       "frac_of_aggregate_measured_hbm": total_bytes / (ms_it * 1e-3) / 1e9 / (peak * world),
                       "frac_of_aggregate_nominal_8000": total_bytes / (ms_it * 1e-3) / 1e9 / (8000.0 * world)},
            "cgnr": {"value": cgnr_its, "unit": "iterations/s", "iterations_per_step": CGNR_ITERS, "ms_per_iteration": ms_it_c,
                     "frac_of_aggregate_measured_hbm": total_bytes / (ms_it_c * 1e-3) / 1e9 / (peak * world)},
            "e2e": {"value": e2e_its, "unit": "iterations/s", "h2d_bytes_per_step": int(b_host.nbytes) * world,
                    "d2h_bytes_per_step": int(x_host.nbytes) * world,
                    "note": "rls.solve_(solver, b_host) per step on every rank: pinned staging + H2D of the rank's rows of b, "
                            "D2H of x"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0, "traffic": traffic,
                         "kernel": f"normal operator A'(A x) on this rank's rows [{AHA.describe()}]",
                         "ms_per_launch": ms_k, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "matrix_layout": A.layout, "ms_per_iteration_in_solve": ms_it,
                         "note": "one launch = one normal-operator apply on the rank's row block, scored against a single read of "
                                 "it (rows x 65536 x 8 B), timed alone right after the solves (the clocks the power cap allows); "
                                 "per iteration a solve adds the epilogue kernel(s) and, at N > 1, the all-reduce"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if parity is not None:
            line["parity"] = parity
        if secondary is not None:
            line["secondary_c2"] = secondary
        if secondary_c4 is not None:
            line["secondary_c4"] = secondary_c4
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--normal", default="auto", choices=["auto", "twopass", "onepass"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    except BaseException:
        # a rank that dies inside a collective job must take the job down at once: the interpreter's normal exit would wait
        # in the NCCL / context destructors for peers that are themselves blocked in a collective (seen: 11 minutes)
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
