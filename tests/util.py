import numpy as np


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    nb = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0)


def rand_matrix(dtype, m, n, seed, scale=None):
    """Philox matrix, identical on host (oracle/philox.py) and device."""
    from oracle.philox import philox_matrix, IH4
    scale = 1.0 / np.sqrt(m) if scale is None else scale
    return philox_matrix(dtype, m, n, seed, IH4, scale), scale


def rand_vector(dtype, n, seed, stream=1):
    from oracle.philox import philox_vector, IH4
    return philox_vector(dtype, n, seed, stream, IH4, 1.0)


def sparse_truth(dtype, n, seed, every=29):
    from oracle.philox import philox_vector, UNIFORM01
    x = np.zeros(n, dtype)
    v = philox_vector(dtype, n, seed, 2, UNIFORM01)
    x[::every] = v[::every]
    return x


def up64(a):
    a = np.asarray(a)
    return a.astype(np.complex128 if a.dtype.kind == "c" else np.float64)


def stepwise_vs_fp64(S, R32, R64, b, iters, tol=1e-5, slack=1.5, each=None):
    """Per-iterate parity where the Float32 oracle's own rounding is of the size of the tolerance: reductions over 65536
    elements in OpenBLAS Float32, and solvers whose steering scalars (CG α, β; ADMM's inner cg!) amplify summation-order
    rounding.  Three runs on the same inputs: the CUDA path S, the oracle in the reference's arithmetic (Float32) R32, the
    oracle in Float64 R64.  At EVERY iterate either the CUDA iterate matches the Float32 oracle to `tol` (1e-5, the
    north-star tolerance), or it is as close to the Float64 recurrence as the Float32 oracle itself is:
        ‖x_gpu − x_64‖ ≤ slack · ‖x_o32 − x_64‖
    (slack 1.5: both sides are roundings of the same size, a strict ≤ at every single iterate would be a coin flip; measured
    on B200 the ratio is 0.73-0.84 in the mean, tools/parity_probe.py) — then the gpu-vs-oracle32 distance is the oracle's
    BLAS rounding, not an error of the CUDA path.  Stopping decisions must be identical.  `each(k)` runs extra checks per
    iterate.  Returns (worst gpu-o32, worst gpu-o64, worst o32-o64, iterates that needed the Float64 criterion)."""
    S.init_(b); R32.init(b); R64.init(up64(b))
    w = [0.0, 0.0, 0.0]
    needed = 0
    for k in range(iters + 2):
        a, r1 = S.iterate(), R32.iterate()
        R64.iterate()
        assert a == r1, f"stopping decision differs at iteration {k}: gpu={a} oracle={r1}"
        if not a:
            break
        if each is not None:
            each(k)
        x, x32, x64 = S.x, R32.x, R64.x
        e = (rel(x, x32), rel(x, x64), rel(x32, x64))
        w = [max(p, q) for p, q in zip(w, e)]
        if not e[0] <= tol:
            needed += 1
            assert e[1] <= slack * e[2], \
                f"iterate {k + 1}: gpu-o32 {e[0]:.2e} > {tol:g} and gpu-o64 {e[1]:.2e} > {slack:g} x o32-o64 {e[2]:.2e}"
    assert S.iteration == R32.iteration
    return w[0], w[1], w[2], needed


def to64(obj):
    """The Float64 twin of an oracle argument: Float32 scalars / arrays widened (same values), regularization terms with
    a widened λ, GradientOp with the wide element type; everything else unchanged."""
    import copy
    if isinstance(obj, np.ndarray):
        return up64(obj) if obj.dtype in (np.float32, np.complex64) else obj
    if isinstance(obj, (np.float32, np.complex64)):
        return np.float64(obj) if isinstance(obj, np.float32) else np.complex128(obj)
    if isinstance(obj, dict):
        return {k: to64(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to64(v) for v in obj)
    if type(obj).__name__ == "GradientOp":
        wide = np.complex128 if np.dtype(obj.dtype).kind == "c" else np.float64
        return type(obj)(wide, obj.shape, obj.dims)
    if hasattr(obj, "lam"):
        c = copy.copy(obj)
        c.lam = to64(obj.lam)
        return c
    if hasattr(obj, "reg") and hasattr(obj, "factor"):
        c = copy.copy(obj)
        c.reg = to64(obj.reg); c.factor = to64(obj.factor)
        return c
    return obj


def assert_close_or_fp64(x, x32, x64_fn, tol=1e-5, slack=1.5, what=""):
    """One iterate: rel-L2 <= tol against the Float32 oracle, or as close to the Float64 oracle as the Float32 oracle is
    (see stepwise_vs_fp64); x64_fn() is only evaluated when needed."""
    e = rel(x, x32)
    if e <= tol:
        return e
    x64 = x64_fn()
    e1, e2 = rel(x, x64), rel(x32, x64)
    assert e1 <= slack * e2, f"{what}: gpu-o32 {e:.2e} > {tol:g} and gpu-o64 {e1:.2e} > {slack:g} x o32-o64 {e2:.2e}"
    return e
