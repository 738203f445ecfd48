"""NumPy restatement of the RegularizedLeastSquares.jl iterative-solver hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference file:line it follows (paths relative to /root/reference).  The
restatement keeps the reference's arithmetic *types*: solver scalars live in
the real type ``rT`` of the problem (Float32 for Float32/ComplexF32 data),
every broadcast is an individually rounded elementwise operation, CGNR's
alpha/beta/zeta are complex-typed for complex data, and buffer
swaps/aliasing follow the reference.  Matrix products go to NumPy (OpenBLAS),
the same BLAS family Julia's ``mul!`` dispatches to.

Third-party arithmetic restated from published sources (not vendored in
/root/reference; only compat bounds exist in Project.toml:27-40):
  * IterativeSolvers.jl 0.9  ``cg!``        -> :func:`cg`
  * LinearOperatorCollection.jl 2 ``GradientOp`` -> :func:`grad_op`, :func:`grad_op_t`
  * LinearOperators.jl 2.3 5-arg ``mul!`` (res = a*Op*v + b*res, product
    rounded before the sum) -> used inline where the reference calls it.
"""
from __future__ import annotations

import itertools
import math
import numpy as np

__all__ = [
    "L1Regularization", "L2Regularization", "L21Regularization", "TVRegularization",
    "NuclearRegularization", "LLRRegularization", "prox_nuclear", "prox_llr",
    "PositiveRegularization", "RealRegularization", "NormalizedRegularization",
    "NoNormalization", "MeasurementBasedNormalization", "SystemMatrixBasedNormalization",
    "prox_", "reg_norm", "lam_of", "grad_op", "grad_op_t", "grad_t_axpy", "grad_rows", "cg", "power_iterations",
    "NormalOp", "FISTA", "CGNR", "POGM", "OptISTA", "ADMM", "SplitBregman", "Kaczmarz", "GradientOp",
    "createLinearSolver", "solve_", "normalize_factor", "enf_real", "enf_pos",
]


# ----------------------------------------------------------------------------
# type helpers
# ----------------------------------------------------------------------------
def real_type(dt):
    dt = np.dtype(dt)
    return {np.dtype(np.float32): np.float32, np.dtype(np.complex64): np.float32,
            np.dtype(np.float64): np.float64, np.dtype(np.complex128): np.float64}[dt]


def _as_julia_scalar(v):
    """A Python float literal is Float64 in Julia; NumPy scalars keep their type."""
    if isinstance(v, (np.floating, np.complexfloating)):
        return v
    if isinstance(v, (float, int)):
        return np.float64(v)
    return v


def _hypot_abs(x):
    """abs() of an array: modulus for complex (hypot evaluated in double for
    single precision, as both Julia's Base.hypot and glibc hypotf do)."""
    if np.iscomplexobj(x):
        if x.dtype == np.complex64:
            return np.sqrt(x.real.astype(np.float64) ** 2 + x.imag.astype(np.float64) ** 2).astype(np.float32)
        return np.hypot(x.real, x.imag)
    return np.abs(x)


def _cdiv(a, b):
    """Julia's generic Complex{T}/Complex{T} (Smith) division, base/complex.jl;
    real operands fall through to plain division."""
    if not (np.iscomplexobj(a) or np.iscomplexobj(b)):
        return a / b
    ct = np.result_type(a, b)
    rt = real_type(ct)
    a = ct.type(a); b = ct.type(b)
    are, aim, bre, bim = rt(a.real), rt(a.imag), rt(b.real), rt(b.imag)
    if abs(bre) <= abs(bim):
        r = bre / bim
        den = bim + r * bre
        return ct.type(complex((are * r + aim) / den, (aim * r - are) / den))
    r = bim / bre
    den = bre + r * bim
    return ct.type(complex((are + aim * r) / den, (aim - are * r) / den))


def _real_over_complex(a, z, dtype):
    """Julia: /(a::Real, z::Complex) = a*inv(z); inv(::ComplexF32) is evaluated in Float64
    (base/complex.jl).  Real dtypes: plain division."""
    dtype = np.dtype(dtype)
    if dtype.kind != "c":
        return dtype.type(a / z)
    if dtype == np.complex64:
        z = np.complex64(z)
        c, d = np.float64(z.real), np.float64(z.imag)
        mag = 1.0 / (c * c + d * d)
        inv = np.complex64(complex(c * mag, -d * mag))
        a = np.float32(a)
        return np.complex64(complex(a * np.float32(inv.real), a * np.float32(inv.imag)))
    return dtype.type(a / z)


def _norm2(x):
    return np.linalg.norm(x.ravel())


def _dot(a, b):
    """LinearAlgebra.dot: conj(a) . b"""
    return np.vdot(a, b)


# ----------------------------------------------------------------------------
# regularization terms (src/Regularization/*.jl, src/proximalMaps/*.jl)
# ----------------------------------------------------------------------------
class _Param:
    def __init__(self, lam):
        self.lam = _as_julia_scalar(lam)


class L1Regularization(_Param):
    """src/proximalMaps/ProxL1.jl:8-11"""


class L2Regularization(_Param):
    """src/proximalMaps/ProxL2.jl:8-11"""


class L21Regularization(_Param):
    """src/proximalMaps/ProxL21.jl:14-18"""
    def __init__(self, lam, slices=1):
        super().__init__(lam)
        self.slices = int(slices)


class TVRegularization(_Param):
    """src/proximalMaps/ProxTV.jl:32-39 (FGP only; see SURVEY 8a TV note)"""
    def __init__(self, lam, shape=(0,), dims=None, iterationsTV=10):
        super().__init__(lam)
        self.shape = tuple(int(s) for s in shape)
        self.dims = tuple(range(1, len(self.shape) + 1)) if dims is None else (
            (int(dims),) if np.isscalar(dims) else tuple(int(d) for d in dims))
        self.iterationsTV = int(iterationsTV)


class NuclearRegularization(_Param):
    """src/proximalMaps/ProxNuclear.jl:15-19"""
    def __init__(self, lam, svtShape=()):
        super().__init__(lam)
        self.svtShape = tuple(int(v) for v in svtShape)


class LLRRegularization(_Param):
    """src/proximalMaps/ProxLLR.jl:20-29.  `shift` replaces the draw `rand(block_idx)` of randshift (:55) so that a
    test can hand the same shift to both sides; None = no shift (randshift=false)."""
    def __init__(self, lam, shape=(0,), blockSize=None, randshift=False, fullyOverlapping=False, L=1, shift=None):
        super().__init__(lam)
        self.shape = tuple(int(v) for v in shape)
        self.blockSize = tuple(int(b) for b in blockSize) if blockSize is not None else (2,) * len(self.shape)
        self.randshift, self.fullyOverlapping, self.L = bool(randshift), bool(fullyOverlapping), int(L)
        self.shift = shift


class PositiveRegularization:
    """src/proximalMaps/ProxPositive.jl:8-9"""


class RealRegularization:
    """src/proximalMaps/ProxReal.jl:8-9"""


class NormalizedRegularization:
    """src/Regularization/NormalizedRegularization.jl:30-38"""
    def __init__(self, reg, factor):
        self.reg = reg
        self.factor = factor


class NoNormalization:
    pass


class MeasurementBasedNormalization:
    pass


class SystemMatrixBasedNormalization:
    pass


def _is_projection(reg):
    return isinstance(reg, (PositiveRegularization, RealRegularization))


def _sink(reg):
    while isinstance(reg, NormalizedRegularization):
        reg = reg.reg
    return reg


def lam_of(reg):
    """lambda(reg): Regularization.jl:29, ScaledRegularization.jl:23"""
    if isinstance(reg, NormalizedRegularization):
        return lam_of(reg.reg) * reg.factor
    if _is_projection(reg):
        return None
    return reg.lam


def normalize_factor(norm, A, b):
    """NormalizedRegularization.jl:40-59"""
    if isinstance(norm, NoNormalization):
        return None
    if isinstance(norm, MeasurementBasedNormalization):
        if b is None:
            return real_type(A.dtype)(1)
        rt = real_type(b.dtype)
        return rt(np.sum(_hypot_abs(b), dtype=rt) / rt(b.size))
    if isinstance(norm, SystemMatrixBasedNormalization):
        if A is None:
            raise ValueError("SystemMatrixBasedNormalization requires supplying A to the constructor of the solver")
        rt = real_type(A.dtype)
        M, N = A.shape
        energy = np.sqrt(np.sum((A.real.astype(rt) ** 2 + (A.imag.astype(rt) ** 2 if np.iscomplexobj(A) else 0)), axis=1, dtype=rt))
        return rt(_norm2(energy) ** 2 / rt(N))
    raise TypeError(norm)


def _normalize_reg(reg, factor):
    """NormalizedRegularization.jl:69-78"""
    if factor is None or _is_projection(reg):
        return reg
    if isinstance(reg, NormalizedRegularization):
        return NormalizedRegularization(reg.reg, factor)
    return NormalizedRegularization(reg, factor)


def enf_real(x):
    """Utils.jl:114-124"""
    if np.iscomplexobj(x):
        x.imag[...] = 0
    return x


def enf_pos(x):
    """Utils.jl:129-144 (complex: re<0 -> im*i; real: max(x,0))"""
    if np.iscomplexobj(x):
        neg = x.real < 0
        x.real[neg] = 0
    else:
        x[x < 0] = 0
    return x


# ---- GradientOp (LinearOperatorCollection.jl GradientOp.jl; restated) --------
def grad_rows(shape, dims):
    shape = tuple(shape)
    tot = int(np.prod(shape))
    return sum((shape[d - 1] - 1) * tot // shape[d - 1] for d in dims)


def grad_op(img, shape, dims):
    """res_d[i] = img[i] - img[i+e_d], stacked over dims; column-major (Julia) shape."""
    im = img.reshape(shape, order="F")
    out = []
    for d in dims:
        a = d - 1
        lo = [slice(None)] * len(shape); hi = [slice(None)] * len(shape)
        lo[a] = slice(0, shape[a] - 1); hi[a] = slice(1, shape[a])
        out.append((im[tuple(lo)] - im[tuple(hi)]).ravel(order="F"))
    return np.concatenate(out) if out else np.zeros(0, img.dtype)


def grad_op_t(g, shape, dims):
    """adjoint/transpose: res[i] = g[i]; res[i+e_d] -= g[i]; summed over dims
    (vcat of operators -> sum of the per-dim adjoints, accumulated dim by dim)."""
    tot = int(np.prod(shape))
    res = None
    off = 0
    for d in dims:
        a = d - 1
        sh = list(shape); sh[a] -= 1
        cnt = int(np.prod(sh))
        gd = g[off:off + cnt].reshape(sh, order="F")
        off += cnt
        r = np.zeros(shape, dtype=g.dtype, order="F")
        lo = [slice(None)] * len(shape); hi = [slice(None)] * len(shape)
        lo[a] = slice(0, shape[a] - 1); hi[a] = slice(1, shape[a])
        r[tuple(lo)] = gd
        r[tuple(hi)] -= gd
        r = r.ravel(order="F")
        res = r if res is None else res + r
    return res if res is not None else np.zeros(tot, g.dtype)


def grad_t_axpy(a, g, shape, dims, res):
    """5-arg mul!(res, transpose(vcat(A_1..A_k)), g, a, 1) as LinearOperators composes it
    (cat.jl vcat_ctprod!): res = a*(A_d' g_d) + res, block after block, each product
    rounded before the sum."""
    off = 0
    res = res.copy()
    for d in dims:
        a_ = d - 1
        sh = list(shape); sh[a_] -= 1
        cnt = int(np.prod(sh))
        r = grad_op_t(g[off:off + cnt], shape, (d,))
        off += cnt
        res = a * r + res
    return res


class GradientOp:
    """regTrafo for ADMM (ADMM.jl:74): forward differences on `shape` along `dims`."""
    def __init__(self, dtype, shape, dims=None):
        self.shape = tuple(int(s) for s in shape)
        self.dims = tuple(range(1, len(self.shape) + 1)) if dims is None else tuple(dims)
        self.dtype = np.dtype(dtype)
        self.rows = grad_rows(self.shape, self.dims)
        self.cols = int(np.prod(self.shape))

    def mul(self, x):
        return grad_op(x, self.shape, self.dims)

    def tmul(self, g):
        return grad_op_t(g, self.shape, self.dims)


class _Eye:
    def __init__(self, n):
        self.rows = self.cols = n

    def mul(self, x):
        return x.copy()

    def tmul(self, g):
        return g.copy()


# ---- proximal maps -------------------------------------------------------------
def prox_l1(x, lam):
    """ProxL1.jl:18-22: max(|x|-lam,0) * (x+eps)/(|x|+eps); eps added to the real part."""
    T = real_type(x.dtype)
    lam = T(lam)
    eps = np.finfo(T).eps
    ax = _hypot_abs(x)
    x[...] = np.maximum(ax - lam, T(0)) * (x + eps) / (ax + eps)
    return x


def prox_l2(x, lam):
    """ProxL2.jl:18-21: x *= 1/(1+2lam), factor evaluated in Float64."""
    T = real_type(x.dtype)
    lam = T(lam)
    f = 1.0 / (1.0 + 2.0 * np.float64(lam))
    x[...] = (x * f).astype(x.dtype)
    return x


def prox_l21(x, lam, slices):
    """ProxL21.jl:30-35: groups are the strided sets x[j::L]; 0/0 -> NaN kept."""
    T = real_type(x.dtype)
    lam = T(lam)
    L = x.size // slices
    if L == 0:
        if x.size == 0:
            return x
        raise ValueError("step cannot be zero")          # x[i:0:end] throws upstream
    xv = x.reshape(-1)
    # group j = x[j:L:end]: `slices` elements, one more where length(x) is not a multiple of slices
    g = np.array([_norm2(xv[j::L]) for j in range(L)], dtype=T)
    with np.errstate(invalid="ignore", divide="ignore"):
        f = np.maximum((g - lam) / g, T(0))
    for j in range(L):
        xv[j::L] = xv[j::L] * f[j]
    return x


def prox_tv_fgp(x, lam, shape, dims, iterationsTV=10):
    """ProxTV.jl:89-125 Fast Gradient Projection with the reference's buffer
    rotation (pq aliases rs during the gradient step)."""
    T = real_type(x.dtype)
    lam = T(lam)
    rows = grad_rows(shape, dims)
    pq = np.zeros(rows, x.dtype); rs = np.zeros(rows, x.dtype); pqOld = np.zeros(rows, x.dtype)
    t = T(1)
    neg_lam = -lam
    inv8 = T(1) / (T(8) * lam)
    for _ in range(iterationsTV):
        pqTmp = pqOld
        pqOld = pq
        pq = rs
        # xTmp = x; mul!(xTmp, grad', rs, -lam, 1)
        xTmp = grad_t_axpy(neg_lam, rs, shape, dims, x)
        # mul!(pq, grad, xTmp, 1/(8lam), 1)   (pq is rs's storage)
        pq[...] = inv8 * grad_op(xTmp, shape, dims) + pq
        # restrict magnitude (per component)
        pq[...] = pq / np.maximum(T(1), _hypot_abs(pq))
        tOld = t
        t = (T(1) + np.sqrt(T(1) + T(4) * (tOld * tOld))) / T(2)
        t2 = (tOld - T(1)) / t
        t3 = T(1) + t2
        rs = pqTmp
        rs[...] = t3 * pq - t2 * pqOld
    x[...] = grad_t_axpy(neg_lam, pq, shape, dims, x)
    return x


def prox_nuclear(x, lam, svtShape):
    """ProxNuclear.jl:27-32: U,S,V = svd(reshape(x, svtShape)); prox!(L1Regularization, S, lam); x = vec(U*Diagonal(S)*V').
    LAPACK gesdd in the element type, as Julia's `svd`."""
    T = real_type(x.dtype)
    X = x.reshape(svtShape, order="F")
    U, S, Vh = np.linalg.svd(X, full_matrices=False)
    S = prox_l1(S.astype(T), lam)
    x[...] = ((U * S) @ Vh).astype(x.dtype).reshape(-1, order="F")
    return x


def _llr_non_overlapping(xs, lam, shape, blockSize):
    """ProxLLR.jl:44-90 on the (already shifted) array xs of size shape x K, in place."""
    T = real_type(xs.dtype)
    lam = T(lam)
    K = xs.shape[-1]
    for origin in itertools.product(*[range(0, shape[d], blockSize[d]) for d in range(len(shape))][::-1]):
        origin = origin[::-1]                                   # CartesianIndices iterate the first dimension fastest
        # idx = idx_center .+ block_idx (1-based offsets 1..blockSize on 0-based centres), cut at the image border (:60-63)
        sl = tuple(slice(origin[d], min(origin[d] + blockSize[d], shape[d])) for d in range(len(shape)))
        blk = xs[sl]
        X = np.zeros((int(np.prod(blockSize)), K), dtype=xs.dtype)
        npx = int(np.prod(blk.shape[:-1]))
        X[:npx, :] = blk.reshape((npx, K), order="F")
        if np.any(X != 0):
            G = (X.conj().T @ X).astype(xs.dtype)                # mul!(x2, x', x)
            ub = np.sqrt(np.max(_hypot_abs(G)))                  # sqrt(norm(x2, Inf)): Julia's norm of a matrix = largest |entry|
            if lam >= ub:
                xs[sl] = 0
            else:
                U, S, Vh = np.linalg.svd(X, full_matrices=False)
                S = prox_l1(S.astype(T), lam)
                Vh = Vh * S[:, None]                             # SVDec.Vt .*= SVDec.S
                X = (U @ Vh).astype(xs.dtype)
                xs[sl] = X[:npx, :].reshape(blk.shape, order="F")
    return xs


def prox_llr(x, lam, shape, blockSize, shift=None, fullyOverlapping=False):
    """ProxLLR.jl:36-38.  shift = the draw of `rand(block_idx)` (:55; None = randshift off).  ShiftedArrays.circshift is a
    lazy view (writes go through to x); here: shift a copy, threshold, shift back."""
    nd = len(shape)
    K = x.size // int(np.prod(shape))
    X = x.reshape(tuple(shape) + (K,), order="F")
    sh = tuple(int(v) for v in shift) if shift is not None else (0,) * nd
    ax = tuple(range(nd))
    if not fullyOverlapping:
        xs = np.roll(X, sh, axis=ax)                             # circshift(x, s): xs[i] = x[i - s]
        _llr_non_overlapping(xs, lam, shape, blockSize)
        x[...] = np.roll(xs, tuple(-v for v in sh), axis=ax).reshape(-1, order="F")
        return x
    # ProxLLR.jl:163-199; the padded branch (:170-173) reshapes the padded array with the unpadded shape inside the
    # nested call (:50) and throws, so only shapes that are multiples of blockSize exist
    if any(shape[d] % blockSize[d] for d in range(nd)):
        raise ValueError("DimensionMismatch: fully overlapping LLR needs shape to be a multiple of blockSize")
    xp = X.copy()
    out = np.zeros_like(X)
    for is_ in itertools.product(*[range(1, blockSize[d] + 1) for d in range(nd)][::-1]):
        is_ = is_[::-1]
        tot = tuple(is_[d] + sh[d] for d in range(nd))           # the nested call applies its own randshift on top (:55)
        xs = np.roll(xp, tot, axis=ax)
        _llr_non_overlapping(xs, lam, shape, blockSize)
        out = out + np.roll(xs, tuple(-v for v in tot), axis=ax)
    out = out / real_type(x.dtype)(int(np.prod(blockSize)))
    x[...] = out.astype(x.dtype).reshape(-1, order="F")
    return x


def prox_(reg, x, lam=None):
    """prox!(reg, x[, lam]) dispatch: Regularization.jl:17,31; NestedRegularization.jl:26-29"""
    if isinstance(reg, type):
        raise TypeError("construct the regularization term first")
    if isinstance(reg, NormalizedRegularization):
        return prox_(reg.reg, x, lam_of(reg) if lam is None else lam)
    if isinstance(reg, PositiveRegularization):
        enf_real(x); enf_pos(x); return x
    if isinstance(reg, RealRegularization):
        enf_real(x); return x
    if lam is None:
        lam = reg.lam
    if isinstance(reg, L1Regularization):
        return prox_l1(x, lam)
    if isinstance(reg, L2Regularization):
        return prox_l2(x, lam)
    if isinstance(reg, L21Regularization):
        return prox_l21(x, lam, reg.slices)
    if isinstance(reg, TVRegularization):
        return prox_tv_fgp(x, lam, reg.shape, reg.dims, reg.iterationsTV)
    if isinstance(reg, NuclearRegularization):
        return prox_nuclear(x, lam, reg.svtShape)
    if isinstance(reg, LLRRegularization):
        return prox_llr(x, lam, reg.shape, reg.blockSize, reg.shift, reg.fullyOverlapping)
    raise TypeError(reg)


def reg_norm(reg, x, lam=None):
    """norm(reg, x, lam): ProxL1.jl:29-32, ProxL2.jl:28, ProxL21.jl:42-46, ProxTV.jl:152-155"""
    reg0 = _sink(reg)
    if lam is None:
        lam = lam_of(reg)
    if isinstance(reg0, L1Regularization):
        return lam * np.sum(_hypot_abs(x))
    if isinstance(reg0, L2Regularization):
        return lam * _norm2(x) ** 2
    if isinstance(reg0, L21Regularization):
        L = x.size // reg0.slices
        xv = x.reshape(-1)
        return lam * sum(_norm2(xv[j::L]) for j in range(L))
    if isinstance(reg0, TVRegularization):
        return lam * np.sum(_hypot_abs(grad_op(x, reg0.shape, reg0.dims)))
    if isinstance(reg0, NuclearRegularization):                  # ProxNuclear.jl:39-42
        return lam * np.sum(np.linalg.svd(x.reshape(reg0.svtShape, order="F"), compute_uv=False))
    raise TypeError(reg)


# ----------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------
def adjoint_mul(A, y):
    """mul!(x, adjoint(A), y) = BLAS gemv('C'): A' y without materialising conj(A).  For complex A NumPy's
    `A.conj().T @ y` first writes a conjugated copy of the whole matrix; conj(Aᵀ conj(y)) is the same sum term by term
    (conjugation is exact) through the transposed gemv on A itself — what Julia's BLAS call does."""
    if np.iscomplexobj(A):
        return np.conj(A.T @ np.conj(y))
    return A.T @ y


class NormalOp:
    """AHA.  mode='lazy': A'(A x) as two gemv (LinearOperatorCollection.normalOperator);
    mode='gram': the materialised A'*A the reference builds by default for a dense
    Matrix (FISTA.jl:58, CGNR.jl:49, ADMM.jl:81, POGM.jl:76, OptISTA.jl:63)."""
    def __init__(self, A=None, G=None, mode="lazy"):
        self.A = A
        self.mode = mode
        if G is not None:
            self.G = G; self.mode = "gram"
        elif mode == "gram":
            self.G = A.conj().T @ A
        self.dtype = (self.G if self.mode == "gram" else A).dtype
        self.n = (self.G if self.mode == "gram" else A).shape[1]

    def apply(self, x):
        if self.mode == "gram":
            return self.G @ x
        return adjoint_mul(self.A, self.A @ x)


def _make_normal(A, AHA, normal):
    if AHA is None:
        return NormalOp(A=A, mode=normal)
    if isinstance(AHA, NormalOp):
        return AHA
    return NormalOp(G=np.asarray(AHA))


def power_iterations(AHA, b0=None, rtol=1e-3, maxiter=30, rng=None):
    """Utils.jl:262-287.  `b0` injects the start vector (the reference draws randn)."""
    if not isinstance(AHA, NormalOp):
        AHA = NormalOp(G=np.asarray(AHA))
    n = AHA.n
    if b0 is None:
        rng = np.random.default_rng(0) if rng is None else rng
        b0 = rng.standard_normal(n).astype(real_type(AHA.dtype))
        if np.issubdtype(AHA.dtype, np.complexfloating):
            b0 = (b0 + 1j * rng.standard_normal(n)).astype(AHA.dtype)
    b = np.array(b0, dtype=AHA.dtype, copy=True)
    lam = np.inf
    for _ in range(maxiter):
        b = b / _norm2(b)
        bold = b
        b = AHA.apply(bold)
        lam_old = lam
        lam = abs(_dot(bold, b))
        if abs(lam / lam_old - 1) < rtol:
            return lam
    return lam


def cg(x, op, b, maxiter, reltol, abstol=0.0):
    """IterativeSolvers.jl 0.9 cg!(x, A, b; Pl=Identity(), maxiter, reltol, statevars)
    (cg.jl: cg_iterator!, iterate(::CGIterable)).  Warm start; returns #iterations."""
    T = real_type(x.dtype)
    u = np.zeros_like(x)
    r = b.copy()
    c = op(x)
    r -= c
    residual = T(_norm2(r))
    tol = max(T(reltol) * residual, T(abstol))
    prev = T(1)
    k = 0
    while k < maxiter and not (residual <= tol):
        beta = residual * residual / (prev * prev)
        u[...] = r + beta * u
        c = op(u)
        alpha = _real_over_complex(residual * residual, _dot(u, c), x.dtype)
        x += alpha * u
        r -= alpha * c
        prev = residual
        residual = T(_norm2(r))
        k += 1
    return k


# ----------------------------------------------------------------------------
# solvers
# ----------------------------------------------------------------------------
def _split_regs(reg, solver_name, exactly_one=True):
    regs = list(reg) if isinstance(reg, (list, tuple)) else [reg]
    proj = [r for r in regs if _is_projection(_sink(r))]
    regs = [r for r in regs if not _is_projection(_sink(r))]
    if exactly_one and len(regs) != 1:
        raise ValueError(f"{solver_name} does not allow for more additional regularization terms, found {len(regs)}")
    return regs, proj


class _Base:
    def _adjoint_b(self, b):
        if self.A is None:
            return np.array(b, dtype=self.T, copy=True)
        return adjoint_mul(self.A, b).astype(self.T, copy=False)

    def solve(self, b, callbacks=None, **kw):
        """solve!(solver, b; callbacks) RegularizedLeastSquares.jl:103-117 (vector b) and
        MultiThreading.jl:30-80 (matrix b, sequential scheduler)."""
        b = np.asarray(b)
        if b.ndim == 2:
            cols = []
            for i in range(b.shape[1]):
                cols.append(self.solve(b[:, i].copy(), **kw).copy())
            return np.stack(cols, axis=1)
        cbs = [] if callbacks is None else (callbacks if isinstance(callbacks, (list, tuple)) else [callbacks])
        self.init(b, **kw)
        for cb in cbs:
            cb(self, 0)
        it = 0
        while self.iterate():
            it += 1
            for cb in cbs:
                cb(self, it)
        return self.x


class FISTA(_Base):
    """src/FISTA.jl:57-189"""
    def __init__(self, A, *, AHA=None, reg=None, normalizeReg=None, iterations=50, verbose=False,
                 rho=None, theta=1, relTol=None, restart="none", normal="lazy", rho_start=None):
        self.A = A
        self.AHA = _make_normal(A, AHA, normal)
        self.T = np.dtype(self.AHA.dtype)
        rT = self.rT = real_type(self.T)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        if reg is None:
            reg = L1Regularization(rT(0))
        regs, self.proj = _split_regs(reg, type(self).__name__)
        f = normalize_factor(self.normalizeReg, A, None)
        self.reg = _normalize_reg(regs[0], f)
        if rho is None:
            rho = 0.95 / power_iterations(self.AHA, rho_start)
        self.rho = rT(rho)
        self.theta0 = rT(theta)
        self.relTol = rT(np.finfo(rT).eps if relTol is None else relTol)
        self.restart = restart
        self.iterations = int(iterations)
        self.verbose = verbose
        n = self.AHA.n
        self.x = np.zeros(n, self.T); self.x0 = np.zeros(n, self.T)
        self.xold = np.zeros(n, self.T); self.res = np.zeros(n, self.T)
        self.iteration = 0
        self.rel_res_norm = rT(np.inf)

    def _init_common(self, b, x0):
        rT = self.rT
        self.x0 = self._adjoint_b(b)                      # FISTA.jl:111-115
        self.norm_x0 = rT(_norm2(self.x0))                # :118
        self.x = np.zeros(self.AHA.n, self.T); self.x[...] = x0   # :120
        self.res = np.full(self.AHA.n, np.inf, self.T)    # :123
        self.rel_res_norm = rT(np.inf)
        self.iteration = 0
        if isinstance(self.normalizeReg, MeasurementBasedNormalization):   # :128 (passes x0 = A'b)
            self.reg = _normalize_reg(self.reg, normalize_factor(self.normalizeReg, self.A, self.x0))

    def init(self, b, x0=0, theta=1):
        self._init_common(b, x0)
        self.xold = np.zeros(self.AHA.n, self.T)          # :121
        self.theta = self.rT(theta); self.thetaold = self.rT(theta)

    def done(self):
        return bool(self.rel_res_norm < self.relTol) or self.iteration >= self.iterations

    def iterate(self):
        rT = self.rT
        if self.done():
            return False
        # momentum (FISTA.jl:144-148)
        self.x, self.xold = self.xold, self.x
        self.x *= (1 - self.thetaold) / self.theta
        self.x += ((self.thetaold - 1) / self.theta + 1) * self.xold
        # gradient step (:152-156)
        self.res = self.AHA.apply(self.x).astype(self.T, copy=False)
        self.res -= self.x0
        self.x -= self.rho * self.res
        self.rel_res_norm = rT(rT(_norm2(self.res)) / self.norm_x0)
        # prox (:164-168)
        prox_(self.reg, self.x, self.rho * lam_of(self.reg))
        for p in self.proj:
            prox_(p, self.x)
        # restart (:171-176)
        if self.restart == "gradient":
            if np.real(_dot(self.res, self.x - self.xold)) > 0:
                self.theta = rT(1)
        # :179-180
        self.thetaold = self.theta
        self.theta = (1 + np.sqrt(1 + 4 * (self.thetaold * self.thetaold))) / 2
        self.iteration += 1
        return True

    def convergence(self):
        return {"residual": _norm2(self.res)}


class POGM(FISTA):
    """src/POGM.jl:75-241"""
    def __init__(self, A, *, sigma_fac=1, **kw):
        super().__init__(A, **kw)
        rT = self.rT
        self.sigma_fac = rT(sigma_fac)
        self.alpha = rT(0); self.beta = rT(1); self.gamma = rT(1); self.gammaold = rT(1); self.sigma = rT(1)
        n = self.AHA.n
        self.y = np.zeros(n, self.T); self.z = np.zeros(n, self.T); self.w = np.zeros(n, self.T)

    def init(self, b, x0=0, theta=1):
        self._init_common(b, x0)
        n = self.AHA.n
        self.xold = np.zeros(n, self.T)
        self.y = np.zeros(n, self.T); self.z = np.zeros(n, self.T)
        if self.restart != "none":
            self.w = np.zeros(n, self.T)
        self.theta = self.rT(theta); self.thetaold = self.rT(theta)
        self.sigma = self.rT(1)                            # gamma NOT reset (POGM.jl:138-164)

    def iterate(self):
        rT = self.rT
        if self.done():
            return False
        self.xold[...] = self.x                            # :180
        self.res = self.AHA.apply(self.x).astype(self.T, copy=False)
        self.res -= self.x0
        self.x -= self.rho * self.res
        self.rel_res_norm = rT(rT(_norm2(self.res)) / self.norm_x0)
        # inertial parameters (:189-202)
        self.thetaold = self.theta
        th2 = self.thetaold * self.thetaold
        if self.iteration == self.iterations - 1 and self.restart != "none":
            self.theta = (1 + np.sqrt(1 + 8 * th2)) / 2
        else:
            self.theta = (1 + np.sqrt(1 + 4 * th2)) / 2
        self.alpha = (self.thetaold - 1) / self.theta
        self.beta = self.sigma * self.thetaold / self.theta
        self.gammaold = self.gamma
        if self.restart == "gradient":
            self.gamma = self.rho * (1 + self.alpha + self.beta)
        else:
            self.gamma = self.rho * (2 * self.thetaold + self.theta - 1) / self.theta
        # inertia (:206-213)
        self.x, self.y = self.y, self.x
        self.x *= -self.alpha
        self.x += (1 + self.alpha + self.beta) * self.y
        self.x -= (self.beta + self.rho * self.alpha / self.gammaold) * self.xold
        self.x += (self.rho * self.alpha / self.gammaold) * self.z
        self.z[...] = self.x
        # prox (:216-219)
        prox_(self.reg, self.x, self.gamma * lam_of(self.reg))
        for p in self.proj:
            prox_(p, self.x)
        # restart (:222-232)
        if self.restart == "gradient":
            self.w += self.y + (self.rho / self.gamma) * (self.x - self.z)
            if np.real((_dot(self.w, self.x) - _dot(self.w, self.z)) / self.gamma - _dot(self.w, self.res)) < 0:
                self.sigma = rT(1); self.theta = rT(1)
            else:
                self.sigma = self.sigma * self.sigma_fac
            self.w[...] = (self.rho / self.gamma) * (self.z - self.x) - self.y
        self.iteration += 1
        return True


class OptISTA(FISTA):
    """src/OptISTA.jl:62-207 (projections are stored but never applied, quirk 10)"""
    def __init__(self, A, **kw):
        kw.pop("restart", None)
        super().__init__(A, **kw)
        n = self.AHA.n
        self.y = np.zeros(n, self.T); self.z = np.zeros(n, self.T); self.zold = np.zeros(n, self.T)

    def init(self, b, x0=0, theta=1):
        rT = self.rT
        self._init_common(b, x0)
        self.y = self.x.copy(); self.z = self.x.copy(); self.zold = self.x.copy()
        self.theta = rT(theta); self.thetaold = rT(theta)
        tn = rT(theta)
        for _ in range(self.iterations - 1):                # OptISTA.jl:145-149
            tn = (1 + np.sqrt(1 + 4 * (tn * tn))) / 2
        tn = (1 + np.sqrt(1 + 8 * (tn * tn))) / 2
        self.thetan = rT(tn)
        self.alpha = rT(0); self.beta = rT(1); self.gamma = rT(1)

    def iterate(self):
        rT = self.rT
        if self.done():
            return False
        th = self.theta; tn2 = self.thetan * self.thetan
        self.gamma = 2 * th / tn2 * (tn2 - 2 * (th * th) + th)          # :168
        self.thetaold = th
        if self.iteration == self.iterations - 1:
            self.theta = (1 + np.sqrt(1 + 8 * (th * th))) / 2
        else:
            self.theta = (1 + np.sqrt(1 + 4 * (th * th))) / 2
        self.alpha = (self.thetaold - 1) / self.theta
        self.beta = self.thetaold / self.theta
        self.zold[...] = self.z                             # :180-184
        self.z[...] = self.y
        self.res = self.AHA.apply(self.x).astype(self.T, copy=False)
        self.res -= self.x0
        self.y -= (self.rho * self.gamma) * self.res
        self.rel_res_norm = rT(rT(_norm2(self.res)) / self.norm_x0)
        prox_(self.reg, self.y, self.rho * self.gamma * lam_of(self.reg))   # :190
        self.z /= -self.gamma                               # :195-199
        self.z += self.x + self.y / self.gamma
        self.x *= -self.beta
        self.x += (1 + self.alpha + self.beta) * self.z
        self.x -= self.alpha * self.zold
        self.iteration += 1
        return True


class CGNR(_Base):
    """src/CGNR.jl:48-185"""
    def __init__(self, A, *, AHA=None, reg=None, normalizeReg=None, iterations=10, relTol=None, normal="lazy"):
        self.A = A
        self.AHA = _make_normal(A, AHA, normal)
        self.T = np.dtype(self.AHA.dtype)
        rT = self.rT = real_type(self.T)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        regs = [] if reg is None else (list(reg) if isinstance(reg, (list, tuple)) else [reg])
        f = normalize_factor(self.normalizeReg, A, None)
        regs = [_normalize_reg(r, f) for r in regs]
        l2 = [r for r in regs if isinstance(_sink(r), L2Regularization)]
        if len(l2) > 1:
            raise ValueError("Cannot unambigiously retrieve reg term of type L2Regularization")
        self.L2 = l2[0] if l2 else L2Regularization(self.T.type(0))
        rest = [r for r in regs if not isinstance(_sink(r), L2Regularization)]
        self.constr = [r for r in rest if isinstance(_sink(r), RealRegularization)] + \
                      [r for r in rest if isinstance(_sink(r), PositiveRegularization)]
        rest = [r for r in rest if not _is_projection(_sink(r))]
        if rest:
            raise ValueError(f"CGNR does not allow for more additional regularization terms, found {len(rest)}")
        self.iterations = int(iterations)
        self.relTol = rT(np.finfo(rT).eps if relTol is None else relTol)
        n = self.AHA.n
        self.x = np.zeros(n, self.T); self.x0 = np.zeros(n, self.T)
        self.pl = np.zeros(n, self.T); self.vl = np.zeros(n, self.T)
        self.iteration = 0; self.z0 = rT(0)

    def init(self, b, x0=0):
        n = self.AHA.n
        if np.any(np.asarray(x0) != 0):
            raise NotImplementedError("CGNR x0 != 0 path is broken in the reference (CGNR.jl:119)")
        self.pl = np.zeros(n, self.T); self.vl = np.zeros(n, self.T)
        self.alpha = self.T.type(0); self.beta = self.T.type(0); self.zeta = self.T.type(0)
        self.iteration = 0
        self.x = np.zeros(n, self.T)
        self.x0 = self._adjoint_b(b)                       # :123
        self.z0 = self.rT(_norm2(self.x0))                 # :125
        self.pl[...] = self.x0                             # :126
        if isinstance(self.normalizeReg, MeasurementBasedNormalization):   # :129 (passes b)
            self.L2 = _normalize_reg(self.L2, normalize_factor(self.normalizeReg, self.A, b))

    def converged(self):
        return bool(self.rT(_norm2(self.x0)) / self.z0 <= self.relTol)

    def done(self):
        return self.converged() or self.iteration >= min(self.iterations, self.AHA.n)

    def iterate(self):
        rT = self.rT; Tc = self.T.type
        if self.done():
            for r in self.constr:
                prox_(r, self.x)
            return False
        self.vl = self.AHA.apply(self.pl).astype(self.T, copy=False)    # :151
        nr = rT(_norm2(self.x0))
        self.zeta = Tc(nr * nr)                            # :153
        normvl = _dot(self.pl, self.vl)                    # :154
        lam = lam_of(self.L2)
        if lam > 0:
            npl = rT(_norm2(self.pl))
            self.alpha = Tc(_cdiv(self.zeta, normvl + lam * (npl * npl)))
        else:
            self.alpha = Tc(_cdiv(self.zeta, normvl))
        self.x += self.pl * self.alpha                     # :163
        self.x0 += self.vl * (-self.alpha)                 # :165
        if lam > 0:
            self.x0 += ((self.pl * (-lam)) * self.alpha).astype(self.T)   # :168
        self.beta = Tc(_cdiv(_dot(self.x0, self.x0), self.zeta))        # :171
        self.pl *= self.beta                               # :173-174
        self.pl += self.x0
        self.iteration += 1
        return True

    def convergence(self):
        return {"residual": _norm2(self.x0)}


class ADMM(_Base):
    """src/ADMM.jl:80-332 (precon = Identity only)"""
    def __init__(self, A, *, AHA=None, reg=None, regTrafo=None, normalizeReg=None, rho=1e-1, vary_rho="none",
                 iterations=10, iterationsCG=10, absTol=None, relTol=None, tolInner=1e-5, verbose=False,
                 normal="lazy"):
        self.A = A
        self.AHA = _make_normal(A, AHA, normal)
        self.T = np.dtype(self.AHA.dtype)
        rT = self.rT = real_type(self.T)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        if reg is None:
            reg = L1Regularization(rT(0))
        regs, self.proj = _split_regs(reg, "ADMM", exactly_one=False)
        n = self.AHA.n
        if regTrafo is None:
            trafo = [_Eye(n) for _ in regs]
        else:
            trafo = list(regTrafo) if isinstance(regTrafo, (list, tuple)) else [regTrafo]
            trafo = [_Eye(n) if t is None else t for t in trafo]
        assert len(regs) == len(trafo), "reg and regTrafo must have the same length"
        self.regTrafo = trafo
        if np.isscalar(rho):
            self.rho = np.array([rT(rho) for _ in regs], dtype=rT)
        else:
            self.rho = np.asarray(rho, dtype=rT).copy()
        f = normalize_factor(self.normalizeReg, A, None)
        self.reg = [_normalize_reg(r, f) for r in regs]
        self.vary_rho = vary_rho
        self.iterations = int(iterations); self.iterationsCG = int(iterationsCG)
        eps = np.finfo(rT).eps
        self.absTol = rT(eps if absTol is None else absTol)
        self.relTol = rT(eps if relTol is None else relTol)
        self.tolInner = rT(tolInner)
        self.verbose = verbose
        self.x = np.zeros(n, self.T)
        self.iteration = 0
        self.cg_iters = []

    def init(self, b, x0=0):
        rT = self.rT; n = self.AHA.n; k = len(self.reg)
        self.x = np.zeros(n, self.T); self.x[...] = x0     # :192
        self.beta_y = self._adjoint_b(b)                   # :195-199
        self.z = [np.asarray(self.regTrafo[i].mul(self.x), dtype=self.T) for i in range(k)]
        self.u = [np.zeros_like(self.z[i]) for i in range(k)]
        self.zold = [np.zeros_like(self.z[i]) for i in range(k)]
        self.uold = [np.zeros_like(self.z[i]) for i in range(k)]
        self.xold = np.zeros(n, self.T)
        self.rk = np.full(k, np.inf, rT); self.sk = np.full(k, np.inf, rT)
        self.eps_pri = np.zeros(k, rT); self.eps_dua = np.zeros(k, rT)
        self.sigma_abs = rT(np.sqrt(rT(len(b))) * self.absTol)    # :212
        self.delta = np.full(k, np.inf, rT)
        self.rho_state = self.rho.copy()                   # :215
        self.iteration = 0
        self.cg_iters = []
        if isinstance(self.normalizeReg, MeasurementBasedNormalization):   # :219 (passes b)
            f = normalize_factor(self.normalizeReg, self.A, b)
            self.reg = [_normalize_reg(r, f) for r in self.reg]

    def converged(self):
        for i in range(len(self.reg)):
            if self.rk[i] >= self.sigma_abs + self.relTol * self.eps_pri[i]:
                return False
            if self.sk[i] >= self.sigma_abs + self.relTol * self.eps_dua[i]:
                return False
        return True

    def done(self):
        return self.converged() or self.iteration >= self.iterations

    def _composite(self, v):
        """AHA + sum_i rho_i Phi_i' Phi_i as LinearOperators composes it (ADMM.jl:141-159):
        res = AHA v; then res = rho_i*(N_i v) + res for each term."""
        res = self.AHA.apply(v).astype(self.T, copy=False)
        for i, t in enumerate(self.regTrafo):
            if isinstance(t, _Eye):
                res = self.rho_state[i] * v + res
            else:
                res = grad_t_axpy(self.rho_state[i], t.mul(v), t.shape, t.dims, res)
        return res

    def iterate(self):
        rT = self.rT
        if self.done():
            return False
        k = len(self.reg)
        # 1. x update (:236-244)
        self.beta = self.beta_y.copy()
        for i in range(k):
            t = self.regTrafo[i]
            if isinstance(t, _Eye):
                self.beta = self.rho_state[i] * self.z[i] + self.beta
                self.beta = (-self.rho_state[i]) * self.u[i] + self.beta
            else:
                self.beta = grad_t_axpy(self.rho_state[i], self.z[i], t.shape, t.dims, self.beta)
                self.beta = grad_t_axpy(-self.rho_state[i], self.u[i], t.shape, t.dims, self.beta)
        self.xold[...] = self.x
        self.cg_iters.append(cg(self.x, self._composite, self.beta, self.iterationsCG, self.tolInner))
        for p in self.proj:
            prox_(p, self.x)
        # 2./3. z and u updates + residuals (:251-309)
        for i in range(k):
            t = self.regTrafo[i]
            self.zold[i], self.z[i] = self.z[i], self.zold[i]
            self.z[i][...] = t.mul(self.x)
            self.z[i] += self.u[i]
            if self.rho_state[i] != 0:
                prox_(self.reg[i], self.z[i], lam_of(self.reg[i]) / (2 * self.rho_state[i]))
            self.uold[i][...] = self.u[i]
            self.u[i][...] = t.mul(self.x) + self.u[i]
            self.u[i] -= self.z[i]
            self.xold[...] = self.x - self.xold
            self.zold[i][...] = self.z[i] - self.zold[i]
            self.uold[i][...] = self.u[i] - self.uold[i]
            delta_old = self.delta[i]
            self.delta[i] = rT(_norm2(self.xold)) + rT(_norm2(self.zold[i])) + rT(_norm2(self.uold[i]))
            self.xold[...] = t.tmul(self.zold[i])
            self.sk[i] = self.rho_state[i] * rT(_norm2(self.xold))
            self.zold[i][...] = t.mul(self.x)
            self.eps_pri[i] = max(rT(_norm2(self.zold[i])), rT(_norm2(self.z[i])))
            self.zold[i] -= self.z[i]
            self.rk[i] = rT(_norm2(self.zold[i]))
            self.xold[...] = t.tmul(self.u[i])
            self.eps_dua[i] = self.rho_state[i] * rT(_norm2(self.xold))
            with np.errstate(invalid="ignore", divide="ignore"):
                if (self.vary_rho == "balance" and self.rk[i] / self.eps_pri[i] > 10 * self.sk[i] / self.eps_dua[i]) or \
                   (self.vary_rho == "PnP" and self.delta[i] / delta_old > 0.9):
                    self.rho_state[i] *= 2
                    self.u[i] /= 2
                elif self.vary_rho == "balance" and self.sk[i] / self.eps_dua[i] > 10 * self.rk[i] / self.eps_pri[i]:
                    self.rho_state[i] /= 2
                    self.u[i] *= 2
        self.iteration += 1
        return True

    def convergence(self):
        return {"primal": self.rk.copy(), "dual": self.sk.copy()}


class SplitBregman(_Base):
    """src/SplitBregman.jl:80-290 (precon = Identity only).  Constrained split Bregman: `iterations` outer
    (Bregman) iterations of at most `iterationsInner` ADMM-like inner iterations each; `state.iteration` counts
    inner iterations from 1 and is reset at every outer update (:258-268)."""
    def __init__(self, A, *, AHA=None, reg=None, regTrafo=None, normalizeReg=None, rho=1e-1, iterations=10,
                 iterationsInner=10, iterationsCG=10, absTol=None, relTol=None, tolInner=1e-5, verbose=False,
                 normal="lazy"):
        self.A = A
        self.AHA = _make_normal(A, AHA, normal)
        self.T = np.dtype(self.AHA.dtype)
        rT = self.rT = real_type(self.T)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        if reg is None:
            reg = L1Regularization(rT(0))
        regs, self.proj = _split_regs(reg, "SplitBregman", exactly_one=False)
        n = self.AHA.n
        if regTrafo is None:
            trafo = [_Eye(n) for _ in regs]
        else:
            trafo = list(regTrafo) if isinstance(regTrafo, (list, tuple)) else [regTrafo]
            trafo = [_Eye(n) if t is None else t for t in trafo]
        assert len(regs) == len(trafo), "reg and regTrafo must have the same length"   # :108
        self.regTrafo = trafo
        if np.isscalar(rho):
            self.rho = np.array([rT(rho) for _ in regs], dtype=rT)                    # :110-114
        else:
            self.rho = np.asarray(rho, dtype=rT).copy()
        f = normalize_factor(self.normalizeReg, A, None)                               # :140
        self.reg = [_normalize_reg(r, f) for r in regs]
        self.iterations = int(iterations); self.iterationsInner = int(iterationsInner); self.iterationsCG = int(iterationsCG)
        eps = np.finfo(rT).eps
        self.absTol = rT(eps if absTol is None else absTol)
        self.relTol = rT(eps if relTol is None else relTol)
        self.tolInner = rT(tolInner)
        self.verbose = verbose
        self.x = np.zeros(n, self.T)
        self.iteration = 1; self.iter_cnt = 1
        self.total_iterations = 0
        self.cg_iters = []

    def init(self, b, x0=0):
        rT = self.rT; n = self.AHA.n; k = len(self.reg)
        self.x = np.zeros(n, self.T); self.x[...] = x0                      # :171
        self.beta_y = self._adjoint_b(b)                                   # :174-178
        self.y = self.beta_y.copy()                                        # :179
        self.z = [np.asarray(self.regTrafo[i].mul(self.x), dtype=self.T) for i in range(k)]   # :183
        self.u = [np.zeros_like(self.z[i]) for i in range(k)]
        self.zold = [np.zeros_like(self.z[i]) for i in range(k)]
        self.rk = np.full(k, np.inf, rT); self.sk = np.full(k, np.inf, rT)
        self.eps_pri = np.zeros(k, rT); self.eps_dua = np.zeros(k, rT)
        self.sigma_abs = rT(np.sqrt(rT(len(b))) * self.absTol)              # :192
        if isinstance(self.normalizeReg, MeasurementBasedNormalization):   # :195
            f = normalize_factor(self.normalizeReg, self.A, b)
            self.reg = [_normalize_reg(r, f) for r in self.reg]
        self.iter_cnt = 1                                                  # :198-199
        self.iteration = 1
        self.total_iterations = 0
        self.cg_iters = []

    def converged(self):                                                   # :281-287
        for i in range(len(self.reg)):
            if self.rk[i] >= self.sigma_abs + self.relTol * self.eps_pri[i]:
                return False
            if self.sk[i] >= self.sigma_abs + self.relTol * self.eps_dua[i]:
                return False
        return True

    def done(self):                                                        # :289
        return self.converged() or (self.iteration == 1 and self.iter_cnt > self.iterations)

    def _composite(self, v):
        res = self.AHA.apply(v).astype(self.T, copy=False)
        for i, t in enumerate(self.regTrafo):
            if isinstance(t, _Eye):
                res = self.rho[i] * v + res
            else:
                res = grad_t_axpy(self.rho[i], t.mul(v), t.shape, t.dims, res)
        return res

    def iterate(self):
        rT = self.rT
        if self.done():
            return False
        k = len(self.reg)
        # x update (:209-218)
        self.beta = self.beta_y.copy()
        for i in range(k):
            t = self.regTrafo[i]
            if isinstance(t, _Eye):
                self.beta = self.rho[i] * self.z[i] + self.beta
                self.beta = (-self.rho[i]) * self.u[i] + self.beta
            else:
                self.beta = grad_t_axpy(self.rho[i], self.z[i], t.shape, t.dims, self.beta)
                self.beta = grad_t_axpy(-self.rho[i], self.u[i], t.shape, t.dims, self.beta)
        self.cg_iters.append(cg(self.x, self._composite, self.beta, self.iterationsCG, self.tolInner))
        for p in self.proj:                                                # :220-222
            prox_(p, self.x)
        for i in range(k):                                                 # :225-254
            t = self.regTrafo[i]
            self.zold[i], self.z[i] = self.z[i], self.zold[i]
            self.z[i][...] = t.mul(self.x)
            self.z[i] += self.u[i]
            if self.rho[i] != 0:
                prox_(self.reg[i], self.z[i], lam_of(self.reg[i]) / self.rho[i])
            self.u[i][...] = t.mul(self.x) + self.u[i]                     # mul!(u, Φ, x, 1, 1)
            self.u[i] -= self.z[i]
            phix = np.asarray(t.mul(self.x), dtype=self.T)
            self.rk[i] = rT(_norm2(phix - self.z[i]))
            self.sk[i] = rT(_norm2(self.rho[i] * np.asarray(t.tmul(self.z[i] - self.zold[i]), dtype=self.T)))
            self.eps_pri[i] = max(rT(_norm2(phix)), rT(_norm2(self.z[i])))
            self.eps_dua[i] = rT(_norm2(self.rho[i] * np.asarray(t.tmul(self.u[i]), dtype=self.T)))
        if self.converged() or self.iteration >= self.iterationsInner:     # :257-268
            self.beta_y += self.y
            self.beta_y[...] = -self.AHA.apply(self.x).astype(self.T, copy=False) + self.beta_y   # mul!(β_y, AHA, x, -1, 1)
            for i in range(k):
                self.z[i][...] = self.regTrafo[i].mul(self.x)
                self.u[i][...] = 0
            self.iter_cnt += 1
            self.iteration = 0
        self.iteration += 1
        self.total_iterations += 1
        return True

    def convergence(self):
        return {"primal": self.rk.copy(), "dual": self.sk.copy()}


class Kaczmarz(_Base):
    """src/Kaczmarz.jl:73-317 (constructor :73-159, init! :178-216, iterate :264-283, row step :305-310,
    initkaczmarz :365-392, rowProbabilities :326-334), src/Utils.jl:16-23 (rownorm²), :59-105 (dot_with_matrix_row),
    Kaczmarz.jl:432-436 (kaczmarz_update!).  The greedy-randomised variant (:231-262, :394-428) is not restated
    (the reference itself excludes it on GPU arrays, test/testKaczmarz.jl:114).

    Row order: Julia's RNG stream cannot be reproduced, so ``shuffleRows`` / ``randomized`` draw from
    ``numpy.random.default_rng(seed)`` — a permutation at init!, and per iteration a weighted sample without
    replacement of ``subMatrixSize`` rows (StatsBase.sample! :267-269)."""
    def __init__(self, A, *, reg=None, normalizeReg=None, randomized=False, subMatrixFraction=0.15, shuffleRows=False,
                 seed=1234, iterations=10, greedy_randomized=False):
        if greedy_randomized:
            raise NotImplementedError("greedy randomised Kaczmarz is not restated")
        A = np.asarray(A)
        self.T = np.dtype(A.dtype)
        rT = self.rT = real_type(self.T)
        self.normalizeReg = NoNormalization() if normalizeReg is None else normalizeReg
        regs = [L2Regularization(rT(0))] if reg is None else (list(reg) if isinstance(reg, (list, tuple)) else [reg])
        f = normalize_factor(self.normalizeReg, A, None)                    # :82 (type-based: also SystemMatrixBased)
        regs = [_normalize_reg(r, f) for r in regs]
        l2 = [r for r in regs if isinstance(_sink(r), L2Regularization)]
        if len(l2) > 1:
            raise ValueError("Cannot unambigiously retrieve reg term of type L2Regularization")
        self.L2 = l2[0] if l2 else L2Regularization(rT(0))                  # :83-89
        rest = [r for r in regs if not isinstance(_sink(r), L2Regularization)]
        lam = lam_of(self.L2)
        if np.ndim(lam) > 0 and not isinstance(self.normalizeReg, (NoNormalization, SystemMatrixBasedNormalization)):
            raise ValueError("Tikhonov matrix for Kaczmarz is only valid with no or system matrix based normalization")
        other = [r for r in rest if _is_projection(_sink(r))]               # :95-97
        rest = [r for r in rest if not _is_projection(_sink(r))]
        if len(rest) == 1:
            other.append(rest[0])
        elif len(rest) > 1:
            raise ValueError(f"Kaczmarz does not allow for more than one additional regularization term, found {len(rest)}")
        self.reg = other
        self.A, self.denom, self.rowindex = self._initkaczmarz(A, lam)      # :118
        M, N = self.A.shape
        self.randomized = bool(randomized); self.shuffleRows = bool(shuffleRows); self.seed = int(seed)
        self.subMatrixSize = int(np.round(subMatrixFraction * M))           # :121 (round half to even, as Julia)
        self.rowIndexCycle = np.arange(len(self.rowindex))
        self.probabilities = self._row_probabilities() if self.randomized else None
        self.usedIndices = np.zeros(self.subMatrixSize, np.int64) if self.randomized else self.rowIndexCycle
        self.iterations = int(iterations)
        self.u = np.zeros(M, self.T); self.x = np.zeros(N, self.T); self.vl = np.zeros(M, self.T)
        self.eps_w = self.T.type(0); self.iteration = 0

    # -- rownorm² (Utils.jl:16-23): sequential sum of abs2 in the real type
    def _rownorm2(self, A):
        rT = self.rT
        a2 = (A.real.astype(rT) ** 2 + A.imag.astype(rT) ** 2) if np.iscomplexobj(A) else A.astype(rT) ** 2
        return np.sum(a2, axis=1, dtype=rT)

    def _initkaczmarz(self, A, lam):
        """:365-392.  denom = T(1.0 / (s² + λ)): the quotient is formed in Float64 and stored in the real type."""
        rT = self.rT
        if np.ndim(lam) > 0:                                                # Tikhonov matrix :377-392
            lam = np.asarray(lam).astype(rT)
            A = (A * (rT(1) / np.sqrt(lam))[None, :]).astype(self.T)
            lam = rT(1)
        s2 = self._rownorm2(A)
        keep = np.nonzero(s2 > 0)[0]
        tot = s2[keep] + lam                                                # Float32 + Float32, or promoted by a Float64 λ
        denom = (1.0 / tot.astype(np.float64)).astype(rT)
        self._s2 = s2
        return A, denom, keep

    def _row_probabilities(self):
        """:326-334 (Float64 accumulation vector, converted to T at :125)"""
        s2 = self._s2
        tot = self.rT(np.sum(s2, dtype=self.rT))
        return (s2[self.rowindex].astype(np.float64) / np.float64(tot)).astype(self.rT)

    def init(self, b, x0=0):
        rT = self.rT
        lam_prev = lam_of(self.L2)
        if not isinstance(self.normalizeReg, SystemMatrixBasedNormalization):   # NormalizedRegularization.jl:84
            f = normalize_factor(self.normalizeReg, self.A, b)
            self.L2 = _normalize_reg(self.L2, f)
            self.reg = [_normalize_reg(r, f) for r in self.reg]
        lam = lam_of(self.L2)
        if np.ndim(lam) == 0 and lam != lam_prev:                           # :186-193
            _, self.denom, self.rowindex = self._initkaczmarz(self.A, lam)
            self.rowIndexCycle = np.arange(len(self.rowindex))
            if self.randomized:
                self.probabilities = self._row_probabilities()
        self._rng = np.random.default_rng(self.seed)                        # :195-197
        if self.randomized:
            self.usedIndices = np.zeros(self.subMatrixSize, np.int64)
        elif self.shuffleRows:
            self.rowIndexCycle = self._rng.permutation(self.rowIndexCycle)  # :201
            self.usedIndices = self.rowIndexCycle
        else:
            self.usedIndices = self.rowIndexCycle
        self.x = np.zeros(self.A.shape[1], self.T); self.x[...] = x0        # :206
        self.vl = np.zeros(self.A.shape[0], self.T)
        self.u = np.array(b, dtype=self.T, copy=True)
        self.eps_w = self.T.type(1) if np.ndim(lam) > 0 else self.T.type(np.sqrt(lam))   # :210-214
        self.iteration = 0

    def done(self):
        return self.iteration >= self.iterations                           # :315

    def iterate(self):
        if self.done():
            return False
        T = self.T.type
        if self.randomized:                                                 # :267-269
            p = self.probabilities.astype(np.float64)
            self.usedIndices = self._rng.choice(len(self.rowIndexCycle), size=self.subMatrixSize, replace=False,
                                                p=p / p.sum())
        A = self.A
        for i in self.usedIndices:                                          # :270-273
            row = self.rowindex[i]
            a = A[row]
            tau = T(np.dot(a, self.x))                                      # dotu: no conjugation (Utils.jl:73-105)
            alpha = T(self.denom[i] * (self.u[row] - tau - self.eps_w * self.vl[row]))   # :307
            self.x += alpha * np.conj(a)                                    # :432-436
            self.vl[row] += alpha * self.eps_w                              # :309
        for r in self.reg:                                                  # :275-277
            prox_(r, self.x)
        self.iteration += 1
        return True

    def solution(self):
        """solversolution :253-256: the Tikhonov-matrix form returns x ./ sqrt.(λ)"""
        lam = lam_of(self.L2)
        if np.ndim(lam) > 0:
            return (self.x * (self.rT(1) / np.sqrt(np.asarray(lam).astype(self.rT)))).astype(self.T)
        return self.x

    def solve(self, b, callbacks=None, **kw):
        super().solve(b, callbacks=callbacks, **kw)
        return self.solution()

    def convergence(self):
        return {"residual": _norm2(self.A @ self.solution() - self.u)}


def createLinearSolver(solver, A, **kw):
    """RegularizedLeastSquares.jl:288-294 (unknown kwargs are dropped with a warning upstream)."""
    import inspect
    names = set()
    for klass in solver.__mro__:
        if klass is object:
            continue
        names |= set(inspect.signature(klass.__init__).parameters)
    kept = {k: v for k, v in kw.items() if k in names}
    return solver(A, **kept)


def solve_(solver, b, **kw):
    return solver.solve(b, **kw)
