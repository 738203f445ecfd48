"""NumPy mirror of csrc/rls_svt.cu (test infrastructure): the same decomposition — Float64 Gram matrix of the short side,
two-sided Jacobi with the kernel's round-robin ordering, rotation formulas and tolerance, W = V diag(max(s-λ,0)/s) V', out = X·W — and
the same index arithmetic (svt_offset) for the LLR patch views.  It lets the CPU suite check the ALGORITHM of the device
path against the oracle's LAPACK-SVD restatement of the reference; the -m gpu tests then check the kernels themselves."""
import numpy as np

MAXQ = 64


def _down(a, dtype):
    return a.astype(dtype) if np.iscomplexobj(np.empty(0, dtype)) else np.real(a).astype(dtype)


def pairs_of_round(q, rr):
    """the round-robin ("circle") ordering of svt_eig_kernel: round rr of a sweep rotates these disjoint pairs at once"""
    qe = q + (q & 1)
    n1 = qe - 1
    out = []
    for k in range(qe // 2):
        a, b = (rr, n1) if k == 0 else ((rr + k) % n1, (rr - k + n1) % n1)
        if a < q and b < q:                    # the padding index of an odd q sits out
            out.append((min(a, b), max(a, b)))
    return out


def jacobi_w(G, thr):
    """svt_eig_kernel: eigen-decomposition of the Hermitian G (complex128) by Jacobi rotations — a sweep is q-1 rounds of
    q/2 disjoint pairs, each round G <- J' G J, V <- V J with J the direct sum of the pairs' rotations — and W = V D V'"""
    q = G.shape[0]
    G = G.astype(np.complex128).copy()
    V = np.eye(q, dtype=np.complex128)
    tr = float(np.real(np.trace(G)))
    tol2 = (1e-15 * tr) ** 2
    qe = q + (q & 1)
    for _ in range(40):
        rotated = False
        for rr in range(qe - 1):
            rots = []
            for (p, r) in pairs_of_round(q, rr):
                b = G[p, r]
                ab2 = b.real * b.real + b.imag * b.imag
                if ab2 <= tol2:
                    continue
                d = G[r, r].real - G[p, p].real
                h = np.sqrt(d * d + 4.0 * ab2)
                w = 1.0 / (abs(d) + h)
                c = 1.0 / np.sqrt(4.0 * ab2 * w * w + 1.0)
                k = (2.0 if d >= 0 else -2.0) * w * c
                rots.append((p, r, c, k * b))
            if not rots:
                continue
            rotated = True
            for p, r, c, sph in rots:              # columns of G and V
                sphc = np.conj(sph)
                gp, gr = G[:, p].copy(), G[:, r].copy()
                G[:, p] = c * gp - sphc * gr
                G[:, r] = sph * gp + c * gr
                vp, vr = V[:, p].copy(), V[:, r].copy()
                V[:, p] = c * vp - sphc * vr
                V[:, r] = sph * vp + c * vr
            for p, r, c, sph in rots:              # rows of G
                sphc = np.conj(sph)
                rp, rw = G[p, :].copy(), G[r, :].copy()
                G[p, :] = c * rp - sph * rw
                G[r, :] = sphc * rp + c * rw
            for p, r, c, sph in rots:
                G[p, r] = 0.0
                G[r, p] = 0.0
                G[p, p] = G[p, p].real
                G[r, r] = G[r, r].real
        if not rotated:
            break
    ev = np.real(np.diag(G))
    sv = np.where(ev > 0, np.sqrt(np.maximum(ev, 0)), 0.0)
    D = np.where(sv > 0, np.maximum(sv - thr, 0.0) / np.where(sv > 0, sv, 1.0), 0.0)
    return (V * D) @ V.conj().T


def svt_tall(Y, thr, llr=False, rowmax_ub=False):
    """Y: long x short view (complex128 / float64 holding Float32 values).  Returns Y·W (or zeros by the LLR shortcut)."""
    G = Y.conj().T @ Y
    if not np.any(G):
        return np.zeros_like(Y)
    if llr:
        g2 = np.max(np.sum(np.abs(Y) ** 2, axis=1)) if rowmax_ub else np.max(np.abs(G))
        if np.float32(thr) >= np.sqrt(np.float32(g2)):
            return np.zeros_like(Y)
    return Y @ jacobi_w(G, float(np.float32(thr)))


def prox_nuclear(x, lam, rows, cols):
    """rls_prox_nuclear_launch: mode 0 (short side = columns) or mode 1 (the conjugate-transposed view)"""
    assert min(rows, cols) <= MAXQ
    up = np.complex128 if np.iscomplexobj(x) else np.float64
    X = x.reshape((rows, cols), order="F").astype(up)
    if cols <= rows:
        out = svt_tall(X, lam)
    else:
        out = svt_tall(X.conj().T, lam).conj().T
    return _down(out, x.dtype).reshape(-1, order="F")


def _llr_pass(x, lam, shape, block, shift):
    """llr_pass + svt_offset (modes 2 / 3): returns the thresholded copy (every element belongs to exactly one patch)"""
    nd = len(shape)
    npix = int(np.prod(shape))
    K = x.size // npix
    ppix = int(np.prod(block))
    stride = [int(np.prod(shape[:d])) for d in range(nd)]
    nblk = [(shape[d] + block[d] - 1) // block[d] for d in range(nd)]
    sh = [((shift[d] if shift is not None else 0) % shape[d] + shape[d]) % shape[d] for d in range(nd)]
    transposed = not (K <= ppix)
    assert (ppix if transposed else K) <= MAXQ
    up = np.complex128 if np.iscomplexobj(x) else np.float64
    out = x.copy()
    for prob in range(int(np.prod(nblk))):
        offs = np.full(ppix, -1, dtype=np.int64)
        for l in range(ppix):
            pr, lr, off, ok = prob, l, 0, True
            for d in range(nd):
                o = (pr % nblk[d]) * block[d]
                pr //= nblk[d]
                t = lr % block[d]
                lr //= block[d]
                c = o + t
                if c >= shape[d]:
                    ok = False
                    break
                c -= sh[d]
                if c < 0:
                    c += shape[d]
                off += c * stride[d]
            if ok:
                offs[l] = off
        valid = offs >= 0
        X = np.zeros((ppix, K), dtype=up)                         # pixels x frames
        for i in range(K):
            X[valid, i] = x[offs[valid] + i * npix]
        if transposed:
            R = svt_tall(X.conj().T, lam, llr=True, rowmax_ub=True).conj().T
        else:
            R = svt_tall(X, lam, llr=True)
        for i in range(K):
            out[offs[valid] + i * npix] = _down(R[valid, i], x.dtype)
    return out


def prox_llr(x, lam, shape, block, shift=None, fully_overlapping=False):
    if not fully_overlapping:
        return _llr_pass(x, lam, shape, block, shift)
    nd = len(shape)
    assert all(shape[d] % block[d] == 0 for d in range(nd))
    acc = np.zeros_like(x)
    nshift = int(np.prod(block))
    for sidx in range(nshift):
        r, sh = sidx, []
        for d in range(nd):
            sh.append(1 + r % block[d] + (shift[d] if shift is not None else 0))
            r //= block[d]
        acc = acc + _llr_pass(x, lam, shape, block, sh)           # Float32 running sum, as svt_apply_kernel's acc
    rt = np.float32
    return (acc / rt(nshift)).astype(x.dtype)
