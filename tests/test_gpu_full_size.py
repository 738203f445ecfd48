"""Size-independent properties at BASELINE.json's full single-GPU shape (configs[1]:
Float32 A 16384 x 65536), where the NumPy oracle is too slow to be the checker:
linearity and adjointness of the operator kernels, one-pass == two-pass, and a
row-subsampled oracle comparison (rows of the same Philox matrix)."""
import numpy as np
import pytest

from util import rel

pytestmark = pytest.mark.gpu


def test_c2_operator_properties(rls, ctx):
    m, n = 16384, 65536
    A = rls.B200Matrix.philox(np.float32, m, n, seed=12345, scale=1.0 / np.sqrt(m), ctx=ctx)
    x = rls.B200Vector(ctx, np.float32, n).fill_philox(1, stream=1, dist=1)
    y = rls.B200Vector(ctx, np.float32, m).fill_philox(2, stream=2, dist=1)
    Ax = A.mul(x)
    Aty = A.adjoint_mul(y)
    # adjointness  <A x, y> == <x, A' y>
    lhs, rhs = Ax.dot(y), x.dot(Aty)
    assert abs(lhs - rhs) < 1e-5 * max(1.0, abs(lhs))
    two = rls.B200NormalOp(A, form="twopass").apply(x).to_numpy()
    one = rls.B200NormalOp(A, form="onepass").apply(x).to_numpy()
    assert rel(one, two) < 2e-6
    # the column-major storage (what rls_mat_wrap_device adopts) holds the same matrix and gives the same result
    assert A.layout == "row"
    Ac = rls.B200Matrix.philox(np.float32, m, n, seed=12345, scale=1.0 / np.sqrt(m), ctx=ctx, layout="col")
    assert rel(rls.B200NormalOp(Ac, form="twopass").apply(x).to_numpy(), one) < 2e-6
    assert rel(rls.B200NormalOp(Ac, form="onepass").apply(x).to_numpy(), one) < 2e-6
    del Ac
    # x' (A'A x) == ||A x||^2
    assert abs(float(np.dot(x.to_numpy().astype(np.float64), two.astype(np.float64))) - Ax.norm() ** 2) < 1e-5 * Ax.norm() ** 2
    # row subsample against NumPy on the same Philox entries
    from oracle.philox import philox_matrix, IH4
    rows = philox_matrix(np.float32, 64, n, 12345, IH4, 1.0 / np.sqrt(m), row_offset=4096, m_global=m)
    ref = rows.astype(np.float64) @ x.to_numpy().astype(np.float64)
    assert rel(Ax.to_numpy()[4096:4160], ref) < 2e-6


def test_c2_fista_onepass_equals_twopass(rls, ctx):
    m, n = 16384, 65536
    A = rls.B200Matrix.philox(np.float32, m, n, seed=12345, scale=1.0 / np.sqrt(m), ctx=ctx)
    b = rls.B200Vector(ctx, np.float32, m).fill_philox(3, stream=3, dist=1).to_numpy()
    out = {}
    for form in ("twopass", "onepass"):
        S = rls.FISTA(A, reg=rls.L1Regularization(np.float32(1e-3)), iterations=20, rho=np.float32(0.1), relTol=0.0,
                      normal=form)
        out[form] = rls.solve_(S, b)
        assert S.iteration == 20
    assert rel(out["onepass"], out["twopass"]) < 1e-5
