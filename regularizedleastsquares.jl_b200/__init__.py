"""regularizedleastsquares.jl_b200 — B200-native (sm_100a) implementation of the
RegularizedLeastSquares.jl iterative-solver inner loop behind the reference's own API.

The directory name contains a dot, so import it through the root-level alias module:

    import rls_b200 as rls
    S = rls.createLinearSolver(rls.FISTA, A, reg=rls.L1Regularization(np.float32(1e-3)), iterations=200, rho=rho)
    x = rls.solve_(S, b)

Layout: csrc/ (CUDA kernels + C ABI, built into lib/librls_b200.so), _capi.py (ctypes),
arrays.py / regularization.py / prox.py / solvers.py (host mirror of the Julia API),
dist.py (row sharding, one process per GPU), julia/RLSB200.jl (the ccall shim).
Importing this package dlopens the library and fails loudly when it is missing: there is
no CPU fallback anywhere in the product path.
"""
from . import _capi

_capi.load()   # fail loudly if the CUDA library has not been built

from ._capi import RlsError  # noqa: E402
from .arrays import B200Context, B200Group, B200GroupMatrix, B200Matrix, B200NormalOp, B200Vector  # noqa: E402
from .regularization import (AbstractParameterizedRegularization, AbstractProjectionRegularization,  # noqa: E402
                             AbstractRegularization, GradientOp, L1Regularization, L2Regularization,
                             L21Regularization, LLRRegularization, MeasurementBasedNormalization, NoNormalization, NuclearRegularization,
                             NormalizedRegularization, PositiveRegularization, RealRegularization,
                             SystemMatrixBasedNormalization, TVRegularization, findsink, findsinks, lam, sink)
from .operators import FFTOp, LinearOperator, SamplingOp  # noqa: E402
from .prox import prox_  # noqa: E402
from .solvers import (ADMM, CGNR, FISTA, POGM, Kaczmarz, OptISTA, SplitBregman, AbstractLinearSolver, createLinearSolver, init_, iterate,  # noqa: E402
                      linearSolverList, solve_, solverconvergence, solversolution, solverstate)
from .callbacks import CompareSolutionCallback, StoreConvergenceCallback, StoreSolutionCallback, nrmsd  # noqa: E402
from . import dist  # noqa: E402

ABI_VERSION = _capi.load().rls_abi_version()
