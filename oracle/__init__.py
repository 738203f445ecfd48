"""CPU oracle for the RegularizedLeastSquares.jl hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`regularizedleastsquares.jl_b200/`) may import, call, link or execute this
directory; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do, and only as the checker / baseline.

Parity status: the Julia reference cannot be executed in this image (no
`julia`), so the oracle is pinned against the reference's own known answers:
the `solve!` docstring KAT (src/RegularizedLeastSquares.jl:44-61, reproduced to
1e-12) plus the acceptance properties of test/testSolvers.jl,
test/testProxMaps.jl, test/testMultiThreading.jl and
docs/src/literate/examples/getting_started.jl (see tests/test_oracle_*.py).
"""
from .rls_oracle import *  # noqa: F401,F403
