"""Callbacks of `solve!` (src/Callbacks.jl) for the host-side mirror: they only use the public accessors
`solversolution` / `solverconvergence`, so they work unchanged on device-resident solvers (the solution is
downloaded once per invocation)."""
from __future__ import annotations

import numpy as np


def nrmsd(I, Ireco):
    """src/Utils.jl:230-242 — RMS deviation after calibrating a global scale away, normalised by the range of |I|."""
    I = np.asarray(I).ravel()
    Ireco = np.asarray(Ireco).ravel()
    N = I.size
    if np.linalg.norm(Ireco) > 0:
        alpha = (np.vdot(I, Ireco) + np.vdot(Ireco, I)) / (2 * np.vdot(Ireco, Ireco))
    else:
        alpha = 1.0
    rms = 1.0 / np.sqrt(N) * np.linalg.norm(I - Ireco * alpha)
    return float(np.real(rms / (np.max(np.abs(I)) - np.min(np.abs(I)))))


class CompareSolutionCallback:
    """CompareSolutionCallback(ref, cmp = nrmsd)  Callbacks.jl:1-18"""
    def __init__(self, ref, cmp=nrmsd):
        self.ref = np.asarray(ref)
        self.cmp = cmp
        self.results = []

    def __call__(self, solver, _):
        self.results.append(float(self.cmp(self.ref, np.asarray(solver.x))))


class StoreSolutionCallback:
    """StoreSolutionCallback(T)  Callbacks.jl:20-33"""
    def __init__(self, T=None):
        self.solutions = []

    def __call__(self, solver, _):
        self.solutions.append(np.array(solver.x, copy=True))


class StoreConvergenceCallback:
    """StoreConvergenceCallback()  Callbacks.jl:35-52"""
    def __init__(self):
        self.convMeas = {}

    def __call__(self, solver, _):
        for key, val in solver.convergence().items():
            self.convMeas.setdefault(key, []).append(val)
