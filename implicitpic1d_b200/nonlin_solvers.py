"""mod_nonLinSolvers: initializeSolver / solvePotential / solveDivAmpereAnderson / funcNitsol
(OpenMP/SharedModules/src/mod_nonLinSolvers.f90:41-81,85-160,198-215,218-239).

The module-level counters of the reference (iterNumPicard, iterNumAdaptiveSteps, amountTimeSplits, potTimer,
moverTimer; :29-30) are exposed through SolverState."""
import ctypes as C
import numpy as np
from . import _lib


class SolverState:
    """Snapshot of the protected module variables of mod_nonLinSolvers."""

    def __init__(self, solver):
        st = _lib.SolverStats()
        _lib.check(_lib.load().ipic_get_solver_stats(solver._h, C.byref(st)))
        self.iterNumPicard = st.iterNumPicard
        self.iterNumAdaptiveSteps = st.iterNumAdaptiveSteps
        self.amountTimeSplits = st.amountTimeSplits
        self.sweeps, self.substeps = st.sweeps, st.substeps
        self.moverTimer, self.potTimer = st.moverSeconds, st.potSeconds
        self.lastResidualNorm = st.lastResidualNorm
        out = np.zeros(1024)
        n = C.c_int32()
        _lib.check(_lib.load().ipic_get_norm_history(solver._h, out.ctypes.data_as(_lib._dp), 1024, C.byref(n)))
        self.normHistory = out[:n.value].copy()


def initializeSolver(solver, eps_a, eps_r, solverType, m_Anderson, Beta_k, maxIter):
    """Values of SolverState.inp in deck order (mod_nonLinSolvers.f90:49-54)."""
    _lib.check(_lib.load().ipic_initialize_solver(solver._h, eps_a, eps_r, solverType, m_Anderson, Beta_k, maxIter))


def solveDivAmpereAnderson(solver, particleList, world, del_t):
    solver._bind(particleList)
    _lib.check(_lib.load().ipic_solve_div_ampere_anderson(solver._h, del_t))
    for p in particleList:
        p._refresh_stats()


def solvePotential(solver, particleList, world, del_t, remainDel_t, currDel_t, timeCurrent):
    """One implicit step.  remainDel_t / currDel_t are intent(in out) in the reference; returned as a tuple here."""
    solver._bind(particleList)
    r, c = C.c_double(remainDel_t), C.c_double(currDel_t)
    _lib.check(_lib.load().ipic_solve_potential(solver._h, del_t, C.byref(r), C.byref(c), timeCurrent))
    for p in particleList:
        p._refresh_stats()
    return r.value, c.value


def funcNitsol(solver, particleList, xcur, del_t):
    """NITSOL residual callback: fcur = xcur - G(xcur), one full particle sweep (mod_nonLinSolvers.f90:218-239)."""
    solver._bind(particleList)
    x = np.ascontiguousarray(xcur, dtype=np.float64)
    f = np.zeros_like(x)
    _lib.check(_lib.load().ipic_func_nitsol(solver._h, x.size, x.ctypes.data_as(_lib._dp), f.ctypes.data_as(_lib._dp), del_t))
    return f
