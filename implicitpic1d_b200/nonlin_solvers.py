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


class NitsolInfo:
    """info(1:6) of nitsol (mod_nitsol.f90:3430-3435) + iterm"""

    def __init__(self, v):
        self.nfe, self.njve, self.nrpre, self.nli, self.nni, self.nbt, self.iterm = [int(x) for x in v]

    def __repr__(self):
        return "NitsolInfo(nfe=%d, njve=%d, nrpre=%d, nli=%d, nni=%d, nbt=%d, iterm=%d)" % (
            self.nfe, self.njve, self.nrpre, self.nli, self.nni, self.nbt, self.iterm)


def solveJFNK(solver, particleList, world, del_t):
    """solveJFNK (mod_nonLinSolvers.f90:255-282): NITSOL on F(phi_f) = phi_f - G(phi_f), then moveParticles on success.
    Needs initializeSolver(..., solverType=1, ...)."""
    solver._bind(particleList)
    _lib.check(_lib.load().ipic_solve_jfnk(solver._h, del_t))
    for p in particleList:
        p._refresh_stats()
    v = np.zeros(7, dtype=np.int32)
    _lib.check(_lib.load().ipic_get_nitsol_info(solver._h, v.ctypes.data_as(_lib._ip)))
    return NitsolInfo(v)


def nitsol(f, x0, abstol, reltol, stptol=1.e-12, maxIter=100, kdmax=10):
    """The library's NITSOL driver on a host callback f(x) -> F(x) (numpy arrays); returns (x, NitsolInfo)."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    CB = C.CFUNCTYPE(C.c_int, C.c_int32, _lib._dp, _lib._dp, C.c_void_p)

    def cb(nn, px, pf, user):
        try:
            xv = np.ctypeslib.as_array(px, shape=(nn,))
            fv = np.ctypeslib.as_array(pf, shape=(nn,))
            fv[:] = f(xv.copy())
            return 0
        except Exception:
            return 1
    info = np.zeros(6, dtype=np.int32)
    iterm = C.c_int32()
    _lib.check(_lib.load().ipic_nitsol_host(n, x.ctypes.data_as(_lib._dp), CB(cb), None, abstol, reltol, stptol, maxIter, kdmax,
                                            info.ctypes.data_as(_lib._ip), C.byref(iterm)))
    return x, NitsolInfo(list(info) + [iterm.value])
