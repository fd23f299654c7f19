"""mod_particleMover: depositJ / moveParticles with the reference's argument lists
(OpenMP/NGP/src/mod_particleMover.f90:17,67-74,162-167 ; OpenMP/CIC/src/mod_particleMoverPicard.f90:67-70,77-84,198-203)."""
import ctypes as C
from . import _lib


def allocateParticleMoverData():
    """NGP: no-op (:17-19).  CIC allocates particleTimeStep (:67-70); here the per-cell cap table lives on the device."""
    return None


def depositJ(solver, particleList, world, del_t):
    """Pushes every particle through del_t in E^{n+1/2} without storing it and leaves J in solver.J."""
    solver._bind(particleList)
    _lib.check(_lib.load().ipic_deposit_J(solver._h, del_t))


def moveParticles(solver, particleList, world, del_t):
    """Same push, storing (l_f, v_f), absorbing at walls with loss sums, stable compaction, sub-step statistics."""
    solver._bind(particleList)
    _lib.check(_lib.load().ipic_move_particles(solver._h, del_t))
    for p in particleList:
        p._refresh_stats()


def depositSubsteps(solver, particle):
    n = C.c_int64()
    _lib.check(_lib.load().ipic_get_deposit_substeps(solver._h, particle._index, C.byref(n)))
    return n.value
