"""Host-side `type(potentialSolver)` (OpenMP/NGP/src/mod_potentialSolver.f90:17-41, OpenMP/CIC/src/mod_potentialSolver.f90:17-41).

Owns the device context.  Field vectors (phi, phi_f, J, rho, EField, tridiagonal) live on the device and are
fetched on attribute access."""
import ctypes as C
import numpy as np
from . import _lib
from .domain import Domain


class potentialSolver:
    def __init__(self, world: Domain, leftVoltage, rightVoltage, RF_frequency, timeCurrent, device=0):
        L = _lib.load()
        self.world = world
        self.device = int(device)
        self._h = C.c_void_p()
        g = np.ascontiguousarray(world.grid, dtype=np.float64)
        dx = np.ascontiguousarray(world.dx_dl, dtype=np.float64)
        bc = np.ascontiguousarray(world.boundaryConditions, dtype=np.int32)
        _lib.check(L.ipic_create(C.byref(self._h), world.scheme, world.NumberXNodes, g.ctypes.data_as(_lib._dp), dx.ctypes.data_as(_lib._dp),
                                 bc.ctypes.data_as(_lib._ip), int(world.gridSmoothBool), device))
        _lib.check(L.ipic_init_solver(self._h, leftVoltage, rightVoltage, RF_frequency, timeCurrent))
        self._particles = []
        # RF wall (bc 4): solver%RF_rad_frequency (NGP mod_potentialSolver.f90:62-84)
        self.RF_rad_frequency = 2.0 * np.pi * float(RF_frequency) if 4 in (int(bc[0]), int(bc[-1])) else 0.0

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().ipic_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- particle binding
    def _bind(self, particleList):
        L = _lib.load()
        for p in particleList:
            if p._solver is self:
                continue
            if p._solver is not None:
                raise RuntimeError("particle list already bound to another solver")
            idx = C.c_int32()
            _lib.check(L.ipic_add_species(self._h, p.mass, p.q, p.w_p, C.byref(idx)))
            p._solver, p._index = self, idx.value
            self._particles.append(p)
            if p.phaseSpace is not None:
                p.upload()

    # ---- field access
    def _get(self, which, n):
        out = np.zeros(n)
        _lib.check(_lib.load().ipic_get_field(self._h, which, out.ctypes.data_as(_lib._dp), n))
        return out

    def _set(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        _lib.check(_lib.load().ipic_set_field(self._h, which, a.ctypes.data_as(_lib._dp), a.size))

    phi = property(lambda s: s._get(_lib.F_PHI, s.world.NumberXNodes), lambda s, v: s._set(_lib.F_PHI, v))
    phi_f = property(lambda s: s._get(_lib.F_PHI_F, s.world.NumberXNodes), lambda s, v: s._set(_lib.F_PHI_F, v))
    J = property(lambda s: s._get(_lib.F_J, s.world.NumberXHalfNodes), lambda s, v: s._set(_lib.F_J, v))
    rho = property(lambda s: s._get(_lib.F_RHO, s.world.NumberXNodes), lambda s, v: s._set(_lib.F_RHO, v))
    EField = property(lambda s: s._get(_lib.F_EFIELD, s.world.NumberXHalfNodes), lambda s, v: s._set(_lib.F_EFIELD, v))
    EField_f = property(lambda s: s._get(_lib.F_EFIELD_F, s.world.NumberXHalfNodes), lambda s, v: s._set(_lib.F_EFIELD_F, v))
    EField_half = property(lambda s: s._get(_lib.F_EFIELD_HALF, s.world.NumberXHalfNodes))
    a_tri = property(lambda s: s._get(_lib.F_A_TRI, s.world.NumberXNodes - 1))
    b_tri = property(lambda s: s._get(_lib.F_B_TRI, s.world.NumberXNodes))
    c_tri = property(lambda s: s._get(_lib.F_C_TRI, s.world.NumberXNodes - 1))

    def set_tridiag_mode(self, mode):
        _lib.check(_lib.load().ipic_set_tridiag_mode(self._h, mode))

    def set_option(self, name, value):
        _lib.check(_lib.load().ipic_set_option(self._h, name.encode(), float(value)))

    # ---- type-bound procedures
    def depositRho(self, particleList, world=None):
        self._bind(particleList)
        _lib.check(_lib.load().ipic_deposit_rho(self._h))

    def solve_tridiag_Poisson(self, world=None, timeCurrent=0.0):
        _lib.check(_lib.load().ipic_solve_tridiag_poisson(self._h, timeCurrent))

    def solveInitialPotential(self, particleList, world=None, timeCurrent=0.0):
        self._bind(particleList)
        _lib.check(_lib.load().ipic_solve_initial_potential(self._h, timeCurrent))

    def makeConstSourceTerm(self, world=None):
        _lib.check(_lib.load().ipic_make_const_source_term(self._h))

    def makeHalfTimeEField(self, workArray=None, world=None):
        _lib.check(_lib.load().ipic_make_half_time_efield(self._h))

    def solve_tridiag_Ampere(self, world=None, del_t=0.0):
        _lib.check(_lib.load().ipic_solve_tridiag_ampere(self._h, del_t))

    def setRFVoltage(self, world=None, timeFuture=0.0):
        _lib.check(_lib.load().ipic_set_rf_voltage(self._h, timeFuture))

    def resetVoltage(self):
        _lib.check(_lib.load().ipic_reset_voltage(self._h))

    def _scalar(self, fn, *args):
        out = C.c_double()
        _lib.check(fn(self._h, *args, C.byref(out)))
        return out.value

    def getTotalPE(self, world=None, future=False):
        return self._scalar(_lib.load().ipic_get_total_PE, int(future))

    def getJdotE(self, world=None, del_t=0.0):
        return self._scalar(_lib.load().ipic_get_JdotE, del_t)

    def getEnergyFromBoundary(self, world=None, del_t=0.0):
        return self._scalar(_lib.load().ipic_get_energy_from_boundary, del_t)

    def getError_tridiag_Poisson(self, world=None):
        return self._scalar(_lib.load().ipic_get_error_tridiag_poisson)

    def getChargeContinuityError(self, rho_i, world=None, del_t=0.0):
        r = np.ascontiguousarray(rho_i, dtype=np.float64)
        return self._scalar(_lib.load().ipic_get_charge_continuity_error, r.ctypes.data_as(_lib._dp), del_t)

    # ---- introspection
    def timing(self, name):
        ms, n = C.c_double(), C.c_int64()
        _lib.check(_lib.load().ipic_get_timing(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def reset_timing(self):
        _lib.check(_lib.load().ipic_reset_timing(self._h))

    def synchronize(self):
        _lib.check(_lib.load().ipic_synchronize(self._h))
