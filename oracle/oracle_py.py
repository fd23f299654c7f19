"""ctypes wrapper of the CPU oracle (oracle/ipic_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by implicitpic1d_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# physical constants, constants.f90:7-15
c_light = 299792458.0
eps_0 = 8.8541878128e-12
m_p = 1.67262192369e-27
m_amu = 1.66053906660e-27
m_e = 9.1093837015e-31
k_B = 1.380649e-23
e = 1.602176634e-19
pi = 4.0 * np.arctan(1.0)


def build(native=False):
    target = "native" if native else "all"
    subprocess.check_call(["make", "-s", "-C", _HERE, target])
    sub = "_build/native" if native else "_build"
    return os.path.join(_HERE, sub, "libipic_oracle.so")


_libs = {}


def lib(native=False):
    if native in _libs:
        return _libs[native]
    sub = "_build/native" if native else "_build"
    path = os.path.join(_HERE, sub, "libipic_oracle.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "ipic_oracle.cpp")):
        path = build(native)
    L = C.CDLL(path)
    d, i, p, ll = C.c_double, C.c_int, C.c_void_p, C.c_longlong
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "orc_create": (p, [i, i, d, i, d, i, i, i, i]),
        "orc_destroy": (None, [p]),
        "orc_Nx": (i, [p]), "orc_Nh": (i, [p]), "orc_T": (i, [p]), "orc_L_domain": (d, [p]),
        "orc_get_grid": (None, [p, dp, dp, ip]),
        "orc_get_slices": (None, [p, ip, ip, ip, ip]),
        "orc_getLFromX": (d, [p, d]), "orc_getXFromL": (d, [p, d]),
        "orc_init_solver": (i, [p, d, d, d, d]),
        "orc_add_species": (i, [p, d, d, i, i]),
        "orc_initialize_n_ave": (None, [p, i, d]),
        "orc_set_weight": (None, [p, i, d]), "orc_get_weight": (d, [p, i]),
        "orc_load_particles": (None, [p, i, C.POINTER(C.c_int32), d, d, i, d]),
        "orc_phase_space": (dp, [p, i]), "orc_final_idx": (i, [p, i]),
        "orc_get_Np": (None, [p, i, ip]), "orc_set_Np": (None, [p, i, ip]),
        "orc_get_counters": (None, [p, i, ip, ip, dp, dp, dp]),
        "orc_get_ref_record": (None, [p, i, i, ip, i]),
        "orc_get_move_stats": (None, [p, i, dp, dp, C.POINTER(ll), C.POINTER(ll), C.POINTER(ll)]),
        "orc_config_solver": (None, [p, d, d, i, i, d, i]),
        "orc_deposit_rho": (None, [p]), "orc_solve_tridiag_poisson": (None, [p, d]),
        "orc_solve_initial_potential": (None, [p, d]), "orc_make_const_source_term": (None, [p]),
        "orc_make_half_time_efield": (None, [p]), "orc_deposit_J": (None, [p, d]),
        "orc_solve_tridiag_ampere": (None, [p, d]), "orc_move_particles": (None, [p, d]),
        "orc_reset_voltage": (None, [p]), "orc_set_rf_voltage": (None, [p, d]),
        "orc_solve_potential": (i, [p, d, dp, dp, d]),
        "orc_solve_div_ampere_anderson": (None, [p, d]),
        "orc_get_field": (i, [p, i, dp]), "orc_set_field": (i, [p, i, dp]),
        "orc_get_cic_walls": (None, [p, dp]), "orc_get_rf": (None, [p, dp, dp, ip, ip]),
        "orc_get_particle_time_step": (None, [p, i, dp]),
        "orc_total_PE": (d, [p, i]), "orc_total_KE": (d, [p, i]), "orc_KE_ave": (d, [p, i]),
        "orc_total_momentum": (None, [p, i, dp]), "orc_JdotE": (d, [p, d]),
        "orc_energy_from_boundary": (d, [p, d]), "orc_gauss_error": (d, [p]),
        "orc_charge_error": (d, [p, dp, d]), "orc_lost_energy_sum": (d, [p, i]),
        "orc_get_solver_stats": (None, [p, ip, ip, ip, C.POINTER(ll), C.POINTER(ll), dp, dp]),
        "orc_get_norm_history": (i, [p, dp, i]),
        "orc_solve_normal_equation": (None, [dp, dp, i, i, dp]),
        "orc_ran2": (d, [C.POINTER(C.c_int32)]), "orc_randPCG": (d, [C.POINTER(C.c_uint64)]),
        "orc_species_scalars": (None, [p, i, dp]),
        "orc_max_threads": (i, []),
        "orc_p_init_solver": (i, [p, d, d, d]),
        "orc_p_deposit_rho": (None, [p]), "orc_p_solve_tridiag_poisson": (None, [p, d]),
        "orc_p_deposit_J": (None, [p, d]), "orc_p_move_particles": (None, [p, d]),
        "orc_p_solve_tridiag_ampere": (None, [p, d]), "orc_p_reset_voltage": (None, [p]),
        "orc_p_solve_potential": (i, [p, d, dp, dp]), "orc_p_solve_div_ampere_anderson": (None, [p, d]),
        "orc_p_total_PE": (d, [p, i]), "orc_p_gauss_error": (d, [p]), "orc_p_charge_error": (d, [p, dp, d]),
        "orc_p_aveJ": (d, [p]), "orc_reset_sweep_counters": (None, [p]),
        "orc_velocity_extrema": (None, [p, i, dp]), "orc_velocity_histograms": (None, [p, i, i, d, d, d, dp, dp, dp]),
        "orc_load_particle_density": (None, [p, i]), "orc_particle_density_array": (None, [p, i, dp]),
        "orc_add_maxwellian_lost_particles": (i, [p, d, d, C.POINTER(C.c_int32), dp]),
        "orc_reflux_particles": (None, [p, d, d, C.POINTER(C.c_int32), dp]),
        "orc_inject_at_boundary": (i, [p, d, d, i, i, d, C.POINTER(C.c_int32)]),
        "orc_inject_uniform_flux": (i, [p, d, d, i, i, d, C.POINTER(C.c_int32)]),
        "orc_maxwellian_heating": (None, [p, i, d, d, C.POINTER(C.c_int32)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _libs[native] = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


FIELDS = {"phi": 0, "phi_f": 1, "J": 2, "rho": 3, "EField": 4, "a_tri": 5, "b_tri": 6, "c_tri": 7, "EField_f": 8, "EField_half": 9}
NGP, CIC, NGP_PERIODIC = 0, 1, 2


def default_seeds(T):
    """the deterministic seeding commented at mod_BasicFunctions.f90:119: 12345*k + 11*(k-1), k = 1..T"""
    return np.array([12345 * k + 11 * (k - 1) for k in range(1, T + 1)], dtype=np.int32)


class Oracle:
    """One reference program state: globalWorld + globalSolver + globalParticleList (mod_nonLinSolvers.f90:26-28)."""

    def __init__(self, scheme, Nx, L_domain, gridType=0, del_x=0.0, bcL=1, bcR=1, smooth=False, T=1, native=False):
        self.L = lib(native)
        # NGP_PERIODIC (E-field form, OpenMP/NGP_Periodic): the NGP grid without smoothing + the E-form solver/mover routines
        self.eform = scheme == NGP_PERIODIC
        if self.eform:
            scheme, smooth = NGP, False
        self.h = self.L.orc_create(scheme, Nx, L_domain, gridType, del_x, bcL, bcR, int(smooth), T)
        if not self.h:
            raise ValueError("oracle: grid construction failed (the reference would stop)")
        self.scheme, self.Nx, self.Nh, self.T = scheme, Nx, self.L.orc_Nh(self.h), T
        self.L_domain = self.L.orc_L_domain(self.h)
        self.n_species = 0

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- world
    def grid(self):
        nd = self.Nh if self.scheme == NGP else self.Nx
        nb = self.Nx if self.scheme == NGP else self.Nh
        g, dx, bc = np.zeros(self.Nx), np.zeros(nd), np.zeros(nb, dtype=np.int32)
        self.L.orc_get_grid(self.h, _dp(g), _dp(dx), _ip(bc))
        return g, dx, bc

    def slices(self):
        hl, nl = np.zeros(2 * self.T, dtype=np.int32), np.zeros(2 * self.T, dtype=np.int32)
        a, b = C.c_int(), C.c_int()
        self.L.orc_get_slices(self.h, _ip(hl), _ip(nl), C.byref(a), C.byref(b))
        return hl[:2 * a.value].reshape(-1, 2), nl[:2 * b.value].reshape(-1, 2)

    def getLFromX(self, x):
        return self.L.orc_getLFromX(self.h, x)

    def getXFromL(self, l):
        return self.L.orc_getXFromL(self.h, l)

    # ---- solver
    def init_solver(self, leftV, rightV, RF_frequency=0.0, t0=0.0):
        if self.eform:
            rc = self.L.orc_p_init_solver(self.h, leftV, rightV, RF_frequency)
        else:
            rc = self.L.orc_init_solver(self.h, leftV, rightV, RF_frequency, t0)
        if rc:
            raise ValueError("oracle: potentialSolver constructor would stop (%d)" % rc)

    def config_solver(self, eps_a=1e-8, eps_r=0.0, solverType=0, m_Anderson=10, Beta_k=0.5, maxIter=100):
        self.L.orc_config_solver(self.h, eps_a, eps_r, solverType, m_Anderson, Beta_k, maxIter)

    def get(self, name):
        out = np.zeros(max(self.Nx, self.Nh) + 2)
        n = self.L.orc_get_field(self.h, FIELDS[name], _dp(out))
        return out[:n].copy()

    def set(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self.L.orc_set_field(self.h, FIELDS[name], _dp(a))

    def cic_walls(self):
        o = np.zeros(5)
        self.L.orc_get_cic_walls(self.h, _dp(o))
        return dict(boundPhi=o[0:2].copy(), sourceTermVals=o[2:4].copy(), RF_voltage=o[4])

    def rf(self):
        a, b, c, d = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        self.L.orc_get_rf(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return dict(RF_rad_frequency=a.value, RF_half_amplitude=b.value, RF_bool=bool(c.value), RF_indx=d.value)

    # ---- particles
    def add_species(self, mass, q, N_p_per_thread, factor=1):
        sp = self.L.orc_add_species(self.h, mass, q, N_p_per_thread, N_p_per_thread * factor)
        self.n_species = sp + 1
        return sp

    def initialize_n_ave(self, sp, n_ave):
        self.L.orc_initialize_n_ave(self.h, sp, n_ave)

    def set_weight(self, sp, w_p):
        self.L.orc_set_weight(self.h, sp, w_p)

    def load_particles(self, sp, seeds, T_eV, alpha=0.0, distType=0, v_drift=0.0):
        assert seeds.dtype == np.int32 and seeds.size == self.T
        self.L.orc_load_particles(self.h, sp, seeds.ctypes.data_as(C.POINTER(C.c_int32)), T_eV, alpha, distType, v_drift)

    def phase_space(self, sp):
        """numpy VIEW of phaseSpace(4, finalIdx, T) as a C-ordered [T, finalIdx, 4] array"""
        fi = self.L.orc_final_idx(self.h, sp)
        ptr = self.L.orc_phase_space(self.h, sp)
        return np.ctypeslib.as_array(ptr, shape=(self.T, fi, 4))

    def N_p(self, sp):
        n = np.zeros(self.T, dtype=np.int32)
        self.L.orc_get_Np(self.h, sp, _ip(n))
        return n

    def set_N_p(self, sp, n):
        n = np.ascontiguousarray(n, dtype=np.int32)
        self.L.orc_set_Np(self.h, sp, _ip(n))

    def counters(self, sp):
        T = self.T
        de, re_ = np.zeros(T, dtype=np.int32), np.zeros(T, dtype=np.int32)
        wl, el, ml = np.zeros((T, 2)), np.zeros((T, 2)), np.zeros((T, 2))
        self.L.orc_get_counters(self.h, sp, _ip(de), _ip(re_), _dp(wl), _dp(el), _dp(ml))
        return dict(delIdx=de, refIdx=re_, wallLoss=wl, energyLoss=el, momentumLoss=ml)

    def ref_record(self, sp, t, n):
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.L.orc_get_ref_record(self.h, sp, t, _ip(out), n)
        return out[:n]

    def move_stats(self, sp):
        a, b = C.c_double(), C.c_double()
        c, d, f = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self.L.orc_get_move_stats(self.h, sp, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(f))
        return dict(numSubStepsAve=a.value, numFuncEvalAve=b.value, substepsMove=c.value, funcEvalsMove=d.value, substepsDeposit=f.value)

    def species_scalars(self, sp):
        o = np.zeros(5)
        self.L.orc_species_scalars(self.h, sp, _dp(o))
        return dict(mass=o[0], q=o[1], w_p=o[2], q_over_m=o[3], q_times_wp=o[4])

    # ---- EEDF / velocity histograms (mod_simulation.f90:433-500)
    def velocityExtrema(self, sp):
        o = np.zeros(3)
        self.L.orc_velocity_extrema(self.h, sp, _dp(o))
        return o                                       # VMax, EMax, Emin

    def velocityHistograms(self, sp, binNumber, VMax, Emin, EMax, VHist, EHist):
        """adds one pass to VHist[2*binNumber], EHist[binNumber]; returns E_grid"""
        E_grid = np.zeros(binNumber)
        self.L.orc_velocity_histograms(self.h, sp, binNumber, VMax, Emin, EMax, _dp(VHist), _dp(EHist), _dp(E_grid))
        return E_grid

    # ---- density diagnostic (mod_Scheme.f90: loadParticleDensity, WriteParticleDensity)
    def loadParticleDensity(self, reset=True):
        self.L.orc_load_particle_density(self.h, int(bool(reset)))

    def particleDensityArray(self, sp):
        """the array WriteParticleDensity writes for species sp (consumes the loaded densities, like the reference)"""
        out = np.zeros(self.Nx)
        self.L.orc_particle_density_array(self.h, sp, _dp(out))
        return out

    # ---- mod_particleInjection.f90 (irand: the per-thread ran2 states, int32[T], advanced in place)
    def _ir(self, irand):
        assert irand.dtype == np.int32 and irand.size == self.T
        return irand.ctypes.data_as(C.POINTER(C.c_int32))

    def addMaxwellianLostParticles(self, T_e, T_i, irand):
        """returns what the call adds to SUM(energyAddColl)"""
        en = C.c_double()
        if self.L.orc_add_maxwellian_lost_particles(self.h, T_e, T_i, self._ir(irand), C.byref(en)):
            raise RuntimeError("oracle: addMaxwellianLostParticles overflows a slab")
        return en.value

    def refluxParticles(self, T_e, T_i, irand):
        en = C.c_double()
        self.L.orc_reflux_particles(self.h, T_e, T_i, self._ir(irand), C.byref(en))
        return en.value

    def injectAtBoundary(self, T_e, del_t, numLow, numHigh, injectionR, irand):
        if self.L.orc_inject_at_boundary(self.h, T_e, del_t, numLow, numHigh, injectionR, self._ir(irand)):
            raise RuntimeError("oracle: injectAtBoundary overflows a slab")

    def injectUniformFlux(self, T_e, T_i, numLow, numHigh, injectionR, irand):
        if self.L.orc_inject_uniform_flux(self.h, T_e, T_i, numLow, numHigh, injectionR, self._ir(irand)):
            raise RuntimeError("oracle: injectUniformFlux overflows a slab")

    def maxwellianHeating(self, sp, compValue, T_eV, irand):
        self.L.orc_maxwellian_heating(self.h, sp, compValue, T_eV, self._ir(irand))

    def particle_time_step(self, sp):
        o = np.zeros(self.Nx)
        self.L.orc_get_particle_time_step(self.h, sp, _dp(o))
        return o

    # ---- the path (E-form routines when self.eform)
    def depositRho(self): (self.L.orc_p_deposit_rho if self.eform else self.L.orc_deposit_rho)(self.h)
    def solve_tridiag_Poisson(self, t=0.0): (self.L.orc_p_solve_tridiag_poisson if self.eform else self.L.orc_solve_tridiag_poisson)(self.h, t)

    def solveInitialPotential(self, t=0.0):
        if self.eform:
            self.depositRho(); self.solve_tridiag_Poisson(t)
        else:
            self.L.orc_solve_initial_potential(self.h, t)

    def makeConstSourceTerm(self):
        if not self.eform:
            self.L.orc_make_const_source_term(self.h)

    def makeHalfTimeEField(self): self.L.orc_make_half_time_efield(self.h)
    def depositJ(self, del_t): (self.L.orc_p_deposit_J if self.eform else self.L.orc_deposit_J)(self.h, del_t)
    def solve_tridiag_Ampere(self, del_t): (self.L.orc_p_solve_tridiag_ampere if self.eform else self.L.orc_solve_tridiag_ampere)(self.h, del_t)
    def moveParticles(self, del_t): (self.L.orc_p_move_particles if self.eform else self.L.orc_move_particles)(self.h, del_t)
    def resetVoltage(self): (self.L.orc_p_reset_voltage if self.eform else self.L.orc_reset_voltage)(self.h)
    def setRFVoltage(self, t):
        if not self.eform:
            self.L.orc_set_rf_voltage(self.h, t)
    def solveDivAmpereAnderson(self, del_t): (self.L.orc_p_solve_div_ampere_anderson if self.eform else self.L.orc_solve_div_ampere_anderson)(self.h, del_t)

    def solvePotential(self, del_t, remainDel_t, currDel_t, timeCurrent):
        r, c = C.c_double(remainDel_t), C.c_double(currDel_t)
        if self.eform:
            rc = self.L.orc_p_solve_potential(self.h, del_t, C.byref(r), C.byref(c))
        else:
            rc = self.L.orc_solve_potential(self.h, del_t, C.byref(r), C.byref(c), timeCurrent)
        if rc:
            raise RuntimeError("oracle: solvePotential would stop (%d)" % rc)
        return r.value, c.value

    def solver_stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        d, f = C.c_longlong(), C.c_longlong()
        g, hh = C.c_double(), C.c_double()
        self.L.orc_get_solver_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(f), C.byref(g), C.byref(hh))
        return dict(iterNumPicard=a.value, iterNumAdaptiveSteps=b.value, amountTimeSplits=c.value, sweeps=d.value, substeps=f.value,
                    moverSeconds=g.value, potSeconds=hh.value)

    def norm_history(self):
        o = np.zeros(1024)
        n = self.L.orc_get_norm_history(self.h, _dp(o), 1024)
        return o[:n].copy()

    # ---- JFNK (mod_nonLinSolvers.f90:218-282) with the numpy NITSOL restatement (oracle/nitsol_np.py)
    def funcNitsol(self, xcur, del_t):
        """fcur = xcur - G(xcur); leaves phi_f (E-form: EField_f) = G(xcur) like the reference
        (mod_nonLinSolvers.f90:226-236 ; NGP_Periodic/src/mod_nonLinSolvers.f90:209-221)"""
        unk = "EField_f" if self.eform else "phi_f"
        self.set(unk, xcur)
        self.depositJ(del_t)
        self.solve_tridiag_Ampere(del_t)
        return np.asarray(xcur) - self.get(unk)

    def solveJFNK(self, del_t, eps_a, eps_r, maxIter, m_Anderson):
        from . import nitsol_np
        self.makeConstSourceTerm()
        nit = nitsol_np.Nitsol(maxIter, m_Anderson)
        self.jfnk_substeps = 0
        def f(x):
            r = self.funcNitsol(x, del_t)
            self.jfnk_substeps += sum(self.move_stats(sp)["substepsDeposit"] for sp in range(self.n_species))
            return r
        x, iterm = nit.solve(f, self.get("EField_f" if self.eform else "phi_f"), eps_a, eps_r, 1.e-12)
        if iterm == 0:
            if self.eform:                       # NGP_Periodic solveJFNK re-deposits J before the move (:251-254)
                self.depositJ(del_t)
            self.moveParticles(del_t)
        return dict(nit.info, iterm=iterm, fnorms=list(nit.fnorms))

    # ---- diagnostics (main.f90:58-114)
    def getTotalPE(self, future): return (self.L.orc_p_total_PE if self.eform else self.L.orc_total_PE)(self.h, int(future))
    def getTotalKE(self, sp): return self.L.orc_total_KE(self.h, sp)
    def getKEAve(self, sp): return self.L.orc_KE_ave(self.h, sp)
    def getJdotE(self, del_t): return self.L.orc_JdotE(self.h, del_t)
    def getEnergyFromBoundary(self, del_t): return 0.0 if self.eform else self.L.orc_energy_from_boundary(self.h, del_t)
    def getError_tridiag_Poisson(self): return (self.L.orc_p_gauss_error if self.eform else self.L.orc_gauss_error)(self.h)
    def aveJ(self): return self.L.orc_p_aveJ(self.h)
    def lostEnergySum(self, sp): return self.L.orc_lost_energy_sum(self.h, sp)

    def localTemperature(self, sp, numCells):
        out = np.zeros(numCells)
        self.L.orc_local_temperature(self.h, sp, numCells, _dp(out))
        return out

    def getTotalMomentum(self, sp):
        o = np.zeros(3)
        self.L.orc_total_momentum(self.h, sp, _dp(o))
        return o

    def getChargeContinuityError(self, rho_i, del_t):
        r = np.ascontiguousarray(rho_i, dtype=np.float64)
        return (self.L.orc_p_charge_error if self.eform else self.L.orc_charge_error)(self.h, _dp(r), del_t)


def plasma_del_t(n_ave, fractionFreq, del_t_deck):
    """constants.f90:106-107"""
    w_p = np.sqrt(n_ave * (e ** 2) / m_e / eps_0)
    return min(fractionFreq * 1.0 / w_p, del_t_deck)


def debye_length(T_e, n_e):
    """mod_BasicFunctions.f90:16-20"""
    return np.sqrt(eps_0 * T_e / n_e / e)


def make_deck_case(scheme, Nx, L_domain, n_ave, T_e, T_i, ppt, T, bcL, bcR, Vl, Vr, RF_freq, smooth, ion_amu=4.002602,
                   gridType=0, del_x=0.0, eps_a=1e-8, eps_r=0.0, m_Anderson=10, Beta_k=0.5, maxIter=100, t0=0.0, factor=1,
                   seeds=None, native=False, ion_charge=1.0, alpha=0.0, distType=0, v_drift=0.0, solverType=0):
    """Builds the program state exactly as main.f90:25-40 does from a deck (electrons + one ion species),
    with the deterministic seeds of mod_BasicFunctions.f90:119."""
    del_x = max(0.1 * debye_length(T_e, n_ave), del_x)     # mod_domain.f90:371
    o = Oracle(scheme, Nx, L_domain, gridType, del_x, bcL, bcR, smooth, T, native=native)
    o.init_solver(Vl, Vr, RF_freq, t0)
    seeds = default_seeds(T) if seeds is None else seeds.copy()
    sp_e = o.add_species(m_e, -1.0 * e, ppt, factor)
    o.initialize_n_ave(sp_e, n_ave)
    o.load_particles(sp_e, seeds, T_e, alpha, distType, v_drift)
    m_ion = ion_amu * m_amu - ion_charge * m_e              # mod_particle.f90:334
    sp_i = o.add_species(m_ion, ion_charge * e, ppt, factor)
    o.initialize_n_ave(sp_i, n_ave)
    o.load_particles(sp_i, seeds, T_i, alpha, distType, v_drift)
    o.config_solver(eps_a, eps_r, solverType, m_Anderson, Beta_k, maxIter)
    o.solveInitialPotential(t0)
    return o
