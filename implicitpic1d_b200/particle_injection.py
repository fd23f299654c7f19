"""Host-side mirror of `mod_particleInjection` (OpenMP/SharedModules/src/mod_particleInjection.f90): same procedure names and
argument lists; the bodies call the C ABI, the work happens on the device where the mover left the particles
(csrc/injection.cu).  `irand` (the per-thread ran2 states of the reference) is accepted for signature parity and ignored: the
device draws from counter-based PCG32 streams (`setInjectionSeed`).

Module state as the reference's (`:9-12`): the five switches, injectionFlux / FractionFreqHeating and the per-call expectation
of injected pairs; `energyAddColl` here is the running SUM(energyAddColl) (the reference keeps one entry per OpenMP thread and
only ever uses the sum, mod_simulation.f90:190)."""
import ctypes as C
import math
import os
from . import _lib

addLostPartBool = refluxPartBool = injectionBool = uniformInjectionBool = heatingBool = False
injectionFlux = 0.0
FractionFreqHeating = 1.0
numFluxPart = 0.0                       # expected pairs per call for the whole process (numThread x the reference's per-thread figure)
numFluxParticlesLow = numFluxParticlesHigh = 0
injectionR = 0.0
energyAddColl = 0.0


def _solver(particleList):
    s = particleList[0]._solver if isinstance(particleList, (list, tuple)) else particleList._solver
    if s is None:
        raise RuntimeError("particles are not bound to a potentialSolver")
    return s


def _refresh_counts(particleList):
    L = _lib.load()
    for p in particleList:
        n = C.c_int64()
        _lib.check(L.ipic_particle_count(p._solver._h, p._index, C.byref(n)))
        p.N_p_total = n.value


def setInjectionSeed(solver, seed):
    _lib.check(_lib.load().ipic_set_injection_seed(solver._h, int(seed) & 0xFFFFFFFFFFFFFFFF))


def addMaxwellianLostParticles(particleList, T_e, T_i, irand=None, world=None):
    """:16-40.  Returns the number of electron-ion pairs added (= the electrons' SUM(delIdx) of the last moveParticles)."""
    global energyAddColl
    s = _solver(particleList)
    en, n = C.c_double(), C.c_int64()
    _lib.check(_lib.load().ipic_add_maxwellian_lost_particles(s._h, float(T_e), float(T_i), C.byref(en), C.byref(n)))
    energyAddColl += en.value
    _refresh_counts(particleList[:2])
    return n.value


def refluxParticles(particleList, T_e, T_i, irand=None, world=None):
    """:42-69"""
    global energyAddColl
    s = _solver(particleList)
    en = C.c_double()
    _lib.check(_lib.load().ipic_reflux_particles(s._h, float(T_e), float(T_i), C.byref(en)))
    energyAddColl += en.value


def injectAtBoundary(particleList, T_e, T_i, irand=None, world=None, del_t=None):
    """:71-130.  Returns the number of pairs injected."""
    if del_t is None:
        raise TypeError("injectAtBoundary needs del_t")
    s = _solver(particleList)
    n = C.c_int64()
    _lib.check(_lib.load().ipic_inject_at_boundary(s._h, float(T_e), float(T_i), float(del_t), float(numFluxPart), C.byref(n)))
    _refresh_counts(particleList[:2])
    return n.value


def injectUniformFlux(particleList, T_e, T_i, irand=None, world=None):
    """:132-158"""
    s = _solver(particleList)
    n = C.c_int64()
    _lib.check(_lib.load().ipic_inject_uniform_flux(s._h, float(T_e), float(T_i), float(numFluxPart), C.byref(n)))
    _refresh_counts(particleList[:2])
    return n.value


def maxwellianHeating(species, irand=None, FracFreq=1.0, T=0.0, currDel_t=1.0, del_t=1.0):
    """:160-177.  Returns the number of particles whose velocity was redrawn."""
    s = _solver(species)
    compValue = FracFreq * (currDel_t / del_t) / FractionFreqHeating
    n = C.c_int64()
    _lib.check(_lib.load().ipic_maxwellian_heating(s._h, species._index, float(compValue), float(T), C.byref(n)))
    return n.value


def setInjectionInputs(addLost=False, reflux=False, injection=False, uniformInjection=False, heating=False, flux=0.0,
                       fractionFreqHeating=1.0, del_t=0.0, w_p=1.0, numThread=1):
    """The state readInjectionInputs leaves behind (:205-243), from values instead of a file."""
    global addLostPartBool, refluxPartBool, injectionBool, uniformInjectionBool, heatingBool, injectionFlux, FractionFreqHeating
    global numFluxPart, numFluxParticlesLow, numFluxParticlesHigh, injectionR
    addLostPartBool, refluxPartBool, injectionBool = bool(addLost), bool(reflux), bool(injection)
    uniformInjectionBool = bool(uniformInjection) and not injectionBool      # only read when injectionBool is off (:203-206)
    heatingBool = bool(heating)
    injectionFlux, FractionFreqHeating = float(flux), float(fractionFreqHeating)
    numFluxPart = 0.0
    numFluxParticlesLow = numFluxParticlesHigh = 0
    injectionR = 0.0
    if injectionBool or uniformInjectionBool:
        per_thread = injectionFlux * del_t / w_p / float(numThread)         # :232, :241
        numFluxParticlesLow = int(math.floor(per_thread))
        numFluxParticlesHigh = numFluxParticlesLow + 1
        injectionR = per_thread - float(numFluxParticlesLow)
        numFluxPart = per_thread * numThread


def readInjectionInputs(InjFilename, w_p, del_t=0.0, numThread=1, input_dir="../InputData"):
    """readInjectionInputs (:181-252): five lines `flag [value]`; the uniform-injection line is only read when boundary injection
    is off.  del_t and numThread are module globals in the reference (constants.f90)."""
    with open(os.path.join(input_dir, InjFilename)) as f:
        lines = [ln.replace(",", " ").split() for ln in f if ln.strip()]
    it = iter(lines)
    addLost = int(next(it)[0]) == 1
    reflux = int(next(it)[0]) == 1
    tok = next(it)
    injection, flux = int(tok[0]) == 1, float(tok[1].lower().replace("d", "e"))
    uniform = False
    if not injection:
        tok = next(it)
        uniform, flux = int(tok[0]) == 1, float(tok[1].lower().replace("d", "e"))
    tok = next(it)
    heating, frac = int(tok[0]) == 1, float(tok[1].lower().replace("d", "e"))
    setInjectionInputs(addLost, reflux, injection, uniform, heating, flux, frac, del_t, w_p, numThread)


def writeParticleInjectionInputs(dirName):
    """:254-261"""
    def L(b):
        return "T" if b else "F"
    with open(os.path.join(dirName, "ParticleInjectionInput.dat"), "w") as f:
        f.write("AddLostPartBool, refluxPartBool, injectionBool, uniformInjectionBool, heatingBool, injectionFlux, FractionFreqHeating\n")
        f.write(" ".join(L(b) for b in (addLostPartBool, refluxPartBool, injectionBool, uniformInjectionBool, heatingBool)))
        f.write(" %16.8E %16.8E\n" % (injectionFlux, FractionFreqHeating))
