"""Host-side mirror of `mod_targetParticle` and `mod_NullCollision` (OpenMP/SharedModules/src/mod_targetParticle.f90,
mod_NullCollision.f90): `targetParticle`, `nullCollision` (constructor + generateCollision) and `readNullCollisionInputs`.

The collision work itself runs on the device (csrc/collisions.cu) through the C ABI; this file holds the table reader (host
logic, numpy) and the objects a caller of the reference would expect.  There is no CPU collision path here."""
import ctypes as C
import os
import re
import numpy as np
from . import _lib
from . import constants as K

COLLISION_TYPES = {"ELASTIC": 1, "IONIZATION": 2, "EXCITATION": 3, "CHARGEEXCHANGE": 4, "DISSOCIATION": 5}


class targetParticle:
    """type(targetParticle), mod_targetParticle.f90:13-37: mass [kg], density [m^-3], temperature [K]"""

    def __init__(self, particleName, mass, density, temperature):
        self.name = particleName.strip()
        self.mass, self.density, self.temperature = float(mass), float(density), float(temperature)
        self.v_therm = np.sqrt(self.temperature * K.k_B / self.mass)
        self._index, self._solver = -1, None

    def _bind(self, solver):
        if self._solver is solver:
            return
        idx = C.c_int32()
        _lib.check(_lib.load().ipic_add_target_particle(solver._h, self.mass, self.density, self.temperature, C.byref(idx)))
        self._index, self._solver = idx.value, solver


def _num(tok):
    return float(tok.replace("d", "e").replace("D", "E"))


def parse_collision_table(path):
    """One LXCat-style table file as the reference reads it (mod_NullCollision.f90:693-811): a list of reactions, each
    {reactants: [names], products: [names], threshold, type, E (eV, scaled), sigma (m^2, scaled)}."""
    with open(path) as f:
        lines = [ln.strip() for ln in f if ln.strip()]
    out, i = [], 0
    while i < len(lines):
        head = lines[i].split()[0]
        if head[:3].upper() == "END":
            break
        if head[:4].upper() != "REAC":
            i += 1
            continue
        eq = lines[i + 1].split()[0]
        if not eq.startswith("["):
            raise ValueError("%s: no chemical reaction after the REACTION label" % path)
        lhs, rhs = eq.split(">", 1)
        reactants, products = re.findall(r"\[([^\]]+)\]", lhs), re.findall(r"\[([^\]]+)\]", rhs)
        threshold = _num(lines[i + 2].replace(",", " ").split()[0])
        sc = lines[i + 3].replace(",", " ").split()
        e_scale, s_scale = _num(sc[0]), _num(sc[1])
        kind = lines[i + 4].split()[0].upper()
        if kind not in COLLISION_TYPES:
            raise ValueError("%s: unknown collision type %r" % (path, kind))
        j = i + 4
        while not lines[j].startswith("----"):
            j += 1
        j += 1
        E, S = [], []
        while not lines[j].startswith("----"):
            a, b = lines[j].replace(",", " ").split()[:2]
            E.append(_num(a) * e_scale); S.append(_num(b) * s_scale)
            j += 1
        out.append(dict(reactants=reactants, products=products, threshold=threshold, type=COLLISION_TYPES[kind],
                        E=np.array(E), sigma=np.array(S)))
        i = j + 1
    return out


def _interp_reference(E_new, E_old, s_old, threshold):
    """cross section on the merged energy axis (:858-885): zero below the threshold, linear inside the table, constant
    beyond its last point"""
    if E_new.size and E_old.size and np.any((E_new >= threshold) & (E_new < E_old[0])):
        raise ValueError("cross-section table starts above its threshold (the reference reads out of bounds here)")
    # np.interp evaluates lo + slope * (x - x_lo); the reference blends s_lo (1 - t) + s_hi t: do the blend explicitly
    idx = np.clip(np.searchsorted(E_old, E_new, side="right") - 1, 0, E_old.size - 2)
    t = (E_new - E_old[idx]) / (E_old[idx + 1] - E_old[idx])
    val = s_old[idx] * (1 - t) + s_old[idx + 1] * t
    exact = E_old[idx] == E_new
    val = np.where(exact, s_old[idx], val)
    val = np.where(E_new >= E_old[-1], s_old[-1], val)
    return np.where(threshold > E_new, 0.0, val)


class nullCollision:
    """type(nullCollision), mod_NullCollision.f90:18-108.  Constructor arguments as the reference's (arrays 0-based numpy,
    sigmaArray shaped (lengthArrays, numberCollisions), reactantsIndx / productsIndx 1-based numbers as in the Fortran)."""

    def __init__(self, numberReactants, numberCollisions, lengthArrays, red_mass, sumMass, redMassTripleProducts, energyArray, sigmaArray,
                 energyThreshold, collisionType, reactantsIndx, numberProducts, productsIndx, seed=0x1D1C0111):
        self.numberReactants, self.numberCollisions, self.lengthArrays = int(numberReactants), int(numberCollisions), int(lengthArrays)
        self.reducedMass, self.sumMass, self.reducedMassIonization = float(red_mass), float(sumMass), float(redMassTripleProducts)
        self._energy = np.ascontiguousarray(energyArray, dtype=np.float64)
        self._sigma = np.asfortranarray(sigmaArray, dtype=np.float64)
        self._thr = np.ascontiguousarray(energyThreshold, dtype=np.float64)
        self._type = np.ascontiguousarray(collisionType, dtype=np.int32)
        self.reactantsIndx = np.ascontiguousarray(reactantsIndx, dtype=np.int32)
        self._nprod = np.ascontiguousarray(numberProducts, dtype=np.int32)
        self._pidx = np.asfortranarray(np.asarray(productsIndx, dtype=np.int32).reshape(3, self.numberCollisions, order="F"))
        if self._sigma.shape != (self.lengthArrays, self.numberCollisions):
            raise ValueError("sigmaArray must have shape (lengthArrays, numberCollisions)")
        self.seed = int(seed)
        self._set, self._solver = -1, None

    def _bind(self, solver):
        """creates the device-side collision set; the species / target numbers refer to the solver's particle list and to the
        target particles bound to it"""
        if self._solver is solver:
            return
        idx = C.c_int32()
        L = _lib.load()
        _lib.check(L.ipic_null_collision_create(solver._h, self.numberCollisions, self.lengthArrays, self.reducedMass, self.sumMass,
                                                self.reducedMassIonization, self._energy.ctypes.data_as(_lib._dp),
                                                self._sigma.ctypes.data_as(_lib._dp), self._thr.ctypes.data_as(_lib._dp),
                                                self._type.ctypes.data_as(_lib._ip), self.reactantsIndx.ctypes.data_as(_lib._ip),
                                                self._nprod.ctypes.data_as(_lib._ip), self._pidx.ctypes.data_as(_lib._ip), self.seed, C.byref(idx)))
        self._set, self._solver = idx.value, solver
        self._refresh()

    def _refresh(self):
        st = _lib.CollisionStats()
        _lib.check(_lib.load().ipic_get_collision_stats(self._solver._h, self._set, C.byref(st)))
        n = self.numberCollisions
        self.collisionType = np.array(st.collisionType[:n])                 # constructor order (largest sigma*v first)
        self.energyThreshold = np.array(st.energyThreshold[:n])
        self.sigmaVMax, self.minEnergy, self.maxEnergy = st.sigmaVMax, st.minEnergy, st.maxEnergy
        self.P_null, self.numberSelected, self.ionizations = st.P_null, st.numberSelected, st.ionizations
        self.totalAmountCollisions = np.array(st.totalAmountCollisions[:n], dtype=np.int64)
        self.totalEnergyLoss = np.array(st.totalEnergyLoss[:n])
        self.totalIncidentEnergy = np.array(st.totalIncidentEnergy[:n])
        self.totalNumberCollidableParticles = st.totalNumberCollidableParticles
        return st

    def generateCollision(self, particleList, targetParticleList, numberChargedParticles=None, numberBinaryCollisions=None, irand=None,
                          del_t=None):
        """generateCollision (mod_NullCollision.f90:110-237); `irand` is accepted for signature parity and ignored (the device
        draws from its own counter-based streams)."""
        if del_t is None:
            raise TypeError("generateCollision needs del_t")
        solver = particleList[self.reactantsIndx[0] - 1]._solver
        if solver is None:
            raise RuntimeError("particles are not bound to a potentialSolver")
        for t in targetParticleList:
            t._bind(solver)
        self._bind(solver)
        _lib.check(_lib.load().ipic_generate_collision(solver._h, self._set, float(del_t)))
        self._refresh()

    # ---- diagnostics files (mod_NullCollision.f90:564-632); collisions are numbered in the constructor's order (largest sigma*v first)
    def _order(self):
        """the constructor's permutation (:55-70): ascending stable sort of max(sigma * v), reversed"""
        v = np.sqrt(2.0 * self._energy * K.e / self.reducedMass)
        svmax = np.max(self._sigma * v[:, None], axis=0)
        return np.argsort(svmax, kind="stable")[::-1]

    def _dir(self, particleList, targetParticleList, dirName):
        return os.path.join(dirName, "BinaryCollisions", "%s_on_%s" % (particleList[self.reactantsIndx[0] - 1].name,
                                                                      targetParticleList[self.reactantsIndx[1] - 1].name))

    def writeCollisionCrossSection(self, part, dirName):
        """:564-577: raw stream files of energyArray and sigmaArray (column-major, constructor order)"""
        os.makedirs(os.path.join(dirName, "CrossSections"), exist_ok=True)
        self._energy.astype("<f8").tofile(os.path.join(dirName, "CrossSections", "IncidentPart_%s_energy.dat" % part.name))
        np.asfortranarray(self._sigma[:, self._order()]).astype("<f8").T.tofile(os.path.join(dirName, "CrossSections", "IncidentPart_%s_sigma.dat" % part.name))

    def writeCollisionProperties(self, particleList, targetParticleList, dirName):
        """:579-609"""
        d = self._dir(particleList, targetParticleList, dirName)
        os.makedirs(d, exist_ok=True)
        order = self._order()
        with open(os.path.join(d, "CollisionProperties.dat"), "w") as f:
            f.write("Coll #, collType, E_thres (eV), maxSigma (m^2), EatMaxSigma (eV)\n")
            for i, j in enumerate(order):
                k = int(np.argmax(self._sigma[:, j]))
                f.write("%3d %3d %16.8E %16.8E %16.8E\n" % (i + 1, self._type[j], self._thr[j], self._sigma[k, j], self._energy[k]))
        for i in range(self.numberCollisions):
            with open(os.path.join(d, "CollisionDiag_%d.dat" % (i + 1)), "w") as f:
                f.write("CollRatio, AveEnergyLoss (eV), AveIncidentEnergy (eV), P_loss(W/m^2), aveCollFreq (Hz/m^2)\n")

    def writeCollisionDiag(self, particleList, targetParticleList, dirName, timeInterval):
        """:611-632: one line per collision appended to its CollisionDiag_<i>.dat"""
        d = self._dir(particleList, targetParticleList, dirName)
        p = particleList[self.reactantsIndx[0] - 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            for i in range(self.numberCollisions):
                n = float(self.totalAmountCollisions[i])
                vals = (n / float(self.totalNumberCollidableParticles), self.totalEnergyLoss[i] * 0.5 / K.e / np.float64(n),
                        self.totalIncidentEnergy[i] * p.mass * 0.5 / K.e / np.float64(n), self.totalEnergyLoss[i] * 0.5 * p.w_p / timeInterval,
                        n * p.w_p / timeInterval)
                with open(os.path.join(d, "CollisionDiag_%d.dat" % (i + 1)), "a") as f:
                    f.write(" ".join("%16.8E" % v for v in vals) + "\n")

    def writeAveCollisionDiag(self, particleList, targetParticleList, dirName, timeInterval):
        """AveCollisionDiag.dat of solveSimulationFinalAverage (mod_simulation.f90:401-412): one line per collision over the averaging window,
        format (I3, 1x), 5(es16.8, 1x)"""
        d = self._dir(particleList, targetParticleList, dirName)
        os.makedirs(d, exist_ok=True)
        p = particleList[self.reactantsIndx[0] - 1]
        with np.errstate(divide="ignore", invalid="ignore"), open(os.path.join(d, "AveCollisionDiag.dat"), "w") as f:
            f.write("Coll #, CollRatio, AveEnergyLoss (eV), AveIncidentEnergy (eV), P_loss(W/m^2), aveCollFreq (Hz/m^2)\n")
            for i in range(self.numberCollisions):
                n = np.float64(self.totalAmountCollisions[i])
                vals = (n / np.float64(self.totalNumberCollidableParticles), self.totalEnergyLoss[i] * 0.5 / K.e / n,
                        self.totalIncidentEnergy[i] * p.mass * 0.5 / K.e / n, self.totalEnergyLoss[i] * 0.5 * p.w_p / timeInterval,
                        n * p.w_p / timeInterval)
                f.write("%3d " % (i + 1) + " ".join("%16.8E" % v for v in vals) + " \n")

    def reset_totals(self):
        _lib.check(_lib.load().ipic_reset_collision_stats(self._solver._h, self._set))
        self._refresh()


def readNullCollisionInputs(filename, particleList, targetParticleList, collision_dir=None, input_dir=None, seed=0x1D1C0111):
    """readNullCollisionInputs (mod_NullCollision.f90:666-939).  `filename` is the collision.inp-style list of table names (looked
    up in `input_dir`, the reference's ../InputData) or directly a list of table file names; tables are looked up in
    `collision_dir` (the reference's ../../CollisionData).  Returns nullCollisionList: one nullCollision per (charged, neutral)
    reactant pair, its collisions interpolated onto the merged energy axis of the pair."""
    if isinstance(filename, (list, tuple)):
        names = list(filename)
    else:
        with open(os.path.join(input_dir or ".", filename)) as f:
            names = []
            for ln in f:
                tok = ln.split()
                if not tok:
                    continue
                if tok[0][:3].upper() == "END":
                    break
                names.append(tok[0])
    charged = {p.name: k for k, p in enumerate(particleList, 1)}
    neutral = {t.name: k for k, t in enumerate(targetParticleList, 1)}
    reactions = []
    for nm in names:
        path = nm if os.path.isabs(nm) else os.path.join(collision_dir or ".", nm)
        if not os.path.exists(path):
            continue                                          # the reference skips table files it cannot find (:694-695)
        reactions += parse_collision_table(path)
    groups = {}                                               # (charged #, neutral #) -> reactions, in first-seen order
    for r in reactions:
        for nm in r["reactants"] + r["products"]:
            if nm not in charged and nm not in neutral:
                raise ValueError("Could not find particle for collision: " + nm)
        # a name that is both a charged species and a neutral resolves to the neutral (the second loop of :725-738 runs last)
        def num(nm):
            return neutral[nm] if nm in neutral else charged[nm]
        key = tuple(num(nm) for nm in r["reactants"][:2])
        r["productsIndx"] = [num(nm) for nm in r["products"]] + [0] * (3 - len(r["products"]))
        groups.setdefault(key, []).append(r)
    out = []
    for (ip, it), rs in groups.items():
        m1, m2 = particleList[ip - 1].mass, targetParticleList[it - 1].mass
        red_mass, sumMass = m1 * m2 / (m1 + m2), m1 + m2
        E = np.unique(np.concatenate([r["E"] for r in rs]))
        sigma = np.stack([_interp_reference(E, r["E"], r["sigma"], r["threshold"]) for r in rs], axis=1)
        triple = 0.0
        for r in rs:
            if r["type"] == 2:
                a, b, c3 = (particleList[k - 1].mass for k in r["productsIndx"])
                triple = 1.0 / (1.0 / a + 1.0 / b + 1.0 / c3)
        out.append(nullCollision(2, len(rs), E.size, red_mass, sumMass, triple, E, sigma, [r["threshold"] for r in rs],
                                 [r["type"] for r in rs], [ip, it], [len(r["products"]) for r in rs],
                                 np.array([r["productsIndx"] for r in rs], dtype=np.int32).T, seed=seed + 7919 * len(out)))
    return out
