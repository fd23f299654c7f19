"""Oracle side of the null-collision MCC: a loop-by-loop restatement of readNullCollisionInputs
(OpenMP/SharedModules/src/mod_NullCollision.f90:666-939) and the ctypes glue to the C++ generateCollision restatement
(oracle/ipic_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- never imported by implicitpic1d_b200/ (the product has its own, differently written reader in
implicitpic1d_b200/null_collision.py; tests compare the two).
"""
import ctypes as C
import numpy as np

from . import oracle_py as op

ELASTIC, IONIZATION, EXCITATION, CHARGEEXCHANGE, DISSOCIATION = 1, 2, 3, 4, 5


def _first_token(line):
    """what a list-directed `read(unit,*) string` returns: the first blank/comma delimited token of the record"""
    t = line.replace(",", " ").split()
    return t[0] if t else ""


def index_sort(arr):
    """indexSortArray (mod_BasicFunctions.f90:316-403) through the oracle library; 0-based permutation"""
    a = np.ascontiguousarray(arr, dtype=np.float64)
    out = np.zeros(a.size, dtype=np.int32)
    op.lib().orc_index_sort(a.size, op._dp(a), op._ip(out))
    return out


def read_null_collision_inputs(table_paths, charged, neutrals):
    """readNullCollisionInputs :666-939.
    table_paths: the table files named by collision.inp, in order.  charged: [(name, mass)], neutrals: [(name, mass)].
    Returns one dict per reactant pair (the arguments of nullCollision_constructor, :905-906), species numbers 1-based."""
    numberBinaryCollisions, numberCollisions = 0, 0
    E_threshold, E_temp, sigma_temp, collisionType = [], [], [], []
    numberReactants, numberProducts, reactantsIndx, productsIndx, reactionIndx = [], [], [], [], []
    pairs = []                                               # collisionReactantsIndx(:, 1:numberBinaryCollisions)
    for path in table_paths:
        with open(path) as f:
            lines = [ln.rstrip("\n") for ln in f]
        pos = 0

        def rd():
            nonlocal pos
            while True:
                ln = lines[pos]
                pos += 1
                if ln.strip():                               # list-directed reads skip blank records
                    return ln
        string = _first_token(rd())
        while string[:3] not in ("END", "end"):
            if string[:4] in ("REAC", "Reac", "reac"):
                numberCollisions += 1
                nR, nP, rI, pI = 0, 0, [0, 0], [0, 0, 0]
                string = _first_token(rd())
                if string[:1] != "[":
                    raise ValueError("Do not have chemical reaction after REACTION label")
                lowerIndx = higherIndx = -1
                reactantBool = True
                for i, ch in enumerate(string):
                    if ch == "[":
                        lowerIndx = i
                    if ch == "]":
                        higherIndx = i
                    if ch == ">":
                        reactantBool = False
                    if higherIndx > lowerIndx:
                        name = string[lowerIndx + 1:higherIndx]
                        found, incident, j, k = False, False, 0, 0
                        for j, (nm, _) in enumerate(charged, 1):
                            if nm == name:
                                found, incident = True, True
                                break
                        for k, (nm, _) in enumerate(neutrals, 1):
                            if nm == name:
                                found, incident = True, False
                                break
                        if not found:
                            raise ValueError("Could not find particle for collision: " + name)
                        if reactantBool:
                            rI[nR] = j if incident else k
                            nR += 1
                        else:
                            pI[nP] = j if incident else k
                            nP += 1
                        lowerIndx = higherIndx = -1
                # same reactant pair as an earlier set?
                old = False
                for i, pr in enumerate(pairs, 1):
                    if pr == tuple(rI):
                        reactionIndx.append(i)
                        old = True
                        break
                if not old:
                    pairs.append(tuple(rI))
                    numberBinaryCollisions += 1
                    reactionIndx.append(numberBinaryCollisions)
                numberReactants.append(nR); numberProducts.append(nP); reactantsIndx.append(rI); productsIndx.append(pI)
                t = rd().replace(",", " ").replace("d", "e").replace("D", "E").split()
                E_threshold.append(float(t[0]))
                t = rd().replace(",", " ").replace("d", "e").replace("D", "E").split()
                E_scaling, sigma_scaling = float(t[0]), float(t[1])
                string = _first_token(rd()).strip()
                ctype = 0
                if string in ("ELASTIC", "elastic"): ctype = ELASTIC
                if string in ("IONIZATION", "ionization"): ctype = IONIZATION
                if string in ("EXCITATION", "excitation"): ctype = EXCITATION
                if string in ("CHARGEEXCHANGE", "chargeexchange"): ctype = CHARGEEXCHANGE
                if string in ("DISSOCIATION", "dissociation"): ctype = DISSOCIATION
                collisionType.append(ctype)
                while string[:4] != "----":
                    string = _first_token(rd())
                Es, ss = [], []
                while True:
                    ln = rd()
                    if ln.lstrip()[:4] == "----" or ln[:4] == "----":
                        break
                    t = ln.replace(",", " ").replace("d", "e").replace("D", "E").split()
                    Es.append(float(t[0]) * E_scaling)
                    ss.append(float(t[1]) * sigma_scaling)
                E_temp.append(Es); sigma_temp.append(ss)
            string = _first_token(rd())

    out = []
    for j in range(1, numberBinaryCollisions + 1):
        m1, m2 = charged[pairs[j - 1][0] - 1][1], neutrals[pairs[j - 1][1] - 1][1]
        red_mass = m1 * m2 / (m1 + m2)
        sumMass = (m1 + m2)
        redMassTripleProducts = 0.0
        members = [i for i in range(numberCollisions) if reactionIndx[i] == j]
        E_cat_temp = np.array([x for i in members for x in E_temp[i]], dtype=np.float64)
        indxSort = index_sort(E_cat_temp)
        E_cat = []
        tempVar1 = 1.e10
        for i in range(E_cat_temp.size):
            if E_cat_temp[indxSort[i]] != tempVar1:
                E_cat.append(E_cat_temp[indxSort[i]])
                tempVar1 = E_cat[-1]
        E_cat = np.array(E_cat)
        n = E_cat.size
        sigma_cat = np.zeros((n, len(members)), order="F")
        types, thr, nprod, pidx = [], [], [], []
        for h, i in enumerate(members):
            types.append(collisionType[i]); thr.append(E_threshold[i]); nprod.append(numberProducts[i]); pidx.append(productsIndx[i])
            if collisionType[i] == IONIZATION:
                redMassTripleProducts = 1.0 / (1.0 / charged[productsIndx[i][0] - 1][1] + 1.0 / charged[productsIndx[i][1] - 1][1] +
                                               1.0 / charged[productsIndx[i][2] - 1][1])
            Et, st = E_temp[i], sigma_temp[i]
            lowerIndx = 1                                    # 1-based into Et
            for k in range(n):
                if E_threshold[i] > E_cat[k]:
                    sigma_cat[k, h] = 0.0
                elif E_cat[k] < Et[-1]:
                    while Et[lowerIndx - 1] <= E_cat[k]:
                        lowerIndx += 1
                    lowerIndx -= 1
                    if lowerIndx < 1:
                        raise ValueError("energy below the first table point (the reference would read out of bounds)")
                    if Et[lowerIndx - 1] == E_cat[k]:
                        sigma_cat[k, h] = st[lowerIndx - 1]
                    else:
                        tv1 = E_cat[k] - Et[lowerIndx - 1]
                        tv2 = tv1 / (Et[lowerIndx] - Et[lowerIndx - 1])
                        sigma_cat[k, h] = (st[lowerIndx - 1] * (1 - tv2) + st[lowerIndx] * (tv2))
                else:
                    sigma_cat[k, h] = st[-1]
        out.append(dict(primary=pairs[j - 1][0], target=pairs[j - 1][1], numberCollisions=len(members), lengthArrays=n, red_mass=red_mass,
                        sumMass=sumMass, redMassTripleProducts=redMassTripleProducts, energyArray=E_cat, sigmaArray=sigma_cat,
                        energyThreshold=np.array(thr), collisionType=np.array(types, dtype=np.int32),
                        numberProducts=np.array(nprod, dtype=np.int32), productsIndx=np.array(pidx, dtype=np.int32)))
    return out


_sigs_done = {}


def _sig(L):
    if id(L) in _sigs_done:
        return
    d, i, p, ll = C.c_double, C.c_int, C.c_void_p, C.c_longlong
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "orc_add_target": (i, [p, d, d, d]),
        "orc_add_null_collision": (i, [p, i, i, i, i, dp, dp, dp, ip, ip, ip, d]),
        "orc_generate_collision": (i, [p, i, C.POINTER(C.c_int32), d]),
        "orc_get_collision_scalars": (None, [p, i, dp]),
        "orc_get_collision_tables": (None, [p, i, dp, ip, dp]),
        "orc_get_collision_totals": (None, [p, i, C.POINTER(ll), dp, dp, C.POINTER(ll)]),
        "orc_reset_collision_totals": (None, [p, i]),
        "orc_set_num_to_collide": (None, [p, i]),
        "orc_index_sort": (None, [i, dp, ip]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _sigs_done[id(L)] = True


_sig(op.lib())


class NullCollisionOracle:
    """nullCollisionList(j) of an oracle program state"""

    def __init__(self, o, spec):
        self.o, self.spec = o, spec
        _sig(o.L)
        E = np.ascontiguousarray(spec["energyArray"], dtype=np.float64)
        S = np.asfortranarray(spec["sigmaArray"], dtype=np.float64)
        thr = np.ascontiguousarray(spec["energyThreshold"], dtype=np.float64)
        ct = np.ascontiguousarray(spec["collisionType"], dtype=np.int32)
        npd = np.ascontiguousarray(spec["numberProducts"], dtype=np.int32)
        pi_ = np.ascontiguousarray(spec["productsIndx"], dtype=np.int32)        # [nc][3] C order == (3, nc) Fortran order
        self.nc = int(spec["numberCollisions"])
        self.set = o.L.orc_add_null_collision(o.h, spec["primary"] - 1, spec["target"] - 1, self.nc, int(spec["lengthArrays"]), op._dp(E),
                                              S.ctypes.data_as(C.POINTER(C.c_double)), op._dp(thr), op._ip(ct), op._ip(npd), op._ip(pi_),
                                              float(spec["redMassTripleProducts"]))

    def scalars(self):
        s = np.zeros(7)
        self.o.L.orc_get_collision_scalars(self.o.h, self.set, op._dp(s))
        return dict(sigmaVMax=s[0], reducedMass=s[1], sumMass=s[2], reducedMassIonization=s[3], minEnergy=s[4], maxEnergy=s[5], P_null=s[6])

    def tables(self):
        n = int(self.spec["lengthArrays"])
        thr, ct, S = np.zeros(self.nc), np.zeros(self.nc, dtype=np.int32), np.zeros((n, self.nc), order="F")
        self.o.L.orc_get_collision_tables(self.o.h, self.set, op._dp(thr), op._ip(ct), S.ctypes.data_as(C.POINTER(C.c_double)))
        return dict(energyThreshold=thr, collisionType=ct, sigmaArray=S)

    def generateCollision(self, irand, del_t):
        assert irand.dtype == np.int32 and irand.size == self.o.T
        rc = self.o.L.orc_generate_collision(self.o.h, self.set, irand.ctypes.data_as(C.POINTER(C.c_int32)), del_t)
        if rc:
            raise RuntimeError("oracle: generateCollision would stop (%d)" % rc)

    def totals(self):
        a = np.zeros(self.nc, dtype=np.int64)
        el, ie = np.zeros(self.nc), np.zeros(self.nc)
        tot = C.c_longlong()
        self.o.L.orc_get_collision_totals(self.o.h, self.set, a.ctypes.data_as(C.POINTER(C.c_longlong)), op._dp(el), op._dp(ie), C.byref(tot))
        return dict(totalAmountCollisions=a, totalEnergyLoss=el, totalIncidentEnergy=ie, totalNumberCollidableParticles=tot.value)

    def reset_totals(self):
        self.o.L.orc_reset_collision_totals(self.o.h, self.set)


def add_target(o, mass, density, temperature):
    _sig(o.L)
    return o.L.orc_add_target(o.h, mass, density, temperature)


def set_num_to_collide(o, sp):
    _sig(o.L)
    o.L.orc_set_num_to_collide(o.h, sp)
