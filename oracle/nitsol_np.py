"""numpy restatement of the NITSOL path the reference's JFNK solver executes -- TEST INFRASTRUCTURE ONLY
(same rules as oracle/ipic_oracle.cpp: imported by tests/, smoke() and bench.py's CPU legs, never by the product).

Follows OpenMP/SharedModules/src/mod_nitsol.f90: initializeNitsol :110-255 (settings :128-140), nitfd :257-286 (the
first-order branch, the only one reachable with ijacv = 0), nitjv :516-722, nitgm :725-1224, nitbt :2108-2304,
nitdrv :2305-2963, nitsol :2964-3443.  Written independently of implicitpic1d_b200/csrc/nitsol.hpp (different language,
different structure: explicit Hessenberg columns as numpy arrays, 0-based) so that agreement between the two is evidence.

dlaic1 (LAPACK incremental condition estimation; the reference links MKL's) follows the published LAPACK 3 routine, source
in the reference tree at Serial/TestNitsol/nitsol-master/Lapack/dlaic1.f.

PARITY STATUS: unpinned against the Fortran (no compiler here); pinned on the reference's own NITSOL test problem
(Serial/TestNitsol/TestNitsol/nitsolTest.f90: 62-unknown Poisson against the direct tridiagonal solve) in
tests/test_nitsol.py."""
import math
import numpy as np

EPS = np.finfo(np.float64).eps * 0.5          # dlamch('e')


def _nrm(x):
    return math.sqrt(float(np.sum(x * x)))    # SQRT(SUM(x**2))


def dlaic1(job, x, sest, w, gamma):
    """returns (sestpr, s, c)"""
    alpha = float(np.dot(x, w))
    absalp, absgam, absest = abs(alpha), abs(gamma), abs(sest)
    sgn = lambda v: -1.0 if math.copysign(1.0, v) < 0 else 1.0
    if job == 1:
        if sest == 0.0:
            s1 = max(absgam, absalp)
            if s1 == 0.0:
                return 0.0, 0.0, 1.0
            s, c = alpha / s1, gamma / s1
            tmp = math.sqrt(s * s + c * c)
            return s1 * tmp, s / tmp, c / tmp
        if absgam <= EPS * absest:
            tmp = max(absest, absalp)
            s1, s2 = absest / tmp, absalp / tmp
            return tmp * math.sqrt(s1 * s1 + s2 * s2), 1.0, 0.0
        if absalp <= EPS * absest:
            return (absest, 1.0, 0.0) if absgam <= absest else (absgam, 0.0, 1.0)
        if absest <= EPS * absalp or absest <= EPS * absgam:
            s1, s2 = absgam, absalp
            if s1 <= s2:
                tmp = s1 / s2
                s = math.sqrt(1.0 + tmp * tmp)
                return s2 * s, sgn(alpha) / s, (gamma / s2) / s
            tmp = s2 / s1
            c = math.sqrt(1.0 + tmp * tmp)
            return s1 * c, (alpha / s1) / c, sgn(gamma) / c
        z1, z2 = alpha / absest, gamma / absest
        b = (1.0 - z1 * z1 - z2 * z2) * 0.5
        cc = z1 * z1
        t = cc / (b + math.sqrt(b * b + cc)) if b > 0.0 else math.sqrt(b * b + cc) - b
        sine, cosine = -z1 / t, -z2 / (1.0 + t)
        tmp = math.sqrt(sine * sine + cosine * cosine)
        return math.sqrt(t + 1.0) * absest, sine / tmp, cosine / tmp
    # job 2
    if sest == 0.0:
        if max(absgam, absalp) == 0.0:
            sine, cosine = 1.0, 0.0
        else:
            sine, cosine = -gamma, alpha
        s1 = max(abs(sine), abs(cosine))
        s, c = sine / s1, cosine / s1
        tmp = math.sqrt(s * s + c * c)
        return 0.0, s / tmp, c / tmp
    if absgam <= EPS * absest:
        return absgam, 0.0, 1.0
    if absalp <= EPS * absest:
        return (absgam, 0.0, 1.0) if absgam <= absest else (absest, 1.0, 0.0)
    if absest <= EPS * absalp or absest <= EPS * absgam:
        s1, s2 = absgam, absalp
        if s1 <= s2:
            tmp = s1 / s2
            c = math.sqrt(1.0 + tmp * tmp)
            return absest * (tmp / c), -(gamma / s2) / c, sgn(alpha) / c
        tmp = s2 / s1
        s = math.sqrt(1.0 + tmp * tmp)
        return absest / s, -sgn(gamma) / s, (alpha / s1) / s
    z1, z2 = alpha / absest, gamma / absest
    norma = max(1.0 + z1 * z1 + abs(z1 * z2), abs(z1 * z2) + z2 * z2)
    test = 1.0 + 2.0 * (z1 - z2) * (z1 + z2)
    if test >= 0.0:
        b = (z1 * z1 + z2 * z2 + 1.0) * 0.5
        cc = z2 * z2
        t = cc / (b + math.sqrt(abs(b * b - cc)))
        sine, cosine = z1 / (1.0 - t), -z2 / t
        sestpr = math.sqrt(t + 4.0 * EPS * EPS * norma) * absest
    else:
        b = (z2 * z2 + z1 * z1 - 1.0) * 0.5
        cc = z1 * z1
        t = -cc / (b + math.sqrt(b * b + cc)) if b >= 0.0 else b - math.sqrt(b * b + cc)
        sine, cosine = -z1 / t, -z2 / (1.0 + t)
        sestpr = math.sqrt(1.0 + t + 4.0 * EPS * EPS * norma) * absest
    tmp = math.sqrt(sine * sine + cosine * cosine)
    return sestpr, sine / tmp, cosine / tmp


class Nitsol:
    def __init__(self, maxIter, m_Anderson):
        # initializeNitsol :128-140
        self.nnimax = maxIter if maxIter > 0 else 200
        self.kdmax = m_Anderson if m_Anderson > 0 else 20
        self.iksmax = 2 * m_Anderson if m_Anderson > 0 else 1000
        self.ibtmax = 10
        self.etamax, self.alpha, self.gamma = 0.5, 1.5, 0.9
        self.eta_cutoff, self.thmin, self.thmax = 0.1, 0.1, 0.5
        self.epsmach = 2.0 * EPS
        self.cndmax = 1.0 / 100.0 / self.epsmach
        self.info = dict(nfe=0, njve=0, nrpre=0, nli=0, nni=0, nbt=0)
        self.fnorms = []

    # ---- nitjv(itask=0) -> nitfd first-order branch
    def _jv(self, f, xcur, fcur, v):
        I = self.info
        eps = _nrm(v)
        I["njve"] += 1
        if eps == 0.0:
            return None
        eps = math.sqrt((1.0 + _nrm(xcur)) * self.epsmach) / eps
        z = f(xcur + eps * v)
        I["nfe"] += 1
        return (z - fcur) / eps

    # ---- nitgm with irpre = 0, iresup = 0
    def _gmres(self, f, xcur, fcur, fcnrm, eta):
        I = self.info
        n, kdmax = xcur.size, self.kdmax
        step = np.zeros(n)
        rsnrm0 = fcnrm
        abstol = eta * rsnrm0
        V = [(-1.0 / fcnrm) * fcur]                      # V[0] = vv(:,1): current (normalised) linear residual
        igm = 0
        while True:
            del V[1:]
            R = np.zeros((kdmax, kdmax))
            wv = np.zeros(kdmax)
            svbig, svsml = np.zeros(kdmax), np.zeros(kdmax)
            kd = 0
            rsnrm = 1.0
            big = small = 0.0
            while True:
                kd += 1
                I["nli"] += 1
                igm += 1
                z = self._jv(f, xcur, fcur, V[kd - 1])
                if z is None:
                    return 1, step, V[0], rsnrm
                for i in range(1, kd):
                    h = float(np.dot(V[i], z))
                    R[i - 1, kd - 1] = h
                    z = z - h * V[i]
                R[kd - 1, kd - 1] = _nrm(z)
                if kd == 1:
                    big = small = R[0, 0]
                    svbig[0] = svsml[0] = 1.0
                else:
                    sestpr, sn, cs = dlaic1(1, svbig[:kd - 1], big, R[:kd - 1, kd - 1], R[kd - 1, kd - 1])
                    big = sestpr
                    svbig[:kd - 1] *= sn
                    svbig[kd - 1] = cs
                    sestpr, sn, cs = dlaic1(2, svsml[:kd - 1], small, R[:kd - 1, kd - 1], R[kd - 1, kd - 1])
                    small = sestpr
                    svsml[:kd - 1] *= sn
                    svsml[kd - 1] = cs
                if big >= small * self.cndmax:
                    if kd == 1:
                        V[0] = V[0] * rsnrm0
                        return 5, step, V[0], rsnrm0 * rsnrm
                    V.append(z)                          # keeps indices uniform; dropped below
                    kd -= 1
                    V[0] = V[0] + wv[kd - 1] * V[kd]
                    last = V[kd]
                    break
                z = z * (1.0 / R[kd - 1, kd - 1])
                V.append(z)
                wv[kd - 1] = float(np.dot(V[0], z))
                temp = max(min(wv[kd - 1] / rsnrm, 1.0), -1.0)
                rsnrm = rsnrm * math.sin(math.acos(temp))
                last = z
                if rsnrm0 * rsnrm <= abstol or kd == kdmax or igm >= self.iksmax:
                    break
                V[0] = V[0] - wv[kd - 1] * z
            # back substitution and step update
            saved = wv[:kd].copy()
            y = wv.copy()
            for i in range(kd, 0, -1):
                y[i - 1] = y[i - 1] / R[i - 1, i - 1]
                if i > 1:
                    y[:i - 1] -= y[i - 1] * R[:i - 1, i - 1]
            corr = V[0] * y[0]
            if kd > 1:
                y[1:kd] += y[0] * saved[:kd - 1]
                for i in range(2, kd + 1):
                    corr = corr + y[i - 1] * V[i - 1]
            V[0] = V[0] - saved[kd - 1] * last
            step = step + rsnrm0 * corr
            itrmks = None
            if rsnrm0 * rsnrm <= abstol:
                itrmks = 0
            elif igm >= self.iksmax:
                itrmks = 3
            elif big >= small * self.cndmax:
                itrmks = 5
            else:
                temp = float(kd) * math.log(abstol / (rsnrm0 * rsnrm)) / math.log(rsnrm / (1.0 + 10.0 * self.epsmach))
                if temp >= 1000.0 * float(self.iksmax - igm):
                    itrmks = 4
            if itrmks is not None:
                V[0] = V[0] * rsnrm0
                return itrmks, step, V[0], rsnrm0 * rsnrm
            rsnrm0 = rsnrm0 * rsnrm
            V[0] = V[0] * (1.0 / rsnrm)

    # ---- nitsol / nitdrv
    def solve(self, f, x0, abstol, reltol, stptol=1e-12):
        """f(x) -> F(x) (numpy arrays).  Returns (x, iterm)."""
        I = self.info
        for k in I:
            I[k] = 0
        xcur = np.array(x0, dtype=np.float64)
        eta = 0.5
        fcur = f(xcur)
        I["nfe"] += 1
        fcnrm = _nrm(fcur)
        ftol = abstol + reltol * fcnrm
        stpnrm, itrmks = 0.0, 0
        self.fnorms = [fcnrm]
        while True:
            if fcnrm <= ftol:
                return xcur, 0
            if I["nni"] > 0 and stpnrm <= stptol and itrmks == 0:
                return xcur, 0
            if I["nni"] >= self.nnimax:
                return xcur, 1
            itrmks, step, rlin, rsnrm = self._gmres(f, xcur, fcur, fcnrm, eta)
            if itrmks in (1, 2):
                return xcur, itrmks + 2
            if itrmks >= 3:
                if rsnrm / fcnrm >= 1.0:
                    return xcur, 5
                temp = math.log(ftol / fcnrm) / math.log(rsnrm / ((1.0 + 10.0 * self.epsmach) * fcnrm))
                if temp > 1000.0 * float(self.nnimax - I["nni"]):
                    return xcur, 5
            oftlm = -float(np.dot(fcur, rlin))
            oftjs = oftlm - fcnrm ** 2
            # nitbt
            ibt, redfac = 0, 1.0
            while True:
                xpls = xcur + step
                fpls = f(xpls)
                I["nfe"] += 1
                fpnrm = _nrm(fpls)
                if fpnrm <= (1.0 - 1.e-4 * (1.0 - eta)) * fcnrm:
                    break
                ibt += 1
                if ibt > self.ibtmax:
                    return xcur, 6
                theta = -(oftjs * redfac) / (fpnrm ** 2 - fcnrm ** 2 - 2.0 * oftjs * redfac)
                theta = min(max(theta, self.thmin), self.thmax)
                step = theta * step
                eta = 1.0 - theta * (1.0 - eta)
                redfac = theta * redfac
            I["nbt"] += ibt
            etamin = self.gamma * eta ** self.alpha
            eta = self.gamma * (fpnrm / fcnrm) ** self.alpha
            if etamin <= self.eta_cutoff:
                etamin = 0.0
            eta = max(eta, etamin)
            eta = min(eta, self.etamax)
            if eta * fpnrm <= 2.0 * ftol:
                eta = (0.8 * ftol) / fpnrm
            xcur, fcur, fcnrm = xpls, fpls, fpnrm
            stpnrm = _nrm(step)
            I["nni"] += 1
            self.fnorms.append(fcnrm)
