// Host-side inexact Newton / GMRES / backtracking driver: the part of the reference's modernised NITSOL
// that its JFNK path actually executes, restated (never copied) from
//   OpenMP/SharedModules/src/mod_nitsol.f90
//     initializeNitsol :110-255   nitfd :257-514   nitjv :516-722   nitgm :725-1224
//     nitbt :2108-2304            nitdrv :2305-2963                 nitsol :2964-3443
// with the settings initializeNitsol hard-wires (:128-140): GMRES, no preconditioner, first-order forward
// difference J*v, at most 10 backtracks, forcing term choice 2 (gamma 0.9, alpha 1.5, eta_max 0.5).
// BiCGSTAB / TFQMR (input(3) = 1, 2) are never selected by the reference and are not restated.
//
// Third-party arithmetic the reference takes from MKL: dlamch('e'), dlamch('s') and dlaic1 (incremental
// condition estimation, LAPACK 3.x; source in the reference tree at
// Serial/TestNitsol/nitsol-master/Lapack/dlaic1.f).  dlaic1 only decides GMRES breakdown
// (big >= small * cndmax); it never touches the iterates.
//
// The Krylov vectors stay on the host: n <= 8192 and every GMRES decision (Gram-Schmidt coefficients,
// breakdown test, residual estimate) is host control flow, so device-resident vectors would cost one
// synchronising D2H per dot product (~kd per inner iteration) against two 32 KB copies per residual call here.
#pragma once
#include <cmath>
#include <cfloat>
#include <functional>
#include <vector>
#include <algorithm>

namespace ipic {
namespace nitsol {

struct Info { int nfe = 0, njve = 0, nrpre = 0, nli = 0, nni = 0, nbt = 0; };

// f(n, x, fx) -> itrmf (0 = ok)
using Residual = std::function<int(int, const double *, double *)>;

struct Params {                 // initializeNitsol(maxIter, m_Anderson, n)
    int nnimax = 200;           // input(1)
    int kdmax = 20;             // input(4)
    int iksmax = 1000;          // input(6)
    int ibtmax = 10;            // input(9)
    int ieta = 2;               // input(10)
    double etamax = 0.5, choice2_exp = 1.5, choice2_coef = 0.9;
    double eta_cutoff = 0.1, thmin = 0.1, thmax = 0.5;
    double epsmach = 2.0 * (DBL_EPSILON * 0.5);        // 2 * dlamch('e')
    double cndmax = 0.0;
    void init(int maxIter, int m_Anderson) {
        nnimax = maxIter > 0 ? maxIter : 200;
        kdmax = m_Anderson > 0 ? m_Anderson : 20;
        iksmax = m_Anderson > 0 ? 2 * m_Anderson : 1000;
        cndmax = 1.0 / 100.0 / epsmach;                // :116
    }
};

inline double nrm2(int n, const double *x) {           // SQRT(SUM(x**2)), not dnrm2 (as the reference's array syntax)
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * x[i];
    return std::sqrt(s);
}
inline double dot(int n, const double *x, const double *y) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
inline double sign1(double b) { return std::signbit(b) ? -1.0 : 1.0; }

// LAPACK dlaic1: one step of incremental condition estimation.
// job 1: largest singular value, job 2: smallest.  x, w have j entries.
inline void dlaic1(int job, int j, const double *x, double sest, const double *w, double gamma, double &sestpr, double &s, double &c) {
    const double eps = DBL_EPSILON * 0.5;
    const double alpha = dot(j, x, w);
    const double absalp = std::fabs(alpha), absgam = std::fabs(gamma), absest = std::fabs(sest);
    if (job == 1) {
        if (sest == 0.0) {
            const double s1 = std::max(absgam, absalp);
            if (s1 == 0.0) { s = 0.0; c = 1.0; sestpr = 0.0; }
            else {
                s = alpha / s1; c = gamma / s1;
                const double tmp = std::sqrt(s * s + c * c);
                s /= tmp; c /= tmp; sestpr = s1 * tmp;
            }
        } else if (absgam <= eps * absest) {
            s = 1.0; c = 0.0;
            const double tmp = std::max(absest, absalp);
            const double s1 = absest / tmp, s2 = absalp / tmp;
            sestpr = tmp * std::sqrt(s1 * s1 + s2 * s2);
        } else if (absalp <= eps * absest) {
            if (absgam <= absest) { s = 1.0; c = 0.0; sestpr = absest; }
            else { s = 0.0; c = 1.0; sestpr = absgam; }
        } else if (absest <= eps * absalp || absest <= eps * absgam) {
            const double s1 = absgam, s2 = absalp;
            if (s1 <= s2) {
                const double tmp = s1 / s2;
                s = std::sqrt(1.0 + tmp * tmp);
                sestpr = s2 * s;
                c = (gamma / s2) / s;
                s = sign1(alpha) / s;
            } else {
                const double tmp = s2 / s1;
                c = std::sqrt(1.0 + tmp * tmp);
                sestpr = s1 * c;
                s = (alpha / s1) / c;
                c = sign1(gamma) / c;
            }
        } else {
            const double zeta1 = alpha / absest, zeta2 = gamma / absest;
            const double b = (1.0 - zeta1 * zeta1 - zeta2 * zeta2) * 0.5;
            const double cc = zeta1 * zeta1;
            const double t = (b > 0.0) ? cc / (b + std::sqrt(b * b + cc)) : std::sqrt(b * b + cc) - b;
            const double sine = -zeta1 / t, cosine = -zeta2 / (1.0 + t);
            const double tmp = std::sqrt(sine * sine + cosine * cosine);
            s = sine / tmp; c = cosine / tmp;
            sestpr = std::sqrt(t + 1.0) * absest;
        }
    } else {
        if (sest == 0.0) {
            sestpr = 0.0;
            double sine, cosine;
            if (std::max(absgam, absalp) == 0.0) { sine = 1.0; cosine = 0.0; }
            else { sine = -gamma; cosine = alpha; }
            const double s1 = std::max(std::fabs(sine), std::fabs(cosine));
            s = sine / s1; c = cosine / s1;
            const double tmp = std::sqrt(s * s + c * c);
            s /= tmp; c /= tmp;
        } else if (absgam <= eps * absest) {
            s = 0.0; c = 1.0; sestpr = absgam;
        } else if (absalp <= eps * absest) {
            if (absgam <= absest) { s = 0.0; c = 1.0; sestpr = absgam; }
            else { s = 1.0; c = 0.0; sestpr = absest; }
        } else if (absest <= eps * absalp || absest <= eps * absgam) {
            const double s1 = absgam, s2 = absalp;
            if (s1 <= s2) {
                const double tmp = s1 / s2;
                c = std::sqrt(1.0 + tmp * tmp);
                sestpr = absest * (tmp / c);
                s = -(gamma / s2) / c;
                c = sign1(alpha) / c;
            } else {
                const double tmp = s2 / s1;
                s = std::sqrt(1.0 + tmp * tmp);
                sestpr = absest / s;
                c = (alpha / s1) / s;
                s = -sign1(gamma) / s;
            }
        } else {
            const double zeta1 = alpha / absest, zeta2 = gamma / absest;
            const double norma = std::max(1.0 + zeta1 * zeta1 + std::fabs(zeta1 * zeta2), std::fabs(zeta1 * zeta2) + zeta2 * zeta2);
            const double test = 1.0 + 2.0 * (zeta1 - zeta2) * (zeta1 + zeta2);
            double sine, cosine;
            if (test >= 0.0) {
                const double b = (zeta1 * zeta1 + zeta2 * zeta2 + 1.0) * 0.5;
                const double cc = zeta2 * zeta2;
                const double t = cc / (b + std::sqrt(std::fabs(b * b - cc)));
                sine = zeta1 / (1.0 - t);
                cosine = -zeta2 / t;
                sestpr = std::sqrt(t + 4.0 * eps * eps * norma) * absest;
            } else {
                const double b = (zeta2 * zeta2 + zeta1 * zeta1 - 1.0) * 0.5;
                const double cc = zeta1 * zeta1;
                const double t = (b >= 0.0) ? -cc / (b + std::sqrt(b * b + cc)) : b - std::sqrt(b * b + cc);
                sine = -zeta1 / t;
                cosine = -zeta2 / (1.0 + t);
                sestpr = std::sqrt(1.0 + t + 4.0 * eps * eps * norma) * absest;
            }
            const double tmp = std::sqrt(sine * sine + cosine * cosine);
            s = sine / tmp; c = cosine / tmp;
        }
    }
}

class Solver {
public:
    Params P;
    int krystat = 0;
    double avrate = 1.0, fcurnrm = 0.0;

    void initialize(int maxIter, int m_Anderson, int n) {
        P.init(maxIter, m_Anderson);
        n_ = n;
        const int kd = P.kdmax;
        vv.assign((size_t)n * (kd + 1), 0.0);
        rr.assign((size_t)kd * kd, 0.0);
        svbig.assign(kd, 0.0); svsml.assign(kd, 0.0); w.assign(kd, 0.0);
        fcur.assign(n, 0.0); xpls.assign(n, 0.0); fpls.assign(n, 0.0); step.assign(n, 0.0); rwork.assign(n, 0.0);
    }

    // nitsol + nitdrv: returns iterm
    int solve(int n, double *xcur, const Residual &f, double abstol, double reltol, double stptol, Info &info) {
        if (n != n_) initialize(P.nnimax, P.kdmax, n);
        info = Info();
        const double alpha = P.choice2_exp, gamma = P.choice2_coef;
        double eta = 0.5, etamin, fcnrm, fpnrm = 0.0, oftjs, oftlm, redfac = 1.0, rsnrm = 0.0, stpnrm = 0.0;
        int itrmks = 0, iterm = 0;
        avrate = 1.0;
        if (f(n, xcur, fcur.data()) != 0) return 2;
        info.nfe += 1;
        fcnrm = nrm2(n, fcur.data());
        const double ftol = abstol + reltol * fcnrm;
        for (;;) {
            if (fcnrm <= ftol) { iterm = 0; break; }
            if (info.nni > 0 && stpnrm <= stptol && itrmks == 0) { iterm = 0; break; }
            if (info.nni >= P.nnimax) { iterm = 1; break; }
            fcurnrm = fcnrm;
            itrmks = gmres(n, xcur, fcnrm, eta, f, info, rsnrm);
            krystat = itrmks;
            avrate = std::pow(rsnrm / fcnrm, 1.0 / (double)info.nli);
            if (itrmks == 1 || itrmks == 2) { iterm = itrmks + 2; break; }
            if (itrmks >= 3) {
                if (rsnrm / fcnrm >= 1.0) { iterm = 5; break; }
                const double temp = std::log(ftol / fcnrm) / std::log(rsnrm / ((1.0 + 10.0 * P.epsmach) * fcnrm));
                if (temp > 1000.0 * (double)(P.nnimax - info.nni)) { iterm = 5; break; }
            }
            // the first n entries of the reference's work array are vv(:,1): the final linear residual (:2826)
            oftlm = -dot(n, fcur.data(), vv.data());
            oftjs = oftlm - fcnrm * fcnrm;
            int ibt = 0;
            const int itrmbt = backtrack(n, xcur, fcnrm, eta, fpnrm, oftjs, redfac, ibt, f, info);
            if (itrmbt == 1) { iterm = 6; break; }
            if (itrmbt == 2) { iterm = 2; break; }
            info.nbt += ibt;
            // forcing term, choice 2 (:2861-2873)
            etamin = gamma * std::pow(eta, alpha);
            eta = gamma * std::pow(fpnrm / fcnrm, alpha);
            if (etamin <= P.eta_cutoff) etamin = 0.0;
            if (eta < etamin) eta = etamin;
            if (eta > P.etamax) eta = P.etamax;
            if (eta * fpnrm <= 2.0 * ftol) eta = (0.8 * ftol) / fpnrm;
            std::copy(xpls.begin(), xpls.end(), xcur);
            fcur = fpls;
            fcnrm = fpnrm;
            stpnrm = nrm2(n, step.data());
            info.nni += 1;
        }
        return iterm;
    }

private:
    int n_ = 0;
    std::vector<double> vv, rr, svbig, svsml, w, fcur, xpls, fpls, step, rwork;
    double *V(int k) { return vv.data() + (size_t)(k - 1) * n_; }            // vv(:,k), 1-based
    double &RR(int i, int j) { return rr[(size_t)(i - 1) + (size_t)P.kdmax * (j - 1)]; }

    // nitjv(itask = 0) + nitfd, first-order branch (:271-286): z = (f(x + eps v) - f(x)) / eps ; v is overwritten
    int jacvec(int n, const double *xcur, const Residual &f, Info &info, const double *v, double *z) {
        std::copy(v, v + n, rwork.data());
        double eps = nrm2(n, rwork.data());
        if (eps == 0.0) { info.njve += 1; return 1; }
        eps = std::sqrt((1.0 + nrm2(n, xcur)) * P.epsmach) / eps;
        for (int i = 0; i < n; ++i) rwork[i] = xcur[i] + eps * rwork[i];
        if (f(n, rwork.data(), z) != 0) { info.njve += 1; return 1; }
        info.nfe += 1;
        for (int i = 0; i < n; ++i) z[i] = (z[i] - fcur[i]) / eps;
        info.njve += 1;
        return 0;
    }

    // nitgm (:725-1224) with irpre = 0, iresup = 0
    int gmres(int n, const double *xcur, double fcnrm, double eta, const Residual &f, Info &info, double &rsnrm) {
        const int kdmax = P.kdmax;
        std::fill(step.begin(), step.end(), 0.0);
        int igm = 0, itrmks = 0, kd = 0, kdp1 = 1;
        double rsnrm0 = fcnrm, big = 0.0, small = 0.0;
        const double abstol = eta * rsnrm0;
        {
            const double temp = -1.0 / fcnrm;
            double *v1 = V(1);
            for (int i = 0; i < n; ++i) v1[i] = temp * fcur[i];
        }
        bool failed = false;
        for (;;) {                                         // label 100: (re)start
            kd = 0;
            rsnrm = 1.0;
            bool solve_now = false;
            while (!solve_now) {                           // label 200
                kd += 1; kdp1 = kd + 1;
                info.nli += 1; igm += 1;
                const int itrmjv = jacvec(n, xcur, f, info, V(kd), V(kdp1));
                if (itrmjv > 0) { itrmks = itrmjv; failed = true; break; }
                double *vn = V(kdp1);
                for (int i = 2; i <= kd; ++i) {
                    const double *vi = V(i);
                    const double h = dot(n, vi, vn);
                    RR(i - 1, kd) = h;
                    for (int k = 0; k < n; ++k) vn[k] -= h * vi[k];
                }
                RR(kd, kd) = nrm2(n, vn);
                if (kd == 1) {
                    big = RR(1, 1); small = big;
                    svbig[0] = 1.0; svsml[0] = 1.0;
                } else {
                    double sestpr, sn, cs;
                    dlaic1(1, kd - 1, svbig.data(), big, &RR(1, kd), RR(kd, kd), sestpr, sn, cs);
                    big = sestpr;
                    for (int i = 0; i < kd - 1; ++i) svbig[i] *= sn;
                    svbig[kd - 1] = cs;
                    dlaic1(2, kd - 1, svsml.data(), small, &RR(1, kd), RR(kd, kd), sestpr, sn, cs);
                    small = sestpr;
                    for (int i = 0; i < kd - 1; ++i) svsml[i] *= sn;
                    svsml[kd - 1] = cs;
                }
                if (big >= small * P.cndmax) {
                    if (kd == 1) { itrmks = 5; failed = true; break; }
                    kdp1 = kd; kd -= 1;
                    double *v1 = V(1); const double *vk = V(kdp1);
                    for (int k = 0; k < n; ++k) v1[k] += w[kd - 1] * vk[k];
                    solve_now = true;
                    break;
                }
                {
                    const double temp = 1.0 / RR(kd, kd);
                    for (int k = 0; k < n; ++k) vn[k] *= temp;
                }
                w[kd - 1] = dot(n, V(1), vn);
                {
                    const double temp = std::max(std::min(w[kd - 1] / rsnrm, 1.0), -1.0);
                    rsnrm = rsnrm * std::sin(std::acos(temp));
                }
                if (rsnrm0 * rsnrm <= abstol || kd == kdmax || igm >= P.iksmax) { solve_now = true; break; }
                double *v1 = V(1);
                for (int k = 0; k < n; ++k) v1[k] -= w[kd - 1] * vn[k];
            }
            if (failed) break;
            // label 300: triangular solve and step update
            for (int i = 0; i < kd; ++i) svbig[i] = w[i];
            for (int i = kd; i >= 1; --i) {
                w[i - 1] = w[i - 1] / RR(i, i);
                for (int k = 1; k <= i - 1; ++k) w[k - 1] -= w[i - 1] * RR(k, i);
            }
            {
                const double *v1 = V(1);
                for (int k = 0; k < n; ++k) rwork[k] = v1[k] * w[0];
            }
            if (kd > 1) {
                for (int i = 2; i <= kd; ++i) w[i - 1] += w[0] * svbig[i - 2];
                for (int i = 2; i <= kd; ++i) {
                    const double *vi = V(i);
                    for (int k = 0; k < n; ++k) rwork[k] += w[i - 1] * vi[k];
                }
            }
            {   // iresup == 0: residual vector updated by linear combination
                double *v1 = V(1); const double *vk = V(kdp1);
                for (int k = 0; k < n; ++k) v1[k] -= svbig[kd - 1] * vk[k];
            }
            for (int k = 0; k < n; ++k) step[k] += rsnrm0 * rwork[k];
            if (rsnrm0 * rsnrm <= abstol) { itrmks = 0; break; }
            if (igm >= P.iksmax) { itrmks = 3; break; }
            if (big >= small * P.cndmax) { itrmks = 5; break; }
            {
                const double temp = (double)kd * std::log(abstol / (rsnrm0 * rsnrm)) / std::log(rsnrm / (1.0 + 10.0 * P.epsmach));
                if (temp >= 1000.0 * (double)(P.iksmax - igm)) { itrmks = 4; break; }
            }
            rsnrm0 = rsnrm0 * rsnrm;
            {
                const double temp = 1.0 / rsnrm;
                double *v1 = V(1);
                for (int k = 0; k < n; ++k) v1[k] *= temp;
            }
        }
        // label 900
        if (itrmks != 1 && itrmks != 2) {
            double *v1 = V(1);
            for (int k = 0; k < n; ++k) v1[k] *= rsnrm0;
            rsnrm = rsnrm0 * rsnrm;
        }
        return itrmks;
    }

    // nitbt (:2108-2304)
    int backtrack(int n, const double *xcur, double fcnrm, double &eta, double &fpnrm, double oftjs, double &redfac, int &ibt, const Residual &f,
                  Info &info) {
        const double t = 1.e-4;
        ibt = 0;
        redfac = 1.0;
        for (;;) {
            for (int i = 0; i < n; ++i) xpls[i] = xcur[i] + step[i];
            if (f(n, xpls.data(), fpls.data()) != 0) return 2;
            info.nfe += 1;
            fpnrm = nrm2(n, fpls.data());
            if (fpnrm <= (1.0 - t * (1.0 - eta)) * fcnrm || P.ibtmax == -1) return 0;
            ibt += 1;
            if (ibt > P.ibtmax) return 1;
            double theta = -(oftjs * redfac) / (fpnrm * fpnrm - fcnrm * fcnrm - 2.0 * oftjs * redfac);
            if (theta < P.thmin) theta = P.thmin;
            if (theta > P.thmax) theta = P.thmax;
            for (int i = 0; i < n; ++i) step[i] *= theta;
            eta = 1.0 - theta * (1.0 - eta);
            redfac = theta * redfac;
        }
    }
};

}  // namespace nitsol
}  // namespace ipic
