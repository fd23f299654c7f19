"""NITSOL restatements: the numpy oracle (oracle/nitsol_np.py) against the reference's own known-answer problem, and the
library's C++ driver (csrc/nitsol.hpp through ipic_nitsol_host) against the numpy oracle on identical problems."""
import numpy as np
import pytest
import implicitpic1d_b200 as ip
from oracle import nitsol_np


def poisson62():
    """Serial/TestNitsol/TestNitsol/nitsolTest.f90: 64-node uniform grid on 0.1 m, rho = e*1e12, phi = 0 at both ends;
    unknowns are the 62 interior nodes, F(x) = A x - b with A = tridiag(1,-2,1)."""
    e, eps_0 = 1.602176634e-19, 8.8541878128e-12
    delx = 0.1 / 63.0
    b = np.full(62, -(e * 1.0e12) * delx ** 2 / eps_0)

    def A(x):
        r = -2.0 * x
        r[:-1] += x[1:]
        r[1:] += x[:-1]
        return r
    # direct Thomas solve (the test's reference answer)
    n = 62
    cp, dp = np.zeros(n), np.zeros(n)
    cp[0], dp[0] = 1.0 / -2.0, b[0] / -2.0
    for i in range(1, n):
        m = -2.0 - cp[i - 1]
        cp[i] = 1.0 / m
        dp[i] = (b[i] - dp[i - 1]) / m
    x = np.zeros(n)
    x[-1] = dp[-1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return (lambda v: A(v) - b), x


def bratu(n=100, lam=3.0):
    """1-D Bratu problem u'' + lam*exp(u) = 0, u(0)=u(1)=0 (the model problem of the NITSOL distribution, in 1-D)"""
    h = 1.0 / (n + 1)

    def F(u):
        r = -2.0 * u
        r[:-1] += u[1:]
        r[1:] += u[:-1]
        return r / h ** 2 + lam * np.exp(u)
    return F


def test_numpy_nitsol_solves_reference_test_problem():
    F, direct = poisson62()
    r0 = np.linalg.norm(F(np.zeros(62)))
    nit = nitsol_np.Nitsol(50, 60)                      # input(1) = 50, kdmax = 60 as the reference test
    x, iterm = nit.solve(F, np.zeros(62), 1e-6 * r0, 0.0, 1e-8)
    assert iterm == 0
    assert np.max(np.abs((x - direct) / direct)) < 1e-5  # the test prints this percentage difference; tolerance = its ftol
    # inexact Newton (eta starts at 0.5, choice-2 forcing): a handful of outer steps even on a linear problem, no backtracks
    assert nit.info["nni"] <= 8 and nit.info["nbt"] == 0
    assert nit.info["nfe"] == 1 + nit.info["nli"] + nit.info["nni"]     # one F per J*v plus one per accepted step


def test_numpy_nitsol_bratu_matches_scipy_root():
    from scipy.optimize import fsolve
    F = bratu(60, 3.0)
    nit = nitsol_np.Nitsol(100, 20)
    x, iterm = nit.solve(F, np.zeros(60), 1e-10, 0.0)
    assert iterm == 0
    xs = fsolve(F, np.zeros(60), xtol=1e-13)
    assert np.max(np.abs(x - xs)) < 1e-8


@pytest.mark.parametrize("case", ["poisson", "bratu", "bratu_restart", "stagnate"])
def test_cxx_nitsol_matches_numpy(case):
    ip.load_library()
    if case == "poisson":
        F, _ = poisson62(); x0 = np.zeros(62); args = (1e-10, 0.0, 1e-12, 50, 60)
    elif case == "bratu":
        F = bratu(100, 3.0); x0 = np.zeros(100); args = (1e-9, 0.0, 1e-12, 100, 20)
    elif case == "bratu_restart":                       # kdmax small: GMRES restarts and iksmax terminations
        F = bratu(100, 3.4); x0 = np.zeros(100); args = (1e-8, 1e-9, 1e-12, 100, 4)
    else:                                               # near the Bratu turning point: backtracking, possible failure
        F = bratu(100, 3.6); x0 = np.full(100, 2.0); args = (1e-9, 0.0, 1e-12, 12, 6)
    abstol, reltol, stptol, maxIter, kd = args
    nit = nitsol_np.Nitsol(maxIter, kd)
    xn, iterm_n = nit.solve(F, x0, abstol, reltol, stptol)
    xc, info = ip.nitsol(F, x0, abstol, reltol, stptol, maxIter, kd)
    assert info.iterm == iterm_n
    assert (info.nfe, info.njve, info.nli, info.nni, info.nbt) == tuple(nit.info[k] for k in ("nfe", "njve", "nli", "nni", "nbt"))
    # identical decisions (all counters); iterates agree to the noise of the forward-difference J*v (eps ~ 1e-8 amplifies the
    # 1e-16 summation-order differences between the two implementations) -- far below the solver tolerance
    tol = 1e-10 if case in ("poisson", "bratu") else 1e-6
    assert np.max(np.abs(xc - xn)) <= tol * max(1.0, np.max(np.abs(xn)))
    if info.iterm == 0:
        assert np.linalg.norm(F(xc)) <= abstol + reltol * np.linalg.norm(F(x0))
