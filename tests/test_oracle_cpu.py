"""CPU tests: the oracle against analytic known answers and the reference's own acceptance invariants
(main.f90:93-114), the host-side Domain against the oracle's, and the C-ABI export list."""
import os
import re
import subprocess
import numpy as np
import pytest
from oracle import oracle_py as op
import implicitpic1d_b200 as ip
from implicitpic1d_b200.domain import readWorld, Domain
from common import DECKS, build_oracle, deck_del_t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ran2_known_answer():
    # Park-Miller minimal standard: seed 1 -> 16807, and the 10000th value is 1043618065 (Park & Miller 1988)
    import ctypes as C
    L = op.lib()
    s = C.c_int32(1)
    r = L.orc_ran2(C.byref(s))
    assert s.value == 16807 and abs(r - 16807 / 2147483647.0) < 1e-18
    for _ in range(9999):
        L.orc_ran2(C.byref(s))
    assert s.value == 1043618065


def test_pcg_matches_reference_c_function():
    # randPCG (mod_BasicFunctions.f90:76-96) uses the PCG32 XSH-RR step of PIC1D/SharedModules/src/PCG_func.c
    import ctypes as C
    L = op.lib()
    st = C.c_uint64(42)
    state = 42
    for _ in range(100):
        old = state
        state = (old * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        rn = ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF
        sgn = rn - (1 << 32) if rn >= (1 << 31) else rn
        expect = (float(sgn) + 2147483648.0) / 4294967296.0
        got = L.orc_randPCG(C.byref(st))
        assert st.value == state and got == expect


@pytest.mark.parametrize("scheme,gridType", [(0, 0), (0, 1), (0, 2), (0, 3), (0, 5), (1, 0), (1, 1), (1, 2), (1, 5)])
def test_domain_matches_oracle(scheme, gridType):
    Nx, L, dx_min = 65, 0.05, 2e-4
    bcl = 1
    o = op.Oracle(scheme, Nx, L, gridType, dx_min, bcl, 1, False, 2)
    g, dx, bc = o.grid()
    w = Domain(scheme, Nx, bcl, 1).constructGrid(dx_min, L, gridType)
    np.testing.assert_allclose(w.grid, g, rtol=0, atol=1e-15 * L)
    np.testing.assert_allclose(w.dx_dl, dx, rtol=1e-12, atol=1e-18)
    assert np.array_equal(w.boundaryConditions, bc)
    assert abs(w.L_domain - o.L_domain) < 1e-15
    rng = np.random.default_rng(1)
    for x in rng.uniform(w.startX, w.endX, 50):
        l = w.getLFromX(x)
        assert abs(l - o.getLFromX(x)) < 1e-12
        assert abs(w.getXFromL(l) - x) < 1e-12 * L + 1e-18


def test_poisson_analytic_parabola():
    # TestCombination.f90's check: uniform rho between grounded walls -> phi(x) = rho x (L - x) / (2 eps0)
    Nx, L = 129, 0.02
    o = op.Oracle(op.NGP, Nx, L, 0, 0.0, 1, 1, False, 1)
    o.init_solver(0.0, 0.0)
    o.add_species(op.m_e, -op.e, 10, 1)
    g, dx, bc = o.grid()
    rho_phys = 1e-6                                    # C/m^3
    rho_logical = rho_phys * np.r_[0.0, 0.5 * (dx[:-1] + dx[1:]), 0.0]   # the solver's rho is charge per node
    o.set("rho", rho_logical)
    o.solve_tridiag_Poisson(0.0)
    phi = o.get("phi")
    exact = rho_phys * g * (L - g) / (2 * op.eps_0)
    assert np.max(np.abs(phi - exact)) < 1e-10 * np.max(exact)


def test_normal_equation_solves_least_squares():
    rng = np.random.default_rng(0)
    m, n = 200, 6
    A = rng.normal(size=(m, n))
    b = rng.normal(size=m)
    out = np.zeros(n)
    Af = np.asfortranarray(A)
    op.lib().orc_solve_normal_equation(Af.ctypes.data_as(op.C.POINTER(op.C.c_double)), b.ctypes.data_as(op.C.POINTER(op.C.c_double)), m, n,
                                       out.ctypes.data_as(op.C.POINTER(op.C.c_double)))
    ref = np.linalg.lstsq(A, b, rcond=None)[0]
    np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-12)


def test_anderson_loop_on_reference_test_problem():
    """Serial/TestAndersonAcceleration/main.f90: 64-node Poisson (62 unknowns) written as the Jacobi fixed point phi = A phi - b, m = 3, mixing 1,
    relative tolerance 1e-8.  The reference prints the Picard and Anderson iteration counts and phi; the known answer is the direct solve.  The
    oracle's Anderson loop (the one solveDivAmpereAnderson runs, normal equations and all) must reach it, far faster than Picard."""
    n, L = 62, 0.1
    delx = L / 63.0
    rho = op.e * 1.0e12
    b = np.full(n, -rho * delx ** 2 / op.eps_0 / 2.0)
    A = np.zeros((n, n))
    A[np.arange(n - 1), np.arange(1, n)] = 0.5
    A[np.arange(1, n), np.arange(n - 1)] = 0.5
    exact = np.linalg.solve(np.eye(n) - A, -b)
    # Picard, as main.f90:32-39
    phi = np.zeros(n)
    r0 = np.linalg.norm(A @ phi - b - phi)
    picard = None
    for i in range(1, 50001):
        phi = A @ phi - b
        if np.linalg.norm(A @ phi - b - phi) < r0 * 1e-8:
            picard = i
            break
    assert picard is not None and picard > 2000
    calls = [0]
    CB = op.C.CFUNCTYPE(None, op.C.c_int, op.C.POINTER(op.C.c_double), op.C.POINTER(op.C.c_double), op.C.c_void_p)

    def G(nn, px, pg, user):
        x = np.ctypeslib.as_array(px, shape=(nn,))
        g = np.ctypeslib.as_array(pg, shape=(nn,))
        g[:] = A @ x - b
        calls[0] += 1
    cb = CB(G)
    L_ = op.lib()
    L_.orc_anderson_fixed_point.restype = op.C.c_int
    for periodic_tree in (0, 1):
        x = np.zeros(n)
        hist = np.zeros(4096)
        nh = op.C.c_int()
        calls[0] = 0
        it = L_.orc_anderson_fixed_point(cb, None, n, x.ctypes.data_as(op.C.POINTER(op.C.c_double)), 3, op.C.c_double(1.0), op.C.c_double(0.0),
                                         op.C.c_double(1e-8), 10000, periodic_tree, hist.ctypes.data_as(op.C.POINTER(op.C.c_double)), 4096, op.C.byref(nh))
        assert 0 < it < 10000 and it < picard // 4, (it, picard)
        # iterNumPicard is i + 1 in the bounded tree (mod_nonLinSolvers.f90:134), i in the periodic tree (:121); G ran 1 + i times
        assert calls[0] == nh.value == it + periodic_tree
        assert hist[nh.value - 1] < 1e-8 * hist[0]
        assert np.max(np.abs(x - exact)) < 1e-6 * np.max(np.abs(exact))


@pytest.mark.parametrize("name", sorted(DECKS))
def test_reference_invariants_one_step(name):
    """main.f90:58-114: energy error ~ solver tolerance, charge/Gauss error ~ round-off after one implicit step."""
    d = DECKS[name]
    o = build_oracle(d)
    del_t = deck_del_t(d)
    rho_i = o.get("rho")
    KE_i = sum(o.getTotalKE(s) for s in range(2))
    remain, curr = o.solvePotential(del_t, del_t, del_t, 0.0)
    PE_i = o.getTotalPE(False)
    KE_f = 0.0
    for s in range(2):
        sc = o.species_scalars(s)
        KE_f += o.getTotalKE(s) + o.lostEnergySum(s) * sc["mass"] * sc["w_p"] * 0.5
    PE_f = o.getTotalPE(True) - o.getEnergyFromBoundary(curr)
    st = o.solver_stats()
    assert st["iterNumPicard"] < d["maxIter"]
    energy_err = abs((PE_i + KE_i - PE_f - KE_f) / (PE_i + KE_i))
    assert energy_err < 1e-6, energy_err
    o.depositRho()
    gauss = o.getError_tridiag_Poisson()
    charge = o.getChargeContinuityError(rho_i, curr)
    assert charge < 1e-6, charge
    assert gauss < 1e-3, gauss
    nh = o.norm_history()
    assert nh[-1] < nh[0]


def test_mover_edge_cases_cpu():
    """absorbed particles still deposit their last partial displacement; Neumann wall reflects; sub-step counts."""
    o = op.Oracle(op.NGP, 11, 1.0, 0, 0.0, 1, 2, False, 1)
    o.init_solver(0.0, 0.0)
    sp = o.add_species(op.m_e, -op.e, 3, 1)
    o.set_weight(sp, 1.0)
    ps = o.phase_space(sp)
    # zero field everywhere is the reference's NaN edge case -> use a tiny uniform field instead
    phi = np.linspace(0.0, 1e-15, 11)
    o.set("phi", phi)
    o.set("phi_f", phi)
    dt = 1e-3
    ps[0, 0] = [1.5, -100.0, 1.0, 2.0]      # hits the Dirichlet wall at l = 1 -> absorbed
    ps[0, 1] = [10.5, 100.0, 0.0, 0.0]      # hits the Neumann wall at l = 11 -> reflected
    ps[0, 2] = [5.25, 1.0, 0.0, 0.0]        # stays in its cell
    o.depositJ(dt)
    J = o.get("J")
    assert abs(J[0] * dt / (-op.e) - (-0.5)) < 1e-9          # absorbed one deposited 1.5 -> 1.0
    o.moveParticles(dt)
    c = o.counters(sp)
    assert c["delIdx"][0] == 1 and c["wallLoss"][0, 0] == 1 and c["refIdx"][0] == 1
    assert abs(c["energyLoss"][0, 0] - (100.0 ** 2 + 1 + 4)) < 1e-3
    assert o.N_p(sp)[0] == 2
    assert ps[0, 0, 1] < 0 and 10.0 < ps[0, 0, 0] < 11.0     # reflected particle moved to slot 1 (stable compaction), v flipped


def test_c_abi_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "ipic1d_b200.h")).read()
    declared = set(re.findall(r"\b(ipic_[a-z_A-Z0-9]+)\s*\(", hdr))
    declared.discard("ipic_ctx")
    L = ip.load_library()
    from implicitpic1d_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", ip.LIB_PATH]).decode()
    exported = set(re.findall(r" T (ipic_[a-zA-Z0-9_]+)", out))
    assert declared <= exported


def test_library_is_sm100a_and_has_tma():
    out = subprocess.run(["cuobjdump", "-lelf", ip.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", ip.LIB_PATH], capture_output=True, text=True).stdout
    push = [blk for blk in sass.split("Function :") if blk.lstrip().startswith("_ZN4ipic6k_push")]
    assert len(push) >= 4              # NGP/CIC x depositJ/moveParticles (x streaming variants)
    deposit = "".join(blk for blk in push if "k_pushILi0ELb0E" in blk or "k_pushILi1ELb0E" in blk)
    assert "REDG.E.ADD.64" in deposit  # int64 fixed-point spill to the global grid: native 64-bit reduction
    assert "DADD" in deposit and "DFMA" in deposit and "MUFU.RSQ64H" in deposit   # fp64 pipe arithmetic


def test_no_gpu_means_loud_failure(gpu_available):
    if gpu_available:
        pytest.skip("GPU present")
    w = readWorld(0, 33, 0.01, 0, 0.0, 1, 1, 0, 2.0, 1e15)
    with pytest.raises(ip.IpicError):
        ip.potentialSolver(w, 0.0, 0.0, 0.0, 0.0)


def test_restart_phase_space_files_round_trip(tmp_path):
    """writePhaseSpace (mod_particle.f90:234-249) / restart read (:371-390): one raw float64 stream per thread slab, N_p from the file length."""
    import implicitpic1d_b200 as ip
    rng = np.random.default_rng(3)
    T, finalIdx = 3, 50
    p = ip.Particle(op.m_e, -op.e, 1.0, 0, finalIdx, "e", T)
    ps = rng.normal(size=(T, finalIdx, 4))
    N_p = np.array([17, 0, 50], dtype=np.int32)
    p.set_phase_space(ps, N_p)
    p.writePhaseSpace(str(tmp_path), download=False)
    for t in range(T):
        fn = tmp_path / "PhaseSpace" / ("phaseSpace_e_thread%d.dat" % (t + 1))
        assert fn.stat().st_size == 32 * N_p[t]
    # Fortran stream order: phaseSpace(1:4, i, t) for i = 1..N_p -> l, v_x, v_y, v_z of particle 1, then particle 2, ...
    first = np.fromfile(tmp_path / "PhaseSpace" / "phaseSpace_e_thread1.dat", dtype="<f8", count=8)
    assert np.array_equal(first, ps[0, :2, :].ravel())
    q = ip.Particle(op.m_e, -op.e, 1.0, 0, finalIdx, "e", T)
    q.readPhaseSpace(str(tmp_path), upload=False)
    assert np.array_equal(q.N_p, N_p)
    for t in range(T):
        assert np.array_equal(q.phaseSpace[t, :N_p[t]], ps[t, :N_p[t]])
    (tmp_path / "PhaseSpace" / "phaseSpace_e_thread2.dat").unlink()
    with pytest.raises(FileNotFoundError):
        q.readPhaseSpace(str(tmp_path), upload=False)


def test_local_temperature_oracle_known_answer():
    """writeLocalTemperature (mod_particle.f90:251-283): per cell SUM(v^2) m / count / 3 / e.  A Maxwellian loaded at T_e gives T_e in every
    cell to sampling noise; a hand-made slab gives the exact per-cell averages."""
    d = DECKS["ngp_rf"]
    o = build_oracle(d)
    Nh = d["Nx"] - 1
    T = o.localTemperature(0, Nh)
    assert T.shape == (Nh,) and np.all(T > 0)
    assert abs(T.mean() - d["T_e"]) < 0.05 * d["T_e"]
    ps = o.phase_space(0)
    n0 = o.N_p(0)
    # exact check against numpy on the same particles
    cells = np.concatenate([ps[t, :n0[t], 0].astype(int) for t in range(d["T"])])
    v2 = np.concatenate([(ps[t, :n0[t], 1:] ** 2).sum(axis=1) for t in range(d["T"])])
    sc = o.species_scalars(0)
    for j in (1, Nh // 2, Nh):
        m = cells == j
        assert abs(T[j - 1] - v2[m].sum() * sc["mass"] / m.sum() / 3.0 / op.e) <= 1e-12 * T[j - 1]
