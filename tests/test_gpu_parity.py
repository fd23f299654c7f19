"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical particle sets.

Tolerances (from BASELINE.json north_star): J / rho within 1e-12 relative to max|.|; identical nonlinear
iteration counts; energy conservation no worse than the oracle.  Integer quantities are compared exactly."""
import numpy as np
import pytest
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib
from implicitpic1d_b200.particle_mover import depositSubsteps
from common import DECKS, build_oracle, build_cuda_from_oracle, deck_del_t, relerr

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

TOL_GRID = 1e-12      # J, rho: relative to max |.|
ALL = sorted(DECKS)


@pytest.fixture(params=ALL)
def pair(request):
    d = DECKS[request.param]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    yield d, o, world, solver, plist
    solver.close()


def test_deposit_rho_and_initial_potential(pair):
    d, o, world, solver, plist = pair
    assert relerr(solver.rho, o.get("rho")) < TOL_GRID
    # Thomas mode reproduces the reference's operation order; the remaining difference is the 1e-16-level rho difference
    assert relerr(solver.phi, o.get("phi")) < 1e-10
    assert relerr(solver.phi_f, o.get("phi_f")) < 1e-10
    for name in ("a_tri", "b_tri", "c_tri"):
        assert relerr(getattr(solver, name), o.get(name)) < 1e-15


def test_scan_solve_matches_thomas(pair):
    """IPIC_TRIDIAG_SCAN (the default on bounded decks): x = T^-1 d through the generators of the constant matrix's inverse, two block scans.
    Against the Thomas sweep on the same right-hand side: 1e-11 of max|phi| (on a realistic right-hand side the generators are closer to
    the exact solution than the double-precision Thomas sweep is)."""
    d, o, world, solver, plist = pair
    solver.set_tridiag_mode(_lib.TRIDIAG_THOMAS)
    solver.solve_tridiag_Poisson(world, 0.0)
    phi_t = solver.phi.copy()
    try:
        solver.set_tridiag_mode(_lib.TRIDIAG_SCAN)
    except _lib.IpicError:
        assert d["bcL"] == 3 or d["bcR"] == 3 or (d["bcL"] == 2 and d["bcR"] == 2)       # periodic / singular: no generators, PCR stays
        return
    solver.solve_tridiag_Poisson(world, 0.0)
    assert relerr(solver.phi, phi_t) < 1e-11, relerr(solver.phi, phi_t)
    solver.set_tridiag_mode(_lib.TRIDIAG_THOMAS)


def test_pcr_matches_thomas(pair):
    d, o, world, solver, plist = pair
    phi_t = solver.phi
    solver.set_tridiag_mode(_lib.TRIDIAG_PCR)
    solver.solve_tridiag_Poisson(world, 0.0)
    assert relerr(solver.phi, phi_t) < 1e-9
    solver.set_tridiag_mode(_lib.TRIDIAG_THOMAS)


def _sync_fields(o, solver):
    """put the oracle's phi / phi_f on the device so that the particle kernels see identical fields"""
    solver.phi = o.get("phi")
    solver.phi_f = o.get("phi_f")
    solver.rho = o.get("rho")


def test_half_time_efield(pair):
    d, o, world, solver, plist = pair
    _sync_fields(o, solver)
    o.makeHalfTimeEField()
    solver.makeHalfTimeEField()
    assert relerr(solver.EField, o.get("EField")) < 1e-14


def test_depositJ(pair):
    d, o, world, solver, plist = pair
    _sync_fields(o, solver)
    del_t = deck_del_t(d)
    o.depositJ(del_t)
    ip.depositJ(solver, plist, world, del_t)
    Jo, Jg = o.get("J"), solver.J
    assert relerr(Jg, Jo) < TOL_GRID, relerr(Jg, Jo)
    for sp, p in enumerate(plist):
        assert depositSubsteps(solver, p) == o.move_stats(sp)["substepsDeposit"]
    # deterministic: bit-identical on repeat, for any scheduling of the sub-steps (register-loop threshold, tile order,
    # grid size): the deposits are exact integers
    ip.depositJ(solver, plist, world, del_t)
    assert np.array_equal(solver.J, Jg)
    for name, vals, default in (("stay_min", (1, 32), 20), ("tile_order", (0, 2), 1), ("push_blocks", (1, 7), 0), ("transit_min", (1, 33), 12),
                                ("tile_subs", (1, 8), 0), ("entry_min", (1, 33), 8)):
        for v in vals:
            solver.set_option(name, v)
            ip.depositJ(solver, plist, world, del_t)
            assert np.array_equal(solver.J, Jg), (name, v)
        solver.set_option(name, default)
    # the entry / exit stages of the fresh path (large-tile kernels of electron-like species): first sub-step to the face ahead or behind and,
    # where the step ends in the cell entered, the last one as well, taken with the whole batch -- the same J bit for bit and the same
    # sub-step counts as the general sub-step (and as the oracle)
    for combo in (dict(tile_subs=8, entry_min=1, exit_stage=1), dict(tile_subs=8, entry_min=1, exit_stage=0), dict(tile_subs=8, entry_min=33, exit_stage=1)):
        for k, v in combo.items():
            solver.set_option(k, v)
        ip.depositJ(solver, plist, world, del_t)
        assert np.array_equal(solver.J, Jg), combo
        for sp, p in enumerate(plist):
            assert depositSubsteps(solver, p) == o.move_stats(sp)["substepsDeposit"], combo
    for k, v in dict(tile_subs=0, entry_min=8, exit_stage=1).items():
        solver.set_option(k, v)


def test_ampere_solve(pair):
    d, o, world, solver, plist = pair
    _sync_fields(o, solver)
    del_t = deck_del_t(d)
    o.makeConstSourceTerm()
    solver.makeConstSourceTerm()
    assert relerr(solver.rho, o.get("rho")) < 1e-13
    o.depositJ(del_t)
    o.solve_tridiag_Ampere(del_t)
    solver.J = o.get("J")
    solver.solve_tridiag_Ampere(world, del_t)
    assert relerr(solver.phi_f, o.get("phi_f")) < 1e-13       # same J, same operation order
    solver.set_tridiag_mode(_lib.TRIDIAG_PCR)
    solver.phi_f = o.get("phi")
    solver.solve_tridiag_Ampere(world, del_t)
    assert relerr(solver.phi_f, o.get("phi_f")) < 1e-8
    solver.set_tridiag_mode(_lib.TRIDIAG_THOMAS)


def test_moveParticles(pair):
    d, o, world, solver, plist = pair
    _sync_fields(o, solver)
    del_t = deck_del_t(d)
    o.moveParticles(del_t)
    ip.moveParticles(solver, plist, world, del_t)
    for sp, p in enumerate(plist):
        c = o.counters(sp)
        ms = o.move_stats(sp)
        assert p.delIdx == int(c["delIdx"].sum())
        assert p.refIdx == int(c["refIdx"].sum())
        assert np.array_equal(p.wallLoss, c["wallLoss"].sum(axis=0))
        assert p.numSubSteps == ms["substepsMove"]
        assert p.numFuncEvals == ms["funcEvalsMove"]
        assert p.numSubStepsAve == ms["numSubStepsAve"]
        el = c["energyLoss"].sum(axis=0)
        np.testing.assert_allclose(p.energyLoss, el, rtol=1e-12, atol=0)
        np.testing.assert_allclose(p.momentumLoss, c["momentumLoss"].sum(axis=0), rtol=1e-11, atol=1e-9)
        ps_g = p.download()
        ps_o = o.phase_space(sp)
        assert np.array_equal(p.N_p, o.N_p(sp))            # per-slab survivor counts: stable compaction
        for t in range(d["T"]):
            n = p.N_p[t]
            lo, lg = ps_o[t, :n, 0], ps_g[t, :n, 0]
            assert np.max(np.abs(lg - lo)) < 1e-9, (sp, t, np.max(np.abs(lg - lo)))
            np.testing.assert_allclose(ps_g[t, :n, 1], ps_o[t, :n, 1], rtol=1e-9, atol=1e-6)
            assert np.array_equal(ps_g[t, :n, 2:], ps_o[t, :n, 2:])   # v_y, v_z are carried bit-for-bit
        assert abs(p.getTotalKE() - o.getTotalKE(sp)) < 1e-12 * abs(o.getTotalKE(sp))
        np.testing.assert_allclose(p.getTotalMomentum(), o.getTotalMomentum(sp), rtol=1e-9, atol=1e-30)


def _one_step(d, o, world, solver, plist):
    del_t = deck_del_t(d)
    rho_i_o, rho_i_g = o.get("rho"), solver.rho
    KE_i_o = sum(o.getTotalKE(s) for s in range(2))
    KE_i_g = sum(p.getTotalKE() for p in plist)
    ro, co = o.solvePotential(del_t, del_t, del_t, 0.0)
    rg, cg = ip.solvePotential(solver, plist, world, del_t, del_t, del_t, 0.0)
    so, sg = o.solver_stats(), ip.SolverState(solver)
    # energy bookkeeping of main.f90:75-93
    PE_i_o, PE_i_g = o.getTotalPE(False), solver.getTotalPE(world, False)
    KE_f_o = KE_f_g = 0.0
    for s, p in enumerate(plist):
        sc = o.species_scalars(s)
        KE_f_o += o.getTotalKE(s) + o.lostEnergySum(s) * sc["mass"] * sc["w_p"] * 0.5
        KE_f_g += p.getTotalKE() + p.energyLoss.sum() * p.mass * p.w_p * 0.5
    PE_f_o = o.getTotalPE(True) - o.getEnergyFromBoundary(co)
    PE_f_g = solver.getTotalPE(world, True) - solver.getEnergyFromBoundary(world, cg)
    e_o = abs((PE_i_o + KE_i_o - PE_f_o - KE_f_o) / (PE_i_o + KE_i_o))
    e_g = abs((PE_i_g + KE_i_g - PE_f_g - KE_f_g) / (PE_i_g + KE_i_g))
    return dict(ro=ro, co=co, rg=rg, cg=cg, so=so, sg=sg, e_o=e_o, e_g=e_g, rho_i_o=rho_i_o, rho_i_g=rho_i_g)


def test_solvePotential_iteration_parity(pair):
    """config 2 of BASELINE.json: same nonlinear iteration counts, energy conservation no worse than the reference path."""
    d, o, world, solver, plist = pair
    r = _one_step(d, o, world, solver, plist)
    so, sg = r["so"], r["sg"]
    assert (r["rg"], r["cg"]) == (r["ro"], r["co"])
    assert sg.iterNumPicard == so["iterNumPicard"], (sg.normHistory, o.norm_history())
    assert sg.sweeps == so["sweeps"]
    assert sg.amountTimeSplits == so["amountTimeSplits"]
    assert sg.substeps == so["substeps"]
    nh_o, nh_g = o.norm_history(), sg.normHistory
    assert len(nh_o) == len(nh_g)
    # the residual history agrees until it reaches the round-off floor of the fixed point
    floor = 1e-9 * nh_o[0]
    for a, b in zip(nh_g, nh_o):
        assert abs(a - b) <= 1e-6 * b + floor
    assert relerr(solver.phi_f, o.get("phi_f")) < 1e-9
    assert relerr(solver.J, o.get("J")) < 1e-8
    assert r["e_g"] <= max(2.0 * r["e_o"], 1e-12), (r["e_g"], r["e_o"])
    for sp, p in enumerate(plist):
        assert p.count() == int(o.N_p(sp).sum())
    # charge continuity and Gauss law after the step (main.f90:102-114)
    o.depositRho()
    solver.depositRho(plist, world)
    assert relerr(solver.rho, o.get("rho")) < 1e-11
    ce_o = o.getChargeContinuityError(r["rho_i_o"], r["co"])
    ce_g = solver.getChargeContinuityError(r["rho_i_g"], world, r["cg"])
    assert ce_g < max(10 * ce_o, 1e-9), (ce_g, ce_o)
    # Gauss residual: an rms of 1 - Ax/d that is dominated by nodes with tiny rho, so only its magnitude is comparable
    assert abs(solver.getError_tridiag_Poisson(world) - o.getError_tridiag_Poisson()) <= 1e-2 * o.getError_tridiag_Poisson() + 1e-12
    assert abs(solver.getJdotE(world, r["cg"]) - o.getJdotE(r["co"])) <= 1e-8 * abs(o.getJdotE(r["co"])) + 1e-30


@pytest.mark.parametrize("mode", [_lib.TRIDIAG_PCR, _lib.TRIDIAG_SCAN])
@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf"])
def test_solvePotential_pcr_mode(name, mode):
    """production modes (scan solve through the inverse's generators, the default; PCR): same iteration count as the oracle on the default decks"""
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=mode)
    try:
        r = _one_step(d, o, world, solver, plist)
        assert r["sg"].iterNumPicard == r["so"]["iterNumPicard"]
        assert r["e_g"] <= max(2.0 * r["e_o"], 1e-12)
        assert relerr(solver.phi_f, o.get("phi_f")) < 1e-7
    finally:
        solver.close()


@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf"])
def test_funcNitsol_residual(name):
    """JFNK residual callback F(x) = x - G(x) (mod_nonLinSolvers.f90:218-239) against the oracle's pieces"""
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    try:
        del_t = deck_del_t(d)
        _sync_fields(o, solver)
        o.makeConstSourceTerm()
        solver.makeConstSourceTerm()
        x = o.get("phi_f") * (1.0 + 1e-3 * np.sin(np.arange(d["Nx"])))
        o.set("phi_f", x)
        o.depositJ(del_t)
        o.solve_tridiag_Ampere(del_t)
        f_o = x - o.get("phi_f")
        f_g = ip.funcNitsol(solver, plist, x, del_t)
        assert relerr(f_g, f_o) < 1e-9
    finally:
        solver.close()


def test_sort_keeps_J_bitwise_and_sorts():
    d = DECKS["ngp_rf"]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o)
    try:
        del_t = deck_del_t(d)
        ip.depositJ(solver, plist, world, del_t)
        J0 = solver.J
        L = ip.load_library()
        for p in plist:
            _lib.check(L.ipic_sort_particles(solver._h, p._index))
        ip.depositJ(solver, plist, world, del_t)
        assert np.array_equal(solver.J, J0)
        ps = plist[0].download()
        l = np.concatenate([ps[t, :plist[0].N_p[t], 0] for t in range(d["T"])])
        assert np.all(np.diff(np.floor(l)) >= 0)
        assert l.size == int(o.N_p(0).sum())
    finally:
        solver.close()


def test_large_load_properties():
    """size-independent properties at a device-generated load: charge continuity, energy conservation, determinism"""
    import ctypes as C
    from implicitpic1d_b200.domain import readWorld
    from implicitpic1d_b200 import constants as K
    n_ave, T_e, Nx, L_dom = 1e16, 4.0, 1025, 0.05
    world = readWorld(0, Nx, L_dom, 0, 0.0, 1, 1, 1, T_e, n_ave)
    solver = ip.potentialSolver(world, 0.0, 0.0, 0.0, 0.0)
    try:
        N = 2_000_000
        L = ip.load_library()
        m_ion = 4.002602 * K.m_amu - K.m_e
        e = ip.Particle(K.m_e, -K.e, n_ave * world.L_domain / N, 0, 0, "e", 1)
        ion = ip.Particle(m_ion, K.e, n_ave * world.L_domain / N, 0, 0, "He+", 1)
        plist = [e, ion]
        solver._bind(plist)
        _lib.check(L.ipic_generate_particles(solver._h, 0, N, T_e, 0.0, 1234))
        _lib.check(L.ipic_generate_particles(solver._h, 1, N, 0.025, 0.0, 1234))
        for p in plist:
            _lib.check(L.ipic_sort_particles(solver._h, p._index))
        ip.initializeSolver(solver, 1e-8, 0.0, 0, 10, 0.5, 100)
        solver.solveInitialPotential(plist, world, 0.0)
        rho_i = solver.rho
        del_t = K.del_t_from_deck(n_ave, 1.0, 1.0)
        KE_i = sum(p.getTotalKE() for p in plist)
        r, c = ip.solvePotential(solver, plist, world, del_t, del_t, del_t, 0.0)
        st = ip.SolverState(solver)
        assert st.iterNumPicard < 100
        PE_i = solver.getTotalPE(world, False)
        KE_f = sum(p.getTotalKE() + p.energyLoss.sum() * p.mass * p.w_p * 0.5 for p in plist)
        PE_f = solver.getTotalPE(world, True) - solver.getEnergyFromBoundary(world, c)
        assert abs((PE_i + KE_i - PE_f - KE_f) / (PE_i + KE_i)) < 1e-9
        solver.depositRho(plist, world)
        assert solver.getChargeContinuityError(rho_i, world, c) < 1e-6
        assert sum(p.count() + p.delIdx for p in plist) == 2 * N
    finally:
        solver.close()


FULL_SIZE = {
    # BASELINE configs[1]: default CIC deck scaled to 1e7 particles (OpenMP/CIC/InputData)
    "c2_cic_1e7": dict(scheme=1, Nx=256, L=0.067, n_ave=5.12e14, T_e=2.59, T_i=0.025, bcL=1, bcR=4, Vl=0.0, Vr=200.0, rf=13.56e6, smooth=1, frac=1.5,
                       dt_deck=9.218289085545723e-11, eps_a=1e-6, m=10, N=5_000_000, e_tol=1e-7),
    # BASELINE configs[4] per-GPU shard: default NGP deck, 4096 cells, 1.25e8 particles
    "c5_ngp_shard": dict(scheme=0, Nx=4097, L=0.067 * 4096 / 1000.0, n_ave=1.6e18, T_e=4.0, T_i=0.025, bcL=1, bcR=4, Vl=0.0, Vr=2000.0, rf=200.0e6, smooth=1,
                         frac=10.0, dt_deck=1.0, eps_a=1e-8, m=10, N=62_500_000, e_tol=1e-7),
}


@pytest.mark.parametrize("name", sorted(FULL_SIZE))
def test_full_size_properties(name):
    """BASELINE.json's own sizes, where the oracle is too slow to be the checker: the size-independent properties the scheme offers
    (main.f90:93-114) -- total energy conserved to the nonlinear tolerance, charge continuity to round-off, every particle accounted
    for -- plus determinism: J is bit-identical from sweep to sweep and for a different tile schedule / CTA count."""
    import ctypes as C
    from implicitpic1d_b200.domain import readWorld
    from implicitpic1d_b200 import constants as K
    w = FULL_SIZE[name]
    world = readWorld(w["scheme"], w["Nx"], w["L"], 0, 0.0, w["bcL"], w["bcR"], w["smooth"], w["T_e"], w["n_ave"])
    solver = ip.potentialSolver(world, w["Vl"], w["Vr"], w["rf"], 0.0)
    try:
        N = w["N"]
        L = ip.load_library()
        m_ion = 4.002602 * K.m_amu - K.m_e
        plist = [ip.Particle(K.m_e, -K.e, w["n_ave"] * world.L_domain / N, 0, 0, "e", 1),
                 ip.Particle(m_ion, K.e, w["n_ave"] * world.L_domain / N, 0, 0, "He+", 1)]
        solver._bind(plist)
        _lib.check(L.ipic_generate_particles(solver._h, 0, N, w["T_e"], 0.0, 2026))
        _lib.check(L.ipic_generate_particles(solver._h, 1, N, w["T_i"], 0.0, 2026))
        for p in plist:
            _lib.check(L.ipic_sort_particles(solver._h, p._index))
        ip.initializeSolver(solver, w["eps_a"], 0.0, 0, w["m"], 0.5, 100)
        solver.solveInitialPotential(plist, world, 0.0)
        del_t = K.del_t_from_deck(w["n_ave"], w["frac"], w["dt_deck"])
        # determinism of the deposition
        ip.depositJ(solver, plist, world, del_t)
        J0 = solver.J.copy()
        ip.depositJ(solver, plist, world, del_t)
        assert np.array_equal(J0, solver.J)
        solver.set_option("tile_order", 0)
        solver.set_option("push_blocks", 97)
        ip.depositJ(solver, plist, world, del_t)
        assert np.array_equal(J0, solver.J)
        solver.set_option("tile_order", 1)
        solver.set_option("push_blocks", 0)
        assert np.all(np.isfinite(J0)) and np.max(np.abs(J0)) > 0
        # two implicit steps: conservation laws
        tcur, remain, curr = 0.0, del_t, del_t
        for step in range(2):
            rho_i = solver.rho if step == 0 else rho_f
            KE_i = sum(p.getTotalKE() for p in plist)
            PE_i = solver.getTotalPE(world, False) if step == 0 else solver.getTotalPE(world, True)
            n_before = sum(p.count() for p in plist)
            remain, curr = ip.solvePotential(solver, plist, world, del_t, remain, curr, tcur)
            tcur += curr
            st = ip.SolverState(solver)
            assert st.iterNumPicard < 100
            KE_f = sum(p.getTotalKE() + p.energyLoss.sum() * p.mass * p.w_p * 0.5 for p in plist)
            PE_f = solver.getTotalPE(world, True) - solver.getEnergyFromBoundary(world, curr)
            assert abs((PE_i + KE_i - PE_f - KE_f) / (PE_i + KE_i)) < w["e_tol"], (step, PE_i, KE_i, PE_f, KE_f)
            solver.depositRho(plist, world)
            rho_f = solver.rho
            assert solver.getChargeContinuityError(rho_i, world, curr) < 1e-6
            assert sum(p.count() + p.delIdx for p in plist) == n_before
    finally:
        solver.close()


def test_two_gpu_sharded_step_is_bitwise_single_gpu():
    """needs >= 2 GPUs on the box: launches tests/multi_gpu_check.py under torchrun"""
    import ctypes as C, os, subprocess, sys
    n = C.c_int(0)
    _lib.check(_lib.load().ipic_device_count(C.byref(n)))
    if n.value < 2:
        pytest.skip("one GPU on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=280)
    assert r.returncode == 0 and "MULTI-GPU CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf", "ngp_neumann", "cic_periodic"])
def test_jfnk_step_matches_oracle(name):
    """solverType 1: the library's NITSOL on the device residual against the numpy NITSOL on the oracle residual
    (mod_nonLinSolvers.f90:218-282).  Newton / GMRES / backtrack counters are compared exactly: the two residuals differ
    by ~1e-13 relative (fixed-point deposition), far from any decision threshold on these decks."""
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    ip.initializeSolver(solver, d["eps_a"], 0.0, 1, d["m"], d["beta"], d["maxIter"])
    del_t = deck_del_t(d)
    io = o.solveJFNK(del_t, d["eps_a"], 0.0, d["maxIter"], d["m"])
    ig = ip.solveJFNK(solver, plist, world, del_t)
    assert ig.iterm == io["iterm"] == 0, (ig, io)
    assert (ig.nfe, ig.njve, ig.nli, ig.nni, ig.nbt) == (io["nfe"], io["njve"], io["nli"], io["nni"], io["nbt"]), (ig, io)
    st = ip.SolverState(solver)
    assert st.iterNumPicard == io["nfe"]                       # iterNumPicard = info(1) (:271)
    assert relerr(solver.phi_f, o.get("phi_f")) < 1e-7
    for sp, p in enumerate(plist):
        assert p.N_p_total == int(o.N_p(sp).sum())
    # the same step through solvePotential (adaptive wrapper :284-314) on a fresh pair gives the same counters
    solver.close()


@pytest.mark.parametrize("mult", [4.0, 400.0])
@pytest.mark.parametrize("name", ["ngp_neumann", "cic_neumann"])
def test_refRecordIdx_matches_oracle(name, mult):
    """moveParticles at a Neumann wall records, per thread slab, the post-compaction index of every particle at its first reflection
    (NGP mod_particleMover.f90:226-231, CIC mod_particleMoverPicard.f90:276-281); the other wall absorbs, so the compaction shifts them."""
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    # mult = 4: enough wall traffic for a meaningful list; mult = 400: electrons cross the whole domain within the step, so some are
    # reflected at the Neumann wall AND absorbed at the other one -- their entries name the successor's slot (reference: i - delIdx)
    del_t = deck_del_t(d) * mult
    solver.phi, solver.phi_f, solver.rho = o.get("phi"), o.get("phi_f"), o.get("rho")
    o.makeHalfTimeEField(); solver.makeHalfTimeEField()
    for step in range(2 if mult < 100.0 else 1):                  # the long step empties the domain of electrons: one move
        o.moveParticles(del_t)
        ip.moveParticles(solver, plist, world, del_t)
        total = 0
        for sp, p in enumerate(plist):
            p.download()
            cnt = o.counters(sp)
            assert np.array_equal(p.N_p, o.N_p(sp))
            refIdx, rec = p.refRecord()
            assert np.array_equal(refIdx, cnt["refIdx"]), (sp, refIdx, cnt["refIdx"])
            assert p.refIdx == int(cnt["refIdx"].sum())
            cap = p.finalIdx // 10                                # refRecordIdx(finalIdx/10, T): the reference drops what does not fit
            for t in range(d["T"]):
                m = min(int(cnt["refIdx"][t]), cap)
                assert np.array_equal(rec[t][:m], o.ref_record(sp, t + 1, m)), (sp, t)
            total += int(refIdx.sum())
            if sp == 0:
                assert cnt["delIdx"].sum() > 0                    # the compaction did shift the electrons' indices
        assert total > 20
    solver.close()


@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf"])
def test_empty_and_ragged_slabs(name):
    """edge cases of the container (mod_particle.f90:15-37: N_p(T) is per thread slab and may be anything up to finalIdx): an empty slab, a
    one-particle slab, a full slab, and a species with no particles at all -- depositJ and moveParticles against the oracle on the same slabs"""
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    del_t = deck_del_t(d)
    T, ppt = d["T"], d["ppt"]
    ragged = [np.array([ppt, 0, 1, ppt // 3][:T], dtype=np.int32), np.zeros(T, dtype=np.int32)]    # electrons ragged, no ions at all
    for sp, p in enumerate(plist):
        o.set_N_p(sp, ragged[sp])
        p.N_p = ragged[sp].copy()
        p.upload()
        assert p.count() == int(ragged[sp].sum())
    solver.phi, solver.phi_f, solver.rho = o.get("phi"), o.get("phi_f"), o.get("rho")
    o.makeHalfTimeEField(); solver.makeHalfTimeEField()
    o.depositJ(del_t)
    ip.depositJ(solver, plist, world, del_t)
    assert relerr(solver.J, o.get("J")) < TOL_GRID
    assert depositSubsteps(solver, plist[0]) == o.move_stats(0)["substepsDeposit"] and depositSubsteps(solver, plist[1]) == 0
    o.moveParticles(del_t)
    ip.moveParticles(solver, plist, world, del_t)
    for sp, p in enumerate(plist):
        ps_g, ps_o = p.download(), o.phase_space(sp)
        assert np.array_equal(p.N_p, o.N_p(sp))
        assert p.delIdx == int(o.counters(sp)["delIdx"].sum())
        for t in range(T):
            n = p.N_p[t]
            assert np.max(np.abs(ps_g[t, :n, 0] - ps_o[t, :n, 0]), initial=0.0) < 1e-9
            assert np.array_equal(ps_g[t, :n, 2:], ps_o[t, :n, 2:])
    assert plist[1].count() == 0 and plist[1].getTotalKE() == 0.0
    # rho of an empty species is zero; the solver still runs
    solver.depositRho(plist, world)
    o.depositRho()
    assert relerr(solver.rho, o.get("rho")) < TOL_GRID
    solver.close()


def test_cic_walls_pushed_by_the_host():
    """level-1 integration of the CIC tree: the host's setRFVoltage / resetVoltage own boundPhi, RF_voltage and sourceTermVals
    (CIC mod_potentialSolver.f90:458-478); ipic_set_cic_walls carries them to the device before depositJ"""
    import ctypes as C
    d = DECKS["cic_rf"]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    del_t = deck_del_t(d)
    t_future = 0.3 / d["rf"]                                       # well into the RF period: walls differ from their t = 0 values
    o.setRFVoltage(t_future)
    o.resetVoltage()
    o.setRFVoltage(t_future + del_t)
    wl = o.cic_walls()
    assert abs(wl["RF_voltage"]) > 1.0 and abs(wl["boundPhi"]).max() > 1.0
    solver.phi, solver.phi_f, solver.rho = o.get("phi"), o.get("phi_f"), o.get("rho")
    bp, stv = np.ascontiguousarray(wl["boundPhi"]), np.ascontiguousarray(wl["sourceTermVals"])
    _lib.check(_lib.load().ipic_set_cic_walls(solver._h, bp.ctypes.data_as(_lib._dp), float(wl["RF_voltage"]), stv.ctypes.data_as(_lib._dp)))
    o.makeHalfTimeEField(); solver.makeHalfTimeEField()
    assert relerr(solver.EField, o.get("EField")) < 1e-13
    o.depositJ(del_t)
    ip.depositJ(solver, plist, world, del_t)
    assert relerr(solver.J, o.get("J")) < TOL_GRID
    o.solve_tridiag_Ampere(del_t); solver.solve_tridiag_Ampere(world, del_t)
    assert relerr(solver.phi_f, o.get("phi_f")) < 1e-10
    solver.close()


@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf"])
def test_local_temperature_matches_oracle(name, tmp_path):
    """writeLocalTemperature (mod_particle.f90:251-283) on the device: the per-cell temperatures equal the oracle's to summation-order
    round-off, before and after a move; the file is the reference's unformatted sequential record"""
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    numCells = min(world.NumberXNodes, world.NumberXHalfNodes)
    del_t = deck_del_t(d)
    for phase in range(2):
        for sp, p in enumerate(plist):
            Tg, To = p.localTemperature(numCells), o.localTemperature(sp, numCells)
            assert np.array_equal(Tg == 0.0, To == 0.0)
            np.testing.assert_allclose(Tg, To, rtol=1e-12, atol=0)
        if phase == 0:
            solver.phi, solver.phi_f, solver.rho = o.get("phi"), o.get("phi_f"), o.get("rho")
            o.makeHalfTimeEField(); solver.makeHalfTimeEField()
            o.moveParticles(del_t)
            ip.moveParticles(solver, plist, world, del_t)
    E = plist[0].writeLocalTemperature(7, str(tmp_path), numCells)
    raw = (tmp_path / "Temperature" / "Temp_e_7.dat").read_bytes()
    assert len(raw) == 8 * numCells + 8
    assert np.frombuffer(raw[:4], dtype="<i4")[0] == 8 * numCells == np.frombuffer(raw[-4:], dtype="<i4")[0]
    assert np.array_equal(np.frombuffer(raw[4:-4], dtype="<f8"), E)
    solver.close()


@pytest.mark.timeout(120)
def test_local_temperature_on_cell_sorted_particles():
    """cell-sorted input puts lanes of one warp into groups of DIFFERENT sizes (two or three cells per warp): the per-group lane sums of
    k_local_temperature must synchronise over the group, not the warp (a full-mask shuffle inside the group loop dead-locks here).
    Result against the oracle's on the same (unsorted) particles: the per-cell sums do not depend on the particle order beyond round-off."""
    d = DECKS["ngp_rf"]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o)
    numCells = min(world.NumberXNodes, world.NumberXHalfNodes)
    for sp, p in enumerate(plist):
        _lib.check(_lib.load().ipic_sort_particles(solver._h, p._index))
        Tg, To = p.localTemperature(numCells), o.localTemperature(sp, numCells)
        assert np.array_equal(Tg == 0.0, To == 0.0)
        np.testing.assert_allclose(Tg, To, rtol=1e-11, atol=0)
    solver.close()


def test_cell_sort_is_a_stable_sort():
    """ipic_sort_particles orders by cell -- by INT(l * sub), cells cut into sub = min(8, 65534 // (max(N_x, N_h) + 2)) slices -- and keeps
    the original order inside a slice, i.e. it equals numpy's stable argsort of that key: deterministic, which the seeded MCC relies on"""
    d = DECKS["cic_rf"]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o)
    try:
        L = ip.load_library()
        for step in range(2):                               # second round: an already sorted input
            for p in plist:
                ps = p.download()
                before = np.concatenate([ps[t, :p.N_p[t]] for t in range(p.numThread)])
                _lib.check(L.ipic_sort_particles(solver._h, p._index))
                ps = p.download()
                after = np.concatenate([ps[t, :p.N_p[t]] for t in range(p.numThread)])
                sub = min(8, 65534 // (max(world.NumberXNodes, world.NumberXHalfNodes) + 2))
                order = np.argsort((before[:, 0] * float(sub)).astype(np.int64), kind="stable")
                assert np.array_equal(after, before[order])
                assert np.all(np.diff(after[:, 0].astype(np.int64)) >= 0)
            if step == 0:
                ip.moveParticles(solver, plist, world, deck_del_t(d))    # shuffle the cells a little
    finally:
        solver.close()


def test_timing_state_stays_bounded_without_queries():
    """The level-1 Fortran drop-in calls depositJ / moveParticles for millions of steps and never reads a timing: brackets whose events have
    completed are folded into the totals as the run goes (no synchronisation), so the pending list and the event pool stay bounded, and
    no bracket is lost."""
    d = DECKS["ngp_rf"]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o)
    del_t = deck_del_t(d)
    solver.reset_timing()
    n = 3000
    for _ in range(n):
        ip.depositJ(solver, plist, world, del_t)
    _, pending = solver.timing("__pending")
    assert pending < 1500, pending                                        # two brackets per species per call + field ops: ~15000 opened
    ms, cnt = solver.timing("depositJ_sp0")
    assert cnt == n and ms > 0.0
    parked, pending = solver.timing("__pending")
    assert pending == 0 and parked < 4000
    solver.close()


@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf", "ngp_neumann"])
def test_anderson_decision_on_the_device_is_independent_of_the_chunk(name):
    """solveDivAmpereAnderson takes its decisions (converged / stagnated / go on, mod_nonLinSolvers.f90:111-134) on the device and the host
    enqueues `anderson_chunk` iterations ahead; iterations behind the one that ended the solve are no-ops.  Whatever the chunk -- 1 (the
    host reads every iteration's outcome), the default, or more than the solve needs -- the step is the same step: iteration counts,
    sweep counts, sub-step counts, residual history and the fields bit for bit, and the oracle's iteration count."""
    d = DECKS[name]
    ref = None
    # field_cluster: the fused iteration's field kernel as one CTA or as a thread-block cluster of 8 (collect, normal equations and push
    # tables split over the CTAs).  The collect and the tables are the same numbers either way; the sums of the normal equations are
    # added up in another order, so the cluster is compared on counts and to 1e-9 on the fields, the chunks bit for bit.
    for chunk, cluster in ((1, 1), (6, 1), (64, 1), (6, 8)):
        o = build_oracle(d)
        world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
        try:
            solver.set_option("anderson_chunk", chunk)
            solver.set_option("field_cluster", cluster)
            del_t = deck_del_t(d)
            res = []
            t, remain, curr = 0.0, del_t, del_t
            total_sweeps = 0
            solver.reset_timing()
            for step in range(3):
                remain, curr = ip.solvePotential(solver, plist, world, del_t, remain, curr, t)
                t += curr
                st = ip.SolverState(solver)
                total_sweeps += st.sweeps
                res.append((remain, curr, st.iterNumPicard, st.sweeps, st.substeps, st.amountTimeSplits, tuple(st.normHistory),
                            solver.phi_f.tobytes(), solver.J.tobytes(), tuple(p.count() for p in plist)))
                if step == 0:
                    o.solvePotential(del_t, del_t, del_t, 0.0)
                    assert st.iterNumPicard == o.solver_stats()["iterNumPicard"] and st.sweeps == o.solver_stats()["sweeps"]
            # timing counts name executed sweeps only
            _, n0 = solver.timing("depositJ_sp0")
            assert n0 == total_sweeps
        finally:
            solver.close()
        if ref is None:
            ref = res
        elif cluster == 1:
            assert res == ref, chunk
        else:
            for a, b in zip(res, ref):
                assert a[:6] == b[:6] and a[9] == b[9], (chunk, cluster)                # remain, curr, iterations, sweeps, sub-steps, splits, counts
                assert len(a[6]) == len(b[6])
                pa, pb = np.frombuffer(a[7]), np.frombuffer(b[7])
                assert np.max(np.abs(pa - pb)) <= 1e-9 * np.max(np.abs(pb))


def test_compaction_by_filling_keeps_the_particle_set():
    """Arrays without slab identity (device loader, after a cell sort) are compacted by moving live particles of the tail into the
    holes the absorbed ones leave (`fill_compaction`, default) instead of the stable scatter: the same set of survivors, the same loss
    sums and counts, another order.  Both modes from the same seed; the following step's J (an integer sum: order-independent) is
    bit-identical."""
    import torch
    from implicitpic1d_b200.domain import readWorld
    from implicitpic1d_b200 import constants as K
    n_ave, T_e, Nx, L_dom = 1e16, 4.0, 513, 0.02
    res = []
    for fill in (1, 0):
        world = readWorld(0, Nx, L_dom, 0, 0.0, 1, 1, 1, T_e, n_ave)
        solver = ip.potentialSolver(world, 0.0, 0.0, 0.0, 0.0)
        try:
            N = 1_000_000
            L = ip.load_library()
            m_ion = 4.002602 * K.m_amu - K.m_e
            plist = [ip.Particle(K.m_e, -K.e, n_ave * world.L_domain / N, 0, 0, "e", 1),
                     ip.Particle(m_ion, K.e, n_ave * world.L_domain / N, 0, 0, "He+", 1)]
            solver._bind(plist)
            solver.set_option("fill_compaction", fill)
            _lib.check(L.ipic_generate_particles(solver._h, 0, N, T_e, 0.0, 99))
            _lib.check(L.ipic_generate_particles(solver._h, 1, N, 0.025, 0.0, 99))
            for p in plist:
                _lib.check(L.ipic_sort_particles(solver._h, p._index))
            ip.initializeSolver(solver, 1e-8, 0.0, 0, 10, 0.5, 100)
            solver.solveInitialPotential(plist, world, 0.0)
            del_t = K.del_t_from_deck(n_ave, 1.0, 1.0)
            out, loss = [], []
            for step in range(3):
                ip.solvePotential(solver, plist, world, del_t, del_t, del_t, step * del_t)
                for p in plist:
                    p._refresh_stats()
                    arr = torch.stack(p.device_arrays(), dim=1).cpu().numpy()
                    assert arr.shape[0] == p.count() and np.all(arr[:, 0] >= 1.0)
                    order = np.lexsort((arr[:, 3], arr[:, 2], arr[:, 1], arr[:, 0]))
                    out.append((p.count(), p.delIdx, p.wallLoss.tobytes(), arr[order].tobytes()))
                    loss.append(np.concatenate([p.energyLoss, p.momentumLoss]))       # fixed-tree sums over the array order: last bits may differ
                out.append(solver.J.tobytes())
            assert out[0][1] > 0                                        # electrons were absorbed in the first step
            res.append((out, np.array(loss)))
        finally:
            solver.close()
    assert res[0][0] == res[1][0]
    assert np.allclose(res[0][1], res[1][1], rtol=1e-12, atol=0.0)


def test_run_ends_in_shared_memory_on_a_grid_larger_than_the_copy():
    """The entry-stage kernels count the ends of runs of traversed cells in a per-CTA shared-memory copy of the difference grid that holds
    5632 cells; on a larger grid the cells beyond go to the global grid as before.  Same J bit for bit with the copy on or off, for large
    and small tiles."""
    from implicitpic1d_b200.domain import readWorld
    from implicitpic1d_b200 import constants as K
    n_ave, T_e, Nx = 1e16, 4.0, 7001
    world = readWorld(0, Nx, 0.05, 0, 0.0, 1, 1, 1, T_e, n_ave)
    solver = ip.potentialSolver(world, 0.0, 0.0, 0.0, 0.0)
    try:
        N = 1_500_000
        L = ip.load_library()
        m_ion = 4.002602 * K.m_amu - K.m_e
        plist = [ip.Particle(K.m_e, -K.e, n_ave * world.L_domain / N, 0, 0, "e", 1), ip.Particle(m_ion, K.e, n_ave * world.L_domain / N, 0, 0, "He+", 1)]
        solver._bind(plist)
        _lib.check(L.ipic_generate_particles(solver._h, 0, N, T_e, 0.0, 7))
        _lib.check(L.ipic_generate_particles(solver._h, 1, N, 0.025, 0.0, 7))
        for p in plist:
            _lib.check(L.ipic_sort_particles(solver._h, p._index))
        ip.initializeSolver(solver, 1e-8, 0.0, 0, 10, 0.5, 100)
        solver.solveInitialPotential(plist, world, 0.0)
        del_t = 4.0 * K.del_t_from_deck(n_ave, 1.0, 1.0)            # electrons cross tens of these small cells
        ref = None
        for subs in (8, 1):
            for sd in (0, 1):
                solver.set_option("tile_subs", subs)
                solver.set_option("sd_runs", sd)
                ip.depositJ(solver, plist, world, del_t)
                if ref is None:
                    ref = solver.J.copy()
                    assert np.all(np.isfinite(ref)) and np.max(np.abs(ref)) > 0
                    assert depositSubsteps(solver, plist[0]) > 5 * N      # runs of several cells
                assert np.array_equal(ref, solver.J), (subs, sd)
    finally:
        solver.close()
