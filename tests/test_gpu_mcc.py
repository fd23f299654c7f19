"""GPU parity tests of the null-collision MCC (csrc/collisions.cu through the C ABI) against the CPU oracle
(mod_NullCollision.f90 restated in oracle/ipic_oracle.cpp).  The random streams differ by design (counter-based PCG32 per
candidate vs per-thread ran2), so the comparison is statistical, on counts and moments, with the tolerances written below;
what does not depend on the stream (conservation, bookkeeping, reproducibility for a seed) is checked exactly."""
import os
import numpy as np
import pytest
from oracle import oracle_py as op, mcc_py as mp
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib
from implicitpic1d_b200.domain import readWorld

pytestmark = pytest.mark.gpu

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collision_data")
M_HE = 4.002602 * op.m_amu
M_AR = 39.948 * op.m_amu


def ar_states(T=4, ppt=50000, T_e=20.0, seed=1234):
    """the same hot argon plasma on both sides: oracle state and CUDA state built from the oracle's particles"""
    o = op.Oracle(op.NGP, 33, 0.05, T=T)
    e = o.add_species(op.m_e, -op.e, ppt, 2)
    o.set_weight(e, 1.0)
    seeds = op.default_seeds(T)
    o.load_particles(e, seeds, T_e)
    ion = o.add_species(M_AR - op.m_e, op.e, ppt, 2)
    o.set_weight(ion, 1.0)
    o.load_particles(ion, seeds, 0.05)
    mp.add_target(o, M_AR, 2.4e21, 300.0)
    specs = mp.read_null_collision_inputs([os.path.join(DATA, "Ar_electron_table.dat")], [("e", op.m_e), ("Ar+", M_AR - op.m_e)], [("Ar", M_AR)])
    ncs_o = [mp.NullCollisionOracle(o, s) for s in specs]

    world = readWorld(0, 33, 0.05, 0, 0.0, 1, 1, 0, T_e, 1e15)
    solver = ip.potentialSolver(world, 0.0, 0.0, 0.0, 0.0)
    plist = []
    for sp, nm in ((e, "e"), (ion, "Ar+")):
        sc = o.species_scalars(sp)
        ps = o.phase_space(sp)
        p = ip.Particle(sc["mass"], sc["q"], sc["w_p"], 0, ps.shape[1], nm, T)
        p.set_phase_space(ps.copy(), o.N_p(sp))
        plist.append(p)
    solver._bind(plist)
    tl = [ip.targetParticle("Ar", M_AR, 2.4e21, 300.0)]
    ncs_g = ip.readNullCollisionInputs(["Ar_electron_table.dat"], plist, tl, collision_dir=DATA, seed=seed)
    return o, (e, ion), seeds, ncs_o, solver, plist, tl, ncs_g


def alive(p):
    ps = p.download()
    return np.concatenate([ps[t, :p.N_p[t]] for t in range(p.numThread)])


def test_tables_sorted_like_the_reference_constructor():
    o, _, _, ncs_o, solver, plist, tl, ncs_g = ar_states(ppt=100)
    for g, r in zip(ncs_g, ncs_o):
        for t in tl:
            t._bind(solver)
        g._bind(solver)
        tb, sc = r.tables(), r.scalars()
        assert np.array_equal(g.collisionType, tb["collisionType"]) and np.array_equal(g.energyThreshold, tb["energyThreshold"])
        assert abs(g.sigmaVMax - sc["sigmaVMax"]) <= 1e-15 * sc["sigmaVMax"]
        assert g.minEnergy == sc["minEnergy"] and g.maxEnergy == sc["maxEnergy"]
    solver.close()


def test_counts_and_energy_loss_match_oracle_statistically():
    """20 collision calls on the same plasma.  Counts per collision type: |N_gpu - N_oracle| < 5 sqrt(N_gpu + N_oracle) (both are
    sums of independent draws); mean energy loss per collision within 2 % (+ 5 sigma of the mean for rare types); mean electron
    energy after the calls within 0.5 %."""
    o, (e, ion), seeds, ncs_o, solver, plist, tl, ncs_g = ar_states()
    del_t = 1e-10
    for _ in range(20):
        for s in (e, ion):
            mp.set_num_to_collide(o, s)
        for r in ncs_o:
            r.generateCollision(seeds, del_t)
        for p in plist:
            _lib.check(_lib.load().ipic_set_num_to_collide(solver._h, p._index, p.count()))     # numToCollide = N_p for every species first
        for g in ncs_g:
            g.generateCollision(plist, tl, del_t=del_t)
    for g, r, mass in zip(ncs_g, ncs_o, (op.m_e, M_AR - op.m_e)):
        to = r.totals()
        assert abs(g.totalNumberCollidableParticles - to["totalNumberCollidableParticles"]) < 1e-2 * to["totalNumberCollidableParticles"]
        assert abs(g.P_null - r.scalars()["P_null"]) < 1e-15
        for c in range(g.numberCollisions):
            ng, no = int(g.totalAmountCollisions[c]), int(to["totalAmountCollisions"][c])
            assert abs(ng - no) < 5.0 * np.sqrt(ng + no + 1), (c, ng, no)
            if min(ng, no) > 200:
                lg, lo = g.totalEnergyLoss[c] / ng, to["totalEnergyLoss"][c] / no
                ig, io = g.totalIncidentEnergy[c] / ng, to["totalIncidentEnergy"][c] / no
                assert abs(ig - io) < (0.02 + 5.0 / np.sqrt(min(ng, no))) * abs(io), (c, ig, io)
                # per-collision losses scatter by up to ~ mass * incident v^2 (elastic recoil, charge exchange): 5 sigma of the mean
                assert abs(lg - lo) < 0.02 * abs(lo) + 5.0 * mass * io / np.sqrt(min(ng, no)), (c, lg, lo)
    # moments after the calls
    ke_g = plist[0].getTotalKE() / plist[0].count()
    ke_o = o.getTotalKE(e) / o.N_p(e).sum()
    assert abs(ke_g - ke_o) < 5e-3 * ke_o
    assert abs(plist[0].count() - o.N_p(e).sum()) < 5.0 * np.sqrt(2.0 * (o.N_p(e).sum() - 4 * 50000) + 1)
    assert plist[0].count() - 4 * 50000 == plist[1].count() - 4 * 50000 > 0
    solver.close()


def test_bookkeeping_is_exact():
    """what one call changes: exactly the collided particles' velocities, positions never; tallies equal the actual change of
    sum m v^2; each ionization appends one electron and one ion at the parent's position, in candidate order"""
    o, (e, ion), seeds, ncs_o, solver, plist, tl, ncs_g = ar_states(ppt=40000)
    before = [alive(p) for p in plist]
    g = ncs_g[0]
    g.generateCollision(plist, tl, del_t=2e-10)
    after = [alive(p) for p in plist]
    n0 = before[0].shape[0]
    nion = int(g.ionizations)
    ncoll = int(g.totalAmountCollisions.sum())
    assert nion == int(g.totalAmountCollisions[g.collisionType == 2].sum()) > 5
    assert after[0].shape[0] == n0 + nion and after[1].shape[0] == before[1].shape[0] + nion
    assert g.numberSelected in (int(g.P_null * n0), int(g.P_null * n0) + 1)
    old = after[0][:n0]
    assert np.array_equal(old[:, 0], before[0][:, 0])                       # nobody moved
    changed = np.any(old[:, 1:] != before[0][:, 1:], axis=1)
    assert changed.sum() == ncoll <= g.numberSelected                       # distinct candidates: one change per collision
    assert np.array_equal(after[1][:before[1].shape[0]], before[1])         # existing ions untouched by the electron set
    new_e, new_i = after[0][n0:], after[1][before[1].shape[0]:]
    assert np.array_equal(new_e[:, 0], new_i[:, 0])
    parents = set(before[0][changed, 0].tolist())
    assert all(x in parents for x in new_e[:, 0].tolist())
    # energy tallies (:208,216) against the particles
    sumv2 = lambda a: (a[:, 1:] ** 2).sum()
    expect = op.m_e * (sumv2(before[0]) - sumv2(old)) - op.m_e * sumv2(new_e) - M_AR * sumv2(new_i)
    assert abs(g.totalEnergyLoss.sum() - expect) <= 1e-9 * abs(expect)
    inc = (before[0][changed, 1:] ** 2).sum()
    assert abs(g.totalIncidentEnergy.sum() - inc) <= 1e-12 * inc
    # momentum: ionization conserves m_e v + M v_target = m_e v1 + m_e v2 + m_ion v_ion; the ion's velocity stays thermal-sized
    assert np.sqrt((new_i[:, 1:] ** 2).sum(axis=1)).max() < 1e4
    # charge exchange on the ion set: the ion takes a thermal neutral's velocity, counts only
    gi = ncs_g[1]
    gi.generateCollision(plist, tl, del_t=5e-9)
    assert gi.totalAmountCollisions.sum() > 100 and gi.ionizations == 0
    assert plist[1].count() == after[1].shape[0]
    solver.close()


def test_reproducible_for_a_seed_and_different_otherwise():
    res = []
    for seed in (77, 77, 78):
        o, _, _, _, solver, plist, tl, ncs_g = ar_states(ppt=20000, seed=seed)
        for _ in range(3):
            for g in ncs_g:
                g.generateCollision(plist, tl, del_t=2e-10)
        res.append([alive(p) for p in plist] + [ncs_g[0].totalEnergyLoss.copy()])
        solver.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert res[0][0].shape != res[2][0].shape or not np.array_equal(res[0][0], res[2][0])


def test_p_null_above_one_half_stops():
    o, _, _, _, solver, plist, tl, ncs_g = ar_states(ppt=100)
    with pytest.raises(_lib.IpicError) as ei:
        ncs_g[0].generateCollision(plist, tl, del_t=1e-6)                   # P_null ~ 1: the reference prints and stops (:126-129)
    assert ei.value.code == _lib.ERR_STOP
    solver.close()


def test_num_to_collide_excludes_particles_born_this_step():
    """numToCollide is the count after the mover (mod_particleMover.f90:263): products of ionization are not collided again
    within the step, by either set"""
    o, _, _, _, solver, plist, tl, ncs_g = ar_states(ppt=20000)
    n_e, n_i = plist[0].count(), plist[1].count()
    ncs_g[0].generateCollision(plist, tl, del_t=2e-10)
    assert plist[0].count() > n_e
    k1 = ncs_g[0].numberSelected
    ncs_g[1].generateCollision(plist, tl, del_t=5e-9)
    assert ncs_g[1].totalNumberCollidableParticles == n_i
    # a second call within the same step sees the range the first one left: the selected particles were taken out of it
    # (mod_NullCollision.f90:221-231 stores the reduced numToCollide), and the newborn are still excluded
    ncs_g[0].generateCollision(plist, tl, del_t=2e-10)
    assert ncs_g[0].totalNumberCollidableParticles == n_e + (n_e - k1)
    n_now = plist[0].count()
    _lib.check(_lib.load().ipic_set_num_to_collide(solver._h, 0, -1))        # < 0: every particle; a new step starts
    ncs_g[0].generateCollision(plist, tl, del_t=2e-10)
    assert ncs_g[0].totalNumberCollidableParticles == n_e + (n_e - k1) + n_now
    solver.close()


def test_gpu_swarm_reproduces_bolsig_collision_frequency():
    """The reference's known-answer test on the device (OpenMP-BField/TestMCC/MCCExample.f90): He, 25 Td, nu*dt = 0.02,
    2 x 25 000 steps on 2e5 electrons; collision frequency within 2 % of BOLSIG after the (1 - exp(-x))/x null-collision factor."""
    import torch
    world = readWorld(0, 33, 1.0, 0, 0.0, 1, 1, 0, 1.5, 1e15)
    solver = ip.potentialSolver(world, 0.0, 0.0, 0.0, 0.0)
    plist = [ip.Particle(op.m_e, -op.e, 1.0, 0, 0, "e", 1), ip.Particle(M_HE - op.m_e, op.e, 1.0, 0, 0, "He+", 1)]
    solver._bind(plist)
    L = _lib.load()
    _lib.check(L.ipic_generate_particles(solver._h, 0, 200000, 2.25 * 2.0 / 3.0, 0.0, 4242))
    tl = [ip.targetParticle("He", M_HE, 1e21, 300.0)]
    nc = ip.readNullCollisionInputs(["turner_benchmark_he_electron_table_noIonization.dat"], plist, tl, collision_dir=DATA, seed=99)[0]
    bolsig, n_gas = 0.7733e-13, 1e21
    del_t = 2e-2 / (bolsig * n_gas)
    accel = -op.e * (25.0 * 1e-21 * n_gas) * del_t / op.m_e
    _, vx, _, _ = plist[0].device_arrays()
    steps = 25000
    for phase in range(2):
        if phase:
            nc.reset_totals()
        for _ in range(steps):
            vx += accel
            torch.cuda.synchronize()
            _lib.check(L.ipic_set_num_to_collide(solver._h, 0, -1))          # what the mover does every step (mod_particleMover.f90:263)
            nc.generateCollision(plist, tl, del_t=del_t)
    freq = nc.totalAmountCollisions.sum() / steps / del_t / plist[0].count()
    x = nc.sigmaVMax * n_gas * del_t
    corrected = freq * x / (1 - np.exp(-x))
    assert abs(corrected / (bolsig * n_gas) - 1.0) < 0.02, (freq, corrected)
    ke = plist[0].getTotalKE() / plist[0].count() / op.e          # w_p = 1: joules per particle -> eV
    assert 6.5 < ke < 9.0
    solver.close()


def test_collisional_time_loop_moments_match_oracle():
    """BASELINE configs[3] in small: the Vass argon deck (NGP_VassBenchmark_Ar_10xParticles: RF 250 V 13.56 MHz, n = 1e16, eps_a = 1e-4,
    Ar 300 K 2.4e21 m^-3) on 64 cells, stepped as mod_simulation.f90:157-160 does -- solvePotential, then generateCollision of every
    reactant pair over the step just taken -- on both sides from identical particles.  The collision streams differ by design, so
    after 30 steps the comparison is on moments: collision counts |N_gpu - N_oracle| < 5 sqrt(N_gpu + N_oracle + 1) per type,
    particle counts within 5 sigma of the ionisation/wall-loss counts, mean electron and ion energy within 1 %, wall losses
    within 5 sigma, and the same number of nonlinear iterations to +-3 per step (the field sees the same plasma up to noise)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import build_oracle, build_cuda_from_oracle, deck_del_t
    d = dict(scheme=0, Nx=65, L=0.025 / 4, n_ave=1.0e16, T_e=2.59, T_i=0.025, ppt=20000, T=4, bcL=1, bcR=4, Vl=0.0, Vr=250.0, rf=13.56e6,
             smooth=True, frac=1.5, dt_deck=7.845352413230403e-12, eps_a=1e-4, m=10, beta=0.5, maxIter=100, ion_amu=39.948)
    o = build_oracle(d, factor=2)
    world, solver, plist = build_cuda_from_oracle(d, o)
    plist[0].name, plist[1].name = "e", "Ar+"
    mp.add_target(o, M_AR, 2.4e21, 300.0)
    specs = mp.read_null_collision_inputs([os.path.join(DATA, "Ar_electron_table.dat")], [("e", op.m_e), ("Ar+", M_AR - op.m_e)], [("Ar", M_AR)])
    ncs_o = [mp.NullCollisionOracle(o, s) for s in specs]
    seeds = op.default_seeds(d["T"])
    tl = [ip.targetParticle("Ar", M_AR, 2.4e21, 300.0)]
    ncs_g = ip.readNullCollisionInputs(["Ar_electron_table.dat"], plist, tl, collision_dir=DATA, seed=2026)
    del_t = deck_del_t(d)
    n0 = [plist[s].count() for s in range(2)]
    ro, co, rg, cg, to, tg = del_t, del_t, del_t, del_t, 0.0, 0.0
    wall_o, wall_g = np.zeros(2), np.zeros(2)
    for step in range(30):
        ro, co = o.solvePotential(del_t, ro, co, to)
        to += co
        for s in range(2):
            wall_o[s] += o.counters(s)["delIdx"].sum()
            mp.set_num_to_collide(o, s)
        for r in ncs_o:
            r.generateCollision(seeds, co)
        rg, cg = ip.solvePotential(solver, plist, world, del_t, rg, cg, tg)
        tg += cg
        assert abs(ip.SolverState(solver).iterNumPicard - o.solver_stats()["iterNumPicard"]) <= 3
        for s in range(2):
            plist[s]._refresh_stats()
            wall_g[s] += plist[s].delIdx
            _lib.check(_lib.load().ipic_set_num_to_collide(solver._h, plist[s]._index, plist[s].count()))
        for g in ncs_g:
            g.generateCollision(plist, tl, del_t=cg)
    nion_g = nion_o = 0
    for g, r in zip(ncs_g, ncs_o):
        tot = r.totals()
        for c in range(g.numberCollisions):
            ng, no = int(g.totalAmountCollisions[c]), int(tot["totalAmountCollisions"][c])
            assert abs(ng - no) < 5.0 * np.sqrt(ng + no + 1), (c, ng, no)
            if g.collisionType[c] == 2:
                nion_g += ng; nion_o += no
    assert ncs_g[0].totalAmountCollisions.sum() > 2500          # the loop did collide
    for s in range(2):
        ng, no = plist[s].count(), int(o.N_p(s).sum())
        assert ng == n0[s] + nion_g - int(wall_g[s]) and no == n0[s] + nion_o - int(wall_o[s])     # exact bookkeeping on both sides
        assert abs(wall_g[s] - wall_o[s]) < 5.0 * np.sqrt(wall_g[s] + wall_o[s] + 1), (s, wall_g, wall_o)
        ke_g, ke_o = plist[s].getTotalKE() / ng, o.getTotalKE(s) / no
        assert abs(ke_g - ke_o) < 1e-2 * ke_o, (s, ke_g, ke_o)
    solver.close()


def test_library_constructor_matches_the_reference_run_log():
    """the library's nullCollision constructor (collision_set_create) on the He table of the reference's default decks reproduces what
    the REAL reference binary printed (OpenMP/NGP/src/output.txt -> tests/golden/reference_run_logs.json): max sigma*v to the digits
    printed, reduced masses, collision order by descending max sigma*v"""
    import json
    from implicitpic1d_b200 import constants as K
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run_logs.json")))["ngp"]
    m_ion, m_He = 4.002602 * K.m_amu - K.m_e, 4.002602 * K.m_amu
    world = readWorld(0, 33, 0.05, 0, 0.0, 1, 1, 0, 4.0, 1e15)
    solver = ip.potentialSolver(world, 0.0, 0.0, 0.0, 0.0)
    plist = []
    for nm, m, q in (("e", K.m_e, -K.e), ("He+", m_ion, K.e)):
        p = ip.Particle(m, q, 1.0, 0, 8, nm, 1)
        p.set_phase_space(np.zeros((1, 8, 4)), np.zeros(1, dtype=np.int32))
        plist.append(p)
    solver._bind(plist)
    tl = [ip.targetParticle("He", m_He, g["target"]["density"], g["target"]["temperature"])]
    ncs = ip.readNullCollisionInputs(["turner_benchmark_he_electron_table.dat"], plist, tl, collision_dir=DATA)
    assert len(ncs) == 2
    for t in tl:
        t._bind(solver)
    for nc, ref in zip(ncs, g["nullCollisions"]):
        nc._bind(solver)
        assert nc.sigmaVMax == pytest.approx(ref["sigmaVMax"], rel=5e-15)
        assert nc.collisionType.tolist() == [c["collisionType"] for c in ref["collisions"]]
        assert nc.energyThreshold.tolist() == [c["energyThreshold"] for c in ref["collisions"]]
        assert nc.lengthArrays == ref["lengthArrays"]
    solver.close()


def test_two_sets_with_one_primary_never_select_a_particle_twice_in_a_step():
    """Two nullCollision sets whose primary is the electrons (electrons on two gases): the reference's serial loop swaps every selected
    particle behind the collidable range and stores the reduced numToCollide (mod_NullCollision.f90:221-231), so the second set cannot
    pick a particle the first one collided, and its totalNumberCollidableParticles counts the reduced range.  Here both sets walk one
    keyed permutation of the step's range, one after the other."""
    o, (e, ion), seeds, ncs_o, solver, plist, tl, ncs_a = ar_states(ppt=20000)
    ncs_b = ip.readNullCollisionInputs(["Ar_electron_table.dat"], plist, tl, collision_dir=DATA, seed=987654321)
    a = [g for g in ncs_a if g.reactantsIndx[0] == 1][0]
    b = [g for g in ncs_b if g.reactantsIndx[0] == 1][0]
    pe = plist[0]
    N = pe.count()
    del_t = 2.5e-10                                                       # P_null ~ 0.28: independent draws would share ~0.08 N particles
    L = _lib.load()
    for step in range(2):
        for p in plist:
            _lib.check(L.ipic_set_num_to_collide(solver._h, p._index, p.count()))
        N = pe.count()
        s0 = alive(pe)[:N, 1:4].copy()
        tot_b0 = getattr(b, "totalNumberCollidableParticles", 0)
        a.generateCollision(plist, tl, del_t=del_t)
        s1 = alive(pe)[:N, 1:4].copy()
        b.generateCollision(plist, tl, del_t=del_t)
        s2 = alive(pe)[:N, 1:4].copy()
        c1 = np.any(s1 != s0, axis=1)
        c2 = np.any(s2 != s1, axis=1)
        assert 0.1 * N < c1.sum() <= a.numberSelected and 0.05 * N < c2.sum() <= b.numberSelected
        assert not np.any(c1 & c2)                                        # disjoint: nobody collides twice in the step
        assert b.totalNumberCollidableParticles - tot_b0 == N - a.numberSelected
        assert abs(b.numberSelected - b.P_null * (N - a.numberSelected)) <= 1.0
    solver.close()
