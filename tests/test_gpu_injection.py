"""GPU tests of wall refluxing / injection / Maxwellian heating (csrc/injection.cu through the C ABI and the Python mirror of
mod_particleInjection) against the CPU oracle (mod_particleInjection.f90 restated in oracle/ipic_oracle.cpp).

The device draws from counter-based PCG32 streams, the reference from per-thread ran2 states, so what depends on the stream is
compared statistically (5-sigma bounds on moments, written where they are used); what does not -- how many particles are added,
which particles change, signs, where new pairs sit, the energy tally against the actual change of sum(m w v^2 / 2), bit
reproducibility for a seed -- is checked exactly."""
import numpy as np
import pytest
from oracle import oracle_py as op
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib, particle_injection as pinj
from common import DECKS, build_oracle, build_cuda_from_oracle, deck_del_t

pytestmark = pytest.mark.gpu

NEUMANN = ["ngp_neumann", "cic_neumann"]


def moved_states(name, mult=8.0, seed=0xBEEF):
    """oracle + CUDA state on identical particles right after one moveParticles with wall traffic"""
    d = DECKS[name]
    o = build_oracle(d, factor=2)
    world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
    del_t = deck_del_t(d) * mult
    solver.phi, solver.phi_f, solver.rho = o.get("phi"), o.get("phi_f"), o.get("rho")
    o.makeHalfTimeEField(); solver.makeHalfTimeEField()
    o.moveParticles(del_t)
    ip.moveParticles(solver, plist, world, del_t)
    pinj.setInjectionSeed(solver, seed)
    return d, o, world, solver, plist, del_t


def alive(p):
    ps = p.download()
    return np.concatenate([ps[t, :p.N_p[t]] for t in range(p.numThread)]).copy()


def ke(a, p):
    return 0.5 * p.mass * p.w_p * np.sum(a[:, 1:4] ** 2, axis=1)


def vth(T_eV, mass):
    return np.sqrt(T_eV * op.e / mass)


def x_of(o, l):
    return np.array([o.getXFromL(float(v)) for v in l])


def walls(o, world, d):
    """x of the two walls: l = 1 and l = N_x (NGP) / N_h (CIC), the latter approached from inside its cell"""
    NB = world.NumberXNodes if d["scheme"] == 0 else world.NumberXNodes + 1
    return o.getXFromL(1.0), o.getXFromL(float(NB) - 1e-9)


@pytest.mark.parametrize("name", NEUMANN)
def test_addMaxwellianLostParticles(name):
    d, o, world, solver, plist, del_t = moved_states(name)
    T_e, T_i = d["T_e"], d["T_i"]
    before = [alive(p) for p in plist]
    lost = int(o.counters(0)["delIdx"].sum())
    assert lost > 30
    e0 = pinj.energyAddColl
    n = pinj.addMaxwellianLostParticles(plist, T_e, T_i, None, world)
    assert n == lost == plist[0].delIdx                                    # one pair per electron absorbed by the last move (:24)
    en_gpu = pinj.energyAddColl - e0
    after = [alive(p) for p in plist]
    for b, a in zip(before, after):
        assert a.shape[0] == b.shape[0] + n
        assert np.array_equal(a[:b.shape[0]], b)                           # nobody else is touched
    ne, ni = after[0][-n:], after[1][-n:]
    assert np.array_equal(ne[:, 0], ni[:, 0])                              # the ion sits on its electron (:33)
    x = x_of(o, ne[:, 0])
    x0, x1 = walls(o, world, d)
    assert np.all((x >= x0) & (x <= x1 * (1.0 + 1e-9)))
    assert abs(np.mean(x) - 0.5 * (x0 + x1)) < 5.0 * (x1 - x0) / np.sqrt(12.0 * n)
    # energy tally = actual energy of the new particles (:29-30)
    assert en_gpu == pytest.approx(ke(ne, plist[0]).sum() + ke(ni, plist[1]).sum(), rel=1e-12)
    # electrons: 3-D Maxwellian at T_e; ions: flux sample at T_i (|v_x| Rayleigh: <v_x^2> = 2 v_th^2)
    ve, vi = vth(T_e, plist[0].mass), vth(T_i, plist[1].mass)
    for k in (1, 2, 3):
        assert abs(np.mean(ne[:, k])) < 5.0 * ve / np.sqrt(n)
        assert abs(np.var(ne[:, k]) / ve ** 2 - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(np.mean(ni[:, 1] ** 2) / (2.0 * vi ** 2) - 1.0) < 5.0 / np.sqrt(n)
    assert abs(np.mean(ni[:, 1] > 0) - 0.5) < 5.0 * 0.5 / np.sqrt(n)
    assert abs(np.mean(ni[:, 2] ** 2 + ni[:, 3] ** 2) / (2.0 * vi ** 2) - 1.0) < 5.0 / np.sqrt(n)
    # the oracle adds the same number of pairs and, within the sampling error of n draws, the same energy
    irand = op.default_seeds(d["T"])
    en_o = o.addMaxwellianLostParticles(T_e, T_i, irand)
    for sp, p in enumerate(plist):
        assert int(o.N_p(sp).sum()) == p.count()
    assert abs(en_gpu / en_o - 1.0) < 6.0 / np.sqrt(n)
    solver.close()


@pytest.mark.parametrize("name", NEUMANN)
def test_refluxParticles(name):
    d, o, world, solver, plist, del_t = moved_states(name)
    T_e, T_i = d["T_e"], d["T_i"]
    before = [alive(p) for p in plist]
    listed = []
    for sp, p in enumerate(plist):
        refIdx, rec = p.refRecord()
        assert np.array_equal(refIdx, o.counters(sp)["refIdx"])
        off = np.concatenate([[0], np.cumsum(p.N_p.astype(np.int64))])
        listed.append(np.concatenate([rec[t].astype(np.int64) - 1 + off[t] for t in range(p.numThread)]))
    assert listed[0].size > 30
    e0 = pinj.energyAddColl
    pinj.refluxParticles(plist, T_e, T_i, None, world)
    en_gpu = pinj.energyAddColl - e0
    after = [alive(p) for p in plist]
    want = 0.0
    left_neumann = d["bcL"] == 2
    for sp, p in enumerate(plist):
        b, a = before[sp], after[sp]
        assert a.shape == b.shape
        changed = np.nonzero(np.any(a != b, axis=1))[0]
        assert np.array_equal(changed, np.sort(listed[sp]))                # exactly the particles of refRecordIdx (:51)
        assert np.array_equal(a[:, 0], b[:, 0])                            # positions stay
        vx = a[listed[sp], 1]
        assert np.all(vx > 0) if left_neumann else np.all(vx < 0)         # away from the Neumann wall (:59-63)
        want += (ke(a, p) - ke(b, p)).sum()
        n = listed[sp].size
        if n > 30:
            v = vth(T_e if sp == 0 else T_i, p.mass)
            assert abs(np.mean(vx ** 2) / (2.0 * v ** 2) - 1.0) < 5.0 / np.sqrt(n)
            assert abs(np.mean(a[listed[sp], 2] ** 2 + a[listed[sp], 3] ** 2) / (2.0 * v ** 2) - 1.0) < 5.0 / np.sqrt(n)
    scale = sum(ke(after[sp][listed[sp]], p).sum() + ke(before[sp][listed[sp]], p).sum() for sp, p in enumerate(plist))
    assert abs(en_gpu - want) <= 1e-12 * scale                             # energyAddColl = new - old (:52, :64)
    # the oracle redraws the same particles
    ob = [o.phase_space(sp).copy() for sp in range(2)]
    o.refluxParticles(T_e, T_i, op.default_seeds(d["T"]))
    for sp, p in enumerate(plist):
        oa = o.phase_space(sp)
        nch = sum(int(np.any(oa[t, :p.N_p[t]] != ob[sp][t, :p.N_p[t]], axis=1).sum()) for t in range(p.numThread))
        assert nch == listed[sp].size
    solver.close()


@pytest.mark.parametrize("name", NEUMANN)
def test_injectAtBoundary(name):
    d, o, world, solver, plist, del_t = moved_states(name)
    T_e, T_i = d["T_e"], d["T_i"]
    T = d["T"]
    cs = np.sqrt(T_e * op.e / plist[1].mass)
    # a flux that injects 80.3 pairs per OpenMP thread and call
    flux = 80.3 * T * plist[0].w_p / del_t
    pinj.setInjectionInputs(injection=True, flux=flux, del_t=del_t, w_p=plist[0].w_p, numThread=T)
    assert (pinj.numFluxParticlesLow, pinj.numFluxParticlesHigh) == (80, 81)
    assert pinj.injectionR == pytest.approx(0.3, abs=1e-9)
    before = [alive(p) for p in plist]
    listed = []
    for p in plist:
        _, rec = p.refRecord()
        off = np.concatenate([[0], np.cumsum(p.N_p.astype(np.int64))])
        listed.append(np.concatenate([rec[t].astype(np.int64) - 1 + off[t] for t in range(p.numThread)]))
    n = pinj.injectAtBoundary(plist, T_e, T_i, None, world, del_t)
    assert n in (int(np.floor(80.3 * T)), int(np.floor(80.3 * T)) + 1)     # floor / floor + 1 of the expectation (:100-104)
    after = [alive(p) for p in plist]
    left_neumann = d["bcL"] == 2
    sgn = 1.0 if left_neumann else -1.0
    for sp, p in enumerate(plist):
        b, a = before[sp], after[sp]
        assert a.shape[0] == b.shape[0] + n
        changed = np.nonzero(np.any(a[:b.shape[0]] != b, axis=1))[0]
        assert np.array_equal(changed, np.sort(listed[sp]))                # reflected particles are redrawn (:80-98)
        assert np.all(sgn * a[listed[sp], 1] > 0)
    # reflected and injected ions leave the wall at the ion sound speed (:87-89, :111-113)
    for rows in (after[1][listed[1]], after[1][-n:]):
        assert np.array_equal(rows[:, 1], np.full(rows.shape[0], sgn * cs))
        assert not rows[:, 2:4].any()
    ne, ni = after[0][-n:], after[1][-n:]
    assert np.array_equal(ne[:, 0], ni[:, 0])
    assert np.all(sgn * ne[:, 1] > 0)
    wall = walls(o, world, d)[0 if left_neumann else 1]
    dist = np.abs(x_of(o, ne[:, 0]) - wall)
    assert np.all(dist <= del_t * cs * (1.0 + 1e-6))                       # within del_t * c_s of the Neumann wall (:108, :118/:123)
    assert abs(np.mean(dist) / (del_t * cs) - 0.5) < 5.0 / np.sqrt(12.0 * n)
    ve = vth(T_e, plist[0].mass)
    assert abs(np.mean(ne[:, 1] ** 2) / (2.0 * ve ** 2) - 1.0) < 5.0 / np.sqrt(n)
    # oracle: every thread injects 80 or 81, same placement statistics
    Np0 = [o.N_p(sp).copy() for sp in range(2)]
    o.injectAtBoundary(T_e, del_t, 80, 81, pinj.injectionR, op.default_seeds(T))
    Np1 = [o.N_p(sp) for sp in range(2)]
    added = int((Np1[0] - Np0[0]).sum())
    assert 80 * T <= added <= 81 * T and np.array_equal(Np1[1] - Np0[1], Np1[0] - Np0[0])
    po = o.phase_space(0)
    lo = np.concatenate([po[t, Np0[0][t]:Np1[0][t], 0] for t in range(T)])
    dist_o = np.abs(x_of(o, lo) - wall)
    assert np.all(dist_o <= del_t * cs * (1.0 + 1e-6))
    assert abs(np.mean(dist_o) - np.mean(dist)) < 5.0 * del_t * cs * np.sqrt(1.0 / (12.0 * n) + 1.0 / (12.0 * lo.size))
    solver.close()


def test_injectUniformFlux():
    name = "ngp_neumann"
    d, o, world, solver, plist, del_t = moved_states(name)
    T_e, T_i, T = d["T_e"], d["T_i"], d["T"]
    flux = 150.6 * T * plist[0].w_p / del_t
    pinj.setInjectionInputs(uniformInjection=True, flux=flux, del_t=del_t, w_p=plist[0].w_p, numThread=T)
    before = [alive(p) for p in plist]
    n = pinj.injectUniformFlux(plist, T_e, T_i, None, world)
    assert n in (int(np.floor(150.6 * T)), int(np.floor(150.6 * T)) + 1)
    after = [alive(p) for p in plist]
    for b, a in zip(before, after):
        assert a.shape[0] == b.shape[0] + n and np.array_equal(a[:b.shape[0]], b)
    ne, ni = after[0][-n:], after[1][-n:]
    assert np.array_equal(ne[:, 0], ni[:, 0])
    ve, vi = vth(20.0 * T_e, plist[0].mass), vth(T_i, plist[1].mass)       # electrons enter at 20 T_e (:149)
    for k in (1, 2, 3):
        assert abs(np.var(ne[:, k]) / ve ** 2 - 1.0) < 5.0 * np.sqrt(2.0 / n)
        assert abs(np.var(ni[:, k]) / vi ** 2 - 1.0) < 5.0 * np.sqrt(2.0 / n)
    x = x_of(o, ne[:, 0])
    x0, x1 = walls(o, world, d)
    assert abs(np.mean(x) - 0.5 * (x0 + x1)) < 5.0 * (x1 - x0) / np.sqrt(12.0 * n)
    n0 = int(o.N_p(0).sum())
    o.injectUniformFlux(T_e, T_i, 150, 151, 0.6, op.default_seeds(T))
    assert 150 * T <= int(o.N_p(0).sum()) - n0 <= 151 * T
    solver.close()


@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf"])
def test_maxwellianHeating(name):
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o)
    pinj.setInjectionSeed(solver, 77)
    pinj.setInjectionInputs(heating=True, fractionFreqHeating=5.0)
    p = plist[0]
    before = alive(p)
    N = before.shape[0]
    T_heat = 3.0 * d["T_e"]
    # compValue = FracFreq * (currDel_t / del_t) / FractionFreqHeating = 2.0 * 0.75 / 5.0 = 0.3 (:167)
    n = pinj.maxwellianHeating(p, None, 2.0, T_heat, 0.75, 1.0)
    after = alive(p)
    changed = np.any(after != before, axis=1)
    assert int(changed.sum()) == n
    assert abs(n - 0.3 * N) < 5.0 * np.sqrt(N * 0.3 * 0.7)
    assert np.array_equal(after[:, 0], before[:, 0])
    v = vth(T_heat, p.mass)
    for k in (1, 2, 3):
        assert abs(np.var(after[changed, k]) / v ** 2 - 1.0) < 5.0 * np.sqrt(2.0 / n)
    # the ions are not touched; the oracle heats the same fraction to the same temperature
    ke0 = o.getTotalKE(0)
    o.maxwellianHeating(0, 0.3, T_heat, op.default_seeds(d["T"]))
    ke_gpu = ke(after, p).sum()
    assert abs(ke_gpu / o.getTotalKE(0) - 1.0) < 0.03 and o.getTotalKE(0) > 1.3 * ke0
    solver.close()


def test_injection_is_bit_reproducible_for_a_seed_and_differs_across_seeds():
    outs = []
    for seed in (5, 5, 6):
        d, o, world, solver, plist, del_t = moved_states("ngp_neumann", seed=seed)
        pinj.setInjectionInputs(injection=True, flux=40.5 * d["T"] * plist[0].w_p / del_t, del_t=del_t, w_p=plist[0].w_p, numThread=d["T"])
        pinj.addMaxwellianLostParticles(plist, d["T_e"], d["T_i"], None, world)
        pinj.refluxParticles(plist, d["T_e"], d["T_i"], None, world)
        pinj.injectAtBoundary(plist, d["T_e"], d["T_i"], None, world, del_t)
        pinj.maxwellianHeating(plist[0], None, 1.0, d["T_e"], 0.1, 1.0)
        outs.append([alive(p) for p in plist])
        solver.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    assert outs[0][0].shape != outs[2][0].shape or not np.array_equal(outs[0][0], outs[2][0])


def test_reinjection_keeps_the_electron_count_over_a_time_loop():
    """mod_simulation.f90:157-163: moveParticles then addMaxwellianLostParticles, every step; the electron count is conserved and the
    ions gain (electrons lost - ions lost) per step; the appended particles take part in the next implicit step"""
    d, o, world, solver, plist, del_t = moved_states("ngp_neumann", mult=2.0)
    pinj.addMaxwellianLostParticles(plist, d["T_e"], d["T_i"], None, world)
    Ne = plist[0].count()
    Ni = plist[1].count()
    t = 0.0
    for step in range(4):
        ip.solvePotential(solver, plist, world, del_t, del_t, del_t, t)
        for p in plist:
            p._refresh_stats()
        lost_e, lost_i = plist[0].delIdx, plist[1].delIdx
        assert plist[0].count() == Ne - lost_e
        n = pinj.addMaxwellianLostParticles(plist, d["T_e"], d["T_i"], None, world)
        assert n == lost_e
        assert plist[0].count() == Ne
        Ni += lost_e - lost_i
        assert plist[1].count() == Ni
        t += del_t
    solver.close()


def test_injection_edge_cases_and_error_codes():
    """no-op and error behaviour of the entry points: nothing to reflux without a Neumann wall, nothing injected at zero flux, the heated
    fraction at compValue 0 and 1, argument errors as status codes with a message (never a crash, never a silent fallback)"""
    import ctypes as C
    L = _lib.load()
    d = DECKS["ngp_rf"]                                            # Dirichlet / RF walls: the mover records no reflections
    o = build_oracle(d, factor=2)
    world, solver, plist = build_cuda_from_oracle(d, o)
    del_t = deck_del_t(d)
    ip.moveParticles(solver, plist, world, del_t)
    before = [alive(p) for p in plist]
    e0 = pinj.energyAddColl
    pinj.refluxParticles(plist, d["T_e"], d["T_i"], None, world)
    assert pinj.energyAddColl == e0
    pinj.setInjectionInputs(injection=True, flux=0.0, del_t=del_t, w_p=plist[0].w_p, numThread=d["T"])
    assert pinj.injectAtBoundary(plist, d["T_e"], d["T_i"], None, world, del_t) == 0
    pinj.setInjectionInputs(uniformInjection=True, flux=0.0, del_t=del_t, w_p=plist[0].w_p, numThread=d["T"])
    assert pinj.injectUniformFlux(plist, d["T_e"], d["T_i"], None, world) == 0
    pinj.setInjectionInputs(heating=True, fractionFreqHeating=1.0)
    assert pinj.maxwellianHeating(plist[0], None, 0.0, d["T_e"], 1.0, 1.0) == 0
    for b, p in zip(before, plist):
        assert np.array_equal(alive(p), b)
    N = plist[0].count()
    assert pinj.maxwellianHeating(plist[0], None, 1.0, d["T_e"], 1.0, 1.0) == N      # R in (0, 1) < compValue = 1 for every particle
    pinj.setInjectionInputs()
    # argument errors
    n = C.c_int64()
    out = np.zeros(3)
    assert L.ipic_maxwellian_heating(solver._h, 7, 0.1, 1.0, C.byref(n)) == _lib.ERR_ARG
    assert L.ipic_velocity_extrema(solver._h, -1, out.ctypes.data_as(_lib._dp)) == _lib.ERR_ARG
    VH, EH = np.zeros(4), np.zeros(2)
    assert L.ipic_velocity_histograms(solver._h, 0, 1, 1.0, 1.0, 2.0, VH.ctypes.data_as(_lib._dp), EH.ctypes.data_as(_lib._dp)) == _lib.ERR_ARG
    assert L.ipic_velocity_histograms(solver._h, 0, 2, 0.0, 1.0, 2.0, VH.ctypes.data_as(_lib._dp), EH.ctypes.data_as(_lib._dp)) == _lib.ERR_ARG
    assert b"histograms" in L.ipic_last_error()
    assert L.ipic_load_particle_density(solver._h, 0, None) == _lib.ERR_ARG
    solver.close()
    # a one-species context cannot re-inject electron-ion pairs
    world1 = ip.domain.readWorld(0, 33, 0.05, 0, 0.0, 1, 1, 0, 2.0, 1e15)
    s1 = ip.potentialSolver(world1, 0.0, 0.0, 0.0, 0.0)
    p1 = ip.Particle(op.m_e, -op.e, 1.0, 0, 16, "e", 1)
    p1.set_phase_space(np.zeros((1, 16, 4)), np.zeros(1, dtype=np.int32))
    s1._bind([p1])
    en = C.c_double()
    assert L.ipic_add_maxwellian_lost_particles(s1._h, 1.0, 1.0, C.byref(en), C.byref(n)) == _lib.ERR_ARG
    # an empty species: density zeros, extrema at their initial values
    dens = np.ones(33)
    _lib.check(L.ipic_load_particle_density(s1._h, 0, dens.ctypes.data_as(_lib._dp)))
    assert not dens.any()
    _lib.check(L.ipic_velocity_extrema(s1._h, 0, out.ctypes.data_as(_lib._dp)))
    assert out[0] == 0.0 and out[1] == 0.0 and out[2] == 0.1 * op.e * 2.0 / op.m_e
    s1.close()


@pytest.mark.parametrize("name", NEUMANN)
def test_refluxParticles_after_a_domain_crossing_step(name):
    """A step long enough for electrons to be reflected at the Neumann wall AND absorbed at the other wall: their refRecordIdx entries name
    the successor's slot (reference: i - delIdx), which may itself be listed -- two entries, one particle.  The device redraws every
    reflected particle that survived exactly once and skips the entries of absorbed ones: no particle outside the list changes, the
    energy tally equals the actual change of sum(m w v^2 / 2) (it would not if two threads raced on one particle), and the result is
    bit-reproducible."""
    runs = []
    for rep in range(2):
        d, o, world, solver, plist, del_t = moved_states(name, mult=400.0)
        T_e, T_i = d["T_e"], d["T_i"]
        before = [alive(p) for p in plist]
        listed, dup = [], 0
        for sp, p in enumerate(plist):
            refIdx, rec = p.refRecord()
            assert np.array_equal(refIdx, o.counters(sp)["refIdx"])
            off = np.concatenate([[0], np.cumsum(p.N_p.astype(np.int64))])
            ent = np.concatenate([rec[t].astype(np.int64) - 1 + off[t] for t in range(p.numThread)]) if refIdx.sum() else np.zeros(0, dtype=np.int64)
            dup += ent.size - np.unique(ent).size
            listed.append(ent)
        assert o.counters(0)["delIdx"].sum() > 0 and listed[0].size > 0
        e0 = pinj.energyAddColl
        pinj.refluxParticles(plist, T_e, T_i, None, world)
        en_gpu = pinj.energyAddColl - e0
        after = [alive(p) for p in plist]
        want, scale = 0.0, 0.0
        for sp, p in enumerate(plist):
            b, a = before[sp], after[sp]
            assert a.shape == b.shape
            changed = np.nonzero(np.any(a != b, axis=1))[0]
            assert np.all(np.isin(changed, listed[sp]))                    # nothing outside refRecordIdx is touched
            assert np.array_equal(a[:, 0], b[:, 0])
            if changed.size:
                vx = a[changed, 1]
                assert np.all(vx > 0) if d["bcL"] == 2 else np.all(vx < 0)
            want += (ke(a, p) - ke(b, p)).sum()
            scale += ke(a[changed], p).sum() + ke(b[changed], p).sum()
        assert abs(en_gpu - want) <= 1e-12 * max(scale, 1e-300)
        runs.append(after)
        solver.close()
    for sp in range(2):
        assert np.array_equal(runs[0][sp], runs[1][sp])
