"""CPU tests of the oracle's restatement of mod_particleInjection.f90 (no golden vectors exist in the reference for these
routines: they are pinned on what the routines state -- distributions of the samplers of mod_BasicFunctions.f90:146-176, the
energy tally against the actual change of sum(m w v^2 / 2), counts, signs) and of the host-side reader / writer of
ParticleInjection.inp (readInjectionInputs :181-252)."""
import numpy as np
import pytest
from oracle import oracle_py as op
from implicitpic1d_b200 import particle_injection as pinj
from common import DECKS, build_oracle, deck_del_t


def moved_oracle(name, mult=8.0):
    d = DECKS[name]
    o = build_oracle(d, factor=2)
    del_t = deck_del_t(d) * mult
    o.makeHalfTimeEField()
    o.moveParticles(del_t)
    return d, o, del_t


def rows(o, sp, lo=None, hi=None):
    ps, Np = o.phase_space(sp), o.N_p(sp)
    lo = np.zeros_like(Np) if lo is None else lo
    hi = Np if hi is None else hi
    return np.concatenate([ps[t, lo[t]:hi[t]] for t in range(o.T)]).copy()


def ke(a, sc):
    return 0.5 * sc["mass"] * sc["w_p"] * np.sum(a[:, 1:4] ** 2, axis=1)


def test_oracle_add_lost_particles_counts_energy_and_moments():
    d, o, del_t = moved_oracle("ngp_neumann")
    lost = o.counters(0)["delIdx"].copy()
    Np0 = [o.N_p(sp).copy() for sp in range(2)]
    assert lost.sum() > 30
    en = o.addMaxwellianLostParticles(d["T_e"], d["T_i"], op.default_seeds(d["T"]))
    for sp in range(2):
        assert np.array_equal(o.N_p(sp), Np0[sp] + lost)            # both species grow by the ELECTRONS' delIdx (:35-36)
    ne, ni = rows(o, 0, Np0[0]), rows(o, 1, Np0[1])
    assert np.array_equal(ne[:, 0], ni[:, 0])
    assert en == pytest.approx(ke(ne, o.species_scalars(0)).sum() + ke(ni, o.species_scalars(1)).sum(), rel=1e-12)
    n = ne.shape[0]
    ve = np.sqrt(d["T_e"] * op.e / o.species_scalars(0)["mass"])
    vi = np.sqrt(d["T_i"] * op.e / o.species_scalars(1)["mass"])
    for k in (1, 2, 3):
        assert abs(np.mean(ne[:, k])) < 5.0 * ve / np.sqrt(n)
        assert abs(np.var(ne[:, k]) / ve ** 2 - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(np.mean(ni[:, 1] ** 2) / (2.0 * vi ** 2) - 1.0) < 5.0 / np.sqrt(n)      # flux sample: |v_x| Rayleigh
    assert 1.0 <= ne[:, 0].min() and ne[:, 0].max() <= d["Nx"]


@pytest.mark.parametrize("name", ["ngp_neumann", "cic_neumann"])
def test_oracle_reflux_redraws_exactly_the_recorded_particles(name):
    d, o, del_t = moved_oracle(name)
    before = [o.phase_space(sp).copy() for sp in range(2)]
    en = o.refluxParticles(d["T_e"], d["T_i"], op.default_seeds(d["T"]))
    want, total = 0.0, 0
    for sp in range(2):
        sc, Np, cnt = o.species_scalars(sp), o.N_p(sp), o.counters(sp)
        after = o.phase_space(sp)
        for t in range(o.T):
            rec = o.ref_record(sp, t + 1, int(cnt["refIdx"][t])) - 1
            ch = np.nonzero(np.any(after[t, :Np[t]] != before[sp][t, :Np[t]], axis=1))[0]
            assert np.array_equal(ch, np.sort(rec))
            vx = after[t, rec, 1]
            assert np.all(vx > 0) if d["bcL"] == 2 else np.all(vx < 0)
            want += (ke(after[t, rec], sc) - ke(before[sp][t, rec], sc)).sum()
            total += rec.size
    assert total > 30
    assert en == pytest.approx(want, rel=1e-9)


def test_oracle_boundary_and_uniform_injection():
    d, o, del_t = moved_oracle("ngp_neumann")
    T = d["T"]
    cs = np.sqrt(d["T_e"] * op.e / o.species_scalars(1)["mass"])
    Np0 = [o.N_p(sp).copy() for sp in range(2)]
    o.injectAtBoundary(d["T_e"], del_t, 50, 51, 0.5, op.default_seeds(T))
    add = o.N_p(0) - Np0[0]
    assert np.all((add == 50) | (add == 51)) and np.array_equal(o.N_p(1) - Np0[1], add)
    ne, ni = rows(o, 0, Np0[0]), rows(o, 1, Np0[1])
    assert np.array_equal(ni[:, 1], np.full(ni.shape[0], cs)) and not ni[:, 2:].any()    # left wall is Neumann: +c_s
    assert np.all(ne[:, 1] > 0)
    x = np.array([o.getXFromL(float(l)) for l in ne[:, 0]])
    assert np.all(x - o.getXFromL(1.0) <= del_t * cs * (1 + 1e-9))
    Np1 = [o.N_p(sp).copy() for sp in range(2)]
    o.injectUniformFlux(d["T_e"], d["T_i"], 200, 201, 0.25, op.default_seeds(T))
    ne = rows(o, 0, Np1[0])
    ve = np.sqrt(20.0 * d["T_e"] * op.e / o.species_scalars(0)["mass"])
    n = ne.shape[0]
    assert 200 * T <= n <= 201 * T
    assert abs(np.var(ne[:, 1]) / ve ** 2 - 1.0) < 5.0 * np.sqrt(2.0 / n)


def test_oracle_maxwellian_heating_fraction():
    d = DECKS["ngp_rf"]
    o = build_oracle(d)
    before = rows(o, 0)
    o.maxwellianHeating(0, 0.3, 10.0, op.default_seeds(d["T"]))
    after = rows(o, 0)
    ch = np.any(after != before, axis=1)
    N = before.shape[0]
    assert abs(ch.sum() - 0.3 * N) < 5.0 * np.sqrt(N * 0.21)
    v = np.sqrt(10.0 * op.e / o.species_scalars(0)["mass"])
    assert abs(np.var(after[ch, 1]) / v ** 2 - 1.0) < 5.0 * np.sqrt(2.0 / ch.sum())


def test_injection_input_reader_and_writer(tmp_path):
    # the reference's ParticleInjection.inp layout (five `flag[, value] ! comment` lines, Fortran d-exponents)
    (tmp_path / "ParticleInjection.inp").write_text(
        "0 ! lost particles re-injected\n1 ! reflux\n1, 5.6d24 ! sonic injection, flux\n0, 100.0d0 ! heating\n")
    pinj.readInjectionInputs("ParticleInjection.inp", w_p=2.0e9, del_t=1.0e-10, numThread=4, input_dir=str(tmp_path))
    assert (pinj.addLostPartBool, pinj.refluxPartBool, pinj.injectionBool, pinj.uniformInjectionBool, pinj.heatingBool) == (
        False, True, True, False, False)
    per_thread = 5.6e24 * 1.0e-10 / 2.0e9 / 4.0                         # :232
    assert pinj.numFluxParticlesLow == int(np.floor(per_thread)) and pinj.numFluxParticlesHigh == pinj.numFluxParticlesLow + 1
    assert pinj.injectionR == pytest.approx(per_thread - np.floor(per_thread))
    assert pinj.numFluxPart == pytest.approx(per_thread * 4)
    assert pinj.FractionFreqHeating == 100.0
    # with boundary injection off the uniform-injection line is read as well (:203-206)
    (tmp_path / "P2.inp").write_text("1\n0\n0, 5.6d24\n1, 3.d20 ! uniform\n1, 50.0d0\n")
    pinj.readInjectionInputs("P2.inp", w_p=1.0e9, del_t=1.0e-10, numThread=2, input_dir=str(tmp_path))
    assert (pinj.addLostPartBool, pinj.refluxPartBool, pinj.injectionBool, pinj.uniformInjectionBool, pinj.heatingBool) == (
        True, False, False, True, True)
    assert pinj.injectionFlux == 3.0e20 and pinj.FractionFreqHeating == 50.0
    pinj.writeParticleInjectionInputs(str(tmp_path))
    txt = (tmp_path / "ParticleInjectionInput.dat").read_text().splitlines()
    assert txt[0].startswith("AddLostPartBool") and txt[1].split()[:5] == ["T", "F", "F", "T", "T"]
    assert float(txt[1].split()[5]) == 3.0e20
    pinj.setInjectionInputs()
