"""CPU tests of the oracle's restatement of the density diagnostic (loadParticleDensity + WriteParticleDensity, NGP
mod_Scheme.f90:68-158, CIC mod_Scheme.f90:68-185).  The reference holds no golden vector for it; the pin is what the routine
states: the written density integrates to the number of physical particles (every interpolation weight is folded back into the
domain, the binomial smoothing and its wall stencils conserve the sum on a uniform grid), and a uniform load gives n_ave."""
import numpy as np
import pytest
from common import DECKS, build_oracle


def node_volumes(o, d):
    g, dx, bc = o.grid()
    Nx = d["Nx"]
    if d["scheme"] == 1:
        return dx[:Nx].copy()                                   # CIC: densities / dx_dl (:163)
    vol = np.zeros(Nx)                                          # NGP: 2 / (dx(j-1) + dx(j)), half cells at the walls (:130-145)
    vol[1:Nx - 1] = 0.5 * (dx[:Nx - 2] + dx[1:Nx - 1])
    if d["bcL"] == 3:
        vol[0] = vol[Nx - 1] = 0.5 * (dx[0] + dx[Nx - 2])
    else:
        vol[0], vol[Nx - 1] = 0.5 * dx[0], 0.5 * dx[Nx - 2]
    return vol


@pytest.mark.parametrize("name", sorted(DECKS))
def test_density_integrates_to_the_particle_number(name):
    d = DECKS[name]
    o = build_oracle(d)
    o.loadParticleDensity(True)
    for sp in range(2):
        w_p = o.species_scalars(sp)["w_p"]
        N = int(o.N_p(sp).sum())
        dens = o.particleDensityArray(sp)
        vol = node_volumes(o, d)
        total = float(np.sum(dens * vol))
        if d["scheme"] == 0 and d["bcL"] == 3:
            total -= dens[-1] * vol[-1]                         # node N_x repeats node 1 on a periodic NGP grid (:146)
        uniform = d.get("gridType", 0) == 0
        if uniform or not d["smooth"]:
            assert total == pytest.approx(N * w_p, rel=1e-12)
        # a uniform load: the mean density is n_ave (w_p = n_ave L / N, mod_particle.f90:76-82)
        assert np.mean(dens[2:-2]) == pytest.approx(d["n_ave"], rel=0.05)


def test_density_accumulates_without_reset():
    d = DECKS["cic_rf"]
    o = build_oracle(d)
    o.loadParticleDensity(True)
    one = o.particleDensityArray(0)
    o.loadParticleDensity(True)
    o.loadParticleDensity(False)
    two = o.particleDensityArray(0)
    assert np.allclose(two, 2.0 * one, rtol=1e-13, atol=0.0)


def test_oracle_velocity_histograms_known_answers():
    """mod_simulation.f90:433-500 restated: every particle strictly inside a range puts total weight 1 on two neighbouring bins;
    the v_x histogram of a Maxwellian load is symmetric and its second moment is v_th^2; the energy grid is logarithmic from
    Emin to EMax; Emin never exceeds 0.1 eV (0.1 e 2 / m)"""
    from oracle import oracle_py as op
    d = DECKS["ngp_rf"]
    o = build_oracle(d)
    bins = 100
    VMax, EMax, Emin = o.velocityExtrema(0)
    m = o.species_scalars(0)["mass"]
    assert Emin <= 0.1 * op.e * 2.0 / m and EMax > Emin and VMax > 0
    VH, EH = np.zeros(2 * bins), np.zeros(bins)
    E_grid = o.velocityHistograms(0, bins, VMax, Emin, EMax, VH, EH)
    N = int(o.N_p(0).sum())
    assert N - 2 <= VH.sum() <= N and N - 2 <= EH.sum() <= N              # the extreme particles sit on the open ends
    assert E_grid[0] == np.exp(np.log(Emin)) and E_grid[-1] == pytest.approx(EMax, rel=1e-12)
    assert np.allclose(np.diff(np.log(E_grid)), (np.log(EMax) - np.log(Emin)) / (bins - 1), rtol=1e-10)
    centers = (np.arange(1, 2 * bins + 1) - bins - 0.5) * 2.0 * VMax / (2.0 * bins - 1)     # v_x of bin k (inverse of :483)
    mean = np.sum(VH * centers) / VH.sum()
    var = np.sum(VH * centers ** 2) / VH.sum()
    vth2 = d["T_e"] * op.e / m
    assert abs(mean) < 5.0 * np.sqrt(vth2 / N)
    assert var / vth2 == pytest.approx(1.0, abs=0.03)


def _raw_density_weights(o, d, sp):
    """SUM(densities, DIM=2) of loadParticleDensity in plain numpy / Python from the oracle's particle positions: NGP (1 - d, d) onto the
    two nodes (mod_Scheme.f90:50-65), CIC the quadratic spline with the weight behind a wall folded into the wall cell (:68-105)"""
    Nx = d["Nx"]
    ps, Np = o.phase_space(sp), o.N_p(sp)
    l = np.concatenate([ps[t, :Np[t], 0] for t in range(o.T)])
    raw = np.zeros(Nx + 2)
    lc = l.astype(np.int64)
    f = l - lc
    if d["scheme"] == 0:
        np.add.at(raw, lc, 1.0 - f)
        np.add.at(raw, lc + 1, f)
        return raw[1:Nx + 1]
    bc = o.grid()[2]                                            # boundaryConditions(1:N_h), 0-based array
    for c, x in zip(lc, f):
        raw[c] += -x * x + x + 0.5
        wl, wr = 0.5 * (1.0 - x) ** 2, 0.5 * x * x
        b = int(bc[c]) - int(bc[c - 1])                         # bc(l_center + 1) - bc(l_center)
        if b == 0:
            raw[c - 1] += wl; raw[c + 1] += wr
        elif b < 0 and b != -3:
            raw[c] += wl; raw[c + 1] += wr
        elif b > 0 and b != 3:
            raw[c] += wr; raw[c - 1] += wl
        elif b == -3:
            raw[Nx] += wl; raw[c + 1] += wr
        else:
            raw[c - 1] += wl; raw[1] += wr
    return raw[1:Nx + 1]


@pytest.mark.parametrize("name", sorted(DECKS))
def test_host_epilogue_of_WriteParticleDensity_matches_the_oracle(name, tmp_path):
    """implicitpic1d_b200/scheme.py (host logic of the product: weight, periodic fold, smoothing, normalisation, record writer) on raw
    weights computed independently in numpy, against the oracle's C++ restatement of the whole routine"""
    from implicitpic1d_b200 import scheme as sch
    from implicitpic1d_b200.domain import readWorld
    from implicitpic1d_b200.particle import Particle
    d = DECKS[name]
    o = build_oracle(d)
    world = readWorld(d["scheme"], d["Nx"], d["L"], d.get("gridType", 0), d.get("del_x", 0.0), d["bcL"], d["bcR"], int(d["smooth"]), d["T_e"], d["n_ave"])
    o.loadParticleDensity(True)
    plist = []
    for sp in range(2):
        sc = o.species_scalars(sp)
        p = Particle(sc["mass"], sc["q"], sc["w_p"], 0, 8, "e" if sp == 0 else "ion", d["T"])
        p.densities = _raw_density_weights(o, d, sp)
        plist.append(p)
    out = sch.WriteParticleDensity(plist, world, 3, False, str(tmp_path))
    for sp in range(2):
        want = o.particleDensityArray(sp)
        assert np.max(np.abs(out[sp] - want)) <= 1e-12 * np.max(np.abs(want))
        raw = np.fromfile(str(tmp_path / "Density" / ("density_%s_3.dat" % plist[sp].name)), dtype=np.uint8)
        assert raw.size == 8 * d["Nx"] + 8 and np.array_equal(np.frombuffer(raw[4:-4].tobytes(), dtype="<f8"), out[sp])


def test_host_energy_grid_and_record_writer_match_the_oracle(tmp_path):
    """simulation.energy_grid (host logic: mod_simulation.f90:453-461) against the oracle's E_grid, and the unformatted-record writer"""
    from implicitpic1d_b200 import simulation as sim
    d = DECKS["cic_rf"]
    o = build_oracle(d)
    VMax, EMax, Emin = o.velocityExtrema(1)
    VH, EH = np.zeros(200), np.zeros(100)
    E_grid = o.velocityHistograms(1, 100, VMax, Emin, EMax, VH, EH)
    diffE, grid = sim.energy_grid(Emin, EMax, 100)
    assert np.array_equal(grid, E_grid)
    sim._record(str(tmp_path / "r.dat"), VH, np.array([VMax]))
    raw = np.fromfile(str(tmp_path / "r.dat"), dtype=np.uint8)
    assert raw.size == 8 * 201 + 8 and int(np.frombuffer(raw[:4].tobytes(), dtype="<i4")[0]) == 8 * 201
    assert np.array_equal(np.frombuffer(raw[4:-4].tobytes(), dtype="<f8"), np.concatenate([VH, [VMax]]))
    assert sim.es(1.6e18) == "  1.60000000E+18" and sim.es(-2.5e-11) == " -2.50000000E-11"      # Fortran es16.8
