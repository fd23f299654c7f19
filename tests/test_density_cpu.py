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
