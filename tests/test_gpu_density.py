"""GPU parity of the density diagnostic (ipic_load_particle_density + the mirror of WriteParticleDensity) against the oracle:
the arrays the reference writes to Density/density_<name>_<step>.dat, within 1e-12 relative, on all wall / smoothing / grid
combinations of the test decks, after a move (particles at the walls, ragged slabs)."""
import os
import numpy as np
import pytest
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import scheme as sch
from common import DECKS, build_oracle, build_cuda_from_oracle, deck_del_t, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(DECKS))
def test_WriteParticleDensity_matches_oracle(name, tmp_path):
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o)
    del_t = deck_del_t(d) * 2.0
    solver.phi, solver.phi_f, solver.rho = o.get("phi"), o.get("phi_f"), o.get("rho")
    o.makeHalfTimeEField(); solver.makeHalfTimeEField()
    for step in range(2):
        o.loadParticleDensity(True)
        sch.loadParticleDensity(plist, world, True)
        out = sch.WriteParticleDensity(plist, world, step + 1, False, str(tmp_path))
        for sp, p in enumerate(plist):
            want = o.particleDensityArray(sp)
            assert relerr(out[sp], want) < 1e-12, (step, sp)
            raw = np.fromfile(os.path.join(str(tmp_path), "Density", "density_%s_%d.dat" % (p.name, step + 1)), dtype=np.uint8)
            assert raw.size == 8 + 8 * d["Nx"]                                     # Fortran unformatted record: marker, N_x doubles, marker
            assert np.array_equal(np.frombuffer(raw[4:-4].tobytes(), dtype="<f8"), out[sp])
            assert int(np.frombuffer(raw[:4].tobytes(), dtype="<i4")[0]) == 8 * d["Nx"]
        o.moveParticles(del_t)
        ip.moveParticles(solver, plist, world, del_t)
    # time average: loads without reset add up (loadParticleDensity(..., reset = .false.), mod_simulation.f90 averaging loop)
    sch.loadParticleDensity(plist, world, True)
    one = [p.densities.copy() for p in plist]
    sch.loadParticleDensity(plist, world, False)
    for p, a in zip(plist, one):
        assert np.array_equal(p.densities, 2.0 * a)
    # the diagnostic leaves the deposition grids clean: rho afterwards is what it was
    solver.depositRho(plist, world)
    o.depositRho()
    assert relerr(solver.rho, o.get("rho")) < 1e-12
    solver.close()
