"""Pins on the only outputs of the REAL reference binary available here: the run logs it ships next to its executables
(OpenMP/NGP/src/output.txt, OpenMP/CIC/src/output.txt -> tests/golden/reference_run_logs.json, made by
tests/golden/make_reference_log_fixture.py).  What the Fortran program printed at start-up for its default He RF decks must come out
of the restated host logic, the oracle and the table reader to the digits printed (list-directed output: 15-16 significant digits):
time step / plasma frequency, grid spacing, species masses and weights, RF constants, and the two null-collision sets it built from
PIC1D/CollisionData (reduced masses, merged energy-axis length, max sigma*v, constructor order of the collisions)."""
import json
import os
import numpy as np
import pytest
from oracle import oracle_py as op, mcc_py as mp
from implicitpic1d_b200 import constants as K
from implicitpic1d_b200.domain import readWorld
from implicitpic1d_b200.null_collision import readNullCollisionInputs, targetParticle
from implicitpic1d_b200.particle import Particle

HERE = os.path.dirname(os.path.abspath(__file__))
LOGS = json.load(open(os.path.join(HERE, "golden", "reference_run_logs.json")))
DATA = os.path.join(HERE, "golden", "collision_data")
TABLE = "turner_benchmark_he_electron_table.dat"
P15 = 2e-15          # a value printed with 15-16 significant digits


@pytest.mark.parametrize("tree", ["ngp", "cic", "ngp_2001"])
def test_startup_constants_match_the_reference_log(tree):
    g = LOGS[tree]
    # constants.f90:106-117: plasmaFreqTemp = SQRT(n_ave e^2 / m_e / eps_0); the log prints del_t * plasmaFreqTemp
    assert g["del_t"] * K.getPlasmaFreq(g["n_ave"]) == pytest.approx(g["fraction_of_inverse_plasma_frequency"], rel=5e-15)
    assert K.del_t_from_deck(g["n_ave"], 10.0, g["del_t"]) == g["del_t"]
    assert 2.0 * K.pi * g["RF_frequency"] == pytest.approx(g["RF_rad_frequency"], rel=5e-15)
    # species: He+ = 4.002602 amu - m_e (mod_particle.f90:334), target He = 4.002602 amu; w_p = n_ave L / N (mod_particle.f90:76-82)
    e, ion = g["particles"]
    assert K.m_e == e["mass"] and -K.e == e["charge"]
    assert 4.002602 * K.m_amu - K.m_e == pytest.approx(ion["mass"], rel=P15)
    assert 4.002602 * K.m_amu == pytest.approx(g["target"]["mass"], rel=P15)
    for p in (e, ion):
        assert g["n_ave"] * g["L_domain"] / p["macroparticles"] == pytest.approx(p["weight"], rel=5e-15)
    # the loader's Maxwellian: mean kinetic energy 1.5 T within 5 sigma of the sample (Var[KE] = 1.5 T^2), log and oracle alike
    for p, T in ((e, g["T_e"]), (ion, g["T_i"])):
        assert abs(p["mean_KE_eV"] - 1.5 * T) < 5.0 * np.sqrt(1.5) * T / np.sqrt(p["macroparticles"])


@pytest.mark.parametrize("tree", ["ngp", "cic", "ngp_2001"])
def test_grid_spacing_matches_the_reference_log(tree):
    """'smallest delX' = MINVAL(dx_dl) of the uniform grid (NGP mod_domain.f90:104-110, CIC :103-110): the rounding of the node positions
    shows in the 13th digit (6.699999999999762E-005, not 6.7E-005), so this pins the construction order of the grid, not just L / N"""
    g = LOGS[tree]
    scheme = 1 if tree == "cic" else 0
    world = readWorld(scheme, g["NumberXNodes"], g["L_domain"], g["gridType"], 0.0, g["bcL"], g["bcR"], 1, g["T_e"], g["n_ave"])
    assert world.NumberXHalfNodes == g["NumberXHalfNodes"]
    assert float(np.min(world.dx_dl)) == pytest.approx(g["smallest_delX"], rel=P15)
    o = op.Oracle(scheme, g["NumberXNodes"], g["L_domain"], gridType=g["gridType"], bcL=g["bcL"], bcR=g["bcR"], smooth=True, T=1)
    grid, dx, bc = o.grid()
    assert float(np.min(dx[:world.dx_dl.size])) == pytest.approx(g["smallest_delX"], rel=P15)
    assert np.array_equal(dx[:world.dx_dl.size], world.dx_dl) and np.array_equal(grid[:world.grid.size], world.grid)


def test_null_collision_sets_match_the_reference_log():
    """readNullCollisionInputs + nullCollision_constructor (mod_NullCollision.f90:44-108, 666-939) on the He table of the default decks
    (turner_benchmark_he_electron_table.dat holds the e + He and the He+ + He reactions): two reactant pairs, merged energy axes of 773
    and 101 points, reduced masses, max sigma*v and the constructor's collision order as the Fortran program printed them"""
    g = LOGS["ngp"]
    assert LOGS["cic"]["nullCollisions"] == g["nullCollisions"] == LOGS["ngp_2001"]["nullCollisions"]      # every logged run prints the same sets
    m_ion, m_He = 4.002602 * K.m_amu - K.m_e, 4.002602 * K.m_amu
    # (1) the oracle's reader + constructor
    o = op.Oracle(op.NGP, 33, 0.05, T=1)
    for m, q in ((K.m_e, -K.e), (m_ion, K.e)):
        sp = o.add_species(m, q, 8, 2)
        o.set_weight(sp, 1.0)
    mp.add_target(o, m_He, g["target"]["density"], g["target"]["temperature"])
    specs = mp.read_null_collision_inputs([os.path.join(DATA, TABLE)], [("e", K.m_e), ("He+", m_ion)], [("He", m_He)])
    assert len(specs) == len(g["nullCollisions"]) == 2
    for spec, ref in zip(specs, g["nullCollisions"]):
        nc = mp.NullCollisionOracle(o, spec)
        sc, tb = nc.scalars(), nc.tables()
        assert int(spec["numberCollisions"]) == ref["numberCollisions"] and int(spec["lengthArrays"]) == ref["lengthArrays"]
        assert sc["reducedMass"] == pytest.approx(ref["reducedMass"], rel=P15)
        assert sc["reducedMassIonization"] == pytest.approx(ref["reducedMassIonization"], rel=P15, abs=0.0)
        assert sc["sigmaVMax"] == pytest.approx(ref["sigmaVMax"], rel=5e-15)
        assert tb["collisionType"].tolist() == [c["collisionType"] for c in ref["collisions"]]
        assert tb["energyThreshold"].tolist() == [c["energyThreshold"] for c in ref["collisions"]]
    # (2) the product's host-side reader (implicitpic1d_b200/null_collision.py; the constructor's sort is mirrored by _order())
    plist = [Particle(K.m_e, -K.e, 1.0, 0, 8, "e", 1), Particle(m_ion, K.e, 1.0, 0, 8, "He+", 1)]
    tl = [targetParticle("He", m_He, g["target"]["density"], g["target"]["temperature"])]
    ncs = readNullCollisionInputs([TABLE], plist, tl, collision_dir=DATA)
    assert len(ncs) == 2
    for nc, ref in zip(ncs, g["nullCollisions"]):
        assert nc.numberCollisions == ref["numberCollisions"] and nc.lengthArrays == ref["lengthArrays"]
        assert nc.reducedMass == pytest.approx(ref["reducedMass"], rel=P15)
        assert nc.reducedMassIonization == pytest.approx(ref["reducedMassIonization"], rel=P15, abs=0.0)
        order = nc._order()
        assert nc._type[order].tolist() == [c["collisionType"] for c in ref["collisions"]]
        assert nc._thr[order].tolist() == [c["energyThreshold"] for c in ref["collisions"]]
        v = np.sqrt(2.0 * nc._energy * K.e / nc.reducedMass)
        assert float(np.max(np.sum(nc._sigma, axis=1) * v)) == pytest.approx(ref["sigmaVMax"], rel=5e-15)      # :88-97
