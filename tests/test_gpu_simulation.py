"""GPU tests of the device-side diagnostics of the time loop (mod_simulation.f90): the velocity / energy histograms of the final
averaging window against the oracle, and the Python mirror of solveSimulation / solveSimulationFinalAverage running a bounded
plasma with re-injection, refluxing and argon MCC entirely on the device -- files in the reference's layouts, conservation
diagnostics within the reference's own warning thresholds (energyError, gaussError > 1e-5 print a WARNING, :214-224)."""
import os
import numpy as np
import pytest
from oracle import oracle_py as op
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib, particle_injection as pinj, simulation as sim
from common import DECKS, build_oracle, build_cuda_from_oracle, deck_del_t, relerr

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collision_data")
M_AR = 39.948 * op.m_amu


def read_record(path):
    raw = np.fromfile(path, dtype=np.uint8)
    n = int(np.frombuffer(raw[:4].tobytes(), dtype="<i4")[0])
    assert raw.size == n + 8 and int(np.frombuffer(raw[-4:].tobytes(), dtype="<i4")[0]) == n
    return np.frombuffer(raw[4:-4].tobytes(), dtype="<f8")


@pytest.mark.parametrize("name", ["ngp_rf", "cic_neumann"])
def test_velocity_histograms_match_oracle(name, tmp_path):
    d = DECKS[name]
    o = build_oracle(d)
    world, solver, plist = build_cuda_from_oracle(d, o)
    bins = 100                                                   # main.f90:117
    h = sim.Histograms(plist, bins)
    for sp, p in enumerate(plist):
        ex = o.velocityExtrema(sp)                               # VMax, EMax, Emin
        assert h.VMax[sp] == ex[0] and h.EMax[sp] == ex[1] and h.Emin[sp] == ex[2]       # bit for bit: v^2 is summed in the reference's order
    h.accumulate(plist)
    h.accumulate(plist)                                          # the window adds one pass per step
    N = [p.count() for p in plist]
    for sp, p in enumerate(plist):
        VH, EH = np.zeros(2 * bins), np.zeros(bins)
        for _ in range(2):
            E_grid = o.velocityHistograms(sp, bins, h.VMax[sp], h.Emin[sp], h.EMax[sp], VH, EH)
        # weights are rounded to 2^-36 on the device: |difference| <= particles in the bin * 2^-37, far below 1e-9 of the peak
        assert np.max(np.abs(h.VHist[sp] - VH)) < 1e-9 * VH.max()
        assert np.max(np.abs(h.EHist[sp] - EH)) < 1e-9 * EH.max()
        assert relerr(h.E_grid[sp], E_grid) < 1e-14
        # every particle strictly inside the range puts weight 1 on its two bins
        assert abs(h.VHist[sp].sum() - VH.sum()) < 1e-6 and VH.sum() > 2 * (N[sp] - 2) - 1e-6
    h.write(plist, str(tmp_path))
    rec = read_record(os.path.join(str(tmp_path), "Temperature", "Temp_e_average.dat"))
    assert rec.size == 2 * bins + 1 and rec[-1] == h.VMax[0] and np.array_equal(rec[:-1], h.VHist[0])
    rec = read_record(os.path.join(str(tmp_path), "Temperature", "TempEnergy_ion_average.dat"))
    assert rec.size == 2 * bins and np.array_equal(rec[:bins], h.EHist[1])
    assert np.allclose(rec[bins:], h.E_grid[1] * 0.5 * plist[1].mass / op.e, rtol=1e-15)
    # bit-reproducible: integer accumulation
    h2 = sim.Histograms(plist, bins)
    h2.accumulate(plist); h2.accumulate(plist)
    assert np.array_equal(h2.VHist, h.VHist) and np.array_equal(h2.EHist, h.EHist)
    solver.close()


def test_solveSimulation_with_reinjection_and_reflux_on_the_device(tmp_path):
    d = DECKS["ngp_neumann"]
    o = build_oracle(d, factor=2)
    world, solver, plist = build_cuda_from_oracle(d, o)
    del_t = deck_del_t(d)
    pinj.setInjectionSeed(solver, 11)
    pinj.setInjectionInputs(addLost=True, reflux=True)
    st = sim.SimulationState(d["T_e"], d["T_i"], d["n_ave"], d["frac"], numDiagnosticSteps=3, directoryName=str(tmp_path), numThread=d["T"])
    Ne0 = plist[0].count()
    rows = sim.solveSimulation(st, solver, plist, [], [], world, del_t, None, 12.0 * del_t)
    pinj.setInjectionInputs()
    assert len(rows) == 3 and st.steps >= 12
    for r in rows:
        assert r[8] < 1e-5 and r[6] < 1e-5 and r[7] < 1e-5       # energyError, gaussError, chargeError below the reference's WARNING level
        assert r[9] >= 1
    # every lost electron was re-injected (one pair per electron absorbed): the electron count never changes
    assert plist[0].count() == Ne0
    g = open(os.path.join(str(tmp_path), "GlobalDiagnosticData.dat")).read().splitlines()
    assert g[0].startswith("Time (s), Collision Loss (W/m^2)") and len(g) == 4 and len(g[1].split()) == 10
    pd = open(os.path.join(str(tmp_path), "ParticleDiagnostic_e.dat")).read().splitlines()
    assert len(pd) == 4 and int(pd[-1].split()[5]) == Ne0
    for k in range(4):
        assert read_record(os.path.join(str(tmp_path), "Density", "density_e_%d.dat" % k)).size == d["Nx"]
        assert read_record(os.path.join(str(tmp_path), "Phi", "phi_%d.dat" % k)).size == d["Nx"]
        assert read_record(os.path.join(str(tmp_path), "Temperature", "Temp_ion_%d.dat" % k)).size == d["Nx"] - 1
    dens = read_record(os.path.join(str(tmp_path), "Density", "density_e_3.dat"))
    assert np.mean(dens[2:-2]) == pytest.approx(d["n_ave"], rel=0.1)
    assert os.path.exists(os.path.join(str(tmp_path), "SimulationFinalData.dat"))
    # the domain files of writeDomain (mod_simulation.f90:147) and the restart files rewritten at every dump (:229)
    assert read_record(os.path.join(str(tmp_path), "domainGrid.dat")).size == d["Nx"]
    assert read_record(os.path.join(str(tmp_path), "domainDxDl.dat")).size == d["Nx"] - 1
    nv = read_record(os.path.join(str(tmp_path), "domainNodeVol.dat"))
    assert nv.size == d["Nx"] and nv.sum() == pytest.approx(d["L"], rel=1e-12)
    for t in range(1, d["T"] + 1):
        f = os.path.join(str(tmp_path), "PhaseSpace", "phaseSpace_%s_thread%d.dat" % (plist[0].name, t))
        assert os.path.getsize(f) == 32 * int(plist[0].N_p[t - 1])
    solver.close()


def test_final_average_with_argon_collisions_on_the_device(tmp_path):
    """configs[3] in small through the mirror of solveSimulationFinalAverage: Vass argon deck, MCC on the device, averaged density /
    phi, then the EEDF window of one RF period"""
    d = dict(scheme=0, Nx=65, L=0.025 / 4, n_ave=1.0e16, T_e=2.59, T_i=0.025, ppt=10000, T=4, bcL=1, bcR=4, Vl=0.0, Vr=250.0, rf=13.56e6,
             smooth=True, frac=1.5, dt_deck=7.845352413230403e-12, eps_a=1e-4, m=10, beta=0.5, maxIter=100, ion_amu=39.948)
    o = build_oracle(d, factor=2)
    world, solver, plist = build_cuda_from_oracle(d, o)
    plist[0].name, plist[1].name = "e", "Ar+"
    tl = [ip.targetParticle("Ar", M_AR, 2.4e21, 300.0)]
    ncs = ip.readNullCollisionInputs(["Ar_electron_table.dat"], plist, tl, collision_dir=DATA, seed=7)
    del_t = deck_del_t(d)
    st = sim.SimulationState(d["T_e"], d["T_i"], d["n_ave"], d["frac"], directoryName=str(tmp_path), numThread=d["T"])
    # a short RF "period" keeps the test in seconds: the window logic is what is exercised
    solver.RF_rad_frequency = 2.0 * np.pi / (8.0 * del_t)
    steps, hist, dens = sim.solveSimulationFinalAverage(st, solver, plist, tl, ncs, world, del_t, None, 10.0 * del_t, 100)
    assert steps >= 10
    assert sum(int(nc.totalAmountCollisions.sum()) for nc in ncs) > 100
    ave = read_record(os.path.join(str(tmp_path), "Density", "density_e_Average.dat"))
    assert np.array_equal(ave, dens[0]) and np.mean(ave[2:-2]) == pytest.approx(d["n_ave"], rel=0.1)
    assert read_record(os.path.join(str(tmp_path), "Phi", "phi_Average.dat")).size == d["Nx"]
    ga = open(os.path.join(str(tmp_path), "GlobalDiagnosticDataAveraged.dat")).read().splitlines()
    assert int(ga[1].split()[0]) == steps and float(ga[1].split()[2]) > 0.0           # collision loss (W/m^2) of the window
    # gaussError is that of the AVERAGED state (mod_simulation.f90:413-417: rho from the averaged densities, phi_f = the averaged potential)
    assert float(ga[1].split()[5]) == pytest.approx(st.gaussError, rel=1e-7)
    assert 0.0 < st.gaussError < 100.0            # averaged net charge against the averaged potential over 10 steps: O(1) particle noise, not O(1e9)
    ac = open(os.path.join(str(tmp_path), "BinaryCollisions", "e_on_Ar", "AveCollisionDiag.dat")).read().splitlines()   # :401-412
    nc_e = [nc for nc in ncs if nc.reactantsIndx[0] == 1][0]
    assert len(ac) == 1 + nc_e.numberCollisions and len(ac[1].split()) == 6 and int(ac[1].split()[0]) == 1
    # 8 steps of histogram passes; each pass puts every electron inside (Emin, EMax) on two bins
    npass = hist.EHist[0].sum() / plist[0].count()
    assert 7.0 < npass < 9.5
    assert read_record(os.path.join(str(tmp_path), "Temperature", "TempEnergy_e_average.dat")).size == 200
    solver.close()


def test_solveSimulation_with_argon_collisions_writes_the_collision_diagnostics(tmp_path):
    d = dict(scheme=0, Nx=65, L=0.025 / 4, n_ave=1.0e16, T_e=2.59, T_i=0.025, ppt=10000, T=4, bcL=1, bcR=4, Vl=0.0, Vr=250.0, rf=13.56e6,
             smooth=True, frac=1.5, dt_deck=7.845352413230403e-12, eps_a=1e-4, m=10, beta=0.5, maxIter=100, ion_amu=39.948)
    o = build_oracle(d, factor=2)
    world, solver, plist = build_cuda_from_oracle(d, o)
    plist[0].name, plist[1].name = "e", "Ar+"
    tl = [ip.targetParticle("Ar", M_AR, 2.4e21, 300.0)]
    ncs = ip.readNullCollisionInputs(["Ar_electron_table.dat"], plist, tl, collision_dir=DATA, seed=3)
    del_t = deck_del_t(d)
    st = sim.SimulationState(d["T_e"], d["T_i"], d["n_ave"], d["frac"], numDiagnosticSteps=2, directoryName=str(tmp_path), numThread=d["T"])
    rows = sim.solveSimulation(st, solver, plist, tl, ncs, world, del_t, None, 8.0 * del_t)
    assert len(rows) == 2 and all(r[1] > 0.0 for r in rows)                 # collision loss (W/m^2) over each diagnostic interval
    for nc in ncs:
        # the host-side permutation equals the device-side constructor order (mod_NullCollision.f90:55-70)
        assert np.array_equal(nc._type[nc._order()], nc.collisionType)
        assert np.array_equal(nc._thr[nc._order()], nc.energyThreshold)
        cd = os.path.join(str(tmp_path), "BinaryCollisions", "e_on_Ar")
        props = open(os.path.join(cd, "CollisionProperties.dat")).read().splitlines()
        assert len(props) == 1 + nc.numberCollisions and int(props[1].split()[1]) == int(nc.collisionType[0])
        for i in range(nc.numberCollisions):
            lines = open(os.path.join(cd, "CollisionDiag_%d.dat" % (i + 1))).read().splitlines()
            assert len(lines) == 3 and len(lines[1].split()) == 5
        break
    solver.close()
