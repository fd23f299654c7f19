"""CPU tests of the null-collision MCC: the oracle restatement (oracle/ipic_oracle.cpp, oracle/mcc_py.py) against the reference's
own known answer and conservation laws, the product's table reader against the oracle's, and the host-side candidate
permutation of the CUDA library.  No GPU needed."""
import ctypes as C
import os
import numpy as np
import pytest
from oracle import oracle_py as op, mcc_py as mp
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib, constants as K

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collision_data")
M_HE = 4.002602 * op.m_amu
M_AR = 39.948 * op.m_amu


def he_swarm(T=8, ppt=2000, table="turner_benchmark_he_electron_table_noIonization.dat"):
    """program state of OpenMP-BField/TestMCC/MCCExample.f90:23-30: electrons at 2.25 * 2/3 eV, He at 300 K, 1e21 m^-3"""
    o = op.Oracle(op.NGP, 33, 1.0, T=T)
    sp = o.add_species(op.m_e, -op.e, ppt, 2)
    o.set_weight(sp, 1.0)
    seeds = np.array([123456 * i * 11 for i in range(1, T + 1)], dtype=np.int32)         # MCCExample.f90:24-26
    o.load_particles(sp, seeds, 2.25 * 2.0 / 3.0)
    o.add_species(M_HE - op.m_e, op.e, 1, 2)
    mp.add_target(o, M_HE, 1e21, 300.0)
    sets = mp.read_null_collision_inputs([os.path.join(DATA, table)], [("e", op.m_e), ("He+", M_HE - op.m_e)], [("He", M_HE)])
    return o, sp, seeds, mp.NullCollisionOracle(o, sets[0])


def test_oracle_reproduces_bolsig_collision_frequency():
    """The reference's known-answer test (OpenMP-BField/TestMCC/MCCExample.f90:13,32-98): an electron swarm in He at E/N = 25 Td
    relaxes under `v_x += q E dt / m ; generateCollision` and its collision frequency is compared with BOLSIG's
    0.7733e-13 m^3/s * n_gas.  Shortened: nu*dt = 0.02 instead of 0.01 and 2 x 50 000 steps instead of 2 x 100 000; the
    null-collision count then sits (1 - exp(-x))/x = 1.1 % below nu at x = sigmaVMax n dt."""
    o, sp, seeds, nc = he_swarm()
    bolsig, n_gas = 0.7733e-13, 1e21
    del_t = 2e-2 / (bolsig * n_gas)
    accel = -op.e * (25.0 * 1e-21 * n_gas) * del_t / op.m_e
    ps = o.phase_space(sp)
    n = int(o.N_p(sp)[0])
    steps = 50000
    for phase in range(2):
        nc.reset_totals()
        for _ in range(steps):
            ps[:, :n, 1] += accel
            mp.set_num_to_collide(o, sp)
            nc.generateCollision(seeds, del_t)
    tot = nc.totals()
    freq = tot["totalAmountCollisions"].sum() / steps / del_t / o.N_p(sp).sum()
    x = nc.scalars()["sigmaVMax"] * n_gas * del_t
    assert abs(nc.scalars()["P_null"] - (1 - np.exp(-x))) < 1e-15
    corrected = freq * x / (1 - np.exp(-x))
    assert abs(corrected / (bolsig * n_gas) - 1.0) < 0.02, (freq, corrected, bolsig * n_gas)
    assert 6.5 < o.getKEAve(sp) < 9.0            # mean energy of the relaxed swarm (eV)


def test_oracle_collisions_conserve_what_they_should():
    """elastic + excitation on Ar: momentum of (electron + target) is conserved by construction, the tallied energy loss equals
    the kinetic energy the electrons actually lost, ionization appends electron + ion at the parent's position"""
    T, ppt = 4, 20000
    o = op.Oracle(op.NGP, 33, 0.05, T=T)
    e = o.add_species(op.m_e, -op.e, ppt, 2)
    o.set_weight(e, 1.0)
    seeds = op.default_seeds(T)
    o.load_particles(e, seeds, 20.0)                                 # hot enough to excite and ionize
    ion = o.add_species(M_AR - op.m_e, op.e, ppt, 2)
    o.set_weight(ion, 1.0)
    o.load_particles(ion, seeds, 0.03)
    mp.add_target(o, M_AR, 2.4e21, 300.0)
    sets = mp.read_null_collision_inputs([os.path.join(DATA, "Ar_electron_table.dat")], [("e", op.m_e), ("Ar+", M_AR - op.m_e)], [("Ar", M_AR)])
    assert [s["numberCollisions"] for s in sets] == [3, 2]
    ncs = [mp.NullCollisionOracle(o, s) for s in sets]
    tb = ncs[0].tables()
    assert sorted(tb["collisionType"].tolist()) == [1, 2, 3]
    ke0 = sum(o.getTotalKE(s) for s in (e, ion))
    n0 = (o.N_p(e).sum(), o.N_p(ion).sum())
    sumv2_e0 = o.getTotalKE(e) / (0.5 * op.m_e)
    for s in (e, ion):
        mp.set_num_to_collide(o, s)
    del_t = 2e-10
    ncs[0].generateCollision(seeds, del_t)
    tot = ncs[0].totals()
    n_ion = tot["totalAmountCollisions"][tb["collisionType"] == 2].sum()
    assert n_ion > 10
    assert o.N_p(e).sum() - n0[0] == n_ion and o.N_p(ion).sum() - n0[1] == n_ion
    # energyLoss tallies are m v^2 sums (:208,216): electrons' lost (1/2 m v^2) * 2, ionization rows also subtract the new pair
    sumv2_e1 = o.getTotalKE(e) / (0.5 * op.m_e)
    new_ion_v2 = 0.0
    for t in range(T):
        a, b = n0[1] // T, o.N_p(ion)[t]
        new_ion_v2 += (o.phase_space(ion)[t, a:b, 1:4] ** 2).sum()
        # appended pairs sit at the same location
        ea, eb = n0[0] // T, o.N_p(e)[t]
        assert np.array_equal(o.phase_space(e)[t, ea:eb, 0], o.phase_space(ion)[t, a:b, 0])
    expect = op.m_e * (sumv2_e0 - sumv2_e1) - new_ion_v2 * M_AR
    assert abs(tot["totalEnergyLoss"].sum() - expect) <= 1e-9 * abs(expect)
    # per collision the loss is the threshold (excitation 11.5 eV, ionization 15.8 eV) up to the recoil of the heavy target
    for c in range(3):
        if tb["collisionType"][c] in (2, 3) and tot["totalAmountCollisions"][c] > 0:
            per = 0.5 * tot["totalEnergyLoss"][c] / tot["totalAmountCollisions"][c] / op.e
            assert abs(per - tb["energyThreshold"][c]) < 0.05 * tb["energyThreshold"][c], (c, per)
    assert abs(sum(o.getTotalKE(s) for s in (e, ion)) - ke0) / ke0 < 0.5


def test_product_reader_matches_oracle_reader():
    """implicitpic1d_b200.readNullCollisionInputs (numpy) against the loop-by-loop restatement of the Fortran reader"""
    for table, ion_name, m_gas, gas in (("Ar_electron_table.dat", "Ar+", M_AR, "Ar"), ("turner_benchmark_he_electron_table.dat", "He+", M_HE, "He")):
        charged = [("e", op.m_e), (ion_name, m_gas - op.m_e)]
        ref = mp.read_null_collision_inputs([os.path.join(DATA, table)], charged, [(gas, m_gas)])
        plist = [ip.Particle(m, 1.0, 1.0, 0, 0, nm, 1) for nm, m in charged]
        tl = [ip.targetParticle(gas, m_gas, 1e21, 300.0)]
        got = ip.readNullCollisionInputs([table], plist, tl, collision_dir=DATA)
        assert len(got) == len(ref)
        for g, r in zip(got, ref):
            assert (g.numberCollisions, g.lengthArrays) == (r["numberCollisions"], r["lengthArrays"])
            assert g.reactantsIndx.tolist() == [r["primary"], r["target"]]
            assert np.array_equal(g._energy, r["energyArray"])
            assert np.allclose(g._sigma, r["sigmaArray"], rtol=1e-14, atol=0.0)
            assert np.array_equal(g._thr, r["energyThreshold"]) and np.array_equal(g._type, r["collisionType"])
            assert np.array_equal(g._pidx.T, r["productsIndx"])
            assert g.reducedMass == r["red_mass"] and g.sumMass == r["sumMass"]
            assert abs(g.reducedMassIonization - r["redMassTripleProducts"]) <= 1e-15 * max(r["redMassTripleProducts"], 1e-300)


def test_collision_list_file_and_missing_table(tmp_path):
    """collision.inp lists table names up to END; names that do not exist are skipped as in the reference (:694-695)"""
    p = tmp_path / "collision.inp"
    p.write_text("Ar_electron_table.dat\nnot_there.dat\nEND\nignored.dat\n")
    plist = [ip.Particle(op.m_e, -op.e, 1.0, 0, 0, "e", 1), ip.Particle(M_AR - op.m_e, op.e, 1.0, 0, 0, "Ar+", 1)]
    tl = [ip.targetParticle("Ar", M_AR, 2.4e21, 300.0)]
    got = ip.readNullCollisionInputs("collision.inp", plist, tl, collision_dir=DATA, input_dir=str(tmp_path))
    assert [g.numberCollisions for g in got] == [3, 2]
    with pytest.raises(ValueError):
        ip.readNullCollisionInputs(["Ar_electron_table.dat"], plist[:1], tl, collision_dir=DATA)     # Ar+ unknown


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 1000, 4097, 65536, 100003])
def test_candidate_permutation_is_a_bijection(n):
    L = ip.load_library()
    out = np.zeros(n, dtype=np.uint64)
    _lib.check(L.ipic_mcc_permutation(n, 0x1234ABCD + n, 0, n, out.ctypes.data_as(C.POINTER(C.c_uint64))))
    assert np.array_equal(np.sort(out), np.arange(n, dtype=np.uint64))


def test_candidate_permutation_prefix_is_uniform():
    """the first k images (the collision candidates) are spread uniformly over the (cell-sorted) particle array, for any key"""
    L = ip.load_library()
    n, k, bins = 1_000_003, 20000, 50
    worst = 0.0
    for key in range(20):
        out = np.zeros(k, dtype=np.uint64)
        _lib.check(L.ipic_mcc_permutation(n, key, 0, k, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        h, _ = np.histogram(out.astype(np.float64), bins=bins, range=(0, n))
        chi2 = ((h - k / bins) ** 2 / (k / bins)).sum()
        worst = max(worst, chi2)
        assert chi2 < 110.0, (key, chi2)          # 49 dof: mean 49, sigma ~10
        # successive candidates are not neighbours in the array
        assert np.mean(np.abs(np.diff(out.astype(np.int64))) < n // 1000) < 0.01
    with pytest.raises(_lib.IpicError):
        _lib.check(L.ipic_mcc_permutation(10, 0, 5, 6, out.ctypes.data_as(C.POINTER(C.c_uint64))))
