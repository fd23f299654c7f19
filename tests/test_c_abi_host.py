"""A compiled C host (tests/c_host/host_driver.c) drives the C ABI in the order of the Fortran shim
(fortran/mod_particleMover_b200.f90) and is checked against the CPU oracle.  Stands in for the Fortran host,
which cannot be compiled in this image."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib
from common import DECKS, build_oracle, deck_del_t, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_host", "host_driver.c")


def _compile(tmp_path):
    exe = str(tmp_path / "host_driver")
    libdir = os.path.dirname(ip.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-o", exe, SRC, "-L" + libdir, "-lipic1d_b200", "-Wl,-rpath," + libdir])
    return exe


def test_c_host_compiles_and_links_against_header(tmp_path):
    """no GPU needed: the header is valid C11 and every symbol the driver uses resolves at link time"""
    ip.load_library()
    exe = _compile(tmp_path)
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 2          # usage error, before any CUDA call


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ngp_rf", "cic_rf", "ngp_periodic"])
def test_c_host_matches_oracle(tmp_path, name):
    d = DECKS[name]
    o = build_oracle(d)
    del_t = deck_del_t(d)
    g, dx, bc = o.grid()
    nsp, T = o.n_species, d["T"]
    finalIdx = o.phase_space(0).shape[1]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([d["scheme"], d["Nx"], int(d["smooth"]), nsp, T, finalIdx], dtype=np.int32).tofile(f)
        np.array([del_t, d["Vl"], d["Vr"], d["rf"]], dtype=np.float64).tofile(f)
        g.tofile(f); dx.tofile(f); bc.astype(np.int32).tofile(f)
        o.get("phi").tofile(f); o.get("phi_f").tofile(f)
        for sp in range(nsp):
            sc = o.species_scalars(sp)
            np.array([sc["mass"], sc["q"], sc["w_p"]]).tofile(f)
            o.N_p(sp).astype(np.int32).tofile(f)
            np.ascontiguousarray(o.phase_space(sp)).tofile(f)
    exe = _compile(tmp_path)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    o.depositJ(del_t)
    Jo = o.get("J")
    o.moveParticles(del_t)
    with open(fout, "rb") as f:
        Jg = np.fromfile(f, dtype=np.float64, count=o.Nh)
        assert relerr(Jg, Jo) < 1e-12
        for sp in range(nsp):
            Np = np.fromfile(f, dtype=np.int32, count=T)
            ps = np.fromfile(f, dtype=np.float64, count=T * finalIdx * 4).reshape(T, finalIdx, 4)
            st = _lib.MoveStats.from_buffer_copy(f.read(C.sizeof(_lib.MoveStats)))
            assert np.array_equal(Np, o.N_p(sp))
            pso = o.phase_space(sp)
            for t in range(T):
                n = Np[t]
                assert np.max(np.abs(ps[t, :n, 0] - pso[t, :n, 0])) < 1e-9
                np.testing.assert_allclose(ps[t, :n, 1], pso[t, :n, 1], rtol=1e-9, atol=1e-6)
                assert np.array_equal(ps[t, :n, 2:], pso[t, :n, 2:])
            c = o.counters(sp)
            assert st.delIdx == int(c["delIdx"].sum()) and st.refIdx == int(c["refIdx"].sum())
            assert st.N_p == int(o.N_p(sp).sum())


@pytest.mark.gpu
def test_c_host_level3_keeps_the_particles_on_the_device(tmp_path):
    """the same compiled host with the level-3 calls between move and download: re-injection of the lost electrons,
    density load and EEDF extrema through the C ABI -- one pair appended per absorbed electron, density integrating to the particle
    number, extrema bracketing the downloaded velocities"""
    d = DECKS["ngp_rf"]
    o = build_oracle(d, factor=2)
    del_t = deck_del_t(d)
    g, dx, bc = o.grid()
    nsp, T = o.n_species, d["T"]
    finalIdx = o.phase_space(0).shape[1]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([d["scheme"], d["Nx"], int(d["smooth"]), nsp, T, finalIdx], dtype=np.int32).tofile(f)
        np.array([del_t, d["Vl"], d["Vr"], d["rf"]], dtype=np.float64).tofile(f)
        g.tofile(f); dx.tofile(f); bc.astype(np.int32).tofile(f)
        o.get("phi").tofile(f); o.get("phi_f").tofile(f)
        for sp in range(nsp):
            sc = o.species_scalars(sp)
            np.array([sc["mass"], sc["q"], sc["w_p"]]).tofile(f)
            o.N_p(sp).astype(np.int32).tofile(f)
            np.ascontiguousarray(o.phase_space(sp)).tofile(f)
    exe = _compile(tmp_path)
    r = subprocess.run([exe, fin, fout, "level3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    o.moveParticles(del_t)
    lost_e = int(o.counters(0)["delIdx"].sum())
    with open(fout, "rb") as f:
        np.fromfile(f, dtype=np.float64, count=o.Nh)
        Np, ps = [], []
        for sp in range(nsp):
            Np.append(np.fromfile(f, dtype=np.int32, count=T))
            ps.append(np.fromfile(f, dtype=np.float64, count=T * finalIdx * 4).reshape(T, finalIdx, 4))
            f.read(C.sizeof(_lib.MoveStats))
        numAdded = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        energyAdd = float(np.fromfile(f, dtype=np.float64, count=1)[0])
        dens = np.fromfile(f, dtype=np.float64, count=d["Nx"])
        extrema = np.fromfile(f, dtype=np.float64, count=3)
    assert numAdded == lost_e > 0 and energyAdd > 0.0
    for sp in range(nsp):
        assert int(Np[sp].sum()) == int(o.N_p(sp).sum()) + numAdded            # the pairs sit behind the last slab
    assert dens.sum() == pytest.approx(float(Np[0].sum()), rel=1e-9)             # NGP weights (1 - d, d): every electron counts once
    ve = np.concatenate([ps[0][t, :Np[0][t], 1:4] for t in range(T)])
    assert extrema[0] == np.max(np.abs(ve[:, 0]))
    v2 = np.sum(ve ** 2, axis=1)
    assert extrema[1] == pytest.approx(v2.max(), rel=1e-15) and extrema[2] <= v2.min() * (1 + 1e-15)
