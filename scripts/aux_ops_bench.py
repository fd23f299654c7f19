"""Times the between-step particle operations at the C5 shard's size on one B200 with CUDA events:
cell sort, density load, local temperature, KE / momentum reduction, velocity extrema + histograms, Maxwellian heating,
re-injection of lost particles -- each with its algorithmic bytes and the fraction of the measured HBM peak.
usage (GPU box): python scripts/aux_ops_bench.py [particles_per_species] > profiles/rN_aux_ops_c5.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib, constants as K, particle_injection as pinj, simulation as sim
from implicitpic1d_b200.domain import readWorld
import bench


def main():
    n_sp = int(sys.argv[1]) if len(sys.argv) > 1 else 62_500_000
    w = bench.WORKLOADS["c5"]
    peak = bench.measured_peak_gbs()[0]
    L = ip.load_library()
    world = readWorld(w["scheme"], w["Nx"], w["L"], 0, 0.0, w["bcL"], w["bcR"], w["smooth"], w["T_e"], w["n_ave"])
    solver = ip.potentialSolver(world, w["Vl"], w["Vr"], w["rf"], 0.0)
    w_p = w["n_ave"] * world.L_domain / n_sp
    plist = [ip.Particle(K.m_e, -K.e, w_p, 0, 0, "e", 1), ip.Particle(w["ion_amu"] * K.m_amu - K.m_e, K.e, w_p, 0, 0, "ion", 1)]
    solver._bind(plist)
    for k, T in ((0, w["T_e"]), (1, w["T_i"])):
        _lib.check(L.ipic_generate_particles_dist(solver._h, k, n_sp, T, 0.0, 0.0, 0, bench.SEED))
    ip.initializeSolver(solver, w["eps_a"], w["eps_r"], 0, w["m"], w["beta"], w["maxIter"])
    for p in plist:
        _lib.check(L.ipic_sort_particles(solver._h, p._index))
    solver.solveInitialPotential(plist, world, 0.0)
    del_t = K.del_t_from_deck(w["n_ave"], w["frac"], w["dt_deck"])
    print("generated + initial potential", file=sys.stderr, flush=True)
    ip.solvePotential(solver, plist, world, del_t, del_t, del_t, 0.0)        # one step: walls have absorbed particles
    print("one implicit step done", file=sys.stderr, flush=True)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        solver.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        solver.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    e = plist[0]
    n = e.count()
    rows = []

    def row(name, ms, nbytes, note=""):
        print("%-60s %10.3f ms" % (name, ms), file=sys.stderr, flush=True)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append(dict(op=name, ms=round(ms, 4), algorithmic_bytes=int(nbytes), GBs=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak, 3), note=note))

    # NOTE: default stream timing brackets library work through solver.synchronize() (the library runs on its own stream)
    row("ipic_sort_particles (electrons)", timed(lambda: _lib.check(L.ipic_sort_particles(solver._h, 0))), n * (8 + 32 + 32),
        "key read 8 B + gather 32 B read + 32 B write per particle; radix passes not counted")
    d = np.zeros(world.NumberXNodes)
    row("ipic_load_particle_density (electrons)", timed(lambda: _lib.check(L.ipic_load_particle_density(solver._h, 0, d.ctypes.data_as(_lib._dp)))), n * 8,
        "reads l")
    row("ipic_local_temperature (electrons)", timed(lambda: e.localTemperature(world.NumberXHalfNodes)), n * 32, "reads l, v_x, v_y, v_z")
    row("ipic_get_total_KE (electrons)", timed(lambda: e.getTotalKE()), n * 32, "reads l, v_x, v_y, v_z")
    o3 = np.zeros(3)
    row("ipic_velocity_extrema (electrons)", timed(lambda: _lib.check(L.ipic_velocity_extrema(solver._h, 0, o3.ctypes.data_as(_lib._dp)))), n * 32,
        "reads l, v_x, v_y, v_z")
    h = sim.Histograms(plist, 100)
    row("ipic_velocity_histograms (electrons, 100 bins)", timed(lambda: _lib.check(L.ipic_velocity_histograms(
        solver._h, 0, 100, h.VMax[0], h.Emin[0], h.EMax[0], h.VHist[0].ctypes.data_as(_lib._dp), h.EHist[0].ctypes.data_as(_lib._dp)))), n * 32,
        "reads l, v_x, v_y, v_z; shared-memory int64 atomics on the two bins of each histogram")
    pinj.setInjectionInputs(heating=True, fractionFreqHeating=1.0)
    row("ipic_maxwellian_heating (electrons, 1 % redrawn)", timed(lambda: pinj.maxwellianHeating(e, None, 0.01, w["T_e"], 1.0, 1.0)), n * 8 + 0.01 * n * 24,
        "reads l, writes v of the redrawn 1 %")
    pinj.setInjectionInputs()
    e._refresh_stats()
    lost = e.delIdx
    ms = timed(lambda: pinj.addMaxwellianLostParticles(plist, w["T_e"], w["T_i"], None, world), reps=1)
    row("ipic_add_maxwellian_lost_particles (%d pairs)" % lost, ms, lost * 64, "writes 2 x 32 B per pair; launch + sync latency dominates")
    print(json.dumps(dict(particles_per_species=n_sp, hbm_peak_GBs=peak, rows=rows), indent=1))
    solver.close()


if __name__ == "__main__":
    main()
