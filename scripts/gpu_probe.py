"""First-contact GPU probe: staged checks with prints (run under gpurun with timeouts)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ctypes as C
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib, constants as K
from implicitpic1d_b200.domain import readWorld
from implicitpic1d_b200.particle_mover import depositSubsteps

stage = sys.argv[1] if len(sys.argv) > 1 else "parity"

if stage == "parity":
    from common import DECKS, build_oracle, build_cuda_from_oracle, deck_del_t, relerr
    for name in sys.argv[2:] or ["ngp_rf", "cic_rf"]:
        d = DECKS[name]
        o = build_oracle(d)
        print(name, "oracle built", flush=True)
        world, solver, plist = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS)
        print(" cuda built; rho err", relerr(solver.rho, o.get("rho")), "phi err", relerr(solver.phi, o.get("phi")), flush=True)
        solver.phi, solver.phi_f, solver.rho = o.get("phi"), o.get("phi_f"), o.get("rho")
        o.makeHalfTimeEField(); solver.makeHalfTimeEField()
        print(" E err", relerr(solver.EField, o.get("EField")), flush=True)
        del_t = deck_del_t(d)
        o.depositJ(del_t)
        ip.depositJ(solver, plist, world, del_t)
        print(" J err", relerr(solver.J, o.get("J")), "substeps", [depositSubsteps(solver, p) for p in plist],
              [o.move_stats(s)["substepsDeposit"] for s in range(2)], flush=True)
        ro = o.solvePotential(del_t, del_t, del_t, 0.0)
        rg = ip.solvePotential(solver, plist, world, del_t, del_t, del_t, 0.0)
        st = ip.SolverState(solver)
        print(" step:", ro, rg, "iters", st.iterNumPicard, o.solver_stats()["iterNumPicard"], "sweeps", st.sweeps, o.solver_stats()["sweeps"],
              "substeps", st.substeps, o.solver_stats()["substeps"], flush=True)
        print(" norms gpu", st.normHistory[:6], "\n norms orc", o.norm_history()[:6], flush=True)
        print(" phi_f err", relerr(solver.phi_f, o.get("phi_f")), flush=True)
        solver.close()
    print("PARITY PROBE DONE", flush=True)

elif stage in ("perf", "prof"):
    Nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4097
    N = int(float(sys.argv[3])) if len(sys.argv) > 3 else 20_000_000
    scheme = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    # default NGP deck (OpenMP/NGP/InputData) with the cell size kept: L = 0.067 m * (Nx-1)/1000
    n_ave, T_e = 1.6e18, 4.0
    L_dom = 0.067 * (Nx - 1) / 1000.0
    world = readWorld(scheme, Nx, L_dom, 0, 0.0, 1, 4, 1, T_e, n_ave)
    solver = ip.potentialSolver(world, 0.0, 2000.0, 200e6, 0.0)
    L = ip.load_library()
    m_ion = 4.002602 * K.m_amu - K.m_e
    plist = [ip.Particle(K.m_e, -K.e, n_ave * world.L_domain / N, 0, 0, "e", 1), ip.Particle(m_ion, K.e, n_ave * world.L_domain / N, 0, 0, "He+", 1)]
    solver._bind(plist)
    _lib.check(L.ipic_generate_particles(solver._h, 0, N, T_e, 0.0, 1234))
    _lib.check(L.ipic_generate_particles(solver._h, 1, N, 0.025, 0.0, 1234))
    ip.initializeSolver(solver, 1e-8, 0.0, 0, 10, 0.5, 100)
    solver.solveInitialPotential(plist, world, 0.0)
    del_t = K.del_t_from_deck(n_ave, 10.0, float(sys.argv[5]) if len(sys.argv) > 5 else 2.5e-11)
    print("del_t", del_t, "N per species", N, "Nx", Nx, "scheme", scheme, flush=True)

    def timeit(label, reps=5):
        ip.depositJ(solver, plist, world, del_t)   # warm
        solver.reset_timing()
        for _ in range(reps):
            ip.depositJ(solver, plist, world, del_t)
        for k in ("depositJ_sp0", "depositJ_sp1"):
            ms, n = solver.timing(k)
            sub = depositSubsteps(solver, plist[0 if k.endswith("0") else 1])
            t = ms / max(n, 1)
            print("  %-22s %-14s %8.3f ms  %7.1f GB/s  substeps/particle %.3f  %.3e substeps/s" % (label, k, t, N * 16 / t / 1e6, sub / N, sub / t * 1e3), flush=True)

    for pp in plist:
        _lib.check(L.ipic_sort_particles(solver._h, pp._index))
    # relax the violent random start: a few implicit steps, then re-sort (steady-state operating point)
    nwarm = int(sys.argv[7]) if len(sys.argv) > 7 else 3
    tcur, remain, curr = 0.0, del_t, del_t
    for k in range(nwarm):
        remain, curr = ip.solvePotential(solver, plist, world, del_t, remain, curr, tcur)
        tcur += curr
        st = ip.SolverState(solver)
        print("warm step", k, "iters", st.iterNumPicard, "substeps/particle/sweep %.3f" % (st.substeps / (st.sweeps + 1) / (2 * N)), "norm0 %.3e" % st.normHistory[0], flush=True)
    for pp in plist:
        _lib.check(L.ipic_sort_particles(solver._h, pp._index))
    if os.environ.get("IPIC_PROF_STEP"):
        # profile one whole implicit step at the operating point reached after `nwarm` steps (run ncu with --profile-from-start off)
        cu = C.CDLL("libcuda.so.1")
        solver.synchronize()
        cu.cuProfilerStart()
        remain, curr = ip.solvePotential(solver, plist, world, del_t, remain, curr, tcur)
        solver.synchronize()
        cu.cuProfilerStop()
        st = ip.SolverState(solver)
        print("profiled step: iters", st.iterNumPicard, "sweeps", st.sweeps, "substeps/particle/sweep %.3f" % (st.substeps / (st.sweeps + 1) / (2 * N)), flush=True)
        print("PROF STEP DONE"); sys.exit(0)
    if stage == "prof":
        solver.set_option("use_tma", int(sys.argv[6]) if len(sys.argv) > 6 else 1)
        ip.depositJ(solver, plist, world, del_t)
        solver.synchronize()
        cu = C.CDLL("libcuda.so.1")
        cu.cuProfilerStart()            # run ncu with --profile-from-start off
        ip.depositJ(solver, plist, world, del_t)
        solver.synchronize()
        cu.cuProfilerStop()
        print("PROF DONE"); sys.exit(0)
    for tma in (1,):
        solver.set_option("use_tma", tma)
        timeit("sorted use_tma=%d" % tma)
    solver.set_option("use_tma", int(sys.argv[6]) if len(sys.argv) > 6 else 1)
    solver.reset_timing()
    t0 = time.time()
    r, c = ip.solvePotential(solver, plist, world, del_t, remain, curr, tcur)
    solver.synchronize()
    wall = time.time() - t0
    st = ip.SolverState(solver)
    print("solvePotential wall %.3f s iters %d sweeps %d substeps %d -> %.3e substeps/s ; mover %.3f s field %.3f s" %
          (wall, st.iterNumPicard, st.sweeps, st.substeps, st.substeps / wall, st.moverTimer, st.potTimer), flush=True)
    for k in ("depositJ_kernels", "moveParticles_kernels", "compaction", "field_kernels"):
        print("   ", k, solver.timing(k))
    print(" norms", st.normHistory)
    solver.close()
    print("PERF PROBE DONE", flush=True)
