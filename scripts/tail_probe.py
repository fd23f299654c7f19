"""Measurement (GPU box, library built with -DIPIC_PUSH_STATS): how long after the average warp the LAST warp of a depositJ launch finishes.
usage: python scripts/tail_probe.py [particles per species] [warm steps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib, constants as K
from implicitpic1d_b200.domain import readWorld

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 62_500_000
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 4
Nx, n_ave, T_e = 4097, 1.6e18, 4.0
world = readWorld(0, Nx, 0.067 * (Nx - 1) / 1000.0, 0, 0.0, 1, 4, 1, T_e, n_ave)
solver = ip.potentialSolver(world, 0.0, 2000.0, 200e6, 0.0)
L = ip.load_library()
m_ion = 4.002602 * K.m_amu - K.m_e
w_p = n_ave * world.L_domain / N
plist = [ip.Particle(K.m_e, -K.e, w_p, 0, 0, "e", 1), ip.Particle(m_ion, K.e, w_p, 0, 0, "He+", 1)]
solver._bind(plist)
_lib.check(L.ipic_generate_particles(solver._h, 0, N, T_e, 0.0, 1234))
_lib.check(L.ipic_generate_particles(solver._h, 1, N, 0.025, 0.0, 1234))
ip.initializeSolver(solver, 1e-8, 0.0, 0, 10, 0.5, 100)
for p in plist:
    _lib.check(L.ipic_sort_particles(solver._h, p._index))
solver.solveInitialPotential(plist, world, 0.0)
del_t = K.del_t_from_deck(n_ave, 10.0, 2.5e-11)
t, remain, curr = 0.0, del_t, del_t
for s in range(warm):
    _lib.check(L.ipic_sort_particles(solver._h, 0))
    remain, curr = ip.solvePotential(solver, plist, world, del_t, remain, curr, t)
    t += curr
# an un-converged iterate: perturb phi_f as the first Anderson iterates do (the field of the previous step against the new RF phase)
solver.set_option("dump_push_stats", 1)
for k, scale in enumerate((1.0, 1.3)):
    phi = solver.phi_f
    solver.phi_f = phi * scale
    solver.makeHalfTimeEField()
    for sp, p in enumerate(plist):
        solver.reset_timing()
        _lib.check(L.ipic_deposit_species(solver._h, sp, del_t)) if hasattr(L, "ipic_deposit_species") else ip.depositJ(solver, [p] if False else plist, world, del_t)
        print("phi_f x %.1f: launch timing e %.3f ms, i %.3f ms" % (scale, solver.timing("depositJ_sp0")[0], solver.timing("depositJ_sp1")[0]), flush=True)
        solver.set_option("dump_push_stats", 1)
        break
    solver.phi_f = phi
solver.close()
