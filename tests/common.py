"""Shared builders for the parity tests: one deck -> (oracle state, CUDA state) on identical particles."""
import numpy as np
from oracle import oracle_py as op
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import constants as K
from implicitpic1d_b200.domain import readWorld

# Scaled-down versions of the reference decks (OpenMP/NGP/InputData, OpenMP/CIC/InputData)
DECKS = {
    # NGP default deck: 1001 nodes, He RF-CCP, bc 1/4, smoothing on; particle count and time step scaled
    "ngp_rf": dict(scheme=0, Nx=201, L=0.067, n_ave=1.6e16, T_e=4.0, T_i=0.025, ppt=2500, T=4, bcL=1, bcR=4, Vl=0.0, Vr=2000.0,
                   rf=200e6, smooth=True, frac=2.0, dt_deck=1.0, eps_a=1e-8, m=10, beta=0.5, maxIter=100, ion_amu=4.002602),
    # CIC default deck: 256 nodes, bc 1/4, smoothing on
    "cic_rf": dict(scheme=1, Nx=128, L=0.067, n_ave=5.12e14, T_e=2.59, T_i=0.025, ppt=2048, T=4, bcL=1, bcR=4, Vl=0.0, Vr=200.0,
                   rf=13.56e6, smooth=True, frac=1.5, dt_deck=9.218289085545723e-11, eps_a=1e-6, m=10, beta=0.5, maxIter=100,
                   ion_amu=4.002602),
    # Dirichlet / Neumann, no smoothing, sine grid
    "ngp_neumann": dict(scheme=0, Nx=129, L=0.05, n_ave=1e15, T_e=3.0, T_i=0.05, ppt=3000, T=3, bcL=2, bcR=1, Vl=0.0, Vr=50.0,
                        rf=0.0, smooth=False, frac=1.0, dt_deck=1.0, eps_a=1e-8, m=5, beta=0.5, maxIter=100, ion_amu=1.00784,
                        gridType=1, del_x=1e-4),
    # periodic NGP
    "ngp_periodic": dict(scheme=0, Nx=65, L=0.01, n_ave=1e15, T_e=2.0, T_i=0.1, ppt=4000, T=2, bcL=3, bcR=3, Vl=0.0, Vr=0.0,
                         rf=0.0, smooth=False, frac=0.5, dt_deck=1.0, eps_a=1e-8, m=10, beta=0.5, maxIter=100, ion_amu=4.002602),
    "cic_neumann": dict(scheme=1, Nx=96, L=0.03, n_ave=2e15, T_e=3.0, T_i=0.05, ppt=3000, T=3, bcL=1, bcR=2, Vl=20.0, Vr=0.0,
                        rf=0.0, smooth=False, frac=1.0, dt_deck=1.0, eps_a=1e-7, m=8, beta=0.5, maxIter=100, ion_amu=4.002602),
    "cic_periodic": dict(scheme=1, Nx=64, L=0.01, n_ave=1e15, T_e=2.0, T_i=0.1, ppt=4000, T=2, bcL=3, bcR=3, Vl=0.0, Vr=0.0,
                         rf=0.0, smooth=True, frac=0.5, dt_deck=1.0, eps_a=1e-7, m=10, beta=0.5, maxIter=100, ion_amu=4.002602),
}


# NGP_Periodic (E-field form) decks: OpenMP/NGP_Periodic/InputData (ion-acoustic shock wave of Chacon 2013: sine-perturbed
# density, drifting Maxwellians, mass ratio 1e4, periodic walls, JFNK eps_r = 1e-8, m = 20), scaled down
PERIODIC_DECKS = {
    "ngpp_iasw": dict(scheme=2, Nx=65, L=0.33245599281881233 / 8, n_ave=1e14, T_e=5.0, T_i=0.01, ppt=4000, T=2, bcL=3, bcR=3, Vl=0.0, Vr=0.0,
                      rf=0.0, smooth=False, frac=1.0, dt_deck=1.0, eps_a=0.0, eps_r=1e-8, m=20, beta=0.5, maxIter=100,
                      ion_amu=5.48634767053312, alpha=0.4, distType=2, v_drift=-9372.932223215392),
    # bounded variant of the same tree: Dirichlet / Neumann walls exercise absorption and reflection in cell-local coordinates
    "ngpp_walls": dict(scheme=2, Nx=97, L=0.02, n_ave=5e14, T_e=3.0, T_i=0.05, ppt=3000, T=3, bcL=2, bcR=1, Vl=0.0, Vr=10.0,
                       rf=0.0, smooth=False, frac=1.0, dt_deck=1.0, eps_a=0.0, eps_r=1e-8, m=10, beta=0.5, maxIter=100,
                       ion_amu=4.002602),
}


def deck_del_t(d):
    return K.del_t_from_deck(d["n_ave"], d["frac"], d["dt_deck"])


def build_oracle(d, factor=1):
    return op.make_deck_case(d["scheme"], d["Nx"], d["L"], d["n_ave"], d["T_e"], d["T_i"], d["ppt"], d["T"], d["bcL"], d["bcR"],
                             d["Vl"], d["Vr"], d["rf"], d["smooth"], ion_amu=d["ion_amu"], gridType=d.get("gridType", 0),
                             del_x=d.get("del_x", 0.0), eps_a=d["eps_a"], eps_r=d.get("eps_r", 0.0), m_Anderson=d["m"], Beta_k=d["beta"],
                             maxIter=d["maxIter"], factor=factor, alpha=d.get("alpha", 0.0), distType=d.get("distType", 0),
                             v_drift=d.get("v_drift", 0.0))


def build_cuda_from_oracle(d, o, tridiag_mode=None, device=0):
    """CUDA-side program state (main.f90:25-40) on the oracle's particle set."""
    world = readWorld(d["scheme"], d["Nx"], d["L"], d.get("gridType", 0), d.get("del_x", 0.0), d["bcL"], d["bcR"], int(d["smooth"]),
                      d["T_e"], d["n_ave"])
    solver = ip.potentialSolver(world, d["Vl"], d["Vr"], d["rf"], 0.0, device=device)
    if tridiag_mode is not None:
        solver.set_tridiag_mode(tridiag_mode)
    plist = []
    for sp in range(o.n_species):
        sc = o.species_scalars(sp)
        ps = o.phase_space(sp)
        p = ip.Particle(sc["mass"], sc["q"], sc["w_p"], 0, ps.shape[1], "e" if sp == 0 else "ion", d["T"])
        p.set_phase_space(ps.copy(), o.N_p(sp))
        plist.append(p)
    ip.initializeSolver(solver, d["eps_a"], d.get("eps_r", 0.0), 0, d["m"], d["beta"], d["maxIter"])
    solver.solveInitialPotential(plist, world, 0.0)
    return world, solver, plist


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / (den if den > 0 else 1.0)
