"""BASELINE configs[2] at scale, GPU against the CPU oracle on IDENTICAL particles (GPU box; the oracle only builds the particle set here).
The oracle's own JFNK counters for the same particle set are produced by `python scripts/c3_parity_probe.py oracle <n per species> <steps>`
on any CPU (minutes per step at 2e7 particles per species); `gpu` prints the device's.  Per step: one tuple (iterm, nfe, nli, nni, nbt)
per solve attempt of adaptiveSolveDivAmpereJFNK (mod_nonLinSolvers.f90:284-314)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from oracle import oracle_py as op

mode = sys.argv[1] if len(sys.argv) > 1 else "gpu"
n_sp = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
w = bench.WORKLOADS["c3"]
T = 8
os.environ.setdefault("OMP_NUM_THREADS", str(T))
del_t = op.plasma_del_t(w["n_ave"], w["frac"], w["dt_deck"])
o = op.make_deck_case(w["scheme"], w["Nx"], w["L"], w["n_ave"], w["T_e"], w["T_i"], n_sp // T, T, w["bcL"], w["bcR"], w["Vl"], w["Vr"],
                      w["rf"], bool(w["smooth"]), ion_amu=w["ion_amu"], eps_a=w["eps_a"], eps_r=w["eps_r"], m_Anderson=w["m"],
                      Beta_k=w["beta"], maxIter=w["maxIter"], native=True, alpha=w.get("alpha", 0.0), distType=w.get("distType", 0),
                      v_drift=w.get("v_drift", 0.0), factor=1, solverType=1)
print("particle set built:", n_sp, "per species, del_t", del_t, flush=True)

if mode == "oracle":
    def solve(dt):
        info = o.solveJFNK(dt, w["eps_a"], w["eps_r"], w["maxIter"], w["m"])
        return (info["iterm"], info["nfe"], info["nli"], info["nni"], info["nbt"])
    def reset():
        o.L.orc_reset_sweep_counters(o.h); o.resetVoltage()
else:
    import implicitpic1d_b200 as ip
    from implicitpic1d_b200 import _lib
    from common import build_cuda_from_oracle
    d = dict(scheme=w["scheme"], Nx=w["Nx"], L=w["L"], bcL=w["bcL"], bcR=w["bcR"], smooth=w["smooth"], T_e=w["T_e"], n_ave=w["n_ave"], Vl=w["Vl"], Vr=w["Vr"],
             rf=w["rf"], T=T, eps_a=w["eps_a"], eps_r=w["eps_r"], m=w["m"], beta=w["beta"], maxIter=w["maxIter"])
    world, solver, plist = build_cuda_from_oracle(d, o)
    ip.initializeSolver(solver, w["eps_a"], w["eps_r"], 1, w["m"], w["beta"], w["maxIter"])
    del o
    def solve(dt):
        i = ip.solveJFNK(solver, plist, world, dt)
        return (i.iterm, i.nfe, i.nli, i.nni, i.nbt)
    def reset():
        solver.resetVoltage()

remain = del_t
for step in range(steps):
    t0 = time.perf_counter()
    reset()
    curr, halv = remain, 0
    hist = [solve(curr)]
    while hist[-1][0] != 0 and halv < 4:
        curr, halv = curr / 2.0, halv + 1
        hist.append(solve(curr))
    remain = del_t if halv == 0 else remain - curr
    print("step", step, "halvings", halv, "secs %.1f" % (time.perf_counter() - t0), hist, flush=True)
    if hist[-1][0] != 0:
        print("STOP: ALREADY REDUCED TIME STEP MORE THAN 3 TIMES")
        break
