#!/usr/bin/env python
"""bench.py -- throughput of the implicit-PIC inner loop (particle-substeps/s per implicit step).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU oracle; rank 0 only)

A "step" is one implicit step = one `solvePotential` call (mod_nonLinSolvers.f90:198-215): K_res residual evaluations
(`depositJ` sweep + field solve + Anderson mixing) and the final `moveParticles`.  The metric counts every executed
sub-step trip (`do while (del_tau > f_tol)` body) of every particle in every sweep and divides by the time of the step.

Workload at every N: the per-GPU shard of BASELINE.json configs[4] (the configuration the target is quoted on: NGP bounded
plasma, 4096 cells, 1e9 particles over 8 GPUs = 1.25e8 particles per GPU, weak scaling), built from the default NGP deck
(OpenMP/NGP/InputData) with the cell size kept.  Particles are generated on the device (counter-based PCG32) because 4 GB per
GPU is too much to draw on the host in a bench; the e2e leg then moves them through HOST buffers every step.

N > 1 (torchrun, one rank per GPU): the J grid is all-reduced once per residual evaluation, through peer memory over NVLink
(csrc/peer.cu; the CUDA IPC handles of the ranks' windows travel through torch.distributed) or, with --no-peer or when the windows
cannot be mapped, through NCCL.

Keys beyond the contract: `particle_sweeps_per_s` (particles x sweeps per second: sub-steps per particle per sweep are physics and
fall with the noise level, i.e. with the GPU count under weak scaling -- DESIGN.md 5), `e2e.upload_ms_per_step` /
`e2e.download_ms_per_step` (the PCIe share of the e2e leg), `e2e_resident`, `roofline.move_kernel` / `ion_sweeps` / `issue_bound`,
`breakdown_ms_per_step`.

Prints ONE JSON line (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ---- workloads (deck values: OpenMP/NGP/InputData/*.inp, OpenMP/CIC/InputData/*.inp) -----------------------------------------
WORKLOADS = {
    # configs[4] per-GPU shard: default NGP deck, 4096 cells at the deck's cell size (L = 0.067 m * 4096/1000)
    "c5": dict(name="C5 shard: NGP bounded He RF plasma, 4096 cells (N_x=4097), 1.25e8 particles/GPU (2 species), default NGP deck, "
                    "Anderson m=10 eps_a=1e-8",
               scheme=0, Nx=4097, L=0.067 * 4096 / 1000.0, n_ave=1.6e18, T_e=4.0, T_i=0.025, bcL=1, bcR=4, Vl=0.0, Vr=2000.0, rf=200.0e6,
               smooth=1, frac=10.0, dt_deck=1.0, eps_a=1e-8, eps_r=0.0, m=10, beta=0.5, maxIter=100, ion_amu=4.002602,
               per_species_per_gpu=62_500_000),
    # configs[1]: default CIC deck scaled to 1e7 particles
    "c2": dict(name="C2: CIC bounded He RF plasma, 256 cells, 1e7 particles (2 species), default CIC deck, Anderson m=10 eps_a=1e-6",
               scheme=1, Nx=256, L=0.067, n_ave=5.12e14, T_e=2.59, T_i=0.025, bcL=1, bcR=4, Vl=0.0, Vr=200.0, rf=13.56e6,
               smooth=1, frac=1.5, dt_deck=9.218289085545723e-11, eps_a=1e-6, eps_r=0.0, m=10, beta=0.5, maxIter=100, ion_amu=4.002602,
               per_species_per_gpu=5_000_000),
    # configs[2]: NGP_Periodic (E-field form), 2048 cells, 1e8 particles, JFNK (deck template OpenMP/NGP_Periodic/InputData:
    # ion-acoustic shock wave, sine-perturbed drifting Maxwellians, eps_r = 1e-8, m = 20), cell size of the deck kept
    "c3": dict(name="C3: NGP_Periodic (E-form) ion-acoustic load, 2048 cells, 1e8 particles (2 species), JFNK (NITSOL) eps_r=1e-8 m=20",
               scheme=2, Nx=2049, L=0.33245599281881233 * 2048 / 512, n_ave=1e14, T_e=5.0, T_i=0.01, bcL=3, bcR=3, Vl=0.0, Vr=0.0, rf=0.0,
               smooth=0, frac=10.0, dt_deck=1.0, eps_a=0.0, eps_r=1e-8, m=20, beta=0.5, maxIter=100, ion_amu=5.48634767053312,
               alpha=0.4, distType=2, v_drift=-9372.932223215392, solver_type=1, per_species_per_gpu=50_000_000),
    # configs[3]: bounded argon plasma with null-collision MCC (OpenMP/NGP/InputDataSets/NGP_VassBenchmark_Ar_10xParticles: 257 nodes, 2.5 cm,
    # RF 250 V 13.56 MHz, n = 1e16, eps_a = 1e-4, Ar 300 K 2.4e21 m^-3, table Ar_electron_table.dat), 1e8 particles over 2 GPUs = 5e7 per GPU;
    # a step = solvePotential + generateCollision of every reactant pair (mod_simulation.f90:157-160)
    "c4": dict(name="C4: NGP bounded Ar RF plasma with MCC (Ar_electron_table.dat), 256 cells, 5e7 particles/GPU (2 species), Vass benchmark deck, "
                    "Anderson m=10 eps_a=1e-4",
               scheme=0, Nx=257, L=0.025, n_ave=1.0e16, T_e=2.59, T_i=0.025, bcL=1, bcR=4, Vl=0.0, Vr=250.0, rf=13.56e6,
               smooth=1, frac=1.5, dt_deck=7.845352413230403e-12, eps_a=1e-4, eps_r=0.0, m=10, beta=0.5, maxIter=100, ion_amu=39.948,
               per_species_per_gpu=25_000_000, mcc=dict(table="Ar_electron_table.dat", neutral=("Ar", 39.948, 2.4e21, 300.0), ion_name="Ar+")),
    # configs[0]: the default NGP deck itself (CPU-runnable case)
    "c1": dict(name="C1: default NGP deck, 1000 cells, 1.536e6 particles (2 species)",
               scheme=0, Nx=1001, L=0.067, n_ave=1.6e18, T_e=4.0, T_i=0.025, bcL=1, bcR=4, Vl=0.0, Vr=2000.0, rf=200.0e6,
               smooth=1, frac=10.0, dt_deck=1.0, eps_a=1e-8, eps_r=0.0, m=10, beta=0.5, maxIter=100, ion_amu=4.002602,
               per_species_per_gpu=768_000),
}
METRIC = "particle-substeps/sec per implicit step"
UNIT = "particle-substeps/s"
SEED = 20261017


TRANSIT_FLOOR_CYCLES = 29.3    # scripts/micro/transit_bench2.cu 1 on B200: the shipped crossing loop alone, 8 warps per SM sub-partition, the lanes of a warp in one cell (cell-sorted particles); 38.7 with every lane in another line of the table (L1 gather-bound)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ---- clocks sampler ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


# ---- the reference arm / cpu_baseline: the CPU oracle timed on the host cores ---------------------------------------------------------
def cpu_reference_run(w, steps, warmup, budget_s):
    """Times the C++/OpenMP restatement of the reference path (oracle/, built -march=native) on a bounded sample of the
    workload: same grid, deck, del_t and tolerances, fewer particles (substeps/s does not depend on the particle count once the
    slabs exceed the caches).  Returns (substeps_per_s, info)."""
    from oracle import oracle_py as op
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    T = cores
    del_t = op.plasma_del_t(w["n_ave"], w["frac"], w["dt_deck"])

    def make(ppt):
        return op.make_deck_case(w["scheme"], w["Nx"], w["L"], w["n_ave"], w["T_e"], w["T_i"], ppt, T, w["bcL"], w["bcR"], w["Vl"], w["Vr"],
                                 w["rf"], bool(w["smooth"]), ion_amu=w["ion_amu"], eps_a=w["eps_a"], eps_r=w["eps_r"], m_Anderson=w["m"],
                                 Beta_k=w["beta"], maxIter=w["maxIter"], native=True, alpha=w.get("alpha", 0.0), distType=w.get("distType", 0),
                                 v_drift=w.get("v_drift", 0.0), factor=2 if w.get("mcc") else 1)   # ionisation appends need head-room

    mcc = w.get("mcc")
    COLL = os.path.join(ROOT, "tests", "golden", "collision_data")

    def add_mcc(o):
        # readNullCollisionInputs + the neutral background, on the oracle side (mod_NullCollision.f90:666-939)
        from oracle import mcc_py as mp
        nm, amu, dens, temp = mcc["neutral"]
        mp.add_target(o, amu * op.m_amu, dens, temp)
        specs = mp.read_null_collision_inputs([os.path.join(COLL, mcc["table"])], [("e", op.m_e), (mcc["ion_name"], w["ion_amu"] * op.m_amu - op.m_e)],
                                              [(nm, amu * op.m_amu)])
        o._ncs = [mp.NullCollisionOracle(o, sp) for sp in specs]
        o._seeds = op.default_seeds(T)
        o._mp = mp

    def run_steps(o, n, tcur, remain, curr):
        sub, secs, iters = 0, 0.0, []
        for _ in range(n):
            t0 = time.perf_counter()
            if w.get("solver_type", 0) == 1:
                # solvePotential with solverType 1: resetVoltage + adaptiveSolveDivAmpereJFNK (step halving on NITSOL failure)
                o.L.orc_reset_sweep_counters(o.h)
                o.resetVoltage()
                curr, halvings = remain, 0
                info = o.solveJFNK(curr, w["eps_a"], w["eps_r"], w["maxIter"], w["m"])
                while info["iterm"] != 0 and halvings < 4:
                    curr, halvings = curr / 2.0, halvings + 1
                    info = o.solveJFNK(curr, w["eps_a"], w["eps_r"], w["maxIter"], w["m"])
                remain = del_t if halvings == 0 else remain - curr
                its = info["nli"]
            else:
                remain, curr = o.solvePotential(del_t, remain, curr, tcur)
                its = None
            if mcc:
                for sp in range(o.n_species):
                    o._mp.set_num_to_collide(o, sp)
                for nc in o._ncs:
                    nc.generateCollision(o._seeds, curr)
            secs += time.perf_counter() - t0
            tcur += curr
            st = o.solver_stats()
            sub += st["substeps"]
            iters.append(st["iterNumPicard"] if its is None else its)
        return sub, secs, iters, tcur, remain, curr

    # calibrate on a small load, then size the sample so that (warmup + steps) fit the budget
    ppt0 = max(2000, 200_000 // T)
    o = make(ppt0)
    if mcc:
        add_mcc(o)
    sub, secs, iters, *_ = run_steps(o, 1, 0.0, del_t, del_t)
    rate0 = sub / max(secs, 1e-9)
    per_particle_step = sub / (2.0 * ppt0 * T)                       # substeps per particle per implicit step
    want = budget_s * rate0 / max(per_particle_step, 1.0) / max(steps + warmup, 1) / 2.0   # particles per species
    ppt = int(min(max(want / T, ppt0), 2_000_000))
    del o
    o = make(ppt)
    if mcc:
        add_mcc(o)
    tcur, remain, curr = 0.0, del_t, del_t
    _, _, _, tcur, remain, curr = run_steps(o, warmup, tcur, remain, curr)
    sub, secs, iters, *_ = run_steps(o, steps, tcur, remain, curr)
    info = dict(cores=cores, kind="port",
                sample="%d particles/species (%d per OpenMP slab x %d threads) of the same grid/deck/del_t, %d timed implicit steps after %d warm-up; "
                       "mean residual evaluations/step %.1f; oracle/ipic_oracle.cpp -O3 -march=native -fopenmp, no fast-math"
                       % (ppt * T, ppt, T, steps, warmup, sum(iters) / max(len(iters), 1)),
                seconds=secs, substeps=sub, ms_per_step=1e3 * secs / max(steps, 1))
    return sub / max(secs, 1e-9), info


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    val, info = cpu_reference_run(w, args.steps, args.warmup, args.cpu_seconds * 4)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": w["name"], "note": "CPU path on a bounded particle sample of the workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---- the CUDA arm -----------------------------------------------------------------------------------------------------------------------
def run_cuda_arm(args, w):
    import numpy as np
    import torch
    import implicitpic1d_b200 as ip
    from implicitpic1d_b200 import _lib, constants as K
    from implicitpic1d_b200.domain import readWorld
    from implicitpic1d_b200.particle_mover import depositSubsteps

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size != args.gpus:
        if world_size == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    peer_on = False
    if world_size > 1:
        # stdout carries exactly one JSON line: whatever NCCL_DEBUG level the environment sets goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/ipic_bench_nccl_%h_%p.log")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L = ip.load_library()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    # ---- program state (main.f90:25-40) ----
    n_sp = args.particles_per_species or w["per_species_per_gpu"]
    N_total_per_species = n_sp * world_size
    world = readWorld(w["scheme"], w["Nx"], w["L"], 0, 0.0, w["bcL"], w["bcR"], w["smooth"], w["T_e"], w["n_ave"])
    solver = ip.potentialSolver(world, w["Vl"], w["Vr"], w["rf"], 0.0, device=local_rank)
    _lib.check(L.ipic_set_stream(solver._h, C.c_void_p(stream.cuda_stream)))
    if world_size > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(L.ipic_nccl_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().numpy().tobytes())
        _lib.check(L.ipic_comm_init(solver._h, world_size, rank, C.c_char_p(raw)))
        # one node: the J all-reduce of every residual evaluation through peer memory (NVLink) instead of NCCL (csrc/peer.cu)
        from implicitpic1d_b200.distributed import peer_init
        peer_on = peer_init(solver, dist, rank, world_size, device=torch.device("cuda", local_rank)) if not args.no_peer else False
    w_p = w["n_ave"] * world.L_domain / N_total_per_species                 # mod_particle.f90:76-82 with SUM(N_p) over all ranks
    m_ion = w["ion_amu"] * K.m_amu - K.m_e                                   # mod_particle.f90:334
    mcc = w.get("mcc")
    plist = [ip.Particle(K.m_e, -K.e, w_p, 0, 0, "e", 1), ip.Particle(m_ion, K.e, w_p, 0, 0, mcc["ion_name"] if mcc else "ion", 1)]
    solver._bind(plist)
    targets, ncs = [], []
    if mcc:
        nm, amu, dens, temp = mcc["neutral"]
        targets = [ip.targetParticle(nm, amu * K.m_amu, dens, temp)]
        # every rank collides its own shard with its own counter-based stream
        ncs = ip.readNullCollisionInputs([mcc["table"]], plist, targets, collision_dir=os.path.join(ROOT, "tests", "golden", "collision_data"),
                                         seed=0x1D1C0111 + 7919 * rank)
    _lib.check(L.ipic_generate_particles_dist(solver._h, 0, n_sp, w["T_e"], w.get("v_drift", 0.0), w.get("alpha", 0.0), w.get("distType", 0), SEED))
    _lib.check(L.ipic_generate_particles_dist(solver._h, 1, n_sp, w["T_i"], w.get("v_drift", 0.0), w.get("alpha", 0.0), w.get("distType", 0), SEED))
    ip.initializeSolver(solver, w["eps_a"], w["eps_r"], w.get("solver_type", 0), w["m"], w["beta"], w["maxIter"])
    if args.thomas:
        solver.set_tridiag_mode(_lib.TRIDIAG_THOMAS)
    for k, v in (args.option or []):
        solver.set_option(k, float(v))
    for p in plist:
        _lib.check(L.ipic_sort_particles(solver._h, p._index))
    solver.solveInitialPotential(plist, world, 0.0)
    del_t = K.del_t_from_deck(w["n_ave"], w["frac"], w["dt_deck"])

    state = dict(t=0.0, remain=del_t, curr=del_t, nstep=0)

    def one_step():
        if args.sort_every > 0 and state["nstep"] > 0:
            # periodic cell sort: electrons cross a cell or two per step and lose their order within a few steps;
            # ions move ~1e-3 cells per step and keep it for hundreds, so they are re-sorted 8 times less often
            for k, p in enumerate(plist):
                period = args.sort_every * (1 if k == 0 else args.sort_ion_factor)
                if state["nstep"] % period == 0:
                    _lib.check(L.ipic_sort_particles(solver._h, p._index))
        try:
            state["remain"], state["curr"] = ip.solvePotential(solver, plist, world, del_t, state["remain"], state["curr"], state["t"])
        except _lib.IpicError:
            if w.get("solver_type", 0) == 1:                        # say what NITSOL saw before the reference's `stop`
                v = np.zeros(7, dtype=np.int32)
                L.ipic_get_nitsol_info(solver._h, v.ctypes.data_as(_lib._ip))
                print("bench.py: step %d failed; last NITSOL solve: %r, norms %s" % (state["nstep"], ip.NitsolInfo(v), ip.SolverState(solver).normHistory[-6:]),
                      file=sys.stderr)
            raise
        if ncs:
            # mod_simulation.f90:157-160: numToCollide = N_p (set by the mover), then every reactant pair collides over the step just taken
            for p in plist:
                _lib.check(L.ipic_set_num_to_collide(solver._h, p._index, p.count()))
            for nc in ncs:
                nc.generateCollision(plist, targets, del_t=state["curr"])
        state["t"] += state["curr"]
        state["nstep"] += 1
        st = ip.SolverState(solver)
        return st

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        n = C.c_int64()
        _lib.check(L.ipic_get_launch_count(solver._h, C.byref(n)))
        return n.value

    # ---- warm-up ----
    for _ in range(args.warmup):
        one_step()
    for p in plist:
        _lib.check(L.ipic_sort_particles(solver._h, p._index))
    state["nstep"] = 0

    # ---- timed region: device-resident particles ----
    solver.reset_timing()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launches()
    e0.record(stream)
    substeps, sweeps, iters, particle_sweeps, dep_sub_e = 0, 0, [], 0, 0
    for _ in range(args.steps):
        n_before = sum(p.count() for p in plist)
        st = one_step()
        substeps += st.substeps
        sweeps += st.sweeps
        dep_sub_e += depositSubsteps(solver, plist[0])
        iters.append(st.iterNumPicard)
        particle_sweeps += n_before * (st.sweeps + 1)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    n_launch = launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device time from the library's own CUDA events (recorded on the launching stream)
    t_dep0, n_dep0 = solver.timing("depositJ_sp0")
    t_dep1, n_dep1 = solver.timing("depositJ_sp1")
    t_pair, n_pair = solver.timing("depositJ_pair")             # both species in one launch (the default for two-species plasmas)
    t_move, _ = solver.timing("moveParticles_kernels")
    t_comp, _ = solver.timing("compaction")
    t_field, n_field = solver.timing("field_kernels")
    t_ar, _ = solver.timing("allreduce_J")
    t_sort, _ = solver.timing("sort")
    t_coll, _ = solver.timing("collisions")
    n_e, n_i = plist[0].count(), plist[1].count()

    vals = torch.tensor([ms, float(substeps), float(particle_sweeps)], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_max, substeps_all, psweeps_all = mx[0].item(), sm[1].item(), sm[2].item()
    else:
        ms_max, substeps_all, psweeps_all = ms, float(substeps), float(particle_sweeps)
    value = substeps_all / (ms_max * 1e-3)

    # ---- e2e: the same step through the reference-facing call with HOST phase-space buffers (mod_particle.f90:23 layout) ----
    e2e = None
    if not args.no_e2e:
        finalIdx = int(max(n_e, n_i) * 1.02) + 1024
        host = []
        for p in plist:
            ptr = C.c_void_p()
            _lib.check(L.ipic_host_alloc(C.byref(ptr), finalIdx * 32))
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(1, finalIdx, 4))
            p.numThread, p.finalIdx = 1, finalIdx
            p.N_p = np.zeros(1, dtype=np.int32)
            p.phaseSpace = arr
            p.download()                                             # host copy of the current state
            host.append(ptr)
        unk = "EField_f" if w["scheme"] == 2 else "phi_f"
        phi_h = getattr(solver, unk)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        solver.reset_timing()
        f0.record(stream)
        sub2, h2d, d2h = 0, 0, 0
        for _ in range(e2e_steps):
            for p in plist:                                          # host -> device: phaseSpace + N_p of every species, phi_f
                p.upload()
                h2d += int(p.N_p.sum()) * 32
            setattr(solver, unk, phi_h)
            h2d += phi_h.nbytes
            st = one_step()
            sub2 += st.substeps
            for p in plist:                                          # device -> host: moved particles, counters, new potential
                p.download()
                d2h += int(p.N_p.sum()) * 32
            phi_h = getattr(solver, unk)
            d2h += phi_h.nbytes
        f1.record(stream)
        barrier()
        ms2 = f0.elapsed_time(f1)
        v2 = torch.tensor([ms2, float(sub2)], dtype=torch.float64, device="cuda")
        if dist is not None:
            mx = v2.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = v2.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms2_max, sub2_all = mx[0].item(), sm[1].item()
        else:
            ms2_max, sub2_all = ms2, float(sub2)
        t_up, n_up = solver.timing("upload")
        t_dn, n_dn = solver.timing("download")
        e2e = {"value": sub2_all / (ms2_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
               "steps": e2e_steps, "ms_per_step": ms2_max / e2e_steps,
               "upload_ms_per_step": t_up / e2e_steps, "download_ms_per_step": t_dn / e2e_steps,
               "path": "pinned host phaseSpace(4,finalIdx,1) -> ipic_upload_particles -> ipic_solve_potential -> ipic_download_particles"}
        for p in plist:
            p.phaseSpace = None
        for ptr in host:
            L.ipic_host_free(ptr)

    # ---- e2e_resident: the level-3 integration -- injection / MCC / diagnostics also run on the device, so the host
    # never touches phaseSpace between steps; per step it pushes its field state and pulls the new fields and the mover's statistics
    e2e_resident = None
    if not args.no_e2e:
        unk = "EField_f" if w["scheme"] == 2 else "phi_f"
        res_steps = max(1, min(args.steps, args.e2e_steps))
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        sub3, h2d3, d2h3 = 0, 0, 0
        phi_h, phi_f_h = solver.phi, getattr(solver, unk)
        for _ in range(res_steps):
            solver.phi = phi_h                                           # host -> device: phi, phi_f (pinned staging inside the library)
            setattr(solver, unk, phi_f_h)
            h2d3 += phi_h.nbytes + phi_f_h.nbytes
            st = one_step()
            sub3 += st.substeps
            phi_h, phi_f_h, J_h = solver.phi, getattr(solver, unk), solver.J     # device -> host: fields + per-species move statistics
            d2h3 += phi_h.nbytes + phi_f_h.nbytes + J_h.nbytes + len(plist) * C.sizeof(_lib.MoveStats)
        r1.record(stream)
        barrier()
        ms3 = r0.elapsed_time(r1)
        v3 = torch.tensor([ms3, float(sub3)], dtype=torch.float64, device="cuda")
        if dist is not None:
            mx = v3.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = v3.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms3_max, sub3_all = mx[0].item(), sm[1].item()
        else:
            ms3_max, sub3_all = ms3, float(sub3)
        e2e_resident = {"value": sub3_all / (ms3_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d3 // res_steps, "d2h_bytes_per_step": d2h3 // res_steps,
                        "steps": res_steps, "ms_per_step": ms3_max / res_steps,
                        "path": "particles resident (injection / MCC / diagnostics on the device): ipic_set_field(phi, phi_f) -> ipic_solve_potential -> "
                                "ipic_get_field(phi, phi_f, J) + ipic_get_move_stats"}


    # ---- roofline of the dominant kernel (the depositJ push sweep) ----
    peak, peak_src = measured_peak_gbs()
    paired = n_pair > 0
    dep_ms = t_pair if paired else t_dep0 + t_dep1
    dep_launches = n_pair if paired else n_dep0 + n_dep1
    # algorithmic bytes: 16 B (l, v_x) per particle per depositJ sweep; units per launch = the particles the launch sweeps
    # (both species when paired, one species otherwise); particle counts change by <0.1% per step (wall losses)
    alg_bytes_total = 16.0 * (n_pair * (n_e + n_i) if paired else (n_dep0 * n_e + n_dep1 * n_i))
    achieved = alg_bytes_total / max(dep_ms * 1e-3, 1e-12) / 1e9
    units = (n_e + n_i) if paired else (n_e + n_i) / 2.0
    kname = {0: "NGP", 1: "CIC", 2: "NGP_Periodic"}[w["scheme"]]
    roofline = {"bound": "hbm", "kernel": ("k_push_pair<%s> (depositJ sweep, both species in one launch)" % kname) if paired
                else ("k_push<%s,deposit> (depositJ sweep, one launch per species)" % kname),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (args.traffic_bytes_per_particle * units) if args.traffic_bytes_per_particle else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": 16.0 * units, "avg_launch_ms": dep_ms / max(dep_launches, 1),
                "per_species_ms": {"both": t_pair / max(n_pair, 1)} if paired else {"electrons": t_dep0 / max(n_dep0, 1), "ions": t_dep1 / max(n_dep1, 1)},
                "kernel_share_of_step": dep_ms / max(ms, 1e-9)}
    # the mover proper (moveParticles: k_push<*, move>, one launch per species per step): 32 B per particle (read l, v_x; write l, v_x)
    if t_move > 0:
        mv = 32.0 * (n_e + n_i) * args.steps / (t_move * 1e-3) / 1e9
        roofline["move_kernel"] = {"kernel": "k_push<%s,move> (moveParticles, one launch per species per step)" % kname, "achieved": mv, "peak": peak,
                                   "unit": "GB/s", "frac": mv / peak, "algorithmic_bytes_per_step": 32.0 * (n_e + n_i), "ms_per_step": t_move / args.steps}
    if not paired and n_dep1 > 0:
        ion = 16.0 * n_i / (t_dep1 / n_dep1 * 1e-3) / 1e9                  # the ion sweeps (s = 1 sub-step per particle) are the bytes-bound launches
        roofline["ion_sweeps"] = {"achieved": ion, "peak": peak, "unit": "GB/s", "frac": ion / peak, "avg_launch_ms": t_dep1 / n_dep1}
    # second bound of the same kernel: the sub-step loop of the electrons is FP64 work, not bytes.  One face-to-face crossing is 10 FP64 + 12.5 other
    # instructions per warp; FP64 instructions issue every other cycle on sm_100 (scripts/micro/fp64_tput.cu), and the loop in isolation at 8 warps
    # per sub-partition runs at 29.3 cycles per crossing when a warp's lanes read one line of the table, as cell-sorted particles make them (38.7 with every
    # lane in another line: L1 gather) (scripts/micro/transit_bench2.cu, profiles/r2_micro_transit_bench2.txt) -- the floor quoted here.
    t_e_sweeps, n_e_sweeps = (t_pair, n_pair) if paired else (t_dep0, n_dep0)     # paired: the electrons share the launch with the ions
    if w["scheme"] != 1 and t_e_sweeps > 0:
        props = torch.cuda.get_device_properties(torch.cuda.current_device())
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
        floor = props.multi_processor_count * 4 * sm_hz / TRANSIT_FLOOR_CYCLES * 32.0
        roofline["issue_bound"] = {"what": "electron sub-steps per second inside the depositJ sweeps vs the measured issue floor of the crossing loop",
                                   "achieved": dep_sub_e / (t_e_sweeps * 1e-3), "peak": floor, "frac": dep_sub_e / (t_e_sweeps * 1e-3) / floor,
                                   "unit": "sub-steps/s", "floor_cycles_per_crossing_per_sm_subpartition": TRANSIT_FLOOR_CYCLES,
                                   "electron_substeps_per_particle_per_sweep": dep_sub_e / max(n_e_sweeps * n_e, 1)}

    line = None
    if rank == 0:
        cpu = None
        if world_size == 1 and not args.no_cpu:
            try:
                v, info = cpu_reference_run(w, 1, 1, args.cpu_seconds)
                cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
            except Exception as ex:                                  # the checker failing must not hide the measurement
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": w["name"], "particles_per_gpu": 2 * n_sp, "particles_total": 2 * n_sp * world_size, "cells": world.NumberXHalfNodes,
                           "del_t": del_t, "residual_evaluations_per_step": sum(iters) / len(iters),
                           # iterNumPicard is the count of the LAST solve of a step; a step whose solve stagnates is split (mod_nonLinSolvers.f90:162-196) and
                           # holds the sweeps of every attempt
                           "sweeps_per_step": sweeps / max(args.steps, 1),
                           "substeps_per_particle_per_sweep": substeps / max(particle_sweeps, 1),
                           "sort_every_steps": args.sort_every, "sort_every_steps_ions": args.sort_every * args.sort_ion_factor, "tridiag": "thomas" if args.thomas else ("pcr" if w["scheme"] == 2 else "scan (generators of the inverse)"),
                           "l2": "inputs larger than L2: %.2f GB particle stream per sweep per GPU vs 126 MB L2" % (32.0 * n_sp / 1e9),
                           "parallelism": "particles sharded over %d GPU(s); int64 fixed-point J grid all-reduced (%s) per residual evaluation; "
                                          "field solve replicated" % (world_size, "peer memory over NVLink, csrc/peer.cu" if (world_size > 1 and peer_on)
                                                                      else "NCCL")},
                # the metric counts sub-steps, and sub-steps per particle per sweep are physics (noise level, RF phase): particles x sweeps per
                # second says how fast the particles are swept whatever they do inside a sweep (what weak scaling measures here)
                "particle_sweeps_per_s": psweeps_all / (ms_max * 1e-3),
                "e2e": e2e, "e2e_resident": e2e_resident, "gpu_launches": n_launch, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "breakdown_ms_per_step": {"depositJ_kernels": dep_ms / args.steps, "moveParticles_kernel": t_move / args.steps,
                                          "compaction": t_comp / args.steps, "field_kernels": t_field / args.steps,
                                          "allreduce_J": t_ar / args.steps, "sort": t_sort / args.steps, "collisions": t_coll / args.steps}}
        print(json.dumps(line), flush=True)
    solver.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS), help="c5 (default, the bench line) | c1 | c2 | c3 | c4: the other BASELINE configs")
    ap.add_argument("--particles-per-species", type=int, default=0, help="override particles per species per GPU (debug)")
    ap.add_argument("--sort-every", type=int, default=None, help="cell-sort the electrons every this many steps inside the timed region (0 = never); default: the workload's own period -- 1 where a step holds tens of sweeps (c1, c3, c5), 5 where it holds two to four (c2, c4)")
    ap.add_argument("--sort-ion-factor", type=int, default=8, help="the ion species is sorted every sort-every x this many steps")
    ap.add_argument("--thomas", action="store_true", help="sequential Thomas solve (bit-parity mode) instead of PCR")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-peer", action="store_true", help="multi-GPU: all-reduce the J grid with NCCL instead of the peer-memory kernels")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU work budget of the cpu_baseline leg")
    ap.add_argument("--traffic-bytes-per-particle", type=float, default=None,
                    help="dram bytes per particle per launch of the dominant kernel from the committed ncu --set full capture")
    ap.add_argument("--option", nargs=2, action="append", metavar=("NAME", "VALUE"), help="ipic_set_option passthrough")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: note: fewer than 3 warm-up steps requested", file=sys.stderr)
    w = WORKLOADS[args.workload]
    if args.sort_every is None:
        args.sort_every = 5 if args.workload in ("c2", "c4") else 1
    if args.traffic_bytes_per_particle is None:
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json), if recorded
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                args.traffic_bytes_per_particle = json.load(f).get(args.workload, {}).get("dram_bytes_per_particle_per_launch")
        except Exception:
            args.traffic_bytes_per_particle = None
    if args.impl == "reference":
        return run_reference_arm(args, w)
    return run_cuda_arm(args, w)


if __name__ == "__main__":
    sys.exit(main())
