"""Run under torchrun with N >= 2 ranks (one GPU each):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Each rank uploads its thread slabs of the same oracle deck; the sharded depositJ / solvePotential must reproduce the oracle
(J to 1e-12, equal iteration counts) and, bit for bit, the single-GPU result (integer deposition + integer all-reduce)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import implicitpic1d_b200 as ip
from implicitpic1d_b200 import _lib, distributed as D, scheme as sch, simulation as sim
from implicitpic1d_b200.domain import readWorld
from common import DECKS, build_oracle, build_cuda_from_oracle, deck_del_t, relerr


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name in ("ngp_rf", "cic_rf", "ngp_periodic"):
        d = DECKS[name]
        o = build_oracle(d)
        del_t = deck_del_t(d)
        # single-GPU result on this rank's GPU (no communicator)
        w1, s1, p1 = build_cuda_from_oracle(d, o, tridiag_mode=_lib.TRIDIAG_THOMAS, device=local)
        ip.depositJ(s1, p1, w1, del_t)
        J_single = s1.J
        r1 = ip.solvePotential(s1, p1, w1, del_t, del_t, del_t, 0.0)
        st1 = ip.SolverState(s1)
        phi_single = s1.phi_f
        # diagnostics of the moved particles on one GPU: density dump arrays, histogram extrema and bins
        sch.loadParticleDensity(p1, w1, True)
        dens_single = [sch.particleDensityArray(p, w1) for p in p1]
        h1 = sim.Histograms(p1, 100)
        h1.accumulate(p1)
        s1.close()
        for mode in ("peer", "nccl"):
            # sharded; the J all-reduce through peer memory (csrc/peer.cu), then NCCL's: both must give the single-GPU bits
            world_dom = readWorld(d["scheme"], d["Nx"], d["L"], d.get("gridType", 0), d.get("del_x", 0.0), d["bcL"], d["bcR"], int(d["smooth"]),
                                  d["T_e"], d["n_ave"])
            solver = ip.potentialSolver(world_dom, d["Vl"], d["Vr"], d["rf"], 0.0, device=local)
            solver.set_tridiag_mode(_lib.TRIDIAG_THOMAS)
            peer_on = D.comm_init(solver, dist, rank, world, device="cuda", peer=(mode == "peer"))
            mine = D.shard_slabs(d["T"], world, rank)
            plist = []
            for sp in range(o.n_species):
                sc = o.species_scalars(sp)
                ps = o.phase_space(sp)[mine].copy()
                p = ip.Particle(sc["mass"], sc["q"], sc["w_p"], 0, ps.shape[1], "e" if sp == 0 else "ion", len(mine))
                p.set_phase_space(ps, o.N_p(sp)[mine])
                plist.append(p)
            ip.initializeSolver(solver, d["eps_a"], 0.0, 0, d["m"], d["beta"], d["maxIter"])
            solver.solveInitialPotential(plist, world_dom, 0.0)
            ip.depositJ(solver, plist, world_dom, del_t)
            J = solver.J
            o.depositJ(del_t)
            e_or = relerr(J, o.get("J"))
            bit = np.array_equal(J, J_single)
            r = ip.solvePotential(solver, plist, world_dom, del_t, del_t, del_t, 0.0)
            st = ip.SolverState(solver)
            bit_phi = np.array_equal(solver.phi_f, phi_single)
            # the density load all-reduces its integer grid inside the library: every rank holds the single-GPU arrays bit for bit;
            # histogram extrema / bins are combined over ranks by the host (max, min, sum) and equal the single-GPU ones
            sch.loadParticleDensity(plist, world_dom, True)
            bit_dens = all(np.array_equal(sch.particleDensityArray(p, world_dom), dens_single[k]) for k, p in enumerate(plist))
            h = sim.Histograms(plist, 100, dist=dist)
            h.accumulate(plist)
            VH = h.VHist.copy(); EH = h.EHist.copy()
            sim._allreduce_np(dist, VH); sim._allreduce_np(dist, EH)
            bit_hist = (np.array_equal(h.VMax, h1.VMax) and np.array_equal(h.EMax, h1.EMax) and np.array_equal(h.Emin, h1.Emin)
                        and np.allclose(VH, h1.VHist, rtol=0, atol=1e-6) and np.allclose(EH, h1.EHist, rtol=0, atol=1e-6))
            n_local = sum(p.count() for p in plist)
            n_all = D.allreduce_scalars(dist, [n_local], device="cuda")[0]
            good = (e_or < 1e-12) and bit and bit_phi and (st.iterNumPicard == st1.iterNumPicard) and bit_dens and bit_hist
            ok = ok and good
            print("rank %d %-12s [%s%s] J vs oracle %.2e | J bitwise == 1-GPU: %s | phi_f bitwise: %s | iters %d (1-GPU %d) | density bitwise: %s | histograms: %s | particles after move %d"
                  % (rank, name, mode, "" if mode == "nccl" or peer_on else ": peer windows unavailable, NCCL used", e_or, bit, bit_phi, st.iterNumPicard, st1.iterNumPicard, bit_dens, bit_hist, int(n_all)), flush=True)
            solver.close()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI-GPU CHECK", "OK" if flag.item() == 1.0 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
