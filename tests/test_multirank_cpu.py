"""world_size-2 gloo tests (CPU) of the N>1 host logic: slab sharding, unique-id exchange, and the property the multi-GPU
path rests on -- J of the whole = sum over ranks of J of the shards at identical fields and weights, so ONE all-reduce of the grid
per residual evaluation with a replicated field solve reproduces the single-process iterate."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from implicitpic1d_b200 import distributed as D
from common import DECKS, build_oracle, deck_del_t, relerr


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, deck, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = DECKS[deck]
        # 1. unique-id exchange (the id maker is injected: no NCCL on a CPU box)
        uid = D.exchange_unique_id(dist, rank, make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 2. every rank builds the same program state, keeps only its slabs
        o = build_oracle(d)
        del_t = deck_del_t(d)
        mine = D.shard_slabs(d["T"], world, rank)
        J_full = None
        if rank == 0:
            o.depositJ(del_t)
            J_full = o.get("J").copy()
        for sp in range(o.n_species):
            n = o.N_p(sp)
            keep = np.zeros_like(n)
            keep[mine] = n[mine]
            ps = o.phase_space(sp)
            for t in range(d["T"]):
                if t not in mine:
                    ps[t, :, :] = 0.0
            o.set_N_p(sp, keep)
        o.depositJ(del_t)                       # weights w_p were fixed from the GLOBAL particle count before the split
        J = torch.from_numpy(o.get("J").copy())
        dist.all_reduce(J)                      # the one collective of the path
        counts = D.allreduce_scalars(dist, [sum(int(o.N_p(sp).sum()) for sp in range(o.n_species))])
        if rank == 0:
            out["err"] = relerr(J.numpy(), J_full)
            out["count"] = counts[0]
            out["expect"] = 2 * d["ppt"] * d["T"]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("deck", ["ngp_rf", "cic_rf"])
def test_two_rank_J_allreduce_matches_single_process(deck):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), deck, out), nprocs=world, join=True)
    assert out["count"] == out["expect"]
    assert out["err"] < 1e-12, out["err"]       # same tolerance as the GPU path: floating-point order differs, integers would not


def _worker_diag(rank, world, port, out):
    """the rank combination of the time loop's diagnostics (simulation.Histograms): histogram extrema by max / min, bins by sum -- what
    the reference's OpenMP reduction(max: VMax, EMax) reduction(min: Emin) reduction(+: VHist, EHist) clauses do over threads
    (mod_simulation.f90:442, :479), here over ranks with the oracle's per-shard results"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from implicitpic1d_b200 import simulation as sim
        d = DECKS["ngp_rf"]
        o = build_oracle(d)
        full_ex = o.velocityExtrema(0)
        VH_full, EH_full = np.zeros(200), np.zeros(100)
        o.velocityHistograms(0, 100, full_ex[0], full_ex[2], full_ex[1], VH_full, EH_full)
        mine = D.shard_slabs(d["T"], world, rank)
        n = o.N_p(0)
        keep = np.zeros_like(n)
        keep[mine] = n[mine]
        o.set_N_p(0, keep)
        ex = o.velocityExtrema(0)
        VMax, EMax, Emin = np.array([ex[0]]), np.array([ex[1]]), np.array([ex[2]])
        sim._allreduce_np(dist, VMax, dist.ReduceOp.MAX)
        sim._allreduce_np(dist, EMax, dist.ReduceOp.MAX)
        sim._allreduce_np(dist, Emin, dist.ReduceOp.MIN)
        VH, EH = np.zeros(200), np.zeros(100)
        o.velocityHistograms(0, 100, VMax[0], Emin[0], EMax[0], VH, EH)
        sim._allreduce_np(dist, VH)
        sim._allreduce_np(dist, EH)
        if rank == 0:
            out["extrema"] = bool(VMax[0] == full_ex[0] and EMax[0] == full_ex[1] and Emin[0] == full_ex[2])
            out["hist"] = max(float(np.max(np.abs(VH - VH_full))), float(np.max(np.abs(EH - EH_full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_histogram_reductions_match_single_process():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_diag, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["extrema"]
    assert out["hist"] < 1e-9                    # sums of weights in [0, 1]: only the summation order differs


def test_shard_helpers():
    for T in (1, 2, 7, 32):
        for w in (1, 2, 3, 8):
            got = sum((D.shard_slabs(T, w, r) for r in range(w)), [])
            assert got == list(range(T))
            sizes = [len(D.shard_slabs(T, w, r)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    assert sum(D.shard_count(10**9 + 3, 8, r) for r in range(8)) == 10**9 + 3
    assert D.global_weight(1.6e18, 0.067, 1.536e6) == 1.6e18 * 0.067 / 1.536e6
