"""Multi-GPU plumbing: one process per GPU, particles sharded by thread slab, one all-reduce of the integer J/rho grids per
deposit (inside the library, NCCL), field solve replicated.  Pattern of the reference's only distributed code,
PIC1D/ExplicitPIC_Cpp/src/potential_solver.cpp:142 (MPI_Allreduce of the grid, replicated solve).

torch.distributed is used for the rendezvous only (broadcast of the 128-byte NCCL unique id, scalar reductions of diagnostics)."""
import ctypes as C
import numpy as np
from . import _lib


def shard_slabs(numThread, world_size, rank):
    """Thread slabs of the reference's phaseSpace(4, finalIdx, numThread) owned by `rank`: contiguous blocks, sizes differing by <= 1."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(numThread, world_size)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def shard_count(n_total, world_size, rank):
    """particles of one species held by `rank` when n_total are dealt evenly"""
    base, extra = divmod(int(n_total), world_size)
    return base + (1 if rank < extra else 0)


def global_weight(n_ave, L_domain, n_total_all_ranks):
    """w_p = n_ave * L / SUM(N_p) with the sum over ALL ranks (mod_particle.f90:76-82)"""
    return n_ave * L_domain / float(n_total_all_ranks)


def exchange_unique_id(dist, rank, make_id=None, device=None):
    """rank 0 creates the NCCL unique id (ipic_nccl_unique_id), everybody receives it through torch.distributed."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        if make_id is None:
            raw = (C.c_char * 128)()
            _lib.check(_lib.load().ipic_nccl_unique_id(raw))
            data = bytes(raw.raw)
        else:
            data = bytes(make_id())
        buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def exchange_peer_handles(dist, world_size, my_handle, device=None):
    """all-gather of the 64-byte CUDA IPC handles of the ranks' peer windows (ipic_peer_export), in rank order"""
    import torch
    mine = torch.frombuffer(bytearray(my_handle), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    out = [torch.zeros_like(mine) for _ in range(world_size)]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def peer_init(solver, dist, rank, world_size, device=None):
    """One node: map every rank's window so that the J all-reduce runs through peer memory (include/ipic1d_b200.h: ipic_peer_export /
    ipic_peer_import).  Returns True when every rank mapped every window; otherwise all ranks stay on NCCL (the choice must be the same
    everywhere: the ranks take the same decisions or they dead-lock)."""
    import torch
    if world_size <= 1:
        return False
    L = _lib.load()
    h = (C.c_char * 64)()
    ok = 1
    try:
        _lib.check(L.ipic_peer_export(solver._h, h))
    except _lib.IpicError:
        ok = 0
    handles = exchange_peer_handles(dist, world_size, bytes(h.raw), device=device)
    if ok:
        try:
            _lib.check(L.ipic_peer_import(solver._h, world_size, rank, C.c_char_p(handles)))
        except _lib.IpicError:
            ok = 0
    t = torch.tensor([ok], dtype=torch.int32)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    all_ok = bool(int(t.cpu().item()))
    solver.set_option("peer_reduce", 1.0 if all_ok else 0.0)
    return all_ok


def comm_init(solver, dist, rank, world_size, device=None, peer=True):
    """ipic_comm_init on every rank with the id made on rank 0; then (one node) the peer-memory windows.  Returns True when the J
    all-reduce runs through peer memory, False when it is NCCL's."""
    if world_size <= 1:
        return False
    uid = exchange_unique_id(dist, rank, device=device)
    _lib.check(_lib.load().ipic_comm_init(solver._h, world_size, rank, C.c_char_p(uid)))
    if peer and dist.get_backend() == "nccl":
        return peer_init(solver, dist, rank, world_size, device=device)
    return False


def allreduce_scalars(dist, values, device=None):
    """sum of per-rank diagnostics (KE, momentum, wall losses, sub-step counters): the small all-reduces of SURVEY.md 8e"""
    import torch
    t = torch.tensor(np.asarray(values, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()
