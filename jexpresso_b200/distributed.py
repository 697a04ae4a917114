"""torch.distributed plumbing for the host-side (setup-time) interface assembly.

``assemble_dist`` is assemble_mpi! (src/kernel/mpi/mpi_communications.jl:260-338) over a
torch.distributed group with host buffers: it is what the reference runs through MPI for
DSS_global_mass!, DSS_global_normals! and conformity4ncf_q! at setup.  The per-stage exchange of
the RHS never comes through here -- that one runs on the GPUs over NCCL inside libjexrhs
(jexrhs.cu: assemble()).  Works on the gloo backend (CPU tests, world_size 2) and, on a GPU box,
on a gloo side-group next to the NCCL world group.
"""
from __future__ import annotations

import numpy as np

__all__ = ["assemble_dist", "host_group", "effective_delta_dist"]

_HOST_GROUP = None


def host_group():
    """A gloo group for CPU tensors (the default group when it already is gloo)."""
    global _HOST_GROUP
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    if dist.get_backend() == "gloo":
        return dist.group.WORLD
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.new_group(backend="gloo")
    return _HOST_GROUP


def _exchange(out_bufs, in_shapes, rank, group):
    """out_bufs: {peer: array}; in_shapes: {peer: shape}.  Returns {peer: array} received."""
    import torch
    import torch.distributed as dist
    got = {}
    if rank in out_bufs and rank in in_shapes:
        got[rank] = out_bufs[rank].copy()                     # the reference's MPI self-send
    ops, keep = [], []
    for peer in sorted(in_shapes):
        if peer == rank:
            continue
        t = torch.empty(in_shapes[peer], dtype=torch.float64)
        keep.append((peer, t))
        ops.append(dist.P2POp(dist.irecv, t, peer, group=group))
    for peer in sorted(out_bufs):
        if peer == rank:
            continue
        t = torch.from_numpy(np.ascontiguousarray(out_bufs[peer]))
        keep.append((None, t))
        ops.append(dist.P2POp(dist.isend, t, peer, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for peer, t in keep:
        if peer is not None:
            got[peer] = t.numpy()
    return got


def assemble_dist(a, lists, group=None):
    """In-place interface sum of ``a`` ([npoin] or [npoin, m]) for this rank's AssemblerLists."""
    a2 = a.reshape(a.shape[0], -1)
    m = a2.shape[1]
    rank = lists.rank
    send = {r: a2[lists.send_i[r] - 1, :].copy() for r in lists.active_send_ranks}
    recv_shapes = {r: (len(lists.recv_idx[r]), m) for r in lists.active_recv_ranks}
    got = _exchange(send, recv_shapes, rank, group)
    for src in sorted(got):                                   # ascending sender rank, list order
        idx = lists.recv_idx[src] - 1
        for j in range(m):
            np.add.at(a2[:, j], idx, got[src][:, j])
    back = {r: a2[lists.recv_idx[r] - 1, :].copy() for r in lists.active_recv_ranks}
    back_shapes = {r: (len(lists.send_i[r]), m) for r in lists.active_send_ranks}
    got = _exchange(back, back_shapes, rank, group)
    for r, buf in got.items():
        a2[lists.recvback_idx[r] - 1, :] = buf
    return a


def effective_delta_dist(mesh, group=None):
    """mesh.Δeffective_l of a partitioned mesh: compute_element_size_driver takes MPI.Allreduce(maximum(Δelem), MAX) before it
    divides by nop (src/kernel/mesh/mesh.jl:5621-5632) -- every rank must hand the same value to jx_set_sgs."""
    import torch
    import torch.distributed as dist
    from .sem.mesh import element_sizes
    t = torch.tensor([float(element_sizes(mesh).max())], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group if group is not None else host_group())
    return float(t.item()) / mesh.nop
