"""Per-rank SEM setup that scales to the benchmark sizes (25 M nodes per GPU, 8 GPUs).

``sem_setup`` (setup.py) mirrors the reference literally: every rank's index arrays are gathered
and the ownership / assembler rules run over all of them.  That is the right oracle for tests but
not for 8 x 25 M nodes.  Here each rank builds ONLY its own mesh and metrics; the ownership and
AssemblerCache rules (mesh.jl:3560-3610, mpi_communications.jl:75-234) are evaluated on the
*interface candidates* alone -- the nodes on sub-box planes that face another part -- which is
exact because a node that no other rank lists is never contested and never enters a send list.
The candidates of every rank follow from the closed-form numbering of the structured box, so no
communication is needed for the integer maps; the floating-point setup assemblies (mass, normals,
IC conditioning) go through ``distributed.assemble_dist``.

tests/test_scalable_cpu.py checks these lists bit-exactly against the literal all-ranks path.
"""
from __future__ import annotations

import numpy as np

from . import basis as _basis
from .mesh import BoxSpec, _numbering_2d, _numbering_3d, part_subbox, structured_box
from .metrics import boundary_normals, build_mass_inverse, build_mass_local, build_metric_terms
from .partition import AssemblerLists, find_gip_owner_all, setup_assembler_all
from .setup import SEM, _snap_normals, sem_setup

__all__ = ["sem_setup_rank", "interface_candidates", "conformity4ncf_q_rank"]


def interface_candidates(spec: BoxSpec, nranks: int, rank: int):
    """(local ids, global ids), 0-based / 1-based resp., ascending local id, of the nodes of ``rank``
    lying on a sub-box plane shared with another part."""
    p = spec.nop
    sub = part_subbox(spec, nranks, rank)
    ex0, ex1, ey0, ey1 = sub
    NX, NY = spec.nel[0], spec.nel[1]
    nx, ny = ex1 - ex0, ey1 - ey0
    if spec.nsd == 3:
        NZ = spec.nel[2]
        ggid, _, _ = _numbering_3d(NX, NY, NZ, p)
        lgid, _, _ = _numbering_3d(nx, ny, NZ, p)
        I, J, K = np.arange(nx * p + 1), np.arange(ny * p + 1), np.arange(NZ * p + 1)
        planes = []
        if ex0 > 0:
            planes.append(np.meshgrid([0], J, K, indexing="ij"))
        if ex1 < NX:
            planes.append(np.meshgrid([nx * p], J, K, indexing="ij"))
        if ey0 > 0:
            planes.append(np.meshgrid(I, [0], K, indexing="ij"))
        if ey1 < NY:
            planes.append(np.meshgrid(I, [ny * p], K, indexing="ij"))
        if not planes:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        Ic = np.concatenate([g[0].reshape(-1) for g in planes])
        Jc = np.concatenate([g[1].reshape(-1) for g in planes])
        Kc = np.concatenate([g[2].reshape(-1) for g in planes])
        lid = lgid(Ic, Jc, Kc)
        gid = ggid(Ic + ex0 * p, Jc + ey0 * p, Kc) + 1
    else:
        ggid, _, _ = _numbering_2d(NX, NY, p)
        lgid, _, _ = _numbering_2d(nx, ny, p)
        I, J = np.arange(nx * p + 1), np.arange(ny * p + 1)
        lines = []
        if ex0 > 0:
            lines.append((np.zeros_like(J), J))
        if ex1 < NX:
            lines.append((np.full_like(J, nx * p), J))
        if ey0 > 0:
            lines.append((I, np.zeros_like(I)))
        if ey1 < NY:
            lines.append((I, np.full_like(I, ny * p)))
        if not lines:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        Ic = np.concatenate([l[0] for l in lines])
        Jc = np.concatenate([l[1] for l in lines])
        lid = lgid(Ic, Jc)
        gid = ggid(Ic + ex0 * p, Jc + ey0 * p) + 1
    lid, first = np.unique(lid, return_index=True)
    return lid.astype(np.int64), gid[first].astype(np.int64)


def _lists_for_rank(spec, nranks, rank):
    cands = [interface_candidates(spec, nranks, r) for r in range(nranks)]
    gids = [c[1] for c in cands]
    owners = find_gip_owner_all(gids)
    cl = setup_assembler_all(gids, owners, only_rank=rank)[0]
    lid = cands[rank][0]

    def to_local(v):
        return lid[np.asarray(v, np.int64) - 1] + 1 if len(v) else np.zeros(0, np.int64)

    lists = AssemblerLists(rank=rank, nranks=nranks,
                           send_i=[to_local(v) for v in cl.send_i],
                           recv_idx=[to_local(v) for v in cl.recv_idx],
                           recvback_idx=[to_local(v) for v in cl.recvback_idx],
                           send_gid=cl.send_gid)
    return lists, lid, owners[rank]


def sem_setup_rank(spec: BoxSpec, rank: int = 0, nranks: int = 1, group="auto"):
    """This rank's :class:`SEM` bundle.  Non-periodic boxes at any rank count; periodic boxes fall
    back to the literal all-ranks path (fine up to a few million nodes)."""
    if any(spec.periodic[:spec.nsd]):
        return sem_setup(spec, nranks)[rank]
    from ..distributed import assemble_dist, host_group
    if group == "auto":
        group = host_group() if nranks > 1 else None
    bs = _basis.build_basis(spec.nop)
    sub = part_subbox(spec, nranks, rank) if nranks > 1 else None
    mesh = structured_box(spec, bs["xi"], sub=sub, rank=rank, nranks=nranks)
    if nranks > 1:
        lists, cand_lid, cand_owner = _lists_for_rank(spec, nranks, rank)
        mesh.gip2owner[cand_lid] = cand_owner
    else:
        e = [np.zeros(0, np.int64)]
        lists = AssemblerLists(rank=0, nranks=1, send_i=list(e), recv_idx=list(e), recvback_idx=list(e), send_gid=list(e))
    met = build_metric_terms(mesh, bs)
    M = build_mass_local(mesh, bs, met["Je"])
    if nranks > 1:
        assemble_dist(M, lists, group)                          # DSS_global_mass!
    nx, ny, nz = boundary_normals(mesh)
    P = mesh.poin_in_bdy_face - 1
    nf = P.shape[0]
    sums = np.zeros((mesh.npoin, spec.nsd), order="F")
    if nf:
        flatP = np.ascontiguousarray(P.reshape(nf, -1)).reshape(-1)
        for d, c in enumerate([nx, ny, nz][:spec.nsd]):
            sums[:, d] = np.bincount(flatP, weights=np.ascontiguousarray(c.reshape(nf, -1)).reshape(-1), minlength=mesh.npoin)
    if nranks > 1:
        assemble_dist(sums, lists, group)                       # DSS_global_normals!
    _snap_normals(mesh, sums, [c for c in (nx, ny, nz) if c is not None])
    return SEM(mesh=mesh, basis=bs, metrics=met, M=M, Minv=build_mass_inverse(M), nx=nx, ny=ny, nz=nz, asm=lists)


def conformity4ncf_q_rank(sem, q, neqs, group="auto"):
    """conformity4ncf_q! (Adaptivity/Projection.jl:2919-2970) for one rank: q <- Minv * DSS(ωJ q)."""
    from ..distributed import assemble_dist, host_group
    m, om = sem.mesh, sem.basis["omega"]
    Je = sem.metrics["Je"]
    if m.nsd == 3:
        w = (om[:, None, None] * om[None, :, None]) * om[None, None, :]
        wJ = (w[None] * Je).reshape(m.nelem, -1)
        conn = (m.connijk - 1).reshape(m.nelem, -1)
    else:
        w = om[:, None] * om[None, :]
        wJ = (w[None] * Je[:, :, :, 0]).reshape(m.nelem, -1)
        conn = (m.connijk[:, :, :, 0] - 1).reshape(m.nelem, -1)
    t = np.zeros((m.npoin, neqs), order="F")
    for ieq in range(neqs):
        t[:, ieq] = np.bincount(conn.reshape(-1), weights=(wJ * q[:, ieq][conn]).reshape(-1), minlength=m.npoin)
    if m.nranks > 1:
        if group == "auto":
            group = host_group()
        assemble_dist(t, sem.asm, group)
    for ieq in range(neqs):
        q[:, ieq] = sem.Minv * t[:, ieq]
