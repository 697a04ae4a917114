"""``sem_setup`` for synthetic boxes: everything ``rhs!`` reads, per rank.

Mirrors the order of src/kernel/infrastructure/sem_setup.jl:132-470 (mesh ->
LGL/basis -> metrics -> matrix_wrapper: mass, global mass DSS, normals, Minv) and
the IC conditioning of params_setup.jl:259-297 (conformity4ncf_q! applied to qn
and qe).  This is setup, executed once on the host with numpy; the per-stage hot
path lives in csrc/.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import basis as _basis
from .mesh import BoxSpec, Mesh, part_subbox, structured_box
from .metrics import boundary_normals, build_mass_inverse, build_mass_local, build_metric_terms
from .partition import AssemblerLists, find_gip_owner_all, restructure4periodicity_all, setup_assembler_all

__all__ = ["SEM", "sem_setup", "assemble_host", "conformity4ncf_q_host"]

METRIC_NAMES_3D = ["dξdx", "dξdy", "dξdz", "dηdx", "dηdy", "dηdz", "dζdx", "dζdy", "dζdz", "Je"]
METRIC_NAMES_2D = ["dξdx", "dξdy", "dηdx", "dηdy", "Je"]


@dataclass
class SEM:
    mesh: Mesh
    basis: dict
    metrics: dict
    M: np.ndarray
    Minv: np.ndarray
    nx: np.ndarray
    ny: np.ndarray
    nz: np.ndarray | None
    asm: AssemblerLists
    extra: dict = field(default_factory=dict)

    @property
    def metric_list(self):
        names = METRIC_NAMES_3D if self.mesh.nsd == 3 else METRIC_NAMES_2D
        return [self.metrics[k] for k in names]


def assemble_host(arrays, lists):
    """assemble_mpi! (mpi_communications.jl:260-338) for all ranks at once, on the host.

    ``arrays[r]`` is rank r's [npoin] or [npoin, m] array (modified in place).  Owners add
    the received contributions in ascending sender rank, list order; the owner's sum is
    then copied back to every sender.
    """
    R = len(arrays)
    a2 = [a.reshape(a.shape[0], -1) for a in arrays]
    bufs = {}
    for src in range(R):
        for dst in lists[src].active_send_ranks:
            bufs[(src, dst)] = a2[src][lists[src].send_i[dst] - 1, :].copy()
    for dst in range(R):
        for src in lists[dst].active_recv_ranks:
            idx = lists[dst].recv_idx[src] - 1
            buf = bufs[(src, dst)]
            for j in range(a2[dst].shape[1]):
                np.add.at(a2[dst][:, j], idx, buf[:, j])       # sequential, list order
    back = {}
    for dst in range(R):
        for src in lists[dst].active_recv_ranks:
            back[(dst, src)] = a2[dst][lists[dst].recv_idx[src] - 1, :].copy()
    for src in range(R):
        for dst in lists[src].active_send_ranks:
            a2[src][lists[src].recvback_idx[dst] - 1, :] = back[(dst, src)]


def conformity4ncf_q_host(sems, qs, neqs):
    """conformity4ncf_q! with ladapt=false (Adaptivity/Projection.jl:2919-2970 3D,
    :2870-2916 2D): q <- Minv * DSS(ωJ q) on the first ``neqs`` columns."""
    tmps = []
    for sem, q in zip(sems, qs):
        m, om = sem.mesh, sem.basis["omega"]
        Je = sem.metrics["Je"]
        if m.nsd == 3:
            w = (om[:, None, None] * om[None, :, None]) * om[None, None, :]     # ωij*ω[k]
            wJ = (w[None] * Je).reshape(m.nelem, -1)
            conn = (m.connijk - 1).reshape(m.nelem, -1)
        else:
            w = om[:, None] * om[None, :]
            wJ = (w[None] * Je[:, :, :, 0]).reshape(m.nelem, -1)
            conn = (m.connijk[:, :, :, 0] - 1).reshape(m.nelem, -1)
        t = np.zeros((m.npoin, neqs), order="F")
        for ieq in range(neqs):
            t[:, ieq] = np.bincount(conn.reshape(-1), weights=(wJ * q[:, ieq][conn]).reshape(-1),
                                    minlength=m.npoin)
        tmps.append(t)
    assemble_host(tmps, [s.asm for s in sems])
    for sem, q, t in zip(sems, qs, tmps):
        for ieq in range(neqs):
            q[:, ieq] = sem.Minv * t[:, ieq]


def sem_setup(spec: BoxSpec, nranks: int = 1, ranks=None):
    """Build the SEM bundle of every rank of an ``nranks``-way xy partition.

    Returns a list with one :class:`SEM` per rank.  Ownership, periodic merging and
    the assembler lists need all ranks' index arrays (the reference gathers them on
    rank 0), so all local meshes are generated here; metrics are only built for the
    ranks listed in ``ranks`` (default all).
    """
    bs = _basis.build_basis(spec.nop)
    xi = bs["xi"]
    meshes = []
    for r in range(nranks):
        sub = part_subbox(spec, nranks, r) if nranks > 1 else None
        meshes.append(structured_box(spec, xi, sub=sub, rank=r, nranks=nranks))
    owners = find_gip_owner_all([m.ip2gip for m in meshes])
    for m, o in zip(meshes, owners):
        m.gip2owner = o
    for d, ax in enumerate("xyz"[:spec.nsd]):
        if spec.periodic[d]:
            restructure4periodicity_all(meshes, "periodic" + ax)
    lists = setup_assembler_all([m.ip2gip for m in meshes], [m.gip2owner for m in meshes])
    want = range(nranks) if ranks is None else ranks
    sems = {}
    Ms = []
    mets = []
    for r in range(nranks):
        met = build_metric_terms(meshes[r], bs)
        mets.append(met)
        Ms.append(build_mass_local(meshes[r], bs, met["Je"]))
    assemble_host(Ms, lists)                                   # DSS_global_mass!
    normals = []
    for r in range(nranks):
        nx, ny, nz = boundary_normals(meshes[r])
        normals.append([nx, ny, nz])
    # DSS_global_normals!: nodal sums need the inter-rank assembly as well
    nsd = spec.nsd
    sums = []
    for r in range(nranks):
        m = meshes[r]
        P = m.poin_in_bdy_face - 1
        nf = P.shape[0]
        s = np.zeros((m.npoin, nsd), order="F")
        if nf:
            flatP = np.ascontiguousarray(P.reshape(nf, -1)).reshape(-1)
            for d in range(nsd):
                s[:, d] = np.bincount(flatP, weights=np.ascontiguousarray(normals[r][d].reshape(nf, -1)).reshape(-1),
                                      minlength=m.npoin)
        sums.append(s)
    assemble_host(sums, lists)
    out = []
    for r in range(nranks):
        nx, ny, nz = normals[r]
        _snap_normals(meshes[r], sums[r], [c for c in (nx, ny, nz) if c is not None])
        out.append(SEM(mesh=meshes[r], basis=bs, metrics=mets[r], M=Ms[r], Minv=build_mass_inverse(Ms[r]),
                       nx=nx, ny=ny, nz=nz, asm=lists[r]))
    return out if ranks is None else [out[r] for r in want]


def _snap_normals(mesh, normals, comps):
    """Second half of DSS_global_normals! (element_matrices.jl:1083-1113)."""
    P = mesh.poin_in_bdy_face - 1
    nf = P.shape[0]
    if nf == 0:
        return
    nsd = mesh.nsd
    sq = normals[P, 0] ** 2 + normals[P, 1] ** 2
    if nsd == 3:
        sq = sq + normals[P, 2] ** 2
    mag = np.sqrt(sq)
    ok = mag > 0
    safe = np.where(ok, mag, 1.0)
    tags = np.array(mesh.bdy_face_type)
    skip = ["periodicx", "periodicy", "periodicz"]
    for d, c in enumerate(comps):
        norm_d = np.where(ok, normals[P, d] / safe, 0.0)
        msk = np.abs(c - norm_d) < 0.25
        msk &= (tags != skip[d]).reshape((nf,) + (1,) * (c.ndim - 1))
        c[msk] = norm_d[msk]
    sq = comps[0] ** 2 + comps[1] ** 2
    if nsd == 3:
        sq = sq + comps[2] ** 2
    mag = np.sqrt(sq)
    ok = mag > 0
    safe = np.where(ok, mag, 1.0)
    for c in comps:
        c[...] = np.where(ok, c / safe, c)
