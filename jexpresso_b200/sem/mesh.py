"""Synthetic structured hex/quad SEM meshes in the reference's array layout.

The reference reads gmsh files through Gridap; neither exists here, and the node
numbering Gridap produces is not pinned by any reference test.
What the RHS path consumes is only the *arrays* (``St_mesh`` fields,
src/kernel/mesh/meshStructs.jl:21-244):

  connijk           Int64[nelem, ngl, ngl, ngl]  (2D: [nelem, ngl, ngl, 1]), 1-based node ids,
                    Julia column-major => element index fastest (mesh.jl:2175, :2109)
  x, y, z / coords  Float64[npoin] / Float64[nsd, npoin]
  poin_in_bdy_face  Int64[nfaces_bdy, ngl, ngl]  (2D: poin_in_bdy_edge Int64[nedges_bdy, ngl])
  bdy_face_type     list of tag strings ("free_slip", "periodicx", ...)
  bdy_face_in_elem  Int64[nfaces_bdy]
  ip2gip            Int64[npoin] local -> global node id (1-based)
  gip2owner         Int64[npoin] owning rank (0-based) of each local node

All numpy arrays returned here are Fortran-ordered so their memory image is the
one Julia would hold, and node ids are 1-based exactly as the C-ABI expects them.

Element-local orientation follows the corner map of mesh.jl:2194-2201 for a
lexicographic (Gridap) hexahedron: local i runs along -x, j along +z, k along +y
(2D, mesh.jl:2128-2131: i along +y, j along -x).  Global numbering is
vertices, then edge-interior, face-interior, volume-interior nodes -- the same
category order the reference uses when inserting high-order nodes
(mesh.jl:4084+, :4555+, :4944+).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

__all__ = ["Mesh", "BoxSpec", "structured_box", "compute_xy_partition", "part_subbox", "element_sizes", "effective_delta_l"]


@dataclass
class BoxSpec:
    nsd: int
    nel: tuple            # (nx, ny[, nz]) global element counts
    nop: int
    lo: tuple             # (xmin, ymin[, zmin])
    hi: tuple
    periodic: tuple = (False, False, False)
    tags: dict = field(default_factory=dict)   # side name -> tag, default "free_slip"
    warp: float = 0.0     # amplitude (fraction of the element size) of the smooth interior warp
    wall_aligned: tuple = ()   # 3D, "ymin" only: the faces of that side are listed in the element's own (i, j) order at local
                               # k = 1 (poin_in_bdy_face[f, i, j] = connijk[e, i, j, 1]) -- what the wall-model loop of
                               # build_custom_bcs_neumann! assumes when it walks connijk[e, i, j, ifirst_wall_node] (BCs.jl:703-717)


@dataclass
class Mesh:
    nsd: int
    nop: int
    ngl: int
    nelem: int
    npoin: int
    npoin_linear: int
    gnpoin: int
    gnelem: int
    connijk: np.ndarray
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    coords: np.ndarray
    poin_in_bdy_face: np.ndarray   # 3D: [nfb,n,n]; 2D: poin_in_bdy_edge [neb,n]
    bdy_face_type: list
    bdy_face_in_elem: np.ndarray
    ip2gip: np.ndarray
    gip2owner: np.ndarray
    el2gel: np.ndarray
    xmin: float
    xmax: float
    ymin: float
    ymax: float
    zmin: float
    zmax: float
    rank: int = 0
    nranks: int = 1
    spec: BoxSpec | None = None
    sub: tuple | None = None

    @property
    def poin_in_bdy_edge(self):
        return self.poin_in_bdy_face

    @property
    def nfaces_bdy(self):
        return self.poin_in_bdy_face.shape[0]


# ---------------------------------------------------------------------------
# closed-form global numbering of the tensor GLL grid of a structured box
# ---------------------------------------------------------------------------
def _numbering_3d(nx, ny, nz, p):
    nvx, nvy, nvz = nx + 1, ny + 1, nz + 1
    nv = nvx * nvy * nvz
    q = p - 1
    nEx, nEy, nEz = nx * nvy * nvz, nvx * ny * nvz, nvx * nvy * nz
    nFxy, nFxz, nFyz = nx * ny * nvz, nx * nvy * nz, nvx * ny * nz
    off_ex = nv
    off_ey = off_ex + nEx * q
    off_ez = off_ey + nEy * q
    off_fxy = off_ez + nEz * q
    off_fxz = off_fxy + nFxy * q * q
    off_fyz = off_fxz + nFxz * q * q
    off_vol = off_fyz + nFyz * q * q
    total = off_vol + nx * ny * nz * q ** 3

    def gid(I, J, K):
        """0-based node id of GLL grid point (I,J,K); arrays broadcast."""
        I, J, K = np.broadcast_arrays(np.asarray(I, np.int64), np.asarray(J, np.int64), np.asarray(K, np.int64))
        ex, rx = np.divmod(I, p)
        ey, ry = np.divmod(J, p)
        ez, rz = np.divmod(K, p)
        a, b, c = rx == 0, ry == 0, rz == 0
        out = np.empty(I.shape, np.int64)
        m = a & b & c
        out[m] = ex[m] + nvx * (ey[m] + nvy * ez[m])
        m = ~a & b & c
        out[m] = off_ex + (ex[m] + nx * (ey[m] + nvy * ez[m])) * q + (rx[m] - 1)
        m = a & ~b & c
        out[m] = off_ey + (ex[m] + nvx * (ey[m] + ny * ez[m])) * q + (ry[m] - 1)
        m = a & b & ~c
        out[m] = off_ez + (ex[m] + nvx * (ey[m] + nvy * ez[m])) * q + (rz[m] - 1)
        m = ~a & ~b & c
        out[m] = off_fxy + (ex[m] + nx * (ey[m] + ny * ez[m])) * q * q + (rx[m] - 1) + q * (ry[m] - 1)
        m = ~a & b & ~c
        out[m] = off_fxz + (ex[m] + nx * (ey[m] + nvy * ez[m])) * q * q + (rx[m] - 1) + q * (rz[m] - 1)
        m = a & ~b & ~c
        out[m] = off_fyz + (ex[m] + nvx * (ey[m] + ny * ez[m])) * q * q + (ry[m] - 1) + q * (rz[m] - 1)
        m = ~a & ~b & ~c
        out[m] = off_vol + (ex[m] + nx * (ey[m] + ny * ez[m])) * q ** 3 + (rx[m] - 1) + q * ((ry[m] - 1) + q * (rz[m] - 1))
        return out

    return gid, total, nv


def _numbering_2d(nx, ny, p):
    nvx, nvy = nx + 1, ny + 1
    nv = nvx * nvy
    q = p - 1
    nEx, nEy = nx * nvy, nvx * ny
    off_ex = nv
    off_ey = off_ex + nEx * q
    off_f = off_ey + nEy * q
    total = off_f + nx * ny * q * q

    def gid(I, J):
        I, J = np.broadcast_arrays(np.asarray(I, np.int64), np.asarray(J, np.int64))
        ex, rx = np.divmod(I, p)
        ey, ry = np.divmod(J, p)
        a, b = rx == 0, ry == 0
        out = np.empty(I.shape, np.int64)
        m = a & b
        out[m] = ex[m] + nvx * ey[m]
        m = ~a & b
        out[m] = off_ex + (ex[m] + nx * ey[m]) * q + (rx[m] - 1)
        m = a & ~b
        out[m] = off_ey + (ex[m] + nvx * ey[m]) * q + (ry[m] - 1)
        m = ~a & ~b
        out[m] = off_f + (ex[m] + nx * ey[m]) * q * q + (rx[m] - 1) + q * (ry[m] - 1)
        return out

    return gid, total, nv


def _axis_coords(n_el, p, lo, hi, xi):
    """Coordinates of the (n_el*p+1) GLL grid points along one axis."""
    h = (hi - lo) / n_el
    e = np.arange(n_el)
    c = lo + h * e[:, None] + 0.5 * h * (xi[None, :p] + 1.0)
    out = np.empty(n_el * p + 1)
    out[:-1] = c.reshape(-1)
    out[-1] = hi
    for k in range(n_el):            # element corners exactly on the uniform grid
        out[k * p] = lo + h * k
    return out


def _bump(s):
    return s * (1.0 - s)             # exactly 0 at s=0 and s=1


# ---------------------------------------------------------------------------
def compute_xy_partition(cx, cy, nparts):
    """Element -> part (1-based) by x-y centroid bins (mesh.jl:1513-1533)."""
    cx = np.asarray(cx, float)
    cy = np.asarray(cy, float)
    lx = cx.max() - cx.min() + 1e-10
    ly = cy.max() - cy.min() + 1e-10
    divisors = np.array([d for d in range(1, nparts + 1) if nparts % d == 0])
    target_nx = np.sqrt(nparts * lx / ly)
    nx = int(divisors[np.argmin(np.abs(divisors - target_nx))])
    ny = nparts // nx
    xi = np.clip(np.floor((cx - cx.min()) / lx * nx).astype(np.int64), 0, nx - 1)
    yi = np.clip(np.floor((cy - cy.min()) / ly * ny).astype(np.int64), 0, ny - 1)
    return xi * ny + yi + 1, nx, ny


def part_subbox(spec: BoxSpec, nparts: int, rank: int):
    """Structured element range (ex0,ex1,ey0,ey1) owned by ``rank`` under the
    reference's xy-bin partitioner applied to the uniform element centroids."""
    nx, ny = spec.nel[0], spec.nel[1]
    hx = (spec.hi[0] - spec.lo[0]) / nx
    hy = (spec.hi[1] - spec.lo[1]) / ny
    cx = spec.lo[0] + hx * (np.arange(nx) + 0.5)
    cy = spec.lo[1] + hy * (np.arange(ny) + 0.5)
    CX, CY = np.meshgrid(cx, cy, indexing="ij")
    part, pnx, pny = compute_xy_partition(CX.reshape(-1), CY.reshape(-1), nparts)
    part = part.reshape(nx, ny)
    mine = np.argwhere(part == rank + 1)
    if mine.size == 0:
        return (0, 0, 0, 0)
    ex0, ey0 = mine.min(axis=0)
    ex1, ey1 = mine.max(axis=0) + 1
    assert (part[ex0:ex1, ey0:ey1] == rank + 1).all() and mine.shape[0] == (ex1 - ex0) * (ey1 - ey0)
    return (int(ex0), int(ex1), int(ey0), int(ey1))


# ---------------------------------------------------------------------------
def structured_box(spec: BoxSpec, xi: np.ndarray, sub=None, rank=0, nranks=1) -> Mesh:
    """Build the (local) mesh of a structured box.

    ``sub=(ex0,ex1,ey0,ey1)`` restricts to a column block of elements (what one
    MPI rank / GPU owns under the xy partition); local node ids then follow the
    same closed-form numbering applied to the sub-box while ``ip2gip`` carries
    the global ids.  ``gip2owner`` is filled with ``rank`` and must be replaced
    by :func:`jexpresso_b200.sem.partition.find_gip_owner` for nranks > 1.
    """
    if spec.nsd == 3:
        return _box3d(spec, xi, sub, rank, nranks)
    return _box2d(spec, xi, sub, rank, nranks)


def _tag(spec, side):
    ax = "xyz"[{"xmin": 0, "xmax": 0, "ymin": 1, "ymax": 1, "zmin": 2, "zmax": 2}[side]]
    if spec.periodic[{"x": 0, "y": 1, "z": 2}[ax]]:
        return "periodic" + ax
    return spec.tags.get(side, "free_slip")


def _box3d(spec, xi, sub, rank, nranks):
    p = spec.nop
    n = p + 1
    NX, NY, NZ = spec.nel
    if sub is None:
        sub = (0, NX, 0, NY)
    ex0, ex1, ey0, ey1 = sub
    nx, ny, nz = ex1 - ex0, ey1 - ey0, NZ
    ggid, gtotal, _ = _numbering_3d(NX, NY, NZ, p)
    lgid, ltotal, lnv = _numbering_3d(nx, ny, nz, p)

    # local GLL grid -> local id / global id / coordinates
    Il, Jl, Kl = np.meshgrid(np.arange(nx * p + 1), np.arange(ny * p + 1), np.arange(nz * p + 1), indexing="ij")
    lid = lgid(Il, Jl, Kl)
    gidg = ggid(Il + ex0 * p, Jl + ey0 * p, Kl)
    ip2gip = np.empty(ltotal, np.int64)
    ip2gip[lid.reshape(-1)] = gidg.reshape(-1) + 1

    xa = _axis_coords(NX, p, spec.lo[0], spec.hi[0], xi)[ex0 * p: ex1 * p + 1]
    ya = _axis_coords(NY, p, spec.lo[1], spec.hi[1], xi)[ey0 * p: ey1 * p + 1]
    za = _axis_coords(NZ, p, spec.lo[2], spec.hi[2], xi)
    X = np.broadcast_to(xa[:, None, None], lid.shape).copy()
    Y = np.broadcast_to(ya[None, :, None], lid.shape).copy()
    Z = np.broadcast_to(za[None, None, :], lid.shape).copy()
    if spec.warp != 0.0:
        Lx, Ly, Lz = (spec.hi[d] - spec.lo[d] for d in range(3))
        sx, sy, sz = (X - spec.lo[0]) / Lx, (Y - spec.lo[1]) / Ly, (Z - spec.lo[2]) / Lz
        B = 64.0 * _bump(sx) * _bump(sy) * _bump(sz)
        hx, hy, hz = Lx / NX, Ly / NY, Lz / NZ
        wx = spec.warp * hx * B * np.sin(2 * np.pi * sy + 0.3) * np.cos(2 * np.pi * sz)
        wy = spec.warp * hy * B * np.sin(2 * np.pi * sz + 0.7) * np.cos(2 * np.pi * sx)
        wz = spec.warp * hz * B * np.sin(2 * np.pi * sx + 1.1) * np.cos(2 * np.pi * sy)
        X, Y, Z = X + wx, Y + wy, Z + wz
    x = np.empty(ltotal)
    y = np.empty(ltotal)
    z = np.empty(ltotal)
    x[lid.reshape(-1)] = X.reshape(-1)
    y[lid.reshape(-1)] = Y.reshape(-1)
    z[lid.reshape(-1)] = Z.reshape(-1)

    # connectivity: element order x fastest; local (i,j,k) -> (-x, +z, +y)
    nelem = nx * ny * nz
    E = np.arange(nelem)
    ex, ey, ez = E % nx, (E // nx) % ny, E // (nx * ny)
    i = np.arange(n)
    Ii = ex[:, None, None, None] * p + (p - i)[None, :, None, None]
    Kk = ez[:, None, None, None] * p + i[None, None, :, None]
    Jj = ey[:, None, None, None] * p + i[None, None, None, :]
    connijk = np.asfortranarray(lid[Ii, Jj, Kk] + 1)
    el2gel = (ex + ex0) + NX * ((ey + ey0) + NY * ez) + 1

    # boundary faces on physical sides of the global box present in this sub-box
    faces, ftype, fel = [], [], []
    a = np.arange(n)

    def add_side(side, fixed_axis, fixed_idx, el_sel, t1_of, t2_of):
        tag = _tag(spec, side)
        for e in el_sel:
            exl, eyl, ezl = e % nx, (e // nx) % ny, e // (nx * ny)
            base = {0: exl * p, 1: eyl * p, 2: ezl * p}
            idx = [None, None, None]
            idx[fixed_axis] = np.full((n, n), fixed_idx)
            idx[t1_of] = (base[t1_of] + a)[:, None] + np.zeros((n, n), np.int64)
            idx[t2_of] = (base[t2_of] + a)[None, :] + np.zeros((n, n), np.int64)
            if side in spec.wall_aligned:
                assert side == "ymin", "wall_aligned: local k runs along +y (mesh.jl:2194-2201), so only the ymin side qualifies"
                faces.append(np.array(connijk[e, :, :, 0]))
            else:
                faces.append(lid[idx[0], idx[1], idx[2]] + 1)
            ftype.append(tag)
            fel.append(e + 1)

    el = np.arange(nelem)
    exl, eyl, ezl = el % nx, (el // nx) % ny, el // (nx * ny)
    add_side("zmin", 2, 0, el[ezl == 0], 0, 1)
    add_side("zmax", 2, nz * p, el[ezl == nz - 1], 0, 1)
    if ey0 == 0:
        add_side("ymin", 1, 0, el[eyl == 0], 0, 2)
    if ey1 == NY:
        add_side("ymax", 1, ny * p, el[eyl == ny - 1], 0, 2)
    if ex0 == 0:
        add_side("xmin", 0, 0, el[exl == 0], 1, 2)
    if ex1 == NX:
        add_side("xmax", 0, nx * p, el[exl == nx - 1], 1, 2)
    pibf = np.asfortranarray(np.stack(faces, axis=0)) if faces else np.zeros((0, n, n), np.int64, order="F")

    coords = np.asfortranarray(np.stack([x, y, z], axis=0))
    m = Mesh(nsd=3, nop=p, ngl=n, nelem=nelem, npoin=ltotal, npoin_linear=lnv, gnpoin=gtotal,
             gnelem=NX * NY * NZ, connijk=connijk, x=x, y=y, z=z, coords=coords,
             poin_in_bdy_face=pibf, bdy_face_type=ftype, bdy_face_in_elem=np.array(fel, np.int64),
             ip2gip=ip2gip, gip2owner=np.full(ltotal, rank, np.int64), el2gel=el2gel,
             xmin=float(spec.lo[0]), xmax=float(spec.hi[0]), ymin=float(spec.lo[1]), ymax=float(spec.hi[1]),
             zmin=float(spec.lo[2]), zmax=float(spec.hi[2]), rank=rank, nranks=nranks, spec=spec, sub=sub)
    return m


def _box2d(spec, xi, sub, rank, nranks):
    p = spec.nop
    n = p + 1
    NX, NY = spec.nel[:2]
    if sub is None:
        sub = (0, NX, 0, NY)
    ex0, ex1, ey0, ey1 = sub
    nx, ny = ex1 - ex0, ey1 - ey0
    ggid, gtotal, _ = _numbering_2d(NX, NY, p)
    lgid, ltotal, lnv = _numbering_2d(nx, ny, p)
    Il, Jl = np.meshgrid(np.arange(nx * p + 1), np.arange(ny * p + 1), indexing="ij")
    lid = lgid(Il, Jl)
    gidg = ggid(Il + ex0 * p, Jl + ey0 * p)
    ip2gip = np.empty(ltotal, np.int64)
    ip2gip[lid.reshape(-1)] = gidg.reshape(-1) + 1
    xa = _axis_coords(NX, p, spec.lo[0], spec.hi[0], xi)[ex0 * p: ex1 * p + 1]
    ya = _axis_coords(NY, p, spec.lo[1], spec.hi[1], xi)[ey0 * p: ey1 * p + 1]
    X = np.broadcast_to(xa[:, None], lid.shape).copy()
    Y = np.broadcast_to(ya[None, :], lid.shape).copy()
    if spec.warp != 0.0:
        Lx, Ly = spec.hi[0] - spec.lo[0], spec.hi[1] - spec.lo[1]
        sx, sy = (X - spec.lo[0]) / Lx, (Y - spec.lo[1]) / Ly
        B = 16.0 * _bump(sx) * _bump(sy)
        X = X + spec.warp * (Lx / NX) * B * np.sin(2 * np.pi * sy + 0.3)
        Y = Y + spec.warp * (Ly / NY) * B * np.sin(2 * np.pi * sx + 1.1)
    x = np.empty(ltotal)
    y = np.empty(ltotal)
    x[lid.reshape(-1)] = X.reshape(-1)
    y[lid.reshape(-1)] = Y.reshape(-1)
    z = np.zeros(ltotal)

    nelem = nx * ny
    E = np.arange(nelem)
    ex, ey = E % nx, E // nx
    i = np.arange(n)
    # local i along +y, local j along -x (mesh.jl:2128-2131 with a lexicographic quad)
    Jj = ey[:, None, None] * p + i[None, :, None]
    Ii = ex[:, None, None] * p + (p - i)[None, None, :]
    connijk = np.asfortranarray((lid[Ii, Jj] + 1).reshape(nelem, n, n, 1))
    el2gel = (ex + ex0) + NX * (ey + ey0) + 1

    edges, etype, eel = [], [], []
    a = np.arange(n)
    for e in E[ey == 0] if ey0 == 0 else []:
        edges.append(lid[(e % nx) * p + a, 0] + 1); etype.append(_tag(spec, "ymin")); eel.append(e + 1)
    for e in E[ey == ny - 1] if ey1 == NY else []:
        edges.append(lid[(e % nx) * p + a, ny * p] + 1); etype.append(_tag(spec, "ymax")); eel.append(e + 1)
    for e in E[ex == 0] if ex0 == 0 else []:
        edges.append(lid[0, (e // nx) * p + a] + 1); etype.append(_tag(spec, "xmin")); eel.append(e + 1)
    for e in E[ex == nx - 1] if ex1 == NX else []:
        edges.append(lid[nx * p, (e // nx) * p + a] + 1); etype.append(_tag(spec, "xmax")); eel.append(e + 1)
    pibe = np.asfortranarray(np.stack(edges, axis=0)) if edges else np.zeros((0, n), np.int64, order="F")
    coords = np.asfortranarray(np.stack([x, y], axis=0))
    return Mesh(nsd=2, nop=p, ngl=n, nelem=nelem, npoin=ltotal, npoin_linear=lnv, gnpoin=gtotal,
                gnelem=NX * NY, connijk=connijk, x=x, y=y, z=z, coords=coords,
                poin_in_bdy_face=pibe, bdy_face_type=etype, bdy_face_in_elem=np.array(eel, np.int64),
                ip2gip=ip2gip, gip2owner=np.full(ltotal, rank, np.int64), el2gel=el2gel,
                xmin=float(spec.lo[0]), xmax=float(spec.hi[0]), ymin=float(spec.lo[1]), ymax=float(spec.hi[1]),
                zmin=0.0, zmax=0.0, rank=rank, nranks=nranks, spec=spec, sub=sub)


def element_sizes(mesh):
    """compute_element_size! (mesh.jl:5646-5711): per element, the extent of the bounding box of its corner nodes, shortest
    direction ("as if it were linear")."""
    n = mesh.ngl
    c = np.asarray(mesh.connijk) - 1
    ends = [0, n - 1]
    if mesh.nsd == 3:
        corners = np.stack([c[:, i, j, k] for i in ends for j in ends for k in ends], axis=1)
        axes = (mesh.x, mesh.y, mesh.z)
    else:
        c = c.reshape(mesh.nelem, n, n)
        corners = np.stack([c[:, i, j] for i in ends for j in ends], axis=1)
        axes = (mesh.x, mesh.y)
    ext = [a[corners].max(axis=1) - a[corners].min(axis=1) for a in axes]
    return np.minimum.reduce(ext)


def effective_delta_l(meshes):
    """mesh.Δeffective_l = max over ALL ranks of Δelem, divided by nop (compute_element_size_driver, mesh.jl:5621-5632: the
    MPI.Allreduce(MAX) is the max over the list here)."""
    meshes = list(meshes) if isinstance(meshes, (list, tuple)) else [meshes]
    return float(max(element_sizes(m).max() for m in meshes) / meshes[0].nop)
