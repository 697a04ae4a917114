"""Metric terms, boundary normals and the diagonal mass matrix (host setup inputs).

Restated from
  * src/kernel/mesh/metric_terms.jl:332-474   build_metric_terms! (3D, volume)
  * src/kernel/mesh/metric_terms.jl:476-607   boundary-face Jacobian / normals (3D)
  * src/kernel/mesh/metric_terms.jl:197-257   build_metric_terms! (2D, volume)
  * src/kernel/infrastructure/element_matrices.jl:173-214, 593-617, 1557-1559
        build_mass_matrix! / DSS_mass! / mass_inverse!  (inexact => diagonal)
  * src/kernel/infrastructure/element_matrices.jl:1059-1115  DSS_global_normals!

With LGL interpolation = quadrature points, psi[i,l] is exactly the identity
(every factor is (x_l-x_j)/(x_i-x_j) = 1 or one factor is exactly 0), so the
reference's triple sums collapse to single sums with exact-zero addends; the
loops below keep the surviving terms in the reference's accumulation order
(ascending node index) with separate multiply and add.

All outputs use the reference's layout: Float64[nelem, ngl, ngl, ngl] Fortran
order (element index fastest, metric_terms.jl:77).
"""
from __future__ import annotations

import numpy as np

__all__ = ["build_metric_terms", "build_mass_local", "build_mass_inverse", "boundary_normals"]


def _elem_coords(mesh, arr):
    # arr[npoin] -> [E, n, n, n] gathered through connijk
    return arr[mesh.connijk - 1]


def build_metric_terms(mesh, basis):
    """Return dict with dξdx..dζdz, Je (3D) or dξdx,dξdy,dηdx,dηdy,Je (2D)."""
    dpsi = basis["dpsi"]
    n = mesh.ngl
    if mesh.nsd == 3:
        X, Y, Z = (_elem_coords(mesh, a) for a in (mesh.x, mesh.y, mesh.z))
        shp = X.shape

        def deriv(C, axis):
            # out[.., l, ..] = sum_i dpsi[i, l] * C[.., i, ..], ascending i, multiply then add
            # (slices are views; no gather copies)
            out = np.empty(shp)
            tmp = np.empty(shp[:axis] + shp[axis + 1:])
            for l in range(n):
                idx = [slice(None)] * 4
                idx[axis] = l
                acc = out[tuple(idx)]
                acc[...] = 0.0
                for i in range(n):
                    idx[axis] = i
                    np.multiply(C[tuple(idx)], dpsi[i, l], out=tmp)
                    np.add(acc, tmp, out=acc)
            return out

        dxdxi, dxdeta, dxdzeta = deriv(X, 1), deriv(X, 2), deriv(X, 3)
        dydxi, dydeta, dydzeta = deriv(Y, 1), deriv(Y, 2), deriv(Y, 3)
        dzdxi, dzdeta, dzdzeta = deriv(Z, 1), deriv(Z, 2), deriv(Z, 3)
        c1 = dydeta * dzdzeta - dydzeta * dzdeta
        c2 = dxdzeta * dzdeta - dxdeta * dzdzeta
        c3 = dxdeta * dydzeta - dxdzeta * dydeta
        c4 = dydzeta * dzdxi - dydxi * dzdzeta
        c5 = dxdxi * dzdzeta - dxdzeta * dzdxi
        c6 = dxdzeta * dydxi - dxdxi * dydzeta
        c7 = dydxi * dzdeta - dydeta * dzdxi
        c8 = dxdeta * dzdxi - dxdxi * dzdeta
        c9 = dxdxi * dydeta - dxdeta * dydxi
        Je = dxdxi * c1 + dydxi * c2 + dzdxi * c3
        Jinv = 1.0 / Je
        out = {"Je": Je,
               "dξdx": c1 * Jinv, "dξdy": c2 * Jinv, "dξdz": c3 * Jinv,
               "dηdx": c4 * Jinv, "dηdy": c5 * Jinv, "dηdz": c6 * Jinv,
               "dζdx": c7 * Jinv, "dζdy": c8 * Jinv, "dζdz": c9 * Jinv}
        return {k: np.asfortranarray(v) for k, v in out.items()}
    # 2D
    conn = mesh.connijk[:, :, :, 0]
    X, Y = mesh.x[conn - 1], mesh.y[conn - 1]
    shp = X.shape

    def deriv2(C, axis):
        out = np.empty(shp)
        tmp = np.empty(shp[:axis] + shp[axis + 1:])
        for l in range(n):
            idx = [slice(None)] * 3
            idx[axis] = l
            acc = out[tuple(idx)]
            acc[...] = 0.0
            for i in range(n):
                idx[axis] = i
                np.multiply(C[tuple(idx)], dpsi[i, l], out=tmp)
                np.add(acc, tmp, out=acc)
        return out

    dxdxi, dxdeta = deriv2(X, 1), deriv2(X, 2)
    dydxi, dydeta = deriv2(Y, 1), deriv2(Y, 2)
    Je = dxdxi * dydeta - dydxi * dxdeta
    Jinv = 1.0 / Je
    out = {"Je": Je, "dξdx": dydeta * Jinv, "dξdy": -dxdeta * Jinv,
           "dηdx": -dydxi * Jinv, "dηdy": dxdxi * Jinv}
    return {k: np.asfortranarray(v.reshape(shp + (1,))) for k, v in out.items()}


def build_mass_local(mesh, basis, Je):
    """Local (un-assembled across ranks) diagonal mass  M[ip] = sum_e (w_i w_j) w_k Je,
    elements in ascending order (DSS_mass!, element_matrices.jl:593-617)."""
    om = basis["omega"]
    if mesh.nsd == 3:
        w = (om[:, None, None] * om[None, :, None]) * om[None, None, :]     # (w_m*w_n)*w_o
        wJ = w[None] * Je
        conn = mesh.connijk
    else:
        w = om[:, None] * om[None, :]
        wJ = w[None] * Je[:, :, :, 0]
        conn = mesh.connijk[:, :, :, 0]
    # bincount accumulates sequentially in input order; element axis first gives the
    # reference's element-ascending order of additions for every node.
    return np.bincount((conn - 1).reshape(mesh.nelem, -1).reshape(-1),
                       weights=np.ascontiguousarray(wJ.reshape(mesh.nelem, -1)).reshape(-1),
                       minlength=mesh.npoin)


def build_mass_inverse(M):
    return 1.0 / M


def boundary_normals(mesh):
    """Outward unit normals at boundary nodes.

    3D: nx,ny,nz Float64[nfaces_bdy, ngl, ngl] (metric_terms.jl:535-600): cross product
    of the differences to the two neighbouring face nodes, flipped to point away from the
    element-interior node connijk[e,2,2,2].
    2D: nx,ny Float64[nedges_bdy, ngl] (metric_terms.jl:262-310): rotated difference to
    the next edge node, flipped away from connijk[e,2,2].
    """
    n = mesh.ngl
    P = mesh.poin_in_bdy_face - 1
    nf = P.shape[0]
    idx = np.arange(n)
    nb = np.where(idx < n - 1, idx + 1, idx - 1)
    x, y, z = mesh.x, mesh.y, mesh.z
    e = mesh.bdy_face_in_elem - 1
    if mesh.nsd == 3:
        nx = np.zeros((nf, n, n), order="F")
        ny = np.zeros((nf, n, n), order="F")
        nz = np.zeros((nf, n, n), order="F")
        if nf == 0:
            return nx, ny, nz
        ip, ip1, ip2 = P, P[:, nb, :], P[:, :, nb]
        dx1, dy1, dz1 = x[ip] - x[ip1], y[ip] - y[ip1], z[ip] - z[ip1]
        dx2, dy2, dz2 = x[ip] - x[ip2], y[ip] - y[ip2], z[ip] - z[ip2]
        cx = dy1 * dz2 - dz1 * dy2
        cy = dz1 * dx2 - dx1 * dz2
        cz = dx1 * dy2 - dy1 * dx2
        ninv = 1.0 / np.sqrt(cx * cx + cy * cy + cz * cz)
        nx[...] = cx * ninv
        ny[...] = cy * ninv
        nz[...] = cz * ninv
        ip3 = mesh.connijk[e, 1, 1, 1] - 1
        dot = (nx * (x[ip3][:, None, None] - x[ip]) + ny * (y[ip3][:, None, None] - y[ip])
               + nz * (z[ip3][:, None, None] - z[ip]))
        flip = dot > 0
        for a in (nx, ny, nz):
            a[flip] = -a[flip]
        return nx, ny, nz
    nx = np.zeros((nf, n), order="F")
    ny = np.zeros((nf, n), order="F")
    if nf == 0:
        return nx, ny, None
    ip, ip1 = P, P[:, nb]
    dx, dy = x[ip] - x[ip1], y[ip] - y[ip1]
    mag_inv = 1.0 / np.sqrt(dx * dx + dy * dy)
    nx[...] = dy * mag_inv
    ny[...] = -dx * mag_inv
    ip2 = mesh.connijk[e, 1, 1, 0] - 1
    dot = nx * (x[ip2][:, None] - x[ip]) + ny * (y[ip2][:, None] - y[ip])
    flip = dot > 0
    nx[flip] = -nx[flip]
    ny[flip] = -ny[flip]
    return nx, ny, None


def boundary_face_jacobian(mesh, basis):
    """metrics.Jef Float64[nfaces_bdy, ngl, ngl] (metric_terms.jl:479-570): the surface Jacobian of every boundary-face node,
    |x_xi x x_eta| with the face coordinates differentiated along the two face directions (psi is the identity at the LGL
    nodes).  An input array of the boundary-flux path (params.metrics.Jef); host-side test support like the rest of sem/."""
    assert mesh.nsd == 3
    P = mesh.poin_in_bdy_face - 1
    nf, n = P.shape[0], mesh.ngl
    Jef = np.zeros((nf, n, n), order="F")
    if nf == 0:
        return Jef
    dpsi = np.asarray(basis["dpsi"])                       # dpsi[i, k] = L'_i(xi_k)
    d = {}
    for name, c in (("x", mesh.x), ("y", mesh.y), ("z", mesh.z)):
        f = c[P]                                           # [face, i, j]
        d[name + "xi"] = np.einsum("ik,fil->fkl", dpsi, f)
        d[name + "eta"] = np.einsum("jl,fkj->fkl", dpsi, f)
    a = d["yeta"] * d["zxi"] - d["yxi"] * d["zeta"]
    b = d["xxi"] * d["zeta"] - d["xeta"] * d["zxi"]
    c_ = d["xeta"] * d["yxi"] - d["xxi"] * d["yeta"]
    Jef[...] = np.sqrt(a * a + b * b + c_ * c_)
    return Jef
