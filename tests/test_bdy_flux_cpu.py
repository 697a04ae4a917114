"""Boundary fluxes with the Monin-Obukhov wall model (second part) in the oracle.  The reference holds no golden
vector of a bdy_fluxes deck, so the C restatement (oracle/jexref.c: apply_boundary_conditions_neumann_3d, CM_MOST, ...) is
cross-checked against an independent numpy transcription written from the Julia sources:

  BCs.jl:655-816              build_custom_bcs_neumann!(::NSD_3D): inside point, tangential velocity, wall distance, F_surf
  CM_MOST.jl:69-100, 144-152, 224-261   Businger-Dyer psi functions, Obukhov length, fixed-point iteration, CM_MOST!
  surface_integral.jl:1-29    compute_surface_integral!, DSS_surface_integral!;  rhs.jl:674-689: RHS .+= S_flux

The wall term is isolated as  du(bdy_fluxes) - du(without)  =  Minv * S_flux.  CPU only."""
import math

import numpy as np
import pytest

from helpers import PHYS, most_case
from oracle import ref

KARMAN, Z0M, Z0H = 0.4, 0.1, 0.01


def _psi_m(z):
    if z < 0:
        x = (1 - 16.0 * z) ** 0.25
        return 2 * math.log((1 + x) / 2) + math.log((1 + x * x) / 2) - 2 * math.atan(x) + math.pi / 2
    return -5.0 * z


def _psi_h(z):
    if z < 0:
        y = (1 - 16.0 * z) ** 0.5
        return 2 * math.log((1 + y) / 2)
    return -5.0 * z


def _L(us, T, QH, cp, g):
    return 1e6 if abs(QH) < 1e-6 else -us ** 3 * T * cp / (KARMAN * g * QH)


def _scales(u, th, z, ths, rho, cp, g):
    us = KARMAN * u / math.log(z / Z0M)
    ts = KARMAN * (th - ths) / math.log(z / Z0H)
    L = _L(us, th, -rho * cp * us * ts, cp, g)
    iters = 0
    for _ in range(20):
        iters += 1
        us2 = KARMAN * u / (math.log(z / Z0M) - _psi_m(z / L) + _psi_m(Z0M / L))
        ts2 = KARMAN * (th - ths) / (math.log(z / Z0H) - _psi_h(z / L) + _psi_h(Z0H / L))
        L2 = _L(us2, th, -rho * cp * us2 * ts2, cp, g)
        err = abs(L2 - L) / max(abs(L), abs(L2))
        us, ts, L = us2, ts2, L2
        if err < 1e-4:
            break
    return us, ts, L, iters


def numpy_wall_term(sem, u_bc, qe, lpert, Jef, ifirst, dhf, hf):
    """Minv * S_flux [neqs, npoin] plus diagnostics (Obukhov lengths, iteration counts)."""
    m = sem.mesh
    n, N = m.ngl, m.npoin
    q = u_bc.reshape(5, N)
    conn = np.asarray(m.connijk) - 1
    P = np.asarray(m.poin_in_bdy_face) - 1
    om = np.asarray(sem.basis["omega"])
    cp, g = PHYS[4], PHYS[2]
    S = np.zeros((5, N))
    Ls, its = [], []
    for f in range(P.shape[0]):
        if m.bdy_face_type[f] != "MOST":
            continue
        e = m.bdy_face_in_elem[f] - 1
        for i in range(n):
            for j in range(n):
                ip, ip1, isf = P[f, i, j], conn[e, i, j, ifirst - 1], conn[e, i, j, 0]
                if lpert:
                    rho = q[0, ip1] + qe[ip1, 0]
                    vel = np.array([(q[1 + d, ip1] + qe[ip1, 1 + d]) / rho for d in range(3)])
                    th = (q[4, ip1] + qe[ip1, 4]) / rho
                    ths = (q[4, isf] + qe[isf, 4]) / (q[0, isf] + qe[isf, 0])
                else:
                    rho = q[0, ip1]
                    vel = q[1:4, ip1] / rho
                    th = q[4, ip1] / rho
                    ths = q[4, isf] / q[0, isf]
                nrm = np.array([sem.nx[f, i, j], sem.ny[f, i, j], sem.nz[f, i, j]])
                vel = vel - (vel @ nrm) * nrm
                z = abs((m.coords[:, ip1] - m.coords[:, isf]) @ nrm)
                umag = math.sqrt(vel @ vel)
                us, ts, L, it = _scales(umag, th, z, ths, rho, cp, g)
                Ls.append(L)
                its.append(it)
                tau = -rho * us * us * (vel / (umag + 2.22e-16))
                wth = -us * ts
                F = np.array([0.0, tau[0], tau[1], tau[2], wth * (1.0 - dhf) + hf * dhf])
                S[:, ip] += om[i] * om[j] * Jef[f, i, j] * F
    return (S * np.asarray(sem.Minv)[None, :]).reshape(-1), np.array(Ls), np.array(its)


@pytest.mark.parametrize("lpert", [False, True])
def test_oracle_most_wall_fluxes_match_numpy_transcription(oracle_lib, lpert):
    sem, qe, u0, Jef = most_case(lpert)
    m = sem.mesh
    N = m.npoin
    assert "MOST" in m.bdy_face_type
    f = m.bdy_face_type.index("MOST")
    e = m.bdy_face_in_elem[f] - 1
    assert np.array_equal(m.poin_in_bdy_face[f], m.connijk[e, :, :, 0])      # wall faces in the element's (i, j) order
    bf = dict(Jef=Jef, ifirst_wall_node_index=3, delta_hf=0.25, user_heatflux=0.12)
    mk = lambda b: ref.RefProblem(sem, qe, eq_id=0, lpert=lpert, lsource=True, lvisc=False, phys=PHYS, pow_mode=1, neqs=5,
                                  bdy_fluxes=b)
    outs = []
    for b in (None, bf):
        prob = mk(b)
        u, RHS = u0.copy(), np.zeros(5 * N)
        prob.build_rhs_local(u, RHS, 0.0)
        prob.divide_by_mass(RHS)
        outs.append((u, RHS))
    assert np.array_equal(outs[0][0], outs[1][0])
    got = outs[1][1] - outs[0][1]
    want, Ls, its = numpy_wall_term(sem, outs[0][0], qe, lpert, Jef, 3, 0.25, 0.12)
    assert (Ls > 0).any() and (Ls < 0).any(), "both the stable and the unstable branch of psi must be exercised"
    assert its.max() > 1 and its.max() < 20
    for eq in range(5):
        sl = slice(eq * N, (eq + 1) * N)
        scale = np.max(np.abs(want[sl]))
        if eq in (0, 2):      # no mass flux; the wall is flat (the warp vanishes on the boundary): the tangential wind has no y part
            assert scale == 0.0 and np.max(np.abs(got[sl])) <= 4e-16 * np.max(np.abs(outs[0][1][sl]))
            continue
        assert scale > 0
        tol = 1e-11 * scale + 4e-16 * np.max(np.abs(outs[0][1][sl]))
        assert np.max(np.abs(got[sl] - want[sl])) <= tol, (eq, np.max(np.abs(got[sl] - want[sl])), scale)
    # only wall nodes receive a flux
    wall = np.zeros(N, bool)
    wall[(m.poin_in_bdy_face[[i for i, t in enumerate(m.bdy_face_type) if t == "MOST"]] - 1).ravel()] = True
    assert not got.reshape(5, N)[:, ~wall].any()
