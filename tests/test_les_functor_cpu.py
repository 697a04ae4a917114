"""The LESICP1 source functor (problems/CompEuler/LESICP1/user_source.jl:1-103: gravity + top sponge towards qe + Coriolis +
geostrophic wind) in the oracle, checked independently: with a diagonal mass matrix the assembled, mass-scaled contribution of
a nodal source is the source itself, so  du(LES functor) - du(plain theta functor)  must equal  S_LES - S_theta  evaluated
with numpy at every node (to rounding).  The reference holds no golden data for this case: PARITY UNPINNED for the functor."""
import numpy as np

from helpers import PHYS, box3d, euler_case
from jexpresso_b200.physics import EQ_EULER_THETA_LES, les_packed
from oracle import ref


def _du(sems, qes, us, eq_id, phys):
    prob = ref.RefProblem(sems[0], qes[0], eq_id=eq_id, lpert=False, lsource=True, lvisc=False, phys=phys, pow_mode=1, neqs=5)
    run = ref.RefRun([prob])
    u, du = [us[0].copy()], [np.zeros_like(us[0])]
    run.rhs(du, u, 0.0)
    return u[0], du[0]


def test_les_source_is_the_nodal_difference_to_the_theta_functor():
    spec = box3d((4, 3, 5), 4, warp=0.05)
    sems, qns, qes, us = euler_case(spec, 1, lpert=False)
    m = sems[0].mesh
    N = m.npoin
    qes[0][:, 1] = 10.0 * qes[0][:, 0]                       # a reference state with a geostrophic wind (U, V) = (10, 2)
    qes[0][:, 2] = 2.0 * qes[0][:, 0]
    zmax, zs, f, alpha = float(m.z.max()), 6000.0, 1.0e-4, 0.5
    u_bc, du0 = _du(sems, qes, us, 0, PHYS)
    for lsponge in (False, True):
        ph = les_packed(zmax, lsponge=lsponge, zsponge=zs, f=f, alpha=alpha)
        assert ph[:8] == list(PHYS)[:8]
        u2, du1 = _du(sems, qes, us, EQ_EULER_THETA_LES, ph)
        assert np.array_equal(u2, u_bc)                       # same free-slip projection
        q = u_bc.reshape(5, N)                                # projected state the sources saw
        qe = qes[0]
        cs = np.where(m.z >= zs, alpha * np.sin(np.pi * (0.5 * (m.z - zs) / (zmax - zs))), 0.0) if lsponge else np.zeros(N)
        dS = np.zeros((5, N))
        dS[1] = -cs * (q[1] - qe[:, 1]) + f * q[2] - q[0] * f * (qe[:, 2] / qe[:, 0])
        dS[2] = -cs * (q[2] - qe[:, 2]) - f * q[1] + q[0] * f * (qe[:, 1] / qe[:, 0])
        dS[3] = -cs * (q[3] - qe[:, 3])
        diff = (du1 - du0).reshape(5, N)
        scale = np.abs(du0).max()
        assert np.abs(diff - dS).max() <= 1e-11 * scale, (lsponge, np.abs(diff - dS).max(), scale)
        assert np.abs(dS).max() > 0 and (not lsponge or np.abs(dS[3]).max() > 0)
        assert not diff[0].any() and not diff[4].any()        # mass and theta equations: no LES source
