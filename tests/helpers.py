"""Shared problem builders for the tests: the same seeded inputs go to the oracle and to the GPU."""
import numpy as np

from jexpresso_b200.physics import PhysicalConst
from jexpresso_b200.sem import BoxSpec, conformity4ncf_q_host, rtb_initial_state, sem_setup

MU3 = [0.0, 125.0, 125.0, 125.0, 125.0]
MU2 = [0.0, 125.0, 125.0, 125.0]


def box3d(nel=(4, 4, 4), nop=4, warp=0.0, periodic=(False, False, False), L=(10000.0, 10000.0, 10000.0)):
    return BoxSpec(nsd=3, nel=tuple(nel), nop=nop, lo=(0.0, 0.0, 0.0), hi=tuple(L), periodic=periodic, warp=warp)


def box2d(nel=(10, 10), nop=4, warp=0.0, periodic=(False, False, False), lo=(-5000.0, 0.0), hi=(5000.0, 10000.0)):
    return BoxSpec(nsd=2, nel=tuple(nel), nop=nop, lo=tuple(lo), hi=tuple(hi), periodic=periodic, warp=warp)


def euler_case(spec, nranks=1, lpert=False, seed=1234, vel_amp=1.0, condition=True):
    """Per-rank (sems, qns, qes, us) for the CompEuler theta rising-bubble state with a seeded
    momentum perturbation, IC conditioned like params_setup.jl:259-297."""
    sems = sem_setup(spec, nranks)
    neqs = spec.nsd + 2
    qns, qes = [], []
    for s in sems:
        qn, qe = rtb_initial_state(s.mesh, lpert, seed=seed, vel_amp=vel_amp)
        qns.append(qn)
        qes.append(qe)
    if condition:
        conformity4ncf_q_host(sems, qns, neqs)
        conformity4ncf_q_host(sems, qes, neqs)
    us = [np.ascontiguousarray(q[:, :neqs].reshape(-1, order="F")) for q in qns]
    return sems, qns, qes, us


def rel_err_per_node(a, b):
    """max |a-b| / max(|b|, eps*||b||_inf) and relative L2 -- the north-star parity measures."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-3 * np.abs(b).max() if b.size else 1.0)
    pn = float(np.max(np.abs(a - b) / scale)) if b.size else 0.0
    l2 = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    return pn, l2


PHYS = PhysicalConst().packed()


def kopriva_case():
    """problems/AdvDiff/kopriva (the reference's CI case, test/ci_cases.jl:75): 10x20 elements, nop 4, periodic in x and y
    (kopriva_periodic.msh), Gaussian at (5, 3) (initialize.jl:17-30), wind (0.5, 1) (user_flux.jl:1-16), AV mu = 0.1,
    PERT with qe = 0, IC conditioned by conformity4ncf_q! (params_setup.jl:259-297 -- this is what makes the periodic
    twins start from one value).  Returns (sem, qe, u0, phys, inputs)."""
    from jexpresso_b200.sem.setup import conformity4ncf_q_host
    spec = box2d((10, 20), 4, periodic=(True, True, False), lo=(0.0, 0.0), hi=(10.0, 20.0))
    sems = sem_setup(spec, 1)
    m = sems[0].mesh
    xc = (m.x.max() + m.x.min()) / 2
    qn = np.zeros((m.npoin, 2), order="F")
    a1, a2 = -((m.x - xc) / 1.0) ** 2, -((m.y - 3.0) / 1.0) ** 2
    qn[:, 0] = 1.0 * np.exp(a1) * np.exp(a2)
    qe = np.zeros((m.npoin, 2), order="F")
    conformity4ncf_q_host(sems, [qn], 1)
    from jexpresso_b200.physics import advdiff_packed
    phys = advdiff_packed(0.5, 1.0)
    inputs = {"SOL_VARS_TYPE": "PERT", "lsource": True, "lvisc": True, "mu": [0.1], "dt": 0.005, "ode_solver": "SSPRK54"}
    return sems[0], qe, np.ascontiguousarray(qn[:, 0]).copy(), phys, inputs


def soliwave_case(nel=(25, 30)):
    """problems/ShallowWater/SoliWaveIsland (the reference's CI case, test/ci_cases.jl:79): [0,25] x [-15,15], 25 x 30
    elements, nop 4, free-slip walls, solitary wave over a conical island (initialize.jl:47-100), TOTAL, AV mu = 0.05.
    Returns (sem, qn, qe, u0, phys, inputs); smaller ``nel`` gives the same set-up on a coarser mesh (functor tests)."""
    from jexpresso_b200.physics import swe_packed
    spec = box2d(nel, 4, lo=(0.0, -15.0), hi=(25.0, 15.0))
    sems = sem_setup(spec, 1)
    m = sems[0].mesh
    g, A, h0, xc_wave = 9.81, 0.064, 0.32, 2.5
    gam = np.sqrt(3.0 * A / (4.0 * h0 ** 3))
    xc, yc, rc, hc, H_dry = 12.5, 0.0, 3.6, 0.93, 1.0e-3
    sech = 1.0 / np.cosh(gam * (m.x - xc_wave))
    eta = A * sech * sech
    dx, dy = m.x - xc, m.y - yc
    r = np.sqrt(dx * dx + dy * dy)
    Hb = np.where(r < rc, hc * (1.0 - r / rc), 0.0)
    Hw = np.maximum(h0 + eta - Hb, H_dry)
    ux = np.where(Hw > H_dry, np.sqrt(g * (h0 + eta)) * eta / (h0 + eta), 0.0)
    qn = np.zeros((m.npoin, 4), order="F")
    qe = np.zeros((m.npoin, 4), order="F")
    qn[:, 0], qn[:, 1] = Hw, Hw * ux
    qe[:, 0] = np.maximum(h0 - Hb, H_dry)
    inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": True, "mu": [0.05, 0.05, 0.05], "dt": 0.01,
              "ode_solver": "SSPRK54"}
    u0 = np.ascontiguousarray(qn[:, :3].reshape(-1, order="F"))
    return sems[0], qn, qe, u0, swe_packed(), inputs
