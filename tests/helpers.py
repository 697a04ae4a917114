"""Shared problem builders for the tests: the same seeded inputs go to the oracle and to the GPU.  The builders the bench
and smoke() use as well live in the package (jexpresso_b200/sem/problems.py); the reference's CI cases stay here."""
import numpy as np

from jexpresso_b200.sem import BoxSpec, conformity4ncf_q_host, rtb_initial_state, sem_setup  # noqa: F401
from jexpresso_b200.sem.problems import MU2, MU3, PHYS, box2d, box3d, euler_case, rel_err_per_node  # noqa: F401


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


def most_spec(nel=(3, 2, 3), nop=4, warp=0.05):
    """A box whose ymin side is the MOST wall, listed in the element's own (i, j) order (BoxSpec.wall_aligned)."""
    return BoxSpec(nsd=3, nel=tuple(nel), nop=nop, lo=(0.0, 0.0, 0.0), hi=(3000.0, 100.0, 3000.0), periodic=(True, False, False),
                   tags={"ymin": "MOST"}, warp=warp, wall_aligned=("ymin",))


def most_case(lpert, seed=1234):
    spec = most_spec()
    sems, qns, qes, us = euler_case(spec, 1, lpert=lpert, seed=seed, vel_amp=6.0)
    sem = sems[0]
    m = sem.mesh
    N = m.npoin
    u0 = us[0]
    # a stably and an unstably stratified half of the wall: warm the air over x < L/2, cool it over x > L/2 (theta only)
    q = u0.reshape(5, N)
    dth = np.where(m.x < 1500.0, 1.5, -0.05) * np.exp(-m.y / 40.0)
    rho = q[0] + (qes[0][:, 0] if lpert else 0.0)
    q[4] += rho * dth
    q[1] += rho * 8.0          # a mean wind along the wall: the fixed-point iteration for (u*, theta*) converges at every node
    from jexpresso_b200.sem.metrics import boundary_face_jacobian
    Jef = boundary_face_jacobian(m, sem.basis)
    return sem, qes[0], u0, Jef


