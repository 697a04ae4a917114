"""Pins the CPU oracle (oracle/jexref.c + oracle/ref.py) against the reference's OWN golden end
state for this path: test/CI-ref/CompEuler/theta (2D theta-form Euler, PERT, AV mu=125, 2000
CarpenterKennedy2N54 steps of dt=0.5 on the 10x10 nop=4 box; reference tolerance atol=1e-5,
test/ci_cases.jl:57,73).  The HDF5 files carry no coordinates and follow Gridap's node numbering,
so the comparison is order free: the sorted nodal values of every variable must agree."""
import os

import numpy as np
import pytest

from helpers import MU2, PHYS, box2d, euler_case
from oracle import ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "CompEuler_theta.npz")


@pytest.fixture(scope="module")
def theta_end_state():
    spec = box2d((10, 10), 4)
    sems, qns, qes, us = euler_case(spec, 1, lpert=True, seed=None)
    prob = ref.RefProblem(sems[0], qes[0], eq_id=0, lpert=True, lsource=True, lvisc=True, visc_coeff=MU2, phys=PHYS, pow_mode=0)
    run = ref.RefRun([prob])
    t = ref.time_loop(run, us, 0.0, 0.5, 2000, scheme="CK2N54")
    return sems[0], qes[0], us[0], t


def test_theta_golden_end_state(theta_end_state):
    sem, qe, u, t = theta_end_state
    g = np.load(GOLD)
    assert abs(t - 1000.0) < 1e-9 and abs(float(g["t_time"][0]) - 1000.0) < 1e-6
    N = sem.mesh.npoin
    assert N == g["q1"].shape[0] == 1681
    worst = 0.0
    for i in range(4):
        mine = np.sort(u[i * N:(i + 1) * N])
        gold = np.sort(g[f"q{i + 1}"])
        worst = max(worst, float(np.max(np.abs(mine - gold))))
        assert np.allclose(mine, gold, rtol=0.0, atol=1e-5), f"variable {i + 1} outside the reference's CI tolerance"
        assert np.allclose(np.sort(qe[:, i]), np.sort(g[f"qe{i + 1}"]), rtol=0.0, atol=1e-9)
    # far inside the reference's own tolerance: the restatement tracks the real Julia run closely
    assert worst < 1e-8, worst


def test_theta_golden_moments(theta_end_state):
    """Order-free moments (sum, sum of squares, extrema) of the end state, per variable."""
    sem, qe, u, t = theta_end_state
    g = np.load(GOLD)
    N = sem.mesh.npoin
    for i in range(4):
        a, b = u[i * N:(i + 1) * N], g[f"q{i + 1}"]
        for f in (np.sum, np.min, np.max, lambda v: np.sum(v * v)):
            fa, fb = float(f(a)), float(f(b))
            assert abs(fa - fb) <= 1e-5 * max(1.0, abs(fb)) * 10


def test_advdiff_kopriva_golden_end_state():
    """test/CI-ref/AdvDiff/kopriva (2D advection-diffusion, doubly periodic, AV mu=0.1, SSPRK54, dt=0.005, tend=10): pins
    the AdvDiff functor, the neqs=1 viscous path, the periodic-twin assembly (incl. the 4-fold corner), the IC
    conditioning and -- per stage -- the SSPRK54 restatement against a real Julia/OrdinaryDiffEq run.  dt is
    Float64(Float32(0.005)) (TimeIntegrators.jl:464-465), so 2000 steps end 2.2e-7 short of tend and the integrator takes
    one clipped step to tend.  Order-free comparison (Gridap numbering); the reference's default CI tolerance applies
    (rtol 1e-5 class), the restatement lands at 1e-12."""
    from helpers import kopriva_case
    sem, qe, u0, phys, inputs = kopriva_case()
    m = sem.mesh
    prob = ref.RefProblem(sem, qe, eq_id=2, lpert=True, lsource=True, lvisc=True, visc_coeff=inputs["mu"], phys=phys,
                          pow_mode=0, neqs=1)
    run = ref.RefRun([prob], ref.setup_assembler([m.ip2gip], [m.gip2owner]))
    us = [u0.copy()]
    t = ref.time_loop(run, us, 0.0, inputs["dt"], 2000, scheme="SSPRK54")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "AdvDiff_kopriva.npz"))
    gold = np.sort(g["q1"])
    assert gold.shape[0] == m.npoin == 3321 and abs(float(g["t_time"][0]) - 10.0) < 1e-9
    assert np.max(np.abs(np.sort(us[0]) - gold)) < 1e-7            # 2.0e-8: the missing last 2.2e-7 of time
    ref.step_ssprk54(run, us, t, 10.0 - t, [np.zeros_like(us[0])])   # the clipped final step
    worst = float(np.max(np.abs(np.sort(us[0]) - gold)))
    assert worst < 1e-10, worst                                    # measured 1.1e-12


def test_shallow_water_soliwave_island_golden_end_state():
    """test/CI-ref/ShallowWater/SoliWaveIsland (2D non-linear shallow water with wet/dry front, AV mu=0.05, SSPRK54,
    dt=0.01 to tend=25, 12 221 nodes): pins the ShallowWater functor (flux, bathymetry source with dry-node relaxation,
    primitives, free-slip walls) within the reference's CI tolerance atol=1e-5 (test/ci_cases.jl:56).  The bulk agrees to
    ~1e-9; the few nodes at the moving wet/dry front carry the largest differences (H 3e-7, Hu 5e-6, Hv 2e-8): the
    `H < H_wet` switch of user_source.jl:60 reacts to the 1e-12 differences between gmsh's node coordinates and the
    closed-form ones used here."""
    from helpers import soliwave_case
    sem, qn, qe, u0, phys, inputs = soliwave_case()
    prob = ref.RefProblem(sem, qe, eq_id=3, lpert=False, lsource=True, lvisc=True, visc_coeff=inputs["mu"], phys=phys,
                          pow_mode=0, neqs=3)
    run = ref.RefRun([prob])
    us = [u0.copy()]
    t = ref.time_loop(run, us, 0.0, inputs["dt"], 2500, scheme="SSPRK54")
    ref.step_ssprk54(run, us, t, 25.0 - t, [np.zeros_like(us[0])])     # the integrator's clipped final step
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ShallowWater_SoliWaveIsland.npz"))
    N = sem.mesh.npoin
    assert N == g["q1"].shape[0] == 12221 and abs(float(g["t_time"][0]) - 25.0) < 1e-9
    for i in range(3):
        d = np.abs(np.sort(us[0][i * N:(i + 1) * N]) - np.sort(g[f"q{i + 1}"]))
        assert np.max(d) < 1e-5, (i, float(np.max(d)))                   # the reference's own CI tolerance
        assert np.median(d) < 1e-7 and np.count_nonzero(d > 1e-6) < 200, (i, float(np.median(d)))
        assert np.max(np.abs(np.sort(qe[:, i]) - np.sort(g[f"qe{i + 1}"]))) < 1e-9
