"""One rhs! of the registered device functors other than CompEuler theta -- ShallowWater (SoliWaveIsland), total-energy
Euler (kelvinHelmholtzChan2022) and AdvDiff (kopriva 2D, 3d_periodic) -- against the CPU oracle, whose AdvDiff and
ShallowWater functors are pinned on the reference's own CI end states (tests/test_oracle_golden.py).  Bars: the north
star's <= 1e-12 per node and <= 1e-10 relative L2 (deterministic DSS).

The last test runs the AdvDiff/kopriva CI case (2000 SSPRK54 steps + the clipped final step) through jx_step and holds the
CUDA path itself to the reference's golden end state.  The file sorts last so that nothing here can disturb the hard
assertions of the other GPU files.

The total-energy functor differentiates (rho, u, v, T) and carries the viscous-work term tau.u on its energy equation
(kelvinHelmholtzChan2022/user_primitives.jl:17-23, rhs.jl:1988, 2018-2041); its oracle twin is cross-checked against an
independent numpy transcription in tests/test_energy_functor_cpu.py.  Every test here is a hard assertion."""
import os

import numpy as np
import pytest

from helpers import MU2, PHYS, box2d, box3d, kopriva_case, rel_err_per_node, soliwave_case
from jexpresso_b200 import rhs as jrhs
from jexpresso_b200.physics import advdiff_packed
from jexpresso_b200.sem import sem_setup
from oracle import ref

pytestmark = pytest.mark.gpu


def _compare(sem, qe, u0, neqs, eq_id, eqs, phys, inputs, caches=None):
    lpert = inputs["SOL_VARS_TYPE"] == "PERT"
    prob = ref.RefProblem(sem, qe, eq_id=eq_id, lpert=lpert, lsource=inputs["lsource"], lvisc=inputs["lvisc"],
                          visc_coeff=np.broadcast_to(np.asarray(inputs["mu"], float), (neqs,)).copy(), phys=phys, pow_mode=1,
                          neqs=neqs)
    run = ref.RefRun([prob], caches)
    uo, duo = [u0.copy()], [np.zeros_like(u0)]
    run.rhs(duo, uo, 0.0)
    p = jrhs.params_setup(sem, qe, inputs, eqs=eqs, phys=phys, pow_mode=1, dss_mode=0)
    try:
        u, du = u0.copy(), np.empty_like(u0)
        jrhs.rhs_bang(du, u, p, 0.0)
    finally:
        p.close()
    assert np.array_equal(u, uo[0]), "boundary-projected state differs from the oracle"
    N = sem.mesh.npoin
    for e in range(neqs):
        sl = slice(e * N, (e + 1) * N)
        if np.max(np.abs(duo[0][sl])) == 0.0:
            assert np.max(np.abs(du[sl])) == 0.0
            continue
        pn, l2 = rel_err_per_node(du[sl], duo[0][sl])
        assert pn <= 1e-12 and l2 <= 1e-10, (eqs, e, pn, l2)


def test_shallow_water_functor_one_rhs():
    sem, qn, qe, u0, phys, inputs = soliwave_case(nel=(10, 12))
    rng = np.random.default_rng(7)
    N = sem.mesh.npoin
    u0 = u0.copy()
    u0[2 * N:] = 0.01 * rng.uniform(-1.0, 1.0, N) * u0[:N]          # some cross-flow so that every flux entry is exercised
    _compare(sem, qe, u0, 3, 3, "ShallowWater", phys, inputs)


def test_euler_energy_functor_one_rhs():
    spec = box2d((6, 5), 4, warp=0.05)
    sem = sem_setup(spec, 1)[0]
    N = sem.mesh.npoin
    rng = np.random.default_rng(11)
    rho = 1.0 + 0.2 * rng.uniform(-1.0, 1.0, N)
    uv = 0.3 * rng.uniform(-1.0, 1.0, (2, N))
    pres = 1.0 + 0.1 * rng.uniform(-1.0, 1.0, N)
    gamma = PHYS[1]
    rE = pres / (gamma - 1.0) + 0.5 * rho * (uv[0] ** 2 + uv[1] ** 2)
    u0 = np.concatenate([rho, rho * uv[0], rho * uv[1], rE])
    qe = np.zeros((N, 5), order="F")
    inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": False, "lvisc": True, "mu": MU2, "dt": 0.1, "ode_solver": "SSPRK54"}
    _compare(sem, qe, u0, 4, 1, "CompEulerEnergy", PHYS, inputs)


def test_advdiff_functor_one_rhs_2d():
    sem, qe, u0, phys, inputs = kopriva_case()
    m = sem.mesh
    _compare(sem, qe, u0, 1, 2, "AdvDiff", phys, inputs, caches=ref.setup_assembler([m.ip2gip], [m.gip2owner]))


def test_advdiff_functor_one_rhs_3d():
    spec = box3d((4, 3, 2), 4, warp=0.05, periodic=(True, True, False))
    sem = sem_setup(spec, 1)[0]
    m = sem.mesh
    u0 = np.exp(-((m.x - 5000.0) / 2000.0) ** 2 - ((m.y - 5000.0) / 2000.0) ** 2 - ((m.z - 3000.0) / 2000.0) ** 2)
    qe = np.zeros((m.npoin, 2), order="F")
    inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": True, "mu": [0.1], "dt": 0.1, "ode_solver": "SSPRK54"}
    _compare(sem, qe, np.ascontiguousarray(u0), 1, 2, "AdvDiff", advdiff_packed(0.5, 1.0, 0.25), inputs,
             caches=ref.setup_assembler([m.ip2gip], [m.gip2owner]))


def test_advdiff_kopriva_golden_end_state_on_gpu():
    """test/CI-ref/AdvDiff/kopriva through the CUDA path: AdvDiff functor, neqs = 1 viscous pass, periodic twins through the
    assembler self lists, SSPRK54 via jx_step, then the integrator's clipped final step to tend = 10 (see
    tests/test_oracle_golden.py::test_advdiff_kopriva_golden_end_state, where the CPU restatement lands at 1e-12)."""
    from jexpresso_b200.physics import SCHEME_SSPRK54
    sem, qe, u0, phys, inputs = kopriva_case()
    p = jrhs.params_setup(sem, qe, inputs, eqs="AdvDiff", phys=phys, pow_mode=1, dss_mode=0)
    try:
        u = u0.copy()
        t = jrhs.time_loop_bang(inputs, p, u, 2000)
        p.ctx.set_state(u)
        p.ctx.step(SCHEME_SSPRK54, t, 10.0 - t, 1)
        u = p.ctx.get_state()
    finally:
        p.close()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "AdvDiff_kopriva.npz"))
    worst = float(np.max(np.abs(np.sort(u) - np.sort(g["q1"]))))
    assert worst < 1e-9, worst


@pytest.mark.parametrize("lvisc", [False, True])
def test_les_source_functor_one_rhs(lvisc):
    """problems/CompEuler/LESICP1 (first part): theta-form fluxes with the sponge + Coriolis + geostrophic source
    (three source components, node coordinates, reference state) on the generic kernel.  Without the sponge every operation is
    IEEE-exact on both sides: bit-identical to the oracle; with it the one sinpi differs by <= 2 ulp between CUDA and libm."""
    from jexpresso_b200.physics import EQ_EULER_THETA_LES, les_packed
    from helpers import MU3, euler_case
    spec = box3d((4, 3, 5), 4, warp=0.05)
    sems, qns, qes, us = euler_case(spec, 1, lpert=False)
    qes[0][:, 1] = 10.0 * qes[0][:, 0]
    qes[0][:, 2] = 2.0 * qes[0][:, 0]
    zmax = float(sems[0].mesh.z.max())
    N = sems[0].mesh.npoin
    for lsponge in (False, True):
        ph = les_packed(zmax, lsponge=lsponge, zsponge=6000.0)
        prob = ref.RefProblem(sems[0], qes[0], eq_id=EQ_EULER_THETA_LES, lpert=False, lsource=True, lvisc=lvisc, visc_coeff=MU3,
                              phys=ph, pow_mode=1, neqs=5)
        run = ref.RefRun([prob])
        uo, duo = [us[0].copy()], [np.zeros_like(us[0])]
        run.rhs(duo, uo, 0.0)
        inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": lvisc, "mu": MU3, "dt": 0.1, "ode_solver": "CarpenterKennedy2N54"}
        for dss in (0, 1):
            p = jrhs.params_setup(sems[0], qes[0], inputs, eqs="CompEulerLES", phys=ph, pow_mode=1, dss_mode=dss)
            try:
                assert p.ctx.kernel_variant() == 0
                u, du = us[0].copy(), np.empty_like(us[0])
                jrhs.rhs_bang(du, u, p, 0.0)
            finally:
                p.close()
            assert np.array_equal(u, uo[0])
            if dss == 0 and not lsponge:
                assert np.array_equal(du, duo[0]), rel_err_per_node(du, duo[0])
            for e in range(5):
                pn, l2 = rel_err_per_node(du[e * N:(e + 1) * N], duo[0][e * N:(e + 1) * N])
                assert pn <= 1e-12 and l2 <= 1e-10, (lsponge, dss, e, pn, l2)
