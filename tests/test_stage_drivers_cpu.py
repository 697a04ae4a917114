"""Known-answer test of the stage drivers.  OrdinaryDiffEq is not in the reference tree, so the update forms of
CarpenterKennedy2N54 / SSPRK54 / SSPRK33 are restated from the literature (oracle/ref.py; jx_step on the device is compared with
them bit for bit in the GPU suite).  CK2N54 and SSPRK54 are pinned end to end by the reference's golden runs; SSPRK33 is not.
Here all three are pinned on an answer that needs no reference run: for a LINEAR right-hand side u' = L u one step of a
Runge-Kutta scheme is its stability polynomial in dt*L,

    SSPRK(3,3)  (Shu & Osher 1988)            1 + z + z^2/2 + z^3/6
    SSPRK(5,4)  (Spiteri & Ruuth 2002)        1 + z + z^2/2 + z^3/6 + z^4/24 + c5 z^5,  c5 = b10 b21 b32 b43 b54 (the one 5-stage chain)
    RK4(5)[2N]  (Carpenter & Kennedy 1994)    1 + z + z^2/2 + z^3/6 + z^4/24 + z^5/200

and the AdvDiff right-hand side of the oracle (constant wind, AV term, doubly periodic box: no projection) is linear.  CPU only."""
import numpy as np
import pytest

from helpers import box2d
from jexpresso_b200.physics import advdiff_packed
from jexpresso_b200.sem import sem_setup
from oracle import ref


def _linear_problem():
    sem = sem_setup(box2d((5, 4), 4, warp=0.05, periodic=(True, True, False), lo=(0.0, 0.0), hi=(10.0, 8.0)), 1)[0]
    m = sem.mesh
    qe = np.zeros((m.npoin, 2), order="F")
    prob = ref.RefProblem(sem, qe, eq_id=2, lpert=False, lsource=False, lvisc=True, visc_coeff=np.array([0.05]),
                          phys=advdiff_packed(0.5, 1.0), pow_mode=1, neqs=1)
    run = ref.RefRun([prob], ref.setup_assembler([m.ip2gip], [m.gip2owner]))
    rng = np.random.default_rng(5)
    # a smooth state that is single valued on the periodic twins
    u0 = np.sin(2 * np.pi * m.x / 10.0) * np.cos(2 * np.pi * m.y / 8.0) + 0.3 * np.cos(4 * np.pi * m.x / 10.0)
    return run, np.ascontiguousarray(u0) + 0.0 * rng.uniform(size=m.npoin)


def _apply_L(run, v):
    u, du = [v.copy()], [np.zeros_like(v)]
    run.rhs(du, u, 0.0)
    assert np.array_equal(u[0], v)
    return du[0]


@pytest.mark.parametrize("scheme", ["SSPRK33", "SSPRK54", "CK2N54"])
def test_one_step_of_a_linear_problem_is_the_stability_polynomial(oracle_lib, scheme):
    run, u0 = _linear_problem()
    # linearity of the right-hand side itself
    a, b = _apply_L(run, u0), _apply_L(run, 2.0 * u0)
    assert np.max(np.abs(b - 2.0 * a)) <= 1e-13 * np.max(np.abs(a))
    dt = 0.06
    C = ref.SSPRK54
    coef = {"SSPRK33": [1.0, 1.0, 0.5, 1.0 / 6.0],
            "SSPRK54": [1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, C["b10"] * C["b21"] * C["b32"] * C["b43"] * C["b54"]],
            "CK2N54": [1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 200.0]}[scheme]
    want, v = np.zeros_like(u0), u0.copy()
    for k, c in enumerate(coef):
        if k:
            v = dt * _apply_L(run, v)
        want = want + c * v
    us = [u0.copy()]
    ks = [np.zeros_like(u0)]
    if scheme == "SSPRK33":
        ref.step_ssprk33(run, us, 0.0, dt, ks)
    elif scheme == "SSPRK54":
        ref.step_ssprk54(run, us, 0.0, dt, ks)
    else:
        ref.step_ck2n54(run, us, 0.0, dt, [np.zeros_like(u0)], ks)
    err = np.max(np.abs(us[0] - want)) / np.max(np.abs(want))
    assert err <= 2e-13, (scheme, err)
    # the step is not the identity and the high-order terms matter at this dt: dropping the last coefficient is visible
    short = want - coef[-1] * v
    assert np.max(np.abs(short - want)) / np.max(np.abs(want)) > 1e-10
