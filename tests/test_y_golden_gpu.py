"""The CUDA path itself against the reference's OWN golden end state (not only through the oracle):
test/CI-ref/CompEuler/theta -- 2D theta-form Euler, PERT, AV mu=125, 2000 CarpenterKennedy2N54 steps of dt=0.5 on the
10x10 nop=4 box, reference tolerance atol=1e-5 (test/ci_cases.jl:57,73) -- run through params_setup / time_loop!
on the GPU (jx_step: all 10 000 stage evaluations on the device).  The HDF5 files follow Gridap's node numbering, so the
comparison is order free, exactly like tests/test_oracle_golden.py does it for the CPU restatement."""
import os

import numpy as np
import pytest

from helpers import MU2, box2d, euler_case
from jexpresso_b200 import capi
from jexpresso_b200 import rhs as jrhs

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "CompEuler_theta.npz")


@pytest.mark.parametrize("graph", [0, 1])
def test_theta_golden_end_state_on_gpu(graph):
    spec = box2d((10, 10), 4)
    sems, qns, qes, us = euler_case(spec, 1, lpert=True, seed=None)
    inputs = {"SOL_VARS_TYPE": "PERT", "lsource": True, "lvisc": True, "mu": MU2, "dt": 0.5,
              "ode_solver": "CarpenterKennedy2N54"}
    p = jrhs.params_setup(sems[0], qes[0], inputs, pow_mode=1, dss_mode=0)
    try:
        p.ctx.set_option(capi.JX_OPT_CUDA_GRAPH, graph)      # 1: every step is a replay of one captured step
        u = us[0].copy()
        t = jrhs.time_loop_bang(inputs, p, u, 2000)
    finally:
        p.close()
    g = np.load(GOLD)
    assert abs(t - 1000.0) < 1e-9
    N = sems[0].mesh.npoin
    assert N == g["q1"].shape[0] == 1681
    worst = 0.0
    for i in range(4):
        mine, gold = np.sort(u[i * N:(i + 1) * N]), np.sort(g[f"q{i + 1}"])
        worst = max(worst, float(np.max(np.abs(mine - gold))))
        assert np.allclose(mine, gold, rtol=0.0, atol=1e-5), f"variable {i + 1} outside the reference's CI tolerance"
    assert worst < 1e-7, worst      # the CPU restatement sits at 7e-11 of the Julia run; the GPU evaluates the same sequence
