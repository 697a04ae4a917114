"""Full-size parity properties (BASELINE configs[4]: 73^3 elements, nop 4, 25.15 M nodes on one B200), where the CPU
oracle would need minutes per evaluation.  Size-independent properties instead (runs last: the file sorts after the
small-size parity tests, which compare every one of these kernels with the oracle bit for bit):

  * two independent implementations of the same arithmetic -- the warp-team kernel on its lane-major pair records and
    the generic thread-per-node kernel on its per-element records, both followed by the deterministic DSS gather --
    must produce IDENTICAL bits on all 125 M degrees of freedom (index arithmetic, record layouts, node -> element CSR
    at a size where 32-bit byte offsets overflow);
  * the throughput configuration the bench times (atomics DSS with M^-1 folded into the scatter weight) must agree with
    the deterministic result to <= 1e-10 relative L2 and <= 1e-12 of the field's max norm, per equation.  (The per-node
    figure relative to max(|ref|, 1e-3 max|ref|), the bar of the small-size tests, is printed, not asserted: the order
    noise of an unordered sum is ~1e-16 of the summed magnitudes, which grow like 1/h against the net value, so at
    h = 137 m the maximum over 1.25e8 nodes sits near that bar by conditioning, not by kernel.)
  * the Dirichlet projection of the state is identical in all three runs and everything is finite.

JX_FULLSIZE_NEL shrinks the box (elements per side) for a quick run.
"""
import os
import sys

import numpy as np
import pytest

from helpers import rel_err_per_node
from jexpresso_b200 import capi
from jexpresso_b200 import rhs as jrhs

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_full_size_cross_kernel_consistency():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    nel = int(os.environ.get("JX_FULLSIZE_NEL", "73"))
    neqs = 5
    spec, sem, qn, qe = bench.build_problem(nel, 4, False, 0, 1)
    N = sem.mesh.npoin
    u0 = np.ascontiguousarray(qn[:, :neqs].reshape(-1, order="F"))
    inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": False, "mu": 0.0, "dt": 0.1,
              "ode_solver": "CarpenterKennedy2N54"}

    def evaluate(elem_kernel, dss_modes):
        out = []
        p = jrhs.params_setup(sem, qe, inputs, pow_mode=1, dss_mode=dss_modes[0], elem_kernel=elem_kernel)
        try:
            for k, mode in enumerate(dss_modes):
                if k:
                    p.ctx.set_option(capi.JX_OPT_DSS_MODE, mode)
                u = u0.copy()
                du = np.empty_like(u)
                jrhs.rhs_bang(du, u, p, 0.0)
                out.append((du, u))
        finally:
            p.close()
        return out

    (du_team, u_team), (du_atom, u_atom) = evaluate(capi_elem_auto(), (0, 1))
    (du_node, u_node), = evaluate(-1, (0,))          # JX_ELEM_GENERIC
    assert np.isfinite(du_team).all() and np.isfinite(du_atom).all() and np.isfinite(du_node).all()
    assert np.array_equal(u_team, u_node) and np.array_equal(u_team, u_atom), "Dirichlet projection differs between runs"
    assert np.array_equal(du_team, du_node), ("team kernel and generic kernel differ at full size",
                                              rel_err_per_node(du_team, du_node))
    worst = 0.0
    for e in range(neqs):
        sl = slice(e * N, (e + 1) * N)
        pn, l2 = rel_err_per_node(du_atom[sl], du_team[sl])
        mx = float(np.max(np.abs(du_atom[sl] - du_team[sl])) / np.max(np.abs(du_team[sl])))
        worst = max(worst, pn)
        assert l2 <= 1e-10 and mx <= 1e-12, (e, mx, l2, pn)
    print(f"full size ({nel}^3 elements, {N} nodes): atomics vs deterministic, worst per-node figure {worst:.2e}")


def capi_elem_auto():
    return 0        # JX_ELEM_AUTO: the warp-team kernel for 3D inviscid nop 4


def _oracle_and_gpu(spec, mu, nsteps, dt):
    """One rhs! and ``nsteps`` CK2N54 steps of the CompEuler theta case on ``spec``: (oracle du, oracle u_end, GPU du, GPU u_end),
    deterministic DSS on the GPU, so the comparison is bit for bit."""
    from helpers import PHYS, euler_case
    from oracle import ref
    sems, qns, qes, us = euler_case(spec, 1, lpert=False)
    prob = ref.RefProblem(sems[0], qes[0], eq_id=0, lpert=False, lsource=True, lvisc=True, visc_coeff=mu, phys=PHYS, pow_mode=1)
    m = sems[0].mesh
    caches = ref.setup_assembler([m.ip2gip], [m.gip2owner]) if any(spec.periodic) else None
    run = ref.RefRun([prob], caches)
    uo, duo = [us[0].copy()], [np.zeros_like(us[0])]
    run.rhs(duo, uo, 0.0)
    ue = [us[0].copy()]
    ref.time_loop(run, ue, 0.0, dt, nsteps, scheme="CK2N54")
    inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": True, "mu": mu, "dt": dt, "ode_solver": "CarpenterKennedy2N54"}
    p = jrhs.params_setup(sems[0], qes[0], inputs, pow_mode=1, dss_mode=0)
    try:
        u = us[0].copy()
        du = np.empty_like(u)
        jrhs.rhs_bang(du, u, p, 0.0)
        ug = us[0].copy()
        jrhs.time_loop_bang(inputs, p, ug, nsteps)
        variant = p.ctx.kernel_variant()
    finally:
        p.close()
    return duo[0], ue[0], du, ug, variant


def test_c3_density_current_box_at_stated_size():
    """BASELINE configs[2] at its stated size: 2D, 128 x 32 elements, nop 5, AV viscous term (the viscous-kernel path,
    rhs.jl:1973-2056): one rhs! and 20 CK2N54 steps, bit-exact against the oracle."""
    from helpers import MU2, box2d
    spec = box2d((128, 32), 5, warp=0.05, lo=(0.0, 0.0), hi=(25600.0, 6400.0))
    duo, ue, du, ug, variant = _oracle_and_gpu(spec, MU2, 20, 0.02)
    assert np.isfinite(du).all() and np.isfinite(ug).all()
    assert np.array_equal(du, duo), rel_err_per_node(du, duo)
    assert np.array_equal(ug, ue), rel_err_per_node(ug, ue)


def test_c4_abl_box_at_stated_size_one_gpu():
    """BASELINE configs[3] at its stated size on one GPU: 3D, 64 x 64 x 24 elements, nop 4, periodic in x and y (the twins
    are summed through the assembler's self lists, restructure_for_periodicity.jl:1387-1577 / mpi_communications.jl:99-112),
    free-slip top and bottom, AV viscous term: one rhs! and the first CK2N54 stages of a step (JX_C4_STEPS, default 0: the
    oracle needs about 25 s per evaluation at this size), bit-exact against the oracle.  (The 2/4/8-rank NCCL runs of the same
    mesh with a whole step: tests/mgpu_parity.py cases 4 and 5, profiles/.)"""
    from helpers import MU3, box3d
    nel = tuple(int(x) for x in os.environ.get("JX_C4_NEL", "64,64,24").split(","))
    spec = box3d(nel, 4, warp=0.05, periodic=(True, True, False), L=(10000.0, 10000.0, 3750.0))
    duo, ue, du, ug, variant = _oracle_and_gpu(spec, MU3, int(os.environ.get("JX_C4_STEPS", "0")), 0.05)
    assert variant == 13, variant      # the warp-team kernels (k_elem_team + k_visc_quad) carry this configuration
    assert np.isfinite(du).all() and np.isfinite(ug).all()
    assert np.array_equal(du, duo), rel_err_per_node(du, duo)
    assert np.array_equal(ug, ue), rel_err_per_node(ug, ue)
