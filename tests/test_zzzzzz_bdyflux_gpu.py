"""Boundary fluxes with the Monin-Obukhov wall model on the GPU (second part; jx_upload_bdy_fluxes, k_most_faces,
k_bdy_flux_add) against the CPU oracle, whose restatement of BCs.jl:655-816 / CM_MOST.jl / surface_integral.jl is cross-checked
against an independent numpy transcription in tests/test_bdy_flux_cpu.py (parity unpinned: the reference holds no golden vector
of a bdy_fluxes deck).  The wall model calls log / atan / pow -- CUDA's here, libm's in the oracle, <= 2 ulp apart -- so the
bars are the north star's 1e-12 per node and 1e-10 relative L2 (not bit equality), in both DSS modes.  The file sorts last."""
import numpy as np
import pytest

from helpers import PHYS, most_case, rel_err_per_node
from jexpresso_b200 import rhs as jrhs
from jexpresso_b200.physics import PhysicalConst
from jexpresso_b200.sem import effective_delta_l
from oracle import ref

pytestmark = pytest.mark.gpu

BF = dict(ifirst_wall_node_index=3, delta_hf=0.25, user_heatflux=0.12)


def _inputs(lpert, **kw):
    d = {"SOL_VARS_TYPE": "PERT" if lpert else "TOTAL", "lsource": True, "lvisc": False, "mu": [0.0] * 5, "dt": 0.01,
         "ode_solver": "CarpenterKennedy2N54", "bdy_fluxes": True, "ifirst_wall_node_index": BF["ifirst_wall_node_index"],
         "delta_hf": BF["delta_hf"], "user_heatflux": BF["user_heatflux"]}
    d.update(kw)
    return d


def _check(du, want, N, bar_pn=1e-12, bar_l2=1e-10):
    for e in range(5):
        sl = slice(e * N, (e + 1) * N)
        pn, l2 = rel_err_per_node(du[sl], want[sl])
        assert pn <= bar_pn and l2 <= bar_l2, (e, pn, l2)


@pytest.mark.parametrize("dss_mode", [0, 1])
@pytest.mark.parametrize("lpert", [False, True])
def test_most_wall_fluxes_one_rhs(lpert, dss_mode):
    sem, qe, u0, Jef = most_case(lpert)
    m = sem.mesh
    N = m.npoin
    caches = ref.setup_assembler([m.ip2gip], [m.gip2owner])               # periodic in x: twins through the self lists
    prob = ref.RefProblem(sem, qe, eq_id=0, lpert=lpert, lsource=True, lvisc=False, phys=PHYS, pow_mode=1, neqs=5,
                          bdy_fluxes=dict(BF, Jef=Jef))
    run = ref.RefRun([prob], caches)
    uo, duo = [u0.copy()], [np.zeros_like(u0)]
    run.rhs(duo, uo, 0.0)
    prob0 = ref.RefProblem(sem, qe, eq_id=0, lpert=lpert, lsource=True, lvisc=False, phys=PHYS, pow_mode=1, neqs=5)
    u1, du1 = [u0.copy()], [np.zeros_like(u0)]
    ref.RefRun([prob0], caches).rhs(du1, u1, 0.0)
    assert np.max(np.abs(duo[0] - du1[0])) > 1e-6 * np.max(np.abs(duo[0])), "the wall fluxes must matter in this state"
    sem.extra["Jef"] = Jef
    p = jrhs.params_setup(sem, qe, _inputs(lpert), pow_mode=1, dss_mode=dss_mode)
    try:
        u, du = u0.copy(), np.empty_like(u0)
        jrhs.rhs_bang(du, u, p, 0.0)
    finally:
        p.close()
    assert np.array_equal(u, uo[0])
    _check(du, duo[0], N)


def test_les_deck_pipeline_smag_most_sponge_and_steps():
    """The LESICP1 deck's whole right-hand side -- theta fluxes, sponge / Coriolis / geostrophic source, SMAG() closure, MOST wall
    fluxes -- one evaluation, then two CK2N54 steps through jx_step against the oracle's stage loop."""
    from jexpresso_b200.physics import EQ_EULER_THETA_LES, les_packed
    sem, qe, u0, Jef = most_case(False)
    m = sem.mesh
    N = m.npoin
    qe = qe.copy(order="F")
    qe[:, 1] = 10.0 * qe[:, 0]
    qe[:, 2] = 2.0 * qe[:, 0]
    ph = les_packed(float(m.z.max()), lsponge=True, zsponge=2000.0)
    PC = PhysicalConst()
    mu = [0.0, 5.0, 5.0, 5.0, 5.0]                                           # problems/CompEuler/LESICP1/user_inputs.jl:34
    delta = effective_delta_l(m)
    sgs = dict(model="SMAG", delta=delta, lrichardson=True, ltheta_eqn=True, consts=PC.sgs_packed())
    caches = ref.setup_assembler([m.ip2gip], [m.gip2owner])
    prob = ref.RefProblem(sem, qe, eq_id=EQ_EULER_THETA_LES, lpert=False, lsource=True, lvisc=True, visc_coeff=np.array(mu), phys=ph,
                          pow_mode=1, neqs=5, sgs=sgs, bdy_fluxes=dict(BF, Jef=Jef))
    run = ref.RefRun([prob], caches)
    uo, duo = [u0.copy()], [np.zeros_like(u0)]
    run.rhs(duo, uo, 0.0)
    inputs = _inputs(False, lvisc=True, mu=mu, visc_model="SMAG", delta_effective=delta)
    sem.extra["Jef"] = Jef
    p = jrhs.params_setup(sem, qe, inputs, eqs="CompEulerLES", phys=ph, pow_mode=1, dss_mode=0)
    try:
        u, du = u0.copy(), np.empty_like(u0)
        jrhs.rhs_bang(du, u, p, 0.0)
        assert np.array_equal(u, uo[0])
        _check(du, duo[0], N)
        us = [u0.copy()]
        ref.time_loop(run, us, 0.0, inputs["dt"], 2, scheme="CK2N54")
        ug = u0.copy()
        jrhs.time_loop_bang(inputs, p, ug, 2)
    finally:
        p.close()
    for e in range(5):
        sl = slice(e * N, (e + 1) * N)
        assert np.max(np.abs(ug[sl] - us[0][sl])) <= 1e-12 * np.max(np.abs(us[0][sl])), e


def test_bdy_flux_refused_configurations():
    from jexpresso_b200 import capi
    sem, qe, u0, Jef = most_case(False)
    m = sem.mesh
    kinds = np.array([1 if t == "MOST" else 0 for t in m.bdy_face_type], np.int32)
    ctx = capi.Context()
    try:
        ctx.set_problem(3, m.ngl, 5, m.nelem, m.npoin, 0, False, True, False, None, PHYS)
        with pytest.raises(capi.JexError) as ei:     # before the mesh
            ctx.upload_bdy_fluxes(m.poin_in_bdy_face, m.bdy_face_in_elem, m.connijk, sem.nx, sem.ny, sem.nz, Jef, sem.basis["omega"],
                                  kinds, 3)
        assert ei.value.code == capi.JX_ESTATE
        ctx.upload_mesh(m.connijk, m.coords, sem.metric_list, sem.basis["dpsi"], sem.basis["omega"], sem.Minv, qe)
        with pytest.raises(capi.JexError):           # the inside point must be 2..ngl
            ctx.upload_bdy_fluxes(m.poin_in_bdy_face, m.bdy_face_in_elem, m.connijk, sem.nx, sem.ny, sem.nz, Jef, sem.basis["omega"],
                                  kinds, 1)
        ctx.upload_bdy_fluxes(m.poin_in_bdy_face, m.bdy_face_in_elem, m.connijk, sem.nx, sem.ny, sem.nz, Jef, sem.basis["omega"],
                              kinds, 3, 0.25, 0.12)
    finally:
        ctx.close()
