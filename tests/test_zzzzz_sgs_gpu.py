"""SGS viscosity on the GPU: Smagorinsky and Vreman closures (jx_set_sgs; k_elem_node<..., VISC = 2>) against the
CPU oracle, whose restatement is cross-checked against an independent numpy transcription in tests/test_sgs_cpu.py (the
reference holds no golden vector of a SMAG()/VREM() deck: parity unpinned).  Every operation of the closure is IEEE-exact on
both sides (+, -, *, /, sqrt, fma, ldexp, fmin; the equation of state through the shared jx_pow), so the deterministic DSS
mode is held to BIT equality; the atomics mode to the north star's 1e-12 per node / 1e-10 relative L2.  The file sorts last."""
import numpy as np
import pytest

from helpers import PHYS, box2d, box3d, euler_case, rel_err_per_node
from jexpresso_b200 import rhs as jrhs
from jexpresso_b200.physics import PhysicalConst
from jexpresso_b200.sem import effective_delta_l, sem_setup
from oracle import ref

pytestmark = pytest.mark.gpu

PC = PhysicalConst()
MU_SGS3 = [0.0, 1.0, 1.0, 1.0, 2.0]          # problems/CompEuler/3d/user_inputs.jl:39 (the shipped VREM deck)
MU_SGS2 = [0.0, 1.0, 1.0, 2.0]


def _inputs(model, mu, lpert, lrich=True, energy="theta", ad_lvl=None, delta=None):
    return {"SOL_VARS_TYPE": "PERT" if lpert else "TOTAL", "lsource": True, "lvisc": True, "mu": mu, "visc_model": model,
            "lrichardson": lrich, "energy_equation": energy, "ad_lvl": ad_lvl, "delta_effective": delta, "dt": 0.4,
            "ode_solver": "CarpenterKennedy2N54"}


def _oracle(sem, qe, inputs, neqs, eq_id=0, phys=PHYS):
    sgs = dict(model=inputs["visc_model"], delta=inputs["delta_effective"], lrichardson=inputs["lrichardson"],
               ltheta_eqn=inputs["energy_equation"] != "energy", consts=PC.sgs_packed(), ad_lvl=inputs["ad_lvl"])
    return ref.RefProblem(sem, qe, eq_id=eq_id, lpert=inputs["SOL_VARS_TYPE"] == "PERT", lsource=True, lvisc=True,
                          visc_coeff=np.array(inputs["mu"], float), phys=phys, pow_mode=1, neqs=neqs, sgs=sgs)


def _one_rhs(sem, qe, u0, inputs, neqs, eqs="CompEuler", eq_id=0, phys=PHYS, dss_mode=0, caches=None):
    run = ref.RefRun([_oracle(sem, qe, inputs, neqs, eq_id, phys)], caches)
    uo, duo = [u0.copy()], [np.zeros_like(u0)]
    run.rhs(duo, uo, 0.0)
    p = jrhs.params_setup(sem, qe, inputs, eqs=eqs, phys=phys, pow_mode=1, dss_mode=dss_mode)
    try:
        assert p.ctx.kernel_variant() == 0                      # the closures run on the generic kernel
        u, du = u0.copy(), np.empty_like(u0)
        jrhs.rhs_bang(du, u, p, 0.0)
    finally:
        p.close()
    assert np.array_equal(u, uo[0]), "boundary-projected state differs from the oracle"
    return du, duo[0]


@pytest.mark.parametrize("model", ["SMAG", "VREM"])
@pytest.mark.parametrize("lpert", [False, True])
def test_sgs_3d_one_rhs_bit_exact(model, lpert):
    spec = box3d((4, 3, 3), 4, warp=0.05)
    sems, qns, qes, us = euler_case(spec, 1, lpert=lpert)
    sem = sems[0]
    ad = np.arange(sem.mesh.nelem) % 3                          # AMR levels 0, 1, 2: Δ_effective = ldexp(Δ, -ad_lvl)
    inputs = _inputs(model, MU_SGS3, lpert, ad_lvl=ad, delta=effective_delta_l(sem.mesh))
    du, want = _one_rhs(sem, qes[0], us[0], inputs, 5)
    assert np.array_equal(du, want), (model, lpert, float(np.max(np.abs(du - want))))
    # the closure is not a no-op here: the same state with AV coefficients gives another right-hand side
    inputs_av = dict(inputs, visc_model="AV")
    p = jrhs.params_setup(sem, qes[0], inputs_av, pow_mode=1, dss_mode=0)
    try:
        u, du_av = us[0].copy(), np.empty_like(us[0])
        jrhs.rhs_bang(du_av, u, p, 0.0)
    finally:
        p.close()
    assert np.max(np.abs(du_av - du)) > 1e-6 * np.max(np.abs(du))


@pytest.mark.parametrize("model", ["SMAG", "VREM"])
def test_sgs_3d_without_richardson_and_other_orders(model):
    for nop, nel in ((2, (3, 3, 2)), (5, (2, 2, 2))):
        spec = box3d(nel, nop, warp=0.05, periodic=(True, True, False))
        sems, qns, qes, us = euler_case(spec, 1, lpert=False)
        sem = sems[0]
        m = sem.mesh
        inputs = _inputs(model, MU_SGS3, False, lrich=False, delta=effective_delta_l(m))
        run = ref.RefRun([_oracle(sem, qes[0], inputs, 5)], ref.setup_assembler([m.ip2gip], [m.gip2owner]))
        uo, duo = [us[0].copy()], [np.zeros_like(us[0])]
        run.rhs(duo, uo, 0.0)
        p = jrhs.params_setup(sem, qes[0], inputs, pow_mode=1, dss_mode=0)
        try:
            u, du = us[0].copy(), np.empty_like(us[0])
            jrhs.rhs_bang(du, u, p, 0.0)
        finally:
            p.close()
        assert np.array_equal(u, uo[0]) and np.array_equal(du, duo[0]), (model, nop)


@pytest.mark.parametrize("model", ["SMAG", "VREM"])
def test_sgs_2d_theta_one_rhs_bit_exact(model):
    for lpert in (False, True):
        spec = box2d((6, 5), 4, warp=0.05)
        sems, qns, qes, us = euler_case(spec, 1, lpert=lpert)
        inputs = _inputs(model, MU_SGS2, lpert, delta=effective_delta_l(sems[0].mesh))
        du, want = _one_rhs(sems[0], qes[0], us[0], inputs, 4)
        assert np.array_equal(du, want), (model, lpert, float(np.max(np.abs(du - want))))


@pytest.mark.parametrize("model", ["SMAG", "VREM"])
def test_sgs_2d_total_energy_viscous_work(model):
    """ltheta_eqn = false: molecular + turbulent diffusivity on T, viscous work with the momentum viscosity on the energy
    equation (rhs.jl:2361-2370), Schmidt-number branch on the density equation (non-zero coefficient there)."""
    spec = box2d((6, 5), 4, warp=0.05)
    sem = sem_setup(spec, 1)[0]
    N = sem.mesh.npoin
    rng = np.random.default_rng(11)
    rho = 1.0 + 0.2 * rng.uniform(-1.0, 1.0, N)
    uv = 0.3 * rng.uniform(-1.0, 1.0, (2, N))
    pres = 1.0 + 0.1 * rng.uniform(-1.0, 1.0, N)
    rE = pres / (PHYS[1] - 1.0) + 0.5 * rho * (uv[0] ** 2 + uv[1] ** 2)
    u0 = np.concatenate([rho, rho * uv[0], rho * uv[1], rE])
    qe = np.zeros((N, 5), order="F")
    inputs = _inputs(model, [0.5, 1.0, 1.0, 2.0], False, energy="energy", delta=effective_delta_l(sem.mesh))
    inputs["lsource"] = True
    du, want = _one_rhs(sem, qe, u0, inputs, 4, eqs="CompEulerEnergy", eq_id=1)
    assert np.array_equal(du, want), (model, float(np.max(np.abs(du - want))))


def test_sgs_les_deck_functor_smag():
    """problems/CompEuler/LESICP1 as shipped: SMAG() closure with the sponge / Coriolis / geostrophic source functor.  The
    sponge's sinpi differs by <= 2 ulp between CUDA and libm, so this one is held to 1e-12 instead of bit equality."""
    from jexpresso_b200.physics import EQ_EULER_THETA_LES, les_packed
    spec = box3d((3, 3, 4), 4, warp=0.05)
    sems, qns, qes, us = euler_case(spec, 1, lpert=False)
    sem = sems[0]
    qes[0][:, 1] = 10.0 * qes[0][:, 0]
    qes[0][:, 2] = 2.0 * qes[0][:, 0]
    ph = les_packed(float(sem.mesh.z.max()), lsponge=True, zsponge=6000.0)
    inputs = _inputs("SMAG", [0.0, 5.0, 5.0, 5.0, 5.0], False, delta=effective_delta_l(sem.mesh))
    du, want = _one_rhs(sem, qes[0], us[0], inputs, 5, eqs="CompEulerLES", eq_id=EQ_EULER_THETA_LES, phys=ph)
    for e in range(5):
        sl = slice(e * sem.mesh.npoin, (e + 1) * sem.mesh.npoin)
        pn, l2 = rel_err_per_node(du[sl], want[sl])
        assert pn <= 1e-12 and l2 <= 1e-10, (e, pn, l2)


def test_sgs_atomics_mode_and_steps():
    """Unordered DSS (red.global.add.f64) within the north-star bars, and three CK2N54 steps of the shipped VREM deck
    configuration through jx_step bit-identical to the oracle's stage loop in the deterministic mode."""
    spec = box3d((4, 4, 3), 4, warp=0.05)
    sems, qns, qes, us = euler_case(spec, 1, lpert=False)
    sem = sems[0]
    N = sem.mesh.npoin
    inputs = _inputs("VREM", MU_SGS3, False, delta=effective_delta_l(sem.mesh))
    du, want = _one_rhs(sem, qes[0], us[0], inputs, 5, dss_mode=1)
    for e in range(5):
        sl = slice(e * N, (e + 1) * N)
        if not want[sl].any():
            assert not du[sl].any()
            continue
        pn, l2 = rel_err_per_node(du[sl], want[sl])
        assert pn <= 1e-12 and l2 <= 1e-10, (e, pn, l2)
    run = ref.RefRun([_oracle(sem, qes[0], inputs, 5)])
    uo = [us[0].copy()]
    ref.time_loop(run, uo, 0.0, inputs["dt"], 3, scheme="CK2N54")
    p = jrhs.params_setup(sem, qes[0], inputs, pow_mode=1, dss_mode=0)
    try:
        ug = us[0].copy()
        jrhs.time_loop_bang(inputs, p, ug, 3)
    finally:
        p.close()
    assert np.array_equal(ug, uo[0]), float(np.max(np.abs(ug - uo[0])))


def test_sgs_c4_abl_box_vrem_at_stated_size():
    """BASELINE configs[3] -- the turbulent ABL-style box, 64 x 64 x 24 elements, nop 4, periodic in x and y -- with the Vreman
    closure (the shipped default of problems/CompEuler/3d/user_inputs.jl:33) at its stated size: one rhs!, bit-exact against
    the oracle (which needs about half a minute per evaluation here).  JX_C4_NEL shrinks it for quick runs."""
    import os
    nel = tuple(int(x) for x in os.environ.get("JX_C4_NEL", "64,64,24").split(","))
    spec = box3d(nel, 4, warp=0.05, periodic=(True, True, False), L=(10000.0, 10000.0, 3750.0))
    sems, qns, qes, us = euler_case(spec, 1, lpert=False)
    m = sems[0].mesh
    inputs = _inputs("VREM", MU_SGS3, False, delta=effective_delta_l(m))
    du, want = _one_rhs(sems[0], qes[0], us[0], inputs, 5, caches=ref.setup_assembler([m.ip2gip], [m.gip2owner]))
    assert np.isfinite(du).all()
    assert np.array_equal(du, want), rel_err_per_node(du, want)


def test_sgs_refused_configurations():
    from jexpresso_b200 import capi
    spec = box3d((2, 2, 2), 4)
    sems, qns, qes, us = euler_case(spec, 1, lpert=False)
    m = sems[0].mesh
    ctx = capi.Context()
    try:
        ctx.set_problem(3, m.ngl, 5, m.nelem, m.npoin, 0, False, True, False, None, PHYS)      # lvisc = 0
        with pytest.raises(capi.JexError) as ei:
            ctx.set_sgs(capi.JX_VISC_VREM, 100.0, True, True, PC.sgs_packed())
        assert ei.value.code == capi.JX_EINVAL
        ctx.set_problem(3, m.ngl, 5, m.nelem, m.npoin, 0, False, True, True, MU_SGS3, PHYS)
        with pytest.raises(capi.JexError):
            ctx.set_sgs(capi.JX_VISC_SMAG, 0.0, True, True, PC.sgs_packed())                    # no effective resolution
        with pytest.raises(capi.JexError):
            ctx.set_sgs(capi.JX_VISC_SMAG, 100.0, True, True, PC.sgs_packed()[:3])              # constants missing
        ctx.set_sgs(capi.JX_VISC_SMAG, 100.0, True, True, PC.sgs_packed())
        assert ctx.kernel_variant() == 0
        ctx.set_sgs(capi.JX_VISC_AV, 0.0)                                                        # back to AV: team kernels again
        assert ctx.kernel_variant() == 13
    finally:
        ctx.close()
