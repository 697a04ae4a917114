"""Analytic anchors of the oracle's 3D operators.  The reference holds no golden vector for 3D, so the restatement of
_expansion_inviscid! / _expansion_visc! / DSS_rhs! / divide_by_mass_matrix! in three dimensions is pinned here on answers that
do not come from the reference at all:

  * exactness: on an affine mesh the collocation derivative of a polynomial of degree <= nop is exact, and the DSS'ed, mass-scaled
    strong-form divergence at a node is the pointwise divergence (M = sum of the omega*J it is divided by) -- so the AdvDiff
    right-hand side of a degree-4 polynomial must equal -(u q_x + v q_y + w q_z) at EVERY node, to rounding;
  * the weak Laplacian (rhs.jl:2794-2867) of a polynomial of degree <= nop - 1 is integrated exactly by the LGL rule, so at every
    node off the boundary the AV term equals mu * laplacian(q), to rounding;
  * structure on a WARPED mesh: the AV operator annihilates constants, is symmetric in the mass inner product and negative
    semi-definite (q^T M L q = -mu * integral |grad q|^2 <= 0) -- the properties the reference's own sphere test checks for its
    surface Laplacian (test/test_sphere_visc.jl:150-176).

The CUDA path is bit-identical to this oracle in the deterministic mode (tests/test_gpu_parity.py, tests/test_zzz_functors_gpu.py),
so the anchors carry over.  CPU only."""
import numpy as np
import pytest

from helpers import box2d, box3d
from jexpresso_b200.physics import advdiff_packed
from jexpresso_b200.sem import sem_setup
from oracle import ref

WIND = (0.5, 1.0, 0.25)


def _du(sem, q, lvisc, mu, neqs=1, wind=WIND):
    N = sem.mesh.npoin
    qe = np.zeros((N, neqs + 1), order="F")
    prob = ref.RefProblem(sem, qe, eq_id=2, lpert=False, lsource=False, lvisc=lvisc, visc_coeff=np.array([mu], float),
                          phys=advdiff_packed(*wind), pow_mode=1, neqs=neqs)
    u, RHS = np.ascontiguousarray(q, dtype=float).copy(), np.zeros(N)
    prob.build_rhs_local(u, RHS, 0.0)
    prob.divide_by_mass(RHS)
    assert np.array_equal(u, q)                      # the AdvDiff hook projects nothing
    return RHS


def _poly3(x, y, z):
    """A polynomial of degree 4 in every variable (nop = 4: still in the element space), its gradient."""
    X, Y, Z = x / 1000.0, y / 1000.0, z / 1000.0
    q = 1.0 + X ** 4 - 2.0 * X * Y ** 3 + 0.5 * Z ** 4 + X ** 2 * Y * Z - 3.0 * Y ** 2 + 0.25 * X * Z ** 3
    qx = (4.0 * X ** 3 - 2.0 * Y ** 3 + 2.0 * X * Y * Z + 0.25 * Z ** 3) / 1000.0
    qy = (-6.0 * X * Y ** 2 + X ** 2 * Z - 6.0 * Y) / 1000.0
    qz = (2.0 * Z ** 3 + X ** 2 * Y + 0.75 * X * Z ** 2) / 1000.0
    return q, qx, qy, qz


@pytest.mark.parametrize("nop,nel", [(4, (3, 2, 2)), (2, (3, 3, 2))])
def test_inviscid_divergence_is_exact_for_polynomials_3d(oracle_lib, nop, nel):
    sem = sem_setup(box3d(nel, nop, warp=0.0), 1)[0]
    m = sem.mesh
    if nop == 4:
        q, qx, qy, qz = _poly3(m.x, m.y, m.z)
    else:       # degree 2 per variable
        X, Y, Z = m.x / 1000.0, m.y / 1000.0, m.z / 1000.0
        q = 1.0 + X * X - 2.0 * X * Y + 0.5 * Z * Z + Y * Z
        qx, qy, qz = (2.0 * X - 2.0 * Y) / 1000.0, (-2.0 * X + Z) / 1000.0, (Z + Y) / 1000.0
    du = _du(sem, q, False, 0.0)
    want = -(WIND[0] * qx + WIND[1] * qy + WIND[2] * qz)
    assert np.max(np.abs(du - want)) <= 1e-11 * np.max(np.abs(want))


def test_inviscid_divergence_is_exact_for_polynomials_2d(oracle_lib):
    sem = sem_setup(box2d((4, 3), 4, warp=0.0), 1)[0]
    m = sem.mesh
    X, Y = m.x / 1000.0, m.y / 1000.0
    q = 2.0 + X ** 4 - X * Y ** 3 + 0.5 * Y ** 4 + X ** 2 * Y
    qx, qy = (4.0 * X ** 3 - Y ** 3 + 2.0 * X * Y) / 1000.0, (-3.0 * X * Y ** 2 + 2.0 * Y ** 3 + X ** 2) / 1000.0
    du = _du(sem, q, False, 0.0, wind=(0.5, 1.0, 0.0))
    want = -(0.5 * qx + 1.0 * qy)
    assert np.max(np.abs(du - want)) <= 1e-11 * np.max(np.abs(want))


def test_weak_laplacian_is_exact_off_the_boundary_3d(oracle_lib):
    sem = sem_setup(box3d((3, 3, 2), 4, warp=0.0), 1)[0]
    m = sem.mesh
    X, Y, Z = m.x / 1000.0, m.y / 1000.0, m.z / 1000.0
    q = X ** 3 - 2.0 * X * Y ** 2 + 0.5 * Z ** 3 + X * Y * Z + Y ** 2 - 0.3 * Z ** 2 * X        # degree <= 3 = nop - 1
    lap = (6.0 * X + 0.0) + (-4.0 * X + 2.0) + (3.0 * Z - 0.6 * X)
    lap = lap / 1.0e6
    mu = 125.0
    term = _du(sem, q, True, mu) - _du(sem, q, False, 0.0)
    bdy = np.zeros(m.npoin, bool)
    bdy[(np.asarray(m.poin_in_bdy_face) - 1).ravel()] = True
    inner = ~bdy
    assert inner.sum() > 100
    assert np.max(np.abs(term[inner] - mu * lap[inner])) <= 1e-9 * np.max(np.abs(mu * lap))
    # on the boundary the weak form keeps the (unimposed) normal-flux term: it must NOT match there
    assert np.max(np.abs(term[bdy] - mu * lap[bdy])) > 1e-3 * np.max(np.abs(mu * lap))


def test_av_operator_structure_on_a_warped_mesh_3d(oracle_lib):
    sem = sem_setup(box3d((3, 2, 3), 4, warp=0.05), 1)[0]
    m = sem.mesh
    N = m.npoin
    M = np.asarray(sem.M) if hasattr(sem, "M") else 1.0 / np.asarray(sem.Minv)
    mu = 3.0

    def L(p):            # the AV term alone: with zero wind the inviscid part vanishes identically
        return _du(sem, p, True, mu, wind=(0.0, 0.0, 0.0))

    assert np.max(np.abs(_du(sem, np.ones(N), False, 0.0, wind=(0.0, 0.0, 0.0)))) == 0.0
    # constants are annihilated (to the rounding of the element sums)
    assert np.max(np.abs(L(np.full(N, 7.5)))) <= 1e-9 * mu * 7.5 / 100.0 ** 2
    rng = np.random.default_rng(3)
    p, q = rng.uniform(-1.0, 1.0, N), rng.uniform(-1.0, 1.0, N)
    Lp, Lq = L(p), L(q)
    a12, a21 = float(q @ (M * Lp)), float(p @ (M * Lq))
    assert abs(a12 - a21) <= 1e-10 * max(abs(a12), abs(a21))                  # symmetric in the mass inner product
    assert float(p @ (M * Lp)) < 0.0 and float(q @ (M * Lq)) < 0.0            # negative definite on non-constants
    # and it is linear
    assert np.max(np.abs(L(2.0 * p - 3.0 * q) - (2.0 * Lp - 3.0 * Lq))) <= 1e-10 * np.max(np.abs(Lp))


@pytest.mark.parametrize("visc", ["none", "AV", "SMAG", "VREM"])
def test_uniform_flow_is_preserved_on_a_warped_mesh_3d(oracle_lib, visc):
    """Free-stream preservation of the CompEuler theta functor on a warped, xy-periodic box: a uniform state has constant fluxes
    and constant primitives, so every contraction is a row sum of dpsi (zero to rounding) and what is left of rhs! is the
    gravity source, du = (0, 0, 0, -rho g, 0) -- with the AV term and with either SGS closure (no strain: mu_t = 0)."""
    from helpers import PHYS
    from jexpresso_b200.physics import PhysicalConst
    from jexpresso_b200.sem import effective_delta_l
    sem = sem_setup(box3d((3, 2, 2), 4, warp=0.05, periodic=(True, True, False)), 1)[0]
    m = sem.mesh
    N = m.npoin
    rho, uu, vv, th = 1.1, 12.0, -7.0, 300.0
    u0 = np.concatenate([np.full(N, rho), np.full(N, rho * uu), np.full(N, rho * vv), np.zeros(N), np.full(N, rho * th)])
    qe = np.zeros((N, 6), order="F")
    sgs = None
    if visc in ("SMAG", "VREM"):
        sgs = dict(model=visc, delta=effective_delta_l(m), lrichardson=True, ltheta_eqn=True, consts=PhysicalConst().sgs_packed())
    prob = ref.RefProblem(sem, qe, eq_id=0, lpert=False, lsource=True, lvisc=visc != "none",
                          visc_coeff=np.array([0.0, 125.0, 125.0, 125.0, 125.0]), phys=PHYS, pow_mode=1, neqs=5, sgs=sgs)
    run = ref.RefRun([prob], ref.setup_assembler([m.ip2gip], [m.gip2owner]))
    u, du = [u0.copy()], [np.zeros_like(u0)]
    run.rhs(du, u, 0.0)
    assert np.array_equal(u[0], u0)                    # w = 0: the free-slip projection of the z faces changes nothing
    d = du[0].reshape(5, N)
    P0 = PHYS[0] * (rho * th) ** PHYS[1]               # the pressure scale the momentum fluxes carry
    h = 10000.0 / 3 / 4                                # node spacing scale
    assert np.max(np.abs(d[0])) <= 1e-11 * rho * 12.0 / h
    for e in (1, 2):
        assert np.max(np.abs(d[e])) <= 1e-11 * P0 / h, e
    assert np.max(np.abs(d[3] + rho * PHYS[2])) <= 1e-11 * P0 / h
    assert np.max(np.abs(d[4])) <= 1e-11 * rho * th * 12.0 / h
