"""Total-energy CompEuler functor (problems/CompEuler/kelvinHelmholtzChan2022): the reference holds no golden vector
for this case, so the C oracle's restatement is cross-checked here against an independent, differently organised
numpy transcription written straight from the Julia sources:

  user_flux.jl:30-48        F, G of the total-energy form
  user_primitives.jl:17-23  differentiated variables (rho, u, v, T = p / (rho Rair))
  rhs.jl:1501-1542          2D inviscid expansion
  rhs.jl:1973-2056          2D AV expansion with the viscous-work term tau.u on the energy equation (:1988, 2018-2041)
  element_matrices.jl:887-900, 972-978   DSS and M^-1

The transcription sums with numpy einsum (another association than the oracle's sequential FMA chains), so agreement is
asserted to 1e-12 of each field's max norm, not bitwise.  CPU only."""
import numpy as np

from helpers import MU2, PHYS, box2d
from jexpresso_b200.sem import sem_setup
from oracle import ref


def _numpy_energy_rhs(sem, u0, mu, phys, with_tau_u=True):
    m = sem.mesh
    n, N, E = m.ngl, m.npoin, m.nelem
    gamma, Rair, gm1 = phys[1], phys[3], phys[7]
    conn = np.asarray(m.connijk).reshape(E, n, n, order="F") - 1          # [iel, i, j]
    dpsi = np.asarray(sem.basis["dpsi"])                                  # dpsi[m, i] = L'_m(xi_i)
    om = np.asarray(sem.basis["omega"])
    xix, xiy, etx, ety, Je = [np.asarray(a).reshape(E, n, n, order="F") for a in sem.metric_list]
    q = u0.reshape(4, N)
    r, ru, rv, rE = q
    # user_flux.jl:30-48
    u, v = ru / r, rv / r
    ke = 0.5 * r * (u * u + v * v)
    P = gm1 * (rE - ke)
    F = np.stack([ru, ru * u + P, rv * u, u * (ke + gamma * P / gm1)])
    G = np.stack([rv, ru * v, rv * v + P, v * (ke + gamma * P / gm1)])
    # user_primitives.jl:17-23
    p2 = gm1 * (rE - 0.5 * (ru * ru + rv * rv) / r)
    prim = np.stack([r, ru / r, rv / r, p2 / (r * Rair)])
    RHS = np.zeros((4, N))
    wJ = om[None, :, None] * om[None, None, :] * Je
    for e in range(4):
        Fe, Ge = F[e][conn], G[e][conn]                                   # [iel, i, j]
        dFdxi = np.einsum("mi,emj->eij", dpsi, Fe)
        dFdeta = np.einsum("mj,eim->eij", dpsi, Fe)
        dGdxi = np.einsum("mi,emj->eij", dpsi, Ge)
        dGdeta = np.einsum("mj,eim->eij", dpsi, Ge)
        dFdx = dFdxi * xix + dFdeta * etx
        dGdy = dGdxi * xiy + dGdeta * ety
        rhs_el = -wJ * (dFdx + dGdy)
        # AV term: weak Laplacian of the primitive variable
        Ue = prim[e][conn]
        dqdxi = np.einsum("mi,emj->eij", dpsi, Ue)
        dqdeta = np.einsum("mj,eim->eij", dpsi, Ue)
        fx = mu[e] * (dqdxi * xix + dqdeta * etx)
        fy = mu[e] * (dqdxi * xiy + dqdeta * ety)
        if e == 3 and with_tau_u:
            Uu, Uv = prim[1][conn], prim[2][conn]
            dudxi, dudeta = np.einsum("mi,emj->eij", dpsi, Uu), np.einsum("mj,eim->eij", dpsi, Uu)
            dvdxi, dvdeta = np.einsum("mi,emj->eij", dpsi, Uv), np.einsum("mj,eim->eij", dpsi, Uv)
            dudx, dudy = dudxi * xix + dudeta * etx, dudxi * xiy + dudeta * ety
            dvdx, dvdy = dvdxi * xix + dvdeta * etx, dvdxi * xiy + dvdeta * ety
            div = dudx + dvdy
            txx = 2.0 * mu[1] * dudx - (2.0 / 3.0) * mu[1] * div
            tyy = 2.0 * mu[1] * dvdy - (2.0 / 3.0) * mu[1] * div
            txy = mu[1] * (dudy + dvdx)
            fx = fx + txx * Uu + txy * Uv
            fy = fy + txy * Uu + tyy * Uv
        gxi = (xix * fx + xiy * fy) * wJ                                  # at quadrature node (k, l)
        geta = (etx * fx + ety * fy) * wJ
        visc = -np.einsum("ik,ekl->eil", dpsi, gxi) - np.einsum("il,ekl->eki", dpsi, geta)
        np.add.at(RHS[e], conn.ravel(), (rhs_el + visc).ravel())
    return (RHS * np.asarray(sem.Minv)[None, :]).reshape(-1)


def _case():
    spec = box2d((5, 4), 4, warp=0.05, periodic=(True, True, False))     # no Dirichlet faces: flux / primitives / tau.u only
    sem = sem_setup(spec, 1)[0]
    N = sem.mesh.npoin
    rng = np.random.default_rng(11)
    rho = 1.0 + 0.2 * rng.uniform(-1.0, 1.0, N)
    uv = 0.3 * rng.uniform(-1.0, 1.0, (2, N))
    pres = 1.0 + 0.1 * rng.uniform(-1.0, 1.0, N)
    rE = pres / (PHYS[1] - 1.0) + 0.5 * rho * (uv[0] ** 2 + uv[1] ** 2)
    return sem, np.concatenate([rho, rho * uv[0], rho * uv[1], rE])


def test_oracle_energy_functor_matches_numpy_transcription(oracle_lib):
    sem, u0 = _case()
    N = sem.mesh.npoin
    qe = np.zeros((N, 5), order="F")
    prob = ref.RefProblem(sem, qe, eq_id=1, lpert=False, lsource=False, lvisc=True, visc_coeff=np.array(MU2, float), phys=PHYS,
                          pow_mode=1, neqs=4)
    # the local (pre-exchange) RHS: the periodic twins keep separate local ids, and the transcription sums per local id too
    u, RHS = u0.copy(), np.zeros(4 * N)
    prob.build_rhs_local(u, RHS, 0.0)
    prob.divide_by_mass(RHS)
    want = _numpy_energy_rhs(sem, u0, MU2, PHYS)
    for e in range(4):
        sl = slice(e * N, (e + 1) * N)
        scale = np.max(np.abs(want[sl]))
        assert scale > 0
        assert np.max(np.abs(RHS[sl] - want[sl])) <= 1e-12 * scale, e
    # the tau.u term is not negligible in this state: without it the energy equation differs visibly
    no_tau = _numpy_energy_rhs(sem, u0, MU2, PHYS, with_tau_u=False)
    sl = slice(3 * N, 4 * N)
    assert np.max(np.abs(no_tau[sl] - want[sl])) > 1e-6 * np.max(np.abs(want[sl]))
