"""SGS viscosity: Smagorinsky and Vreman in the oracle.  The reference holds no golden vector of a SMAG()/VREM()
deck, so the C restatement (oracle/jexref.c: compute_sgs_cache_{2,3}d, viscous_rhs_el_{2,3}d_sgs) is cross-checked here
against an independent, vectorised numpy transcription written straight from the Julia sources:

  SGS.jl:1118-1260 / 1262-1408     compute_sgs_cache! 3D, SMAG / VREM (dry: micro == 1)
  SGS.jl:1416-1535 / 1537-1655     compute_sgs_cache! 2D
  SGS.jl:1087-1109, 1663-1685      cache-reading SGS_diffusion
  rhs.jl:2582-2785 / 2275-2400     cache-reading _expansion_visc! 3D / 2D (incl. the 2D viscous-work term of energy runs)
  rhs.jl:1398-1461, 1255-1330      element drivers (Δ_effective^2 with the AMR level in 3D, Δ^2 in 2D)

The SGS term is isolated as  du(lvisc, model) - du(inviscid)  =  Minv * DSS(viscous element term); the transcription sums with
einsum (another association than the oracle's chains), so agreement is asserted to 1e-11 of the term's max norm.  CPU only."""
import numpy as np
import pytest

from helpers import PHYS, box2d, box3d, euler_case
from jexpresso_b200.physics import PhysicalConst
from jexpresso_b200.sem import effective_delta_l, sem_setup
from oracle import ref

PC = PhysicalConst()
MU_SGS3 = [0.0, 1.0, 1.0, 1.0, 2.0]          # problems/CompEuler/3d/user_inputs.jl:39
MU_SGS2 = [0.0, 1.0, 1.0, 2.0]


def _f_Ri(N2, Sij2, Ri_crit):
    Ri = np.where(Sij2 > 1e-12, N2 / np.where(Sij2 > 1e-12, Sij2, 1.0), 0.0)
    stable = (1.0 - Ri / Ri_crit) ** 2
    unstable = np.minimum(np.sqrt(np.maximum(1.0 - 16.0 * Ri, 0.0)), 3.0)
    return np.where(Ri >= Ri_crit, 0.0, np.where(Ri >= 0.0, stable, unstable)), Ri


def _numpy_sgs_term(sem, prim, mu, model, delta, lrich, ltheta, ad_lvl=None):
    """Minv * DSS of the SGS viscous element term for primitive fields prim[neqs][npoin]."""
    m = sem.mesh
    d, n, N, E = m.nsd, m.ngl, m.npoin, m.nelem
    q = len(prim)
    shape = (E,) + (n,) * d
    conn = np.asarray(m.connijk).reshape(shape, order="F") - 1
    dpsi, om = np.asarray(sem.basis["dpsi"]), np.asarray(sem.basis["omega"])
    met = [np.asarray(a).reshape(shape, order="F") for a in sem.metric_list]
    Je = met[-1]
    # J[a][b] = d xi_a / d x_b
    J = [[met[a * d + b] for b in range(d)] for a in range(d)]
    if d == 3:
        wJ = om[None, :, None, None] * om[None, None, :, None] * om[None, None, None, :] * Je
        fwd = ["ai,eajk->eijk", "aj,eiak->eijk", "ak,eija->eijk"]
        bwd = ["ik,eklm->eilm", "il,eklm->ekim", "im,eklm->ekli"]
    else:
        wJ = om[None, :, None] * om[None, None, :] * Je
        fwd = ["ai,eaj->eij", "aj,eia->eij"]
        bwd = ["ik,ekl->eil", "il,ekl->eki"]

    def grad(f):            # physical gradient of a nodal field at every element node
        fe = f[conn]
        dref = [np.einsum(s, dpsi, fe) for s in fwd]
        return [sum(dref[a] * J[a][b] for a in range(d)) for b in range(d)]

    G = [grad(prim[1 + c]) for c in range(d)]            # G[c][b] = d u_c / d x_b
    div = sum(G[c][c] for c in range(d))
    S = [[0.5 * (G[a][b] + G[b][a]) for b in range(d)] for a in range(d)]
    SijSij = sum(S[a][b] * S[a][b] for a in range(d) for b in range(d))
    Sij2 = 2.0 * SijSij
    itemp = d + 1
    rho, th = prim[0][conn], prim[itemp][conn]
    D = np.full(E, delta)
    if ad_lvl is not None and d == 3:
        D = np.ldexp(D, -np.asarray(ad_lvl))
    D2 = (D * D).reshape((E,) + (1,) * d)
    fRi, Ri = np.ones(shape), np.zeros(shape)
    if lrich:
        dthdz = grad(prim[itemp])[d - 1]                 # z in 3D, y in 2D
        with np.errstate(divide="ignore", invalid="ignore"):     # PERT: theta' is exactly 0 outside the bubble
            N2 = np.where(np.abs(th) > 1e-12, (PC.g / th) * dthdz, 0.0)
        fRi, Ri = _f_Ri(N2, Sij2, PC.Ri_crit)
    if model == "SMAG":
        mu_t = rho * (PC.C_s * PC.C_s) * D2 * np.sqrt(Sij2) * fRi
    else:
        beta = [[D2 * sum(G[a][c] * G[b][c] for c in range(d)) for b in range(d)] for a in range(d)]
        if d == 3:
            B = (beta[0][0] * beta[1][1] + beta[0][0] * beta[2][2] + beta[1][1] * beta[2][2]
                 - (beta[0][1] ** 2 + beta[0][2] ** 2 + beta[1][2] ** 2))
        else:
            B = beta[0][0] * beta[1][1] - beta[0][1] ** 2
        uu = sum(G[a][b] ** 2 for a in range(d) for b in range(d))
        ok = (uu > np.finfo(float).eps) & (B > 0.0)
        mu_t = np.where(ok, rho * (2.5 * PC.C_s * PC.C_s) * np.sqrt(np.where(ok, B / np.where(ok, uu, 1.0), 0.0)), 0.0) * fRi
    out = np.zeros((q, N))
    for e in range(q):
        if 1 <= e <= d:
            ev = (PC.mu_mol + mu_t) * mu[e]
            c = e - 1
            flux = [ev * (G[c][b] + G[b][c]) for b in range(d)]
            flux[c] = 2.0 * ev * G[c][c] - (2.0 / 3.0) * ev * div
        else:
            if e == itemp:
                kt = mu_t / (rho * PC.Pr_t)
                ed = (kt if ltheta else (PC.kappa_mol + kt)) * mu[e]
            else:
                ed = (PC.kappa_mol + mu_t / (rho * PC.Sc_t)) * mu[e]
            gs = grad(prim[e])
            flux = [ed * gs[b] for b in range(d)]
            if e == itemp and d == 2 and not ltheta:      # viscous work, rhs.jl:2361-2370
                ev = (PC.mu_mol + mu_t) * mu[1]
                txx = 2.0 * ev * G[0][0] - (2.0 / 3.0) * ev * div
                tyy = 2.0 * ev * G[1][1] - (2.0 / 3.0) * ev * div
                txy = ev * (G[0][1] + G[1][0])
                ul, vl = prim[1][conn], prim[2][conn]
                flux = [flux[0] + txx * ul + txy * vl, flux[1] + txy * ul + tyy * vl]
        el = np.zeros(shape)
        for a in range(d):
            g = sum(J[a][b] * flux[b] for b in range(d)) * wJ
            el -= np.einsum(bwd[a], dpsi, g)
        np.add.at(out[e], conn.ravel(), el.ravel())
    return (out * np.asarray(sem.Minv)[None, :]).reshape(-1), Ri, mu_t


def _oracle_du(sem, qe, u0, neqs, eq_id, lpert, lvisc, mu, sgs):
    prob = ref.RefProblem(sem, qe, eq_id=eq_id, lpert=lpert, lsource=True, lvisc=lvisc, visc_coeff=np.array(mu, float), phys=PHYS,
                          pow_mode=1, neqs=neqs, sgs=sgs)
    u, RHS = u0.copy(), np.zeros(neqs * sem.mesh.npoin)
    prob.build_rhs_local(u, RHS, 0.0)
    prob.divide_by_mass(RHS)
    return u, RHS, prob


def _theta_prims(sem, u, qe, lpert):
    N, d = sem.mesh.npoin, sem.mesh.nsd
    q = u.reshape(d + 2, N)
    if not lpert:
        return [q[0]] + [q[e] / q[0] for e in range(1, d + 2)]
    r = q[0] + qe[:, 0]
    return [r] + [q[e] / r for e in range(1, d + 1)] + [(q[d + 1] + qe[:, d + 1]) / r - qe[:, d + 1] / qe[:, 0]]


@pytest.mark.parametrize("model", ["SMAG", "VREM"])
@pytest.mark.parametrize("nsd,lpert,lrich", [(3, False, True), (3, True, True), (3, False, False), (2, False, True), (2, True, False)])
def test_oracle_sgs_term_matches_numpy_transcription(oracle_lib, model, nsd, lpert, lrich):
    spec = (box3d((3, 2, 3), 4, warp=0.05, periodic=(True, True, False)) if nsd == 3
            else box2d((5, 4), 4, warp=0.05, periodic=(True, False, False)))
    sems, qns, qes, us = euler_case(spec, 1, lpert=lpert)
    sem, qe, u0 = sems[0], qes[0], us[0]
    N, neqs = sem.mesh.npoin, nsd + 2
    mu = MU_SGS3 if nsd == 3 else MU_SGS2
    delta = effective_delta_l(sem.mesh)
    ad = (np.arange(sem.mesh.nelem) % 3) if nsd == 3 else None      # AMR levels 0, 1, 2: exercises ldexp(Δ, -ad_lvl)
    sgs = dict(model=model, delta=delta, lrichardson=lrich, ltheta_eqn=True, consts=PC.sgs_packed(), ad_lvl=ad)
    u_bc, du_inv, _ = _oracle_du(sem, qe, u0, neqs, 0, lpert, False, mu, None)
    u_bc2, du_sgs, prob = _oracle_du(sem, qe, u0, neqs, 0, lpert, True, mu, sgs)
    assert np.array_equal(u_bc, u_bc2)
    want, Ri, mu_t = _numpy_sgs_term(sem, _theta_prims(sem, u_bc, qe, lpert), mu, model, delta, lrich, True, ad)
    got = du_sgs - du_inv
    for e in range(neqs):
        sl = slice(e * N, (e + 1) * N)
        scale = np.max(np.abs(want[sl]))
        if mu[e] == 0.0:
            assert scale == 0.0 and not got[sl].any()
            continue
        assert scale > 0
        # the difference of two sums carries the rounding of the (larger) inviscid part
        tol = 1e-11 * scale + 4e-16 * np.max(np.abs(du_inv[sl]))
        assert np.max(np.abs(got[sl] - want[sl])) <= tol, (e, np.max(np.abs(got[sl] - want[sl])), scale)
    assert mu_t.max() > 0.0
    if lrich:   # every branch of the Richardson function is exercised
        assert (Ri >= PC.Ri_crit).any() and ((Ri >= 0) & (Ri < PC.Ri_crit)).any() and (Ri < 0).any()
    # the cache the oracle leaves behind: each node holds the value of the last element that wrote it
    cache = prob.sgs_mu_turb()
    conn = np.asarray(sem.mesh.connijk).reshape(sem.mesh.nelem, -1, order="F") - 1
    last = np.zeros(N)
    for iel in range(sem.mesh.nelem):                                # later elements overwrite earlier ones
        last[conn[iel]] = mu_t[iel].reshape(-1, order="F")
    assert np.max(np.abs(cache - last)) <= 1e-10 * mu_t.max()


def test_oracle_sgs_energy_equation_2d_viscous_work(oracle_lib):
    """ltheta_eqn = false (inputs[:energy_equation] == "energy"): molecular + turbulent diffusivity on T and the viscous-work
    term with the momentum viscosity on the energy equation (rhs.jl:2361-2370, SGS.jl:1674-1680)."""
    spec = box2d((5, 4), 4, warp=0.05, periodic=(True, True, False))
    sem = sem_setup(spec, 1)[0]
    N = sem.mesh.npoin
    rng = np.random.default_rng(11)
    rho = 1.0 + 0.2 * rng.uniform(-1.0, 1.0, N)
    uv = 0.3 * rng.uniform(-1.0, 1.0, (2, N))
    pres = 1.0 + 0.1 * rng.uniform(-1.0, 1.0, N)
    rE = pres / (PHYS[1] - 1.0) + 0.5 * rho * (uv[0] ** 2 + uv[1] ** 2)
    u0 = np.concatenate([rho, rho * uv[0], rho * uv[1], rE])
    qe = np.zeros((N, 5), order="F")
    delta = effective_delta_l(sem.mesh)
    mu = [0.5, 1.0, 1.0, 2.0]        # a non-zero coefficient on the density equation: the "other scalars" branch (Sc_t)
    for model in ("SMAG", "VREM"):
        sgs = dict(model=model, delta=delta, lrichardson=True, ltheta_eqn=False, consts=PC.sgs_packed())
        _, du_inv, _ = _oracle_du(sem, qe, u0, 4, 1, False, False, mu, None)
        _, du_sgs, _ = _oracle_du(sem, qe, u0, 4, 1, False, True, mu, sgs)
        r, ru, rv = u0[:N], u0[N:2 * N], u0[2 * N:3 * N]
        p = PHYS[7] * (u0[3 * N:] - 0.5 * (ru * ru + rv * rv) / r)
        prim = [r, ru / r, rv / r, p / (r * PHYS[3])]
        want, _, _ = _numpy_sgs_term(sem, prim, mu, model, delta, True, False)
        got = du_sgs - du_inv
        for e in range(4):
            sl = slice(e * N, (e + 1) * N)
            scale = np.max(np.abs(want[sl]))
            assert scale > 0
            tol = 1e-11 * scale + 4e-16 * np.max(np.abs(du_inv[sl]))
            assert np.max(np.abs(got[sl] - want[sl])) <= tol, (model, e, np.max(np.abs(got[sl] - want[sl])), scale)


def test_oracle_sgs_two_ranks_match_one_rank(oracle_lib):
    """The closure is element-local, so an element partition (two ranks, interface sums through the AssemblerCache lists) must
    reproduce the one-rank right-hand side at every global node -- provided every rank uses the GLOBAL mesh.Δeffective_l
    (MPI.Allreduce(MAX), mesh.jl:5629-5632), which is what effective_delta_l(list of meshes) restates."""
    # periodic in x and y: only the z faces carry the free-slip projection, so no node is projected by two faces (on box edges
    # the result depends on the order in which a rank visits its faces -- in the reference as well)
    spec = box3d((4, 2, 2), 4, warp=0.05, periodic=(True, True, False))
    for model in ("SMAG", "VREM"):
        out = {}
        for R in (1, 2):
            sems, qns, qes, us = euler_case(spec, R, lpert=False)
            delta = effective_delta_l([s.mesh for s in sems])
            sgs = dict(model=model, delta=delta, lrichardson=True, ltheta_eqn=True, consts=PC.sgs_packed())
            probs = [ref.RefProblem(s, qe, eq_id=0, lpert=False, lsource=True, lvisc=True, visc_coeff=np.array(MU_SGS3, float), phys=PHYS,
                                    pow_mode=1, neqs=5, sgs=sgs) for s, qe in zip(sems, qes)]
            caches = ref.setup_assembler([s.mesh.ip2gip for s in sems], [s.mesh.gip2owner for s in sems])
            run = ref.RefRun(probs, caches)
            u = [x.copy() for x in us]
            du = [np.zeros_like(x) for x in us]
            run.rhs(du, u, 0.0)
            glob = {}
            for s, d in zip(sems, du):
                N = s.mesh.npoin
                for e in range(5):
                    glob.setdefault(e, {}).update(zip(s.mesh.ip2gip.tolist(), d[e * N:(e + 1) * N].tolist()))
            out[R] = (glob, delta)
        assert out[1][1] == out[2][1]
        for e in range(5):
            a = np.array([out[1][0][e][g] for g in sorted(out[1][0][e])])
            b = np.array([out[2][0][e][g] for g in sorted(out[1][0][e])])
            if not a.any():
                assert not b.any()
                continue
            # another partition = another summation order at the interface nodes (and conditioned initial states that differ in
            # the last bit): the hydrostatic cancellation amplifies that to ~3e-12 already for the inviscid terms
            assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(a)), (model, e)
