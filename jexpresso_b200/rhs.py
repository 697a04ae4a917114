"""Host-side mirror of the reference's operator surface for the explicit RHS path.

    reference (Julia)                                   here
    ------------------------------------------------    ------------------------------------
    params_setup(sem, qp, inputs, ...)                  params_setup(sem, qe, inputs, ...)
        src/kernel/infrastructure/params_setup.jl:1
    rhs!(du, u, params, time)                           rhs_bang(du, u, params, time)
        src/kernel/operators/rhs.jl:121-134
    time_loop!(inputs, params, u, ...) -> solve(...)    time_loop_bang(inputs, params, u, nsteps)
        src/kernel/solvers/TimeIntegrators.jl:191,597

``inputs`` is the reference's ``Dict{Symbol,Any}`` with string keys (``:lvisc`` -> "lvisc",
``:μ`` -> "mu", ``:Δt`` -> "dt", ``:SOL_VARS_TYPE`` -> "PERT"|"TOTAL", ``:ode_solver`` ->
"CarpenterKennedy2N54"|"SSPRK54"|"SSPRK33", ``:visc_model`` -> "AV"|"SMAG"|"VREM", ``:lrichardson``,
``:energy_equation`` -> "theta"|"energy", ``:bdy_fluxes``, ``:ifirst_wall_node_index``, ``:δhf`` -> "delta_hf",
``:user_heatflux``).  Everything numeric runs in libjexrhs on the GPU
through the C ABI (capi.py); this module only marshals arrays.  Semantics kept from the
reference: ``rhs!`` is in place, returns nothing, and *mutates u* (the Dirichlet projection
writes the ODE state, BCs.jl:651).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import capi
from .physics import (BC_FREE_SLIP, BC_SKIP, EQ_ADVDIFF, EQ_EULER_ENERGY, EQ_EULER_THETA, EQ_EULER_THETA_LES, EQ_SHALLOW_WATER,
                      SCHEME_CK2N54, SCHEME_SSPRK33, SCHEME_SSPRK54, VISC_AV, VISC_SMAG, VISC_VREM, PhysicalConst)

__all__ = ["Params", "params_setup", "rhs_bang", "time_loop_bang", "face_kinds", "float32_dt"]

_PERIODIC_TAGS = {"periodicx", "periodicy", "periodicz", "periodic1", "periodic2", "periodic3", "Laguerre"}
_SCHEMES = {"CarpenterKennedy2N54": SCHEME_CK2N54, "SSPRK54": SCHEME_SSPRK54, "SSPRK33": SCHEME_SSPRK33}
_VISC_MODELS = {"AV": VISC_AV, "SMAG": VISC_SMAG, "VREM": VISC_VREM}
_EQS = {"CompEuler": EQ_EULER_THETA, "CompEulerEnergy": EQ_EULER_ENERGY, "AdvDiff": EQ_ADVDIFF,
        "ShallowWater": EQ_SHALLOW_WATER, "CompEulerLES": EQ_EULER_THETA_LES}


def face_kinds(tags):
    """BCs.jl:621-623 -- faces tagged periodic* are skipped by the Dirichlet loop."""
    return np.array([BC_SKIP if t in _PERIODIC_TAGS else BC_FREE_SLIP for t in tags], np.int32)


def float32_dt(dt):
    """TimeIntegrators.jl:464-465: ``dt = Float32(Δt / 2^ad_lvl_max)`` widened back to Float64."""
    return float(np.float32(dt))


@dataclass
class Params:
    """What ``params`` is to ``rhs!``: here, the handle of the device-resident problem."""
    ctx: capi.Context
    neqs: int
    npoin: int
    inputs: dict
    sem: object = None

    def close(self):
        self.ctx.close()


def params_setup(sem, qe, inputs, *, device=0, rank=0, nranks=1, nccl_uid=None, eqs="CompEuler", phys=None,
                 dss_mode=0, pow_mode=1, elem_kernel=0, overlap=0, device_metrics=False, device_mass=False):
    """Upload one rank's SEM bundle (mesh, metrics, basis, M^-1, reference state, boundary and
    interface lists) to its GPU and return the ``params`` handle ``rhs_bang`` takes."""
    m = sem.mesh
    eq_id = _EQS[eqs] if isinstance(eqs, str) else int(eqs)
    neqs = {EQ_EULER_THETA: m.nsd + 2, EQ_EULER_ENERGY: 4, EQ_ADVDIFF: 1, EQ_SHALLOW_WATER: 3, EQ_EULER_THETA_LES: 5}[eq_id]
    lpert = inputs.get("SOL_VARS_TYPE", "TOTAL") == "PERT"
    lvisc = bool(inputs.get("lvisc", False))
    mu = np.zeros(neqs)
    if lvisc:
        mu[:] = np.broadcast_to(np.asarray(inputs.get("mu", 0.0), float), (neqs,))
    if phys is None:
        phys = PhysicalConst().packed()
    ctx = capi.Context(device=device, rank=rank, nranks=nranks, nccl_uid=nccl_uid, nccl_max_ctas=overlap if nranks > 1 else 0)
    try:
        ctx.set_option(capi.JX_OPT_DSS_MODE, dss_mode)
        ctx.set_option(capi.JX_OPT_POW_MODE, pow_mode)
        ctx.set_option(capi.JX_OPT_ELEM_KERNEL, elem_kernel)
        ctx.set_option(capi.JX_OPT_OVERLAP, overlap)
        ctx.set_problem(m.nsd, m.ngl, neqs, m.nelem, m.npoin, eq_id, lpert, bool(inputs.get("lsource", False)), lvisc, mu, phys)
        visc_model = _VISC_MODELS[inputs.get("visc_model", "AV")]
        if lvisc and visc_model != VISC_AV:
            # params.sgs = allocate_SGS(...) with the flags of params_setup.jl:249-253.  mesh.Δeffective_l is a global maximum
            # (mesh.jl:5629-5632): multi-rank callers pass it as inputs["delta_effective"]
            from .sem import effective_delta_l
            delta = inputs.get("delta_effective")
            if not delta:
                if nranks > 1:
                    raise ValueError("visc_model SMAG/VREM on several ranks: pass the GLOBAL mesh.Δeffective_l as "
                                     "inputs['delta_effective'] (jexpresso_b200.distributed.effective_delta_dist)")
                delta = effective_delta_l(m)
            ctx.set_sgs(visc_model, delta, inputs.get("lrichardson", True), inputs.get("energy_equation", "theta") != "energy",
                        inputs.get("sgs_consts") or PhysicalConst().sgs_packed(), inputs.get("ad_lvl"))
        if device_metrics or device_mass:   # build_metric_terms! on the device (jx_upload_mesh_coords): sem.metrics is not read;
            # device_mass: neither is sem.Minv -- the mass matrix is built, assembled over the halo lists and inverted on the device
            ctx.upload_mesh_coords(m.connijk, m.coords, sem.basis["dpsi"], sem.basis["omega"], None if device_mass else sem.Minv, qe)
        else:
            ctx.upload_mesh(m.connijk, m.coords, sem.metric_list, sem.basis["dpsi"], sem.basis["omega"], sem.Minv, qe)
        if m.poin_in_bdy_face.shape[0] > 0:
            ctx.upload_bcs(m.poin_in_bdy_face, sem.nx, sem.ny, sem.nz, face_kinds(m.bdy_face_type))
        if inputs.get("bdy_fluxes", False):
            # apply_boundary_conditions_neumann! (BCs.jl:35-63, 655-816): MOST wall model on the faces tagged "MOST";
            # metrics.Jef comes with the SEM bundle (sem.extra["Jef"]) or is built here from the face coordinates
            from .sem.metrics import boundary_face_jacobian
            Jef = sem.extra.get("Jef") if sem.extra.get("Jef") is not None else boundary_face_jacobian(m, sem.basis)
            kinds = np.array([capi.JX_FLUX_MOST if t == "MOST" else capi.JX_FLUX_NONE for t in m.bdy_face_type], np.int32)
            ctx.upload_bdy_fluxes(m.poin_in_bdy_face, m.bdy_face_in_elem, m.connijk, sem.nx, sem.ny, sem.nz, Jef, sem.basis["omega"],
                                  kinds, inputs["ifirst_wall_node_index"], inputs.get("delta_hf", 0.0),
                                  inputs.get("user_heatflux", 0.0), inputs.get("most_consts"))
        if sem.asm is not None and not sem.asm.is_trivial():
            ctx.upload_halo(sem.asm.send_i, sem.asm.recv_idx, sem.asm.recvback_idx)
    except Exception:
        ctx.close()
        raise
    return Params(ctx=ctx, neqs=neqs, npoin=m.npoin, inputs=dict(inputs), sem=sem)


def rhs_bang(du, u, params, time):
    """``rhs!(du, u, params, time)`` (rhs.jl:121-134) through the C ABI with host buffers: u is
    uploaded, du downloaded, and the boundary-projected state is written back into u."""
    params.ctx.rhs(time, u=u, du=du, u_back=u)
    return None


def time_loop_bang(inputs, params, u, nsteps, t0=0.0):
    """``solve(prob, inputs[:ode_solver]; dt, adaptive=false)`` for ``nsteps`` fixed steps, all
    stages on the device (jx_step); u (host) is updated in place.  Returns the final time."""
    scheme = _SCHEMES[inputs.get("ode_solver", "SSPRK54")]
    dt = float32_dt(inputs["dt"])
    params.ctx.set_state(u)
    params.ctx.step(scheme, t0, dt, nsteps)
    u[:] = params.ctx.get_state()
    return t0 + nsteps * dt
