"""Initial / reference states for the benchmark cases (inputs of the RHS path).

Restated from problems/CompEuler/3d/initialize.jl:55-116 and
problems/CompEuler/theta/initialize.jl:48-100 (rising thermal bubble), plus the
seeded momentum perturbation SURVEY.md 8d prescribes for synthetic inputs.
"""
from __future__ import annotations

import numpy as np

from ..physics import PhysicalConst

__all__ = ["rtb_initial_state"]


def rtb_initial_state(mesh, lpert: bool, seed=None, xc=None, vel_amp=1.0):
    """Return (qn[npoin, neqs+1], qe[npoin, neqs+1]) Fortran-ordered.

    3D: bubble in the x-z plane (initialize.jl:58-83, zc=2500, r0=2000, θref=300, θc=2);
    2D: bubble in x-y with y vertical.  ``seed`` adds ρu,ρv(,ρw) ~ U(-1,1)·ρ·vel_amp drawn
    per *global* node id so that every rank sees the same value at shared nodes.
    """
    PC = PhysicalConst()
    nsd = mesh.nsd
    neqs = nsd + 2
    x = mesh.x
    zv = mesh.z if nsd == 3 else mesh.y          # vertical coordinate
    if xc is None:
        xc = (mesh.xmax + mesh.xmin) / 2
    zc, r0, thref, thc = 2500.0, 2000.0, 300.0, 2.0
    r = np.sqrt((x - xc) ** 2 + (zv - zc) ** 2)
    dth = np.where(r < r0, thc * (1.0 - r / r0), 0.0)
    th = thref + dth
    p = PC.pref * (1.0 - PC.g * zv / (PC.cp * th)) ** PC.cpoverR
    pref = PC.pref * (1.0 - PC.g * zv / (PC.cp * thref)) ** PC.cpoverR
    rho = (1.0 / th) * (p / PC.C0) ** (1.0 / PC.gamma)
    rhoref = (1.0 / thref) * (pref / PC.C0) ** (1.0 / PC.gamma)
    vel = np.zeros((mesh.npoin, nsd))
    if seed is not None:
        rng = np.random.default_rng(seed)
        table = rng.uniform(-1.0, 1.0, size=(mesh.gnpoin, nsd)) * vel_amp
        vel = table[mesh.ip2gip - 1]
    qn = np.zeros((mesh.npoin, neqs + 1), order="F")
    qe = np.zeros((mesh.npoin, neqs + 1), order="F")
    ie = neqs - 1
    if lpert:
        qn[:, 0] = rho - rhoref
        for d in range(nsd):
            qn[:, 1 + d] = rho * vel[:, d] - rhoref * vel[:, d]
            qe[:, 1 + d] = vel[:, d]
        qn[:, ie] = rho * th - rhoref * thref
    else:
        qn[:, 0] = rho
        for d in range(nsd):
            qn[:, 1 + d] = rho * vel[:, d]
            qe[:, 1 + d] = (rhoref * vel[:, d]) if nsd == 3 else vel[:, d]
        qn[:, ie] = rho * th
    qn[:, neqs] = p
    qe[:, 0] = rhoref
    qe[:, ie] = rhoref * thref
    qe[:, neqs] = pref
    if seed is not None:
        # the synthetic perturbation is a perturbation of the *state*, not of the reference
        qe[:, 1:1 + nsd] = 0.0
        if lpert:
            for d in range(nsd):
                qn[:, 1 + d] = rho * vel[:, d]
    return qn, qe
