"""Synthetic problem builders shared by bench.py, __graft_entry__.smoke() and the tests: box specs of the BASELINE
configurations, the CompEuler theta rising-bubble state with the seeded momentum perturbation SURVEY.md 8d prescribes,
the north-star parity measures.  Host-side input generation only -- nothing here is on the hot path."""
import numpy as np

from ..physics import PhysicalConst
from .cases import rtb_initial_state
from .mesh import BoxSpec
from .setup import conformity4ncf_q_host, sem_setup

MU3 = [0.0, 125.0, 125.0, 125.0, 125.0]
MU2 = [0.0, 125.0, 125.0, 125.0]


def box3d(nel=(4, 4, 4), nop=4, warp=0.0, periodic=(False, False, False), L=(10000.0, 10000.0, 10000.0)):
    return BoxSpec(nsd=3, nel=tuple(nel), nop=nop, lo=(0.0, 0.0, 0.0), hi=tuple(L), periodic=periodic, warp=warp)


def box2d(nel=(10, 10), nop=4, warp=0.0, periodic=(False, False, False), lo=(-5000.0, 0.0), hi=(5000.0, 10000.0)):
    return BoxSpec(nsd=2, nel=tuple(nel), nop=nop, lo=tuple(lo), hi=tuple(hi), periodic=periodic, warp=warp)


def euler_case(spec, nranks=1, lpert=False, seed=1234, vel_amp=1.0, condition=True):
    """Per-rank (sems, qns, qes, us) for the CompEuler theta rising-bubble state with a seeded
    momentum perturbation, IC conditioned like params_setup.jl:259-297."""
    sems = sem_setup(spec, nranks)
    neqs = spec.nsd + 2
    qns, qes = [], []
    for s in sems:
        qn, qe = rtb_initial_state(s.mesh, lpert, seed=seed, vel_amp=vel_amp)
        qns.append(qn)
        qes.append(qe)
    if condition:
        conformity4ncf_q_host(sems, qns, neqs)
        conformity4ncf_q_host(sems, qes, neqs)
    us = [np.ascontiguousarray(q[:, :neqs].reshape(-1, order="F")) for q in qns]
    return sems, qns, qes, us


def rel_err_per_node(a, b):
    """max |a-b| / max(|b|, eps*||b||_inf) and relative L2 -- the north-star parity measures."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-3 * np.abs(b).max() if b.size else 1.0)
    pn = float(np.max(np.abs(a - b) / scale)) if b.size else 0.0
    l2 = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    return pn, l2


PHYS = PhysicalConst().packed()
