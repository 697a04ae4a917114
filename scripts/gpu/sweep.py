#!/usr/bin/env python
"""Kernel experiment harness (not the bench contract): one problem set-up, several element-kernel variants / options,
one JSON line per variant with the per-phase times of K eager evaluations and the time of K graph replays.

    python scripts/gpu/sweep.py --nel 73 --variants 9,10 [--visc] [--nop 4] [--steps 20] [--dss 1] [--lib path]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nel", type=int, default=73)
    ap.add_argument("--nop", type=int, default=4)
    ap.add_argument("--variants", default="9,10")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dss", type=int, default=1)
    ap.add_argument("--visc", action="store_true")
    ap.add_argument("--pert", action="store_true")
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--check", action="store_true", help="compare every variant's du with the first variant's (deterministic DSS)")
    a = ap.parse_args()
    from bench import build_problem
    from jexpresso_b200.sem.problems import MU3
    from jexpresso_b200 import capi
    from jexpresso_b200 import rhs as jrhs
    t0 = time.perf_counter()
    spec, sem, qn, qe = build_problem(a.nel, a.nop, a.pert, 0, 1)
    N = sem.mesh.npoin
    u0 = np.ascontiguousarray(qn[:, :5].reshape(-1, order="F"))
    print(json.dumps({"setup_s": round(time.perf_counter() - t0, 1), "npoin": N, "nelem": sem.mesh.nelem}), flush=True)
    inputs = {"SOL_VARS_TYPE": "PERT" if a.pert else "TOTAL", "lsource": True, "lvisc": a.visc, "mu": MU3, "dt": 0.1,
              "ode_solver": "CarpenterKennedy2N54"}
    ref_du = None
    for v in [int(x) for x in a.variants.split(",")]:
        try:
            p = jrhs.params_setup(sem, qe, inputs, dss_mode=a.dss, pow_mode=1, elem_kernel=v)
        except Exception as e:
            print(json.dumps({"variant": v, "error": str(e)}), flush=True)
            continue
        try:
            ctx = p.ctx
            ctx.set_state(u0)
            ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
            ctx.bench_rhs(3, fused_stage=a.fused, phases=False)
            ms_e, ph = ctx.bench_rhs(a.steps, fused_stage=a.fused, phases=True)
            ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 1)
            ctx.bench_rhs(3, fused_stage=a.fused, phases=False)
            ms_g, _ = ctx.bench_rhs(a.steps, fused_stage=a.fused, phases=False)
            line = {"variant": v, "nel": a.nel, "nop": a.nop, "visc": a.visc, "dss": a.dss, "fused": a.fused,
                    "graph_ms": round(ms_g / a.steps, 4), "eager_ms": round(ms_e / a.steps, 4),
                    "phase_ms": {k: round(x / a.steps, 4) for k, x in zip(("bc", "elem", "dss", "halo", "update", "aux"), ph[:6])},
                    "gdofs": round(N * 5 / (ms_g / a.steps * 1e-3) / 1e9, 2)}
            if a.check:
                ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
                ctx.set_state(u0)
                ctx.rhs(0.0)
                du = ctx.get_du()
                if ref_du is None:
                    ref_du = du
                    line["check"] = "reference"
                else:
                    line["check"] = {"equal": bool(np.array_equal(du, ref_du)),
                                     "max_rel": float(np.max(np.abs(du - ref_du)) / np.max(np.abs(ref_du)))}
            print(json.dumps(line), flush=True)
        except Exception as e:
            print(json.dumps({"variant": v, "error": str(e)}), flush=True)
        finally:
            p.close()


if __name__ == "__main__":
    main()
