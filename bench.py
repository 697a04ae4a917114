#!/usr/bin/env python
"""bench.py -- GDOF/s per explicit-RHS evaluation of 3D CompEuler (theta form, nop=4) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nel 73] [--nop 4] [--visc]

A "step" is ONE evaluation of rhs!(du,u,params,t) (boundary projection, fused per-element flux +
divergence kernel, DSS, interface exchange when N>1, M^-1) on the synthetic weak-scaling mesh of
BASELINE.json configs[4]: nel^3 elements per GPU (default 73^3 -> 25.15 M nodes, 125.8 M DOF per GPU).
`value` times it with the state resident in HBM; `e2e` times the same call through the reference-facing
rhs!(du,u,...) surface with HOST buffers (u H2D and du D2H inside the timed region).  Inputs (3.9 GB of
metric terms per GPU) are far larger than the 126 MB L2, so no L2 flush is needed between steps.

--impl reference times the CPU restatement of the reference's own rhs! (oracle/, the Julia reference
cannot run in this image) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "3D Euler nop=4 RHS GDOF/s"
# algorithmic bytes per unique node of one RHS evaluation (3D, q=5, Float64 data,
# Int64 connectivity), r = ((nop+1)/nop)^3 element-nodes per unique node.
def algorithmic_bytes_per_node(nop, lpert, fused_stage=False):
    q, r = 5, ((nop + 1) / nop) ** 3
    b = 8 * q + 8 * q + 8 + r * (8 * 10 + 8)
    if lpert:
        b += 8 * (q + 1)
    if fused_stage:
        b += 8 * q * 2
    return b


def elem_kernel_bytes_per_node(nop, lpert):
    """Share of the algorithmic bytes the fused element kernel moves: u (+qe) and the element records."""
    q, r = 5, ((nop + 1) / nop) ** 3
    return 8 * q + r * (8 * 10 + 8) + (8 * (q + 1) if lpert else 0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed region by an in-process NVML thread
    (5 ms period; `nvidia-smi -lms` is too coarse for a 100 ms region)."""

    def __init__(self, index):
        import threading
        self.sm, self.reasons, self.smax, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.005)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [], "samples": 0}
        if self._th is None:
            return out
        self._stop.set()
        self._th.join(timeout=2)
        if self.sm:
            out["sm_mhz"] = float(np.median(self.sm))
            out["sm_min_mhz"] = float(np.min(self.sm))
            out["power_w_max"] = float(np.max(self.power)) if self.power else None
        out["reasons"] = sorted(self.reasons)
        out["samples"] = len(self.sm)
        return out


# ----------------------------------------------------------------------------------------------
def build_problem(nel, nop, lpert, rank, nranks, warp=0.05, periodic=False):
    """This rank's SEM bundle + conditioned IC for the weak-scaling box (nel^3 elements per GPU)."""
    from helpers import box3d
    from jexpresso_b200.sem import rtb_initial_state
    from jexpresso_b200.sem.scalable import conformity4ncf_q_rank, sem_setup_rank
    px, py = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[nranks]
    L = 10000.0
    spec = box3d((nel * px, nel * py, nel), nop, warp=warp, L=(L * px, L * py, L),
                 periodic=(True, True, False) if periodic else (False, False, False))
    sem = sem_setup_rank(spec, rank, nranks)
    qn, qe = rtb_initial_state(sem.mesh, lpert, seed=1234)
    conformity4ncf_q_rank(sem, qn, 5)                 # params_setup.jl:259-297 IC conditioning
    conformity4ncf_q_rank(sem, qe, 5)
    return spec, sem, qn, qe


def cpu_sample(nel, nop, lvisc, reps):
    """Time the CPU oracle (port of the reference's rhs!) on one core on an nel^3 sample."""
    from helpers import MU3, PHYS, box3d, euler_case
    from oracle import ref
    spec = box3d((nel, nel, nel), nop, warp=0.05)
    sems, qns, qes, us = euler_case(spec, 1, lpert=False, condition=False)
    prob = ref.RefProblem(sems[0], qes[0], eq_id=0, lpert=False, lsource=True, lvisc=lvisc, visc_coeff=MU3, phys=PHYS, pow_mode=0)
    run = ref.RefRun([prob])
    dus = [np.zeros_like(us[0])]
    run.rhs(dus, us, 0.0)
    t0 = time.perf_counter()
    for _ in range(reps):
        run.rhs(dus, us, 0.0)
    dt = (time.perf_counter() - t0) / reps
    return sems[0].mesh.npoin * 5, dt


def _cpu_worker(args):
    return cpu_sample(*args)


def run_reference(a):
    """--impl reference: the reference's CPU rhs! (oracle port) on all host cores, one independent
    element partition per core (its MPI layout without the negligible interface exchange)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    build_oracle_only()
    cores = os.cpu_count() or 1
    nel = a.ref_nel
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(a.warmup, 0) and 1):
            pool.map(_cpu_worker, [(4, a.nop, a.visc, 1)] * cores)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(nel, a.nop, a.visc, a.steps)] * cores)
        wall = time.perf_counter() - t0
    dofs = sum(r[0] for r in res)
    t_step = max(r[1] for r in res)
    value = dofs / t_step / 1e9
    sample = f"{cores} independent {nel}^3-element nop={a.nop} partitions, {a.steps} rhs! evaluations each"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "cpu_sample": sample, "wall_s": wall},
            "cpu_baseline": {"value": value, "unit": "GDOF/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def build_oracle_only():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


def resolve_overlap(overlap, world, env):
    """--overlap -1 (auto): interface-first split with 4 SMs left to the exchange when there is one (N > 1), else off.
    When the split is on, the exchange's NCCL send/recv kernels must fit on those SMs (one CTA per channel, one CTA
    per SM), so the channel count is capped in ``env`` before any communicator exists."""
    if overlap < 0:
        overlap = 4 if world > 1 else 0
    if overlap > 0:
        for k in ("NCCL_MAX_CTAS", "NCCL_MAX_NCHANNELS", "NCCL_MAX_P2P_NCHANNELS"):
            env.setdefault(k, str(overlap))
    return overlap


def workload_name(a):
    return (f"CompEuler theta 3D TOTAL {'AV mu=125' if a.visc else 'inviscid'} + gravity source, nop={a.nop}, "
            f"{a.nel}^3 elements per GPU (synthetic weak-scaling mesh, BASELINE configs[4])")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nel", type=int, default=73)
    ap.add_argument("--nop", type=int, default=4)
    ap.add_argument("--visc", action="store_true")
    ap.add_argument("--pert", action="store_true")
    ap.add_argument("--dss-mode", type=int, default=int(os.environ.get("JX_DSS_MODE", "1")))
    ap.add_argument("--pow-mode", type=int, default=int(os.environ.get("JX_POW_MODE", "1")))
    ap.add_argument("--elem-kernel", type=int, default=int(os.environ.get("JX_ELEM_KERNEL", "9")))
    ap.add_argument("--graph", type=int, default=int(os.environ.get("JX_BENCH_GRAPH", "1")))
    ap.add_argument("--overlap", type=int, default=int(os.environ.get("JX_OVERLAP", "-1")),
                    help="JX_OPT_OVERLAP: interface groups first, exchange beside the interior launch; value = SMs left to the "
                         "exchange (0 = off; -1 = auto: 4 at N > 1, where there is an exchange to hide)")
    ap.add_argument("--periodic", action="store_true", help="periodic x,y box (self-exchange of the twins; small meshes only)")
    ap.add_argument("--ref-nel", type=int, default=12)
    ap.add_argument("--cpu-nel", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    if a.elem_kernel >= 1 and (a.visc or a.nop not in (2, 4)):
        a.elem_kernel = 0        # the 3D fast paths are inviscid, nop <= 4; everything else runs the generic k_elem_node
    if a.impl == "reference":
        return run_reference(a)
    a.overlap = resolve_overlap(a.overlap, int(os.environ.get("WORLD_SIZE", "1")), os.environ)

    import torch
    import torch.distributed as dist
    from jexpresso_b200 import capi
    from jexpresso_b200 import rhs as jrhs
    from helpers import MU3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    t_setup = time.perf_counter()
    spec, sem, qn, qe = build_problem(a.nel, a.nop, a.pert, rank, world, periodic=a.periodic)
    neqs = 5
    N = sem.mesh.npoin
    inputs = {"SOL_VARS_TYPE": "PERT" if a.pert else "TOTAL", "lsource": True, "lvisc": a.visc, "mu": MU3, "dt": 0.1,
              "ode_solver": "CarpenterKennedy2N54"}
    params = jrhs.params_setup(sem, qe, inputs, device=local, rank=rank, nranks=world, nccl_uid=uid,
                               dss_mode=a.dss_mode, pow_mode=a.pow_mode, elem_kernel=a.elem_kernel, overlap=a.overlap)
    ctx = params.ctx
    split = ctx.split_info()
    u0 = np.ascontiguousarray(qn[:, :neqs].reshape(-1, order="F"))
    ctx.set_state(u0)
    # global unique nodes (weak scaling: shared interface nodes counted once)
    n_owned = int(np.count_nonzero(sem.mesh.gip2owner == rank)) if world > 1 else N
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    total_dofs = sum_over_ranks(float(n_owned)) * neqs

    # ---- device-resident timing ---------------------------------------------------------------
    # The timed region replays ONE captured RHS evaluation (all kernels and, at N > 1, the NCCL send/recv groups of
    # the interface exchange) K times as a CUDA graph: the host's enqueue cost stays out of it.  If the capture is
    # refused the same K evaluations are enqueued eagerly.  Per-phase times come from a second, eager pass of K
    # evaluations with CUDA events around every phase (the element kernel's launch time for the roofline).
    graph = bool(a.graph)
    ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 1 if graph else 0)
    try:
        ctx.bench_rhs(a.warmup, fused_stage=False, phases=False)
    except capi.JexError:
        graph = False
        ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
        ctx.bench_rhs(a.warmup, fused_stage=False, phases=False)
    flag = sum_over_ranks(0.0 if graph else 1.0)          # all ranks take the same path
    if flag > 0 and graph:
        graph = False
        ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
    barrier()
    l0 = ctx.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms, _ = ctx.bench_rhs(a.steps, fused_stage=False, phases=False)
    barrier()
    launches = ctx.launch_count() - l0
    ms = max_over_ranks(ms)
    ms_graph = ms
    t_step = ms / a.steps * 1e-3
    value = total_dofs / t_step / 1e9
    ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
    ms_eager, phases = ctx.bench_rhs(a.steps, fused_stage=False, phases=True)
    barrier()
    clocks = sampler.stop() if sampler else None      # clocks sampled across both timed regions
    ms_eager = max_over_ranks(ms_eager)
    # both regions time exactly K evaluations between barriers; the headline is the faster enqueue mode
    # (the graph wins at N <= 2, eager enqueue at N = 8 where the replayed NCCL groups serialise more)
    timed_mode = "graph" if graph else "eager"
    if ms_eager < ms:
        ms, timed_mode = ms_eager, "eager"
        t_step = ms / a.steps * 1e-3
        value = total_dofs / t_step / 1e9
    # fused low-storage stage (RHS + M^-1 + RK update), reported beside the headline
    ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 1 if graph else 0)
    ctx.bench_rhs(2, fused_stage=True, phases=False)
    barrier()
    ms_f, _ = ctx.bench_rhs(a.steps, fused_stage=True, phases=False)
    ms_f = max_over_ranks(ms_f)
    ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)

    # ---- end to end through rhs!(du, u, params, t) with pinned host buffers ---------------------
    e2e = None
    if not a.no_e2e:
        uh = torch.empty(N * neqs, dtype=torch.float64).pin_memory()
        dh = torch.empty(N * neqs, dtype=torch.float64).pin_memory()
        uh.numpy()[:] = u0
        un, dn = uh.numpy(), dh.numpy()
        for _ in range(2):
            ctx.rhs(0.0, u=un, du=dn)
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(3, min(a.steps, 10))
        for _ in range(k_e2e):
            ctx.rhs(0.0, u=un, du=dn)
        barrier()
        t_e2e = max_over_ranks((time.perf_counter() - t0) / k_e2e)
        checksum = float(dn[:8].sum())
        e2e = {"value": total_dofs / t_e2e / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": int(N * neqs * 8),
               "d2h_bytes_per_step": int(N * neqs * 8), "ms_per_step": t_e2e * 1e3, "checksum": checksum,
               "api": "jexpresso_b200.capi.Context.rhs == jx_rhs(ctx,t,u_host,du_host,NULL)"}

    def shutdown():
        """All ranks tear down together: NCCL communicator first (collective), then the process group."""
        sys.stdout.flush()
        if world > 1:
            dist.barrier()
        params.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        shutdown()
        return 0

    peak, peak_src = peaks()
    traffic = None   # dram bytes of one element-kernel launch from the committed ncu --set full capture of this workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01g_elem_traffic.json")))
        if (tr["elem_kernel"], tr["nel"], tr["nop"], tr["pert"]) == (a.elem_kernel, a.nel, a.nop, bool(a.pert)) and not a.visc:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    elem_ms = phases[1] / a.steps
    elem_bytes = elem_kernel_bytes_per_node(a.nop, a.pert) * N
    achieved = elem_bytes / (elem_ms * 1e-3) / 1e9 if elem_ms > 0 else 0.0
    rhs_bytes = algorithmic_bytes_per_node(a.nop, a.pert) * N
    rhs_gbs = rhs_bytes / t_step / 1e9 * (1.0 if world == 1 else 1.0)
    line = {
        "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "nodes_per_gpu": N, "elements_per_gpu": sem.mesh.nelem, "neqs": neqs,
                   "l2": "inputs (3.9 GB metric records + 1 GB state per GPU) >> 126 MB L2; no flush needed",
                   "dss_mode": a.dss_mode, "pow_mode": a.pow_mode, "elem_kernel": a.elem_kernel, "setup_s": round(setup_s, 1),
                   "overlap": {"sms_left_to_exchange": a.overlap, "interface_groups": split[0], "interior_groups": split[1]},
                   "periodic_xy": bool(a.periodic),
                   "phase_ms_per_step": {k: round(v / a.steps, 4) for k, v in
                                         zip(("bc", "elem", "dss", "halo", "update", "aux"), phases[:6])},
                   "timing": "two timed regions of K RHS evaluations each (CUDA events, barrier + sync on both sides, max over ranks): "
                             "(a) CUDA graph replay of one captured evaluation incl. the NCCL groups, (b) eager enqueue with CUDA "
                             "events around every phase (source of phase_ms_per_step); headline = " + timed_mode,
                   "graph_ms_per_step": (ms_graph / a.steps) if graph else None,
                   "eager_ms_per_step": ms_eager / a.steps,
                   "fused_stage_ms_per_step": ms_f / a.steps,
                   "fused_stage_gdofs": total_dofs / (ms_f / a.steps * 1e-3) / 1e9},
        "roofline": {"bound": "hbm", "kernel": ("k_elem_team (fused flux + divergence per element group; variant %d)" % a.elem_kernel) if a.elem_kernel >= 8
                     else ("k_elem_node (generic fused flux + divergence%s, thread per node)" % (" + AV viscous term" if a.visc else "")) if a.elem_kernel <= 0
                     else "element kernel variant %d" % a.elem_kernel, "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": elem_bytes, "launch_ms": elem_ms,
                     "whole_rhs": {"achieved": rhs_gbs, "frac": rhs_gbs / peak, "bytes_per_node": algorithmic_bytes_per_node(a.nop, a.pert)}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
    }
    if world == 1 and not a.no_cpu:
        build_oracle_only()
        t0 = time.perf_counter()
        dofs, dt = cpu_sample(a.cpu_nel, a.nop, a.visc, 3)
        line["cpu_baseline"] = {"value": dofs / dt / 1e9, "unit": "GDOF/s", "cores": 1, "kind": "port",
                                "sample": f"oracle/jexref.c rhs! on a {a.cpu_nel}^3-element nop={a.nop} box, 3 evaluations, 1 thread",
                                "wall_s": round(time.perf_counter() - t0, 1)}
    print(json.dumps(line))
    shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
