#!/usr/bin/env python
"""bench.py -- GDOF/s per explicit-RHS evaluation of 3D CompEuler (theta form, nop=4) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c5|c2|c3|c4] [--nel 73] [--nop 4] [--visc]

--config picks one of BASELINE.json's configurations: c5 (default) = synthetic weak-scaling mesh, nel^3 elements nop 4 (or 7)
per GPU, the configuration the metric is quoted on; c2 = 3D rising bubble 10^3 elements nop 4 with AV; c3 = 2D density-current
box 128 x 32 elements nop 5 with AV; c4 = 64 x 64 x 24 elements nop 4, periodic x,y, AV, STRONG scaling over the ranks.

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
def build_problem(nel, nop, lpert, rank, nranks, warp=0.05, periodic=False, config="c5"):
    """This rank's SEM bundle + conditioned IC.  c5: weak-scaling box (nel^3 elements per GPU); c2 / c3 / c4: the fixed
    meshes of BASELINE.json configs[1..3], partitioned over the ranks like the reference's _compute_xy_partition."""
    from jexpresso_b200.sem.problems import box2d, box3d
    from jexpresso_b200.sem import rtb_initial_state
    from jexpresso_b200.sem.scalable import conformity4ncf_q_rank, sem_setup_rank
    L = 10000.0
    if config == "c2":
        spec = box3d((10, 10, 10), 4, warp=warp)
    elif config == "c3":
        spec = box2d((128, 32), 5, warp=warp, lo=(0.0, 0.0), hi=(25600.0, 6400.0))
    elif config == "c4":
        spec = box3d((64, 64, 24), 4, warp=warp, L=(L, L, 0.375 * L), periodic=(True, True, False))
    else:
        px, py = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[nranks]
        spec = box3d((nel * px, nel * py, nel), nop, warp=warp, L=(L * px, L * py, L),
                     periodic=(True, True, False) if periodic else (False, False, False))
    sem = sem_setup_rank(spec, rank, nranks)
    neqs = spec.nsd + 2
    qn, qe = rtb_initial_state(sem.mesh, lpert, seed=1234)
    conformity4ncf_q_rank(sem, qn, neqs)              # params_setup.jl:259-297 IC conditioning
    conformity4ncf_q_rank(sem, qe, neqs)
    return spec, sem, qn, qe


CONFIGS = {
    # name: (nsd, nop, AV viscous term, scaling, description)
    "c5": (3, None, None, "weak", "synthetic weak-scaling mesh, BASELINE configs[4]"),
    "c2": (3, 4, True, "strong", "3D rising thermal bubble, 10x10x10 elements, AV mu=125, BASELINE configs[1]"),
    "c3": (2, 5, True, "strong", "2D density-current box, 128x32 elements, AV mu=125, BASELINE configs[2]"),
    "c4": (3, 4, True, "strong", "3D ABL-style box, 64x64x24 elements, periodic x,y, AV mu=125, BASELINE configs[3]"),
}
CK_A1, CK_B1 = -567301805773.0 / 1357537059087.0, 5161836677717.0 / 13612068292357.0   # second CK2N54 stage (jx_bench_rhs fused stage)


def cpu_sample(nel, nop, lvisc, reps, config="c5"):
    """Time the CPU oracle (port of the reference's rhs!) plus the 2N low-storage stage update on ONE core: an nel^3-element
    box of the same discretisation (c5), or the configuration's own mesh cut down to at most nel^nsd elements."""
    from jexpresso_b200.sem.problems import MU2, MU3, PHYS, box2d, box3d, euler_case
    from oracle import ref
    nsd = CONFIGS[config][0]
    if nsd == 2:
        spec = box2d((min(128, nel * 4), min(32, nel)), nop, warp=0.05, lo=(0.0, 0.0), hi=(25600.0, 6400.0))
    else:
        per = (True, True, False) if config == "c4" else (False, False, False)
        spec = box3d((nel, nel, nel), nop, warp=0.05, periodic=per)
    sems, qns, qes, us = euler_case(spec, 1, lpert=False, condition=False)
    neqs = nsd + 2
    prob = ref.RefProblem(sems[0], qes[0], eq_id=0, lpert=False, lsource=True, lvisc=lvisc, visc_coeff=MU3 if nsd == 3 else MU2,
                          phys=PHYS, pow_mode=0)
    m = sems[0].mesh
    caches = ref.setup_assembler([m.ip2gip], [m.gip2owner]) if config == "c4" else None
    run = ref.RefRun([prob], caches)
    dus = [np.zeros_like(us[0])]
    tmp = np.zeros_like(us[0])
    run.rhs(dus, us, 0.0)
    t0 = time.perf_counter()
    for _ in range(reps):
        run.rhs(dus, us, 0.0)
        tmp *= CK_A1                       # a dt = 0 stage: full stage traffic, state unchanged (as jx_bench_rhs)
        tmp += 0.0 * dus[0]
        us[0] += CK_B1 * tmp
    dt = (time.perf_counter() - t0) / reps
    return m.npoin * neqs, dt


def _cpu_worker(args):
    return cpu_sample(*args)


def run_reference(a):
    """--impl reference: the reference's CPU rhs! (oracle port) + stage update on all host cores, one independent element
    partition per core (its MPI layout without the negligible interface exchange).  Every partition (ref_nel^3 elements,
    ~140 MB of metric terms at 24^3) exceeds the per-core share of the last-level cache."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    build_oracle_only()
    cores = os.cpu_count() or 1
    nel = a.ref_nel
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(a.warmup, 0) and 1):
            pool.map(_cpu_worker, [(4, a.nop, a.visc, 1, a.config)] * cores)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(nel, a.nop, a.visc, a.steps, a.config)] * cores)
        wall = time.perf_counter() - t0
    dofs = sum(r[0] for r in res)
    t_step = max(r[1] for r in res)
    value = dofs / t_step / 1e9
    sample = (f"{cores} independent {nel}^{CONFIGS[a.config][0]}-element nop={a.nop} partitions ({res[0][0]} DOF each), "
              f"{a.steps} rhs! evaluations + stage updates each, one process per core")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": CONFIGS[a.config][3],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
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
    The NCCL channel cap that keeps the exchange kernels on those SMs is applied inside jx_init (ncclCommInitRankConfig);
    the environment variables are set as well for NCCL builds that ignore the config fields."""
    if overlap < 0:
        overlap = 4 if world > 1 else 0
    if overlap > 0:
        for k in ("NCCL_MAX_CTAS", "NCCL_MAX_NCHANNELS", "NCCL_MAX_P2P_NCHANNELS"):
            env.setdefault(k, str(overlap))
    return overlap


def workload_name(a):
    nsd, _, _, _, desc = CONFIGS[a.config]
    size = f"{a.nel}^3 elements per GPU" if a.config == "c5" else "whole mesh partitioned over the GPUs"
    return (f"CompEuler theta {nsd}D TOTAL {('AV mu=125' if a.visc_model == 'AV' else a.visc_model + ' SGS closure') if a.visc else 'inviscid'} + gravity source, nop={a.nop}, {size} ({desc})")


def elem_kernel_name(ctx_variant, a):
    if a.visc and ctx_variant == 13:
        return "k_elem_team + k_visc_quad (inviscid warp-team kernel followed by the four-warp AV viscous pass)"
    return {9: "k_elem_team (fused flux + divergence per element pair; variant 9)",
            8: "k_elem_team (variant 8)",
            12: "k_elem_tri (three-role pencil kernel, nop 7)"}.get(
        ctx_variant, "k_elem_node (generic fused flux + divergence%s, thread per node)" % ((" + AV viscous term" if a.visc_model == "AV" else " + %s closure" % a.visc_model) if a.visc else ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--nel", type=int, default=73)
    ap.add_argument("--nop", type=int, default=4)
    ap.add_argument("--visc", action="store_true")
    ap.add_argument("--visc-model", default="AV", choices=["AV", "SMAG", "VREM"],
                    help="with --visc: constant coefficients (AV) or an SGS closure (jx_set_sgs; generic element kernel), "
                         "coefficients of problems/CompEuler/3d/user_inputs.jl:39")
    ap.add_argument("--pert", action="store_true")
    ap.add_argument("--dss-mode", type=int, default=int(os.environ.get("JX_DSS_MODE", "1")))
    ap.add_argument("--pow-mode", type=int, default=int(os.environ.get("JX_POW_MODE", "1")))
    ap.add_argument("--elem-kernel", type=int, default=int(os.environ.get("JX_ELEM_KERNEL", "0")),
                    help="JX_OPT_ELEM_KERNEL: 0 = JX_ELEM_AUTO (fastest exact-order kernel of the configuration), -1 generic, 8 / 9 / 12 / 13 a variant")
    ap.add_argument("--graph", type=int, default=int(os.environ.get("JX_BENCH_GRAPH", "1")),
                    help="1: the timed regions replay one captured evaluation as a CUDA graph (declared mode); 0: eager enqueue")
    ap.add_argument("--overlap", type=int, default=int(os.environ.get("JX_OVERLAP", "-1")),
                    help="JX_OPT_OVERLAP: interface groups first, exchange beside the interior launch; value = SMs left to the "
                         "exchange (0 = off; -1 = auto: 4 at N > 1, where there is an exchange to hide)")
    ap.add_argument("--periodic", action="store_true", help="c5 only: periodic x,y box (self-exchange of the twins; small meshes only)")
    ap.add_argument("--ref-nel", type=int, default=24)
    ap.add_argument("--cpu-nel", type=int, default=24)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    nsd, cnop, cvisc, scaling, _ = CONFIGS[a.config]
    if cnop is not None:
        a.nop, a.visc = cnop, cvisc
    if a.impl == "reference":
        return run_reference(a)
    a.overlap = resolve_overlap(a.overlap, int(os.environ.get("WORLD_SIZE", "1")), os.environ)

    import torch
    import torch.distributed as dist
    from jexpresso_b200 import capi
    from jexpresso_b200 import rhs as jrhs
    from jexpresso_b200.sem.problems import MU2, MU3

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
    spec, sem, qn, qe = build_problem(a.nel, a.nop, a.pert, rank, world, periodic=a.periodic, config=a.config)
    neqs = nsd + 2
    N = sem.mesh.npoin
    inputs = {"SOL_VARS_TYPE": "PERT" if a.pert else "TOTAL", "lsource": True, "lvisc": a.visc, "mu": MU3 if nsd == 3 else MU2,
              "dt": 0.1, "ode_solver": "CarpenterKennedy2N54"}
    if a.visc and a.visc_model != "AV":
        from jexpresso_b200.sem import element_sizes
        inputs.update(visc_model=a.visc_model, mu=[0.0, 1.0, 1.0, 1.0, 2.0] if nsd == 3 else [0.0, 1.0, 1.0, 2.0])
        dmax = float(element_sizes(sem.mesh).max())
        if world > 1:                               # mesh.Δeffective_l is a global maximum (mesh.jl:5629-5632)
            tmax = torch.tensor([dmax], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dmax = float(tmax.item())
        inputs["delta_effective"] = dmax / sem.mesh.nop
    params = jrhs.params_setup(sem, qe, inputs, device=local, rank=rank, nranks=world, nccl_uid=uid,
                               dss_mode=a.dss_mode, pow_mode=a.pow_mode, elem_kernel=a.elem_kernel, overlap=a.overlap)
    ctx = params.ctx
    variant = ctx.kernel_variant()
    split = ctx.split_info()
    u0 = np.ascontiguousarray(qn[:, :neqs].reshape(-1, order="F"))
    ctx.set_state(u0)
    # global unique nodes (shared interface nodes counted once)
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
    # ONE declared enqueue mode for every timed region: replay of one captured evaluation (all kernels and, at N > 1, the
    # NCCL send/recv groups of the interface exchange) as a CUDA graph, or -- if --graph 0, or if the capture is refused on
    # any rank -- eager enqueue.  Timed regions of exactly K evaluations each, CUDA events, barrier + sync on both sides,
    # max over ranks:  (1) rhs! + M^-1 + 2N low-storage stage update = the headline (t_RHS includes the stage
    # axpy);  (2) rhs! alone;  (3) rhs! alone, eager, with CUDA events around every phase (element-kernel launch time for the
    # roofline; never the headline).
    graph = bool(a.graph)
    ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 1 if graph else 0)
    try:
        ctx.bench_rhs(a.warmup, fused_stage=True, phases=False)
    except capi.JexError:
        graph = False
        ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
        ctx.bench_rhs(a.warmup, fused_stage=True, phases=False)
    if sum_over_ranks(0.0 if graph else 1.0) > 0 and graph:   # all ranks take the same path
        graph = False
        ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
    timed_mode = "graph" if graph else "eager"
    barrier()
    l0 = ctx.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_f, _ = ctx.bench_rhs(a.steps, fused_stage=True, phases=False)
    barrier()
    launches = ctx.launch_count() - l0
    ms_f = max_over_ranks(ms_f)
    t_step = ms_f / a.steps * 1e-3
    value = total_dofs / t_step / 1e9
    ctx.bench_rhs(a.warmup, fused_stage=False, phases=False)
    barrier()
    ms_r, _ = ctx.bench_rhs(a.steps, fused_stage=False, phases=False)
    barrier()
    ms_r = max_over_ranks(ms_r)
    ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 0)
    ms_eager, phases = ctx.bench_rhs(a.steps, fused_stage=False, phases=True)
    barrier()
    clocks = sampler.stop() if sampler else None      # clocks sampled across the timed regions
    ms_eager = max_over_ranks(ms_eager)

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
        # achieved copy rates of this rank while every rank copies at once (what saturates at N = 8: see DESIGN.md section 6)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.set_state(un)
        h2d_gbs = 3 * N * neqs * 8 / (time.perf_counter() - t0) / 1e9
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.get_state_into(dn)
        d2h_gbs = 3 * N * neqs * 8 / (time.perf_counter() - t0) / 1e9
        barrier()
        e2e = {"value": total_dofs / t_e2e / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": int(N * neqs * 8),
               "d2h_bytes_per_step": int(N * neqs * 8), "ms_per_step": t_e2e * 1e3, "checksum": checksum,
               "h2d_gbs_this_rank": h2d_gbs, "d2h_gbs_this_rank": d2h_gbs,
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
        for tr in json.load(open(os.path.join(ROOT, "profiles", "elem_traffic.json"))):
            if (tr["config"], tr["elem_kernel"], tr["nel"], tr["nop"], tr["pert"], tr["visc"]) == \
                    (a.config, variant, a.nel, a.nop, bool(a.pert), bool(a.visc)):
                traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    elem_ms = phases[1] / a.steps
    r_el = ((a.nop + 1) / a.nop) ** nsd
    elem_bytes = (8 * neqs + r_el * (8 * (nsd * nsd + 1) + 8) + (8 * (neqs + 1) if a.pert else 0)) * N
    achieved = elem_bytes / (elem_ms * 1e-3) / 1e9 if elem_ms > 0 else 0.0
    rhs_bytes_node = 8 * neqs * 2 + 8 + r_el * (8 * (nsd * nsd + 1) + 8) + (8 * (neqs + 1) if a.pert else 0)
    fused_bytes = (rhs_bytes_node + 8 * neqs * 2) * N
    line = {
        "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "value_rhs_only": total_dofs / (ms_r / a.steps * 1e-3) / 1e9,
        "config": {"workload": workload_name(a), "config": a.config, "nodes_per_gpu": N, "elements_per_gpu": sem.mesh.nelem, "neqs": neqs,
                   "l2": "inputs (metric records + state per GPU) >> 126 MB L2 at c5 / c4 sizes; no flush needed"
                         if N * neqs * 8 > 2.6e8 else "inputs fit the 126 MB L2: small-mesh configuration, launch bound",
                   "dss_mode": a.dss_mode, "pow_mode": a.pow_mode, "elem_kernel": variant, "setup_s": round(setup_s, 1),
                   "overlap": {"sms_left_to_exchange": a.overlap, "interface_groups": split[0], "interior_groups": split[1]},
                   "periodic_xy": bool(a.periodic) or a.config == "c4",
                   "phase_ms_per_step": {k: round(v / a.steps, 4) for k, v in
                                         zip(("bc", "elem", "dss", "halo", "update", "aux"), phases[:6])},
                   "timing": "value = K evaluations of rhs! + M^-1 + 2N low-storage stage update (SURVEY 8d t_RHS), value_rhs_only = K "
                             "evaluations of rhs! alone; both in ONE declared enqueue mode (" + timed_mode + "), CUDA events, barrier + sync "
                             "on both sides, max over ranks; phase_ms_per_step from a third, eager region with events around every phase",
                   "enqueue_mode": timed_mode,
                   "rhs_only_ms_per_step": ms_r / a.steps,
                   "rhs_only_eager_with_phase_events_ms_per_step": ms_eager / a.steps},
        "roofline": {"bound": "hbm", "kernel": elem_kernel_name(variant, a), "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": elem_bytes, "launch_ms": elem_ms,
                     "whole_step": {"achieved": fused_bytes / t_step / 1e9, "frac": fused_bytes / t_step / 1e9 / peak,
                                    "bytes_per_node": fused_bytes / N}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
    }
    if world == 1 and not a.no_cpu:
        build_oracle_only()
        t0 = time.perf_counter()
        reps = 10
        dofs, dt = cpu_sample(a.cpu_nel, a.nop, a.visc, reps, a.config)
        line["cpu_baseline"] = {"value": dofs / dt / 1e9, "unit": "GDOF/s", "cores": 1, "kind": "port",
                                "sample": f"oracle/jexref.c rhs! + stage update on {dofs} DOF of the same discretisation "
                                          f"({a.cpu_nel}^{nsd} elements, nop={a.nop}, working set beyond the last-level cache), "
                                          f"{reps} evaluations, 1 thread",
                                "wall_s": round(time.perf_counter() - t0, 1)}
    print(json.dumps(line))
    shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
