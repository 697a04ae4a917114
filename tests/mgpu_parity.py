"""Multi-GPU parity check (one rank per GPU, NCCL interface exchange inside libjexrhs), launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_parity.py
Every rank evaluates one rhs! and 3 CK2N54 steps of a small periodic-xy / free-slip-z 3D CompEuler box on its own
partition; rank r compares its result with the oracle's all-ranks restatement (computed redundantly on the host).
Deterministic DSS: bit-exact; atomics DSS (bench configuration, team kernel): <= 1e-12 per node, <= 1e-10 L2.
The last case runs the interface-first split (JX_OPT_OVERLAP: exchange on a second stream beside the interior launch,
CK2N54 steps replayed as a CUDA graph).  JX_MGPU_CASES="3" (comma separated indices) selects cases.
Run under the driver by tests/test_zzzz_mgpu_nccl.py (two ranks); the CPU suite covers the same partition / assembler lists over gloo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from helpers import MU3, PHYS, box3d, euler_case, rel_err_per_node
    from jexpresso_b200 import capi
    from jexpresso_b200 import rhs as jrhs
    from oracle import ref

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = (((True, True, False), True, 0, 0, (6, 4, 3), 0), ((False, False, False), False, 0, 9, (6, 4, 3), 0),
             ((True, True, False), False, 1, 9, (6, 4, 3), 0), ((False, False, False), False, 1, 9, (12, 12, 3), 2),
             # BASELINE configs[3] at its stated size (64 x 64 x 24 elements, periodic x,y, AV): opt-in, JX_MGPU_CASES=4,5
             ((True, True, False), True, 0, 0, (64, 64, 24), 0), ((True, True, False), True, 1, 0, (64, 64, 24), 4),
             # Vreman closure (jx_set_sgs; the shipped default of problems/CompEuler/3d) across rank interfaces: bit-exact
             ((True, True, False), "VREM", 0, 0, (6, 4, 3), 0))
    pick = os.environ.get("JX_MGPU_CASES")
    pick = {int(x) for x in pick.split(",")} if pick else (set(range(4)) | {6})
    for ci, (periodic, lvisc, dss, variant, nel, overlap) in enumerate(cases):
        if ci not in pick:
            continue
        visc_model = lvisc if isinstance(lvisc, str) else "AV"
        lvisc = bool(lvisc)
        box = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        big = nel[0] * nel[1] * nel[2] > 10000
        nsteps = 1 if big else 3
        spec = box3d(nel, 4, warp=0.05, periodic=periodic, L=(10000.0, 10000.0, 3750.0)) if big else box3d(nel, 4, warp=0.05, periodic=periodic)
        sems, qns, qes, us = euler_case(spec, world, lpert=False)
        mu = MU3 if visc_model == "AV" else [0.0, 1.0, 1.0, 1.0, 2.0]
        sgs, delta = None, None
        if visc_model != "AV":
            from jexpresso_b200.physics import PhysicalConst
            from jexpresso_b200.sem import effective_delta_l
            delta = effective_delta_l([s.mesh for s in sems])           # mesh.Δeffective_l: the maximum over all ranks
            sgs = dict(model=visc_model, delta=delta, lrichardson=True, ltheta_eqn=True, consts=PhysicalConst().sgs_packed())
        probs = [ref.RefProblem(s, qe, eq_id=0, lpert=False, lsource=True, lvisc=lvisc, visc_coeff=mu, phys=PHYS, pow_mode=1, sgs=sgs)
                 for s, qe in zip(sems, qes)]
        caches = ref.setup_assembler([s.mesh.ip2gip for s in sems], [s.mesh.gip2owner for s in sems])
        run = ref.RefRun(probs, caches)
        uo = [u.copy() for u in us]
        duo = [np.zeros_like(u) for u in us]
        run.rhs(duo, uo, 0.0)
        inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": lvisc, "mu": mu, "dt": 0.05 if big else 0.4,   # 39 m node spacing at C4
                  "ode_solver": "CarpenterKennedy2N54", "visc_model": visc_model, "delta_effective": delta}
        p = jrhs.params_setup(sems[rank], qes[rank], inputs, device=local, rank=rank, nranks=world, nccl_uid=box[0],
                              pow_mode=1, dss_mode=dss, elem_kernel=variant, overlap=overlap)
        split = p.ctx.split_info()
        if overlap:
            p.ctx.set_option(capi.JX_OPT_CUDA_GRAPH, 1)
        try:
            u = us[rank].copy()
            du = np.empty_like(u)
            jrhs.rhs_bang(du, u, p, 0.0)
            ug = us[rank].copy()
            jrhs.time_loop_bang(inputs, p, ug, nsteps)
        finally:
            p.close()
        us2 = [x.copy() for x in us]
        ref.time_loop(run, us2, 0.0, inputs["dt"], nsteps, scheme="CK2N54")
        pn, l2 = rel_err_per_node(du, duo[rank])
        pn2, l22 = rel_err_per_node(ug, us2[rank])
        exact = bool(np.array_equal(du, duo[rank]) and np.array_equal(ug, us2[rank]))
        good = exact if dss == 0 else (pn <= 1e-12 and l2 <= 1e-10 and pn2 <= 1e-12 and l22 <= 1e-10)
        if overlap and world <= 4:          # (with more ranks the small box has no interior pairs left: nothing to split)
            good = good and split[0] > 0 and split[1] > 0
        ok &= good
        print(f"[rank {rank}/{world}] nel={nel} periodic={periodic} visc={visc_model if lvisc else False} dss={dss} kernel={variant} overlap={overlap} split={split}: rhs pn={pn:.2e} l2={l2:.2e} "
              f"{nsteps} steps pn={pn2:.2e} l2={l22:.2e} bit_exact={exact} -> {'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
