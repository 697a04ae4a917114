"""The scalable per-rank setup (interface candidates only, torch.distributed assembly over gloo)
against the literal all-ranks restatement: integer lists bit-exact, assembled mass / normals /
conditioned IC bit-exact (same summation order)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import box2d, box3d
from jexpresso_b200.sem import sem_setup
from jexpresso_b200.sem.scalable import _lists_for_rank


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("dim", [2, 3])
def test_candidate_lists_match_literal(nranks, dim):
    spec = box3d((8, 4, 2), 2) if dim == 3 else box2d((8, 6), 3)
    full = sem_setup(spec, nranks)
    for r in range(nranks):
        lists, lid, owner = _lists_for_rank(spec, nranks, r)
        g = full[r].mesh.gip2owner.copy()
        mine = np.full_like(g, r)
        mine[lid] = owner
        assert np.array_equal(mine, g)
        for name in ("send_i", "recv_idx", "recvback_idx"):
            for peer in range(nranks):
                assert np.array_equal(getattr(lists, name)[peer], getattr(full[r].asm, name)[peer]), (r, name, peer)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import box3d as b3
        from jexpresso_b200.sem import rtb_initial_state
        from jexpresso_b200.sem.scalable import conformity4ncf_q_rank, sem_setup_rank
        spec = b3((4, 4, 2), 3, warp=0.04)
        sem = sem_setup_rank(spec, rank, world)
        qn, qe = rtb_initial_state(sem.mesh, False, seed=7)
        conformity4ncf_q_rank(sem, qn, 5)
        from jexpresso_b200.distributed import effective_delta_dist
        q.put((rank, sem.M.copy(), sem.Minv.copy(), sem.nx.copy(), sem.nz.copy(), qn.copy(), sem.mesh.gip2owner.copy(),
               effective_delta_dist(sem.mesh)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_setup_matches_literal():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        item = q.get(timeout=180)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from helpers import euler_case
    spec = box3d((4, 4, 2), 3, warp=0.04)
    sems, qns, qes, us = euler_case(spec, world, lpert=False, seed=7)
    for r in range(world):
        M, Minv, nx, nz, qn, owner, delta = got[r]
        from jexpresso_b200.sem import effective_delta_l
        assert delta == effective_delta_l([x.mesh for x in sems])      # the Allreduce(MAX) of mesh.jl:5629-5632: one value on all ranks
        assert np.array_equal(owner, sems[r].mesh.gip2owner)
        assert np.array_equal(M, sems[r].M) and np.array_equal(Minv, sems[r].Minv)
        assert np.array_equal(nx, sems[r].nx) and np.array_equal(nz, sems[r].nz)
        assert np.array_equal(qn[:, :5], qns[r][:, :5])
