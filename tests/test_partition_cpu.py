"""Integer host logic (bit-exact bar): the vectorised partition / ownership / assembler-list code the
product uses (jexpresso_b200/sem/partition.py) against the literal dictionary-based restatement of the
reference in oracle/ref.py (mesh.jl:3560-3610, mpi_communications.jl:1-234), plus end-to-end
properties of the interface assembly (assemble_mpi!, mpi_communications.jl:260-338)."""
import numpy as np
import pytest

from helpers import box2d, box3d
from jexpresso_b200.sem import compute_xy_partition, sem_setup
from jexpresso_b200.sem.partition import find_gip_owner_all, setup_assembler_all
from oracle import ref


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("periodic", [(False, False, False), (True, True, False)])
def test_owner_and_assembler_lists_match_literal(nranks, periodic):
    if nranks == 1 and not any(periodic):
        pytest.skip("trivial")
    spec = box3d((4, 4, 2), 2, periodic=periodic)
    sems = sem_setup(spec, nranks)
    ip2gip = [s.mesh.ip2gip for s in sems]
    owners = [s.mesh.gip2owner for s in sems]
    if not any(periodic):
        lit_owner = ref.find_gip_owner(ip2gip)
        for a, b in zip(owners, lit_owner):
            assert np.array_equal(a, b)
    lit = ref.setup_assembler(ip2gip, owners)
    for r in range(nranks):
        for name in ("send_i", "recv_idx", "recvback_idx"):
            mine = getattr(sems[r].asm, name)
            for peer in range(nranks):
                assert np.array_equal(mine[peer], lit[r][name][peer]), (r, name, peer)


def test_xy_partition_rule():
    """mesh.jl:1513-1533: nx = divisor of nparts closest to sqrt(nparts*lx/ly); part = xi*ny+yi+1."""
    cx, cy = np.meshgrid(np.arange(8) + 0.5, np.arange(8) + 0.5, indexing="ij")
    part, nx, ny = compute_xy_partition(cx.reshape(-1), cy.reshape(-1), 8)
    assert (nx, ny) == (2, 4)
    assert part.min() == 1 and part.max() == 8
    assert np.all(np.bincount(part)[1:] == 8)
    part, nx, ny = compute_xy_partition(cx.reshape(-1), cy.reshape(-1), 4)
    assert (nx, ny) == (2, 2)


@pytest.mark.parametrize("nranks", [2, 4])
def test_partitioned_mass_matches_single_rank(nranks):
    """DSS_global_mass!: the assembled diagonal mass of an nranks-way partition equals the one-rank
    mass at the same global node, up to summation order (a few ulp)."""
    spec = box3d((4, 4, 2), 3, warp=0.04)
    one = sem_setup(spec, 1)[0]
    many = sem_setup(spec, nranks)
    ref_M = np.zeros(one.mesh.gnpoin)
    ref_M[one.mesh.ip2gip - 1] = one.M
    for s in many:
        assert np.allclose(s.M, ref_M[s.mesh.ip2gip - 1], rtol=1e-14, atol=0)


def test_oracle_partitioned_rhs_matches_single_rank():
    """assemble_mpi! in the oracle: an R-rank evaluation gives every copy of a shared node the
    identical value, and equals the single-rank RHS to summation-order accuracy."""
    from helpers import MU3, PHYS, euler_case
    spec = box3d((4, 4, 2), 3, warp=0.04)
    out = {}
    for R in (1, 4):
        sems, qns, qes, us = euler_case(spec, R, lpert=False)
        probs = [ref.RefProblem(s, qe, eq_id=0, lpert=False, lsource=True, lvisc=True, visc_coeff=MU3, phys=PHYS)
                 for s, qe in zip(sems, qes)]
        caches = ref.setup_assembler([s.mesh.ip2gip for s in sems], [s.mesh.gip2owner for s in sems]) if R > 1 else None
        run = ref.RefRun(probs, caches)
        dus = [np.zeros_like(u) for u in us]
        run.rhs(dus, us, 0.0)
        g = np.full((sems[0].mesh.gnpoin, 5), np.nan)
        for s, du in zip(sems, dus):
            loc = du.reshape(s.mesh.npoin, 5, order="F")
            prev = g[s.mesh.ip2gip - 1]
            seen = ~np.isnan(prev[:, 0])
            assert np.array_equal(prev[seen], loc[seen]), "copies of a shared node differ"
            g[s.mesh.ip2gip - 1] = loc
        out[R] = g
    # the conditioned ICs of the two partitions already differ by ulps (different DSS orders), and the
    # derivative operators amplify that; 1e-10 of the field scale is summation-order accuracy here
    scale = np.abs(out[1]).max(axis=0)
    assert np.all(np.abs(out[4] - out[1]) <= 1e-10 * scale)


def test_interface_element_groups_cover_every_shared_node():
    """Host restatement of the interface-first split (JX_OPT_OVERLAP): no interior group may touch a node of the
    assembler lists, every group is in exactly one set, and a 2-rank split of a box has both kinds."""
    from helpers import box3d
    from jexpresso_b200.sem import sem_setup
    from jexpresso_b200.sem.partition import interface_element_groups
    for nranks, periodic in ((2, (False, False, False)), (1, (True, True, False)), (4, (True, False, False))):
        sems = sem_setup(box3d((16, 12, 3) if nranks > 1 else (8, 8, 3), 2, periodic=periodic), nranks)
        for s in sems:
            for epb in (1, 2, 3):
                gi, gn = interface_element_groups(s.mesh.connijk, s.asm, epb)
                ngroups = (s.mesh.nelem + epb - 1) // epb
                assert sorted(list(gi) + list(gn)) == list(range(ngroups))
                assert len(gi) > 0 and len(gn) > 0
                shared = set()
                for lists in (s.asm.send_i, s.asm.recv_idx, s.asm.recvback_idx):
                    for v in lists:
                        shared.update(int(x) for x in v)
                conn = s.mesh.connijk.reshape(s.mesh.nelem, -1)
                for g in gn:
                    els = range(g * epb, min((g + 1) * epb, s.mesh.nelem))
                    assert not (set(int(x) for e in els for x in conn[e]) & shared)
