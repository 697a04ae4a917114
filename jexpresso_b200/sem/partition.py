"""Node ownership, periodic merging and the interface-assembly index lists.

Integer (bit-exact) host logic that produces what ``jx_upload_halo`` consumes.
Restated from
  * src/kernel/mesh/mesh.jl:3560-3610            find_gip_owner
  * src/kernel/mesh/restructure_for_periodicity.jl:1387-1577
                                                  restructure4periodicity_3D_sorted!
  * src/kernel/mpi/mpi_communications.jl:1-46     CyclingReverseDict
  * src/kernel/mpi/mpi_communications.jl:75-234   setup_assembler  (AssemblerCache)

The reference runs these with MPI collectives rooted at rank 0.  Here every
function takes the per-rank arrays of *all* ranks (what rank 0 sees after the
Gatherv) and returns the per-rank results (what Scatterv hands back), so the
same code serves a single process simulating R ranks (tests, oracle checks) and
the torch.distributed wrapper in :mod:`jexpresso_b200.distributed`, which
performs the gather/scatter with ``all_gather_object``.

All node ids are 1-based, ranks 0-based, exactly as in the reference.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = ["interface_element_groups", "find_gip_owner_all", "restructure4periodicity_all", "AssemblerLists", "setup_assembler_all"]


def find_gip_owner_all(ip2gip_list):
    """Owner rank of every local node of every rank (mesh.jl:3560-3610).

    First rank to list a global id owns it, unless a later rank currently
    "owns fewer contested ids"; the contest counter is bumped for the winner on
    every repeated sighting (including repeats inside one rank).
    """
    nranks = len(ip2gip_list)
    sizes = [len(a) for a in ip2gip_list]
    flat = np.concatenate([np.asarray(a, np.int64) for a in ip2gip_list]) if nranks else np.zeros(0, np.int64)
    rank_of = np.repeat(np.arange(nranks, dtype=np.int64), sizes)
    # ids seen exactly once keep their first-seen owner and never touch the counters:
    uniq, first_idx, counts = np.unique(flat, return_index=True, return_counts=True)
    owner_of = dict()
    contested = set(uniq[counts > 1].tolist())
    owners_flat = rank_of.copy()
    if contested:
        sel = np.nonzero(np.isin(flat, uniq[counts > 1]))[0]          # flat order preserved
        ownership_counts = [0] * nranks
        for i in sel.tolist():
            el = int(flat[i])
            r = int(rank_of[i])
            cur = owner_of.get(el)
            if cur is None:
                owner_of[el] = r
            else:
                if ownership_counts[r] < ownership_counts[cur]:
                    owner_of[el] = r
                ownership_counts[owner_of[el]] += 1
        owners_flat[sel] = np.array([owner_of[int(flat[i])] for i in sel.tolist()], np.int64)
    out, off = [], 0
    for s in sizes:
        out.append(owners_flat[off:off + s].copy())
        off += s
    return out


def _r5(a):
    return np.round(a, 5)


def restructure4periodicity_all(meshes, direction):
    """Merge the global ids of periodic twins in ``direction`` ("periodicx|y|z").

    Rule (restructure_for_periodicity.jl:1460-1560): boundary nodes on faces tagged
    ``direction`` are gathered, deduplicated by rounded coordinates, and every
    min-side node is paired with the max-side node having the same transverse
    coordinates; the pair member selected by the reference's comp1||comp2||comp3
    test is the master, the other adopts its global id and owner.  ``ip2gip`` and
    ``gip2owner`` of all meshes are updated in place.
    """
    per_ip, xs, ys, zs, gips, owners = [], [], [], [], [], []
    for m in meshes:
        tags = np.array(m.bdy_face_type) if len(m.bdy_face_type) else np.zeros(0, dtype=str)
        sel = np.nonzero(tags == direction)[0]
        ip = np.unique(m.poin_in_bdy_face[sel].reshape(-1)) - 1 if sel.size else np.zeros(0, np.int64)
        per_ip.append(ip)
        xs.append(m.x[ip]); ys.append(m.y[ip]); zs.append(m.z[ip])
        gips.append(m.ip2gip[ip]); owners.append(m.gip2owner[ip])
    x = np.concatenate(xs); y = np.concatenate(ys); z = np.concatenate(zs)
    g = np.concatenate(gips); ow = np.concatenate(owners)
    if x.size == 0:
        return
    key = np.stack([_r5(x), _r5(y), _r5(z)], axis=1)
    _, uidx = np.unique(key, axis=0, return_index=True)
    uidx.sort()
    ux, uy, uz, ug, uo = x[uidx], y[uidx], z[uidx], g[uidx], ow[uidx]
    ax, t1, t2 = {"periodicx": (ux, uy, uz), "periodicy": (uy, ux, uz), "periodicz": (uz, ux, uy)}[direction]
    axmin, axmax = ax.min(), ax.max()
    on_max = np.nonzero(np.abs(ax - axmax) < 1e-4)[0]
    on_min = np.nonzero(np.abs(ax - axmin) < 1e-4)[0]
    lookup = {(float(a), float(b)): int(j) for a, b, j in zip(_r5(t1[on_max]), _r5(t2[on_max]), on_max)}
    changes_ip, changes_owner = {}, {}

    for i in on_min.tolist():
        j = lookup.get((float(_r5(t1[i])), float(_r5(t2[i]))))
        if j is None:
            continue
        gi, gj = int(ug[i]), int(ug[j])
        if changes_ip.get(gi, gi) == gj or changes_ip.get(gj, gj) == gi:
            continue
        xi_, yi_, zi_ = float(ux[i]), float(uy[i]), float(uz[i])
        xt, yt, zt = float(ux[j]), float(uy[j]), float(uz[j])
        # colinearity of (p_i - p_j) with the axis normal holds by construction for box twins
        if yi_ == 0 and yt == 0 and zi_ == 0 and zt == 0:
            comp1 = xi_ < xt
        elif yi_ == 0 and yt == 0:
            comp1 = xi_ * abs(zi_) < xt * abs(zt)
        elif zi_ == 0 and zt == 0:
            comp1 = xi_ * abs(zi_) < xt * abs(yt)
        else:
            comp1 = xi_ * abs(yi_ * zi_) < xt * abs(yt * zt)
        if xi_ == 0 and xt == 0 and zi_ == 0 and zt == 0:
            comp2 = yi_ < yt
        elif xi_ == 0 and xt == 0:
            comp2 = yi_ * abs(zi_) < yt * abs(zt)
        elif zi_ == 0 and zt == 0:
            comp2 = yi_ * abs(xi_) < yt * abs(xt)
        else:
            comp2 = yi_ * abs(xi_ * zi_) < yt * abs(xt * zt)
        if xi_ == 0 and xt == 0 and yi_ == 0 and yt == 0:
            comp3 = zi_ < zt
        elif xi_ == 0 and xt == 0:
            comp3 = zi_ * abs(yi_) < zt * abs(yt)
        elif yi_ == 0 and yt == 0:
            comp3 = zi_ * abs(xi_) < zt * abs(xt)
        else:
            comp3 = zi_ * abs(xi_ * yi_) < zt * abs(xt * yt)
        if comp1 or comp2 or comp3:
            changes_ip[gj] = gi
            if uo[j] != uo[i]:
                changes_owner[gj] = int(uo[i]); changes_owner[gi] = int(uo[i])
        else:
            changes_ip[gi] = gj
            if uo[j] != uo[i]:
                changes_owner[gi] = int(uo[j]); changes_owner[gj] = int(uo[j])
    for m, ip in zip(meshes, per_ip):
        old = m.ip2gip[ip]
        m.gip2owner[ip] = np.array([changes_owner.get(int(v), int(o)) for v, o in zip(old, m.gip2owner[ip])], np.int64)
        m.ip2gip[ip] = np.array([changes_ip.get(int(v), int(v)) for v in old], np.int64)


@dataclass
class AssemblerLists:
    """The index content of the reference's AssemblerCache for one rank
    (mpi_communications.jl:48-73).  Lists are indexed by peer rank; entries are
    1-based local node ids."""
    rank: int
    nranks: int
    send_i: list            # send_i[r]       local ids whose values go to owner r
    recv_idx: list          # recv_idx[r]     local ids (owner side) the values from r add into
    recvback_idx: list      # recvback_idx[r] local ids overwritten by the sum sent back by r
    send_gid: list          # global ids travelling with send_i (setup-time message)

    @property
    def active_send_ranks(self):
        return [r for r in range(self.nranks) if len(self.send_i[r]) > 0]

    @property
    def active_recv_ranks(self):
        return [r for r in range(self.nranks) if len(self.recv_idx[r]) > 0]

    def is_trivial(self):
        return not self.active_send_ranks and not self.active_recv_ranks


def interface_element_groups(connijk, asm, epb):
    """Host restatement of libjexrhs' interface-first split (jexrhs.cu: ensure_split): ids of the groups of ``epb``
    consecutive elements that name a node of the assembler lists, and of the remaining (interior) groups.  Used by the
    tests to check ``jx_split_info``; the library builds the same lists on the device from its element records."""
    nelem = connijk.shape[0]
    npoin = int(connijk.max())
    mask = np.zeros(npoin + 1, bool)
    for lists in (asm.send_i, asm.recv_idx, asm.recvback_idx):
        for v in lists:
            mask[np.asarray(v, np.int64)] = True
    touch = mask[connijk.reshape(nelem, -1)].any(axis=1)
    ngroups = (nelem + epb - 1) // epb
    pad = np.zeros(ngroups * epb, bool)
    pad[:nelem] = touch
    if nelem % epb and mask[1]:
        pad[nelem:] = True            # empty slots of the last record hold node id 1 (0-based 0)
    flag = pad.reshape(ngroups, epb).any(axis=1)
    return np.nonzero(flag)[0], np.nonzero(~flag)[0]


class _CyclingReverseDict:
    """mpi_communications.jl:1-46: global id -> list of local indices, with a
    per-key round-robin cursor for repeated (periodic) ids."""

    def __init__(self, a):
        a = np.asarray(a, np.int64)
        order = np.argsort(a, kind="stable")
        sa = a[order]
        starts = np.nonzero(np.r_[True, sa[1:] != sa[:-1]])[0]
        ends = np.r_[starts[1:], len(sa)]
        self._first = dict(zip(sa[starts].tolist(), (order[starts] + 1).tolist()))
        rep = np.nonzero(ends - starts > 1)[0]
        self._multi = {int(sa[starts[k]]): (order[starts[k]:ends[k]] + 1).tolist() for k in rep.tolist()}
        # repeated_keys in order of the *second* sighting, as the reference pushes them
        second = {k: v[1] for k, v in self._multi.items()}
        self.repeated_keys = sorted(second, key=second.get)
        self._counters = {}

    def all(self, key):
        return self._multi.get(key) or [self._first[key]]

    def first(self, key):
        return self._first[key]

    def next(self, key):
        v = self._multi.get(key)
        if v is None:
            return self._first[key]
        c = self._counters.get(key, 0)
        self._counters[key] = c + 1
        return v[c % len(v)]


def setup_assembler_all(ip2gip_list, owner_list, only_rank=None):
    """Build the AssemblerCache index lists of every rank (mpi_communications.jl:75-234).

    The Alltoall / Isend / Irecv of the reference move ``send_idx`` (global ids) to the
    owners and echo them back; with all ranks in hand that is plain list passing.
    """
    R = len(ip2gip_list)
    g2l, send_idx, send_i = [], [], []
    for rank in range(R):
        index_a = np.asarray(ip2gip_list[rank], np.int64)
        owner_a = np.asarray(owner_list[rank], np.int64)
        sidx = [[] for _ in range(R)]
        si = [[] for _ in range(R)]
        remote = np.nonzero(owner_a != rank)[0]                      # ascending local index
        for r in np.unique(owner_a[remote]).tolist():
            loc = remote[owner_a[remote] == r]
            sidx[r] = index_a[loc].tolist()
            si[r] = (loc + 1).tolist()
        crd = _CyclingReverseDict(index_a)
        for gid in crd.repeated_keys:                                # local periodic twins
            for i in crd.all(gid)[1:]:
                if owner_a[i - 1] == rank:
                    sidx[rank].append(int(gid))
                    si[rank].append(int(i))
        g2l.append(crd); send_idx.append(sidx); send_i.append(si)
    out = []
    for rank in (range(R) if only_rank is None else [only_rank]):
        crd = g2l[rank]
        recv_idx = [[crd.first(g) for g in send_idx[src][rank]] for src in range(R)]
        recvback = []
        for rk in range(R):
            echoed = send_idx[rank][rk]          # owner rk echoes the ids it received from us
            if rk == rank:
                recvback.append(list(send_i[rank][rk]))
            else:
                recvback.append([crd.next(g) for g in echoed])
        out.append(AssemblerLists(rank=rank, nranks=R,
                                  send_i=[np.array(v, np.int64) for v in send_i[rank]],
                                  recv_idx=[np.array(v, np.int64) for v in recv_idx],
                                  recvback_idx=[np.array(v, np.int64) for v in recvback],
                                  send_gid=[np.array(v, np.int64) for v in send_idx[rank]]))
    return out
