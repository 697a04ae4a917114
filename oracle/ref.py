"""oracle/ref.py -- TEST INFRASTRUCTURE (python side of the CPU oracle).

ctypes driver for oracle/jexref.c plus literal python restatements of the pieces
of the reference that are integer logic or orchestration:

  * find_gip_owner          src/kernel/mesh/mesh.jl:3560-3610
  * CyclingReverseDict      src/kernel/mpi/mpi_communications.jl:1-46
  * setup_assembler         src/kernel/mpi/mpi_communications.jl:75-234
  * assemble_mpi!           src/kernel/mpi/mpi_communications.jl:260-338
  * rhs! / _build_rhs!      src/kernel/operators/rhs.jl:121-134, 498-711  (orchestration; loops in C)
  * stage drivers           OrdinaryDiffEq.jl (NOT in the reference tree; restated from the
                            published schemes: Carpenter & Kennedy 1994 2N-storage RK4(5),
                            Spiteri & Ruuth 2002 SSPRK(5,4), Shu & Osher 1988 SSPRK(3,3));
                            call site src/kernel/solvers/TimeIntegrators.jl:597-607, dt rounded
                            to Float32 at :464-465.  Per-stage parity with OrdinaryDiffEq is
                            UNPINNED except through the 2D theta golden end state.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libjexref.so")


def build(force=False):
    src = [os.path.join(_HERE, "jexref.c"), os.path.join(_HERE, "..", "include", "jxpow.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.jxo_work_doubles.restype = ctypes.c_size_t
        _lib.jxo_work_doubles.argtypes = [ctypes.c_void_p]
        _lib.jxo_build_rhs_local.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
        _lib.jxo_divide_by_mass_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.jxo_rhs_single.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
        for f in (_lib.jxo_assemble_add, _lib.jxo_assemble_pack, _lib.jxo_assemble_unpack):
            f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _lib.jxo_pow.restype = ctypes.c_double
        _lib.jxo_pow.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int]
        _lib.jxo_almost_equal.argtypes = [ctypes.c_double, ctypes.c_double]
        _lib.jxo_sizeof_problem.restype = ctypes.c_size_t
    return _lib


class _Problem(ctypes.Structure):
    _fields_ = [("nsd", ctypes.c_int32), ("ngl", ctypes.c_int32), ("neqs", ctypes.c_int32),
                ("eq_id", ctypes.c_int32), ("lpert", ctypes.c_int32), ("lsource", ctypes.c_int32),
                ("lvisc", ctypes.c_int32), ("pow_mode", ctypes.c_int32),
                ("nelem", ctypes.c_int64), ("npoin", ctypes.c_int64),
                ("phys", ctypes.c_double * 16),
                ("visc_coeff", ctypes.c_void_p), ("connijk", ctypes.c_void_p), ("coords", ctypes.c_void_p),
                ("met", ctypes.c_void_p * 10), ("dpsi", ctypes.c_void_p), ("omega", ctypes.c_void_p),
                ("Minv", ctypes.c_void_p), ("qe", ctypes.c_void_p),
                ("nfaces_bdy", ctypes.c_int64), ("poin_in_bdy_face", ctypes.c_void_p),
                ("nx", ctypes.c_void_p), ("ny", ctypes.c_void_p), ("nz", ctypes.c_void_p),
                ("face_kind", ctypes.c_void_p),
                ("xmin", ctypes.c_double), ("xmax", ctypes.c_double), ("ymin", ctypes.c_double),
                ("ymax", ctypes.c_double), ("zmin", ctypes.c_double), ("zmax", ctypes.c_double),
                ("visc_model", ctypes.c_int32), ("lrichardson", ctypes.c_int32), ("ltheta_eqn", ctypes.c_int32),
                ("sgs_pad", ctypes.c_int32), ("sgs", ctypes.c_double * 8), ("ad_lvl", ctypes.c_void_p),
                ("lbdy_fluxes", ctypes.c_int32), ("ifirst_wall_node", ctypes.c_int32),
                ("delta_hf", ctypes.c_double), ("user_heatflux", ctypes.c_double), ("most", ctypes.c_double * 4),
                ("bdy_face_in_elem", ctypes.c_void_p), ("Jef", ctypes.c_void_p), ("face_flux_kind", ctypes.c_void_p)]


def _f64(a):
    a = np.asarray(a, dtype=np.float64)
    return a if a.flags.f_contiguous else np.asfortranarray(a)


def _i64(a):
    a = np.asarray(a, dtype=np.int64)
    return a if a.flags.f_contiguous else np.asfortranarray(a)


def face_kinds(tags):
    """BCs.jl:621-623: faces tagged periodic* are skipped by the Dirichlet loop."""
    per = {"periodicx", "periodicy", "periodicz", "periodic1", "periodic2", "periodic3", "Laguerre"}
    return np.array([0 if t in per else 1 for t in tags], np.int32)


class RefProblem:
    """One rank's `params` as the oracle sees it.  Holds numpy arrays alive for the C struct."""

    def __init__(self, sem, qe, *, eq_id=0, lpert=False, lsource=True, lvisc=False, visc_coeff=None,
                 phys=None, pow_mode=0, neqs=None, sgs=None, bdy_fluxes=None):
        """sgs: None (AV) or a dict(model="SMAG"|"VREM", delta=mesh.Δeffective_l, lrichardson=True, ltheta_eqn=True,
        consts=[Pr_t, Sc_t, mu_mol, kappa_mol, Ri_crit, C_s], ad_lvl=None) -- the SGS struct of sgsStructs.jl.
        bdy_fluxes: None or a dict(Jef=[nfb,n,n], ifirst_wall_node_index=.., delta_hf=.., user_heatflux=.., karman=0.4,
        z0_m=0.1, z0_h=0.01) -- inputs[:bdy_fluxes] with the MOST wall model on the faces tagged "MOST" (BCs.jl:655-816)."""
        m = sem.mesh
        self.sem = sem
        self.neqs = neqs if neqs is not None else m.nsd + 2
        self.npoin = m.npoin
        keep = self._keep = {}
        keep["conn"] = _i64(m.connijk)
        keep["coords"] = _f64(m.coords)
        keep["met"] = [_f64(a) for a in sem.metric_list]
        keep["dpsi"] = _f64(sem.basis["dpsi"])
        keep["omega"] = _f64(sem.basis["omega"])
        keep["Minv"] = _f64(sem.Minv)
        keep["qe"] = _f64(qe)
        assert keep["qe"].shape == (m.npoin, self.neqs + 1)
        keep["pibf"] = _i64(m.poin_in_bdy_face)
        keep["nx"], keep["ny"] = _f64(sem.nx), _f64(sem.ny)
        keep["nz"] = _f64(sem.nz) if sem.nz is not None else None
        keep["kind"] = face_kinds(m.bdy_face_type)
        keep["visc"] = _f64(visc_coeff if visc_coeff is not None else np.zeros(self.neqs))
        P = self.P = _Problem()
        P.nsd, P.ngl, P.neqs = m.nsd, m.ngl, self.neqs
        P.eq_id, P.lpert, P.lsource, P.lvisc, P.pow_mode = eq_id, int(lpert), int(lsource), int(lvisc), pow_mode
        P.nelem, P.npoin = m.nelem, m.npoin
        ph = list(phys) if phys is not None else []
        for i in range(16):
            P.phys[i] = ph[i] if i < len(ph) else 0.0
        P.visc_coeff = keep["visc"].ctypes.data
        P.connijk = keep["conn"].ctypes.data
        P.coords = keep["coords"].ctypes.data
        for i, a in enumerate(keep["met"]):
            P.met[i] = a.ctypes.data
        P.dpsi, P.omega = keep["dpsi"].ctypes.data, keep["omega"].ctypes.data
        P.Minv, P.qe = keep["Minv"].ctypes.data, keep["qe"].ctypes.data
        P.nfaces_bdy = keep["pibf"].shape[0]
        P.poin_in_bdy_face = keep["pibf"].ctypes.data
        P.nx, P.ny = keep["nx"].ctypes.data, keep["ny"].ctypes.data
        P.nz = keep["nz"].ctypes.data if keep["nz"] is not None else None
        P.face_kind = keep["kind"].ctypes.data
        P.xmin, P.xmax, P.ymin, P.ymax, P.zmin, P.zmax = m.xmin, m.xmax, m.ymin, m.ymax, m.zmin, m.zmax
        P.visc_model = 0
        if sgs is not None:
            P.visc_model = {"SMAG": 1, "VREM": 2}[sgs["model"]]
            P.lrichardson, P.ltheta_eqn = int(sgs.get("lrichardson", True)), int(sgs.get("ltheta_eqn", True))
            cs = list(sgs["consts"])
            assert len(cs) == 6
            for i in range(6):
                P.sgs[i] = cs[i]
            P.sgs[6] = float(sgs["delta"])
            if sgs.get("ad_lvl") is not None:
                keep["ad_lvl"] = _i64(sgs["ad_lvl"])
                assert keep["ad_lvl"].shape == (m.nelem,)
                P.ad_lvl = keep["ad_lvl"].ctypes.data
        P.lbdy_fluxes = 0
        if bdy_fluxes is not None:
            assert m.nsd == 3
            P.lbdy_fluxes = 1
            P.ifirst_wall_node = int(bdy_fluxes["ifirst_wall_node_index"])
            assert 2 <= P.ifirst_wall_node <= m.ngl
            P.delta_hf, P.user_heatflux = float(bdy_fluxes.get("delta_hf", 0.0)), float(bdy_fluxes.get("user_heatflux", 0.0))
            P.most[0], P.most[1], P.most[2] = (float(bdy_fluxes.get("karman", 0.4)), float(bdy_fluxes.get("z0_m", 0.1)),
                                               float(bdy_fluxes.get("z0_h", 0.01)))
            keep["bfe"] = _i64(m.bdy_face_in_elem)
            keep["Jef"] = _f64(bdy_fluxes["Jef"])
            assert keep["Jef"].shape == keep["pibf"].shape
            keep["fkind"] = np.array([1 if t == "MOST" else 0 for t in m.bdy_face_type], np.int32)
            P.bdy_face_in_elem, P.Jef, P.face_flux_kind = keep["bfe"].ctypes.data, keep["Jef"].ctypes.data, keep["fkind"].ctypes.data
        assert lib().jxo_sizeof_problem() == ctypes.sizeof(_Problem)
        self.work = np.empty(lib().jxo_work_doubles(ctypes.byref(P)), np.float64)

    def build_rhs_local(self, u, RHS, t):
        lib().jxo_build_rhs_local(ctypes.byref(self.P), u.ctypes.data, RHS.ctypes.data, float(t), self.work.ctypes.data)

    def sgs_mu_turb(self):
        """sgs.μ_turb after the last evaluation (every shared node holds the value of the last element that wrote it)."""
        L = lib()
        L.jxo_sgs_mu_turb.restype = ctypes.c_void_p
        L.jxo_sgs_mu_turb.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        p = L.jxo_sgs_mu_turb(ctypes.byref(self.P), self.work.ctypes.data)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(self.npoin,)).copy()

    def divide_by_mass(self, RHS):
        lib().jxo_divide_by_mass_matrix(ctypes.byref(self.P), RHS.ctypes.data)


# --------------------------------------------------------------------------------------------
# integer restatements (literal, dictionary based, python loops: small cases only)
# --------------------------------------------------------------------------------------------
def find_gip_owner(all_a):
    """mesh.jl:3560-3610, rank-0 view: all_a[r] = ip2gip of rank r.  Returns owners per rank."""
    nranks = len(all_a)
    element_owner_map = {}
    ownership_counts = [0] * nranks
    for r in range(nranks):
        for el in all_a[r]:
            el = int(el)
            owner = r
            if el not in element_owner_map:
                element_owner_map[el] = owner
            else:
                current_owner = element_owner_map[el]
                if ownership_counts[owner] < ownership_counts[current_owner]:
                    element_owner_map[el] = owner
                ownership_counts[element_owner_map[el]] += 1
    return [np.array([element_owner_map[int(el)] for el in a], np.int64) for a in all_a]


class CyclingReverseDict:
    """mpi_communications.jl:1-46"""

    def __init__(self, a):
        self.mapping, self.counters, self.repeated_keys = {}, {}, []
        for i, val in enumerate(a, start=1):
            val = int(val)
            if val in self.mapping:
                if len(self.mapping[val]) == 1:
                    self.repeated_keys.append(val)
                self.mapping[val].append(i)
            else:
                self.mapping[val] = [i]

    def get_vals(self, key, all=False, first=False):
        indices = self.mapping[key]
        if all:
            return indices
        if first:
            return indices[0]
        counter = self.counters.get(key, 0)
        result = indices[counter % len(indices)]
        self.counters[key] = counter + 1
        return result


def setup_assembler(index_a_all, owner_a_all):
    """mpi_communications.jl:75-234 executed for every rank, with the Alltoall/Isend/Irecv of
    global-id lists replaced by direct list hand-over.  Returns per rank a dict with
    send_i, recv_idx_buffers, recvback_idx_buffers (lists indexed by peer rank, 1-based ids)."""
    R = len(index_a_all)
    send_idx_all, send_i_all, g2l_all = [], [], []
    for rank in range(R):
        index_a, owner_a = index_a_all[rank], owner_a_all[rank]
        send_idx = {i: [] for i in range(R)}
        send_i = [[] for _ in range(R)]
        for i, idx in enumerate(index_a, start=1):
            owner = int(owner_a[i - 1])
            if owner != rank:
                send_idx[owner].append(int(idx))
                send_i[owner].append(i)
        a_g2l_idx = CyclingReverseDict(index_a)
        for idx in a_g2l_idx.repeated_keys:
            local_idx = a_g2l_idx.get_vals(idx, all=True)[1:]
            for i in local_idx:
                owner = int(owner_a[i - 1])
                if owner == rank:
                    send_idx[owner].append(idx)
                    send_i[owner].append(i)
        send_idx_all.append(send_idx); send_i_all.append(send_i); g2l_all.append(a_g2l_idx)
    out = []
    for rank in range(R):
        a_g2l_idx = g2l_all[rank]
        recv_idx_buffers = [list(send_idx_all[src][rank]) for src in range(R)]          # tag 1 messages
        recvback_idx_buffers = [list(send_idx_all[rank][dst]) for dst in range(R)]      # tag 3 echo
        for rk in range(R):
            recv_idx_buffers[rk] = [a_g2l_idx.get_vals(idx, first=True) for idx in recv_idx_buffers[rk]]
            if rk == rank:
                recvback_idx_buffers[rk] = [send_i_all[rank][rk][i] for i in range(len(recvback_idx_buffers[rk]))]
            else:
                recvback_idx_buffers[rk] = [a_g2l_idx.get_vals(idx) for idx in recvback_idx_buffers[rk]]
        out.append({"send_i": [np.array(v, np.int64) for v in send_i_all[rank]],
                    "recv_idx": [np.array(v, np.int64) for v in recv_idx_buffers],
                    "recvback_idx": [np.array(v, np.int64) for v in recvback_idx_buffers]})
    return out


def assemble_mpi(arrays, caches):
    """mpi_communications.jl:260-338 for all ranks: pack -> owner adds in ascending sender rank ->
    pack sums -> send back -> overwrite.  arrays[r] is [npoin, m] Fortran ordered."""
    R = len(arrays)
    L = lib()
    m = arrays[0].shape[1] if arrays[0].ndim == 2 else 1
    send = {}
    for src in range(R):
        for dst in range(R):
            idx = caches[src]["send_i"][dst]
            if len(idx):
                buf = np.empty(len(idx) * m)
                L.jxo_assemble_pack(arrays[src].ctypes.data, arrays[src].shape[0], m, idx.ctypes.data, len(idx), buf.ctypes.data)
                send[(src, dst)] = buf
    for dst in range(R):
        for src in range(R):
            idx = caches[dst]["recv_idx"][src]
            if len(idx):
                L.jxo_assemble_add(arrays[dst].ctypes.data, arrays[dst].shape[0], m, idx.ctypes.data, len(idx),
                                   send[(src, dst)].ctypes.data)
    back = {}
    for dst in range(R):
        for src in range(R):
            idx = caches[dst]["recv_idx"][src]
            if len(idx):
                buf = np.empty(len(idx) * m)
                L.jxo_assemble_pack(arrays[dst].ctypes.data, arrays[dst].shape[0], m, idx.ctypes.data, len(idx), buf.ctypes.data)
                back[(dst, src)] = buf
    for src in range(R):
        for dst in range(R):
            idx = caches[src]["recvback_idx"][dst]
            if len(idx):
                L.jxo_assemble_unpack(arrays[src].ctypes.data, arrays[src].shape[0], m, idx.ctypes.data, len(idx),
                                      back[(dst, src)].ctypes.data)


# --------------------------------------------------------------------------------------------
# rhs! for R simulated ranks
# --------------------------------------------------------------------------------------------
class RefRun:
    """R ranks of the reference CPU path in one process."""

    def __init__(self, problems, caches=None):
        self.problems = problems
        self.caches = caches
        self.RHS = [np.zeros((p.npoin, p.neqs), order="F") for p in problems]

    def rhs(self, dus, us, t):
        """rhs!(du,u,params,t) on every rank (rhs.jl:121-134).  us[r] is mutated by the BC."""
        for p, u, R_ in zip(self.problems, us, self.RHS):
            p.build_rhs_local(u, R_, t)
        if self.caches is not None:
            assemble_mpi(self.RHS, self.caches)          # DSS_global_RHS!  rhs.jl:690
        for p, R_, du in zip(self.problems, self.RHS, dus):
            p.divide_by_mass(R_)                         # rhs.jl:698-699
            du[:] = R_.reshape(-1, order="F")            # RHStoDU!  rhs.jl:14-27


CK2N54 = {
    # Carpenter & Kennedy (1994), NASA TM 109112, 2N-storage RK4(5)
    "A": [0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
          -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0],
    "B": [1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
          1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
          2277821191437.0 / 14882151754819.0],
    "c": [0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
          2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0],
}

SSPRK54 = dict(b10=0.391752226571890, a20=0.444370493651235, a21=0.555629506348765, b21=0.368410593050371,
               a30=0.620101851488403, a32=0.379898148511597, b32=0.251891774271694,
               a40=0.178079954393132, a43=0.821920045606868, b43=0.544974750228521,
               a52=0.517231671970585, a53=0.096059710526147, b53=0.063692468666290,
               a54=0.386708617503269, b54=0.226007483236906,
               c1=0.391752226571890, c2=0.586079689311540, c3=0.474542363121400, c4=0.935010630967653)


def float32_dt(dt):
    """TimeIntegrators.jl:464-465: dt = Float32(Δt / 2^ad_lvl_max), widened back to Float64."""
    return float(np.float32(dt))


def step_ck2n54(run, us, t, dt, tmps, ks):
    """One CarpenterKennedy2N54 step, 2N low-storage form: tmp = A_i tmp + dt k; u += B_i tmp."""
    A, B, c = CK2N54["A"], CK2N54["B"], CK2N54["c"]
    for i in range(5):
        run.rhs(ks, us, t + c[i] * dt)
        for u, tmp, k in zip(us, tmps, ks):
            if i == 0:
                tmp[:] = dt * k
            else:
                tmp[:] = A[i] * tmp + dt * k
            u[:] = u + B[i] * tmp


def step_ssprk33(run, us, t, dt, ks):
    """Shu-Osher SSPRK(3,3) in OrdinaryDiffEq's update form.  The first evaluation is the FSAL
    value f(uprev) (computed at the end of the previous step in the integrator; it applies the
    Dirichlet projection to uprev in place, exactly as calling it here does)."""
    run.rhs(ks, us, t)
    uprev = [u.copy() for u in us]
    for u, k in zip(us, ks):
        u[:] = u + dt * k
    run.rhs(ks, us, t + dt)
    for u, up, k in zip(us, uprev, ks):
        u[:] = (3 * up + u + dt * k) / 4
    run.rhs(ks, us, t + dt / 2)
    for u, up, k in zip(us, uprev, ks):
        u[:] = (up + 2 * u + 2 * dt * k) / 3


def step_ssprk54(run, us, t, dt, ks):
    C = SSPRK54
    run.rhs(ks, us, t)
    uprev = [u.copy() for u in us]
    u2 = [up + C["b10"] * dt * k for up, k in zip(uprev, ks)]
    run.rhs(ks, u2, t + C["c1"] * dt)
    u2 = [C["a20"] * up + C["a21"] * a + C["b21"] * dt * k for up, a, k in zip(uprev, u2, ks)]
    run.rhs(ks, u2, t + C["c2"] * dt)
    u3 = [C["a30"] * up + C["a32"] * a + C["b32"] * dt * k for up, a, k in zip(uprev, u2, ks)]
    k3 = [np.empty_like(k) for k in ks]
    run.rhs(k3, u3, t + C["c3"] * dt)
    u4 = [C["a40"] * up + C["a43"] * a + C["b43"] * dt * k for up, a, k in zip(uprev, u3, k3)]
    run.rhs(ks, u4, t + C["c4"] * dt)
    for u, a2, a3, a4, kk3, k in zip(us, u2, u3, u4, k3, ks):
        u[:] = C["a52"] * a2 + C["a53"] * a3 + C["b53"] * dt * kk3 + C["a54"] * a4 + C["b54"] * dt * k


def time_loop(run, us, t0, dt, nsteps, scheme="CK2N54"):
    """solve(...; dt, adaptive=false) restated: nsteps fixed steps; returns final time."""
    dt = float32_dt(dt)
    ks = [np.zeros_like(u) for u in us]
    tmps = [np.zeros_like(u) for u in us]
    t = t0
    for _ in range(nsteps):
        if scheme == "CK2N54":
            step_ck2n54(run, us, t, dt, tmps, ks)
        elif scheme == "SSPRK33":
            step_ssprk33(run, us, t, dt, ks)
        elif scheme == "SSPRK54":
            step_ssprk54(run, us, t, dt, ks)
        else:
            raise ValueError(scheme)
        t = t + dt
    if scheme in ("SSPRK33", "SSPRK54"):
        run.rhs(ks, us, t)        # trailing FSAL evaluation: applies the BC projection to the final state
    return t
