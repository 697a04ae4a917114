"""ctypes binding of libjexrhs (include/jexrhs.h) -- the same C ABI a Julia host reaches with
``ccall`` (julia/rhs_b200.jl).

There is no fallback: if the shared library is missing or no CUDA device is present the calls
raise :class:`JexError`.
"""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JX_LIB") or os.path.join(_HERE, "lib", "libjexrhs.so")   # JX_LIB: alternative build (kernel experiments)
HEADER_PATH = os.path.join(_HERE, "..", "include", "jexrhs.h")

JX_OK, JX_EINVAL, JX_ENODEV, JX_ECUDA, JX_ESTATE, JX_ENCCL, JX_ENOMEM = 0, -1, -2, -3, -4, -5, -6
JX_OPT_DSS_MODE, JX_OPT_POW_MODE, JX_OPT_ELEM_KERNEL, JX_OPT_CUDA_GRAPH, JX_OPT_OVERLAP = 1, 2, 3, 4, 5
JX_ELEM_AUTO, JX_ELEM_GENERIC = 0, -1
JX_VISC_AV, JX_VISC_SMAG, JX_VISC_VREM = 0, 1, 2
JX_FLUX_NONE, JX_FLUX_MOST = 0, 1
_ERRNAMES = {-1: "JX_EINVAL", -2: "JX_ENODEV", -3: "JX_ECUDA", -4: "JX_ESTATE", -5: "JX_ENCCL", -6: "JX_ENOMEM"}


class JexError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}" if msg else f"{_ERRNAMES.get(code, code)}")


def declared_symbols(header=HEADER_PATH):
    """Names of every function the public header declares (used by the CPU export test)."""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jx_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JexError(JX_ENODEV, f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    L.jx_init.argtypes = [i32, i32, i32, vp, ctypes.POINTER(vp)]
    L.jx_init_ex.argtypes = [i32, i32, i32, vp, i32, ctypes.POINTER(vp)]
    L.jx_rhs_dev.argtypes = [vp, dbl, vp, vp]
    L.jx_kernel_variant.argtypes = [vp]
    L.jx_nccl_unique_id.argtypes = [vp]
    L.jx_destroy.argtypes = [vp]
    L.jx_destroy.restype = None
    L.jx_last_error.argtypes = [vp, ctypes.c_char_p, i32]
    L.jx_set_option.argtypes = [vp, i32, i64]
    L.jx_set_problem.argtypes = [vp, i32, i32, i32, i64, i64, i32, i32, i32, i32, vp, vp, i32]
    L.jx_set_sgs.argtypes = [vp, i32, dbl, i32, i32, vp, i32, vp]
    L.jx_upload_mesh.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp]
    L.jx_upload_mesh_coords.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.jx_upload_bcs.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.jx_get_minv.argtypes = [vp, vp]
    L.jx_condition_state.argtypes = [vp, ctypes.c_int]
    L.jx_upload_bdy_fluxes.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, dbl, dbl, vp, i32]
    L.jx_upload_halo.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.jx_set_state.argtypes = [vp, vp]
    L.jx_get_state.argtypes = [vp, vp]
    L.jx_get_du.argtypes = [vp, vp]
    L.jx_rhs.argtypes = [vp, dbl, vp, vp, vp]
    L.jx_step.argtypes = [vp, i32, dbl, dbl, i32]
    L.jx_last_elapsed_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.jx_launch_count.argtypes = [vp]
    L.jx_launch_count.restype = i64
    L.jx_sync.argtypes = [vp]
    L.jx_split_info.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.jx_bench_rhs.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    L.jx_selftest.argtypes = [vp, i32, i64, ctypes.POINTER(i64)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def f64(a):
    """Fortran-ordered float64 view/copy (the memory image Julia holds)."""
    a = np.asarray(a, dtype=np.float64)
    return a if a.flags.f_contiguous else np.asfortranarray(a)


def i64(a):
    a = np.asarray(a, dtype=np.int64)
    return a if a.flags.f_contiguous else np.asfortranarray(a)


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    rc = lib().jx_nccl_unique_id(buf)
    if rc:
        raise JexError(rc, "ncclGetUniqueId failed / NCCL not loadable")
    return buf.raw


class Context:
    """Thin owner of one ``jx_ctx`` (one per rank / GPU)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_uid: bytes | None = None, nccl_max_ctas=0):
        self._h = ctypes.c_void_p()
        uid = ctypes.create_string_buffer(nccl_uid, 128) if nccl_uid is not None else None
        rc = lib().jx_init_ex(device, rank, nranks, uid, int(nccl_max_ctas), ctypes.byref(self._h))
        if rc:
            self._h = ctypes.c_void_p()
            raise JexError(rc, "jx_init failed (no CUDA device / bad arguments / NCCL)")
        self.rank, self.nranks = rank, nranks

    def _ck(self, rc):
        if rc:
            buf = ctypes.create_string_buffer(512)
            lib().jx_last_error(self._h, buf, 512)
            raise JexError(rc, buf.value.decode(errors="replace"))

    def close(self):
        if self._h:
            lib().jx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ---------------------------------------------------------------------
    def set_option(self, key, value):
        self._ck(lib().jx_set_option(self._h, key, int(value)))

    def set_problem(self, nsd, ngl, neqs, nelem, npoin, eq_id, lpert, lsource, lvisc, visc_coeff, phys):
        v = f64(visc_coeff if visc_coeff is not None else np.zeros(neqs))
        p = f64(phys if phys is not None else np.zeros(0))
        self.neqs, self.npoin = neqs, npoin
        self._ck(lib().jx_set_problem(self._h, nsd, ngl, neqs, nelem, npoin, eq_id, int(lpert), int(lsource), int(lvisc),
                                      _ptr(v), _ptr(p), len(p)))

    def set_sgs(self, visc_model, delta_effective, lrichardson=True, ltheta_eqn=True, consts=None, ad_lvl=None):
        """allocate_SGS + the flags of params_setup.jl:249-253 (jx_set_sgs); after set_problem(lvisc=True), before upload_mesh."""
        cs = f64(consts if consts is not None else np.zeros(0))
        lv = i64(ad_lvl) if ad_lvl is not None else None
        self._ck(lib().jx_set_sgs(self._h, int(visc_model), float(delta_effective), int(bool(lrichardson)), int(bool(ltheta_eqn)),
                                  _ptr(cs), len(cs), _ptr(lv)))

    def upload_mesh(self, connijk, coords, metrics, dpsi, omega, Minv, qe):
        mets = [f64(m) for m in metrics]
        arr = (ctypes.c_void_p * len(mets))(*[m.ctypes.data for m in mets])
        c, x, d, o, mi = i64(connijk), f64(coords), f64(dpsi), f64(omega), f64(Minv)
        q = f64(qe) if qe is not None else None
        self._ck(lib().jx_upload_mesh(self._h, _ptr(c), _ptr(x), arr, len(mets), _ptr(d), _ptr(o), _ptr(mi), _ptr(q)))

    def upload_mesh_coords(self, connijk, coords, dpsi, omega, Minv, qe):
        """jx_upload_mesh with the metric terms built on the device from the coordinates."""
        c, x, d, o = i64(connijk), f64(coords), f64(dpsi), f64(omega)
        mi = f64(Minv) if Minv is not None else None        # None: the mass matrix is built on the device too
        q = f64(qe) if qe is not None else None
        self._ck(lib().jx_upload_mesh_coords(self._h, _ptr(c), _ptr(x), _ptr(d), _ptr(o), _ptr(mi), _ptr(q)))

    def get_minv(self):
        mi = np.empty(self.npoin)
        self._ck(lib().jx_get_minv(self._h, _ptr(mi)))
        return mi

    def condition_state(self, which=0):
        """conformity4ncf_q! on the resident state (0) or the reference state (1); needs upload_mesh_coords(Minv=None)."""
        self._ck(lib().jx_condition_state(self._h, int(which)))

    def upload_bcs(self, poin_in_bdy_face, nx, ny, nz, kinds):
        p = i64(poin_in_bdy_face)
        nf = p.shape[0]
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        a, b = f64(nx), f64(ny)
        cz = f64(nz) if nz is not None else None
        self._ck(lib().jx_upload_bcs(self._h, nf, _ptr(p), _ptr(a), _ptr(b), _ptr(cz), _ptr(k)))

    def upload_bdy_fluxes(self, poin_in_bdy_face, bdy_face_in_elem, connijk, nx, ny, nz, Jef, omega, flux_kinds,
                          ifirst_wall_node_index, delta_hf=0.0, user_heatflux=0.0, most_consts=None):
        """inputs[:bdy_fluxes]: the MOST wall model on the faces with flux kind JX_FLUX_MOST (jx_upload_bdy_fluxes)."""
        p, e, cn = i64(poin_in_bdy_face), i64(bdy_face_in_elem), i64(connijk)
        a, b, cz, je, om = f64(nx), f64(ny), f64(nz), f64(Jef), f64(omega)
        k = np.ascontiguousarray(flux_kinds, dtype=np.int32)
        mc = f64(most_consts) if most_consts is not None else None
        self._ck(lib().jx_upload_bdy_fluxes(self._h, p.shape[0], _ptr(p), _ptr(e), _ptr(cn), _ptr(a), _ptr(b), _ptr(cz), _ptr(je),
                                            _ptr(om), _ptr(k), int(ifirst_wall_node_index), float(delta_hf), float(user_heatflux),
                                            _ptr(mc), 0 if mc is None else len(mc)))

    def upload_halo(self, send_i, recv_idx, recvback_idx):
        """Lists indexed by peer rank (1-based local ids), i.e. the AssemblerCache content."""
        def csr(lists):
            ptr = np.zeros(len(lists) + 1, np.int64)
            ptr[1:] = np.cumsum([len(v) for v in lists])
            flat = np.concatenate([np.asarray(v, np.int64) for v in lists]) if ptr[-1] else np.zeros(0, np.int64)
            return ptr, np.ascontiguousarray(flat)
        sp, sv = csr(send_i)
        rp, rv = csr(recv_idx)
        bp, bv = csr(recvback_idx)
        self._ck(lib().jx_upload_halo(self._h, _ptr(sp), _ptr(sv), _ptr(rp), _ptr(rv), _ptr(bp), _ptr(bv)))

    # -- state -------------------------------------------------------------------------------
    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == self.npoin * self.neqs
        self._ck(lib().jx_set_state(self._h, _ptr(u)))

    def get_state(self):
        u = np.empty(self.npoin * self.neqs)
        self._ck(lib().jx_get_state(self._h, _ptr(u)))
        return u

    def get_state_into(self, u):
        assert u.dtype == np.float64 and u.flags.c_contiguous and u.size == self.npoin * self.neqs
        self._ck(lib().jx_get_state(self._h, _ptr(u)))

    def kernel_variant(self):
        v = lib().jx_kernel_variant(self._h)
        if v < 0:
            raise JexError(v, "jx_kernel_variant before jx_set_problem")
        return v

    def rhs_dev(self, t, u_ptr, du_ptr):
        """rhs!(du,u,params,t) on device arrays the caller owns (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        self._ck(lib().jx_rhs_dev(self._h, float(t), ctypes.c_void_p(int(u_ptr)), ctypes.c_void_p(int(du_ptr))))

    def get_du(self):
        du = np.empty(self.npoin * self.neqs)
        self._ck(lib().jx_get_du(self._h, _ptr(du)))
        return du

    # -- hot path ----------------------------------------------------------------------------
    def rhs(self, t=0.0, u=None, du=None, u_back=None):
        for a in (u, du, u_back):
            assert a is None or (a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.npoin * self.neqs)
        self._ck(lib().jx_rhs(self._h, float(t), _ptr(u), _ptr(du), _ptr(u_back)))

    def step(self, scheme, t, dt, nsteps=1):
        self._ck(lib().jx_step(self._h, scheme, float(t), float(dt), int(nsteps)))

    def sync(self):
        self._ck(lib().jx_sync(self._h))

    def last_elapsed_ms(self):
        ms = ctypes.c_float()
        self._ck(lib().jx_last_elapsed_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib().jx_launch_count(self._h))

    def split_info(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        self._ck(lib().jx_split_info(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def selftest(self, which, n):
        bad = ctypes.c_int64()
        self._ck(lib().jx_selftest(self._h, int(which), int(n), ctypes.byref(bad)))
        return bad.value

    def bench_rhs(self, n, fused_stage=False, phases=True):
        tot = ctypes.c_float()
        ph = (ctypes.c_float * 8)()
        self._ck(lib().jx_bench_rhs(self._h, n, int(fused_stage), ctypes.byref(tot), ph if phases else None))
        return tot.value, list(ph)
