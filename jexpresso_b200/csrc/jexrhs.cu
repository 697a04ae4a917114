// jexrhs.cu -- the C ABI of libjexrhs (include/jexrhs.h): context, uploads, the per-stage RHS
// pipeline and the fused stage drivers.  Kernels live in jx_kernels.cuh, their instantiations
// in jx_inst_*.cu.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/jexrhs.h"
#include "jx_internal.h"
#include "jx_bdyflux.cuh"

using namespace jx;

// ------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so the library loads -- and reports a clear error -- on hosts
// without NCCL, and shares the copy PyTorch already mapped when both live in one process.
// ------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclUid { char internal[128]; };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(NcclUid *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclUid, int) = nullptr;
    int (*CommInitRankConfig)(ncclComm_t *, int, NcclUid, int, void *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
constexpr int kNcclFloat64 = 8;   // ncclFloat64 / ncclDouble
// ncclConfig_t as of NCCL 2.18 (nccl.h: size, magic, version, then the user fields); newer libraries accept an older,
// shorter struct by its size/version and default the fields it lacks.
struct NcclConfig218 {
    size_t size;
    unsigned int magic, version;
    int blocking, cgaClusterSize, minCTAs, maxCTAs;
    const char *netName;
    int splitShare;
};
constexpr int kNcclUndefInt = (-2147483647 - 1);

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *names[] = {getenv("JX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n) continue;
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) return api;
#define JX_SYM(field, sym) *(void **)(&api.field) = dlsym(api.h, sym)
    JX_SYM(GetUniqueId, "ncclGetUniqueId");
    JX_SYM(CommInitRank, "ncclCommInitRank");
    JX_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
    JX_SYM(CommDestroy, "ncclCommDestroy");
    JX_SYM(Send, "ncclSend");
    JX_SYM(Recv, "ncclRecv");
    JX_SYM(GroupStart, "ncclGroupStart");
    JX_SYM(GroupEnd, "ncclGroupEnd");
    JX_SYM(GetErrorString, "ncclGetErrorString");
#undef JX_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart &&
             api.GroupEnd;
    return api;
}
}  // namespace

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
enum { PH_BC = 0, PH_ELEM = 1, PH_DSS = 2, PH_HALO = 3, PH_UPDATE = 4, PH_AUX = 5, PH_COUNT = 8 };

struct PeerSeg {
    int peer;
    int64_t off, len;     // offset/length (in nodes) inside the concatenated list
};
struct AddRound {
    int64_t off, len;     // range inside d_add_sel (positions into the concatenated recv list)
};

struct jx_ctx {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    std::string err;
    int num_sms = 148;
    int nccl_max_ctas = 0;           // CTA cap the communicator was created with (0: none)
    int64_t launches = 0;

    // problem (jx_set_problem)
    bool have_problem = false, have_mesh = false;
    int nsd = 0, ngl = 0, neqs = 0, eq_id = 0, lpert = 0, lsource = 0, lvisc = 0;
    int64_t nelem = 0, npoin = 0;
    double visc[8] = {0};
    SgsArgs sgs;                     // jx_set_sgs: SMAG / VREM closure (model 0 = AV); sgs.ad_lvl points at d_ad_lvl
    int32_t *d_ad_lvl = nullptr;
    Phys phys;
    const KernelSet *ks = nullptr;
    int np = 0, nmet = 0, rec_bytes = 0;
    int rec_layout = -1;             // layout the resident element records were built with (-1: none)
    int visc_layout = 0;             // ... and the viscous pair records (KernelSet::visc_layout)

    // options
    int dss_mode = 0, pow_mode = 0, elem_variant = 0;

    // device arrays
    double *u = nullptr, *du = nullptr, *tmp = nullptr, *qe = nullptr, *Minv = nullptr, *coords = nullptr;
    double *rhs_el = nullptr, *rhs_el_visc = nullptr;
    double *aux = nullptr;           // per-node flux ingredient (k_node_aux)
    size_t aux_doubles = 0;
    bool aux_fresh = false;          // k_stage_direct left aux = aux(c->u): the next rhs_core on c->u skips k_node_aux
    bool acc_ready = false;          // c->tmp holds the low-storage register S' = S/dt pre-scaled by acc_A (direct accumulation)
    double acc_A = 0.0;
    int64_t *d_ifn = nullptr;        // unique nodes of the assembler lists (interface / periodic twins), ascending
    int64_t n_ifn = 0;
    double *d_ifbase = nullptr;      // their share of S' while the exchange sums the pure contributions
    char *rec = nullptr;
    char *rec_visc = nullptr;        // pair records of the viscous pass (k_visc_quad), lvisc only
    int32_t *d_eorig = nullptr;      // team records: position -> element id (order_elements); nullptr = identity
    int32_t *d_epos = nullptr;       // element -> position (kept for the weight rows rebuilt by finalize_mass)
    double *massw = nullptr;         // device-built mass: (w_i*w_j)*w_k*Je per element node, [l][iel] (jx_upload_mesh_coords with Minv = NULL)
    bool mass_pending = false;       // c->Minv still holds this rank's un-assembled M: finalize_mass() before first use
    int64_t *n2e_ptr = nullptr;
    uint32_t *n2e_idx = nullptr;
    double dpsi[64] = {0};
    double *ss[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // SSPRK scratch (uprev, u2, u3, u4, k3)

    // boundary fluxes with the MOST wall model (jx_upload_bdy_fluxes)
    int nwn = 0, nwnode = 0;          // wall-face nodes, unique wall nodes
    int32_t *wf_ip1 = nullptr, *wf_ipsfc = nullptr, *wf_node = nullptr, *wf_ptr = nullptr, *wf_hit = nullptr;
    double *wf_normal = nullptr, *wf_wJ = nullptr, *wf_sface = nullptr;
    double most_c[3] = {0.4, 0.1, 0.01}, delta_hf = 0.0, user_heatflux = 0.0;

    // boundary projection lists
    int nb = 0;
    int32_t *bc_node = nullptr, *bc_ptr = nullptr;
    double *bc_normal = nullptr;

    // interface / periodic assembly (AssemblerCache)
    bool have_halo = false;
    std::vector<PeerSeg> send_seg, recv_seg;        // by peer, ascending
    int64_t nsend = 0, nrecv = 0;
    int64_t *d_send_i = nullptr, *d_recv_idx = nullptr, *d_recvback_idx = nullptr;
    double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
    std::vector<std::vector<AddRound>> add_rounds;   // per recv segment: rounds of conflict-free positions
    int64_t *d_add_sel = nullptr;

    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    bool phase_timing = false;
    std::vector<cudaEvent_t> ph_ev;                  // pairs (start, stop) tagged by ph_tag
    std::vector<cudaEvent_t> ph_pool;                // recycled events
    std::vector<int> ph_tag;
    int use_graph = 0;                               // jx_bench_rhs / jx_step replay one captured RHS (JX_OPT_CUDA_GRAPH)

    // interface-first split (JX_OPT_OVERLAP): groups touching a shared node run first, the exchange then runs on a
    // second, high-priority stream beside the launch over the interior groups
    int overlap_sms = 0;                             // 0 = off; n > 0: n SMs kept free of the interior launch
    bool split_ready = false;
    int32_t *d_glist = nullptr;                      // [ngroups]: interface groups, then interior groups
    int n_iface = 0, n_inner = 0;
    int *d_gctr = nullptr;                           // work counters of the two list-driven launches
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

int fail(jx_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(c, e_ == cudaErrorMemoryAllocation ? JX_ENOMEM : JX_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, \
                        #call, cudaGetErrorString(e_));                                                   \
    } while (0)

#define NCK(call)                                                                                               \
    do {                                                                                                        \
        int e_ = (call);                                                                                        \
        if (e_ != 0)                                                                                            \
            return fail(c, JX_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call,                                 \
                        nccl().GetErrorString ? nccl().GetErrorString(e_) : "nccl error");                      \
    } while (0)

template <class T>
int dalloc(jx_ctx *c, T **p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    CK(cudaMalloc((void **)p, count * sizeof(T)));
    return JX_OK;
}
template <class T>
void dfree(T *&p) {
    if (p) cudaFree(p);
    p = nullptr;
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

struct PhaseScope {
    jx_ctx *c;
    bool on;
    PhaseScope(jx_ctx *c_, int tag) : c(c_), on(c_->phase_timing) {
        if (!on) return;
        cudaEvent_t a, b;
        if (c->ph_pool.size() >= 2) {          // events are recycled between jx_bench_rhs calls
            a = c->ph_pool.back(); c->ph_pool.pop_back();
            b = c->ph_pool.back(); c->ph_pool.pop_back();
        } else {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
        }
        c->ph_ev.push_back(a);
        c->ph_ev.push_back(b);
        c->ph_tag.push_back(tag);
        cudaEventRecord(a, c->stream);
    }
    ~PhaseScope() {
        if (on) cudaEventRecord(c->ph_ev.back(), c->stream);
    }
};

void free_split(jx_ctx *c) {
    dfree(c->d_glist);
    c->split_ready = false;
    c->n_iface = c->n_inner = 0;
}

void free_bdy_fluxes(jx_ctx *c) {
    dfree(c->wf_ip1); dfree(c->wf_ipsfc); dfree(c->wf_node); dfree(c->wf_ptr); dfree(c->wf_hit);
    dfree(c->wf_normal); dfree(c->wf_wJ); dfree(c->wf_sface);
    c->nwn = c->nwnode = 0;
}

void free_mesh(jx_ctx *c) {
    free_split(c);
    free_bdy_fluxes(c);              // the wall lists name nodes of the mesh they were built for
    dfree(c->u); dfree(c->du); dfree(c->tmp); dfree(c->qe); dfree(c->Minv); dfree(c->coords);
    dfree(c->rhs_el); dfree(c->rhs_el_visc); dfree(c->aux); c->aux_doubles = 0; dfree(c->rec); dfree(c->rec_visc); dfree(c->d_eorig); dfree(c->d_epos); dfree(c->massw); c->mass_pending = false; dfree(c->n2e_ptr); dfree(c->n2e_idx);
    for (auto &p : c->ss) dfree(p);
    c->have_mesh = false;
    c->rec_layout = -1;
    c->visc_layout = 0;
}
void free_bcs(jx_ctx *c) {
    dfree(c->bc_node); dfree(c->bc_ptr); dfree(c->bc_normal);
    c->nb = 0;
}

void free_halo(jx_ctx *c) {
    free_split(c);
    dfree(c->d_send_i); dfree(c->d_recv_idx); dfree(c->d_recvback_idx); dfree(c->d_sendbuf); dfree(c->d_recvbuf);
    dfree(c->d_add_sel); dfree(c->d_ifn); dfree(c->d_ifbase);
    c->n_ifn = 0;
    c->send_seg.clear(); c->recv_seg.clear(); c->add_rounds.clear();
    c->nsend = c->nrecv = 0;
    c->have_halo = false;
}

int select_kernels(jx_ctx *c) {
    const KernelSet *ks = nullptr;
    const int vcode = c->lvisc ? (c->sgs.model ? 2 : 1) : 0;   // KernelSet::lvisc: 0 inviscid, 1 AV, 2 SGS closure (generic kernel only)
    auto lookup = [&](int variant) -> const KernelSet * {
        if (c->eq_id == JX_EQ_EULER_THETA && c->nsd == 3) return lookup_euler_theta_3d(c->ngl, c->lpert, c->pow_mode, vcode, variant);
        if (c->eq_id == JX_EQ_EULER_THETA && c->nsd == 2) return lookup_euler_theta_2d(c->ngl, c->lpert, c->pow_mode, vcode, variant);
        return lookup_other(c->nsd, c->ngl, c->eq_id, vcode, variant);
    };
    if (c->elem_variant == JX_ELEM_AUTO) {
        // fastest exact-order kernel that exists for this configuration: the warp-team kernels (bit-identical to the
        // generic one), else the generic thread-per-node kernel.  Resident records pin the choice to their layout.
        // 9/8: nop 4/2.  Opt-in only: 12 (k_elem_tri, nop 7: 23.8 GDOF/s against 27.5 for the generic kernel, profiles/r02c)
        // 13 = 9 + the four-warp viscous pass k_visc_quad (AV decks; 21.6 against 18.7 GDOF/s for its removed predecessor k_visc_team, profiles/r02e)
        const int order[] = {13, 9, 8, 0};
        for (int v : order) {
            ks = lookup(v);
            if (ks && (!c->have_mesh || (ks->rec_layout == c->rec_layout && ks->visc_layout == c->visc_layout))) break;
            ks = nullptr;
        }
    } else {
        ks = lookup(c->elem_variant == JX_ELEM_GENERIC ? 0 : c->elem_variant);
    }
    if (!ks)
        return fail(c, JX_EINVAL, "no kernel for nsd=%d ngl=%d eq=%d lpert=%d pow=%d lvisc=%d sgs=%d variant=%d", c->nsd, c->ngl,
                    c->eq_id, c->lpert, c->pow_mode, c->lvisc, c->sgs.model, c->elem_variant);
    if (ks->neq != c->neqs) return fail(c, JX_EINVAL, "equation set %d has %d equations, got neqs=%d", c->eq_id, ks->neq, c->neqs);
    if (c->have_mesh && (ks->rec_layout != c->rec_layout || ks->visc_layout != c->visc_layout))
        return fail(c, JX_ESTATE, "kernel variant %d reads element-record layout %d/%d, the resident records have layout %d/%d: "
                    "set JX_OPT_ELEM_KERNEL before jx_upload_mesh", ks->variant, ks->rec_layout, ks->visc_layout, c->rec_layout, c->visc_layout);
    CK(ks->prepare());
    if (ks->visc_prepare) CK(ks->visc_prepare());
    c->ks = ks;
    return JX_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------
extern "C" int jx_version(void) { return 100; }

extern "C" int jx_nccl_unique_id(void *uid128) {
    if (!uid128) return JX_EINVAL;
    if (!nccl().ok) return JX_ENCCL;
    NcclUid id;
    if (nccl().GetUniqueId(&id) != 0) return JX_ENCCL;
    memcpy(uid128, &id, 128);
    return JX_OK;
}

extern "C" int jx_init(int device, int rank, int nranks, const void *nccl_uid, jx_ctx **out) {
    return jx_init_ex(device, rank, nranks, nccl_uid, 0, out);
}

extern "C" int jx_init_ex(int device, int rank, int nranks, const void *nccl_uid, int nccl_max_ctas, jx_ctx **out) {
    if (!out) return JX_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return JX_ENODEV;   // no CPU fallback
    if (device < 0 || device >= ndev || nranks < 1 || rank < 0 || rank >= nranks) return JX_EINVAL;
    jx_ctx *c = new jx_ctx();
    c->device = device; c->rank = rank; c->nranks = nranks;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return JX_ECUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return JX_ECUDA; }
    c->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return JX_ECUDA; }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    if (nranks > 1) {
        if (!nccl_uid || !nccl().ok) { jx_destroy(c); return JX_ENCCL; }
        NcclUid id;
        memcpy(&id, nccl_uid, 128);
        // The interface-first overlap (JX_OPT_OVERLAP) leaves nccl_max_ctas SMs free of the interior launch; the exchange's
        // send/recv kernels take one CTA per channel, so the communicator is created with that many CTAs at most.
        int rcn = -1;
        if (nccl_max_ctas > 0 && nccl().CommInitRankConfig) {
            NcclConfig218 cfg;
            cfg.size = sizeof(NcclConfig218); cfg.magic = 0xcafebeef; cfg.version = 21803;
            cfg.blocking = kNcclUndefInt; cfg.cgaClusterSize = kNcclUndefInt; cfg.minCTAs = 1; cfg.maxCTAs = nccl_max_ctas;
            cfg.netName = nullptr; cfg.splitShare = kNcclUndefInt;
            rcn = nccl().CommInitRankConfig(&c->comm, nranks, id, rank, &cfg);
            if (rcn != 0) c->comm = nullptr;
        }
        if (rcn != 0 && nccl().CommInitRank(&c->comm, nranks, id, rank) != 0) { jx_destroy(c); return JX_ENCCL; }
        c->nccl_max_ctas = rcn == 0 ? nccl_max_ctas : 0;
    }
    *out = c;
    return JX_OK;
}

extern "C" void jx_destroy(jx_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    free_mesh(c); free_bcs(c); free_halo(c); free_bdy_fluxes(c);
    dfree(c->d_gctr); dfree(c->d_ad_lvl);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    for (auto e : c->ph_ev) cudaEventDestroy(e);
    for (auto e : c->ph_pool) cudaEventDestroy(e);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int jx_last_error(jx_ctx *c, char *buf, int len) {
    if (!c || !buf || len <= 0) return JX_EINVAL;
    snprintf(buf, (size_t)len, "%s", c->err.c_str());
    return JX_OK;
}

extern "C" int jx_set_option(jx_ctx *c, int key, int64_t value) {
    if (!c) return JX_EINVAL;
    if (key != JX_OPT_CUDA_GRAPH) { c->aux_fresh = false; c->acc_ready = false; }
    const int old_dss = c->dss_mode, old_pow = c->pow_mode, old_var = c->elem_variant;
    switch (key) {
        case JX_OPT_DSS_MODE:
            if (value != 0 && value != 1) return fail(c, JX_EINVAL, "dss mode must be 0 or 1");
            c->dss_mode = (int)value;
            break;
        case JX_OPT_POW_MODE:
            if (value != 0 && value != 1) return fail(c, JX_EINVAL, "pow mode must be 0 or 1");
            c->pow_mode = (int)value;
            break;
        case JX_OPT_ELEM_KERNEL: c->elem_variant = (int)value; break;
        case JX_OPT_CUDA_GRAPH: c->use_graph = value ? 1 : 0; return JX_OK;
        case JX_OPT_OVERLAP:
            if (value < 0 || value > 64) return fail(c, JX_EINVAL, "overlap: 0 (off) or the number of SMs kept free (1..64)");
            c->overlap_sms = (int)value;
            free_split(c);
            return JX_OK;
        default: return fail(c, JX_EINVAL, "unknown option %d", key);
    }
    free_split(c);   // the split depends on the DSS mode and on the kernel's record layout: rebuilt at the next evaluation
    if (c->have_problem) {
        const int rc = select_kernels(c);
        if (rc) { c->dss_mode = old_dss; c->pow_mode = old_pow; c->elem_variant = old_var; }   // refused: keep the working set
        return rc;
    }
    return JX_OK;
}

extern "C" int jx_set_problem(jx_ctx *c, int nsd, int ngl, int neqs, int64_t nelem, int64_t npoin, int equation_id, int lpert,
                              int lsource, int lvisc, const double *visc_coeff, const double *phys_consts, int nphys) {
    if (!c) return JX_EINVAL;
    if ((nsd != 2 && nsd != 3) || ngl < 2 || ngl > 8 || neqs < 1 || neqs > 8 || nelem < 0 || npoin < 0 || nphys < 0 || nphys > 16)
        return fail(c, JX_EINVAL, "jx_set_problem: bad dimensions");
    if (npoin >= (int64_t)1 << 31) return fail(c, JX_EINVAL, "npoin must be < 2^31 per rank");
    cudaSetDevice(c->device);
    c->aux_fresh = false; c->acc_ready = false;
    free_mesh(c); free_bcs(c); free_halo(c); free_bdy_fluxes(c);
    c->nsd = nsd; c->ngl = ngl; c->neqs = neqs; c->nelem = nelem; c->npoin = npoin;
    c->eq_id = equation_id; c->lpert = lpert ? 1 : 0; c->lsource = lsource ? 1 : 0; c->lvisc = lvisc ? 1 : 0;
    c->sgs = SgsArgs();              // AV until jx_set_sgs says otherwise
    dfree(c->d_ad_lvl);
    for (int i = 0; i < 8; ++i) c->visc[i] = (visc_coeff && i < neqs) ? visc_coeff[i] : 0.0;
    for (int i = 0; i < 16; ++i) c->phys.v[i] = (phys_consts && i < nphys) ? phys_consts[i] : 0.0;
    c->np = 1;
    for (int d = 0; d < nsd; ++d) c->np *= ngl;
    c->nmet = nsd * nsd + 1;
    const int npp = (c->np + 3) / 4 * 4;
    c->rec_bytes = (c->nmet * c->np * 8 + npp * 4 + 15) / 16 * 16;
    if ((int64_t)c->np * nelem >= ((int64_t)1 << 32)) return fail(c, JX_EINVAL, "nelem*ngl^nsd must be < 2^32 per rank");
    c->have_problem = true;
    return select_kernels(c);
}

// replaces: allocate_SGS (sgsStructs.jl:77-120) + the flags of params_setup.jl:249-253
extern "C" int jx_set_sgs(jx_ctx *c, int visc_model, double delta_effective, int lrichardson, int ltheta_eqn, const double *consts,
                          int nconsts, const int64_t *ad_lvl) {
    if (!c) return JX_EINVAL;
    if (!c->have_problem) return fail(c, JX_ESTATE, "jx_set_sgs before jx_set_problem");
    if (visc_model < JX_VISC_AV || visc_model > JX_VISC_VREM) return fail(c, JX_EINVAL, "visc_model must be JX_VISC_AV, _SMAG or _VREM");
    cudaSetDevice(c->device);
    c->aux_fresh = false; c->acc_ready = false;
    const SgsArgs old = c->sgs;
    SgsArgs g;
    if (visc_model != JX_VISC_AV) {
        if (!c->lvisc) return fail(c, JX_EINVAL, "an SGS closure needs lvisc = 1 in jx_set_problem");
        if (!consts || nconsts < 6) return fail(c, JX_EINVAL, "jx_set_sgs: consts = [Pr_t, Sc_t, mu_mol, kappa_mol, Ri_crit, C_s]");
        if (!(delta_effective > 0.0)) return fail(c, JX_EINVAL, "jx_set_sgs: delta_effective must be positive");
        g.model = visc_model; g.lrichardson = lrichardson ? 1 : 0; g.ltheta_eqn = ltheta_eqn ? 1 : 0;
        g.delta = delta_effective;
        g.Pr_t = consts[0]; g.Sc_t = consts[1]; g.mu_mol = consts[2]; g.kappa_mol = consts[3]; g.Ri_crit = consts[4];
        g.C_s2 = consts[5] * consts[5];                   // T(PhysConst.C_s * PhysConst.C_s)
        g.C_vrem = 2.5 * consts[5] * consts[5];           // T(2.5 * PhysConst.C_s * PhysConst.C_s)
        g.g = c->phys.v[2];
    }
    dfree(c->d_ad_lvl);
    if (visc_model != JX_VISC_AV && ad_lvl && c->nelem > 0) {
        std::vector<int32_t> lv((size_t)c->nelem);
        for (int64_t e = 0; e < c->nelem; ++e) {
            if (ad_lvl[e] < 0 || ad_lvl[e] > 1000) return fail(c, JX_EINVAL, "jx_set_sgs: ad_lvl[%lld] = %lld", (long long)e, (long long)ad_lvl[e]);
            lv[(size_t)e] = (int32_t)ad_lvl[e];
        }
        int rc = dalloc(c, &c->d_ad_lvl, (size_t)c->nelem);
        if (rc) return rc;
        CK(cudaMemcpy(c->d_ad_lvl, lv.data(), (size_t)c->nelem * sizeof(int32_t), cudaMemcpyHostToDevice));
        g.ad_lvl = c->d_ad_lvl;
    }
    c->sgs = g;
    free_split(c);
    const int rc = select_kernels(c);
    if (rc) { c->sgs = old; c->sgs.ad_lvl = nullptr; }    // refused (e.g. resident team-kernel records): the AV / previous set stays
    return rc;
}

// ------------------------------------------------------------------------------------------
// uploads
// ------------------------------------------------------------------------------------------
namespace {

// Record order of the team kernels: elements sorted along a Morton curve through their centres (cells of one mean element
// size), so that the elements a CTA wave works on share faces, edges and vertices -- their q / aux gathers and du RED.ADDs
// then hit L2 instead of DRAM (x-fastest order of a 73^3 box: the z-neighbour comes 5329 elements = 75 MB of records
// later; measured L2 hit rate 38 %, profiles/r01g).  Pure data placement: every element keeps its arithmetic, rhs_el keeps
// the caller's numbering (ElemArgs::eorig) and the deterministic gather its element-ascending order.
// JX_ELEM_ORDER=0 in the environment keeps the caller's order.
int order_elements(jx_ctx *c, const int64_t *connijk, const double *coords, std::vector<int32_t> &epos) {
    const int64_t E = c->nelem;
    const int np = c->np, nsd = c->nsd;
    epos.clear();
    const char *env = getenv("JX_ELEM_ORDER");
    if (!coords || E < 2 || (env && env[0] == '0')) return JX_OK;
    std::vector<double> cen((size_t)E * 3, 0.0);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t iel = 0; iel < E; ++iel) {
        const int64_t a = connijk[(size_t)iel] - 1, b = connijk[(size_t)iel + (size_t)E * (np - 1)] - 1;   // opposite corners
        if (a < 0 || a >= c->npoin || b < 0 || b >= c->npoin) return JX_OK;                                 // reported by the CSR build
        for (int d = 0; d < nsd; ++d) {
            const double x = 0.5 * (coords[(size_t)a * nsd + d] + coords[(size_t)b * nsd + d]);
            cen[(size_t)iel * 3 + d] = x;
            lo[d] = std::min(lo[d], x); hi[d] = std::max(hi[d], x);
        }
    }
    double vol = 1.0;
    int dims = 0;
    for (int d = 0; d < nsd; ++d)
        if (hi[d] > lo[d]) { vol *= hi[d] - lo[d]; ++dims; }
    if (dims == 0) return JX_OK;
    const double h = std::pow(vol / (double)E, 1.0 / dims);
    if (!(h > 0.0) || !std::isfinite(h)) return JX_OK;
    auto spread = [](uint64_t v) {   // 21 bits -> every third bit
        v &= 0x1fffffull;
        v = (v | v << 32) & 0x1f00000000ffffull;
        v = (v | v << 16) & 0x1f0000ff0000ffull;
        v = (v | v << 8) & 0x100f00f00f00f00full;
        v = (v | v << 4) & 0x10c30c30c30c30c3ull;
        v = (v | v << 2) & 0x1249249249249249ull;
        return v;
    };
    std::vector<std::pair<uint64_t, int32_t>> key((size_t)E);
    for (int64_t iel = 0; iel < E; ++iel) {
        uint64_t k = 0;
        for (int d = 0; d < nsd; ++d) {
            const double q = (cen[(size_t)iel * 3 + d] - lo[d]) / h;
            k |= spread((uint64_t)std::min(q < 0 ? 0.0 : q, 2097151.0)) << d;
        }
        key[(size_t)iel] = {k, (int32_t)iel};
    }
    std::sort(key.begin(), key.end());
    epos.resize((size_t)E);
    std::vector<int32_t> eorig((size_t)E);
    for (int64_t p = 0; p < E; ++p) { eorig[(size_t)p] = key[(size_t)p].second; epos[(size_t)key[(size_t)p].second] = (int32_t)p; }
    int rc = dalloc(c, &c->d_eorig, (size_t)E);
    if (rc) return rc;
    CK(cudaMemcpy(c->d_eorig, eorig.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
    return JX_OK;
}

// The rows of the element records that hold M^-1 (folded scatter weight of the team / tri kernels, the Minv row of the viscous
// records) are (re)built from the ids and weights already in the records: at upload, and again once a device-built mass matrix
// has been assembled across the ranks (finalize_mass).
void rebuild_minv_rows(jx_ctx *c) {
    const KernelSet *ks = c->ks;
    const int64_t total = c->nelem * c->np;
    if (total <= 0) return;
    if (ks->rec_layout == 7) {
        TriRetileArgs ta;
        ta.src = nullptr; ta.omega = nullptr; ta.Minv = c->Minv; ta.connijk = nullptr; ta.rec = c->rec; ta.nelem = c->nelem; ta.ngl = c->ngl;
        ta.rec_bytes = ks->group_bytes; ta.zid_off = ks->zid_off; ta.fid_off = ks->fid_off; ta.slot = -2;
        k_retile_tri<<<nblk(total, 256), 256, 0, c->stream>>>(ta);
        c->launches++;
    } else if (ks->rec_layout == 5) {
        GroupRetileArgs ga;
        ga.src = nullptr; ga.omega = nullptr; ga.Minv = c->Minv; ga.connijk = nullptr; ga.epos = c->d_epos; ga.rec = c->rec; ga.nelem = c->nelem;
        ga.ngl = c->ngl; ga.epb = ks->elems_per_block; ga.group_bytes = ks->group_bytes;
        ga.zid_off = ks->zid_off; ga.fid_off = ks->fid_off; ga.z_off = ks->z_off; ga.w_off = ks->w_off; ga.wf_off = ks->wf_off; ga.slot = -2;
        k_retile_group<<<nblk(total, 256), 256, 0, c->stream>>>(ga);
        c->launches++;
        if (ks->launch_visc && c->rec_visc) {
            ViscRetileArgs vr;
            vr.src = nullptr; vr.omega = nullptr; vr.Minv = c->Minv; vr.connijk = nullptr; vr.epos = c->d_epos; vr.rec = c->rec_visc; vr.nelem = c->nelem;
            vr.ngl = c->ngl; vr.epb = ks->elems_per_block; vr.group_bytes = ks->visc_group_bytes;
            vr.zid_off = ks->visc_zid_off; vr.fid_off = ks->visc_fid_off; vr.slot = -2;
            ks->retile_visc(vr, (unsigned)nblk(total, 256), c->stream);
            c->launches++;
        }
    }
}

int assemble(jx_ctx *c, double *a, cudaStream_t s, int ncomp);

// DSS_global_mass! + Minv = 1 ./ M (element_matrices.jl:1160-1174, 1557-1559) for a mass matrix built on the device
int finalize_mass(jx_ctx *c) {
    if (!c->mass_pending) return JX_OK;
    c->mass_pending = false;
    if (c->have_halo) {
        int rc = assemble(c, c->Minv, c->stream, 1);
        if (rc) return rc;
    }
    if (c->npoin > 0) k_invert<<<nblk(c->npoin, 256), 256, 0, c->stream>>>(c->Minv, c->npoin);
    c->launches++;
    rebuild_minv_rows(c);
    CK(cudaGetLastError());
    return JX_OK;
}

// metrics == nullptr: the metric arrays are built on the device from coords (k_build_metric) instead of uploaded
int upload_mesh_impl(jx_ctx *c, const int64_t *connijk, const double *coords, const double *const *metrics,
                     const double *dpsi, const double *omega, const double *Minv, const double *qe) {
    cudaSetDevice(c->device);
    free_mesh(c);
    const int64_t E = c->nelem, N = c->npoin;
    const int np = c->np, q = c->neqs;
    const int64_t total = E * np;
    const size_t nq = (size_t)N * q;
    const bool tri = c->ks->rec_layout == 7;          // per-element pencil-stream records of k_elem_tri
    const bool grouped = c->ks->rec_layout == 5;   // element-group records of the team kernels
    const int64_t ngroups = grouped ? (E + c->ks->elems_per_block - 1) / c->ks->elems_per_block : 0;
    const size_t rec_total = grouped ? (size_t)ngroups * c->ks->group_bytes : (tri ? (size_t)E * c->ks->group_bytes : (size_t)E * c->rec_bytes);
    int rc;
    if ((rc = dalloc(c, &c->u, nq)) || (rc = dalloc(c, &c->du, nq)) || (rc = dalloc(c, &c->tmp, nq)) ||
        (rc = dalloc(c, &c->Minv, (size_t)N)) || (rc = dalloc(c, &c->qe, (size_t)N * (q + 1))) ||
        (rc = dalloc(c, &c->coords, (size_t)N * c->nsd)) || (rc = dalloc(c, &c->rec, rec_total)) ||
        (rc = dalloc(c, &c->n2e_ptr, (size_t)N + 1)) || (rc = dalloc(c, &c->n2e_idx, (size_t)total)))
        return rc;
    CK(cudaMemsetAsync(c->u, 0, nq * 8, c->stream));
    CK(cudaMemsetAsync(c->du, 0, nq * 8, c->stream));
    CK(cudaMemsetAsync(c->tmp, 0, nq * 8, c->stream));
    if (Minv) CK(cudaMemcpyAsync(c->Minv, Minv, (size_t)N * 8, cudaMemcpyHostToDevice, c->stream));
    else CK(cudaMemsetAsync(c->Minv, 0, (size_t)std::max<int64_t>(1, N) * 8, c->stream));      // built below from Je (device mass)
    if (qe) CK(cudaMemcpyAsync(c->qe, qe, (size_t)N * (q + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    else CK(cudaMemsetAsync(c->qe, 0, (size_t)N * (q + 1) * 8, c->stream));
    for (int i = 0; i < c->ngl * c->ngl; ++i) c->dpsi[i] = dpsi[i];

    // staging buffer reused for every element-sized host array
    double *d_stage = nullptr, *d_omega = nullptr, *d_dpsi = nullptr;
    int64_t *d_conn = nullptr;
    int32_t *d_cnt = nullptr;
    int32_t *&d_epos = c->d_epos;
    if ((rc = dalloc(c, &d_stage, (size_t)std::max<int64_t>(total, N * c->nsd))) || (rc = dalloc(c, &d_omega, (size_t)c->ngl)) ||
        (rc = dalloc(c, &d_dpsi, (size_t)c->ngl * c->ngl)) || (rc = dalloc(c, &d_conn, (size_t)total)) ||
        (rc = dalloc(c, &d_cnt, (size_t)N)))
        return rc;
    auto cleanup = [&]() { dfree(d_stage); dfree(d_omega); dfree(d_dpsi); dfree(d_conn); dfree(d_cnt); };
#define CKC(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cleanup();                                                                               \
            return fail(c, JX_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
        }                                                                                            \
    } while (0)
    CKC(cudaMemcpyAsync(d_omega, omega, (size_t)c->ngl * 8, cudaMemcpyHostToDevice, c->stream));
    CKC(cudaMemcpyAsync(d_conn, connijk, (size_t)total * 8, cudaMemcpyHostToDevice, c->stream));
    CKC(cudaMemcpyAsync(d_dpsi, dpsi, (size_t)c->ngl * c->ngl * 8, cudaMemcpyHostToDevice, c->stream));
    // coords: Julia [nsd, N] -> device [nsd][N]
    if (coords && N > 0) {
        std::vector<double> soa((size_t)N * c->nsd);
        for (int d = 0; d < c->nsd; ++d)
            for (int64_t ip = 0; ip < N; ++ip) soa[(size_t)d * N + ip] = coords[(size_t)ip * c->nsd + d];
        CKC(cudaMemcpy(c->coords, soa.data(), soa.size() * 8, cudaMemcpyHostToDevice));
    } else {
        CKC(cudaMemsetAsync(c->coords, 0, (size_t)std::max<int64_t>(1, N * c->nsd) * 8, c->stream));
    }
    // metric array m in the reference's layout -> d_stage: uploaded, or built from the coordinates on the device
    auto stage_metric = [&](int m) -> cudaError_t {
        if (metrics) return cudaMemcpyAsync(d_stage, metrics[m], (size_t)total * 8, cudaMemcpyHostToDevice, c->stream);
        MetricBuildArgs mb;
        mb.connijk = d_conn; mb.coords = c->coords; mb.dpsi = d_dpsi; mb.out = d_stage; mb.nelem = E; mb.npoin = N;
        mb.nsd = c->nsd; mb.ngl = c->ngl; mb.slot = m;
        k_build_metric<<<nblk(total, 256), 256, 0, c->stream>>>(mb);
        c->launches++;
        return cudaGetLastError();
    };
    RetileArgs ra;
    ra.omega = d_omega; ra.connijk = d_conn; ra.rec = c->rec; ra.nelem = E; ra.nsd = c->nsd; ra.ngl = c->ngl; ra.np = np;
    ra.nmet = c->nmet; ra.npp = (np + 3) / 4 * 4; ra.rec_bytes = c->rec_bytes; ra.src = nullptr;
    if (total > 0 && tri) {
        const KernelSet *ks = c->ks;
        TriRetileArgs ta;
        ta.src = nullptr; ta.omega = d_omega; ta.Minv = c->Minv; ta.connijk = d_conn; ta.rec = c->rec; ta.nelem = E; ta.ngl = c->ngl;
        ta.rec_bytes = ks->group_bytes; ta.zid_off = ks->zid_off; ta.fid_off = ks->fid_off;
        CKC(cudaMemsetAsync(c->rec, 0, rec_total, c->stream));
        ta.slot = -1;
        k_retile_tri<<<nblk(total, 256), 256, 0, c->stream>>>(ta);
        for (int m = 0; m < c->nmet; ++m) {
            if (metrics && !metrics[m]) { cleanup(); return fail(c, JX_EINVAL, "metric array %d is null", m); }
            CKC(stage_metric(m));
            ta.src = d_stage; ta.slot = m;
            k_retile_tri<<<nblk(total, 256), 256, 0, c->stream>>>(ta);
            CKC(cudaStreamSynchronize(c->stream));
        }
        c->launches += c->nmet + 1;
    } else if (total > 0 && grouped) {
        const KernelSet *ks = c->ks;
        std::vector<int32_t> epos;
        if ((rc = order_elements(c, connijk, coords, epos))) { cleanup(); return rc; }
        if (!epos.empty()) {
            if ((rc = dalloc(c, &d_epos, (size_t)E))) { cleanup(); return rc; }
            CKC(cudaMemcpyAsync(d_epos, epos.data(), (size_t)E * 4, cudaMemcpyHostToDevice, c->stream));
            CKC(cudaStreamSynchronize(c->stream));
        }
        GroupRetileArgs ga;
        ga.src = nullptr; ga.omega = d_omega; ga.Minv = c->Minv; ga.connijk = d_conn; ga.epos = d_epos; ga.rec = c->rec; ga.nelem = E;
        ga.ngl = c->ngl; ga.epb = ks->elems_per_block; ga.group_bytes = ks->group_bytes;
        ga.zid_off = ks->zid_off; ga.fid_off = ks->fid_off; ga.z_off = ks->z_off; ga.w_off = ks->w_off; ga.wf_off = ks->wf_off;
        CKC(cudaMemsetAsync(c->rec, 0, rec_total, c->stream));
        ga.slot = -1;
        k_retile_group<<<nblk(total, 256), 256, 0, c->stream>>>(ga);
        ViscRetileArgs vr;
        const bool with_visc = ks->launch_visc != nullptr;
        if (with_visc) {                               // records of the viscous pass, built from the same staged arrays
            const size_t vbytes = (size_t)ngroups * ks->visc_group_bytes;
            if ((rc = dalloc(c, &c->rec_visc, vbytes))) { cleanup(); return rc; }
            CKC(cudaMemsetAsync(c->rec_visc, 0, vbytes, c->stream));
            vr.src = nullptr; vr.omega = d_omega; vr.Minv = c->Minv; vr.connijk = d_conn; vr.epos = d_epos; vr.rec = c->rec_visc; vr.nelem = E;
            vr.ngl = c->ngl; vr.epb = ks->elems_per_block; vr.group_bytes = ks->visc_group_bytes;
            vr.zid_off = ks->visc_zid_off; vr.fid_off = ks->visc_fid_off;
            vr.slot = -1;
            ks->retile_visc(vr, (unsigned)nblk(total, 256), c->stream);
        }
        for (int m = 0; m < c->nmet; ++m) {
            if (metrics && !metrics[m]) { cleanup(); return fail(c, JX_EINVAL, "metric array %d is null", m); }
            CKC(stage_metric(m));
            ga.src = d_stage; ga.slot = m;
            k_retile_group<<<nblk(total, 256), 256, 0, c->stream>>>(ga);
            if (with_visc) {
                vr.src = d_stage; vr.slot = m;
                ks->retile_visc(vr, (unsigned)nblk(total, 256), c->stream);
            }
            CKC(cudaStreamSynchronize(c->stream));
        }
        c->launches += (c->nmet + 1) * (with_visc ? 2 : 1);
    } else if (total > 0) {
        CKC(cudaMemsetAsync(c->rec, 0, rec_total, c->stream));
        ra.slot = -1;
        k_retile<<<nblk(total, 256), 256, 0, c->stream>>>(ra);
        for (int m = 0; m < c->nmet; ++m) {
            if (metrics && !metrics[m]) { cleanup(); return fail(c, JX_EINVAL, "metric array %d is null", m); }
            CKC(stage_metric(m));
            ra.src = d_stage; ra.slot = m;
            k_retile<<<nblk(total, 256), 256, 0, c->stream>>>(ra);
            CKC(cudaStreamSynchronize(c->stream));   // host array may be pageable: keep the staging reuse ordered
        }
        c->launches += c->nmet + 1;
    }
    if (!Minv && total > 0) {
        // device-built mass: d_stage still holds the last staged metric array, Je (metric_terms.jl:77 layout)
        if ((rc = dalloc(c, &c->massw, (size_t)total))) { cleanup(); return rc; }
        k_mass_weight<<<nblk(total, 256), 256, 0, c->stream>>>(d_stage, d_omega, E, c->ngl, c->nsd, c->massw);
        c->launches++;
    }
    // node -> (element, local) CSR in DSS_rhs! order (element ascending)
    {
        CKC(cudaMemsetAsync(d_cnt, 0, (size_t)std::max<int64_t>(1, N) * 4, c->stream));
        if (total > 0) k_count_valence<<<nblk(total, 256), 256, 0, c->stream>>>(d_conn, total, d_cnt);
        std::vector<int32_t> cnt((size_t)N);
        CKC(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)N * 4, cudaMemcpyDeviceToHost, c->stream));
        CKC(cudaStreamSynchronize(c->stream));
        std::vector<int64_t> ptr((size_t)N + 1);
        ptr[0] = 0;
        for (int64_t i = 0; i < N; ++i) ptr[i + 1] = ptr[i] + cnt[i];
        if (ptr[N] != total) { cleanup(); return fail(c, JX_EINVAL, "connijk holds ids outside 1..npoin"); }
        CKC(cudaMemcpyAsync(c->n2e_ptr, ptr.data(), ((size_t)N + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        CKC(cudaMemsetAsync(d_cnt, 0, (size_t)std::max<int64_t>(1, N) * 4, c->stream));
        if (total > 0) {
            k_fill_n2e<<<nblk(total, 256), 256, 0, c->stream>>>(d_conn, E, np, c->n2e_ptr, d_cnt, c->n2e_idx);
            k_sort_n2e<<<nblk(N, 128), 128, 0, c->stream>>>(N, c->n2e_ptr, c->n2e_idx);
            c->launches += 3;
        }
        CKC(cudaStreamSynchronize(c->stream));
    }
    c->rec_layout = c->ks->rec_layout;
    c->visc_layout = c->ks->visc_layout;
    if (!Minv) {
        // DSS_mass!: this rank's sums; the global assembly and the inversion wait for the halo lists (finalize_mass)
        if (N > 0) k_mass_gather<<<nblk(N, 128), 128, 0, c->stream>>>(c->n2e_ptr, c->n2e_idx, c->massw, E, np, N, 0, nullptr, c->Minv);
        c->launches++;
        c->mass_pending = true;
    } else {
        rebuild_minv_rows(c);
    }
    CKC(cudaStreamSynchronize(c->stream));
    CKC(cudaGetLastError());
    cleanup();
#undef CKC
    c->have_mesh = true;
    c->aux_fresh = false; c->acc_ready = false;
    return JX_OK;
}

}  // namespace

extern "C" int jx_upload_mesh(jx_ctx *c, const int64_t *connijk, const double *coords, const double *const *metrics,
                              int nmetrics, const double *dpsi, const double *omega, const double *Minv, const double *qe) {
    if (!c) return JX_EINVAL;
    if (!c->have_problem) return fail(c, JX_ESTATE, "jx_upload_mesh before jx_set_problem");
    if (!connijk || !metrics || !dpsi || !omega || !Minv) return fail(c, JX_EINVAL, "jx_upload_mesh: null array");
    if (nmetrics != c->nmet) return fail(c, JX_EINVAL, "expected %d metric arrays, got %d", c->nmet, nmetrics);
    return upload_mesh_impl(c, connijk, coords, metrics, dpsi, omega, Minv, qe);
}

extern "C" int jx_upload_mesh_coords(jx_ctx *c, const int64_t *connijk, const double *coords, const double *dpsi,
                                     const double *omega, const double *Minv, const double *qe) {
    if (!c) return JX_EINVAL;
    if (!c->have_problem) return fail(c, JX_ESTATE, "jx_upload_mesh_coords before jx_set_problem");
    if (!connijk || !coords || !dpsi || !omega) return fail(c, JX_EINVAL, "jx_upload_mesh_coords: null array");
    return upload_mesh_impl(c, connijk, coords, nullptr, dpsi, omega, Minv, qe);
}

extern "C" int jx_upload_bcs(jx_ctx *c, int64_t nfaces, const int64_t *poin_in_bdy_face, const double *nx, const double *ny,
                             const double *nz, const int32_t *face_bc_kind) {
    if (!c) return JX_EINVAL;
    if (!c->have_problem) return fail(c, JX_ESTATE, "jx_upload_bcs before jx_set_problem");
    cudaSetDevice(c->device);
    free_bcs(c);
    if (nfaces <= 0) return JX_OK;
    if (!poin_in_bdy_face || !nx || !ny || !face_bc_kind || (c->nsd == 3 && !nz)) return fail(c, JX_EINVAL, "jx_upload_bcs: null array");
    const int n = c->ngl;
    const int per_face = (c->nsd == 3) ? n * n : n;
    // hits in the reference's visiting order: face ascending, then i outer / j inner (BCs.jl:621-626)
    struct Hit { int32_t node; int64_t seq; double nx, ny, nz; };
    std::vector<Hit> hits;
    hits.reserve((size_t)nfaces * per_face);
    int64_t seq = 0;
    for (int64_t f = 0; f < nfaces; ++f) {
        if (face_bc_kind[f] == JX_BC_SKIP) continue;
        if (face_bc_kind[f] != JX_BC_FREE_SLIP) return fail(c, JX_EINVAL, "face %lld: unknown bc kind %d", (long long)f, face_bc_kind[f]);
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < (c->nsd == 3 ? n : 1); ++b) {
                const int64_t off = (c->nsd == 3) ? f + nfaces * (a + (int64_t)n * b) : f + nfaces * (int64_t)a;
                const int64_t ip = poin_in_bdy_face[off] - 1;
                if (ip < 0 || ip >= c->npoin) return fail(c, JX_EINVAL, "poin_in_bdy_face holds ids outside 1..npoin");
                hits.push_back({(int32_t)ip, seq++, nx[off], ny[off], nz ? nz[off] : 0.0});
            }
    }
    std::stable_sort(hits.begin(), hits.end(), [](const Hit &x, const Hit &y) { return x.node < y.node; });
    std::vector<int32_t> node, ptr;
    std::vector<double> normal(hits.size() * 3);
    for (size_t h = 0; h < hits.size(); ++h) {
        if (h == 0 || hits[h].node != hits[h - 1].node) { node.push_back(hits[h].node); ptr.push_back((int32_t)h); }
        normal[3 * h] = hits[h].nx; normal[3 * h + 1] = hits[h].ny; normal[3 * h + 2] = hits[h].nz;
    }
    ptr.push_back((int32_t)hits.size());
    c->nb = (int)node.size();
    if (c->nb == 0) return JX_OK;
    int rc;
    if ((rc = dalloc(c, &c->bc_node, node.size())) || (rc = dalloc(c, &c->bc_ptr, ptr.size())) ||
        (rc = dalloc(c, &c->bc_normal, normal.size())))
        return rc;
    CK(cudaMemcpy(c->bc_node, node.data(), node.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->bc_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->bc_normal, normal.data(), normal.size() * 8, cudaMemcpyHostToDevice));
    return JX_OK;
}

// replaces: inputs[:bdy_fluxes] -> build_custom_bcs_neumann!(::NSD_3D) (BCs.jl:655-816), the surface integrals and RHS .+= S_flux
extern "C" int jx_upload_bdy_fluxes(jx_ctx *c, int64_t nfaces, const int64_t *poin_in_bdy_face, const int64_t *bdy_face_in_elem,
                                    const int64_t *connijk, const double *nx, const double *ny, const double *nz, const double *Jef,
                                    const double *omega, const int32_t *face_flux_kind, int ifirst_wall_node_index, double delta_hf,
                                    double user_heatflux, const double *most_consts, int nconsts) {
    if (!c) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_upload_bdy_fluxes before jx_upload_mesh");
    if (c->nsd != 3 || c->neqs < 5) return fail(c, JX_EINVAL, "boundary fluxes: 3D CompEuler equation sets only (rho, rho u, rho v, rho w, rho theta)");
    if (c->eq_id != JX_EQ_EULER_THETA && c->eq_id != JX_EQ_EULER_THETA_LES) return fail(c, JX_EINVAL, "boundary fluxes: theta-form CompEuler only");
    cudaSetDevice(c->device);
    c->aux_fresh = false; c->acc_ready = false;
    free_bdy_fluxes(c);
    free_split(c);
    if (nfaces <= 0) return JX_OK;
    if (!poin_in_bdy_face || !bdy_face_in_elem || !connijk || !nx || !ny || !nz || !Jef || !omega || !face_flux_kind)
        return fail(c, JX_EINVAL, "jx_upload_bdy_fluxes: null array");
    const int n = c->ngl;
    if (ifirst_wall_node_index < 2 || ifirst_wall_node_index > n) return fail(c, JX_EINVAL, "ifirst_wall_node_index must be in 2..ngl");
    const int64_t E = c->nelem, N = c->npoin;
    if (most_consts && nconsts >= 3) { c->most_c[0] = most_consts[0]; c->most_c[1] = most_consts[1]; c->most_c[2] = most_consts[2]; }
    else { c->most_c[0] = 0.4; c->most_c[1] = 0.1; c->most_c[2] = 0.01; }   // PhysConst.karman; z0_m, z0_h of BCs.jl:770-772
    c->delta_hf = delta_hf; c->user_heatflux = user_heatflux;
    // wall-face nodes in the reference's visiting order (face ascending, i outer, j inner) and, per unique wall node, its hits
    struct Hit { int32_t node; int32_t w; };
    std::vector<int32_t> ip1, ipsfc;
    std::vector<double> normal, wJ;
    std::vector<Hit> hits;
    const int kw = ifirst_wall_node_index - 1;
    for (int64_t f = 0; f < nfaces; ++f) {
        if (face_flux_kind[f] == JX_FLUX_NONE) continue;      // user_bc_neumann!: F_surf stays zero
        if (face_flux_kind[f] != JX_FLUX_MOST) return fail(c, JX_EINVAL, "face %lld: unknown flux kind %d", (long long)f, face_flux_kind[f]);
        const int64_t e = bdy_face_in_elem[f] - 1;
        if (e < 0 || e >= E) return fail(c, JX_EINVAL, "bdy_face_in_elem holds ids outside 1..nelem");
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                const int64_t off = f + nfaces * (i + (int64_t)n * j);
                const int64_t ip = poin_in_bdy_face[off] - 1;
                const int64_t a1 = connijk[e + E * (i + n * (j + n * (int64_t)kw))] - 1;
                const int64_t as = connijk[e + E * (i + n * (int64_t)j)] - 1;
                if (ip < 0 || ip >= N || a1 < 0 || a1 >= N || as < 0 || as >= N) return fail(c, JX_EINVAL, "node ids outside 1..npoin");
                hits.push_back({(int32_t)ip, (int32_t)ip1.size()});
                ip1.push_back((int32_t)a1); ipsfc.push_back((int32_t)as);
                normal.push_back(nx[off]); normal.push_back(ny[off]); normal.push_back(nz[off]);
                const double w = omega[i] * omega[j];
                wJ.push_back(w * Jef[off]);                    // surface_integral.jl:4: ω[i]*ω[j]*Jac_face[i,j]
            }
    }
    if (ip1.empty()) return JX_OK;
    std::stable_sort(hits.begin(), hits.end(), [](const Hit &x, const Hit &y) { return x.node < y.node; });
    std::vector<int32_t> node, ptr, hit(hits.size());
    for (size_t h = 0; h < hits.size(); ++h) {
        if (h == 0 || hits[h].node != hits[h - 1].node) { node.push_back(hits[h].node); ptr.push_back((int32_t)h); }
        hit[h] = hits[h].w;
    }
    ptr.push_back((int32_t)hits.size());
    int rc;
    if ((rc = dalloc(c, &c->wf_ip1, ip1.size())) || (rc = dalloc(c, &c->wf_ipsfc, ipsfc.size())) || (rc = dalloc(c, &c->wf_node, node.size())) ||
        (rc = dalloc(c, &c->wf_ptr, ptr.size())) || (rc = dalloc(c, &c->wf_hit, hit.size())) || (rc = dalloc(c, &c->wf_normal, normal.size())) ||
        (rc = dalloc(c, &c->wf_wJ, wJ.size())) || (rc = dalloc(c, &c->wf_sface, ip1.size() * 4)))
        return rc;
    CK(cudaMemcpy(c->wf_ip1, ip1.data(), ip1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->wf_ipsfc, ipsfc.data(), ipsfc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->wf_node, node.data(), node.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->wf_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->wf_hit, hit.data(), hit.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->wf_normal, normal.data(), normal.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->wf_wJ, wJ.data(), wJ.size() * 8, cudaMemcpyHostToDevice));
    c->nwn = (int)ip1.size(); c->nwnode = (int)node.size();
    return JX_OK;
}

extern "C" int jx_upload_halo(jx_ctx *c, const int64_t *send_ptr, const int64_t *send_i, const int64_t *recv_ptr,
                              const int64_t *recv_idx, const int64_t *recvback_ptr, const int64_t *recvback_idx) {
    if (!c) return JX_EINVAL;
    if (!c->have_problem) return fail(c, JX_ESTATE, "jx_upload_halo before jx_set_problem");
    cudaSetDevice(c->device);
    free_halo(c);
    if (!send_ptr || !recv_ptr || !recvback_ptr) return fail(c, JX_EINVAL, "jx_upload_halo: null pointer array");
    if (c->massw && c->have_mesh && !c->mass_pending)
        return fail(c, JX_ESTATE, "the device-built mass matrix was already assembled without these lists: call jx_upload_halo "
                    "before the first evaluation / jx_get_minv / jx_condition_state");
    const int R = c->nranks;
    const int64_t ns = send_ptr[R], nr = recv_ptr[R];
    if (recvback_ptr[R] != ns) return fail(c, JX_EINVAL, "recvback lists must mirror the send lists");
    for (int r = 0; r < R; ++r) {
        if (send_ptr[r + 1] - send_ptr[r] != recvback_ptr[r + 1] - recvback_ptr[r])
            return fail(c, JX_EINVAL, "recvback list of peer %d differs in length from the send list", r);
        if (send_ptr[r + 1] > send_ptr[r]) c->send_seg.push_back({r, send_ptr[r], send_ptr[r + 1] - send_ptr[r]});
        if (recv_ptr[r + 1] > recv_ptr[r]) c->recv_seg.push_back({r, recv_ptr[r], recv_ptr[r + 1] - recv_ptr[r]});
    }
    const int64_t self_s = send_ptr[c->rank + 1] - send_ptr[c->rank], self_r = recv_ptr[c->rank + 1] - recv_ptr[c->rank];
    if (self_s != self_r) return fail(c, JX_EINVAL, "self send/recv lists differ in length");
    if (ns == 0 && nr == 0) return JX_OK;
    if (R > 1 && !c->comm && (ns != self_s || nr != self_r)) return fail(c, JX_ENCCL, "remote peers listed but no communicator");
    auto to0 = [&](const int64_t *src, int64_t n, std::vector<int64_t> &dst) -> bool {
        dst.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            dst[i] = src[i] - 1;
            if (dst[i] < 0 || dst[i] >= c->npoin) return false;
        }
        return true;
    };
    std::vector<int64_t> s0, r0, b0;
    if (!to0(send_i, ns, s0) || !to0(recv_idx, nr, r0) || !to0(recvback_idx, ns, b0))
        return fail(c, JX_EINVAL, "halo list holds ids outside 1..npoin");
    // owner-side adds keep list order per node: entries of one peer list are split into rounds
    // (round k = k-th occurrence of a node in that list); each round is conflict free.
    std::vector<int64_t> sel;
    c->add_rounds.resize(c->recv_seg.size());
    {
        std::vector<int32_t> occ_count((size_t)c->npoin, 0);
        for (size_t s = 0; s < c->recv_seg.size(); ++s) {
            const PeerSeg &sg = c->recv_seg[s];
            std::vector<int32_t> occ((size_t)sg.len);
            int32_t maxocc = 0;
            for (int64_t i = 0; i < sg.len; ++i) { occ[i] = occ_count[r0[sg.off + i]]++; maxocc = std::max(maxocc, occ[i]); }
            for (int64_t i = 0; i < sg.len; ++i) occ_count[r0[sg.off + i]] = 0;
            for (int32_t k = 0; k <= maxocc; ++k) {
                AddRound rd{(int64_t)sel.size(), 0};
                for (int64_t i = 0; i < sg.len; ++i)
                    if (occ[i] == k) sel.push_back(sg.off + i);
                rd.len = (int64_t)sel.size() - rd.off;
                c->add_rounds[s].push_back(rd);
            }
        }
    }
    c->nsend = ns; c->nrecv = nr;
    int rc;
    if ((rc = dalloc(c, &c->d_send_i, (size_t)ns)) || (rc = dalloc(c, &c->d_recv_idx, (size_t)nr)) ||
        (rc = dalloc(c, &c->d_recvback_idx, (size_t)ns)) || (rc = dalloc(c, &c->d_sendbuf, (size_t)ns * c->neqs)) ||
        (rc = dalloc(c, &c->d_recvbuf, (size_t)nr * c->neqs)) || (rc = dalloc(c, &c->d_add_sel, sel.size())))
        return rc;
    CK(cudaMemcpy(c->d_send_i, s0.data(), (size_t)ns * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_recv_idx, r0.data(), (size_t)nr * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_recvback_idx, b0.data(), (size_t)ns * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_add_sel, sel.data(), sel.size() * 8, cudaMemcpyHostToDevice));
    {   // unique nodes of the lists (direct accumulation sets their share of the low-storage register aside, rhs_core)
        std::vector<int64_t> un;
        un.reserve(s0.size() + r0.size() + b0.size());
        un.insert(un.end(), s0.begin(), s0.end()); un.insert(un.end(), r0.begin(), r0.end()); un.insert(un.end(), b0.begin(), b0.end());
        std::sort(un.begin(), un.end());
        un.erase(std::unique(un.begin(), un.end()), un.end());
        c->n_ifn = (int64_t)un.size();
        if ((rc = dalloc(c, &c->d_ifn, un.size())) || (rc = dalloc(c, &c->d_ifbase, un.size() * c->neqs))) return rc;
        CK(cudaMemcpy(c->d_ifn, un.data(), un.size() * 8, cudaMemcpyHostToDevice));
    }
    c->acc_ready = false;
    c->have_halo = true;
    return JX_OK;
}

extern "C" int jx_set_state(jx_ctx *c, const double *u) {
    if (!c || !u) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_set_state before jx_upload_mesh");
    cudaSetDevice(c->device);
    c->aux_fresh = false;
    CK(cudaMemcpyAsync(c->u, u, (size_t)c->npoin * c->neqs * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return JX_OK;
}
extern "C" int jx_get_state(jx_ctx *c, double *u) {
    if (!c || !u) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_get_state before jx_upload_mesh");
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(u, c->u, (size_t)c->npoin * c->neqs * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return JX_OK;
}
extern "C" int jx_get_du(jx_ctx *c, double *du) {
    if (!c || !du) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_get_du before jx_upload_mesh");
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(du, c->du, (size_t)c->npoin * c->neqs * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return JX_OK;
}

// ------------------------------------------------------------------------------------------
// the per-stage pipeline
// ------------------------------------------------------------------------------------------
namespace {

// assemble_mpi! (mpi_communications.jl:260-338) on the device: pack -> owners add in ascending
// sender rank, list order -> owners pack the sums -> send back -> non-owners overwrite.
int assemble(jx_ctx *c, double *a, cudaStream_t s, int ncomp) {
    const int m = ncomp > 0 ? ncomp : c->neqs;
    const int64_t N = c->npoin;
    if (c->nsend > 0) {
        k_pack<<<nblk(c->nsend * m, 256), 256, 0, s>>>(a, N, m, c->d_send_i, c->nsend, c->d_sendbuf);
        c->launches++;
    }
    auto exchange = [&](bool back) -> int {
        // forward: sendbuf segments -> owners' recvbuf segments; back: recvbuf segments -> sendbuf segments
        const std::vector<PeerSeg> &out = back ? c->recv_seg : c->send_seg;
        const std::vector<PeerSeg> &in = back ? c->send_seg : c->recv_seg;
        double *obuf = back ? c->d_recvbuf : c->d_sendbuf;
        double *ibuf = back ? c->d_sendbuf : c->d_recvbuf;
        bool remote = false;
        for (const PeerSeg &g : out) remote |= (g.peer != c->rank);
        for (const PeerSeg &g : in) remote |= (g.peer != c->rank);
        if (remote) {
            NCK(nccl().GroupStart());
            for (const PeerSeg &g : out)
                if (g.peer != c->rank) NCK(nccl().Send(obuf + g.off * m, (size_t)(g.len * m), kNcclFloat64, g.peer, c->comm, s));
            for (const PeerSeg &g : in)
                if (g.peer != c->rank) NCK(nccl().Recv(ibuf + g.off * m, (size_t)(g.len * m), kNcclFloat64, g.peer, c->comm, s));
            NCK(nccl().GroupEnd());
        }
        for (const PeerSeg &g : out)
            if (g.peer == c->rank)
                for (const PeerSeg &h : in)
                    if (h.peer == c->rank)   // the reference's MPI self-send (periodic twins on one rank)
                        CK(cudaMemcpyAsync(ibuf + h.off * m, obuf + g.off * m, (size_t)(g.len * m) * 8, cudaMemcpyDeviceToDevice, s));
        return JX_OK;
    };
    int rc = exchange(false);
    if (rc) return rc;
    for (size_t sgi = 0; sgi < c->recv_seg.size(); ++sgi)          // ascending sender rank
        for (const AddRound &rd : c->add_rounds[sgi]) {
            if (rd.len == 0) continue;
            k_add_sel<<<nblk(rd.len * m, 256), 256, 0, s>>>(a, N, m, c->d_recv_idx, c->d_add_sel + rd.off, rd.len, c->d_recvbuf);
            c->launches++;
        }
    if (c->nrecv > 0) {
        k_pack<<<nblk(c->nrecv * m, 256), 256, 0, s>>>(a, N, m, c->d_recv_idx, c->nrecv, c->d_recvbuf);
        c->launches++;
    }
    rc = exchange(true);
    if (rc) return rc;
    if (c->nsend > 0) {
        k_unpack<<<nblk(c->nsend * m, 256), 256, 0, s>>>(a, N, m, c->d_recvback_idx, c->nsend, c->d_sendbuf);
        c->launches++;
    }
    return JX_OK;
}

// Interface-first split ("boundary-elements-first kernel ordering to overlap with the interior kernel").
// Built lazily once mesh, halo lists and options are in place: marks the nodes of the assembler lists, flags every
// element group whose record names one of them, and lays the group ids out as [interface groups | interior groups].
// Only the list-driven team kernel (KernelSet::has_dyn) in atomics mode takes part; everything else keeps the
// single launch followed by the exchange.
int ensure_split(jx_ctx *c) {
    if (c->split_ready) return JX_OK;
    c->split_ready = true;
    c->n_iface = c->n_inner = 0;
    if (c->overlap_sms <= 0 || !c->have_halo || !c->have_mesh || !c->ks || !c->ks->has_dyn || c->dss_mode != 1) return JX_OK;
    if (c->nsend + c->nrecv == 0 || c->nelem == 0) return JX_OK;
    const KernelSet *ks = c->ks;
    const int64_t ngroups = (c->nelem + ks->elems_per_block - 1) / ks->elems_per_block;
    uint8_t *d_mask = nullptr, *d_flag = nullptr;
    int rc;
    if ((rc = dalloc(c, &d_mask, (size_t)c->npoin)) || (rc = dalloc(c, &d_flag, (size_t)ngroups))) { dfree(d_mask); return rc; }
    auto cleanup = [&]() { dfree(d_mask); dfree(d_flag); };
    cudaStream_t s = c->stream;
    cudaError_t e = cudaMemsetAsync(d_mask, 0, (size_t)c->npoin, s);
    if (c->nsend > 0) {
        k_mark_nodes<<<nblk(c->nsend, 256), 256, 0, s>>>(d_mask, c->d_send_i, c->nsend);
        k_mark_nodes<<<nblk(c->nsend, 256), 256, 0, s>>>(d_mask, c->d_recvback_idx, c->nsend);
    }
    if (c->nrecv > 0) k_mark_nodes<<<nblk(c->nrecv, 256), 256, 0, s>>>(d_mask, c->d_recv_idx, c->nrecv);
    k_flag_groups<<<nblk(ngroups, 128), 128, 0, s>>>(c->rec, ks->group_bytes, ks->fid_off, ks->elems_per_block * c->np, ngroups, d_mask, d_flag);
    c->launches += 4;
    std::vector<uint8_t> flag((size_t)ngroups);
    if (e == cudaSuccess) e = cudaMemcpyAsync(flag.data(), d_flag, (size_t)ngroups, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) return fail(c, JX_ECUDA, "interface split: %s", cudaGetErrorString(e));
    std::vector<int32_t> list;
    list.reserve((size_t)ngroups);
    for (int64_t g = 0; g < ngroups; ++g) if (flag[g]) list.push_back((int32_t)g);
    const int ni = (int)list.size();
    for (int64_t g = 0; g < ngroups; ++g) if (!flag[g]) list.push_back((int32_t)g);
    if (ni == 0 || ni == ngroups) return JX_OK;              // nothing to overlap with
    if ((rc = dalloc(c, &c->d_glist, (size_t)ngroups))) return rc;
    if (!c->d_gctr && (rc = dalloc(c, &c->d_gctr, 4))) return rc;
    CK(cudaMemcpy(c->d_glist, list.data(), (size_t)ngroups * 4, cudaMemcpyHostToDevice));
    if (!c->stream2) {
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, greatest));
        CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    c->n_iface = ni;
    c->n_inner = (int)(ngroups - ni);
    return JX_OK;
}

struct StageUpdate {
    int kind = 0;          // 0: du = Minv*RHS only; 1: 2N low-storage update of (u, tmp) behind it; 2: direct accumulation --
                           // the element kernels add into the low-storage register itself (k_stage_direct), no du
    double A = 0, B = 0, dt = 0, Anext = 0;
    int first = 0;
};

// kind 2 is possible with the unordered scatter on the context's own state when the kernel set has the per-node pre-pass
bool direct_ok(const jx_ctx *c) { return c->dss_mode == 1 && c->ks && c->ks->launch_aux && c->ks->launch_stage; }

// rhs!(du, u, params, t) on the device (rhs.jl:121-134, 498-711).  `u` is projected in place by
// the Dirichlet kernel; the mass-scaled result lands in `du`; with upd.kind == 1 the low-storage
// stage update is applied to (u, tmp) as well (fused into the DSS gather when no exchange follows).
int rhs_core(jx_ctx *c, double *u, double *du, const StageUpdate &upd) {
    if (c->mass_pending) {
        int rc = finalize_mass(c);
        if (rc) return rc;
    }
    if (!c->split_ready) {
        int rc = ensure_split(c);
        if (rc) return rc;
    }
    const KernelSet *ks = c->ks;
    cudaStream_t s = c->stream;
    const int64_t N = c->npoin, E = c->nelem;
    const int q = c->neqs;
    const bool atomics = c->dss_mode == 1;
    const bool direct = upd.kind == 2;                                   // caller checked direct_ok(c) and passes u == c->u
    double *acc = direct ? c->tmp : du;                                  // where the scatter lands
    // the previous stage's sweep (k_stage_direct) already evaluated aux on this state
    const bool fresh = c->aux_fresh && atomics && u == c->u && ks->launch_aux && c->aux;
    c->aux_fresh = false;
    if (c->nb > 0) {                                                     // rhs.jl:558
        PhaseScope ps(c, PH_BC);
        BcArgs b;
        b.u = u; b.qe = c->qe; b.node = c->bc_node; b.ptr = c->bc_ptr; b.normal = c->bc_normal; b.npoin = N; b.nb = c->nb;
        b.aux = fresh ? c->aux : nullptr; b.phys = c->phys;
        ks->launch_bc(b, s);
        c->launches++;
    }
    auto stage_update = [&]() {
        PhaseScope ps(c, PH_UPDATE);
        if (direct) {
            StageArgs sa;
            sa.u = u; sa.acc = acc; sa.qe = c->qe; sa.aux = c->aux; sa.npoin = N;
            sa.Bdt = upd.B * upd.dt; sa.Anext = upd.Anext; sa.phys = c->phys;
            ks->launch_stage(sa, 0, s);
            c->aux_fresh = true;
            c->acc_ready = true; c->acc_A = upd.Anext;
        } else {
            k_lsrk_update<<<nblk(N * q, 256), 256, 0, s>>>(u, c->tmp, du, N * q, upd.A, upd.B, upd.dt, upd.first);
            c->acc_ready = false;
        }
        c->launches++;
    };
    if (!atomics && (!c->rhs_el || (c->lvisc && !c->rhs_el_visc))) {
        int rc;
        if (!c->rhs_el && (rc = dalloc(c, &c->rhs_el, (size_t)E * c->np * q))) return rc;
        if (c->lvisc && !c->rhs_el_visc) {
            if ((rc = dalloc(c, &c->rhs_el_visc, (size_t)E * c->np * q))) return rc;
            CK(cudaMemsetAsync(c->rhs_el_visc, 0, (size_t)E * c->np * q * 8, s));   // equations with mu = 0 are never written by k_visc_quad
        }
    }
    ElemArgs ea;
    ea.u = u; ea.qe = c->qe; ea.rec = c->rec; ea.rhs_el = c->rhs_el; ea.rhs_el_visc = c->rhs_el_visc; ea.du = du;
    ea.Minv = c->Minv; ea.coords = c->coords; ea.elist = nullptr; ea.eorig = c->d_eorig; ea.nelem = E; ea.npoin = N;
    ea.atomics = atomics ? 1 : 0; ea.lsource = c->lsource; ea.phys = c->phys;
    for (int i = 0; i < 8; ++i) ea.visc[i] = c->visc[i];
    for (int i = 0; i < 64; ++i) ea.dpsi[i] = c->dpsi[i];
    ea.sgs = c->sgs;
    // atomics mode folds M^-1 into the scatter weight.  With an interface exchange the ranks then sum already
    // scaled partials (M^-1 is the globally assembled value, identical on every copy of a node): one rounding per
    // contribution instead of one per node, inside the tolerance the unordered mode has anyway, and no extra
    // pass over du.  The deterministic mode keeps the reference order (exchange, then divide_by_mass_matrix!).
    const bool fold_minv = atomics;
    ea.aux = nullptr;
    if (ks->launch_aux) {                                                // per-node flux ingredient (+ zero-fill of du)
        if (!c->aux || c->aux_doubles < (size_t)N * 4) {                 // up to 4 values per node
            int rc = dalloc(c, &c->aux, (size_t)N * 4);
            if (rc) return rc;
            c->aux_doubles = (size_t)N * 4;
        }
        if (!fresh) {
            PhaseScope ps(c, PH_AUX);
            AuxArgs aa;
            aa.u = u; aa.qe = c->qe; aa.aux = c->aux; aa.zero = (atomics && !direct) ? du : nullptr; aa.npoin = N; aa.phys = c->phys;
            ks->launch_aux(aa, (int)std::min<int64_t>((N + 255) / 256, (int64_t)c->num_sms * 8), s);
            c->launches++;
        } else if (atomics && !direct) {
            PhaseScope ps(c, PH_AUX);
            CK(cudaMemsetAsync(du, 0, (size_t)N * q * 8, s));
        }
        ea.aux = c->aux;
    } else if (atomics) {
        PhaseScope ps(c, PH_DSS);
        CK(cudaMemsetAsync(du, 0, (size_t)N * q * 8, s));
    }
    if (direct) {
        // the low-storage register must hold A_i S'_{i-1}: left so by the previous stage's sweep, else the accumulation
        // restarts from zero (first stage of a step, A_1 = 0; or a stand-alone stage of jx_bench_rhs)
        if (!(c->acc_ready && c->acc_A == upd.A)) {
            PhaseScope ps(c, PH_AUX);
            CK(cudaMemsetAsync(acc, 0, (size_t)N * q * 8, s));
        }
        c->acc_ready = false;
        if (c->have_halo && c->n_ifn > 0) {
            PhaseScope ps(c, PH_HALO);
            k_if_save_zero<<<nblk(c->n_ifn * q, 256), 256, 0, s>>>(acc, N, q, c->d_ifn, c->n_ifn, c->d_ifbase);
            c->launches++;
        }
    }
    ea.du = acc;
    auto restore_iface = [&]() {
        if (direct && c->have_halo && c->n_ifn > 0) {
            k_if_restore<<<nblk(c->n_ifn * q, 256), 256, 0, s>>>(acc, N, q, c->d_ifn, c->n_ifn, c->d_ifbase);
            c->launches++;
        }
    };
    ea.glist = nullptr; ea.gctr = nullptr; ea.nlist = 0; ea.reserve_sms = 0; ea.exit_ctr = nullptr; ea.exit_budget = 0;
    const bool bdy = c->nwn > 0;                                         // boundary fluxes: added before the exchange, rhs.jl:674-689
    auto bdy_fluxes = [&]() {
        PhaseScope ps(c, PH_BC);
        MostArgs ma;
        ma.u = u; ma.qe = c->qe; ma.coords = c->coords; ma.ip1 = c->wf_ip1; ma.ipsfc = c->wf_ipsfc; ma.normal = c->wf_normal;
        ma.wJ = c->wf_wJ; ma.sface = c->wf_sface; ma.npoin = N; ma.nwn = c->nwn; ma.lpert = c->lpert;
        ma.karman = c->most_c[0]; ma.z0_m = c->most_c[1]; ma.z0_h = c->most_c[2]; ma.cp = c->phys.v[4]; ma.g = c->phys.v[2];
        ma.delta_hf = c->delta_hf; ma.user_heatflux = c->user_heatflux;
        k_most_faces<<<nblk(c->nwn, 128), 128, 0, s>>>(ma);
        FluxAddArgs fa;
        fa.rhs = acc; fa.Minv = fold_minv ? c->Minv : nullptr; fa.sface = c->wf_sface; fa.node = c->wf_node; fa.ptr = c->wf_ptr;
        fa.hit = c->wf_hit; fa.npoin = N; fa.nnode = c->nwnode;
        k_bdy_flux_add<<<nblk(c->nwnode, 128), 128, 0, s>>>(fa);
        c->launches += 2;
    };
    const bool split = atomics && c->have_halo && c->split_ready && c->n_iface > 0 && ks->has_dyn && E > 0 && !bdy;
    if (split) {
        // interface groups first; their sums are final once that launch ends, so the exchange (second stream, high
        // priority) runs beside the launch over the interior groups, which leaves overlap_sms SMs to it
        const int per_sm = std::max(1, ks->max_blocks_per_sm());
        const int cap = c->num_sms * per_sm;
        CK(cudaMemsetAsync(c->d_gctr, 0, 4 * sizeof(int), s));
        // the AV viscous pass (k_visc_quad) walks the same lists behind each inviscid launch
        ViscArgs va;
        va.rec = c->rec_visc; va.out_el = c->rhs_el_visc; va.nv = 0;
        if (c->lvisc && ks->launch_visc)
            for (int e = 0; e < q && e < 8; ++e)
                if (c->visc[e] != 0.0) va.ve[va.nv++] = e;
        for (int i = va.nv; i < 8; ++i) va.ve[i] = 0;
        const int vper = va.nv > 0 ? std::max(1, ks->visc_max_blocks()) : 1;
        {
            PhaseScope ps(c, PH_ELEM);
            ea.glist = c->d_glist; ea.nlist = c->n_iface; ea.gctr = c->d_gctr; ea.reserve_sms = 0;
            ks->launch_elem(ea, std::min(c->n_iface, cap), s);
            if (va.nv > 0) {
                ks->launch_visc(ea, va, std::min(c->n_iface, c->num_sms * vper), s);
                c->launches++;
            }
            CK(cudaEventRecord(c->ev_fork, s));
            ea.glist = c->d_glist + c->n_iface; ea.nlist = c->n_inner; ea.gctr = c->d_gctr + 1;
            ea.reserve_sms = std::min(c->overlap_sms, c->num_sms / 2);
            // the launch carries reserve_sms * per_sm surplus CTAs; only that many may leave a reserved SM unworked
            ea.exit_ctr = c->d_gctr + 2; ea.exit_budget = ea.reserve_sms * per_sm;
            ks->launch_elem(ea, (int)std::min<int64_t>((int64_t)c->n_inner + (int64_t)ea.reserve_sms * per_sm, cap), s);
            if (va.nv > 0) {
                // the viscous pass walks its list with a static stride, so none of its CTAs may leave: it keeps no SM free
                // (the exchange has the whole inviscid interior launch to itself and competes for SMs only if it outlasts it)
                ea.reserve_sms = 0; ea.exit_ctr = nullptr; ea.exit_budget = 0;
                ks->launch_visc(ea, va, (int)std::min<int64_t>((int64_t)c->n_inner, (int64_t)c->num_sms * vper), s);
                c->launches++;
            }
            c->launches += 2;
        }
        CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        int rc = assemble(c, acc, c->stream2, 0);                           // DSS_global_RHS!, rhs.jl:690
        if (rc) return rc;
        CK(cudaEventRecord(c->ev_join, c->stream2));
        {
            PhaseScope ps(c, PH_HALO);                                   // what is left exposed after the interior launch
            CK(cudaStreamWaitEvent(s, c->ev_join, 0));
            restore_iface();
        }
        if (upd.kind != 0) stage_update();
        CK(cudaGetLastError());
        return JX_OK;
    }
    if (E > 0) {
        PhaseScope ps(c, PH_ELEM);
        const int64_t ngroups = (E + ks->elems_per_block - 1) / ks->elems_per_block;
        const int per_sm = std::max(1, ks->max_blocks_per_sm());
        const int grid = (int)std::min<int64_t>(ngroups, (int64_t)c->num_sms * per_sm);
        ks->launch_elem(ea, grid, s);                                    // rhs.jl:611, 659
        c->launches++;
        if (c->lvisc && ks->launch_visc) {                               // AV viscous term as a pass of its own, rhs.jl:659-672
            ViscArgs va;
            va.rec = c->rec_visc; va.out_el = c->rhs_el_visc; va.nv = 0;
            for (int e = 0; e < q && e < 8; ++e)
                if (c->visc[e] != 0.0) va.ve[va.nv++] = e;
            for (int i = va.nv; i < 8; ++i) va.ve[i] = 0;
            if (va.nv > 0) {
                const int vper = std::max(1, ks->visc_max_blocks());
                ks->launch_visc(ea, va, (int)std::min<int64_t>(ngroups, (int64_t)c->num_sms * vper), s);
                c->launches++;
            }
        }
    }
    if (!atomics) {
        PhaseScope ps(c, PH_DSS);
        GatherArgs g;
        g.rhs_el = c->rhs_el; g.rhs_el_visc = c->lvisc ? c->rhs_el_visc : nullptr; g.ptr = c->n2e_ptr; g.idx = c->n2e_idx;
        g.Minv = c->Minv; g.out = du; g.u = u; g.tmp = c->tmp; g.npoin = N; g.np = c->np; g.neqs = q;
        g.A = upd.A; g.B = upd.B; g.dt = upd.dt; g.first_stage = upd.first;
        g.mode = (c->have_halo || bdy) ? 0 : (upd.kind == 1 ? 2 : 1);   // rhs.jl:624, 671-672, 698-699
        ks->launch_gather(g, s);
        c->launches++;
        if (!c->have_halo && !bdy) { CK(cudaGetLastError()); return JX_OK; }
    }
    if (bdy) bdy_fluxes();                                               // apply_boundary_conditions_neumann!, rhs.jl:674-689
    if (c->have_halo) {                                                  // DSS_global_RHS!, rhs.jl:690
        PhaseScope ps(c, PH_HALO);
        int rc = assemble(c, acc, s, 0);
        if (rc) return rc;
        restore_iface();
    }
    if ((c->have_halo || bdy) && !fold_minv) {
        PhaseScope ps(c, PH_UPDATE);
        k_scale_minv<<<nblk(N * q, 256), 256, 0, s>>>(du, c->Minv, N, q);   // rhs.jl:698-699
        c->launches++;
    }
    if (upd.kind != 0) stage_update();
    CK(cudaGetLastError());
    return JX_OK;
}

int lincomb(jx_ctx *c, double *y, int form, int nterms, const double *const *x, const double *coef) {
    LinArgs a;
    a.y = y; a.nterms = nterms; a.n = c->npoin * c->neqs; a.form = form;
    for (int i = 0; i < 5; ++i) { a.x[i] = i < nterms ? x[i] : nullptr; a.c[i] = i < nterms ? coef[i] : 0.0; }
    if (a.n > 0) k_lincomb<<<nblk(a.n, 256), 256, 0, c->stream>>>(a);
    c->launches++;
    return JX_OK;
}

// Carpenter & Kennedy (1994) 2N-storage RK4(5) -- OrdinaryDiffEq CarpenterKennedy2N54
const double CK_A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                        -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
const double CK_B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                        1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                        2277821191437.0 / 14882151754819.0};

int ensure_scratch(jx_ctx *c, int n) {
    for (int i = 0; i < n; ++i)
        if (!c->ss[i]) {
            int rc = dalloc(c, &c->ss[i], (size_t)c->npoin * c->neqs);
            if (rc) return rc;
        }
    return JX_OK;
}

}  // namespace

extern "C" int jx_get_minv(jx_ctx *c, double *Minv) {
    if (!c || !Minv) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_get_minv before jx_upload_mesh");
    cudaSetDevice(c->device);
    int rc = finalize_mass(c);
    if (rc) return rc;
    CK(cudaMemcpyAsync(Minv, c->Minv, (size_t)c->npoin * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return JX_OK;
}

extern "C" int jx_condition_state(jx_ctx *c, int which) {
    if (!c || (which != 0 && which != 1)) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_condition_state before jx_upload_mesh");
    if (!c->massw) return fail(c, JX_ESTATE, "jx_condition_state needs the device-built mass weights: jx_upload_mesh_coords with Minv = NULL");
    cudaSetDevice(c->device);
    int rc = finalize_mass(c);
    if (rc) return rc;
    const int64_t N = c->npoin;
    const int q = c->neqs;
    double *x = which == 0 ? c->u : c->qe, *out = c->du;
    if (N > 0) k_mass_gather<<<nblk(N, 128), 128, 0, c->stream>>>(c->n2e_ptr, c->n2e_idx, c->massw, c->nelem, c->np, N, q, x, out);
    c->launches++;
    if (c->have_halo && (rc = assemble(c, out, c->stream, 0))) return rc;       // DSS_global_RHS!
    if (N > 0) k_scale_minv<<<nblk(N * q, 256), 256, 0, c->stream>>>(out, c->Minv, N, q);   // divide_by_mass_matrix!
    c->launches++;
    CK(cudaMemcpyAsync(x, out, (size_t)N * q * 8, cudaMemcpyDeviceToDevice, c->stream));
    c->aux_fresh = false;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return JX_OK;
}

extern "C" int jx_rhs(jx_ctx *c, double t, const double *u_host, double *du_host, double *u_back_host) {
    (void)t;   // none of the registered hooks depends on time
    if (!c) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_rhs before jx_upload_mesh");
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)c->npoin * c->neqs * 8;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (u_host) {
        c->aux_fresh = false;
        CK(cudaMemcpyAsync(c->u, u_host, bytes, cudaMemcpyHostToDevice, c->stream));
    }
    StageUpdate upd;
    int rc = rhs_core(c, c->u, c->du, upd);
    if (rc) return rc;
    if (du_host) CK(cudaMemcpyAsync(du_host, c->du, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (u_back_host) CK(cudaMemcpyAsync(u_back_host, c->u, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    if (u_host || du_host || u_back_host) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    }
    return JX_OK;
}

extern "C" int jx_rhs_dev(jx_ctx *c, double t, double *u_dev, double *du_dev) {
    (void)t;
    if (!c || !u_dev || !du_dev) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_rhs_dev before jx_upload_mesh");
    cudaSetDevice(c->device);
    cudaPointerAttributes pa, pb;
    if (cudaPointerGetAttributes(&pa, u_dev) != cudaSuccess || cudaPointerGetAttributes(&pb, du_dev) != cudaSuccess ||
        pa.type != cudaMemoryTypeDevice || pb.type != cudaMemoryTypeDevice || pa.device != c->device || pb.device != c->device) {
        cudaGetLastError();
        return fail(c, JX_EINVAL, "jx_rhs_dev: u and du must be device memory of device %d", c->device);
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    StageUpdate upd;
    int rc = rhs_core(c, u_dev, du_dev, upd);     // the Dirichlet projection writes u_dev in place, like rhs! does with u
    if (rc) return rc;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    return JX_OK;
}

extern "C" int jx_kernel_variant(jx_ctx *c) {
    if (!c || !c->ks) return JX_ESTATE;
    return c->ks->variant;
}

extern "C" int jx_step(jx_ctx *c, int scheme, double t, double dt, int nsteps) {
    (void)t;
    if (!c) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_step before jx_upload_mesh");
    if (nsteps < 0) return fail(c, JX_EINVAL, "nsteps < 0");
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)c->npoin * c->neqs * 8;
    int rc;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (scheme == JX_SCHEME_CK2N54) {
        auto one_step = [&]() -> int {
            for (int i = 0; i < 5; ++i) {
                StageUpdate upd;
                upd.kind = direct_ok(c) ? 2 : 1; upd.A = CK_A[i]; upd.B = CK_B[i]; upd.dt = dt; upd.first = (i == 0);
                upd.Anext = CK_A[(i + 1) % 5];
                const int r = rhs_core(c, c->u, c->du, upd);
                if (r) return r;
            }
            return JX_OK;
        };
        if (c->use_graph && nsteps > 2) {
            // the five stages of one step (about 20 launches, plus the NCCL groups at N > 1) are captured once and
            // replayed: small meshes are launch bound (C2: ~50 us of device work per stage)
            if ((rc = one_step())) return rc;                   // eager first step: lazy allocations stay outside the capture
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t gexec = nullptr;
            const int64_t l0 = c->launches;
            CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            rc = one_step();
            cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
            const int64_t per = c->launches - l0;
            c->launches = l0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(c, JX_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(c, JX_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
            for (int n = 1; n < nsteps; ++n) {
                ce = cudaGraphLaunch(gexec, c->stream);
                if (ce != cudaSuccess) { cudaGraphExecDestroy(gexec); return fail(c, JX_ECUDA, "cudaGraphLaunch: %s", cudaGetErrorString(ce)); }
            }
            c->launches += per * (nsteps - 1);
            ce = cudaStreamSynchronize(c->stream);
            cudaGraphExecDestroy(gexec);
            if (ce != cudaSuccess) return fail(c, JX_ECUDA, "jx_step graph replay: %s", cudaGetErrorString(ce));
        } else {
            for (int n = 0; n < nsteps; ++n)
                if ((rc = one_step())) return rc;
        }
    } else if (scheme == JX_SCHEME_SSPRK33) {
        // OrdinaryDiffEq SSPRK33 perform_step! (FSAL): k = f(uprev) is evaluated on the state itself
        if ((rc = ensure_scratch(c, 1))) return rc;
        double *up = c->ss[0], *u = c->u, *k = c->du;
        StageUpdate none;
        for (int n = 0; n < nsteps; ++n) {
            if ((rc = rhs_core(c, u, k, none))) return rc;
            CK(cudaMemcpyAsync(up, u, bytes, cudaMemcpyDeviceToDevice, c->stream));
            { const double *x[2] = {u, k}; double cf[2] = {1.0, dt}; lincomb(c, u, 0, 2, x, cf); }
            if ((rc = rhs_core(c, u, k, none))) return rc;
            { const double *x[3] = {up, u, k}; double cf[3] = {3.0, 1.0, dt}; lincomb(c, u, 1, 3, x, cf); }
            if ((rc = rhs_core(c, u, k, none))) return rc;
            { const double *x[3] = {up, u, k}; double cf[3] = {1.0, 2.0, 2 * dt}; lincomb(c, u, 2, 3, x, cf); }
        }
        if (nsteps > 0 && (rc = rhs_core(c, u, k, none))) return rc;   // trailing FSAL evaluation projects the final state
    } else if (scheme == JX_SCHEME_SSPRK54) {
        // Spiteri & Ruuth (2002) SSPRK(5,4) in OrdinaryDiffEq's update form
        if ((rc = ensure_scratch(c, 5))) return rc;
        const double b10 = 0.391752226571890, a20 = 0.444370493651235, a21 = 0.555629506348765, b21 = 0.368410593050371,
                     a30 = 0.620101851488403, a32 = 0.379898148511597, b32 = 0.251891774271694, a40 = 0.178079954393132,
                     a43 = 0.821920045606868, b43 = 0.544974750228521, a52 = 0.517231671970585, a53 = 0.096059710526147,
                     b53 = 0.063692468666290, a54 = 0.386708617503269, b54 = 0.226007483236906;
        double *up = c->ss[0], *u2 = c->ss[1], *u3 = c->ss[2], *u4 = c->ss[3], *k3 = c->ss[4], *u = c->u, *k = c->du;
        StageUpdate none;
        for (int n = 0; n < nsteps; ++n) {
            if ((rc = rhs_core(c, u, k, none))) return rc;
            CK(cudaMemcpyAsync(up, u, bytes, cudaMemcpyDeviceToDevice, c->stream));
            { const double *x[2] = {up, k}; double cf[2] = {1.0, b10 * dt}; lincomb(c, u2, 0, 2, x, cf); }
            if ((rc = rhs_core(c, u2, k, none))) return rc;
            { const double *x[3] = {up, u2, k}; double cf[3] = {a20, a21, b21 * dt}; lincomb(c, u2, 0, 3, x, cf); }
            if ((rc = rhs_core(c, u2, k, none))) return rc;
            { const double *x[3] = {up, u2, k}; double cf[3] = {a30, a32, b32 * dt}; lincomb(c, u3, 0, 3, x, cf); }
            if ((rc = rhs_core(c, u3, k3, none))) return rc;
            { const double *x[3] = {up, u3, k3}; double cf[3] = {a40, a43, b43 * dt}; lincomb(c, u4, 0, 3, x, cf); }
            if ((rc = rhs_core(c, u4, k, none))) return rc;
            { const double *x[5] = {u2, u3, k3, u4, k}; double cf[5] = {a52, a53, b53 * dt, a54, b54 * dt}; lincomb(c, u, 0, 5, x, cf); }
        }
        if (nsteps > 0 && (rc = rhs_core(c, u, k, none))) return rc;
    } else {
        return fail(c, JX_EINVAL, "unknown scheme %d", scheme);
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    return JX_OK;
}

extern "C" int jx_last_elapsed_ms(jx_ctx *c, float *ms) {
    if (!c || !ms) return JX_EINVAL;
    *ms = c->last_ms;
    return JX_OK;
}
extern "C" int64_t jx_launch_count(jx_ctx *c) { return c ? c->launches : 0; }
extern "C" int jx_sync(jx_ctx *c) {
    if (!c) return JX_EINVAL;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    return JX_OK;
}

extern "C" int jx_split_info(jx_ctx *c, int64_t *interface_groups, int64_t *interior_groups) {
    if (!c || !interface_groups || !interior_groups) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_split_info before jx_upload_mesh");
    cudaSetDevice(c->device);
    if (!c->split_ready) {
        int rc = ensure_split(c);
        if (rc) return rc;
    }
    const bool on = c->dss_mode == 1 && c->have_halo && c->n_iface > 0;
    *interface_groups = on ? c->n_iface : 0;
    *interior_groups = on ? c->n_inner : 0;
    return JX_OK;
}

extern "C" int jx_selftest(jx_ctx *c, int which, int64_t n, int64_t *failures) {
    if (!c || !failures || n < 0) return JX_EINVAL;
    if (which != 0) return fail(c, JX_EINVAL, "unknown self test %d", which);
    cudaSetDevice(c->device);
    unsigned long long *d_bad = nullptr, h_bad = 0;
    CK(cudaMalloc((void **)&d_bad, 8));
    CK(cudaMemsetAsync(d_bad, 0, 8, c->stream));
    if (n > 0) k_selftest_div<<<c->num_sms * 8, 256, 0, c->stream>>>(n, 0x1234abcdull, d_bad);
    c->launches++;
    cudaError_t e = cudaMemcpyAsync(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_bad);
    if (e != cudaSuccess) return fail(c, JX_ECUDA, "jx_selftest: %s", cudaGetErrorString(e));
    *failures = (int64_t)h_bad;
    return JX_OK;
}

extern "C" int jx_bench_rhs(jx_ctx *c, int n, int fused_stage, float *total_ms, float *phase_ms) {
    if (!c || n <= 0) return JX_EINVAL;
    if (!c->have_mesh) return fail(c, JX_ESTATE, "jx_bench_rhs before jx_upload_mesh");
    cudaSetDevice(c->device);
    for (auto e : c->ph_ev) c->ph_pool.push_back(e);
    c->ph_ev.clear(); c->ph_tag.clear();
    c->phase_timing = phase_ms != nullptr;
    StageUpdate upd;
    if (fused_stage) {   // a dt = 0 CK2N54 stage: full stage traffic, state unchanged
        upd.kind = direct_ok(c) ? 2 : 1; upd.A = CK_A[1]; upd.B = CK_B[1]; upd.dt = 0.0; upd.first = 0; upd.Anext = CK_A[1];
    }
    int rc = JX_OK;
    // CUDA graph path (no per-phase events): one RHS evaluation -- kernels, NCCL send/recv groups, copies -- is
    // captured once and replayed n times, so the host's enqueue cost (dominant at N > 1: two NCCL groups and six
    // small kernels per evaluation) leaves the timed region.
    cudaGraphExec_t gexec = nullptr;
    if (c->use_graph && !c->phase_timing) {
        rc = rhs_core(c, c->u, c->du, upd);                 // eager once: lazy allocations happen outside the capture
        if (rc) return rc;
        cudaGraph_t graph = nullptr;
        const int64_t l0 = c->launches;
        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        rc = rhs_core(c, c->u, c->du, upd);
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        const int64_t per = c->launches - l0;
        c->launches = l0;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(c, JX_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail(c, JX_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < n; ++i) {
            ce = cudaGraphLaunch(gexec, c->stream);
            if (ce != cudaSuccess) { cudaGraphExecDestroy(gexec); return fail(c, JX_ECUDA, "cudaGraphLaunch: %s", cudaGetErrorString(ce)); }
        }
        c->launches += per * n;
    } else {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < n && rc == JX_OK; ++i) rc = rhs_core(c, c->u, c->du, upd);
    }
    c->phase_timing = false;
    if (rc) return rc;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    if (gexec) cudaGraphExecDestroy(gexec);
    if (total_ms) *total_ms = c->last_ms;
    if (phase_ms) {
        for (int i = 0; i < PH_COUNT; ++i) phase_ms[i] = 0.f;
        for (size_t k = 0; k < c->ph_tag.size(); ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->ph_ev[2 * k], c->ph_ev[2 * k + 1]);
            phase_ms[c->ph_tag[k]] += ms;
        }
    }
    return JX_OK;
}
