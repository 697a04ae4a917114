// jx_kernels.cuh -- hand-written sm_100a FP64 kernels of the explicit RHS path.
//
// Data layout in HBM (built once by jx_upload_mesh, resident across stages):
//   u, du, tmp        double[neqs][npoin]      (the reference's flat ODE vector, rhs.jl:29-47)
//   qe                double[neqs+1][npoin]
//   Minv              double[npoin]
//   element records   per element: double met[NMET][NP] (metric terms; slot NMET-1 holds
//                     ωJac = ω_i*(ω_j*ω_k)*Je, the expression of rhs.jl:1641-1643) followed by
//                     int32 conn[NPP] (0-based node ids, padded) -- one contiguous, 16-byte
//                     aligned block fetched with ONE cp.async.bulk (TMA) per element group
//   rhs_el            double[nelem][neqs][NP]  (deterministic DSS mode only)
//   n2e_ptr / n2e_idx CSR node -> (element*NP + local), ascending element = DSS_rhs! order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "jx_functors.cuh"

namespace jx {

constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }
constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <int NSD, int NGL>
struct Geo {
    static constexpr int NP = ipow(NGL, NSD);
    static constexpr int NMET = NSD * NSD + 1;
    static constexpr int NPP = round_up(NP, 4);                               // int32 conn padding
    static constexpr int REC_BYTES = round_up(NMET * NP * 8 + NPP * 4, 16);   // 16 B multiple for TMA bulk
    static constexpr int REC_DOUBLES = REC_BYTES / 8;
};

// ------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// setup kernels (run once at upload)
// ------------------------------------------------------------------------------------------
struct RetileArgs {
    const double *src;        // one metric array [E, n, n, n|1], element fastest (device copy)
    const double *omega;
    const int64_t *connijk;   // [E, n, n, n|1] 1-based (slot < 0 only)
    char *rec;
    int64_t nelem;
    int nsd, ngl, np, nmet, npp, rec_bytes;
    int slot;                 // metric slot; nmet-1 = Je (stored as ωJac); -1 = connectivity
};

// element-fastest Julia arrays [E, n, n, n] -> per-element records; thread = (element, local node)
static __global__ void k_retile(RetileArgs a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = a.nelem * a.np;
    if (tid >= total) return;
    const int64_t iel = tid % a.nelem;     // element fastest: coalesced reads
    const int l = (int)(tid / a.nelem);
    const int n = a.ngl;
    double *rm = reinterpret_cast<double *>(a.rec + (size_t)iel * a.rec_bytes);
    int32_t *rc = reinterpret_cast<int32_t *>(a.rec + (size_t)iel * a.rec_bytes + (size_t)a.nmet * a.np * 8);
    const size_t src = (size_t)iel + (size_t)a.nelem * l;
    if (a.slot < 0) {
        rc[l] = (int32_t)(a.connijk[src] - 1);
        if (l == 0)
            for (int p = a.np; p < a.npp; ++p) rc[p] = 0;
    } else if (a.slot < a.nmet - 1) {
        rm[a.slot * a.np + l] = a.src[src];
    } else {
        const double Je = a.src[src];
        double wJ;
        if (a.nsd == 3) {
            const int i = l % n, j = (l / n) % n, k = l / (n * n);
            const double wjk = a.omega[j] * a.omega[k];      // rhs.jl:1636-1643
            wJ = a.omega[i] * wjk * Je;
        } else {
            const int i = l % n, j = l / n;
            wJ = a.omega[i] * a.omega[j] * Je;               // rhs.jl:1515-1516
        }
        rm[(a.nmet - 1) * a.np + l] = wJ;
    }
}

static __global__ void k_count_valence(const int64_t *connijk, int64_t total, int32_t *cnt) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < total) atomicAdd(&cnt[connijk[tid] - 1], 1);
}

// fill CSR lists (unordered), thread = (element, local node) with element fastest
static __global__ void k_fill_n2e(const int64_t *connijk, int64_t nelem, int np, const int64_t *ptr, int32_t *cursor,
                           uint32_t *idx) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nelem * np) return;
    const int64_t iel = tid % nelem;
    const int l = (int)(tid / nelem);
    const int64_t ip = connijk[tid] - 1;
    const int pos = atomicAdd(&cursor[ip], 1);
    idx[ptr[ip] + pos] = (uint32_t)(iel * np + l);
}

// sort every node's list ascending (element-major key = DSS_rhs! order, element_matrices.jl:905-916)
static __global__ void k_sort_n2e(int64_t npoin, const int64_t *ptr, uint32_t *idx) {
    const int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npoin) return;
    const int64_t b = ptr[ip], e = ptr[ip + 1];
    for (int64_t i = b + 1; i < e; ++i) {
        const uint32_t v = idx[i];
        int64_t j = i - 1;
        while (j >= b && idx[j] > v) { idx[j + 1] = idx[j]; --j; }
        idx[j + 1] = v;
    }
}

// ------------------------------------------------------------------------------------------
// Dirichlet boundary projection.  BCs.jl:610-652 (3D) / :183-280 (2D) + uaux2u!.
// The reference visits boundary faces in ascending order and applies the hook to the *current*
// value, so nodes on box edges/corners are projected by each adjacent face in turn.  One thread
// owns one unique boundary node and replays that node's (face-ordered) hit list sequentially:
// same result, no race (the reference's own KA kernel races here, rhs_gpu.jl:699-701).
// ------------------------------------------------------------------------------------------
struct BcArgs {
    double *u;
    const double *qe;
    const int32_t *node;     // [nb] unique boundary nodes (0-based)
    const int32_t *ptr;      // [nb+1]
    const double *normal;    // [nhits][3]
    int64_t npoin;
    int nb;
    double *aux = nullptr;   // non-null: the per-node flux ingredient of the projected nodes is re-evaluated here (the fused
                             // stage update k_stage_fused produced aux from the unprojected state)
    Phys phys;
};

template <class EQ>
static __global__ void k_bc_dirichlet(BcArgs a) {
    constexpr int NEQ = EQ::NEQ;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nb) return;
    const int64_t ip = a.node[t];
    double q[NEQ], qe[NEQ + 1], qbdy[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) q[e] = a.u[(size_t)e * a.npoin + ip];
#pragma unroll
    for (int e = 0; e <= NEQ; ++e) qe[e] = EQ::NEEDS_QE ? a.qe[(size_t)e * a.npoin + ip] : 0.0;
    for (int h = a.ptr[t]; h < a.ptr[t + 1]; ++h) {
#pragma unroll
        for (int e = 0; e < NEQ; ++e) qbdy[e] = 4325789.0;
        EQ::bc_dirichlet(q, qe, a.normal[3 * h], a.normal[3 * h + 1], a.normal[3 * h + 2], qbdy);
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
            if (!AlmostEqual(qbdy[e], q[e]) && !AlmostEqual(qbdy[e], 4325789.0)) q[e] = qbdy[e];
    }
#pragma unroll
    for (int e = 0; e < NEQ; ++e) a.u[(size_t)e * a.npoin + ip] = q[e];
    if constexpr (EQ::HAS_AUX) {
        if (a.aux) {
            double ax[EQ::NAUX > 0 ? EQ::NAUX : 1];
            EQ::aux(a.phys, q, qe, ax);
#pragma unroll
            for (int x = 0; x < EQ::NAUX; ++x) a.aux[(size_t)x * a.npoin + ip] = ax[x];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fused per-element kernel: flux/source/primitives at the LGL nodes, sum-factorised inviscid
// divergence (rhs.jl:1615-1698 3D, :1501-1542 2D) and AV viscous term (rhs.jl:2794-2867 3D,
// :1973-2056 2D).  Variant "node": one thread per element node, EPB elements per CTA iteration,
// element records double-buffered through shared memory by TMA bulk copies.
// ------------------------------------------------------------------------------------------
// SGS viscosity: the scalar content of the reference's SGS_SMAG / SGS_VREM structs (sgsStructs.jl:6-72, filled by
// allocate_SGS :77-120 and params_setup.jl:249-253).  The per-node caches of the reference (sgs.μ_turb ...) are element-local in
// effect -- compute_sgs_cache! refills them for every element before its equation loop -- so here μ_turb lives in a register.
struct SgsArgs {
    int model = 0;                     // 0: AV (sgs === nothing); 1: SMAG; 2: VREM
    int lrichardson = 0, ltheta_eqn = 1;
    double delta = 0.0;                // mesh.Δeffective_l
    double Pr_t = 0, Sc_t = 0, mu_mol = 0, kappa_mol = 0, Ri_crit = 0, C_s2 = 0, C_vrem = 0, g = 0;
    const int32_t *ad_lvl = nullptr;   // mesh.ad_lvl [nelem] (3D: Δ_effective = ldexp(Δ, -ad_lvl), rhs.jl:1416) or nullptr
};

struct ElemArgs {
    const double *u;
    const double *qe;
    const char *rec;
    double *rhs_el;        // deterministic mode: [E][NEQ][NP]
    double *rhs_el_visc;   // deterministic mode, viscous part kept separate (DSS'ed separately, rhs.jl:671-672)
    double *du;            // atomics mode target
    const double *Minv;
    const double *coords;  // [nsd][npoin] (only read by functors with NEEDS_XYZ)
    const double *aux;     // [NAUX][npoin] per-node part of the flux (k_node_aux), kernels with EQ::HAS_AUX only
    const int32_t *elist;  // optional element subset (interface / interior split); nullptr = all
    const int32_t *eorig;  // team kernels: record position -> element id (the records are laid out in order_elements() order;
                           // rhs_el keeps the caller's element numbering, which the deterministic gather walks); nullptr = identity
    const int32_t *glist;  // k_elem_team<DYN>: list of element groups this launch processes (interface or interior set)
    int *gctr;             // k_elem_team<DYN>: work counter (zeroed before the launch); CTAs take list positions from it
    int nlist;             // length of glist
    int reserve_sms;       // k_elem_team<DYN>: CTAs landing on SMs with %smid < reserve_sms exit at once (SMs kept free for
                           // the interface exchange running beside the interior launch)
    int *exit_ctr;         // k_elem_team<DYN>: how many CTAs have left a reserved SM so far (zeroed before the launch)
    int exit_budget;       // ... and how many may: the surplus CTAs of the launch.  Every CTA beyond it works wherever it lands,
                           // so the list is always covered, whatever the block scheduler does with the other SMs
    int64_t nelem, npoin;  // nelem = number of elements this launch processes
    int atomics;
    int lsource;
    Phys phys;
    double visc[8];
    double dpsi[64];       // Julia dψ[m,i] column-major: dpsi[m + NGL*i]
    SgsArgs sgs;           // k_elem_node<..., VISC = 2> only
};

// functors that add the viscous-work term to one equation of the 2D AV pass declare TAU_U_EQ
template <class EQ, class = void>
struct has_tau_u : std::false_type {};
template <class EQ>
struct has_tau_u<EQ, std::void_t<decltype(EQ::TAU_U_EQ)>> : std::true_type {};
template <class EQ, bool = has_tau_u<EQ>::value>
struct tau_u_eq { static constexpr int value = -1; };
template <class EQ>
struct tau_u_eq<EQ, true> { static constexpr int value = EQ::TAU_U_EQ; };

// Richardson stability function, SGS.jl:1241-1253
__device__ __forceinline__ double sgs_f_Ri(const SgsArgs &sg, double N2_val, double Sij2_val) {
    const double Ri = Sij2_val > 1e-12 ? N2_val / Sij2_val : 0.0;
    if (Ri >= sg.Ri_crit) return 0.0;
    if (Ri >= 0.0) {
        const double ratio = Ri / sg.Ri_crit;
        return (1.0 - ratio) * (1.0 - ratio);
    }
    return fmin(sqrt(1.0 - 16.0 * Ri), 3.0);
}

// cache-reading SGS_diffusion (SGS.jl:1087-1109 3D, :1663-1685 2D); e 0-based, IT = the temperature / energy equation
template <int IT>
__device__ __forceinline__ double sgs_diffusion(const SgsArgs &sg, const double *visc, int e, double rho, double mu_turb) {
    if (e >= 1 && e < IT) return (sg.mu_mol + mu_turb) * visc[e];
    if (e == IT) {
        const double k_turb = mu_turb / (rho * sg.Pr_t);
        if (sg.ltheta_eqn) return k_turb * visc[e];
        return (sg.kappa_mol + k_turb) * visc[e];
    }
    const double k_turb_scalar = mu_turb / (rho * sg.Sc_t);
    return (sg.kappa_mol + k_turb_scalar) * visc[e];
}

// Node-local step of the SGS viscous pass for the node (i,j,k) = l of one element: compute_sgs_cache! (SGS.jl:1118-1408 3D,
// :1416-1655 2D; dry, micro == 1) -- gradients accumulated with separately rounded multiply and add, as the plain Julia loop
// does -- then the cache-reading _expansion_visc! (rhs.jl:2582-2785 3D, :2275-2400 2D) -- gradients as FMA chains (@turbo) --
// for every equation.  U: primitives [NEQ][NP] of the element in shared memory; Gv: [NEQ][NSD][NP] weak-form fluxes out.
template <int NSD, int NGL, int NEQ>
__device__ __forceinline__ void sgs_node_fluxes(const SgsArgs &sg, double D2, const double *visc, const double *sD, const double *U,
                                                double *Gv, const double *mt, double wJ, int i, int j, int k, int l) {
    static_assert(NEQ >= NSD + 2, "SGS closures need (rho, velocity, temperature)");
    constexpr int NP = Geo<NSD, NGL>::NP;
    constexpr int IT = NSD + 1;
    int off[NSD];          // first node of the xi-, eta- (zeta-) line through l, with strides 1, NGL, NGL^2
    off[0] = NGL * (j + NGL * k);
    off[1] = i + NGL * NGL * k;
    if constexpr (NSD == 3) off[2] = i + NGL * j;
    const int str[3] = {1, NGL, NGL * NGL};
    const int ijk[3] = {i, j, k};
    // --- compute_sgs_cache!: velocity and temperature gradients, a += b*c separately rounded
    double gc[NSD][NSD], gt[NSD];
#pragma unroll
    for (int c = 0; c < NSD; ++c) {
        gt[c] = 0.0;
#pragma unroll
        for (int a = 0; a < NSD; ++a) gc[c][a] = 0.0;
    }
#pragma unroll
    for (int m = 0; m < NGL; ++m) {
#pragma unroll
        for (int a = 0; a < NSD; ++a) {
            const double dm = sD[m + NGL * ijk[a]];
            const int ln = off[a] + str[a] * m;
#pragma unroll
            for (int c = 0; c < NSD; ++c) gc[c][a] = gc[c][a] + dm * U[(1 + c) * NP + ln];
            gt[a] = gt[a] + dm * U[IT * NP + ln];
        }
    }
    double Dc[NSD][NSD];   // Dc[c][b] = d u_c / d x_b
#pragma unroll
    for (int c = 0; c < NSD; ++c)
#pragma unroll
        for (int b = 0; b < NSD; ++b) {
            if constexpr (NSD == 3) Dc[c][b] = gc[c][0] * mt[b] + gc[c][1] * mt[3 + b] + gc[c][2] * mt[6 + b];
            else Dc[c][b] = gc[c][0] * mt[b] + gc[c][1] * mt[2 + b];
        }
    double Sij2_val;
    if constexpr (NSD == 3) {
        const double S12 = 0.5 * (Dc[0][1] + Dc[1][0]), S13 = 0.5 * (Dc[0][2] + Dc[2][0]), S23 = 0.5 * (Dc[1][2] + Dc[2][1]);
        const double SS = Dc[0][0] * Dc[0][0] + Dc[1][1] * Dc[1][1] + Dc[2][2] * Dc[2][2] + 2.0 * (S12 * S12 + S13 * S13 + S23 * S23);
        Sij2_val = 2.0 * SS;
    } else {
        const double S12 = 0.5 * (Dc[0][1] + Dc[1][0]);
        const double SS = Dc[0][0] * Dc[0][0] + Dc[1][1] * Dc[1][1] + 2.0 * S12 * S12;
        Sij2_val = 2.0 * SS;
    }
    double f_Ri_val = 1.0;
    if (sg.lrichardson) {
        const double th = U[IT * NP + l];
        double dtdz;                                   // vertical: z in 3D, y in 2D
        if constexpr (NSD == 3) dtdz = gt[0] * mt[2] + gt[1] * mt[5] + gt[2] * mt[8];
        else dtdz = gt[0] * mt[1] + gt[1] * mt[3];
        const double N2_val = fabs(th) > 1e-12 ? (sg.g / th) * dtdz : 0.0;
        f_Ri_val = sgs_f_Ri(sg, N2_val, Sij2_val);
    }
    const double rho = U[l];
    double mu_turb;
    if (sg.model == 1) {
        mu_turb = rho * sg.C_s2 * D2 * sqrt(Sij2_val) * f_Ri_val;
    } else {
        double B_b, uu;
        if constexpr (NSD == 3) {
            const double b11 = D2 * (Dc[0][0] * Dc[0][0] + Dc[0][1] * Dc[0][1] + Dc[0][2] * Dc[0][2]);
            const double b12 = D2 * (Dc[0][0] * Dc[1][0] + Dc[0][1] * Dc[1][1] + Dc[0][2] * Dc[1][2]);
            const double b13 = D2 * (Dc[0][0] * Dc[2][0] + Dc[0][1] * Dc[2][1] + Dc[0][2] * Dc[2][2]);
            const double b22 = D2 * (Dc[1][0] * Dc[1][0] + Dc[1][1] * Dc[1][1] + Dc[1][2] * Dc[1][2]);
            const double b23 = D2 * (Dc[1][0] * Dc[2][0] + Dc[1][1] * Dc[2][1] + Dc[1][2] * Dc[2][2]);
            const double b33 = D2 * (Dc[2][0] * Dc[2][0] + Dc[2][1] * Dc[2][1] + Dc[2][2] * Dc[2][2]);
            B_b = b11 * b22 + b11 * b33 + b22 * b33 - (b12 * b12 + b13 * b13 + b23 * b23);
            uu = Dc[0][0] * Dc[0][0] + Dc[0][1] * Dc[0][1] + Dc[0][2] * Dc[0][2] + Dc[1][0] * Dc[1][0] + Dc[1][1] * Dc[1][1] +
                 Dc[1][2] * Dc[1][2] + Dc[2][0] * Dc[2][0] + Dc[2][1] * Dc[2][1] + Dc[2][2] * Dc[2][2];
        } else {
            const double b11 = D2 * (Dc[0][0] * Dc[0][0] + Dc[0][1] * Dc[0][1]);
            const double b12 = D2 * (Dc[0][0] * Dc[1][0] + Dc[0][1] * Dc[1][1]);
            const double b22 = D2 * (Dc[1][0] * Dc[1][0] + Dc[1][1] * Dc[1][1]);
            B_b = b11 * b22 - b12 * b12;
            uu = Dc[0][0] * Dc[0][0] + Dc[0][1] * Dc[0][1] + Dc[1][0] * Dc[1][0] + Dc[1][1] * Dc[1][1];
        }
        const double mu_base = (uu > 2.220446049250313e-16 && B_b > 0.0) ? rho * sg.C_vrem * sqrt(B_b / uu) : 0.0;
        mu_turb = mu_base * f_Ri_val;
    }
    // --- cache-reading _expansion_visc!: velocity gradients again, this time as the @turbo FMA chains
#pragma unroll
    for (int c = 0; c < NSD; ++c)
#pragma unroll
        for (int a = 0; a < NSD; ++a) gc[c][a] = 0.0;
#pragma unroll
    for (int m = 0; m < NGL; ++m) {
#pragma unroll
        for (int a = 0; a < NSD; ++a) {
            const double dm = sD[m + NGL * ijk[a]];
            const int ln = off[a] + str[a] * m;
#pragma unroll
            for (int c = 0; c < NSD; ++c) gc[c][a] = fma(dm, U[(1 + c) * NP + ln], gc[c][a]);
        }
    }
#pragma unroll
    for (int c = 0; c < NSD; ++c)
#pragma unroll
        for (int b = 0; b < NSD; ++b) {
            if constexpr (NSD == 3) Dc[c][b] = gc[c][0] * mt[b] + gc[c][1] * mt[3 + b] + gc[c][2] * mt[6 + b];
            else Dc[c][b] = gc[c][0] * mt[b] + gc[c][1] * mt[2 + b];
        }
    double div_u;
    if constexpr (NSD == 3) div_u = Dc[0][0] + Dc[1][1] + Dc[2][2];
    else div_u = Dc[0][0] + Dc[1][1];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) {
        double flux[NSD];
        if (e >= 1 && e <= NSD) {
            const int c = e - 1;
            const double ev = sgs_diffusion<IT>(sg, visc, e, rho, mu_turb);
#pragma unroll
            for (int b = 0; b < NSD; ++b) {
                if (b == c) flux[b] = 2.0 * ev * Dc[c][c] - (2.0 / 3.0) * ev * div_u;
                else flux[b] = ev * (Dc[c < b ? c : b][c < b ? b : c] + Dc[c < b ? b : c][c < b ? c : b]);
            }
        } else {
            double gs[NSD];
#pragma unroll
            for (int a = 0; a < NSD; ++a) gs[a] = 0.0;
#pragma unroll
            for (int m = 0; m < NGL; ++m) {
#pragma unroll
                for (int a = 0; a < NSD; ++a) gs[a] = fma(sD[m + NGL * ijk[a]], U[e * NP + off[a] + str[a] * m], gs[a]);
            }
            const double ed = sgs_diffusion<IT>(sg, visc, e, rho, mu_turb);
#pragma unroll
            for (int b = 0; b < NSD; ++b) {
                double dsdx;
                if constexpr (NSD == 3) dsdx = gs[0] * mt[b] + gs[1] * mt[3 + b] + gs[2] * mt[6 + b];
                else dsdx = gs[0] * mt[b] + gs[1] * mt[2 + b];
                flux[b] = ed * dsdx;
            }
            if constexpr (NSD == 2) {
                if (e == IT && !sg.ltheta_eqn) {      // total-energy equation: viscous work with the momentum viscosity, rhs.jl:2361-2370
                    const double ev = sgs_diffusion<IT>(sg, visc, 1, rho, mu_turb);
                    const double txx = 2.0 * ev * Dc[0][0] - (2.0 / 3.0) * ev * div_u;
                    const double tyy = 2.0 * ev * Dc[1][1] - (2.0 / 3.0) * ev * div_u;
                    const double txy = ev * (Dc[0][1] + Dc[1][0]);
                    const double ul = U[1 * NP + l], vl = U[2 * NP + l];
                    flux[0] = flux[0] + (txx * ul + txy * vl);
                    flux[1] = flux[1] + (txy * ul + tyy * vl);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NSD; ++a) {
            double s;
            if constexpr (NSD == 3) s = mt[3 * a] * flux[0] + mt[3 * a + 1] * flux[1] + mt[3 * a + 2] * flux[2];
            else s = mt[2 * a] * flux[0] + mt[2 * a + 1] * flux[1];
            Gv[(e * NSD + a) * NP + l] = s * wJ;       // sigma_mu = 1.0 (rhs.jl:2615): the further factor is exact
        }
    }
}

// VISC: 0 = inviscid only; 1 = AV (constant coefficients, sgs === nothing); 2 = SGS closure (SMAG / VREM, ElemArgs::sgs)
template <int NSD, int NGL, class EQ, int VISC, int EPB>
struct ElemNodeCfg {
    using G = Geo<NSD, NGL>;
    static constexpr int NEQ = EQ::NEQ;
    static constexpr int NP = G::NP;
    static constexpr int NT = round_up(EPB * NP, 32);
    static constexpr int REC_D = G::REC_DOUBLES;
    static constexpr int SM_REC = 2 * EPB * REC_D;
    static constexpr int SM_FLUX = EPB * NSD * NEQ * NP;
    static constexpr int SM_PRIM = VISC ? EPB * NEQ * NP : 0;
    static constexpr int SM_GV = VISC ? EPB * NEQ * NSD * NP : 0;
    static constexpr int SM_D = NGL * NGL;
    static constexpr size_t SMEM_BYTES = (size_t)(SM_REC + SM_FLUX + SM_PRIM + SM_GV + SM_D) * 8 + 16;
};

template <int NSD, int NGL, class EQ, int VISC, int EPB>
static __global__ void __launch_bounds__(ElemNodeCfg<NSD, NGL, EQ, VISC, EPB>::NT, (VISC == 2 && ElemNodeCfg<NSD, NGL, EQ, VISC, EPB>::NT <= 256) ? 2 : 1)
k_elem_node(const __grid_constant__ ElemArgs a) {
    using C = ElemNodeCfg<NSD, NGL, EQ, VISC, EPB>;
    using G = Geo<NSD, NGL>;
    constexpr int NEQ = C::NEQ, NP = C::NP, NMET = G::NMET, REC_D = C::REC_D, REC_BYTES = G::REC_BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sRec = reinterpret_cast<double *>(smem_raw);
    double *sFl = sRec + C::SM_REC;
    double *sU = sFl + C::SM_FLUX;
    double *sGv = sU + C::SM_PRIM;
    double *sD = sGv + C::SM_GV;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sD + C::SM_D);

    const int t = threadIdx.x;
    const bool active = t < EPB * NP;
    const int slot = active ? t / NP : 0;
    const int l = active ? t % NP : 0;
    const int i = l % NGL, j = (l / NGL) % NGL, k = (NSD == 3) ? l / (NGL * NGL) : 0;

    for (int x = t; x < NGL * NGL; x += blockDim.x) sD[x] = a.dpsi[x];
    if (t == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_proxy_async();
    }
    __syncthreads();

    const int64_t ngroups = (a.nelem + EPB - 1) / EPB;
    int64_t g = blockIdx.x;
    uint32_t phase[2] = {0, 0};
    int buf = 0;
    // one thread issues the TMA bulk copies of a group: contiguous records in one copy, or one
    // copy per element when an element list is given
    auto issue = [&](int64_t grp, int b) {
        const int64_t e0 = grp * EPB;
        const int cnt = (int)(a.nelem - e0 < EPB ? a.nelem - e0 : EPB);
        mbar_expect_tx(&bar[b], (uint32_t)(cnt * REC_BYTES));
        double *dst = sRec + (size_t)b * EPB * REC_D;
        if (a.elist == nullptr) {
            tma_bulk_g2s(dst, a.rec + (size_t)e0 * REC_BYTES, (uint32_t)(cnt * REC_BYTES), &bar[b]);
        } else {
            for (int s = 0; s < cnt; ++s)
                tma_bulk_g2s(dst + (size_t)s * REC_D, a.rec + (size_t)a.elist[e0 + s] * REC_BYTES, REC_BYTES, &bar[b]);
        }
    };
    if (t == 0 && g < ngroups) issue(g, 0);
    for (; g < ngroups; g += gridDim.x, buf ^= 1) {
        const int64_t gn = g + gridDim.x;
        if (t == 0 && gn < ngroups) issue(gn, buf ^ 1);   // prefetch the next group into the other buffer
        mbar_wait(&bar[buf], phase[buf]);
        phase[buf] ^= 1;

        const int64_t pos = g * EPB + slot;
        const bool live = active && pos < a.nelem;
        const int64_t iel = live ? (a.elist ? (int64_t)a.elist[pos] : pos) : 0;
        const double *rec = sRec + ((size_t)buf * EPB + slot) * REC_D;
        const int32_t *conn = reinterpret_cast<const int32_t *>(rec + NMET * NP);
        double *F = sFl + (size_t)slot * NSD * NEQ * NP;   // [d][eq][NP]
        double S[NEQ];
        int64_t ip = 0;
        if (live) {
            ip = conn[l];
            double q[NEQ], qe[NEQ + 1], f[NEQ], gg[NEQ], h[NEQ];
#pragma unroll
            for (int e = 0; e < NEQ; ++e) q[e] = a.u[(size_t)e * a.npoin + ip];
#pragma unroll
            for (int e = 0; e <= NEQ; ++e) qe[e] = EQ::NEEDS_QE ? a.qe[(size_t)e * a.npoin + ip] : 0.0;
            EQ::flux(a.phys, q, qe, f, gg, h);
#pragma unroll
            for (int e = 0; e < NEQ; ++e) {
                F[(0 * NEQ + e) * NP + l] = f[e];
                F[(1 * NEQ + e) * NP + l] = gg[e];
                if constexpr (NSD == 3) F[(2 * NEQ + e) * NP + l] = h[e];
            }
            if (a.lsource) {
                double xyz[3] = {0.0, 0.0, 0.0};
                if constexpr (EQ::NEEDS_XYZ) {
#pragma unroll
                    for (int d = 0; d < NSD; ++d) xyz[d] = a.coords[(size_t)d * a.npoin + ip];
                }
                EQ::source(a.phys, q, qe, xyz, S);
            } else {
#pragma unroll
                for (int e = 0; e < NEQ; ++e) S[e] = 0.0;
            }
            if constexpr (VISC) {
                double up[NEQ];
                EQ::primitives(a.phys, q, qe, up);
#pragma unroll
                for (int e = 0; e < NEQ; ++e) sU[((size_t)slot * NEQ + e) * NP + l] = up[e];
            }
        }
        __syncthreads();

        double out[NEQ];
        double mt[NMET];
        if (live) {
#pragma unroll
            for (int m = 0; m < NMET; ++m) mt[m] = rec[m * NP + l];
            const double wJ = mt[NMET - 1];
#pragma unroll
            for (int e = 0; e < NEQ; ++e) {
                const double *Fe = F + (0 * NEQ + e) * NP, *Ge = F + (1 * NEQ + e) * NP;
                if constexpr (NSD == 3) {
                    const double *He = F + (2 * NEQ + e) * NP;
                    double dFdxi = 0, dFdeta = 0, dFdzeta = 0, dGdxi = 0, dGdeta = 0, dGdzeta = 0, dHdxi = 0, dHdeta = 0,
                           dHdzeta = 0;
#pragma unroll
                    for (int m = 0; m < NGL; ++m) {
                        const double di = sD[m + NGL * i], dj = sD[m + NGL * j], dk = sD[m + NGL * k];
                        const int lx = m + NGL * (j + NGL * k), ly = i + NGL * (m + NGL * k), lz = i + NGL * (j + NGL * m);
                        dFdxi = fma(di, Fe[lx], dFdxi);
                        dFdeta = fma(dj, Fe[ly], dFdeta);
                        dFdzeta = fma(dk, Fe[lz], dFdzeta);
                        dGdxi = fma(di, Ge[lx], dGdxi);
                        dGdeta = fma(dj, Ge[ly], dGdeta);
                        dGdzeta = fma(dk, Ge[lz], dGdzeta);
                        dHdxi = fma(di, He[lx], dHdxi);
                        dHdeta = fma(dj, He[ly], dHdeta);
                        dHdzeta = fma(dk, He[lz], dHdzeta);
                    }
                    const double dFdx = dFdxi * mt[0] + dFdeta * mt[3] + dFdzeta * mt[6];
                    const double dGdy = dGdxi * mt[1] + dGdeta * mt[4] + dGdzeta * mt[7];
                    const double dHdz = dHdxi * mt[2] + dHdeta * mt[5] + dHdzeta * mt[8];
                    out[e] = 0.0 - wJ * ((dFdx + dGdy + dHdz) - S[e]);
                } else {
                    double dFdxi = 0, dFdeta = 0, dGdxi = 0, dGdeta = 0;
#pragma unroll
                    for (int m = 0; m < NGL; ++m) {
                        const double di = sD[m + NGL * i], dj = sD[m + NGL * j];
                        const int lx = m + NGL * j, ly = i + NGL * m;
                        dFdxi = fma(di, Fe[lx], dFdxi);
                        dFdeta = fma(dj, Fe[ly], dFdeta);
                        dGdxi = fma(di, Ge[lx], dGdxi);
                        dGdeta = fma(dj, Ge[ly], dGdeta);
                    }
                    const double dFdx = dFdxi * mt[0] + dFdeta * mt[2];
                    const double dGdy = dGdxi * mt[1] + dGdeta * mt[3];
                    out[e] = 0.0 - wJ * ((dFdx + dGdy) - S[e]);
                }
            }
        }

        double outv[NEQ];
        if constexpr (VISC) {
            if constexpr (VISC == 2) {
                if (live) {
                    double D = a.sgs.delta;
                    if constexpr (NSD == 3) {
                        if (a.sgs.ad_lvl) D = ldexp(D, -a.sgs.ad_lvl[iel]);      // calculate_effective_delta, mesh.jl:5749-5751
                    }
                    sgs_node_fluxes<NSD, NGL, NEQ>(a.sgs, D * D, a.visc, sD, sU + (size_t)slot * NEQ * NP,
                                                   sGv + (size_t)slot * NEQ * NSD * NP, mt, mt[NMET - 1], i, j, k, l);
                }
            } else if (live) {
                const double wJ = mt[NMET - 1];
#pragma unroll
                for (int e = 0; e < NEQ; ++e) {
                    const double *Ue = sU + ((size_t)slot * NEQ + e) * NP;
                    double *Gv = sGv + ((size_t)slot * NEQ + e) * NSD * NP;
                    const double mu = a.visc[e];
                    if constexpr (NSD == 3) {
                        double dqdxi = 0, dqdeta = 0, dqdzeta = 0;
#pragma unroll
                        for (int m = 0; m < NGL; ++m) {
                            dqdxi = fma(sD[m + NGL * i], Ue[m + NGL * (j + NGL * k)], dqdxi);
                            dqdeta = fma(sD[m + NGL * j], Ue[i + NGL * (m + NGL * k)], dqdeta);
                            dqdzeta = fma(sD[m + NGL * k], Ue[i + NGL * (j + NGL * m)], dqdzeta);
                        }
                        double auxi = dqdxi * mt[0] + dqdeta * mt[3] + dqdzeta * mt[6];
                        const double dqdx = mu * auxi;
                        auxi = dqdxi * mt[1] + dqdeta * mt[4] + dqdzeta * mt[7];
                        const double dqdy = mu * auxi;
                        auxi = dqdxi * mt[2] + dqdeta * mt[5] + dqdzeta * mt[8];
                        const double dqdz = mu * auxi;
                        Gv[0 * NP + l] = (mt[0] * dqdx + mt[1] * dqdy + mt[2] * dqdz) * wJ;
                        Gv[1 * NP + l] = (mt[3] * dqdx + mt[4] * dqdy + mt[5] * dqdz) * wJ;
                        Gv[2 * NP + l] = (mt[6] * dqdx + mt[7] * dqdy + mt[8] * dqdz) * wJ;
                    } else {
                        double dqdxi = 0, dqdeta = 0;
#pragma unroll
                        for (int m = 0; m < NGL; ++m) {
                            dqdxi = fma(sD[m + NGL * i], Ue[m + NGL * j], dqdxi);
                            dqdeta = fma(sD[m + NGL * j], Ue[i + NGL * m], dqdeta);
                        }
                        double auxi = dqdxi * mt[0] + dqdeta * mt[2];
                        const double dqdx = mu * auxi;
                        auxi = dqdxi * mt[1] + dqdeta * mt[3];
                        const double dqdy = mu * auxi;
                        double flux_x = dqdx, flux_y = dqdy;
                        if constexpr (has_tau_u<EQ>::value) {
                            if (e == tau_u_eq<EQ>::value) {          // total-energy form: viscous work, rhs.jl:1988, 2018-2041
                                const double *Uu = sU + ((size_t)slot * NEQ + 1) * NP, *Uv = sU + ((size_t)slot * NEQ + 2) * NP;
                                double dudxi = 0, dudeta = 0, dvdxi = 0, dvdeta = 0;
#pragma unroll
                                for (int m = 0; m < NGL; ++m) {
                                    dudxi = fma(sD[m + NGL * i], Uu[m + NGL * j], dudxi);
                                    dudeta = fma(sD[m + NGL * j], Uu[i + NGL * m], dudeta);
                                    dvdxi = fma(sD[m + NGL * i], Uv[m + NGL * j], dvdxi);
                                    dvdeta = fma(sD[m + NGL * j], Uv[i + NGL * m], dvdeta);
                                }
                                const double dudx = dudxi * mt[0] + dudeta * mt[2], dudy = dudxi * mt[1] + dudeta * mt[3];
                                const double dvdx = dvdxi * mt[0] + dvdeta * mt[2], dvdy = dvdxi * mt[1] + dvdeta * mt[3];
                                const double div_u = dudx + dvdy;
                                const double mu2 = a.visc[1];
                                const double txx = 2.0 * mu2 * dudx - (2.0 / 3.0) * mu2 * div_u;
                                const double tyy = 2.0 * mu2 * dvdy - (2.0 / 3.0) * mu2 * div_u;
                                const double txy = mu2 * (dudy + dvdx);
                                const double ul = Uu[l], vl = Uv[l];
                                flux_x = flux_x + (txx * ul + txy * vl);
                                flux_y = flux_y + (txy * ul + tyy * vl);
                            }
                        }
                        Gv[0 * NP + l] = (mt[0] * flux_x + mt[1] * flux_y) * wJ;
                        Gv[1 * NP + l] = (mt[2] * flux_x + mt[3] * flux_y) * wJ;
                    }
                }
            }
            __syncthreads();
            if (live) {
#pragma unroll
                for (int e = 0; e < NEQ; ++e) {
                    const double *Gv = sGv + ((size_t)slot * NEQ + e) * NSD * NP;
                    // weak-form scatter of the reference written as a gather with the same
                    // accumulation order (ascending quadrature index), x -= a*b as fma(-a,b,x)
                    double axi = 0, aeta = 0, azeta = 0;
                    if constexpr (NSD == 3) {
#pragma unroll
                        for (int m = 0; m < NGL; ++m) {
                            axi = fma(-sD[i + NGL * m], Gv[0 * NP + m + NGL * (j + NGL * k)], axi);
                            aeta = fma(-sD[j + NGL * m], Gv[1 * NP + i + NGL * (m + NGL * k)], aeta);
                            azeta = fma(-sD[k + NGL * m], Gv[2 * NP + i + NGL * (j + NGL * m)], azeta);
                        }
                        outv[e] = axi + aeta + azeta;
                    } else {
#pragma unroll
                        for (int m = 0; m < NGL; ++m) {
                            axi = fma(-sD[i + NGL * m], Gv[0 * NP + m + NGL * j], axi);
                            aeta = fma(-sD[j + NGL * m], Gv[1 * NP + i + NGL * m], aeta);
                        }
                        outv[e] = axi + aeta;
                    }
                }
            }
        }

        if (live) {
            if (!a.atomics) {
#pragma unroll
                for (int e = 0; e < NEQ; ++e) {
                    a.rhs_el[((size_t)iel * NEQ + e) * NP + l] = out[e];
                    if constexpr (VISC) a.rhs_el_visc[((size_t)iel * NEQ + e) * NP + l] = outv[e];
                }
            } else {
                const double mi = a.Minv[ip];
#pragma unroll
                for (int e = 0; e < NEQ; ++e) {
                    double v = out[e];
                    if constexpr (VISC) v = v + outv[e];
                    atomicAdd(&a.du[(size_t)e * a.npoin + ip], v * mi);
                }
            }
        }
        __syncthreads();   // all reads of this buffer / flux tiles done before they are refilled
    }
}

// L2 prefetch of a contiguous range (cp.async.bulk.prefetch: a uniform-datapath instruction, issued lane by lane)
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------
// Per-unique-node pre-pass of the team kernels: the part of user_flux! that depends on the node
// only (the equation of state, one pow per node) is evaluated once per node instead of once per
// element-node (x1.95 at nop=4), and -- in atomics mode -- the scatter target is zeroed in the same
// sweep (replaces the memset).  Runs after the Dirichlet projection, like the flux evaluation it feeds.
// ------------------------------------------------------------------------------------------
struct AuxArgs {
    const double *u, *qe;
    double *aux;      // [NAUX][npoin]
    double *zero;     // du to clear (atomics mode) or nullptr
    int64_t npoin;
    Phys phys;
};

template <class EQ>
static __global__ void k_node_aux(const __grid_constant__ AuxArgs a) {
    constexpr int NEQ = EQ::NEQ;
    for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < a.npoin; ip += (int64_t)gridDim.x * blockDim.x) {
        double q[NEQ], qe[NEQ + 1];
#pragma unroll
        for (int e = 0; e < NEQ; ++e) q[e] = ((EQ::AUX_MASK >> e) & 1u) ? a.u[(size_t)e * a.npoin + ip] : 0.0;
#pragma unroll
        for (int e = 0; e <= NEQ; ++e) qe[e] = (EQ::NEEDS_QE && (e == NEQ || ((EQ::AUX_MASK >> e) & 1u))) ? a.qe[(size_t)e * a.npoin + ip] : 0.0;
        if constexpr (EQ::HAS_AUX) {
            double ax[EQ::NAUX > 0 ? EQ::NAUX : 1];
            EQ::aux(a.phys, q, qe, ax);
#pragma unroll
            for (int x = 0; x < EQ::NAUX; ++x) a.aux[(size_t)x * a.npoin + ip] = ax[x];
        }
        if (a.zero) {
#pragma unroll
            for (int e = 0; e < NEQ; ++e) a.zero[(size_t)e * a.npoin + ip] = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Atomics mode, 2N low-storage schemes (Williamson form S_i = A_i S_{i-1} + dt F(u), u += B_i S_i): the element kernels
// RED.ADD their mass-scaled contributions DIRECTLY into the low-storage register, kept as S' = S / dt and pre-scaled by
// the stage's A_i (S'_i = A_i S'_{i-1} + F).  What is left of the stage is ONE sweep:
//     u += (B_i dt) S'_i;   S' <- A_{i+1} S'_i (exact zeros when A_{i+1} = 0: first stage of the next step);   aux = EOS(u)
// 168 B per node instead of 288 (k_lsrk_update + k_node_aux with its zero-fill): no du array, no zero-fill, no separate
// equation-of-state pass.  The next rhs! starts at the Dirichlet projection, which re-evaluates aux at the nodes it changes.
// (Sums differ from tmp = A tmp + dt du by the association of the dt factor: inside the bars of the unordered mode.)
// ------------------------------------------------------------------------------------------
struct StageArgs {
    double *u, *acc;      // state and low-storage register S'
    const double *qe;
    double *aux;          // [NAUX][npoin]
    int64_t npoin;
    double Bdt, Anext;
    Phys phys;
};

template <class EQ>
static __global__ void __launch_bounds__(256) k_stage_direct(const __grid_constant__ StageArgs a) {
    constexpr int NEQ = EQ::NEQ;
    for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < a.npoin; ip += (int64_t)gridDim.x * blockDim.x) {
        double s[NEQ], q[NEQ], qe[NEQ + 1];
#pragma unroll
        for (int e = 0; e < NEQ; ++e) s[e] = a.acc[(size_t)e * a.npoin + ip];
#pragma unroll
        for (int e = 0; e < NEQ; ++e) q[e] = a.u[(size_t)e * a.npoin + ip];
#pragma unroll
        for (int e = 0; e <= NEQ; ++e) qe[e] = (EQ::NEEDS_QE && EQ::HAS_AUX && (e == NEQ || ((EQ::AUX_MASK >> e) & 1u))) ? a.qe[(size_t)e * a.npoin + ip] : 0.0;
#pragma unroll
        for (int e = 0; e < NEQ; ++e) {
            q[e] = q[e] + a.Bdt * s[e];
            a.u[(size_t)e * a.npoin + ip] = q[e];
            a.acc[(size_t)e * a.npoin + ip] = a.Anext == 0.0 ? 0.0 : a.Anext * s[e];
        }
        if constexpr (EQ::HAS_AUX) {
            double ax[EQ::NAUX > 0 ? EQ::NAUX : 1];
            EQ::aux(a.phys, q, qe, ax);
#pragma unroll
            for (int x = 0; x < EQ::NAUX; ++x) a.aux[(size_t)x * a.npoin + ip] = ax[x];
        }
    }
}

// interface / periodic-twin nodes under the direct accumulation: their share of S' is set aside and zeroed before the
// element kernels, so that the exchange sums pure contributions, and added back afterwards (thread = (node of the list, equation))
static __global__ void k_if_save_zero(double *a, int64_t npoin, int m, const int64_t *idx, int64_t len, double *base) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len * m) return;
    const size_t o = (size_t)(t % m) * npoin + idx[t / m];
    base[t] = a[o];
    a[o] = 0.0;
}
static __global__ void k_if_restore(double *a, int64_t npoin, int m, const int64_t *idx, int64_t len, const double *base) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len * m) return;
    const size_t o = (size_t)(t % m) * npoin + idx[t / m];
    a[o] = base[t] + a[o];
}

// cp.async (LDGSTS) helpers
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

struct GroupRetileArgs {
    const double *src;        // one metric array [E, n, n, n], element fastest (device copy); slot >= 0
    const double *omega;
    const double *Minv;       // slot -2
    const int64_t *connijk;   // slot -1
    const int32_t *epos;      // element -> position in record order (nullptr: identity); see order_elements()
    char *rec;
    int64_t nelem;
    int ngl, epb, group_bytes, zid_off, fid_off, z_off, w_off, wf_off;
    int slot;                 // 0..8 metric term, 9 = Je (stored as omega*J), -1 = node ids, -2 = -(omega*J*Minv)
};

// element-fastest Julia arrays -> element-group records of the team kernels (layout 5, ElemTeamCfg); thread = (element,
// local node), element fastest.  Element iel lives in group pos/epb, slot pos%epb, pos = epos[iel].
static __global__ void k_retile_group(GroupRetileArgs a) {
    const int n = a.ngl, nc = n * n, np = nc * n;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= a.nelem * np) return;
    const int64_t iel = tid % a.nelem;
    const int l = (int)(tid / a.nelem);
    const int i = l % n, j = (l / n) % n, k = l / nc;
    const int64_t pos = a.epos ? a.epos[iel] : iel;
    const int64_t g = pos / a.epb;
    const int s = (int)(pos % a.epb);
    char *rec = a.rec + (size_t)g * a.group_bytes;
    const int npl = 3 * n * a.epb, zrow = a.epb * nc, c = i + n * j;
    double *pl = reinterpret_cast<double *>(rec);
    double *zs = reinterpret_cast<double *>(rec + a.z_off);
    double *w = reinterpret_cast<double *>(rec + a.w_off), *wf = reinterpret_cast<double *>(rec + a.wf_off);
    int32_t *zid = reinterpret_cast<int32_t *>(rec + a.zid_off);
    int32_t *fid = reinterpret_cast<int32_t *>(rec + a.fid_off);
    const size_t src = (size_t)iel + (size_t)a.nelem * l;
    const int zpos = k * zrow + s * nc + c;
    if (a.slot == -1) {
        const int32_t ip = (int32_t)(a.connijk[src] - 1);
        zid[zpos] = ip;
        fid[s * np + l] = ip;
    } else if (a.slot == -2) {
        wf[zpos] = -(w[zpos] * a.Minv[zid[zpos]]);
    } else if (a.slot < 6) {
        const int X = a.slot % 3, lane = k + n * (X + 3 * s);
        pl[(size_t)((a.slot < 3 ? 0 : nc) + n * j + i) * npl + lane] = a.src[src];
    } else if (a.slot < 9) {
        zs[(size_t)(a.slot - 6) * n * zrow + zpos] = a.src[src];
    } else {
        const double wjk = a.omega[j] * a.omega[k];      // rhs.jl:1636-1643
        w[zpos] = a.omega[i] * wjk * a.src[src];
    }
}

// ------------------------------------------------------------------------------------------
// Fused per-element kernel, variant "team" (3D, inviscid, exact order): specialised warps of one CTA work on a
// group of EPB elements (2 at nop=4).  profiles/r01f: the round-1 pencil kernels (one thread per LGL line; removed in round 2) were bound by the LSU data pipe --
// 24 doubles cross shared memory per node and equation (9 line loads, 12 partial-product exchanges, 3 flux
// stores).  Here
//   * the PLANE warp: lane (slot, X in {F,G,H}, k) keeps the 25 values of plane k of field X_e in registers and
//     computes BOTH in-plane derivatives from them (25 loads feed 250 FMAs), combines them with its register-
//     resident metric terms xi_X, eta_X and writes the exact partial  B_X = dX/dxi*xi_X + dX/deta*eta_X
//     (rhs.jl:1679-1687, first two products of each sum, left to right);
//   * the ZETA warp: lane (i,j) of each element adds the zeta term from its line, (B_X + dX/dzeta*zeta_X),
//     sums F,G,H parts left to right, applies the source and the quadrature weight and scatters (RED.ADD or
//     rhs_el store) -- the reference's order, bit for bit;
//   * the two roles run skewed by one equation on double-buffered B tiles: one block barrier per equation,
//     15 doubles per node and equation through shared memory instead of 24;
//   * flux tiles are ordered (equation, X) and padded to GB = 3 (mod 16) doubles so the 30 plane lanes of a
//     warp-wide access fall on every bank pair at most twice (scripts/analysis/bank_search_team.py).
// ------------------------------------------------------------------------------------------
// ZW = number of zeta-role warps (1: one warp walks all element slots; EPB: one warp per slot).
// MODE = 0: rhs_el store (deterministic DSS); 1: RED.ADD of omega*J-weighted values; 2: RED.ADD with M^-1 pre-folded
#ifndef JX_TEAM_L2PF
#define JX_TEAM_L2PF 1          // pull the next group's record into L2 one group ahead
#endif
#ifndef JX_TEAM_CHUNK
#define JX_TEAM_CHUNK 5         // plane role: outputs in flight (2 FMA chains each)
#endif
template <int NGL, class EQ, int ZW = 1, int PW = 1>
struct ElemTeamCfg {
    static constexpr int N = NGL, NC = NGL * NGL, NP = NGL * NGL * NGL, NEQ = EQ::NEQ;
    static constexpr int EPB = 32 / (3 * NGL);                  // elements per group: 3*N*EPB plane lanes <= 32
    static_assert(EPB >= 1 && NC <= 32, "team kernel: nop <= 4");
    static constexpr int NPL = 3 * NGL * EPB;
    static_assert(ZW == 1 || ZW == EPB, "one zeta warp, or one per element slot");
    static constexpr int SPW = EPB / ZW;                        // element slots per zeta warp
    static_assert(PW == 1 || PW == 2, "one plane warp, or two sharing the nodes of every plane");
    static constexpr int NT = 32 * (PW + ZW);
    static constexpr int MAXREG = NT == 64 ? 200 : (NT == 96 ? 168 : 128);   // 4-5 CTAs per SM
    static constexpr int NSPLIT = PW == 1 ? NC : (NC + 1) / 2;              // plane warp 0 owns nodes [0,NSPLIT), warp 1 the rest
    static constexpr int NNODE = EPB * NP;
    static constexpr int R = (NNODE + NT - 1) / NT;
    static constexpr int GB = (EPB * NP + 12) / 16 * 16 + 3;    // >= EPB*NP, = 3 (mod 16)
    static constexpr int NFLD = 3 * NEQ;
    static constexpr int NTILE = NFLD + 6 + (EQ::SRC_EQ >= 0 ? 1 : 0);
    static constexpr size_t SMEM_BYTES = (size_t)NTILE * GB * 8;
    static constexpr int NQ = EQ::NEQ - (EQ::FLUX_QMASK == ((1u << (EQ::NEQ - 1)) - 1u) ? 1 : 0);
    static constexpr int NCOMP = NQ + EQ::NAUX;
    // pair records (layout 5), every row packed to the lanes that read it:
    //   [0, Z_OFF)        2*NC plane rows of NPL doubles: xi_X (rows n), eta_X (rows NC + n) at plane node n, lane k + N*(X + 3*slot)
    //   [Z_OFF, W_OFF)    3*N zeta rows of ZROW = EPB*NC doubles: zeta_q at node m of the zeta line, row q*N + m, position slot*NC + c
    //   [W_OFF, WF_OFF)   N rows of ZROW doubles: omega*J          (MODE 0 / 1)
    //   [WF_OFF, ZID_OFF) N rows of ZROW doubles: -(omega*J*Minv)  (MODE 2); a launch reads ONE of the two weight blocks
    //   [ZID_OFF, FID_OFF) int32 zeta-view node ids, N rows of ZROW; then int32 flux-view node ids [EPB*NP]
    static constexpr int ZROW = EPB * NC;
    static constexpr int Z_OFF = round_up(2 * NC * NPL * 8, 16);
    static constexpr int W_OFF = Z_OFF + 3 * NGL * ZROW * 8;
    static constexpr int WBYTES = round_up(NGL * ZROW * 8, 16);
    static constexpr int WF_OFF = W_OFF + WBYTES;
    static constexpr int ZID_OFF = WF_OFF + WBYTES;
    static constexpr int FID_OFF = ZID_OFF + round_up(NGL * ZROW * 4, 16);
    static constexpr int GROUP_BYTES = round_up(FID_OFF + NNODE * 4, 128);
};

// DYN = true: the launch walks a LIST of groups (a.glist) handed out through an atomic counter instead of the static
// blockIdx stride, and CTAs that land on the first a.reserve_sms SMs leave at once -- the interior launch of the
// interface-first split keeps those SMs free for the exchange kernels of the second stream.
template <int NGL, class EQ, int ZW, int MODE, int PW, bool DYN = false>
static __global__ void __maxnreg__((ElemTeamCfg<NGL, EQ, ZW, PW>::MAXREG))
k_elem_team(const __grid_constant__ ElemArgs a) {
    using C = ElemTeamCfg<NGL, EQ, ZW, PW>;
    constexpr int SPW = C::SPW;
    constexpr int N = NGL, NC = C::NC, NP = C::NP, NEQ = C::NEQ, NT = C::NT, R = C::R, GB = C::GB, EPB = C::EPB;
    constexpr int NQ = C::NQ, NCOMP = C::NCOMP, ZROW = C::ZROW, NPL = C::NPL;
    static_assert(EQ::SRC_EQ >= -1, "team kernels keep at most one source component");
    static_assert(EQ::HAS_AUX, "the team kernels use the two-stage flux functors");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *X = reinterpret_cast<double *>(smem_raw);     // [NEQ][3][GB]: F_e, G_e, H_e adjacent
    double *B = X + (size_t)C::NFLD * GB;                 // [2][3][GB]
    double *Sf = B + 6 * GB;                              // [GB]

    const int t = threadIdx.x, lane = t & 31;
    const bool plane_warp = t < 32 * PW;
    // plane role: lane = k + N*(X + 3*slot)
    const int pk = lane % N, pX = (lane / N) % 3, ps = lane / (3 * N);
    const bool pact = lane < C::NPL;
    const int plane_l = pact ? lane : 0;                  // inactive lanes re-read lane 0's column (rows are packed to NPL doubles)
    const int poff = ps * NP + NC * pk;                   // plane k of element slot ps inside a tile
    // zeta role: lane c = i + N*j
    const bool zact = lane < NC;
    const int c = zact ? lane : 0;
    const int zw = (t >> 5) - PW;                         // zeta warp index (role branch only)
    constexpr bool fold = MODE == 2;
#define JX_D(m, i) a.dpsi[(m) + NGL * (i)]
    const int64_t ngroups = (a.nelem + EPB - 1) / EPB;
    auto fid_of = [&](int64_t g) { return reinterpret_cast<const int32_t *>(a.rec + (size_t)g * C::GROUP_BYTES + C::FID_OFF); };
    // group sequence of this CTA: static stride, or (DYN) list positions taken from the work counter one group ahead;
    // "no more work" is the group id ngroups, which every consumer below already treats as out of range
    __shared__ int s_grp[4];
    int64_t gfirst = blockIdx.x;
    int gnext = 0;
    if constexpr (DYN) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if ((int)smid < a.reserve_sms) {
            if (t == 0) s_grp[3] = atomicAdd(a.exit_ctr, 1) < a.exit_budget ? 1 : 0;
            __syncthreads();
            if (s_grp[3]) return;
        }
        if (t == 0) {
            const int p0 = atomicAdd(a.gctr, 1), p1 = atomicAdd(a.gctr, 1);
            s_grp[0] = p0 < a.nlist ? a.glist[p0] : (int)ngroups;
            s_grp[1] = p1 < a.nlist ? a.glist[p1] : (int)ngroups;
        }
        __syncthreads();
        gfirst = s_grp[0];
        gnext = s_grp[1];
    }
    auto next_of = [&](int64_t g) -> int64_t {
        if constexpr (DYN) return gnext;
        else return g + gridDim.x;
    };
    auto fetch_ahead = [&]() {      // DYN: thread 0, once per group, between two block barriers of the step loop
        if constexpr (DYN) {
            const int p = atomicAdd(a.gctr, 1);
            s_grp[2] = p < a.nlist ? a.glist[p] : (int)ngroups;
        }
    };
    auto advance = [&](int64_t g) -> int64_t {   // after the last block barrier of a group
        if constexpr (DYN) {
            const int64_t r = gnext;
            gnext = s_grp[2];
            return r;
        } else return g + gridDim.x;
    };
    int fidn[R];
    if (gfirst < ngroups) {
        const int32_t *fi = fid_of(gfirst);
#pragma unroll
        for (int r = 0; r < R; ++r) fidn[r] = r * NT + t < C::NNODE ? __ldcs(fi + r * NT + t) : 0;
    }
    // (the first group's gathers are issued below, once issue_gathers is defined)
    // ---- pieces shared by the roles (inlined into each role's branch so that every role has its own register
    //      allocation: the plane metrics and the zeta metrics never coexist).  All warps take part in the flux
    //      phase (giving it to the zeta warps alone was measured 35 % slower, profiles/r01f).
    // q / aux gathers of the NEXT group, staged asynchronously (cp.async, no registers, no scoreboard) in the
    // flux tiles of the first NST equations, which are dead once their zeta step is done: issued by every warp
    // at the start of step ISSUE_STEP and consumed by the next group's flux phase, so the gather latency is
    // covered by the remaining equation steps.  Staged component x of group node n lives at X[x*GB + n]; the
    // flux phase reads its own node's values before it overwrites position n of every tile.
    constexpr int NST = (NCOMP + 2) / 3, ISSUE_STEP = NST + 1;
    static_assert(ISSUE_STEP <= NEQ, "not enough dead flux tiles to stage the gathers");
    auto issue_gathers = [&](int64_t g, const int(&fid)[R]) {
        if (g < ngroups) {
            const int cnt = (int)(a.nelem - g * EPB < EPB ? a.nelem - g * EPB : EPB);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = r * NT + t;
                if (n < cnt * NP) {
#pragma unroll
                    for (int e = 0; e < NQ; ++e) cp_async8(X + e * GB + n, a.u + (size_t)e * a.npoin + fid[r]);
#pragma unroll
                    for (int x = 0; x < EQ::NAUX; ++x) cp_async8(X + (NQ + x) * GB + n, a.aux + (size_t)x * a.npoin + fid[r]);
                }
            }
        }
        cp_async_commit();
    };
    auto prefetch_next = [&](int64_t gn, int(&fid)[R]) {   // next group: flux-view node ids -> registers, record -> L2
        if (gn < ngroups) {
            const int32_t *fi = fid_of(gn);
#pragma unroll
            for (int r = 0; r < R; ++r) fid[r] = r * NT + t < C::NNODE ? __ldcs(fi + r * NT + t) : 0;
#if JX_TEAM_L2PF
            // everything this launch reads of the record except the flux-view ids (loaded above): the plane and zeta rows,
            // ONE of the two weight blocks, the zeta-view ids -- two contiguous ranges
            // The instruction runs on the uniform datapath, lane by lane: 22 one-kilobyte chunks issued by the first lanes of
            // warp 0 made that warp the last to reach the group barrier, with the three others waiting for it (8 % of all warp
            // samples, profiles/r02h_team_source_stalls.md).  Now lane 0 of every warp pulls an equal slice of the first
            // range and lane 1 of the last warp the second: NW + 1 bulk prefetches per group, spread over the warps.
            constexpr int A_END = fold ? C::W_OFF : C::WF_OFF, B_BEG = fold ? C::WF_OFF : C::ZID_OFF;
            constexpr int NW = NT / 32;
            constexpr int SL = ((A_END + NW - 1) / NW + 15) / 16 * 16;
            const char *rn = a.rec + (size_t)gn * C::GROUP_BYTES;
            const int w = t >> 5;
            if (lane == 0) {
                const int off = w * SL;
                const int sz = ((A_END - off) < SL ? (A_END - off) : SL) & ~15;
                if (sz > 0) prefetch_l2_bulk(rn + off, (uint32_t)sz);
            } else if (lane == 1 && w == NW - 1) {
                prefetch_l2_bulk(rn + B_BEG, (uint32_t)((C::FID_OFF - B_BEG) & ~15));
            }
#endif
        }
    };
    // flux / source at every node of the group, node-parallel (group node n = slot*NP + l)
    // flux / source at every node of the group, node-parallel (group node n = slot*NP + l), from the staged gathers
    auto flux_phase = [&](int cnt) {
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ad = r * NT + t;
            if (ad < cnt * NP) {
                double q[NEQ], ax[EQ::NAUX], f[NEQ], gg[NEQ], h[NEQ];
#pragma unroll
                for (int e = 0; e < NEQ; ++e) q[e] = e < NQ ? X[(e < NQ ? e : 0) * GB + ad] : 1.0;
#pragma unroll
                for (int x = 0; x < EQ::NAUX; ++x) ax[x] = X[(NQ + x) * GB + ad];
                EQ::flux_aux(a.phys, q, ax, f, gg, h);
#pragma unroll
                for (int e = 0; e < NEQ; ++e) {
                    X[(e * 3 + 0) * GB + ad] = f[e];
                    X[(e * 3 + 1) * GB + ad] = gg[e];
                    X[(e * 3 + 2) * GB + ad] = h[e];
                }
                if constexpr (EQ::SRC_EQ >= 0) Sf[ad] = a.lsource ? EQ::source_aux(a.phys, q, ax) : 0.0;
            }
        }
    };
    // block barrier reached from both role branches (bar.sync counts warps, not program locations)
    auto block_sync = [&]() { asm volatile("bar.sync 0;" ::: "memory"); };

    // plane role over the node range [LO, HI) of every plane (PW = 2: the two plane warps split the 25 outputs;
    // both read the whole plane, each keeps only its own metric terms -> half the registers, half the step time)
    auto plane_role = [&](auto lo_c, auto hi_c) {
        constexpr int LO = decltype(lo_c)::value, HI = decltype(hi_c)::value, NN = HI - LO;
        constexpr int CHUNK = JX_TEAM_CHUNK;     // outputs in flight: 2*CHUNK independent FMA chains
        for (int64_t g = gfirst; g < ngroups; g = advance(g)) {
            const int cnt = (int)(a.nelem - g * EPB < EPB ? a.nelem - g * EPB : EPB);
            const double *pl = reinterpret_cast<const double *>(a.rec + (size_t)g * C::GROUP_BYTES);
            double mxi[NN], met[NN];             // xi_X, eta_X at this warp's nodes of plane k (lane-major streams)
#pragma unroll
            for (int n = 0; n < NN; ++n) { mxi[n] = __ldcs(pl + (LO + n) * NPL + plane_l); met[n] = __ldcs(pl + (NC + LO + n) * NPL + plane_l); }
            flux_phase(cnt);
            prefetch_next(next_of(g), fidn);
            block_sync();
            const bool live = pact && ps < cnt;
#pragma unroll 1
            for (int step = 0; step <= NEQ; ++step) {
                if (step == ISSUE_STEP) issue_gathers(next_of(g), fidn);
                if constexpr (DYN && LO == 0) {
                    if (step == 1 && t == 0) fetch_ahead();
                }
                if (step < NEQ && live) {
                    const double *T = X + (size_t)(step * 3 + pX) * GB + poff;
                    double *Bo = B + (size_t)((step & 1) * 3 + pX) * GB + poff;
                    double v[NC];
#pragma unroll
                    for (int n = 0; n < NC; ++n) v[n] = T[n];
#pragma unroll
                    for (int c0 = 0; c0 < NN; c0 += CHUNK) {
                        double dx[CHUNK], de[CHUNK];
#pragma unroll
                        for (int u = 0; u < CHUNK; ++u) { dx[u] = 0.0; de[u] = 0.0; }
#pragma unroll
                        for (int m = 0; m < N; ++m)
#pragma unroll
                            for (int u = 0; u < CHUNK; ++u) {
                                if (c0 + u < NN) {
                                    const int n = LO + c0 + u, i = n % N, j = n / N;
                                    dx[u] = fma(JX_D(m, i), v[N * j + m], dx[u]);
                                    de[u] = fma(JX_D(m, j), v[N * m + i], de[u]);
                                }
                            }
#pragma unroll
                        for (int u = 0; u < CHUNK; ++u)
                            if (c0 + u < NN) Bo[LO + c0 + u] = dx[u] * mxi[c0 + u] + de[u] * met[c0 + u];
                    }
                }
                block_sync();   // B[step&1] complete; the zeta warps are done with B[(step-1)&1]
            }
        }
    };
    issue_gathers(gfirst, fidn);             // first group: nothing to hide behind
    if (plane_warp) {
        // =============================== PLANE ROLE ===============================
        if constexpr (PW == 1) plane_role(std::integral_constant<int, 0>{}, std::integral_constant<int, NC>{});
        else {
            if (t < 32) plane_role(std::integral_constant<int, 0>{}, std::integral_constant<int, C::NSPLIT>{});
            else plane_role(std::integral_constant<int, C::NSPLIT>{}, std::integral_constant<int, NC>{});
        }
    } else {
        // =============================== ZETA ROLE ===============================
        for (int64_t g = gfirst; g < ngroups; g = advance(g)) {
            const int64_t e0 = g * EPB;
            const int cnt = (int)(a.nelem - e0 < EPB ? a.nelem - e0 : EPB);
            const char *rec = a.rec + (size_t)g * C::GROUP_BYTES;
            const double *zs = reinterpret_cast<const double *>(rec + C::Z_OFF);
            const double *ws = reinterpret_cast<const double *>(rec + (fold ? C::WF_OFF : C::W_OFF));
            const int32_t *zid = reinterpret_cast<const int32_t *>(rec + C::ZID_OFF);
            double mz[SPW][3][N], wj[SPW][N];
            int ip[SPW][N];
#pragma unroll
            for (int sl = 0; sl < SPW; ++sl) {
                const int zp = (zw * SPW + sl) * NC + c;
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int m = 0; m < N; ++m) mz[sl][q][m] = __ldcs(zs + (q * N + m) * ZROW + zp);
#pragma unroll
                for (int m = 0; m < N; ++m) {
                    wj[sl][m] = __ldcs(ws + m * ZROW + zp);
                    ip[sl][m] = __ldcs(zid + m * ZROW + zp);
                }
            }
            flux_phase(cnt);
            prefetch_next(next_of(g), fidn);
            block_sync();
#pragma unroll 1
            for (int step = 0; step <= NEQ; ++step) {
                if (step == ISSUE_STEP) issue_gathers(next_of(g), fidn);
                if (step >= 1) {
                    const int e = step - 1;
#pragma unroll
                    for (int sl = 0; sl < SPW; ++sl) {
                        const int s = zw * SPW + sl;
                        if (zact && s < cnt) {
                            const double *Fe = X + (size_t)(e * 3) * GB + s * NP + c, *Ge = Fe + GB, *He = Ge + GB;
                            const double *Bf = B + (size_t)((e & 1) * 3) * GB + s * NP + c;
                            double f[N], gg[N], h[N], b[3][N], Sv[N];
#pragma unroll
                            for (int m = 0; m < N; ++m) { f[m] = Fe[NC * m]; gg[m] = Ge[NC * m]; h[m] = He[NC * m]; }
#pragma unroll
                            for (int q = 0; q < 3; ++q)
#pragma unroll
                                for (int k = 0; k < N; ++k) b[q][k] = Bf[q * GB + NC * k];
#pragma unroll
                            for (int k = 0; k < N; ++k) Sv[k] = 0.0;
                            if constexpr (EQ::SRC_EQ >= 0) {
                                if (e == EQ::SRC_EQ) {
#pragma unroll
                                    for (int k = 0; k < N; ++k) Sv[k] = Sf[s * NP + c + NC * k];
                                }
                            }
                            double dF[N], dG[N], dH[N];
#pragma unroll
                            for (int o = 0; o < N; ++o) { dF[o] = 0.0; dG[o] = 0.0; dH[o] = 0.0; }
#pragma unroll
                            for (int m = 0; m < N; ++m)
#pragma unroll
                                for (int o = 0; o < N; ++o) {
                                    dF[o] = fma(JX_D(m, o), f[m], dF[o]);
                                    dG[o] = fma(JX_D(m, o), gg[m], dG[o]);
                                    dH[o] = fma(JX_D(m, o), h[m], dH[o]);
                                }
                            double *due = a.du + (size_t)e * a.npoin;
                            double *rhe = nullptr;
                            if constexpr (MODE == 0) {
                                const int64_t eo = a.eorig ? (int64_t)__ldg(a.eorig + e0 + s) : e0 + s;
                                rhe = a.rhs_el + ((size_t)eo * NEQ + e) * NP + c;
                            }
#pragma unroll
                            for (int k = 0; k < N; ++k) {
                                const double dFdx = b[0][k] + dF[k] * mz[sl][0][k];
                                const double dGdy = b[1][k] + dG[k] * mz[sl][1][k];
                                const double dHdz = b[2][k] + dH[k] * mz[sl][2][k];
                                const double r = (dFdx + dGdy) + dHdz;
                                if constexpr (MODE == 2) atomicAdd(due + ip[sl][k], wj[sl][k] * (r - Sv[k]));   // wj = -(omega*J*Minv)
                                else {
                                    const double out = 0.0 - wj[sl][k] * (r - Sv[k]);
                                    if constexpr (MODE == 0) rhe[NC * k] = out;
                                    else atomicAdd(due + ip[sl][k], out);
                                }
                            }
                        }
                    }
                }
                block_sync();
            }
        }
    }
#undef JX_D
}

// ------------------------------------------------------------------------------------------
// self test of Recip::div against the compiler's correctly rounded `/` (jx_selftest(ctx, 0, n, &bad)).
// Samples: b = density-like values in [1e-3, 1e3] and wide-range values, a = momentum-like values of
// both signs, exact zeros, powers of two and near-overflow / near-underflow magnitudes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t jx_mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
static __global__ void k_selftest_div(int64_t n, uint64_t seed, unsigned long long *bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h1 = jx_mix64(seed + 2 * (uint64_t)i), h2 = jx_mix64(seed + 2 * (uint64_t)i + 1);
        // mantissas from the hash, exponents from a small menu
        const int mode = (int)(h1 & 7);
        const int eb = mode < 5 ? (int)((h2 >> 52) % 21) - 10 : (int)((h2 >> 52) % 1200) - 600;
        const int ea = mode < 5 ? (int)((h1 >> 52) % 41) - 20 : (int)((h1 >> 52) % 1800) - 900;
        double b = __longlong_as_double((long long)((h2 & 0x000fffffffffffffull) | ((uint64_t)(1023 + eb) << 52)));
        double a = __longlong_as_double((long long)((h1 >> 3 & 0x000fffffffffffffull) | ((uint64_t)(1023 + ea) << 52)));
        if (h1 & 8) a = -a;
        if (mode == 6 && (h2 & 16)) b = -b;
        if ((h1 >> 4 & 63) == 0) a = 0.0;
        if ((h1 >> 10 & 63) == 0) a = ldexp(1.0, ea);
        const Recip rc(b);
        const double q1 = rc.div(a), q2 = a / b;
        if (__double_as_longlong(q1) != __double_as_longlong(q2) && !(q1 != q1 && q2 != q2)) atomicAdd(bad, 1ull);
    }
}

// ------------------------------------------------------------------------------------------
// DSS gather + M^-1 + stage update.
//   RHS[ip] = Σ_e rhs_el (ascending element)  [+ Σ_e rhs_el_visc]     DSS_rhs!, rhs.jl:624, 671-672
//   mode 0: out = RHS (un-scaled; the halo path continues)            DSS_global_RHS! follows
//   mode 1: du  = Minv*RHS                                            divide_by_mass_matrix!, RHStoDU!
//   mode 2: 2N low-storage stage: tmp = A*tmp + dt*(Minv*RHS); u += B*tmp   (ArrayFuse, rhs.jl:15-20)
// ------------------------------------------------------------------------------------------
struct GatherArgs {
    const double *rhs_el, *rhs_el_visc;
    const int64_t *ptr;
    const uint32_t *idx;
    const double *Minv;
    double *out;      // du or RHS
    double *u, *tmp;  // mode 2
    int64_t npoin;
    int np, neqs, mode, first_stage;
    double A, B, dt;
};

template <int NEQ>
static __global__ void k_gather(GatherArgs a) {
    const int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= a.npoin) return;
    const int64_t b = a.ptr[ip], e = a.ptr[ip + 1];
    double acc[NEQ], accv[NEQ];
#pragma unroll
    for (int q = 0; q < NEQ; ++q) { acc[q] = 0.0; accv[q] = 0.0; }
    for (int64_t p = b; p < e; ++p) {
        const uint32_t v = a.idx[p];
        const uint32_t el = v / (uint32_t)a.np, l = v % (uint32_t)a.np;
        const size_t base = (size_t)el * NEQ * a.np + l;
#pragma unroll
        for (int q = 0; q < NEQ; ++q) acc[q] = acc[q] + a.rhs_el[base + (size_t)q * a.np];
        if (a.rhs_el_visc) {
#pragma unroll
            for (int q = 0; q < NEQ; ++q) accv[q] = accv[q] + a.rhs_el_visc[base + (size_t)q * a.np];
        }
    }
    const double mi = a.Minv[ip];
#pragma unroll
    for (int q = 0; q < NEQ; ++q) {
        double r = acc[q];
        if (a.rhs_el_visc) r = r + accv[q];
        const size_t o = (size_t)q * a.npoin + ip;
        if (a.mode == 0) a.out[o] = r;
        else {
            const double k = mi * r;
            if (a.mode == 1) a.out[o] = k;
            else {
                const double tm = a.first_stage ? a.dt * k : a.A * a.tmp[o] + a.dt * k;
                a.tmp[o] = tm;
                a.u[o] = a.u[o] + a.B * tm;
            }
        }
    }
}

// node-wise passes used when a halo / periodic exchange sits between DSS and M^-1, by the
// atomics mode, and by the SSPRK stage forms
static __global__ void k_scale_minv(double *rhs, const double *Minv, int64_t npoin, int neqs) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npoin * neqs) return;
    rhs[t] = Minv[t % npoin] * rhs[t];
}

// 2N low-storage update with du already mass-scaled
static __global__ void k_lsrk_update(double *u, double *tmp, const double *du, int64_t n, double A, double B, double dt, int first) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double tm = first ? dt * du[t] : A * tmp[t] + dt * du[t];
    tmp[t] = tm;
    u[t] = u[t] + B * tm;
}

// generic linear combination  y = c0*x0 + c1*x1 + c2*x2 + c3*x3 + c4*x4 (left to right), used by SSPRK forms
struct LinArgs {
    double *y;
    const double *x[5];
    double c[5];
    int nterms;
    int64_t n;
    int form;   // 0: plain left-to-right sum of products; 1: (3*x0 + x1 + c2*x2)/4 ; 2: (x0 + 2*x1 + c2*x2)/3
};
static __global__ void k_lincomb(LinArgs a) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    if (a.form == 1) { a.y[t] = (3 * a.x[0][t] + a.x[1][t] + a.c[2] * a.x[2][t]) / 4; return; }
    if (a.form == 2) { a.y[t] = (a.x[0][t] + 2 * a.x[1][t] + a.c[2] * a.x[2][t]) / 3; return; }
    double s = a.c[0] * a.x[0][t];
    for (int i = 1; i < a.nterms; ++i) s = s + a.c[i] * a.x[i][t];
    a.y[t] = s;
}

// ------------------------------------------------------------------------------------------
// interface / periodic assembly (assemble_mpi!, mpi_communications.jl:260-338)
// buffers are node-major interleaved: buf[i*m + j] = a[idx[i], j]
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// build_metric_terms! on the device (metric_terms.jl:332-474 3D, :197-257 2D).  Thread = (element, local
// node), element fastest; every thread differentiates the element's coordinates along its three (two) LGL lines --
// psi is the identity at the LGL points, so the reference's triple sums collapse to single sums over the line --
// in ascending node order with separate multiply and add, forms the cofactors and the determinant with the
// reference's association, and writes ONE of the metric arrays (slot) in the reference's element-fastest layout;
// the record builders (k_retile / k_retile_group) then consume it exactly like a host-supplied array.
// ------------------------------------------------------------------------------------------
struct MetricBuildArgs {
    const int64_t *connijk;   // [E, n, n, n|1], 1-based
    const double *coords;     // [nsd][npoin]
    const double *dpsi;       // dpsi[m + n*i] = L'_m(xi_i)
    double *out;              // [E, n, n, n|1]
    int64_t nelem, npoin;
    int nsd, ngl, slot;       // 3D: 0..8 = dxi/dx dxi/dy dxi/dz deta/dx ... dzeta/dz, 9 = Je; 2D: 0..3, 4 = Je
};

static __global__ void k_build_metric(MetricBuildArgs a) {
    const int n = a.ngl, nsd = a.nsd;
    const int np = nsd == 3 ? n * n * n : n * n;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= a.nelem * np) return;
    const int64_t iel = tid % a.nelem;
    const int l = (int)(tid / a.nelem);
    const int i = l % n, j = (l / n) % n, k = l / (n * n);
    const int lidx[3] = {i, j, k};
    const int stride[3] = {1, n, n * n};
    double d[3][3];           // d[c][dir] = d(coordinate c) / d(reference direction dir)
    for (int dir = 0; dir < nsd; ++dir) {
        double acc[3] = {0.0, 0.0, 0.0};
        const int base = l - lidx[dir] * stride[dir];
        for (int m = 0; m < n; ++m) {
            const int64_t ip = a.connijk[iel + a.nelem * (int64_t)(base + m * stride[dir])] - 1;
            const double w = a.dpsi[m + n * lidx[dir]];
            for (int c = 0; c < nsd; ++c) {
                const double t = a.coords[(size_t)c * a.npoin + ip] * w;
                acc[c] = acc[c] + t;
            }
        }
        for (int c = 0; c < nsd; ++c) d[c][dir] = acc[c];
    }
    double v;
    if (nsd == 3) {
        const double dxdxi = d[0][0], dxdeta = d[0][1], dxdzeta = d[0][2];
        const double dydxi = d[1][0], dydeta = d[1][1], dydzeta = d[1][2];
        const double dzdxi = d[2][0], dzdeta = d[2][1], dzdzeta = d[2][2];
        const double c1 = dydeta * dzdzeta - dydzeta * dzdeta;
        const double c2 = dxdzeta * dzdeta - dxdeta * dzdzeta;
        const double c3 = dxdeta * dydzeta - dxdzeta * dydeta;
        const double Je = (dxdxi * c1 + dydxi * c2) + dzdxi * c3;
        if (a.slot == 9) v = Je;
        else {
            const double Jinv = 1.0 / Je;
            double cf;
            switch (a.slot) {
                case 0: cf = c1; break;
                case 1: cf = c2; break;
                case 2: cf = c3; break;
                case 3: cf = dydzeta * dzdxi - dydxi * dzdzeta; break;
                case 4: cf = dxdxi * dzdzeta - dxdzeta * dzdxi; break;
                case 5: cf = dxdzeta * dydxi - dxdxi * dydzeta; break;
                case 6: cf = dydxi * dzdeta - dydeta * dzdxi; break;
                case 7: cf = dxdeta * dzdxi - dxdxi * dzdeta; break;
                default: cf = dxdxi * dydeta - dxdeta * dydxi; break;
            }
            v = cf * Jinv;
        }
    } else {
        const double dxdxi = d[0][0], dxdeta = d[0][1], dydxi = d[1][0], dydeta = d[1][1];
        const double Je = dxdxi * dydeta - dydxi * dxdeta;
        if (a.slot == 4) v = Je;
        else {
            const double Jinv = 1.0 / Je;
            v = a.slot == 0 ? dydeta * Jinv : a.slot == 1 ? (-dxdeta) * Jinv : a.slot == 2 ? (-dydxi) * Jinv : dxdxi * Jinv;
        }
    }
    a.out[tid] = v;
}

// interface-first split: nodes named by the assembler lists, then the element groups whose records name one of them
// ------------------------------------------------------------------------------------------
// Setup on the device: diagonal mass matrix and IC conditioning.
//   DSS_mass! (element_matrices.jl:593-617) of build_mass_matrix! (:173-214, psi = identity at the LGL points):
//       M[ip] = sum over (element ascending, local node ascending) of (w_m*w_n)*w_o * Je          [2D: w_m*w_n * Je]
//   conformity4ncf_q! (Adaptivity/Projection.jl:2919-2970):  q[ip] <- Minv[ip] * DSS(wJac * q[ip]),  wJac = (w_i*w_j)*w_k*Je
// Both walk the node -> (element, local) CSR, which is sorted in the reference's visiting order.
// ------------------------------------------------------------------------------------------
static __global__ void k_mass_weight(const double *Je, const double *omega, int64_t nelem, int ngl, int nsd, double *w) {
    const int np = nsd == 3 ? ngl * ngl * ngl : ngl * ngl;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nelem * np) return;
    const int l = (int)(tid / nelem);
    const int i = l % ngl, j = (l / ngl) % ngl, k = l / (ngl * ngl);
    const double wij = omega[i] * omega[j];
    const double w3 = nsd == 3 ? wij * omega[k] : wij;
    w[tid] = w3 * Je[tid];
}
// neq == 0: M[ip] = sum of weights;  neq > 0: out[e][ip] = sum of (weight * q[e][ip]) for the first neq columns of q
static __global__ void k_mass_gather(const int64_t *ptr, const uint32_t *idx, const double *w, int64_t nelem, int np, int64_t npoin,
                                     int neq, const double *q, double *out) {
    const int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npoin) return;
    const int64_t p0 = ptr[ip], p1 = ptr[ip + 1];
    if (neq == 0) {
        double s = 0.0;
        for (int64_t p = p0; p < p1; ++p) { const uint32_t x = idx[p]; s += w[(size_t)(x / np) + (size_t)nelem * (x % np)]; }
        out[ip] = s;
    } else {
        for (int e = 0; e < neq; ++e) {
            const double qv = q[(size_t)e * npoin + ip];
            double s = 0.0;
            for (int64_t p = p0; p < p1; ++p) { const uint32_t x = idx[p]; s += w[(size_t)(x / np) + (size_t)nelem * (x % np)] * qv; }
            out[(size_t)e * npoin + ip] = s;
        }
    }
}
static __global__ void k_invert(double *a, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] = 1.0 / a[t];
}

static __global__ void k_mark_nodes(uint8_t *mask, const int64_t *idx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[idx[i]] = 1;
}
static __global__ void k_flag_groups(const char *rec, int group_bytes, int fid_off, int nnode, int64_t ngroups, const uint8_t *mask,
                                     uint8_t *flag) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const int32_t *fid = reinterpret_cast<const int32_t *>(rec + (size_t)g * group_bytes + fid_off);
    uint8_t f = 0;
    for (int n = 0; n < nnode; ++n) f |= mask[fid[n]];
    flag[g] = f;
}

static __global__ void k_pack(const double *a, int64_t npoin, int m, const int64_t *idx, int64_t len, double *buf) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len * m) return;
    const int64_t i = t / m;
    const int j = (int)(t % m);
    buf[t] = a[(size_t)j * npoin + idx[i]];
}
static __global__ void k_unpack(double *a, int64_t npoin, int m, const int64_t *idx, int64_t len, const double *buf) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len * m) return;
    const int64_t i = t / m;
    const int j = (int)(t % m);
    a[(size_t)j * npoin + idx[i]] = buf[t];
}
// owner-side add (mpi_communications.jl:292-300).  A node may occur several times in one peer's
// list (periodic corner twins); jx_upload_halo splits every list into rounds (round k = k-th
// occurrence of a node), so within one launch all targets are distinct and successive launches
// reproduce the reference's per-node summation order.  sel holds positions into idx / buf.
static __global__ void k_add_sel(double *a, int64_t npoin, int m, const int64_t *idx, const int64_t *sel, int64_t len,
                          const double *buf) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len * m) return;
    const int64_t i = sel[t / m];
    const int j = (int)(t % m);
    const size_t o = (size_t)j * npoin + idx[i];
    a[o] = a[o] + buf[i * m + j];
}

}  // namespace jx

#include "jx_tri.cuh"
#include "jx_visc.cuh"
