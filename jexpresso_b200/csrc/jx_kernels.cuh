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
}

// ------------------------------------------------------------------------------------------
// Fused per-element kernel: flux/source/primitives at the LGL nodes, sum-factorised inviscid
// divergence (rhs.jl:1615-1698 3D, :1501-1542 2D) and AV viscous term (rhs.jl:2794-2867 3D,
// :1973-2056 2D).  Variant "node": one thread per element node, EPB elements per CTA iteration,
// element records double-buffered through shared memory by TMA bulk copies.
// ------------------------------------------------------------------------------------------
struct ElemArgs {
    const double *u;
    const double *qe;
    const char *rec;
    double *rhs_el;        // deterministic mode: [E][NEQ][NP]
    double *rhs_el_visc;   // deterministic mode, viscous part kept separate (DSS'ed separately, rhs.jl:671-672)
    double *du;            // atomics mode target
    const double *Minv;
    const double *coords;  // [nsd][npoin] (only read by functors with NEEDS_XYZ)
    const int32_t *elist;  // optional element subset (interface / interior split); nullptr = all
    int64_t nelem, npoin;  // nelem = number of elements this launch processes
    int atomics;
    int lsource;
    Phys phys;
    double visc[8];
    double dpsi[64];       // Julia dψ[m,i] column-major: dpsi[m + NGL*i]
};

template <int NSD, int NGL, class EQ, bool VISC, int EPB>
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

template <int NSD, int NGL, class EQ, bool VISC, int EPB>
static __global__ void __launch_bounds__(ElemNodeCfg<NSD, NGL, EQ, VISC, EPB>::NT)
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
            if (live) {
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
                        Gv[0 * NP + l] = (mt[0] * dqdx + mt[1] * dqdy) * wJ;
                        Gv[1 * NP + l] = (mt[2] * dqdx + mt[3] * dqdy) * wJ;
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
