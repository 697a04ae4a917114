// jx_visc.cuh -- k_visc_quad: the AV viscous term (rhs.jl:1374-1461, 2794-2867) of 3D nop = 4 elements as a warp-team pass
// of its own, launched after the inviscid element kernel; bit-identical to the viscous pass of k_elem_node.
//
//     q_e = user_primitives!(u)                                      per node, equation e with mu_e != 0
//     (dq/dxi, dq/deta, dq/dzeta) = D-contractions of q_e            sequential FMA chains, ascending m
//     dqdx = mu*(dq/dxi*xi_x + dq/deta*eta_x + dq/dzeta*zeta_x), dqdy, dqdz
//     G_xi = (xi_x*dqdx + xi_y*dqdy + xi_z*dqdz)*wJ, G_eta, G_zeta    all nine metric terms of the node meet here
//     rhs_visc(i,j,k) = a_xi + a_eta + a_zeta,  a_xi = sum_m fma(-D[i,m], G_xi(m,j,k), .)  etc.
//
// (Its predecessor k_visc_team -- line-owner warps holding all nine metric terms of a zeta line in registers, 168 registers,
// three warps per CTA -- took 3.78 ms at 73^3 elements against 2.49 ms for k_visc_quad and was removed; profiles/r02e.)
#pragma once

namespace jx {

struct ViscRetileArgs {
    const double *src;        // one metric array [E, n, n, n], element fastest (device copy); slot >= 0
    const double *omega;
    const double *Minv;       // slot -2
    const int64_t *connijk;   // slot -1
    const int32_t *epos;      // element -> position in record order (nullptr: identity)
    char *rec;
    int64_t nelem;
    int ngl, epb, group_bytes, zid_off, fid_off;
    int slot;                 // 0..8 metric term, 9 = Je (stored as omega*J), -1 = node ids, -2 = Minv
};

struct ViscArgs {
    const char *rec;          // viscous pair records
    double *out_el;           // MODE 0: rhs_el_visc [E][NEQ][NP]
    int nv;                   // number of equations with mu != 0
    int ve[8];                // their indices, ascending
};

// ==========================================================================================================
// k_visc_quad (variant 13): four warps per element pair, every phase on all of them.
//
// The node-local step -- the only place where all nine metric terms of a node meet -- is
// NODE PARALLEL (thread = node, like the flux phase of the inviscid kernel): a thread loads the ten metric values of its two
// nodes when it needs them (coalesced rows of the node-ordered record, pulled into L2 one pair ahead) and drops them again.
// The contractions stay with the lanes that feed 250 (plane) or 25 (line) FMAs per load:
//     P0  all threads   primitives of the pair's nodes (state staged one pair ahead with cp.async into the U tiles)
//     P1  warps 0,1     plane lanes (slot = warp; lane = k + 5 v): dq/dxi, dq/deta of all viscous equations v   -> A_v, B_v
//         warps 2,3     line lanes  (slot = warp - 2; lane = i + 5 j): dq/dzeta                                  -> C_v
//     P2  all threads   node-local: (A,B,C)_v[n] <- (G_xi, G_eta, G_zeta)            rhs.jl:2826-2845, same expressions
//     P3  warps 0,1     A_v <- a_xi + a_eta  (backward in-plane contractions);   warps 2,3   C_v <- a_zeta
//     P4  all threads   out = A_v[n] + C_v[n] = (a_xi + a_eta) + a_zeta  -> RED.ADD (x Minv) or rhs_el_visc store
// Four block barriers per pair, 16 doubles per node and equation through shared memory, 128 registers, 4 CTAs = 16 warps
// per SM.  Every sum keeps the reference's order: bit-identical to k_elem_node<VISC> and the oracle.
//
// Records (layout of its own, one per pair): double [11][EPB*NP] node ordered (metric m of node l of slot s at
// m*EPB*NP + s*NP + l; m = 9: omega*J, 10: Minv), then int32 node ids [EPB*NP].
// ==========================================================================================================
template <int NGL, class EQ>
struct ViscQuadCfg {
    static constexpr int N = NGL, NC = NGL * NGL, NP = NGL * NGL * NGL, NEQ = EQ::NEQ;
    static constexpr int EPB = 2;
    static_assert(NGL == 5, "k_visc_quad: nop = 4");
    static constexpr int NT = 128, NNODE = EPB * NP, R = (NNODE + NT - 1) / NT;
    // tile stride = 13 (mod 16) doubles: the plane lanes (k + 5 v) of a warp then fall on distinct bank pairs in both halves of a
    // warp-wide access (25 k + GB v mod 16; with 3 (mod 16), the inviscid kernel's stride, v = 3 collides with v = 0)
    static constexpr int GB = (EPB * NP + 2) / 16 * 16 + 13;
    static_assert(GB >= EPB * NP && GB % 16 == 13, "tile stride");
    static constexpr int NMET = 11;
    static constexpr int ID_OFF = round_up(NMET * NNODE * 8, 16);
    static constexpr int GROUP_BYTES = round_up(ID_OFF + NNODE * 4, 128);
    static constexpr int NTILE = 4 * NEQ;
    static constexpr size_t SMEM_BYTES = (size_t)NTILE * GB * 8;
    static constexpr int MAXREG = 128;
};

static __global__ void k_retile_visc_quad(ViscRetileArgs a) {
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
    double *ms = reinterpret_cast<double *>(rec);
    int32_t *id = reinterpret_cast<int32_t *>(rec + a.zid_off);
    const size_t src = (size_t)iel + (size_t)a.nelem * l;
    const int nn = a.epb * np, p = s * np + l;
    if (a.slot == -1) id[p] = (int32_t)(a.connijk[src] - 1);
    else if (a.slot == -2) ms[(size_t)10 * nn + p] = a.Minv[id[p]];
    else if (a.slot < 9) ms[(size_t)a.slot * nn + p] = a.src[src];
    else {
        const double wjk = a.omega[j] * a.omega[k];
        ms[(size_t)9 * nn + p] = a.omega[i] * wjk * a.src[src];
    }
}

template <int NGL, class EQ, int MODE>
static __global__ void __maxnreg__((ViscQuadCfg<NGL, EQ>::MAXREG))
k_visc_quad(const __grid_constant__ ElemArgs a, const __grid_constant__ ViscArgs va) {
    using C = ViscQuadCfg<NGL, EQ>;
    constexpr int N = NGL, NC = C::NC, NP = C::NP, NEQ = C::NEQ, NT = C::NT, R = C::R, GB = C::GB, EPB = C::EPB, NN = C::NNODE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *S0 = reinterpret_cast<double *>(smem_raw);
    auto Ut = [&](int v) { return S0 + (size_t)v * GB; };
    auto At = [&](int v) { return S0 + (size_t)(NEQ + v) * GB; };
    auto Bt = [&](int v) { return S0 + (size_t)(2 * NEQ + v) * GB; };
    auto Ct = [&](int v) { return S0 + (size_t)(3 * NEQ + v) * GB; };
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nv = va.nv;
    const bool plane = warp < 2;
    const int slot = warp & 1;
    // plane lanes: lane = k + N*v
    const int pk = lane % N, pv = lane / N;
    // line lanes: lane c = i + N*j
    const bool lact = lane < NC;
    const int c = lact ? lane : 0;
#define JX_D(m, i) a.dpsi[(m) + NGL * (i)]
    const int64_t ngroups = (a.nelem + EPB - 1) / EPB;
    auto ids_of = [&](int64_t gg) { return reinterpret_cast<const int32_t *>(va.rec + (size_t)gg * C::GROUP_BYTES + C::ID_OFF); };
    int idc[R], idn[R];         // node ids of this thread's nodes: current pair, next pair
    auto load_ids = [&](int64_t gg, int(&id)[R]) {
        if (gg < ngroups) {
            const int32_t *p = ids_of(gg);
#pragma unroll
            for (int r = 0; r < R; ++r) id[r] = r * NT + t < NN ? __ldcs(p + r * NT + t) : 0;
        }
    };
    auto issue_gathers = [&](int64_t gg, const int(&id)[R]) {
        if (gg < ngroups) {
            const int nn = (int)(a.nelem - gg * EPB < EPB ? a.nelem - gg * EPB : EPB) * NP;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = r * NT + t;
                if (n < nn) {
#pragma unroll
                    for (int e = 0; e < NEQ; ++e) cp_async8(Ut(e) + n, a.u + (size_t)e * a.npoin + id[r]);
                }
            }
        }
        cp_async_commit();
    };
    // pair sequence of this CTA: static stride over all pairs, or over the list a.glist (interface-first split, DESIGN.md
    // section 5: interface pairs first, interior pairs beside the exchange).  Static stride: every CTA of the launch must run
    // (no early exit on reserved SMs here -- a CTA that left would take its pairs with it).
    const int64_t nl = a.glist ? (int64_t)a.nlist : ngroups;
    auto pair_at = [&](int64_t li) -> int64_t { return li < nl ? (a.glist ? (int64_t)a.glist[li] : li) : ngroups; };
    load_ids(pair_at(blockIdx.x), idn);
    if constexpr (!EQ::NEEDS_QE) issue_gathers(pair_at(blockIdx.x), idn);
    for (int64_t li = blockIdx.x; li < nl; li += gridDim.x) {
        const int64_t g = pair_at(li), gnx = pair_at(li + gridDim.x);
        const int cnt = (int)(a.nelem - g * EPB < EPB ? a.nelem - g * EPB : EPB);
        const int nn = cnt * NP;
        const char *rec = va.rec + (size_t)g * C::GROUP_BYTES;
        const double *ms = reinterpret_cast<const double *>(rec);
#pragma unroll
        for (int r = 0; r < R; ++r) idc[r] = idn[r];
        load_ids(gnx, idn);
        if (t < 2 && gnx < ngroups)      // next pair's record towards L2
            prefetch_l2_bulk(va.rec + (size_t)gnx * C::GROUP_BYTES + t * (C::GROUP_BYTES / 2 / 16 * 16), C::GROUP_BYTES / 2 / 16 * 16);
        // ---------------- P0: primitives, node parallel ----------------
        if constexpr (!EQ::NEEDS_QE) cp_async_wait<0>();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = r * NT + t;
            if (n < nn) {
                double q[NEQ], qe[NEQ + 1], up[NEQ];
                if constexpr (!EQ::NEEDS_QE) {
#pragma unroll
                    for (int e = 0; e < NEQ; ++e) q[e] = Ut(e)[n];
#pragma unroll
                    for (int e = 0; e <= NEQ; ++e) qe[e] = 0.0;
                } else {
#pragma unroll
                    for (int e = 0; e < NEQ; ++e) q[e] = __ldg(a.u + (size_t)e * a.npoin + idc[r]);
#pragma unroll
                    for (int e = 0; e <= NEQ; ++e) qe[e] = __ldg(a.qe + (size_t)e * a.npoin + idc[r]);
                }
                EQ::primitives(a.phys, q, qe, up);
                // U tiles are indexed by EQUATION (static stores; the consumers pick tile ve[v]): selecting up[ve[v]] here
                // turns into a local-memory array (profiles/r02e)
#pragma unroll
                for (int e = 0; e < NEQ; ++e) Ut(e)[n] = up[e];
            }
        }
        __syncthreads();
        // ---------------- P1: forward contractions ----------------
        if (plane) {
            if (slot < cnt && pv < nv) {
                const int poff = slot * NP + NC * pk;
                const double *T = Ut(va.ve[pv]) + poff;
                double *Ox = At(pv) + poff, *Oe = Bt(pv) + poff;
                double w[NC];
#pragma unroll
                for (int n = 0; n < NC; ++n) w[n] = T[n];
#pragma unroll
                for (int c0 = 0; c0 < NC; c0 += N) {
                    double dx[N], de[N];
#pragma unroll
                    for (int u = 0; u < N; ++u) { dx[u] = 0.0; de[u] = 0.0; }
#pragma unroll
                    for (int m = 0; m < N; ++m)
#pragma unroll
                        for (int u = 0; u < N; ++u) {
                            const int n = c0 + u, i = n % N, j = n / N;
                            dx[u] = fma(JX_D(m, i), w[N * j + m], dx[u]);
                            de[u] = fma(JX_D(m, j), w[N * m + i], de[u]);
                        }
#pragma unroll
                    for (int u = 0; u < N; ++u) { Ox[c0 + u] = dx[u]; Oe[c0 + u] = de[u]; }
                }
            }
        } else if (lact && slot < cnt) {
            const int lo = slot * NP + c;
#pragma unroll 1
            for (int v = 0; v < nv; ++v) {
                const double *Uq = Ut(va.ve[v]) + lo;
                double uq[N], dz[N];
#pragma unroll
                for (int m = 0; m < N; ++m) uq[m] = Uq[NC * m];
#pragma unroll
                for (int k = 0; k < N; ++k) dz[k] = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m)
#pragma unroll
                    for (int k = 0; k < N; ++k) dz[k] = fma(JX_D(m, k), uq[m], dz[k]);
                double *Cz = Ct(v) + lo;
#pragma unroll
                for (int k = 0; k < N; ++k) Cz[NC * k] = dz[k];
            }
        }
        // metric terms of this thread's nodes for P2: issued ahead of the barrier, their latency overlaps the wait
        double M[R][10];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = r * NT + t < NN ? r * NT + t : 0;
#pragma unroll
            for (int m = 0; m < 10; ++m) M[r][m] = __ldcs(ms + m * NN + n);
        }
        __syncthreads();
        // the U tiles are dead: stage the next pair's state there
        if constexpr (!EQ::NEEDS_QE) issue_gathers(gnx, idn);
        // ---------------- P2: node-local step, node parallel ----------------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = r * NT + t;
            if (n < nn) {
#pragma unroll 1
                for (int v = 0; v < nv; ++v) {
                    const double mu = a.visc[va.ve[v]];
                    double *pa = At(v) + n, *pb = Bt(v) + n, *pc = Ct(v) + n;
                    const double dqdxi = *pa, dqdeta = *pb, dqdzeta = *pc;
                    double auxi = dqdxi * M[r][0] + dqdeta * M[r][3] + dqdzeta * M[r][6];
                    const double dqdx = mu * auxi;
                    auxi = dqdxi * M[r][1] + dqdeta * M[r][4] + dqdzeta * M[r][7];
                    const double dqdy = mu * auxi;
                    auxi = dqdxi * M[r][2] + dqdeta * M[r][5] + dqdzeta * M[r][8];
                    const double dqdz = mu * auxi;
                    const double wJ = M[r][9];
                    *pa = (M[r][0] * dqdx + M[r][1] * dqdy + M[r][2] * dqdz) * wJ;
                    *pb = (M[r][3] * dqdx + M[r][4] * dqdy + M[r][5] * dqdz) * wJ;
                    *pc = (M[r][6] * dqdx + M[r][7] * dqdy + M[r][8] * dqdz) * wJ;
                }
            }
        }
        __syncthreads();
        // ---------------- P3: backward contractions ----------------
        if (plane) {
            if (slot < cnt && pv < nv) {
                const int poff = slot * NP + NC * pk;
                double *Gx = At(pv) + poff;
                const double *Ge = Bt(pv) + poff;
                double w[NC];
#pragma unroll
                for (int n = 0; n < NC; ++n) w[n] = Ge[n];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double gx[N], ax[N], ae[N];
#pragma unroll
                    for (int m = 0; m < N; ++m) gx[m] = Gx[N * j + m];
#pragma unroll
                    for (int i = 0; i < N; ++i) { ax[i] = 0.0; ae[i] = 0.0; }
#pragma unroll
                    for (int m = 0; m < N; ++m)
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            ax[i] = fma(-JX_D(i, m), gx[m], ax[i]);
                            ae[i] = fma(-JX_D(j, m), w[N * m + i], ae[i]);
                        }
#pragma unroll
                    for (int i = 0; i < N; ++i) Gx[N * j + i] = ax[i] + ae[i];
                }
            }
        } else if (lact && slot < cnt) {
            const int lo = slot * NP + c;
#pragma unroll 1
            for (int v = 0; v < nv; ++v) {
                double *Cz = Ct(v) + lo;
                double gz[N], az[N];
#pragma unroll
                for (int m = 0; m < N; ++m) gz[m] = Cz[NC * m];
#pragma unroll
                for (int k = 0; k < N; ++k) az[k] = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m)
#pragma unroll
                    for (int k = 0; k < N; ++k) az[k] = fma(-JX_D(k, m), gz[m], az[k]);
#pragma unroll
                for (int k = 0; k < N; ++k) Cz[NC * k] = az[k];
            }
        }
        double minv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) minv[r] = MODE == 2 ? __ldcs(ms + 10 * NN + (r * NT + t < NN ? r * NT + t : 0)) : 1.0;
        __syncthreads();
        // ---------------- P4: (a_xi + a_eta) + a_zeta, scatter, node parallel ----------------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = r * NT + t;
            if (n < nn) {
                int64_t eo = 0;
                if constexpr (MODE == 0) {
                    const int64_t pos = g * EPB + n / NP;
                    eo = a.eorig ? (int64_t)__ldg(a.eorig + pos) : pos;
                }
#pragma unroll 1
                for (int v = 0; v < nv; ++v) {
                    const int e = va.ve[v];
                    const double outv = At(v)[n] + Ct(v)[n];
                    if constexpr (MODE == 0) va.out_el[((size_t)eo * NEQ + e) * NP + n % NP] = outv;
                    else atomicAdd(a.du + (size_t)e * a.npoin + idc[r], outv * minv[r]);
                }
            }
        }
        // (the next P0 writes only U tiles; the barrier behind it orders P4's reads of A, C against the next P1)
    }
#undef JX_D
}

}  // namespace jx
