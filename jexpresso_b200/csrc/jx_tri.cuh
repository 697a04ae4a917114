// jx_tri.cuh -- k_elem_tri: three-role pencil kernel for nop = 7 (ngl = 8), 3D, inviscid, exact order.
//
// At ngl = 8 a plane has 64 nodes: the plane role of the team kernels (25 values of a plane in registers feed both
// in-plane derivatives) would need 128 registers for the plane alone.  Lines do fit: a thread that owns one LGL line
// keeps its 8 values in registers and produces the 8 derivatives along the line with 64 FMAs (8 per shared-memory load).
// One CTA works on one element (512 nodes) with three specialised roles of 64 lanes each:
//     xi role   lane (j,k):  A_X  = dX/dxi  * xi_X                         -> partial tile A
//     eta role  lane (i,k):  A_X += dX/deta * eta_X                        (in place)
//     zeta role lane (i,j):  dXdx = A_X + dX/dzeta * zeta_X;  r = (dFdx + dGdy) + dHdz;  rhs = 0 - wJ*(r - S)
// i.e. (dF/dxi*xi_x + dF/deta*eta_x) + dF/dzeta*zeta_x, the left-to-right order of rhs.jl:1679-1696, every derivative a
// sequential FMA chain in ascending m: bit-identical to k_elem_node and to the oracle.  Each role keeps only ITS
// direction's metric terms in registers for all equations (24 doubles; 40 + 8 node ids for the zeta role).  The roles
// run skewed on three rotating partial buffers and hand them on with producer/consumer named barriers (xi -> eta: A-x-full,
// eta -> zeta: A-y-full, zeta -> xi: A-empty), so no role waits for a step it does not depend on; two block barriers per
// element bracket the flux phase.  The state of the NEXT element's nodes is gathered into registers before the steps.
//
// Shared-memory tiles hold node (i,j,k) at  pos = (i ^ j) + 8 j + 66 k  (528 doubles per tile): with the xi and eta lanes
// ordered k-fastest every warp-wide access of all three roles and of the node-parallel flux phase is bank-conflict free
// (searched exhaustively over strides and XOR swizzles; a plain (i + S j + P k) layout always leaves one role 2-way
// conflicted).
//
// Records (layout 7), one per element, every stream one coalesced run of 64 doubles in the consuming role's lane order:
//   [0,3N)    xi_q   at node m of the xi-pencil    stream q*N+m        lane k + N*j
//   [3N,6N)   eta_q  at node m of the eta-pencil                        lane k + N*i
//   [6N,9N)   zeta_q at node m of the zeta-pencil                       lane i + N*j
//   [9N,10N)  omega*J, [10N,11N) -(omega*J*Minv) at node m of the zeta-pencil
//   then int32 zeta-view node ids [N][64] and flux-view node ids [512].
#pragma once

namespace jx {

template <int NGL, class EQ>
struct ElemTriCfg {
    static constexpr int N = NGL, NC = NGL * NGL, NP = NGL * NGL * NGL, NEQ = EQ::NEQ;
    static_assert(NGL == 8, "k_elem_tri: nop = 7");
    static constexpr int NT = 3 * NC;
    static constexpr int TP = 528;                               // doubles per tile: max pos = 7 + 56 + 462 = 525
    static constexpr int NFLD = 3 * NEQ;
    static constexpr int R = (NP + NT - 1) / NT;
    static constexpr int OFF_A = NFLD * TP, OFF_S = OFF_A + 9 * TP;
    static constexpr size_t SMEM_BYTES = (size_t)(OFF_S + TP) * 8;
    static constexpr int NSTREAM = 11 * NGL;
    static constexpr int ZID_OFF = NSTREAM * NC * 8, FID_OFF = ZID_OFF + NGL * NC * 4;
    static constexpr int REC_BYTES = round_up(FID_OFF + NP * 4, 128);
};

__host__ __device__ inline int tri_pos(int i, int j, int k) { return (i ^ j) + 8 * j + 66 * k; }

struct TriRetileArgs {
    const double *src;        // one metric array [E, n, n, n], element fastest (device copy); slot >= 0
    const double *omega;
    const double *Minv;       // slot -2
    const int64_t *connijk;   // slot -1
    char *rec;
    int64_t nelem;
    int ngl, rec_bytes, zid_off, fid_off;
    int slot;                 // 0..8 metric term, 9 = Je (stored as omega*J), -1 = node ids, -2 = -(omega*J*Minv)
};

static __global__ void k_retile_tri(TriRetileArgs a) {
    const int n = a.ngl, nc = n * n, np = nc * n;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= a.nelem * np) return;
    const int64_t iel = tid % a.nelem;
    const int l = (int)(tid / a.nelem);
    const int i = l % n, j = (l / n) % n, k = l / nc;
    char *rec = a.rec + (size_t)iel * a.rec_bytes;
    double *st = reinterpret_cast<double *>(rec);
    int32_t *zid = reinterpret_cast<int32_t *>(rec + a.zid_off);
    int32_t *fid = reinterpret_cast<int32_t *>(rec + a.fid_off);
    const size_t src = (size_t)iel + (size_t)a.nelem * l;
    const int cz = i + n * j;
    if (a.slot == -1) {
        const int32_t ip = (int32_t)(a.connijk[src] - 1);
        zid[k * nc + cz] = ip;
        fid[l] = ip;
    } else if (a.slot == -2) {
        st[(size_t)(10 * n + k) * nc + cz] = -(st[(size_t)(9 * n + k) * nc + cz] * a.Minv[zid[k * nc + cz]]);
    } else if (a.slot < 3) {
        st[(size_t)(a.slot * n + i) * nc + (k + n * j)] = a.src[src];
    } else if (a.slot < 6) {
        st[(size_t)(3 * n + (a.slot - 3) * n + j) * nc + (k + n * i)] = a.src[src];
    } else if (a.slot < 9) {
        st[(size_t)(6 * n + (a.slot - 6) * n + k) * nc + cz] = a.src[src];
    } else {
        const double wjk = a.omega[j] * a.omega[k];      // rhs.jl:1636-1643
        st[(size_t)(9 * n + k) * nc + cz] = a.omega[i] * wjk * a.src[src];
    }
}

// MODE = 0: rhs_el store (deterministic DSS); 1: RED.ADD of omega*J-weighted values; 2: RED.ADD with M^-1 pre-folded
template <int NGL, class EQ, int MODE>
static __global__ void __launch_bounds__(ElemTriCfg<NGL, EQ>::NT, 2)
k_elem_tri(const __grid_constant__ ElemArgs a) {
    using C = ElemTriCfg<NGL, EQ>;
    constexpr int N = NGL, NC = C::NC, NP = C::NP, NEQ = C::NEQ, NT = C::NT, TP = C::TP, R = C::R;
    constexpr int NQ = NEQ - (EQ::FLUX_QMASK == ((1u << (NEQ - 1)) - 1u) ? 1 : 0);
    static_assert(EQ::SRC_EQ >= -1 && EQ::HAS_AUX, "k_elem_tri uses the two-stage flux functors with one source component");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *X = reinterpret_cast<double *>(smem_raw);     // [NEQ][3][TP]: F_e, G_e, H_e
    double *A = X + C::OFF_A;                             // [3 buffers][3][TP] partial sums
    double *Sf = X + C::OFF_S;                            // [TP] source of equation SRC_EQ

    const int t = threadIdx.x;
    const int role = t / NC;                              // 0: xi, 1: eta, 2: zeta (two warps each)
    const int c = t % NC, c0 = c % N, c1 = c / N;
    // position of node m of this lane's line: base + off(m)
    //   xi   (j = c1, k = c0): (m ^ j) + 8 j + 66 k
    //   eta  (i = c1, k = c0): (i ^ m) + 8 m + 66 k
    //   zeta (i = c0, j = c1): (i ^ j) + 8 j + 66 m
    int lp[N];
#pragma unroll
    for (int m = 0; m < N; ++m)
        lp[m] = role == 0 ? (m ^ c1) + 8 * c1 + 66 * c0 : (role == 1 ? (c1 ^ m) + 8 * m + 66 * c0 : (c0 ^ c1) + 8 * c1 + 66 * m);
    auto lpos = [&](int m) -> int { return lp[m]; };
    constexpr bool fold = MODE == 2;
#define JX_D(m, i) a.dpsi[(m) + NGL * (i)]
    auto fid_of = [&](int64_t el) { return reinterpret_cast<const int32_t *>(a.rec + (size_t)el * C::REC_BYTES + C::FID_OFF); };
    int fidn[R];
    if ((int64_t)blockIdx.x < a.nelem) {
        const int32_t *fi = fid_of(blockIdx.x);
#pragma unroll
        for (int r = 0; r < R; ++r) fidn[r] = r * NT + t < NP ? __ldcs(fi + r * NT + t) : 0;
    }
    // named barriers (128 threads each: the 64 producers arrive, the 64 consumers wait); buffer b = global step % 3
    constexpr int BAR_XF = 1, BAR_YF = 4, BAR_EM = 7;
#define JX_BAR_SYNC3(BASE, b)                                                                  \
    do {                                                                                       \
        if ((b) == 0) asm volatile("bar.sync %0, 128;" ::"n"(BASE) : "memory");                 \
        else if ((b) == 1) asm volatile("bar.sync %0, 128;" ::"n"(BASE + 1) : "memory");        \
        else asm volatile("bar.sync %0, 128;" ::"n"(BASE + 2) : "memory");                      \
    } while (0)
#define JX_BAR_ARRIVE3(BASE, b)                                                                \
    do {                                                                                       \
        if ((b) == 0) asm volatile("bar.arrive %0, 128;" ::"n"(BASE) : "memory");               \
        else if ((b) == 1) asm volatile("bar.arrive %0, 128;" ::"n"(BASE + 1) : "memory");      \
        else asm volatile("bar.arrive %0, 128;" ::"n"(BASE + 2) : "memory");                    \
    } while (0)
    if (role == 2) {            // all three partial buffers start empty
        JX_BAR_ARRIVE3(BAR_EM, 0);
        JX_BAR_ARRIVE3(BAR_EM, 1);
        JX_BAR_ARRIVE3(BAR_EM, 2);
    }
    constexpr int NCMP = NQ + EQ::NAUX;
    double qa[R][NCMP];         // state of this thread's flux-view nodes of the element about to be processed
    auto gather = [&](const int(&id)[R]) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r * NT + t < NP) {
#pragma unroll
                for (int e = 0; e < NQ; ++e) qa[r][e] = __ldg(a.u + (size_t)e * a.npoin + id[r]);
#pragma unroll
                for (int x = 0; x < EQ::NAUX; ++x) qa[r][NQ + x] = __ldg(a.aux + (size_t)x * a.npoin + id[r]);
            }
        }
    };
    if ((int64_t)blockIdx.x < a.nelem) gather(fidn);
    int sg = 0;                 // global step counter modulo 3 at equation 0 of the current element
    for (int64_t el = blockIdx.x; el < a.nelem; el += gridDim.x) {
        const char *rec = a.rec + (size_t)el * C::REC_BYTES;
        const double *st = reinterpret_cast<const double *>(rec);
        double M[3][N], wj[N];
        int ip[N];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int m = 0; m < N; ++m) M[q][m] = __ldcs(st + (size_t)(role * 3 * N + q * N + m) * NC + c);
        if (role == 2) {
            const int32_t *zid = reinterpret_cast<const int32_t *>(rec + C::ZID_OFF);
#pragma unroll
            for (int m = 0; m < N; ++m) {
                wj[m] = __ldcs(st + (size_t)((fold ? 10 : 9) * N + m) * NC + c);
                ip[m] = __ldcs(zid + m * NC + c);
            }
        }
        const int64_t en = el + gridDim.x;          // next element: flux-view ids -> registers, record -> L2
        if (en < a.nelem) {
            const int32_t *fi = fid_of(en);
#pragma unroll
            for (int r = 0; r < R; ++r) fidn[r] = r * NT + t < NP ? __ldcs(fi + r * NT + t) : 0;
            if (t < 3) prefetch_l2_bulk(a.rec + (size_t)en * C::REC_BYTES + t * (C::REC_BYTES / 3), C::REC_BYTES / 3);
        }
        __syncthreads();        // every role is done with the flux tiles of the previous element
        // ---- flux / source at every node, node-parallel (node l = r*NT + t), from the gathered registers ----
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int l = r * NT + t;
            if (l < NP) {
                double q[NEQ], ax[EQ::NAUX], f[NEQ], gg[NEQ], h[NEQ];
#pragma unroll
                for (int e = 0; e < NEQ; ++e) q[e] = e < NQ ? qa[r][e < NQ ? e : 0] : 1.0;
#pragma unroll
                for (int x = 0; x < EQ::NAUX; ++x) ax[x] = qa[r][NQ + x];
                EQ::flux_aux(a.phys, q, ax, f, gg, h);
                const int p = tri_pos(l % N, (l / N) % N, l / NC);
#pragma unroll
                for (int e = 0; e < NEQ; ++e) {
                    X[(e * 3 + 0) * TP + p] = f[e];
                    X[(e * 3 + 1) * TP + p] = gg[e];
                    X[(e * 3 + 2) * TP + p] = h[e];
                }
                if constexpr (EQ::SRC_EQ >= 0) Sf[p] = a.lsource ? EQ::source_aux(a.phys, q, ax) : 0.0;
            }
        }
        if (en < a.nelem) gather(fidn);             // in flight during the equation steps
        __syncthreads();        // flux tiles complete
        // ---- NEQ steps per role, handed on through the partial buffers ----
#pragma unroll 1
        for (int e = 0; e < NEQ; ++e) {
            int b = sg + e;
            b -= b >= 3 ? 3 : 0;
            b -= b >= 3 ? 3 : 0;
            double *Ab = A + (size_t)(b * 3) * TP;
            if (role == 0) JX_BAR_SYNC3(BAR_EM, b);
            else if (role == 1) JX_BAR_SYNC3(BAR_XF, b);
            else JX_BAR_SYNC3(BAR_YF, b);
            double racc[N];
#pragma unroll
            for (int Xf = 0; Xf < 3; ++Xf) {
                const double *T = X + (size_t)(e * 3 + Xf) * TP;
                double f[N], d[N];
#pragma unroll
                for (int m = 0; m < N; ++m) f[m] = T[lpos(m)];
#pragma unroll
                for (int o = 0; o < N; ++o) d[o] = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m)
#pragma unroll
                    for (int o = 0; o < N; ++o) d[o] = fma(JX_D(m, o), f[m], d[o]);
                double *Ax = Ab + (size_t)Xf * TP;
                if (role == 0) {
#pragma unroll
                    for (int o = 0; o < N; ++o) Ax[lpos(o)] = d[o] * M[Xf][o];
                } else if (role == 1) {
#pragma unroll
                    for (int o = 0; o < N; ++o) Ax[lpos(o)] = Ax[lpos(o)] + d[o] * M[Xf][o];
                } else {
#pragma unroll
                    for (int o = 0; o < N; ++o) {
                        const double dXdx = Ax[lpos(o)] + d[o] * M[Xf][o];
                        racc[o] = Xf == 0 ? dXdx : racc[o] + dXdx;      // (dFdx + dGdy) + dHdz
                    }
                }
            }
            if (role == 0) JX_BAR_ARRIVE3(BAR_XF, b);
            else if (role == 1) JX_BAR_ARRIVE3(BAR_YF, b);
            else {
                double Sv[N];
#pragma unroll
                for (int o = 0; o < N; ++o) Sv[o] = 0.0;
                if constexpr (EQ::SRC_EQ >= 0) {
                    if (e == EQ::SRC_EQ) {
#pragma unroll
                        for (int o = 0; o < N; ++o) Sv[o] = Sf[lpos(o)];
                    }
                }
                JX_BAR_ARRIVE3(BAR_EM, b);          // the partial buffer is consumed (racc holds the sums)
                double *due = a.du + (size_t)e * a.npoin;
                double *rhe = MODE == 0 ? a.rhs_el + ((size_t)el * NEQ + e) * NP + c : nullptr;
#pragma unroll
                for (int o = 0; o < N; ++o) {
                    if constexpr (MODE == 2) atomicAdd(due + ip[o], wj[o] * (racc[o] - Sv[o]));   // wj = -(omega*J*Minv)
                    else {
                        const double out = 0.0 - wj[o] * (racc[o] - Sv[o]);
                        if constexpr (MODE == 0) rhe[NC * o] = out;
                        else atomicAdd(due + ip[o], out);
                    }
                }
            }
        }
        sg += NEQ % 3;
        sg -= sg >= 3 ? 3 : 0;
    }
#undef JX_BAR_SYNC3
#undef JX_BAR_ARRIVE3
#undef JX_D
}

}  // namespace jx
