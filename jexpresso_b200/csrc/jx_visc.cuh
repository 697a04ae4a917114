// jx_visc.cuh -- k_visc_team: the AV viscous term (rhs.jl:1374-1461, 2794-2867) of 3D nop = 4 elements as a warp-team
// pass of its own, launched after the inviscid element kernel; bit-identical to the viscous pass of k_elem_node.
//
//     q_e = user_primitives!(u)                                      per node, equation e with mu_e != 0
//     (dq/dxi, dq/deta, dq/dzeta) = D-contractions of q_e            sequential FMA chains, ascending m
//     dqdx = mu*(dq/dxi*xi_x + dq/deta*eta_x + dq/dzeta*zeta_x), dqdy, dqdz
//     G_xi = (xi_x*dqdx + xi_y*dqdy + xi_z*dqdz)*wJ, G_eta, G_zeta    all nine metric terms of the node meet here
//     rhs_visc(i,j,k) = a_xi + a_eta + a_zeta,  a_xi = sum_m fma(-D[i,m], G_xi(m,j,k), .)  etc.
//
// The node-local step needs every metric term of the node in one thread, which neither role of the inviscid team kernel
// has (its plane lanes hold xi_X, eta_X of one direction X, its zeta lanes zeta_X and wJ).  Here the roles are
//   * LINE OWNERS (two warps, one per element of the pair; lane (i,j)): all nine metric terms, wJ and M^-1 at the five
//     nodes of their zeta line in registers (55 doubles) for all equations; they take dq/dxi, dq/deta from shared memory,
//     form dq/dzeta from their line, do the node-local step, write G_xi, G_eta over dq/dxi, dq/deta (same thread, same
//     nodes), contract G_zeta along their own line in registers, and later add a_zeta to (a_xi + a_eta) and scatter;
//   * PLANE LANES (one warp; lane (element, equation of the half, k)): metric free -- the 25 values of plane k in
//     registers feed both in-plane contractions (250 FMAs per 25 loads), forward on q_e, backward on G_xi, G_eta.
// 13 doubles per node and equation cross shared memory.  The viscous equations are split into two halves A, B and the
// roles alternate on them:   P: F(A) F(B) B(A) B(B) -      L: -  C(A) C(B) E(A) E(B)     (F forward, C combine,
// B backward, E end), one block barrier per slot.  Equations with mu_e = 0 are skipped: their reference contribution is
// +-0 (eq. 1 of every CompEuler deck), so the sums keep their values.
//
// Viscous records, one per element pair (built by k_retile_visc only when lvisc is set):
//   per element slot 11*N lane streams of 32 doubles: xi_x xi_y xi_z eta_x eta_y eta_z zeta_x zeta_y zeta_z wJ Minv at
//   node k of zeta line c = i + N*j; then int32 zeta-view node ids [EPB][N][32] and flux-view node ids [EPB*NP].
#pragma once

namespace jx {

template <int NGL, class EQ>
struct ViscTeamCfg {
    static constexpr int N = NGL, NC = NGL * NGL, NP = NGL * NGL * NGL, NEQ = EQ::NEQ;
    static constexpr int EPB = 2;
    static_assert(NGL == 5, "k_visc_team: nop = 4");
    static constexpr int NT = 96, NNODE = EPB * NP, R = (NNODE + NT - 1) / NT;
    static constexpr int GB = (EPB * NP + 12) / 16 * 16 + 3;     // as the team kernel: conflict-free plane and line accesses
    static constexpr int NMET = 11;
    // viscous pair records: NMET*N rows of ZROW = EPB*NC doubles (row m*N + k: metric m at node k of the zeta line, position
    // slot*NC + c), then int32 zeta-view node ids (N rows of ZROW) and flux-view node ids [EPB*NP]; rows packed to their lanes
    static constexpr int ZROW = EPB * NC;
    static constexpr int ZID_OFF = round_up(NMET * NGL * ZROW * 8, 16), FID_OFF = ZID_OFF + round_up(NGL * ZROW * 4, 16);
    static constexpr int GROUP_BYTES = round_up(FID_OFF + NNODE * 4, 128);
    static constexpr int NTILE = 4 * NEQ;                        // U, D_xi/G_xi/O, D_eta/G_eta, A_zeta per equation
    static constexpr size_t SMEM_BYTES = (size_t)NTILE * GB * 8;
    static constexpr int MAXREG = 168;
};

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

static __global__ void k_retile_visc(ViscRetileArgs a) {
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
    double *zs = reinterpret_cast<double *>(rec);
    int32_t *zid = reinterpret_cast<int32_t *>(rec + a.zid_off);
    int32_t *fid = reinterpret_cast<int32_t *>(rec + a.fid_off);
    const size_t src = (size_t)iel + (size_t)a.nelem * l;
    const int zrow = a.epb * nc, zpos = k * zrow + s * nc + i + n * j;
    if (a.slot == -1) {
        const int32_t ip = (int32_t)(a.connijk[src] - 1);
        zid[zpos] = ip;
        fid[s * np + l] = ip;
    } else if (a.slot == -2) {
        zs[(size_t)10 * n * zrow + zpos] = a.Minv[zid[zpos]];
    } else if (a.slot < 9) {
        zs[(size_t)a.slot * n * zrow + zpos] = a.src[src];
    } else {
        const double wjk = a.omega[j] * a.omega[k];      // the weight every record layout stores: omega_i*(omega_j*omega_k)*Je
        zs[(size_t)9 * n * zrow + zpos] = a.omega[i] * wjk * a.src[src];
    }
}

struct ViscArgs {
    const char *rec;          // viscous pair records
    double *out_el;           // MODE 0: rhs_el_visc [E][NEQ][NP]
    int nv;                   // number of equations with mu != 0
    int ve[8];                // their indices, ascending
};

// MODE 0: store rhs_el_visc (deterministic DSS); MODE 2: RED.ADD of Minv-scaled values into du
template <int NGL, class EQ, int MODE>
static __global__ void __maxnreg__((ViscTeamCfg<NGL, EQ>::MAXREG))
k_visc_team(const __grid_constant__ ElemArgs a, const __grid_constant__ ViscArgs va) {
    using C = ViscTeamCfg<NGL, EQ>;
    constexpr int N = NGL, NC = C::NC, NP = C::NP, NEQ = C::NEQ, NT = C::NT, R = C::R, GB = C::GB, EPB = C::EPB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *S0 = reinterpret_cast<double *>(smem_raw);
    // tiles of one kind are adjacent (stride GB = 3 mod 16), so the plane lanes (k, equation of the half, element) fall on
    // distinct banks exactly like the (k, X, element) lanes of the inviscid team kernel
    auto Ut = [&](int v) { return S0 + (size_t)v * GB; };                // primitives of viscous equation v
    auto Dx = [&](int v) { return S0 + (size_t)(NEQ + v) * GB; };        // dq/dxi  -> G_xi -> a_xi + a_eta
    auto De = [&](int v) { return S0 + (size_t)(2 * NEQ + v) * GB; };    // dq/deta -> G_eta
    auto Az = [&](int v) { return S0 + (size_t)(3 * NEQ + v) * GB; };    // a_zeta

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool owner = warp < 2;                     // line owners: warp = element slot
    const bool lact = lane < NC;
    const int c = lact ? lane : 0;
    const int nv = va.nv, nA = (nv + 1) / 2;         // halves: A = [0, nA), B = [nA, nv)
    // plane lanes: lane = k + N*(vh + VH*slot), vh = equation within the half
    const int VH = nA;                                // lanes per slot = VH*N (<= 15 for nv <= 6)
    const int pk = lane % N, pvh = (lane / N) % (VH > 0 ? VH : 1), ps = lane / (N * (VH > 0 ? VH : 1));
    const bool pact = lane < EPB * VH * N;
#define JX_D(m, i) a.dpsi[(m) + NGL * (i)]
    const int64_t ngroups = (a.nelem + EPB - 1) / EPB;
    bool staged = false;
    int fidn[R];              // flux-view node ids of the NEXT pair (loaded a pair ahead, so that the staging never waits for them)
    auto load_fid = [&](int64_t gg) {
        if (gg < ngroups) {
            const int32_t *fid = reinterpret_cast<const int32_t *>(va.rec + (size_t)gg * C::GROUP_BYTES + C::FID_OFF);
#pragma unroll
            for (int r = 0; r < R; ++r) fidn[r] = r * NT + t < C::NNODE ? __ldcs(fid + r * NT + t) : 0;
        }
    };
    auto issue_gathers = [&](int64_t gg) {
        const int nn = (int)(a.nelem - gg * EPB < EPB ? a.nelem - gg * EPB : EPB) * NP;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = r * NT + t;
            if (n < nn) {
#pragma unroll
                for (int e = 0; e < NEQ; ++e) cp_async8(Ut(e) + n, a.u + (size_t)e * a.npoin + fidn[r]);
            }
        }
        cp_async_commit();
        staged = true;
    };
    load_fid(blockIdx.x);
    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int cnt = (int)(a.nelem - g * EPB < EPB ? a.nelem - g * EPB : EPB);
        const char *rec = va.rec + (size_t)g * C::GROUP_BYTES;
        // ---- line owners: metric terms of the line (in flight during the primitives phase and F(A)) ----
        double M[10][N], minv[N];
        int ip[N];
        if (owner) {
            const double *zs = reinterpret_cast<const double *>(rec) + warp * NC + c;
            const int32_t *zid = reinterpret_cast<const int32_t *>(rec + C::ZID_OFF) + warp * NC + c;
#pragma unroll
            for (int m = 0; m < 10; ++m)
#pragma unroll
                for (int k = 0; k < N; ++k) M[m][k] = __ldcs(zs + (m * N + k) * C::ZROW);
#pragma unroll
            for (int k = 0; k < N; ++k) {
                minv[k] = MODE == 2 ? __ldcs(zs + (10 * N + k) * C::ZROW) : 1.0;
                ip[k] = __ldcs(zid + k * C::ZROW);
            }
        }
        // ---- primitives at every node of the pair, node-parallel ----
        // The state of the pair's nodes was gathered one pair ahead with cp.async into the U tiles (component e of node n at
        // U(e)[n], dead after the second combine slot): no registers, no exposed gather latency.  Every thread reads its own
        // nodes' staged values and overwrites position n of the tiles with the primitives.  (PERT functors read the
        // reference state as well: direct loads.)
        {
            if constexpr (!EQ::NEEDS_QE) {
                if (!staged) issue_gathers(g);       // first pair of this CTA
                load_fid(g + gridDim.x);             // ids of the next pair: used by the staging in slot 3
                cp_async_wait<0>();
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = r * NT + t;
                if (n < cnt * NP) {
                    double q[NEQ], qe[NEQ + 1], up[NEQ];
                    if constexpr (!EQ::NEEDS_QE) {
#pragma unroll
                        for (int e = 0; e < NEQ; ++e) q[e] = Ut(e)[n];
#pragma unroll
                        for (int e = 0; e <= NEQ; ++e) qe[e] = 0.0;
                    } else {
                        const int64_t node = __ldcs(reinterpret_cast<const int32_t *>(rec + C::FID_OFF) + n);
#pragma unroll
                        for (int e = 0; e < NEQ; ++e) q[e] = __ldg(a.u + (size_t)e * a.npoin + node);
#pragma unroll
                        for (int e = 0; e <= NEQ; ++e) qe[e] = __ldg(a.qe + (size_t)e * a.npoin + node);
                    }
                    EQ::primitives(a.phys, q, qe, up);
                    // (this thread's other nodes sit at other positions of the tiles: overwriting position n is safe)
                    for (int v = 0; v < nv; ++v) {
                        double val = up[0];
#pragma unroll
                        for (int e = 1; e < NEQ; ++e) val = va.ve[v] == e ? up[e] : val;
                        Ut(v)[n] = val;
                    }
                }
            }
            if (t < 2 && g + gridDim.x < ngroups)      // next pair's record towards L2
                prefetch_l2_bulk(va.rec + (size_t)(g + gridDim.x) * C::GROUP_BYTES + t * (C::GROUP_BYTES / 2 / 16 * 16), C::GROUP_BYTES / 2 / 16 * 16);
        }
        __syncthreads();
#pragma unroll 1
        for (int slot_t = 0; slot_t < 5; ++slot_t) {
            if constexpr (!EQ::NEEDS_QE) {
                if (slot_t == 3) {        // the U tiles are dead (both combine slots are over): stage the next pair's state there
                    if (g + gridDim.x < ngroups) issue_gathers(g + gridDim.x);
                    else staged = false;
                }
            }
            if (!owner) {
                // =============================== PLANE LANES ===============================
                // slot 0: F(A), 1: F(B), 2: B(A), 3: B(B)
                const int half = slot_t & 1, v = half * nA + pvh;
                const bool live = pact && ps < cnt && slot_t < 4 && v < nv && (half == 0 || pvh < nv - nA);
                if (live) {
                    const int poff = ps * NP + NC * pk;
                    double w[NC];
                    if (slot_t < 2) {
                        const double *T = Ut(v) + poff;
                        double *Ox = Dx(v) + poff, *Oe = De(v) + poff;
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
                    } else {
                        double *Gx = Dx(v) + poff;
                        const double *Ge = De(v) + poff;
                        double ax[NC];
#pragma unroll
                        for (int n = 0; n < NC; ++n) w[n] = Gx[n];
#pragma unroll
                        for (int n = 0; n < NC; ++n) ax[n] = 0.0;
#pragma unroll
                        for (int m = 0; m < N; ++m)
#pragma unroll
                            for (int n = 0; n < NC; ++n) ax[n] = fma(-JX_D(n % N, m), w[N * (n / N) + m], ax[n]);
#pragma unroll
                        for (int n = 0; n < NC; ++n) w[n] = Ge[n];
#pragma unroll
                        for (int c0 = 0; c0 < NC; c0 += N) {
                            double ae[N];
#pragma unroll
                            for (int u = 0; u < N; ++u) ae[u] = 0.0;
#pragma unroll
                            for (int m = 0; m < N; ++m)
#pragma unroll
                                for (int u = 0; u < N; ++u) {
                                    const int n = c0 + u, i = n % N, j = n / N;
                                    ae[u] = fma(-JX_D(j, m), w[N * m + i], ae[u]);
                                }
#pragma unroll
                            for (int u = 0; u < N; ++u) Gx[c0 + u] = ax[c0 + u] + ae[u];
                        }
                    }
                }
            } else if (slot_t >= 1) {
                // =============================== LINE OWNERS ===============================
                // slot 1: C(A), 2: C(B), 3: E(A), 4: E(B)
                const bool combine = slot_t < 3;
                const int half = (slot_t - 1) & 1;
                const int v0 = half == 0 ? 0 : nA, v1 = half == 0 ? nA : nv;
                if (lact && warp < cnt) {
                    for (int v = v0; v < v1; ++v) {
                        const int e = va.ve[v];
                        const int lo = warp * NP + c;
                        if (combine) {
                            const double mu = a.visc[e];
                            const double *Uq = Ut(v) + lo;
                            double *Gx = Dx(v) + lo, *Ge = De(v) + lo;
                            double uq[N], dz[N], gz[N];
#pragma unroll
                            for (int m = 0; m < N; ++m) uq[m] = Uq[NC * m];
#pragma unroll
                            for (int k = 0; k < N; ++k) dz[k] = 0.0;
#pragma unroll
                            for (int m = 0; m < N; ++m)
#pragma unroll
                                for (int k = 0; k < N; ++k) dz[k] = fma(JX_D(m, k), uq[m], dz[k]);
#pragma unroll
                            for (int k = 0; k < N; ++k) {
                                const double dqdxi = Gx[NC * k], dqdeta = Ge[NC * k], dqdzeta = dz[k];
                                double auxi = dqdxi * M[0][k] + dqdeta * M[3][k] + dqdzeta * M[6][k];
                                const double dqdx = mu * auxi;
                                auxi = dqdxi * M[1][k] + dqdeta * M[4][k] + dqdzeta * M[7][k];
                                const double dqdy = mu * auxi;
                                auxi = dqdxi * M[2][k] + dqdeta * M[5][k] + dqdzeta * M[8][k];
                                const double dqdz = mu * auxi;
                                const double wJ = M[9][k];
                                Gx[NC * k] = (M[0][k] * dqdx + M[1][k] * dqdy + M[2][k] * dqdz) * wJ;
                                Ge[NC * k] = (M[3][k] * dqdx + M[4][k] * dqdy + M[5][k] * dqdz) * wJ;
                                gz[k] = (M[6][k] * dqdx + M[7][k] * dqdy + M[8][k] * dqdz) * wJ;
                            }
                            double az[N];
#pragma unroll
                            for (int k = 0; k < N; ++k) az[k] = 0.0;
#pragma unroll
                            for (int m = 0; m < N; ++m)
#pragma unroll
                                for (int k = 0; k < N; ++k) az[k] = fma(-JX_D(k, m), gz[m], az[k]);
                            double *Azp = Az(v) + lo;
#pragma unroll
                            for (int k = 0; k < N; ++k) Azp[NC * k] = az[k];
                        } else {
                            const double *Ox = Dx(v) + lo, *Azp = Az(v) + lo;
                            double *due = a.du + (size_t)e * a.npoin;
                            double *oel = nullptr;
                            if constexpr (MODE == 0) {
                                const int64_t eo = a.eorig ? (int64_t)__ldg(a.eorig + g * EPB + warp) : g * EPB + warp;
                                oel = va.out_el + ((size_t)eo * NEQ + e) * NP + c;
                            }
#pragma unroll
                            for (int k = 0; k < N; ++k) {
                                const double outv = Ox[NC * k] + Azp[NC * k];      // (a_xi + a_eta) + a_zeta
                                if constexpr (MODE == 0) oel[NC * k] = outv;
                                else atomicAdd(due + ip[k], outv * minv[k]);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
#undef JX_D
}


// ==========================================================================================================
// k_visc_quad -- second form of the AV viscous pass (variant 13): four warps per element pair, every phase on all of them.
//
// k_visc_team keeps all nine metric terms of a zeta line in its two line-owner warps (110 registers of metrics, 168 per
// thread, three warps per CTA, one plane warp of 20 lanes on the critical path: barrier 16 %, long scoreboard and local-memory
// reloads on top, profiles/r02c).  Here the node-local step -- the only place where all nine metric terms of a node meet -- is
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
// per SM.  Every sum keeps the reference's order: bit-identical to k_visc_team, k_elem_node<VISC> and the oracle.
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
    // section 5: interface pairs first, interior pairs beside the exchange).  CTAs of the interior launch that land on one of
    // the first a.reserve_sms SMs leave at once, at most a.exit_budget of them (the launch's surplus).
    const int64_t nl = a.glist ? (int64_t)a.nlist : ngroups;
    auto pair_at = [&](int64_t li) -> int64_t { return li < nl ? (a.glist ? (int64_t)a.glist[li] : li) : ngroups; };
    if (a.reserve_sms > 0) {
        __shared__ int s_exit;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if ((int)smid < a.reserve_sms) {
            if (t == 0) s_exit = atomicAdd(a.exit_ctr, 1) < a.exit_budget ? 1 : 0;
            __syncthreads();
            if (s_exit) return;
        }
    }
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
