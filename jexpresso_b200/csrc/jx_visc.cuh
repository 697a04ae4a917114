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

}  // namespace jx
