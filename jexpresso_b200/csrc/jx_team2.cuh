// jx_team2.cuh -- k_elem_team2: the round-2 warp-team element kernel (3D, inviscid, exact order).
//
// Same arithmetic and the same element-pair records (layout 5) as k_elem_team variant 9 -- two plane warps, two
// zeta warps, partial sums B_X = dX/dxi*xi_X + dX/deta*eta_X handed from the plane role to the zeta role, which
// finishes (B_X + dX/dzeta*zeta_X), sums F,G,H parts left to right and scatters (rhs.jl:1615-1698) -- so results are
// bit-identical to it, to k_elem_node and to the oracle.  What changed is how data moves (profiles/r01i_lsu_breakdown.md:
// 1499 LSU wavefronts per element, of which 310 for the cp.async-staged gathers; barrier stalls 20 %):
//
//  * NODE IMAGE + TMA.  The pre-pass (k_node_image) writes one 16-byte-multiple row per unique node,
//    w[node] = (q components the flux reads, per-node equation-of-state values), and the kernel pulls the rows of the
//    next pair into shared memory with one cp.async.bulk (TMA, mbarrier completion) per node: no LSU wavefronts for the
//    gather at all (scripts/micro/gather.cu: 66 wavefronts per element for reading the rows back against 308).
//  * RING OF EQUATION SLOTS.  Flux tiles live in a ring of NEQ+1 slots of three tiles (F_e, G_e, H_e), equation e of
//    the CTA's it-th pair in slot (NEQ*it + e) mod (NEQ+1): the flux phase of pair it+1 never touches the slot of
//    equation NEQ-1 of pair it, so the zeta role's last step of a pair runs beside the plane role's first step of the
//    next one -- the roles stay skewed by one step ACROSS pairs and neither idles one step in NEQ+1.
//  * The rows of pair it+1 land in the two slots that are already dead at step 2 of pair it (the spare slot and the
//    slot of equation 0), which are exactly the slots equations 0 and 1 of pair it+1 will occupy.
//  * PRODUCER/CONSUMER BARRIERS.  One block-wide barrier per pair (between reading the rows and writing the flux
//    tiles); everything else is bar.arrive / bar.sync pairs on named barriers (B-full, B-empty, X-full), so a role
//    never waits for a step it does not depend on.
#pragma once

namespace jx {

// WT = false: node-image rows land in the (dead) ring slots the pair's equations 0 and 1 will occupy -- 51.8 KB of shared
//              memory, 4 CTAs of 128 registers per SM, one block barrier between reading the rows and writing the tiles;
// WT = true:  dedicated row tile -- 64 KB, 3 CTAs of 168 registers per SM, no barrier inside the flux phase.
template <int NGL, class EQ, bool WT = false>
struct ElemTeam2Cfg {
    using L = ElemTeamCfg<NGL, EQ, 32 / (3 * NGL), 2>;          // record layout 5 (shared with k_elem_team variant 9)
    static constexpr int N = NGL, NC = NGL * NGL, NP = NGL * NGL * NGL, NEQ = EQ::NEQ;
    static constexpr int EPB = L::EPB;
    static_assert(EPB == 2, "k_elem_team2: element pairs (nop = 4)");
    static constexpr int NT = 128, NNODE = EPB * NP, R = (NNODE + NT - 1) / NT;
    static constexpr int GB = L::GB;
    static constexpr int SLOT_D = 3 * GB;                        // doubles per equation slot (F_e, G_e, H_e tiles)
    static constexpr int NXS = NEQ + 1;                          // flux-tile ring
    static constexpr int NQ = L::NQ, NCOMP = L::NCOMP;
    static constexpr int ROWD = (NCOMP + 1) / 2 * 2;             // doubles per node-image row (16-byte multiple)
    static constexpr int ROWB = ROWD * 8;
    static_assert(NP * ROWB + 8 <= SLOT_D * 8, "the node-image rows of one element must fit one equation slot");
    static_assert(EQ::SRC_EQ < NEQ - 1, "the source tile is single-buffered: the last equation must be source-free");
    static constexpr int OFF_B = NXS * SLOT_D, OFF_S = OFF_B + 2 * SLOT_D, OFF_W = (OFF_S + GB + 1) / 2 * 2;
    static constexpr int OFF_BAR = WT ? OFF_W + NNODE * ROWD : OFF_S + GB;
    static constexpr size_t SMEM_BYTES = (size_t)(OFF_BAR + 2) * 8;
    static constexpr int MAXREG = WT ? 168 : 128;
    static constexpr bool WTILE = WT;
};

// per-unique-node pre-pass of k_elem_team2: node-image rows (+ zero-fill of the scatter target in atomics mode)
struct ImageArgs {
    const double *u, *qe;
    double *img;      // [npoin][ROWD]
    double *zero;     // du to clear (atomics mode) or nullptr
    int64_t npoin;
    int rowd;
    Phys phys;
};

template <class EQ>
static __global__ void __launch_bounds__(256) k_node_image(const __grid_constant__ ImageArgs a) {
    constexpr int NEQ = EQ::NEQ;
    constexpr int NQ = NEQ - (EQ::FLUX_QMASK == ((1u << (NEQ - 1)) - 1u) ? 1 : 0);
    constexpr int NCOMP = NQ + EQ::NAUX, ROWD = (NCOMP + 1) / 2 * 2;
    // rows are staged through shared memory so that a block stores its 256 rows as one contiguous, fully coalesced run
    // (a thread writing its own 48-byte row directly leaves every 32-byte sector half written per instruction)
    __shared__ double2 srow[256 * ROWD / 2];
    double *sr = reinterpret_cast<double *>(srow);
    const int t = threadIdx.x;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < a.npoin; base += (int64_t)gridDim.x * 256) {
        const int64_t ip = base + t;
        if (ip < a.npoin) {
            double q[NEQ], qe[NEQ + 1];
#pragma unroll
            for (int e = 0; e < NEQ; ++e) q[e] = (e < NQ || ((EQ::AUX_MASK >> e) & 1u)) ? a.u[(size_t)e * a.npoin + ip] : 0.0;
#pragma unroll
            for (int e = 0; e <= NEQ; ++e) qe[e] = (EQ::NEEDS_QE && (e == NEQ || ((EQ::AUX_MASK >> e) & 1u))) ? a.qe[(size_t)e * a.npoin + ip] : 0.0;
            double ax[EQ::NAUX > 0 ? EQ::NAUX : 1];
            EQ::aux(a.phys, q, qe, ax);
            // row t at sr[t*ROWD]: the 48-byte-stride column writes are 2-way conflicted, the 16-byte row-major reads below are not
#pragma unroll
            for (int e = 0; e < NQ; ++e) sr[t * ROWD + e] = q[e];
#pragma unroll
            for (int x = 0; x < EQ::NAUX; ++x) sr[t * ROWD + NQ + x] = ax[x];
#pragma unroll
            for (int x = NCOMP; x < ROWD; ++x) sr[t * ROWD + x] = 0.0;
            if (a.zero) {
#pragma unroll
                for (int e = 0; e < NEQ; ++e) a.zero[(size_t)e * a.npoin + ip] = 0.0;
            }
        }
        __syncthreads();
        const int64_t nrow = a.npoin - base < 256 ? a.npoin - base : 256;
        double2 *dst = reinterpret_cast<double2 *>(a.img + (size_t)base * ROWD);
        for (int x = t; x < nrow * (ROWD / 2); x += 256) dst[x] = srow[x];
        __syncthreads();
    }
}

// named barriers with immediate ids (a register id makes ptxas reserve all 16 barriers)
template <int ID>
__device__ __forceinline__ void bar_sync_id() { asm volatile("bar.sync %0, 128;" ::"n"(ID) : "memory"); }
template <int ID>
__device__ __forceinline__ void bar_arrive_id() { asm volatile("bar.arrive %0, 128;" ::"n"(ID) : "memory"); }
template <int ID0>
__device__ __forceinline__ void bar_sync_par(int par) {
    if (par) bar_sync_id<ID0 + 1>();
    else bar_sync_id<ID0>();
}
template <int ID0>
__device__ __forceinline__ void bar_arrive_par(int par) {
    if (par) bar_arrive_id<ID0 + 1>();
    else bar_arrive_id<ID0>();
}

// MODE = 0: rhs_el store (deterministic DSS); 1: RED.ADD of omega*J-weighted values; 2: RED.ADD with M^-1 pre-folded.
// DYN: list-driven launch of the interface-first split (see k_elem_team).
//
// Code layout: ONE loop over the CTA's pairs for both roles, so the flux phase and each role's step exist once in the
// instruction stream (the first version had three copies of the flux phase and of the zeta step: 55 KB of SASS against a
// 32 KB instruction cache, 15 % of the warp samples waiting for instructions).  The role metrics share one register
// array: the plane role reads it as xi_X, eta_X at its nodes of plane k, the zeta role as zeta_{x,y,z} and the weight
// at the nodes of its line.
template <int NGL, class EQ, int MODE, bool DYN = false, bool WT = false>
static __global__ void __maxnreg__((ElemTeam2Cfg<NGL, EQ, WT>::MAXREG))
k_elem_team2(const __grid_constant__ ElemArgs a) {
    using C = ElemTeam2Cfg<NGL, EQ, WT>;
    using L = typename C::L;
    constexpr int N = NGL, NC = C::NC, NP = C::NP, NEQ = C::NEQ, NT = C::NT, R = C::R, GB = C::GB, EPB = C::EPB;
    constexpr int NQ = C::NQ, NSTRZ = L::NSTRZ, SLOT_D = C::SLOT_D, NXS = C::NXS, ROWD = C::ROWD, ROWB = C::ROWB;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_XFULL = 5;     // named barriers (0 = whole CTA)
    constexpr int NSPLIT = L::NSPLIT, NMREG = 2 * NSPLIT > 4 * N ? 2 * NSPLIT : 4 * N;
    static_assert(EQ::SRC_EQ >= -1 && EQ::HAS_AUX, "team kernels use the two-stage flux functors with one source component");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *S0 = reinterpret_cast<double *>(smem_raw);
    double *Bt = S0 + C::OFF_B;                                   // [2][3][GB] partial sums
    double *Sf = S0 + C::OFF_S;                                   // [GB] source of equation SRC_EQ
    uint64_t *wbar = reinterpret_cast<uint64_t *>(S0 + C::OFF_BAR);

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool plane_warp = warp < 2;
    // plane role: lane = k + N*(X + 3*slot)
    const int pk = lane % N, pX = (lane / N) % 3, ps = lane / (3 * N);
    const bool pact = lane < L::NPL;
    const int poff = ps * NP + NC * pk;
    // zeta role: lane c = i + N*j of element slot zsl
    const bool zact = lane < NC;
    const int c = zact ? lane : 0;
    const int zsl = warp & 1;
#define JX_D(m, i) a.dpsi[(m) + NGL * (i)]
    const int64_t ngroups = (a.nelem + EPB - 1) / EPB;
    auto cnt_of = [&](int64_t gg) { return (int)(a.nelem - gg * EPB < EPB ? a.nelem - gg * EPB : EPB); };
    auto slot_ptr = [&](int xb, int e) -> double * {
        int sl = xb + e;
        sl -= sl >= NXS ? NXS : 0;
        return S0 + (size_t)sl * SLOT_D;
    };
    // rows of element slot s of a pair live in the slot its equation s will occupy (16-byte aligned)
    auto wrow = [&](int xb, int s, int row) -> unsigned char * {
        if constexpr (WT) return reinterpret_cast<unsigned char *>(S0 + C::OFF_W) + (size_t)(s * NP + row) * ROWB;
        unsigned char *p = reinterpret_cast<unsigned char *>(slot_ptr(xb, s));
        p += (smem_u32(p) & 8u);
        return p + (size_t)row * ROWB;
    };

    __shared__ int s_grp[4];
    int64_t g = blockIdx.x;
    int gnext = 0;
    if constexpr (DYN) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if ((int)smid < a.reserve_sms) return;
        if (t == 0) {
            const int p0 = atomicAdd(a.gctr, 1), p1 = atomicAdd(a.gctr, 1);
            s_grp[0] = p0 < a.nlist ? a.glist[p0] : (int)ngroups;
            s_grp[1] = p1 < a.nlist ? a.glist[p1] : (int)ngroups;
        }
    }
    if (t == 0) {
        mbar_init(wbar, 1);
        fence_proxy_async();
    }
    __syncthreads();
    if constexpr (DYN) {
        g = s_grp[0];
        gnext = s_grp[1];
    }
    auto next_of = [&](int64_t gg) -> int64_t {
        if constexpr (DYN) return gnext;
        else return gg + gridDim.x;
    };
    // Row runs of element slot zsl of group gg (zeta warp zsl fetches its own element's rows): lane j takes runs j, j+32,
    // ...  A bulk copy is a uniform-datapath instruction, issued lane by lane, so the number of copies matters:
    // consecutive node ids are merged on the host (build_row_runs).
    constexpr int RR = (L::MAXRUN + 31) / 32;
    int run_id[RR], run_rl[RR], run_n = 0;
    auto load_runs = [&](int64_t gg) {
        run_n = 0;
        if (gg < ngroups && zsl < cnt_of(gg)) {
            const char *rec = a.rec + (size_t)gg * L::GROUP_BYTES;
            run_n = __ldcs(reinterpret_cast<const int32_t *>(rec + L::NRUN_OFF) + zsl);
            const int32_t *ri = reinterpret_cast<const int32_t *>(rec + L::RUNI_OFF) + zsl * L::MAXRUN;
            const unsigned char *rr = reinterpret_cast<const unsigned char *>(rec + L::RUNR_OFF) + zsl * L::MAXRUN;
            const unsigned char *rl = reinterpret_cast<const unsigned char *>(rec + L::RUNL_OFF) + zsl * L::MAXRUN;
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const int j = r * 32 + lane;
                if (j < run_n) { run_id[r] = __ldcs(ri + j); run_rl[r] = (int)__ldcs(rr + j) | ((int)__ldcs(rl + j) << 8); }
            }
        }
    };
    // (the target slots were last touched through the generic proxy by steps this warp has itself completed or that the
    // hand-shakes have ordered before this point; like a TMA producer refilling a consumed stage, no proxy fence)
    auto issue_rows = [&](int64_t gg, int xb) {
        if (warp == 2 && lane == 0) mbar_expect_tx(wbar, (uint32_t)(cnt_of(gg) * NP * ROWB));
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            if (r * 32 + lane < run_n)
                tma_bulk_g2s(wrow(xb, zsl, run_rl[r] & 255), reinterpret_cast<const unsigned char *>(a.aux) + (size_t)run_id[r] * ROWB,
                             (uint32_t)((run_rl[r] >> 8) * ROWB), wbar);
        }
    };
    // row (within its element's tile) of this thread's flux-view nodes of group gg
    int wpn[R];
    auto load_wpos = [&](int64_t gg) {
        if (gg < ngroups) {
            const unsigned char *w = reinterpret_cast<const unsigned char *>(a.rec + (size_t)gg * L::GROUP_BYTES + L::WPOS_OFF);
#pragma unroll
            for (int r = 0; r < R; ++r) wpn[r] = r * NT + t < C::NNODE ? (int)__ldcs(w + r * NT + t) : 0;
        }
    };

    double M[NMREG];          // role metrics of the current pair (see above)
    int ipz[N];               // zeta role: node ids of the line
    int64_t e0z = 0;          // zeta role: first element of the pair the metrics belong to
    int cntz = 0;
    int xbase = 0, xprev = 0; // slot of equation 0 of the current / previous pair: (NEQ*it) mod NXS
    load_wpos(g);
    if (!plane_warp) {
        bar_arrive_id<BAR_EMPTY>();
        bar_arrive_id<BAR_EMPTY + 1>();
        load_runs(g);
        if (g < ngroups) issue_rows(g, 0);
    }
    for (int it = 0;; ++it) {
        const bool have = g < ngroups;
        if (!have && (plane_warp || it == 0)) break;
        const int cnt = have ? cnt_of(g) : 0;
        const int64_t gn = next_of(g);
        if (have) {
            if (plane_warp) {   // xi_X, eta_X at this warp's nodes of plane k (lane-major streams): in flight during the flux phase
                const double *pl = reinterpret_cast<const double *>(a.rec + (size_t)g * L::GROUP_BYTES);
                const int lo = warp == 0 ? 0 : NSPLIT;
#pragma unroll
                for (int n = 0; n < NSPLIT; ++n) {
                    const int nn = lo + n < NC ? lo + n : NC - 1;
                    M[n] = __ldcs(pl + nn * 32 + lane);
                    M[NSPLIT + n] = __ldcs(pl + (NC + nn) * 32 + lane);
                }
            }
            // ---------------- joint flux phase ----------------
            const int nn = cnt * NP;
            double comp[R][ROWD];
            if constexpr (WT) asm volatile("bar.sync 0;" ::: "memory");   // every step of the previous pair that used the target slots is done
            mbar_wait(wbar, (uint32_t)(it & 1));
            if constexpr (!WT) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int n = r * NT + t;
                    if (n < nn) {
                        const double2 *row = reinterpret_cast<const double2 *>(wrow(xbase, n / NP, wpn[r]));
#pragma unroll
                        for (int x = 0; x < ROWD / 2; ++x) {
                            const double2 v = row[x];
                            comp[r][2 * x] = v.x;
                            comp[r][2 * x + 1] = v.y;
                        }
                    }
                }
                load_wpos(gn);
            }
            // next group's record towards L2 in three pieces that leave out the weight stream this scatter mode does not
            // read (omega*J or its M^-1-folded twin); three lanes only -- a bulk prefetch is issued lane by lane as well
            if (warp == 1 && lane < 3 && gn < ngroups) {
                constexpr int ZB = NSTRZ * 256, SB = N * 256, U = MODE == 2 ? 3 : 4;   // bytes per slot block / stream set; unused set
                const int lo = lane == 0 ? 0 : L::Z_OFF + (lane - 1) * ZB + (U + 1) * SB;
                const int hi = lane == 2 ? L::GROUP_BYTES : L::Z_OFF + lane * ZB + U * SB;
                if (hi > lo) prefetch_l2_bulk(a.rec + (size_t)gn * L::GROUP_BYTES + lo, (uint32_t)(hi - lo));
            }
            if constexpr (!WT) asm volatile("bar.sync 0;" ::: "memory");   // rows read; every step of the previous pair that used the target slots is done
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = r * NT + t;
                if (n < nn) {
                    if constexpr (WT) {
                        const double2 *row = reinterpret_cast<const double2 *>(wrow(xbase, n / NP, wpn[r]));
#pragma unroll
                        for (int x = 0; x < ROWD / 2; ++x) {
                            const double2 v = row[x];
                            comp[r][2 * x] = v.x;
                            comp[r][2 * x + 1] = v.y;
                        }
                    }
                    double q[NEQ], ax[EQ::NAUX], f[NEQ], gg2[NEQ], h[NEQ];
#pragma unroll
                    for (int e = 0; e < NEQ; ++e) q[e] = e < NQ ? comp[r][e < NQ ? e : 0] : 1.0;
#pragma unroll
                    for (int x = 0; x < EQ::NAUX; ++x) ax[x] = comp[r][NQ + x];
                    EQ::flux_aux(a.phys, q, ax, f, gg2, h);
#pragma unroll
                    for (int e = 0; e < NEQ; ++e) {
                        double *T = slot_ptr(xbase, e) + n;
                        T[0] = f[e];
                        T[GB] = gg2[e];
                        T[2 * GB] = h[e];
                    }
                    if constexpr (EQ::SRC_EQ >= 0) Sf[n] = a.lsource ? EQ::source_aux(a.phys, q, ax) : 0.0;
                }
            }
            if constexpr (WT) load_wpos(gn);
        }
        if (plane_warp) {
            // =============================== PLANE ROLE ===============================
            bar_sync_id<BAR_XFULL>();          // the zeta warps' share of the flux tiles is in place too
            const bool live = pact && ps < cnt;
            auto plane_step = [&](auto lo_c, auto hi_c, int e, int par) {
                constexpr int LO = decltype(lo_c)::value, HI = decltype(hi_c)::value, NN = HI - LO;
                constexpr int CHUNK = JX_TEAM_CHUNK;
                const double *T = slot_ptr(xbase, e) + (size_t)pX * GB + poff;
                double *Bo = Bt + (size_t)(par * 3 + pX) * GB + poff;
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
                        if (c0 + u < NN) Bo[LO + c0 + u] = dx[u] * M[c0 + u] + de[u] * M[NSPLIT + c0 + u];
                }
            };
#pragma unroll 1
            for (int e = 0; e < NEQ; ++e) {
                const int par = (it + e) & 1;   // (NEQ*it + e) & 1 with NEQ odd
                bar_sync_par<BAR_EMPTY>(par);
                if constexpr (DYN) {
                    if (e == 1 && t == 0) {
                        const int p = atomicAdd(a.gctr, 1);
                        s_grp[2] = p < a.nlist ? a.glist[p] : (int)ngroups;
                    }
                }
                if (live) {
                    if (warp == 0) plane_step(std::integral_constant<int, 0>{}, std::integral_constant<int, NSPLIT>{}, e, par);
                    else plane_step(std::integral_constant<int, NSPLIT>{}, std::integral_constant<int, NC>{}, e, par);
                }
                bar_arrive_par<BAR_FULL>(par);
            }
        } else {
            // =============================== ZETA ROLE ===============================
            constexpr bool fold = MODE == 2;
            if (have) {
                if constexpr (WT) {
                    // dedicated row tile: once all four warps have left the flux phase the tile is free for the next
                    // pair's rows, which then have the whole pair to land
                    bar_sync_id<BAR_XFULL>();
                    load_runs(gn);
                    if (gn < ngroups) issue_rows(gn, 0);
                } else bar_arrive_id<BAR_XFULL>();
            }
            // steps of this iteration: the last equation of the previous pair (k = -1), then equations 0..NEQ-2 of this one
#pragma unroll 1
            for (int k = it > 0 ? -1 : 0; k < (have ? NEQ - 1 : 0); ++k) {
                const int e = k < 0 ? NEQ - 1 : k;
                const int xb = k < 0 ? xprev : xbase;
                const int par = (k < 0 ? it - 1 + NEQ - 1 : it + k) & 1;
                if (k == 0) {   // this pair's metrics and ids (the previous pair's last step is done with the old ones)
                    const char *rec = a.rec + (size_t)g * L::GROUP_BYTES;
                    const double *zs = reinterpret_cast<const double *>(rec + L::Z_OFF);
                    const int32_t *zid = reinterpret_cast<const int32_t *>(rec + L::ZID_OFF);
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int m = 0; m < N; ++m) M[q * N + m] = __ldcs(zs + (zsl * NSTRZ + q * N + m) * 32 + lane);
#pragma unroll
                    for (int m = 0; m < N; ++m) {
                        M[3 * N + m] = __ldcs(zs + (zsl * NSTRZ + (fold ? 4 : 3) * N + m) * 32 + lane);
                        ipz[m] = __ldcs(zid + (zsl * N + m) * 32 + lane);
                    }
                    e0z = g * EPB;
                    cntz = cnt;
                    if constexpr (!WT) load_runs(gn);
                }
                bar_sync_par<BAR_FULL>(par);
                if (zact && zsl < cntz) {
                    const double *Fe = slot_ptr(xb, e) + zsl * NP + c, *Ge = Fe + GB, *He = Ge + GB;
                    const double *Bf = Bt + (size_t)(par * 3) * GB + zsl * NP + c;
                    double f[N], gg2[N], h[N], b[3][N], Sv[N];
#pragma unroll
                    for (int m = 0; m < N; ++m) { f[m] = Fe[NC * m]; gg2[m] = Ge[NC * m]; h[m] = He[NC * m]; }
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int kk = 0; kk < N; ++kk) b[q][kk] = Bf[q * GB + NC * kk];
#pragma unroll
                    for (int kk = 0; kk < N; ++kk) Sv[kk] = 0.0;
                    if constexpr (EQ::SRC_EQ >= 0) {
                        if (e == EQ::SRC_EQ) {
#pragma unroll
                            for (int kk = 0; kk < N; ++kk) Sv[kk] = Sf[zsl * NP + c + NC * kk];
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
                            dG[o] = fma(JX_D(m, o), gg2[m], dG[o]);
                            dH[o] = fma(JX_D(m, o), h[m], dH[o]);
                        }
                    double *due = a.du + (size_t)e * a.npoin;
                    double *rhe = MODE == 0 ? a.rhs_el + ((size_t)(e0z + zsl) * NEQ + e) * NP + c : nullptr;
#pragma unroll
                    for (int kk = 0; kk < N; ++kk) {
                        const double dFdx = b[0][kk] + dF[kk] * M[kk];
                        const double dGdy = b[1][kk] + dG[kk] * M[N + kk];
                        const double dHdz = b[2][kk] + dH[kk] * M[2 * N + kk];
                        const double r = (dFdx + dGdy) + dHdz;
                        if constexpr (MODE == 2) atomicAdd(due + ipz[kk], M[3 * N + kk] * (r - Sv[kk]));   // weight = -(omega*J*Minv)
                        else {
                            const double out = 0.0 - M[3 * N + kk] * (r - Sv[kk]);
                            if constexpr (MODE == 0) rhe[NC * kk] = out;
                            else atomicAdd(due + ipz[kk], out);
                        }
                    }
                }
                bar_arrive_par<BAR_EMPTY>(par);
                // rows of the next pair: its equation-0 slot (the spare, free since the step k = -1) and its equation-1
                // slot (equation 0 of this pair, free now) -- both read for the last time by this very warp's role
                if (!WT && k == 0 && gn < ngroups) {
                    int xb1 = xbase + NEQ;
                    xb1 -= xb1 >= NXS ? NXS : 0;
                    asm volatile("bar.sync 6, 64;" ::: "memory");   // both zeta warps are done with step 0 (and with k = -1)
                    issue_rows(gn, xb1);
                }
            }
        }
        if (!have) break;
        xprev = xbase;
        xbase += NEQ;
        xbase -= xbase >= NXS ? NXS : 0;
        if constexpr (DYN) {
            // thread 0 wrote s_grp[2] at step 1; the hand-shakes of the later steps order that write before this read in
            // every warp, and the block barrier of the next flux phase orders this read before thread 0's next write
            g = gnext;
            gnext = s_grp[2];
        } else g = gn;
    }
#undef JX_D
}

}  // namespace jx
