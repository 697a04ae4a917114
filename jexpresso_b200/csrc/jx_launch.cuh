// jx_launch.cuh -- turns one template combination into a KernelSet of plain function pointers.
#pragma once
#include "jx_internal.h"

namespace jx {

template <int NSD, int NGL>
constexpr int default_epb() {
    constexpr int np = Geo<NSD, NGL>::NP;
    return np >= 256 ? 1 : (256 / np);
}

template <int NSD, int NGL, class EQ, bool VISC>
struct NodeKernel {
    static constexpr int EPB = default_epb<NSD, NGL>();
    using C = ElemNodeCfg<NSD, NGL, EQ, VISC, EPB>;
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(k_elem_node<NSD, NGL, EQ, VISC, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)C::SMEM_BYTES);
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_elem_node<NSD, NGL, EQ, VISC, EPB>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, int grid, cudaStream_t s) {
        k_elem_node<NSD, NGL, EQ, VISC, EPB><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
    }
};

template <class EQ>
void launch_bc_t(const BcArgs &a, cudaStream_t s) {
    if (a.nb <= 0) return;
    k_bc_dirichlet<EQ><<<(a.nb + 127) / 128, 128, 0, s>>>(a);
}

template <class EQ>
void launch_aux_t(const AuxArgs &a, int grid, cudaStream_t s) {
    if (a.npoin <= 0) return;
    k_node_aux<EQ><<<grid, 256, 0, s>>>(a);
}

template <int NEQ>
void launch_gather_t(const GatherArgs &a, cudaStream_t s) {
    if (a.npoin <= 0) return;
    k_gather<NEQ><<<(unsigned)((a.npoin + 255) / 256), 256, 0, s>>>(a);
}

template <int NSD, int NGL, class EQ, bool VISC>
KernelSet make_node_set(int eq_id, int lpert, int jxpow) {
    using K = NodeKernel<NSD, NGL, EQ, VISC>;
    KernelSet ks;
    ks.nsd = NSD; ks.ngl = NGL; ks.eq_id = eq_id; ks.lpert = lpert; ks.jxpow = jxpow; ks.lvisc = VISC; ks.variant = 0;
    ks.neq = EQ::NEQ;
    ks.elems_per_block = K::EPB;
    ks.rec_layout = 0;
    ks.nthreads = K::C::NT;
    ks.smem_bytes = K::C::SMEM_BYTES;
    ks.prepare = &K::prepare;
    ks.max_blocks_per_sm = &K::max_blocks;
    ks.launch_elem = &K::launch;
    ks.launch_bc = &launch_bc_t<EQ>;
    ks.launch_gather = &launch_gather_t<EQ::NEQ>;
    ks.launch_aux = nullptr;
    return ks;
}

// variant 1 (EXACT) / 2 (single partial): pencil kernel, 3D inviscid
template <int NGL, class EQ, bool EXACT>
struct PencilKernel {
    static constexpr int NC = NGL * NGL;
    static constexpr int EPB = NC >= 128 ? 1 : (128 / NC);
    using C = ElemPencilCfg<NGL, EQ, EPB, EXACT>;
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(k_elem_pencil<NGL, EQ, EPB, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)C::SMEM_BYTES);
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_elem_pencil<NGL, EQ, EPB, EXACT>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, int grid, cudaStream_t s) {
        k_elem_pencil<NGL, EQ, EPB, EXACT><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
    }
};

template <int NGL, class EQ, bool EXACT>
KernelSet make_pencil_set(int eq_id, int lpert, int jxpow) {
    using K = PencilKernel<NGL, EQ, EXACT>;
    KernelSet ks;
    ks.nsd = 3; ks.ngl = NGL; ks.eq_id = eq_id; ks.lpert = lpert; ks.jxpow = jxpow; ks.lvisc = 0; ks.variant = EXACT ? 1 : 2;
    ks.neq = EQ::NEQ;
    ks.elems_per_block = K::EPB;
    ks.rec_layout = 1;
    ks.nthreads = K::C::NT;
    ks.smem_bytes = K::C::SMEM_BYTES;
    ks.prepare = &K::prepare;
    ks.max_blocks_per_sm = &K::max_blocks;
    ks.launch_elem = &K::launch;
    ks.launch_bc = &launch_bc_t<EQ>;
    ks.launch_gather = &launch_gather_t<EQ::NEQ>;
    ks.launch_aux = nullptr;
    return ks;
}

// variant 3 (EXACT) / 4 (single partial): one element per CTA pencil kernel, 3D inviscid
template <int NGL, class EQ, bool EXACT>
struct WPencilKernel {
    using C = ElemWPencilCfg<NGL, EQ, EXACT>;
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(k_elem_wpencil<NGL, EQ, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)C::SMEM_BYTES);
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_elem_wpencil<NGL, EQ, EXACT>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, int grid, cudaStream_t s) {
        k_elem_wpencil<NGL, EQ, EXACT><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
    }
};

template <int NGL, class EQ, bool EXACT>
KernelSet make_wpencil_set(int eq_id, int lpert, int jxpow) {
    using K = WPencilKernel<NGL, EQ, EXACT>;
    KernelSet ks;
    ks.nsd = 3; ks.ngl = NGL; ks.eq_id = eq_id; ks.lpert = lpert; ks.jxpow = jxpow; ks.lvisc = 0; ks.variant = EXACT ? 3 : 4;
    ks.neq = EQ::NEQ;
    ks.elems_per_block = 1;
    ks.rec_layout = 1;
    ks.nthreads = K::C::NT;
    ks.smem_bytes = K::C::SMEM_BYTES;
    ks.prepare = &K::prepare;
    ks.max_blocks_per_sm = &K::max_blocks;
    ks.launch_elem = &K::launch;
    ks.launch_bc = &launch_bc_t<EQ>;
    ks.launch_gather = &launch_gather_t<EQ::NEQ>;
    ks.launch_aux = EQ::HAS_AUX ? &launch_aux_t<EQ> : nullptr;
    return ks;
}

// variant 5 (group of elements per CTA, lanes full) / 6 (one element per CTA): group-pencil kernel, 3D inviscid,
// exact order
template <int NGL, class EQ, int EPB>
struct GPencilKernel {
    using C = ElemGPencilCfg<NGL, EQ, EPB>;
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(k_elem_gpencil<NGL, EQ, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)C::SMEM_BYTES);
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_elem_gpencil<NGL, EQ, EPB>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, int grid, cudaStream_t s) {
        k_elem_gpencil<NGL, EQ, EPB><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
    }
};

template <int NGL, class EQ, int EPB>
KernelSet make_gpencil_set(int eq_id, int lpert, int jxpow, int variant) {
    using K = GPencilKernel<NGL, EQ, EPB>;
    using C = typename K::C;
    using L = typename C::L;
    KernelSet ks;
    ks.nsd = 3; ks.ngl = NGL; ks.eq_id = eq_id; ks.lpert = lpert; ks.jxpow = jxpow; ks.lvisc = 0; ks.variant = variant;
    ks.neq = EQ::NEQ;
    ks.elems_per_block = EPB;
    ks.rec_layout = 4;
    ks.nthreads = C::NT;
    ks.smem_bytes = C::SMEM_BYTES;
    ks.prepare = &K::prepare;
    ks.max_blocks_per_sm = &K::max_blocks;
    ks.launch_elem = &K::launch;
    ks.launch_bc = &launch_bc_t<EQ>;
    ks.launch_gather = &launch_gather_t<EQ::NEQ>;
    ks.launch_aux = &launch_aux_t<EQ>;
    ks.group_bytes = C::GROUP_BYTES; ks.group_nt = C::NT; ks.zid_off = C::ZID_OFF; ks.fid_off = C::FID_OFF;
    const DigitOrder ord[3] = {L::XI, L::ETA, L::ZETA};
    for (int ps = 0; ps < 3; ++ps)
        for (int d = 0; d < 3; ++d) ks.group_mult[ps][d] = gp_mult(ord[ps], d, NGL, EPB);
    return ks;
}

// variant 8 (one plane warp, one zeta warp per element slot) / 9 (two plane warps): plane-role + zeta-pencil-role warp team per element
// group, 3D inviscid, exact order; the scatter mode (rhs_el store / RED.ADD / RED.ADD with folded M^-1) is a
// template parameter chosen per launch
template <int NGL, class EQ, int ZW, int PW>
struct TeamKernel {
    using C = ElemTeamCfg<NGL, EQ, ZW, PW>;
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(k_elem_team<NGL, EQ, ZW, 0, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_team<NGL, EQ, ZW, 1, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_team<NGL, EQ, ZW, 2, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_team<NGL, EQ, ZW, 2, PW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        return e;
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_elem_team<NGL, EQ, ZW, 2, PW>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, int grid, cudaStream_t s) {
        if (a.gctr != nullptr) {   // interface-first split: list-driven launch (atomics DSS with folded M^-1 only)
            if (a.atomics && a.Minv != nullptr) k_elem_team<NGL, EQ, ZW, 2, PW, true><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        } else if (!a.atomics) k_elem_team<NGL, EQ, ZW, 0, PW><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        else if (a.Minv == nullptr) k_elem_team<NGL, EQ, ZW, 1, PW><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        else k_elem_team<NGL, EQ, ZW, 2, PW><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
    }
};

template <int NGL, class EQ, int ZW, int PW>
KernelSet make_team_set(int eq_id, int lpert, int jxpow, int variant) {
    using K = TeamKernel<NGL, EQ, ZW, PW>;
    using C = typename K::C;
    KernelSet ks;
    ks.nsd = 3; ks.ngl = NGL; ks.eq_id = eq_id; ks.lpert = lpert; ks.jxpow = jxpow; ks.lvisc = 0; ks.variant = variant;
    ks.neq = EQ::NEQ;
    ks.elems_per_block = C::EPB;
    ks.rec_layout = 5;
    ks.nthreads = C::NT;
    ks.smem_bytes = C::SMEM_BYTES;
    ks.prepare = &K::prepare;
    ks.max_blocks_per_sm = &K::max_blocks;
    ks.launch_elem = &K::launch;
    ks.launch_bc = &launch_bc_t<EQ>;
    ks.launch_gather = &launch_gather_t<EQ::NEQ>;
    ks.launch_aux = &launch_aux_t<EQ>;
    ks.group_bytes = C::GROUP_BYTES; ks.group_nt = 32; ks.zid_off = C::ZID_OFF; ks.fid_off = C::FID_OFF; ks.z_off = C::Z_OFF;
    ks.has_dyn = 1;
    return ks;
}

template <class EQ>
void launch_image_t(const ImageArgs &a, int grid, cudaStream_t s) {
    if (a.npoin <= 0) return;
    k_node_image<EQ><<<grid, 256, 0, s>>>(a);
}

// variant 10: k_elem_team2 (node image + TMA row gathers, ring of equation slots, producer/consumer barriers)
template <int NGL, class EQ, bool WT>
struct Team2Kernel {
    using C = ElemTeam2Cfg<NGL, EQ, WT>;
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(k_elem_team2<NGL, EQ, 0, false, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_team2<NGL, EQ, 1, false, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_team2<NGL, EQ, 2, false, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_team2<NGL, EQ, 2, true, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        return e;
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_elem_team2<NGL, EQ, 2, false, WT>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, int grid, cudaStream_t s) {
        if (a.gctr != nullptr) {
            if (a.atomics && a.Minv != nullptr) k_elem_team2<NGL, EQ, 2, true, WT><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        } else if (!a.atomics) k_elem_team2<NGL, EQ, 0, false, WT><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        else if (a.Minv == nullptr) k_elem_team2<NGL, EQ, 1, false, WT><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        else k_elem_team2<NGL, EQ, 2, false, WT><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
    }
};

template <int NGL, class EQ, bool WT>
KernelSet make_team2_set(int eq_id, int lpert, int jxpow, int variant) {
    using K = Team2Kernel<NGL, EQ, WT>;
    using C = typename K::C;
    using L = typename C::L;
    KernelSet ks;
    ks.nsd = 3; ks.ngl = NGL; ks.eq_id = eq_id; ks.lpert = lpert; ks.jxpow = jxpow; ks.lvisc = 0; ks.variant = variant;
    ks.neq = EQ::NEQ;
    ks.elems_per_block = C::EPB;
    ks.rec_layout = 5;
    ks.nthreads = C::NT;
    ks.smem_bytes = C::SMEM_BYTES;
    ks.prepare = &K::prepare;
    ks.max_blocks_per_sm = &K::max_blocks;
    ks.launch_elem = &K::launch;
    ks.launch_bc = &launch_bc_t<EQ>;
    ks.launch_gather = &launch_gather_t<EQ::NEQ>;
    ks.launch_aux = nullptr;
    ks.launch_image = &launch_image_t<EQ>;
    ks.img_rowd = C::ROWD;
    ks.group_bytes = L::GROUP_BYTES; ks.group_nt = 32; ks.zid_off = L::ZID_OFF; ks.fid_off = L::FID_OFF; ks.z_off = L::Z_OFF;
    ks.wpos_off = L::WPOS_OFF; ks.runi_off = L::RUNI_OFF; ks.runr_off = L::RUNR_OFF; ks.runl_off = L::RUNL_OFF; ks.nrun_off = L::NRUN_OFF;
    ks.maxrun = L::MAXRUN;
    ks.has_dyn = 1;
    return ks;
}

}  // namespace jx
