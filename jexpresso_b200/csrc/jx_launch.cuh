// jx_launch.cuh -- turns one template combination into a KernelSet of plain function pointers.
#pragma once
#include "jx_internal.h"

namespace jx {

template <int NSD, int NGL>
constexpr int default_epb() {
    constexpr int np = Geo<NSD, NGL>::NP;
    return np >= 256 ? 1 : (256 / np);
}

template <int NSD, int NGL, class EQ, int VISC>
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

template <class EQ>
void launch_stage_t(const StageArgs &a, int grid, cudaStream_t s) {
    if (a.npoin <= 0) return;
    (void)grid;   // one node per thread: 1.07 ms at 25 M nodes against 1.5-1.7 ms for a persistent grid (scripts/micro/stage_bench.cu)
    k_stage_direct<EQ><<<(unsigned)((a.npoin + 255) / 256), 256, 0, s>>>(a);
}

template <int NEQ>
void launch_gather_t(const GatherArgs &a, cudaStream_t s) {
    if (a.npoin <= 0) return;
    k_gather<NEQ><<<(unsigned)((a.npoin + 255) / 256), 256, 0, s>>>(a);
}

template <int NSD, int NGL, class EQ, int VISC>
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
    ks.launch_stage = &launch_stage_t<EQ>;
    ks.group_bytes = C::GROUP_BYTES; ks.group_nt = 32; ks.zid_off = C::ZID_OFF; ks.fid_off = C::FID_OFF; ks.z_off = C::Z_OFF;
    ks.w_off = C::W_OFF; ks.wf_off = C::WF_OFF;
    ks.has_dyn = 1;
    return ks;
}

// variant 12: k_elem_tri (nop = 7): xi-, eta- and zeta-pencil roles of 64 lanes each, one element per CTA
template <int NGL, class EQ>
struct TriKernel {
    using C = ElemTriCfg<NGL, EQ>;
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(k_elem_tri<NGL, EQ, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_tri<NGL, EQ, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_elem_tri<NGL, EQ, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        return e;
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_elem_tri<NGL, EQ, 2>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, int grid, cudaStream_t s) {
        if (!a.atomics) k_elem_tri<NGL, EQ, 0><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        else if (a.Minv == nullptr) k_elem_tri<NGL, EQ, 1><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
        else k_elem_tri<NGL, EQ, 2><<<grid, C::NT, C::SMEM_BYTES, s>>>(a);
    }
};

template <int NGL, class EQ>
KernelSet make_tri_set(int eq_id, int lpert, int jxpow, int variant) {
    using K = TriKernel<NGL, EQ>;
    using C = typename K::C;
    KernelSet ks;
    ks.nsd = 3; ks.ngl = NGL; ks.eq_id = eq_id; ks.lpert = lpert; ks.jxpow = jxpow; ks.lvisc = 0; ks.variant = variant;
    ks.neq = EQ::NEQ;
    ks.elems_per_block = 1;
    ks.rec_layout = 7;
    ks.nthreads = C::NT;
    ks.smem_bytes = C::SMEM_BYTES;
    ks.prepare = &K::prepare;
    ks.max_blocks_per_sm = &K::max_blocks;
    ks.launch_elem = &K::launch;
    ks.launch_bc = &launch_bc_t<EQ>;
    ks.launch_gather = &launch_gather_t<EQ::NEQ>;
    ks.launch_aux = &launch_aux_t<EQ>;
    ks.launch_stage = &launch_stage_t<EQ>;
    ks.group_bytes = C::REC_BYTES; ks.zid_off = C::ZID_OFF; ks.fid_off = C::FID_OFF;
    return ks;
}

// variant 13: the inviscid team kernel followed by k_visc_quad (four warps per pair, node-parallel node-local step)
template <int NGL, class EQ>
struct ViscQuadKernel {
    using C = ViscQuadCfg<NGL, EQ>;
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(k_visc_quad<NGL, EQ, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_visc_quad<NGL, EQ, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        return e;
    }
    static int max_blocks() {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_visc_quad<NGL, EQ, 2>, C::NT, C::SMEM_BYTES);
        return nb;
    }
    static void launch(const ElemArgs &a, const ViscArgs &v, int grid, cudaStream_t s) {
        if (!a.atomics) k_visc_quad<NGL, EQ, 0><<<grid, C::NT, C::SMEM_BYTES, s>>>(a, v);
        else k_visc_quad<NGL, EQ, 2><<<grid, C::NT, C::SMEM_BYTES, s>>>(a, v);
    }
};

template <int NGL, class EQ, int ZW, int PW>
KernelSet make_team_visc_quad_set(int eq_id, int lpert, int jxpow, int variant) {
    KernelSet ks = make_team_set<NGL, EQ, ZW, PW>(eq_id, lpert, jxpow, variant);
    using V = ViscQuadKernel<NGL, EQ>;
    ks.lvisc = 1;
    ks.has_dyn = 1;                       // the viscous pass walks the same pair lists (ElemArgs::glist) as the inviscid launch
    ks.launch_visc = &V::launch;
    ks.visc_max_blocks = &V::max_blocks;
    ks.visc_prepare = &V::prepare;
    ks.visc_group_bytes = V::C::GROUP_BYTES; ks.visc_zid_off = V::C::ID_OFF; ks.visc_fid_off = V::C::ID_OFF;
    ks.visc_layout = 2;
    ks.retile_visc = [](const ViscRetileArgs &a, unsigned grid, cudaStream_t s) { k_retile_visc_quad<<<grid, 256, 0, s>>>(a); };
    return ks;
}

}  // namespace jx
