// jx_internal.h -- host-side glue between the C ABI (jexrhs.cu) and the kernel instantiation
// units (jx_inst_*.cu).  Not part of the public interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "jx_kernels.cuh"

namespace jx {

// One (nsd, ngl, equation set, PERT/TOTAL, pow mode, viscous on/off) combination = one set of
// compiled kernels.  The launchers are plain function pointers so jexrhs.cu never sees the templates.
struct KernelSet {
    int nsd, ngl, eq_id, lpert, jxpow, lvisc, variant;
    int neq;
    int elems_per_block;
    int rec_layout;                                             // element-record layout the kernel reads (RetileArgs::layout)
    int nthreads;
    size_t smem_bytes;
    cudaError_t (*prepare)();                                   // cudaFuncSetAttribute(s)
    int (*max_blocks_per_sm)();                                 // occupancy of the element kernel
    void (*launch_elem)(const ElemArgs &, int grid, cudaStream_t);
    void (*launch_bc)(const BcArgs &, cudaStream_t);
    void (*launch_gather)(const GatherArgs &, cudaStream_t);
    void (*launch_aux)(const AuxArgs &, int grid, cudaStream_t);   // nullptr unless the element kernel reads ElemArgs::aux
    // atomics mode, 2N schemes: u += B dt S', S' <- A_next S', next aux in one sweep (k_stage_direct) behind element kernels
    // that accumulate into S' directly; set for the kernels with launch_aux
    void (*launch_stage)(const StageArgs &, int grid, cudaStream_t) = nullptr;
    // AV viscous term as a pass of its own (k_visc_quad) behind an inviscid element kernel: its launcher and the layout of
    // its pair records (nullptr: the element kernel itself carries the viscous pass, or lvisc = 0)
    void (*launch_visc)(const ElemArgs &, const ViscArgs &, int grid, cudaStream_t) = nullptr;
    int (*visc_max_blocks)() = nullptr;
    cudaError_t (*visc_prepare)() = nullptr;
    void (*retile_visc)(const ViscRetileArgs &, unsigned grid, cudaStream_t) = nullptr;   // builds the viscous pair records
    int visc_group_bytes = 0, visc_zid_off = 0, visc_fid_off = 0;
    int visc_layout = 0;                                        // 0: no viscous records; 2: k_visc_quad (node ordered)
    // rec_layout 5 (element-group records of the team kernels): bytes per group and stream / id offsets
    int group_bytes = 0, group_nt = 0, zid_off = 0, fid_off = 0, z_off = 0, w_off = 0, wf_off = 0;
    int has_dyn = 0;                                            // launch_elem honours ElemArgs::glist/gctr (interface-first split)
};

// each instantiation unit exports one lookup; returns nullptr if it does not hold the combination
const KernelSet *lookup_euler_theta_3d(int ngl, int lpert, int jxpow, int lvisc, int variant);
const KernelSet *lookup_euler_theta_2d(int ngl, int lpert, int jxpow, int lvisc, int variant);
const KernelSet *lookup_other(int nsd, int ngl, int eq_id, int lvisc, int variant);

}  // namespace jx
