/*
 * jexrhs.h -- C ABI of libjexrhs, the B200-native (sm_100a, FP64) explicit-RHS engine that
 * drops in behind Jexpresso's `rhs!(du, u, params, t)` callback.
 *
 * The reference has no FFI: its boundary is Julia multiple dispatch
 *   rhs!(du,u,params,time)                       src/kernel/operators/rhs.jl:121-134
 *   _build_rhs!(RHS,u,params,time)               src/kernel/operators/rhs.jl:498-711
 * with `params` built by params_setup (src/kernel/infrastructure/params_setup.jl:442-479).
 * Each entry point below names the reference interface it replaces.  A Julia host binds
 * them with `ccall` (julia/rhs_b200.jl); the tests bind them with ctypes.
 *
 * Conventions
 *  - every function returns 0 on success or a negative JX_E* code; nothing throws across the
 *    ABI; jx_last_error() returns the message of the last failure on that context;
 *  - arrays are passed exactly as Julia holds them: column-major, Float64 / Int64, node and
 *    element ids 1-based; host pointers are borrowed for the duration of the call only;
 *  - the context owns all device memory (state, metrics, connectivity, M^-1, work arrays stay
 *    resident in HBM across stages); one context per rank/GPU; not thread-safe;
 *  - there is NO CPU fallback: without a CUDA device jx_init fails with JX_ENODEV.
 */
#ifndef JEXRHS_H
#define JEXRHS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jx_ctx jx_ctx;

/* error codes */
#define JX_OK 0
#define JX_EINVAL (-1)   /* bad argument / unsupported configuration */
#define JX_ENODEV (-2)   /* no CUDA device */
#define JX_ECUDA (-3)    /* CUDA runtime error (message in jx_last_error) */
#define JX_ESTATE (-4)   /* call order violated (e.g. jx_rhs before jx_upload_mesh) */
#define JX_ENCCL (-5)    /* NCCL error */
#define JX_ENOMEM (-6)

/* equation sets = registered device functors mirroring the user_flux!/user_source!/
 * user_primitives!/user_bc_dirichlet! hooks of a case directory (problems/<eqs>/<case>/) */
#define JX_EQ_EULER_THETA 0    /* problems/CompEuler/3d, problems/CompEuler/theta */
#define JX_EQ_EULER_ENERGY 1   /* problems/CompEuler/kelvinHelmholtzChan2022 (2D) */
#define JX_EQ_ADVDIFF 2        /* problems/AdvDiff/kopriva (2D), problems/AdvDiff/3d_periodic */
#define JX_EQ_SHALLOW_WATER 3  /* problems/ShallowWater/SoliWaveIsland (2D) */
#define JX_EQ_EULER_THETA_LES 4 /* problems/CompEuler/LESICP1 (3D, TOTAL): theta-form fluxes, sponge + Coriolis + geostrophic source;
                                  phys[8..12] = lsponge, zsponge, zmax, f, alpha */

/* stage drivers (OrdinaryDiffEq algorithms used at src/kernel/solvers/TimeIntegrators.jl:597-607) */
#define JX_SCHEME_CK2N54 0
#define JX_SCHEME_SSPRK54 1
#define JX_SCHEME_SSPRK33 2

/* boundary face kinds (src/kernel/boundaryconditions/BCs.jl:621-623) */
#define JX_BC_SKIP 0           /* periodic* tags: not touched by the Dirichlet loop */
#define JX_BC_FREE_SLIP 1      /* user_bc_dirichlet! free-slip projection */

/* jx_set_option keys */
#define JX_OPT_DSS_MODE 1      /* 0 = deterministic gather in the reference's element-ascending
                                  order (default); 1 = atomics (red.global.add.f64), M^-1 folded */
#define JX_OPT_POW_MODE 2      /* 0 = CUDA pow() (default); 1 = jx_pow (include/jxpow.h), bit-identical
                                  to the oracle's jx_pow */
#define JX_OPT_ELEM_KERNEL 3   /* element-kernel variant: JX_ELEM_AUTO (default) picks the fastest exact-order kernel that
                                  exists for the configuration (3D inviscid nop 2/4: the warp-team kernel, else the generic
                                  one); JX_ELEM_GENERIC forces the generic thread-per-node kernel; 8 / 9 name the warp-team kernel
                                  (one / two plane warps), 13 = 9 followed by the four-warp viscous pass k_visc_quad (the default with AV),
                                  12 the three-role pencil kernel of nop 7 (opt-in).
                                  All variants keep the reference's order of every sum: bit-identical results. */
#define JX_ELEM_AUTO 0
#define JX_ELEM_GENERIC (-1)
#define JX_OPT_CUDA_GRAPH 4    /* 1: jx_bench_rhs captures one RHS evaluation, and jx_step(CK2N54) one whole step (five stages:
                                  kernels + NCCL groups), in a CUDA graph and replays it; results are identical */
#define JX_OPT_OVERLAP 5       /* n > 0 (atomics DSS + warp-team element kernel + interface lists): the element groups that touch
                                  a shared node run first and the interface exchange (DSS_global_RHS!, rhs.jl:690) then runs on
                                  a second stream beside the launch over the interior groups, which leaves n SMs free for it.
                                  0 (default): one element launch, then the exchange.  May be set at any time */

/* replaces: MPI.Init / get_mpi_comm (src/run.jl:74-88).  nccl_uid: 128-byte ncclUniqueId shared
 * by all ranks (see jx_nccl_unique_id) or NULL when nranks == 1. */
int jx_init(int device, int rank, int nranks, const void *nccl_uid, jx_ctx **out);
/* jx_init with the communicator capped to nccl_max_ctas CTAs (ncclCommInitRankConfig, maxCTAs): the NCCL send/recv kernels of
 * the interface exchange then fit on the SMs JX_OPT_OVERLAP = nccl_max_ctas leaves free of the interior element launch, whatever
 * the host process's environment says.  0 = no cap (= jx_init). */
int jx_init_ex(int device, int rank, int nranks, const void *nccl_uid, int nccl_max_ctas, jx_ctx **out);
int jx_nccl_unique_id(void *uid128);
void jx_destroy(jx_ctx *);
int jx_last_error(jx_ctx *, char *buf, int len);
int jx_set_option(jx_ctx *, int key, int64_t value);
int jx_version(void);

/* replaces: the scalar part of `params` (params_setup.jl:442-479): SD, neqs, ngl, inputs[:lsource],
 * inputs[:lvisc], inputs[:SOL_VARS_TYPE] (lpert), visc_coeff = inputs[:μ] (params_setup.jl:307-315),
 * PhysicalConst (globalConstantsPhysics.jl:3-62) packed as
 *   phys[0..7] = C0, γ, g, Rair, cp, cv, pref, γ-1 ; phys[8..10] = AdvDiff wind (u,v,w). */
int jx_set_problem(jx_ctx *, int nsd, int ngl, int neqs, int64_t nelem, int64_t npoin, int equation_id, int lpert,
                   int lsource, int lvisc, const double *visc_coeff, const double *phys_consts, int nphys);

/* replaces: inputs[:visc_model] = AV() | SMAG() | VREM() -> params.sgs = allocate_SGS(npoin, T, backend, PhysConst, visc_model)
 * (src/kernel/physics/sgsStructs.jl:77-120) with the flags set at params_setup.jl:249-253, read by compute_sgs_cache! and the
 * cache-reading SGS_diffusion / _expansion_visc! (SGS.jl:1087-1109, 1118-1685; rhs.jl:2275-2400, 2582-2785).  Call after
 * jx_set_problem(lvisc = 1) and before jx_upload_mesh.  consts = PhysicalConst's [Pr_t, Sc_t, μ_mol, κ_mol, Ri_crit, C_s]
 * (globalConstantsPhysics.jl:15-31; g is phys[2]); delta_effective = params.mesh.Δeffective_l (mesh.jl:5632); lrichardson =
 * get(inputs, :lrichardson, true); ltheta_eqn = !(inputs[:energy_equation] == "energy"); ad_lvl = params.mesh.ad_lvl
 * (Int64[nelem], 3D only: Δ_effective = ldexp(Δ, -ad_lvl[iel]), rhs.jl:1416) or NULL.  Dry runs only (size(mp.Tabs,1) == 1).
 * The closures run on the generic element kernel, in both DSS modes, bit-identical to the oracle in the deterministic one. */
#define JX_VISC_AV 0
#define JX_VISC_SMAG 1
#define JX_VISC_VREM 2
int jx_set_sgs(jx_ctx *, int visc_model, double delta_effective, int lrichardson, int ltheta_eqn, const double *consts,
               int nconsts, const int64_t *ad_lvl);

/* replaces: params.mesh.connijk, params.mesh.coords, params.metrics.{dξdx..dζdz,Je}, params.basis.dψ,
 * params.ω, params.Minv, params.qp.qe.  metrics: 3D 10 arrays (dξdx dξdy dξdz dηdx dηdy dηdz dζdx dζdy dζdz Je),
 * 2D 5 arrays (dξdx dξdy dηdx dηdy Je), each Float64[nelem, ngl, ngl, ngl|1] element-fastest. */
int jx_upload_mesh(jx_ctx *, const int64_t *connijk, const double *coords, const double *const *metrics, int nmetrics,
                   const double *dpsi, const double *omega, const double *Minv, const double *qe);

/* replaces: build_metric_terms! (src/kernel/mesh/metric_terms.jl:332-474 3D, :197-257 2D) + jx_upload_mesh: the same upload
 * with the metric terms dξdx..dζdz, Je built on the device from connijk and coords (required here) -- each coordinate
 * differentiated along the LGL lines in ascending node order, cofactors and determinant with the reference's association --
 * so the 10 (5) element-sized host arrays need not exist.  Results are bit-identical to the host arrays of the oracle. */
int jx_upload_mesh_coords(jx_ctx *, const int64_t *connijk, const double *coords, const double *dpsi, const double *omega,
                          const double *Minv, const double *qe);
/* Minv == NULL in jx_upload_mesh_coords: the diagonal mass matrix is built on the device as well -- replaces matrix_wrapper's
 * build_mass_matrix! + DSS_mass! + DSS_global_mass! + Minv = 1 ./ M (src/kernel/infrastructure/element_matrices.jl:173-214,
 * 593-617, 1160-1174, 1557-1559): M[ip] = sum over elements ascending of (w_i*w_j)*w_k*Je, summed across the ranks through the
 * lists of jx_upload_halo (upload them before the first evaluation), inverted.  Bit-identical to the host arrays of the oracle.
 * jx_get_minv returns the assembled M^-1 (Float64[npoin]). */
int jx_get_minv(jx_ctx *, double *Minv);
/* replaces: conformity4ncf_q! (src/kernel/Adaptivity/Projection.jl:2919-2970; params_setup.jl:259-297), the conditioning of the
 * initial state that gives shared / periodic-twin nodes one value:  q <- Minv * DSS_global(wJac * q) on the neqs columns of the
 * resident state (which = 0, after jx_set_state) or of the reference state qe (which = 1).  Needs the device-built mass. */
int jx_condition_state(jx_ctx *, int which);

/* replaces: params.mesh.poin_in_bdy_face / poin_in_bdy_edge, params.metrics.nx/ny/nz, bdy_face_type
 * (tags mapped to JX_BC_* kinds by the host).  3D arrays are [nfaces, ngl, ngl], 2D [nedges, ngl]. */
int jx_upload_bcs(jx_ctx *, int64_t nfaces, const int64_t *poin_in_bdy_face, const double *nx, const double *ny,
                  const double *nz, const int32_t *face_bc_kind);

/* replaces: inputs[:bdy_fluxes] = true -> apply_boundary_conditions_neumann! -> build_custom_bcs_neumann!(::NSD_3D)
 * (src/kernel/boundaryconditions/BCs.jl:35-63, 655-816) with the Monin-Obukhov wall model CM_MOST! (src/kernel/physics/CM_MOST.jl:
 * 69-100, 144-152, 224-261) on the faces tagged "MOST", compute_surface_integral! / DSS_surface_integral!
 * (src/kernel/boundaryconditions/surface_integral.jl:1-29) and RHS .+= S_flux (rhs.jl:674-689).  3D theta-form CompEuler, dry
 * (size(mp.Tabs,1) == 1), inputs[:bulk_fluxes] = false.  Arrays as Julia holds them: params.mesh.poin_in_bdy_face [nf,n,n],
 * params.mesh.bdy_face_in_elem [nf], params.mesh.connijk [E,n,n,n], params.metrics.nx/ny/nz and params.metrics.Jef [nf,n,n],
 * params.ω [n]; face_flux_kind[f] = JX_FLUX_MOST where bdy_face_type[f] == "MOST", JX_FLUX_NONE elsewhere (user_bc_neumann!
 * leaves F_surf zero in the shipped decks); inputs[:ifirst_wall_node_index], inputs[:δhf], inputs[:user_heatflux];
 * most_consts = [PhysConst.karman, z0_m, z0_h] (NULL: 0.4, 0.1, 0.01 -- the literals of BCs.jl:770-772).  After jx_upload_mesh.
 * Transcendentals are CUDA's: 1e-12 against the oracle, not bit equality.  The interface-first overlap is not used with it. */
#define JX_FLUX_NONE 0
#define JX_FLUX_MOST 1
int jx_upload_bdy_fluxes(jx_ctx *, int64_t nfaces, const int64_t *poin_in_bdy_face, const int64_t *bdy_face_in_elem,
                         const int64_t *connijk, const double *nx, const double *ny, const double *nz, const double *Jef,
                         const double *omega, const int32_t *face_flux_kind, int ifirst_wall_node_index, double delta_hf,
                         double user_heatflux, const double *most_consts, int nconsts);

/* replaces: AssemblerCache built by setup_assembler (src/kernel/mpi/mpi_communications.jl:48-234).
 * CSR by peer rank: send_i (local ids sent to owner r), recv_idx (owner-side local ids the values from r add
 * into), recvback_idx (local ids overwritten with the owner's sum).  Self entries (r == rank) carry the
 * periodic-twin pairs the reference exchanges by self-send. */
int jx_upload_halo(jx_ctx *, const int64_t *send_ptr, const int64_t *send_i, const int64_t *recv_ptr,
                   const int64_t *recv_idx, const int64_t *recvback_ptr, const int64_t *recvback_idx);

/* state: u is the ODE state vector Float64[npoin*neqs], index (ieq-1)*npoin + ip (rhs.jl:29-47) */
int jx_set_state(jx_ctx *, const double *u);
int jx_get_state(jx_ctx *, double *u);
/* du of the last jx_rhs / jx_bench_rhs(fused_stage = 0) evaluation.  jx_step does not produce it in the atomics mode: there the
 * element kernels accumulate straight into the 2N low-storage register. */
int jx_get_du(jx_ctx *, double *du);

/* replaces: rhs!(du,u,params,time).  u_host != NULL: upload u first; du_host != NULL: download du after.
 * u_back_host != NULL: download the state after the Dirichlet projection (rhs! mutates u, BCs.jl:651).
 * All NULL => fully device resident (state set by jx_set_state / advanced by jx_step). */
int jx_rhs(jx_ctx *, double t, const double *u_host, double *du_host, double *u_back_host);

/* rhs!(du,u,params,time) on DEVICE arrays the caller owns (a CuArray state under OrdinaryDiffEq: no PCIe copy): u_dev and
 * du_dev are Float64[npoin*neqs] in the same flat layout, on the context's device; u_dev is projected in place (rhs! mutates
 * u), du_dev receives the mass-scaled right-hand side.  Returns after the context's stream has drained. */
int jx_rhs_dev(jx_ctx *, double t, double *u_dev, double *du_dev);

/* element-kernel variant in use after JX_ELEM_AUTO resolution (0 = generic k_elem_node); negative JX_E* before jx_set_problem */
int jx_kernel_variant(jx_ctx *);

/* replaces: OrdinaryDiffEq perform_step! for the fixed-step explicit schemes the decks use; all stages on the
 * device, M^-1 fused with the stage update, no host round trip.  dt is used as given (the Julia side passes
 * Float64(Float32(Δt)), TimeIntegrators.jl:464-465). */
int jx_step(jx_ctx *, int scheme, double t, double dt, int nsteps);

/* device-side timing of the last jx_rhs / jx_step call (CUDA events on the context's stream), in ms */
int jx_last_elapsed_ms(jx_ctx *, float *ms);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
int64_t jx_launch_count(jx_ctx *);
int jx_sync(jx_ctx *);
/* interface-first split in force for the next evaluation (JX_OPT_OVERLAP): numbers of element groups in the launch that
 * precedes the exchange and in the launch that runs beside it; both 0 when the evaluation uses one launch.  (No reference
 * counterpart: the reference overlaps nothing, mpi_communications.jl:260-338 blocks in MPI.Waitall.) */
int jx_split_info(jx_ctx *, int64_t *interface_groups, int64_t *interior_groups);

/* benchmarking helper: `n` device-resident RHS evaluations back to back on the context's stream, timed
 * with CUDA events.  fused_stage != 0 evaluates a full low-storage RK stage (RHS + M^-1 + stage update with
 * dt = 0, so the state is unchanged) instead of rhs! alone.  phase_ms (optional, 8 floats) receives the
 * summed device time per phase: [0] boundary projection, [1] element kernel, [2] DSS gather (or the
 * zero-fill in atomics mode), [3] interface exchange, [4] M^-1 / stage update passes. */
int jx_bench_rhs(jx_ctx *, int n, int fused_stage, float *total_ms, float *phase_ms);

/* device self tests (diagnostics for the parity suite, no reference counterpart).  which = 0: the shared-
 * reciprocal division used by the two-stage flux functors against the compiler's correctly rounded `/` on n
 * pseudo-random operand pairs; *failures receives the number of bitwise mismatches. */
int jx_selftest(jx_ctx *, int which, int64_t n, int64_t *failures);

#ifdef __cplusplus
}
#endif
#endif /* JEXRHS_H */
