# rhs_b200.jl -- reference-side binding of libjexrhs (include/jexrhs.h) for Jexpresso.
#
# `include` this file inside `module Jexpresso` after src/kernel/operators/rhs.jl.  It adds ONE backend branch to the
# reference: when `inputs[:backend]` is `B200()` the ODEProblem is built on `rhs_b200!` with `B200Params(params, ctx)` as its
# parameter object instead of `rhs!` with `params` (src/kernel/solvers/TimeIntegrators.jl:210-213); everything else (mesh,
# setup, integrator, callbacks, I/O) is untouched.  It cannot be executed in the authoring image (no Julia); its ctypes twin
# jexpresso_b200/{capi,rhs}.py is what the tests run, and tests/test_julia_shim_cpu.py checks that every `params.*` field
# and `inputs[:*]` key read below exists in the reference (params_setup.jl:442-479, mod_inputs.jl, run.jl:160-170).
module JexRHSB200

const LIB = get(ENV, "JEXRHS_LIB", "libjexrhs.so")
const Ctx = Ptr{Cvoid}
const J = parentmodule(@__MODULE__)                # Jexpresso: PHYS_CONST (rhs.jl:11), PERT (SOL_VARS_TYPE tags)

struct B200 end                                    # value for inputs[:backend]

"What the ODEProblem carries instead of `params`: the reference's NamedTuple plus the device context and a host buffer."
struct B200Params{P}
    params::P
    ctx::Ctx
    dubuf::Vector{Float64}                         # staging for `du` objects without a pointer (ArrayFuse, see rhs_b200!)
end

const SCHEMES = Dict("CarpenterKennedy2N54" => 0, "SSPRK54" => 1, "SSPRK33" => 2)
const PERIODIC = ("periodicx", "periodicy", "periodicz", "periodic1", "periodic2", "periodic3", "Laguerre")

function check(ctx::Ctx, rc::Cint)
    rc == 0 && return
    buf = Vector{UInt8}(undef, 512)
    ccall((:jx_last_error, LIB), Cint, (Ctx, Ptr{UInt8}, Cint), ctx, buf, 512)
    error("libjexrhs error $rc: " * unsafe_string(pointer(buf)))
end

"""
Equation-set id of include/jexrhs.h.  The reference names the equation directory in `inputs[:_parsed_equations]`
(run.jl:167); the total-energy form is CompEuler with `inputs[:energy_equation] == "energy"` (mod_inputs.jl:1027-1028,
run.jl:305-314); the LESICP1 case has a source hook of its own (sponge, Coriolis, geostrophic wind).
"""
function equation_id(inputs)
    eqs = inputs[:_parsed_equations]
    if eqs == "CompEuler"
        inputs[:_parsed_case_name] == "LESICP1" && return 4          # theta fluxes + sponge / Coriolis / geostrophic source
        return inputs[:energy_equation] == "theta" ? 0 : 1
    end
    eqs == "AdvDiff"      && return 2
    eqs == "ShallowWater" && return 3
    error("libjexrhs has no device functor registered for problems/$eqs")
end

"""
The 16 packed constants of jx_set_problem: [C0, γ, g, Rair, cp, cv, pref, γ-1] from PhysicalConst{Float64}() (rhs.jl:11,
globalConstantsPhysics.jl:3-62), then the constants the case hooks hard-code (they are literals inside user_flux.jl /
user_source.jl, not inputs): slots 8-10 the AdvDiff wind; slots 9-15 the ShallowWater cone and wet/dry constants
(jexpresso_b200/physics.py: advdiff_packed, swe_packed).  `inputs[:b200_case_constants]` (a Vector{Float64} for slots
8-15) overrides the table below for cases it does not list.
"""
function packed_constants(inputs, mesh)
    PC = J.PHYS_CONST
    phys = zeros(Float64, 16)
    phys[1:8] .= (PC.C0, PC.γ, PC.g, PC.Rair, PC.cp, PC.cv, PC.pref, PC.γm1)
    if inputs[:_parsed_equations] == "CompEuler" && inputs[:_parsed_case_name] == "LESICP1"
        # problems/CompEuler/LESICP1/user_source.jl:30-103: sponge switch and base height from the deck, zmax from the mesh,
        # f = 1.0e-4 and alpha = 0.5 are literals of the hook
        phys[9:13] .= (inputs[:lsponge] ? 1.0 : 0.0, inputs[:zsponge], mesh.zmax, 1.0e-4, 0.5)
        return phys
    end
    if haskey(inputs, :b200_case_constants)
        phys[9:16] .= inputs[:b200_case_constants]
    else
        eqs, case = inputs[:_parsed_equations], inputs[:_parsed_case_name]
        if eqs == "AdvDiff"
            wind = case == "kopriva" ? (0.5, 1.0, 0.0) :                       # problems/AdvDiff/kopriva/user_flux.jl:1-16
                   case == "3d_periodic" ? (0.2, 0.2, 0.0) :                    # problems/AdvDiff/3d_periodic/user_flux.jl:1-16
                   error("AdvDiff case $case: pass the wind in inputs[:b200_case_constants]")
            phys[9:11] .= wind
        elseif eqs == "ShallowWater"
            case == "SoliWaveIsland" || error("ShallowWater case $case: pass the constants in inputs[:b200_case_constants]")
            # user_flux.jl:44-86, user_source.jl:34-73: cone height, dry relaxation, g, film depth, cone centre and radius
            phys[10:16] .= (0.93, 25.0, 9.81, 1.0e-3, 12.5, 0.0, 3.6)
        end
    end
    return phys
end

"Build the device-resident problem from what params_setup returns (params_setup.jl:442-479)."
function b200_setup(params, inputs; device = 0, rank = 0, nranks = 1, uid = C_NULL)
    overlap = get(inputs, :b200_overlap, 0)
    ctx = Ref{Ctx}(C_NULL)
    # jx_init_ex caps the NCCL communicator to `overlap` CTAs (ncclCommInitRankConfig), so the exchange kernels fit on the SMs
    # the interface-first overlap leaves free -- no environment variables needed
    rc = ccall((:jx_init_ex, LIB), Cint, (Cint, Cint, Cint, Ptr{Cvoid}, Cint, Ref{Ctx}), device, rank, nranks, uid,
               nranks > 1 ? overlap : 0, ctx)
    rc == 0 || error("jx_init_ex failed ($rc): no CUDA device / NCCL")
    c = ctx[]
    mesh, metrics, basis = params.mesh, params.metrics, params.basis
    nsd = mesh.nsd; ngl = mesh.ngl; neqs = params.neqs
    phys = packed_constants(inputs, mesh)
    # engine options (include/jexrhs.h): inputs[:b200_dss] = :gather (deterministic, reference summation order, default)
    # or :atomics (throughput mode); the element kernel is chosen by the library (JX_ELEM_AUTO) unless
    # inputs[:b200_elem_kernel] names a variant; the shared jx_pow keeps the equation of state reproducible
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 1, get(inputs, :b200_dss, :gather) == :atomics ? 1 : 0))
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 2, get(inputs, :b200_pow, 1)))
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 3, get(inputs, :b200_elem_kernel, 0)))
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 5, overlap))
    μ = Float64.(params.visc_coeff)                                          # inputs[:μ], params_setup.jl:307-315
    lpert = inputs[:SOL_VARS_TYPE] == J.PERT() ? 1 : 0
    check(c, ccall((:jx_set_problem, LIB), Cint,
                   (Ctx, Cint, Cint, Cint, Int64, Int64, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Cint),
                   c, nsd, ngl, neqs, mesh.nelem, mesh.npoin, equation_id(inputs), lpert,
                   inputs[:lsource] ? 1 : 0, inputs[:lvisc] ? 1 : 0, μ, phys, length(phys)))
    # SGS closure (inputs[:visc_model] = SMAG() | VREM(); params.sgs = allocate_SGS(...), params_setup.jl:246-253): the scalar
    # content of the SGS struct; the per-node caches of the reference live in registers on the device.  Dry runs only.
    vm = inputs[:visc_model]
    if inputs[:lvisc] && (vm isa J.SMAG || vm isa J.VREM)
        size(params.mp.Tabs, 1) == 1 || error("libjexrhs: the SGS closures are implemented for dry runs (no microphysics)")
        PC = J.PHYS_CONST
        sgsc = Float64[PC.Pr_t, PC.Sc_t, PC.μ_mol, PC.κ_mol, PC.Ri_crit, PC.C_s]
        check(c, ccall((:jx_set_sgs, LIB), Cint, (Ctx, Cint, Float64, Cint, Cint, Ptr{Float64}, Cint, Ptr{Int64}),
                       c, vm isa J.SMAG ? 1 : 2, Float64(mesh.Δeffective_l), get(inputs, :lrichardson, true) ? 1 : 0,
                       inputs[:energy_equation] == "energy" ? 0 : 1, sgsc, length(sgsc), Int64.(mesh.ad_lvl)))
    end
    mets = nsd == 3 ?
        [metrics.dξdx, metrics.dξdy, metrics.dξdz, metrics.dηdx, metrics.dηdy, metrics.dηdz,
         metrics.dζdx, metrics.dζdy, metrics.dζdz, metrics.Je] :
        [metrics.dξdx, metrics.dξdy, metrics.dηdx, metrics.dηdy, metrics.Je]
    if get(inputs, :b200_device_metrics, false)
        # build_metric_terms! on the device from connijk + coords (metric_terms.jl:332-474): same bits, no host arrays read
        check(c, ccall((:jx_upload_mesh_coords, LIB), Cint,
                       (Ctx, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       c, mesh.connijk, mesh.coords, basis.dψ, params.ω, params.Minv, params.qp.qe))
    else
        GC.@preserve mets begin
            ptrs = Ptr{Float64}[pointer(m) for m in mets]
            check(c, ccall((:jx_upload_mesh, LIB), Cint,
                           (Ctx, Ptr{Int64}, Ptr{Float64}, Ptr{Ptr{Float64}}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                           c, mesh.connijk, mesh.coords, ptrs, length(ptrs), basis.dψ, params.ω, params.Minv, params.qp.qe))
        end
    end
    kinds = Int32[t in PERIODIC ? 0 : 1 for t in mesh.bdy_face_type]        # BCs.jl:621-623
    if nsd == 3
        check(c, ccall((:jx_upload_bcs, LIB), Cint, (Ctx, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                       c, size(mesh.poin_in_bdy_face, 1), mesh.poin_in_bdy_face, metrics.nx, metrics.ny, metrics.nz, kinds))
    else
        check(c, ccall((:jx_upload_bcs, LIB), Cint, (Ctx, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                       c, size(mesh.poin_in_bdy_edge, 1), mesh.poin_in_bdy_edge, metrics.nx, metrics.ny, C_NULL, kinds))
    end
    # Boundary fluxes (inputs[:bdy_fluxes]; apply_boundary_conditions_neumann!, BCs.jl:35-63, 655-816): the MOST wall model on the
    # faces tagged "MOST", surface integrals and RHS .+= S_flux on the device.  Dry runs, bulk_fluxes = false.
    if nsd == 3 && inputs[:bdy_fluxes]
        inputs[:bulk_fluxes] && error("libjexrhs: bulk surface fluxes need the microphysics state (not on the device path)")
        size(params.mp.Tabs, 1) == 1 || error("libjexrhs: the wall model is implemented for dry runs (no microphysics)")
        PC = J.PHYS_CONST
        fk = Int32[t == "MOST" ? 1 : 0 for t in mesh.bdy_face_type]
        mostc = Float64[PC.karman, 0.1, 0.01]                                 # z0_m, z0_h: the literals of BCs.jl:770-772
        check(c, ccall((:jx_upload_bdy_fluxes, LIB), Cint,
                       (Ctx, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                        Ptr{Float64}, Ptr{Int32}, Cint, Float64, Float64, Ptr{Float64}, Cint),
                       c, size(mesh.poin_in_bdy_face, 1), mesh.poin_in_bdy_face, mesh.bdy_face_in_elem, mesh.connijk,
                       metrics.nx, metrics.ny, metrics.nz, metrics.Jef, params.ω, fk, inputs[:ifirst_wall_node_index],
                       Float64(inputs[:δhf]), Float64(inputs[:user_heatflux]), mostc, length(mostc)))
    end
    # AssemblerCache (mpi_communications.jl:48-73).  Uploaded whenever its lists are non-empty -- also on ONE rank, where
    # periodic twins are summed by the reference's MPI self-send (mpi_communications.jl:99-112).
    cache = params.g_dss_cache
    if cache !== nothing && any(!isempty, cache.send_i)
        csr(v) = (Int64[0; cumsum(length.(v))], Int64.(reduce(vcat, v; init = Int64[])))
        sp, sv = csr(cache.send_i); rp, rv = csr(cache.recv_idx_buffers); bp, bv = csr(cache.recvback_idx_buffers)
        check(c, ccall((:jx_upload_halo, LIB), Cint, (Ctx, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                       c, sp, sv, rp, rv, bp, bv))
    end
    return B200Params(params, c, Vector{Float64}(undef, mesh.npoin * neqs))
end

"""
Same signature and semantics as rhs!(du,u,params,time) (rhs.jl:121-134): in place, returns nothing, mutates u.
`du` may be the ArrayFuse of OrdinaryDiffEq's low-storage methods (CarpenterKennedy2N54 with its default
`williamson_condition = true`; rhs.jl:14-27), which has no pointer and only scalar `setindex!`: then the result goes through
a host buffer and is stored element by element exactly like RHStoDU!, so the fused stage update still happens in the store.
"""
function rhs_b200!(du, u, p::B200Params, time)
    c = p.ctx
    if du isa Array{Float64}
        check(c, ccall((:jx_rhs, LIB), Cint, (Ctx, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c, time, u, du, u))
    else
        buf = p.dubuf
        check(c, ccall((:jx_rhs, LIB), Cint, (Ctx, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c, time, u, buf, u))
        @inbounds for i in eachindex(buf)
            du[i] = buf[i]
        end
    end
    return nothing
end

"""
rhs! on a device-resident state (`u`, `du` CuArray{Float64} on the context's device): no PCIe copy; `pointer(::CuArray)`
is a CuPtr whose bits are the device address jx_rhs_dev expects.
"""
function rhs_b200_dev!(du, u, p::B200Params, time)
    up = reinterpret(Ptr{Float64}, UInt(pointer(u))); dp = reinterpret(Ptr{Float64}, UInt(pointer(du)))
    check(p.ctx, ccall((:jx_rhs_dev, LIB), Cint, (Ctx, Float64, Ptr{Float64}, Ptr{Float64}), p.ctx, time, up, dp))
    return nothing
end

"Replaces solve(prob, alg; dt, adaptive=false) for the fixed-step explicit schemes: all stages on the device."
function step_b200!(u, p::B200Params, inputs, t, nsteps)
    c = p.ctx
    dt = Float64(Float32(inputs[:Δt]))                                        # TimeIntegrators.jl:464-465
    check(c, ccall((:jx_set_state, LIB), Cint, (Ctx, Ptr{Float64}), c, u))
    check(c, ccall((:jx_step, LIB), Cint, (Ctx, Cint, Float64, Float64, Cint), c, SCHEMES[string(nameof(typeof(inputs[:ode_solver])))], t, dt, nsteps))
    check(c, ccall((:jx_get_state, LIB), Cint, (Ctx, Ptr{Float64}), c, u))
    return t + nsteps * dt
end

b200_free(p::B200Params) = ccall((:jx_destroy, LIB), Cvoid, (Ctx,), p.ctx)

end # module
