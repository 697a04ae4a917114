# rhs_b200.jl -- reference-side binding of libjexrhs (include/jexrhs.h) for Jexpresso.
#
# Drop this file into src/kernel/operators/ and `include` it after rhs.jl.  It adds ONE backend
# branch to the reference: when `inputs[:backend]` is `B200()` the ODEProblem is built on
# `rhs_b200!` instead of `rhs!` (src/kernel/solvers/TimeIntegrators.jl:210-213), everything else
# (mesh, setup, integrator, callbacks, I/O) is untouched.  Cannot be executed in the authoring
# image (no Julia); its ctypes twin jexpresso_b200/{capi,rhs}.py is what the tests run.
module JexRHSB200

const LIB = get(ENV, "JEXRHS_LIB", "libjexrhs.so")
const Ctx = Ptr{Cvoid}

struct B200 end                                    # value for inputs[:backend]

const EQ_IDS = Dict("CompEuler" => 0, "CompEulerEnergy" => 1, "AdvDiff" => 2, "ShallowWater" => 3)
const SCHEMES = Dict("CarpenterKennedy2N54" => 0, "SSPRK54" => 1, "SSPRK33" => 2)
const PERIODIC = ("periodicx", "periodicy", "periodicz", "periodic1", "periodic2", "periodic3", "Laguerre")

function check(ctx::Ctx, rc::Cint)
    rc == 0 && return
    buf = Vector{UInt8}(undef, 512)
    ccall((:jx_last_error, LIB), Cint, (Ctx, Ptr{UInt8}, Cint), ctx, buf, 512)
    error("libjexrhs error $rc: " * unsafe_string(pointer(buf)))
end

"Build the device-resident problem from what params_setup returns (params_setup.jl:442-479)."
function b200_setup(params, inputs; device = 0, rank = 0, nranks = 1, uid = C_NULL)
    ctx = Ref{Ctx}(C_NULL)
    rc = ccall((:jx_init, LIB), Cint, (Cint, Cint, Cint, Ptr{Cvoid}, Ref{Ctx}), device, rank, nranks, uid, ctx)
    rc == 0 || error("jx_init failed ($rc): no CUDA device / NCCL")
    c = ctx[]
    mesh, metrics, basis = params.mesh, params.metrics, params.basis
    nsd = mesh.nsd; ngl = mesh.ngl; neqs = params.neqs
    PhysConst = params.PhysConst
    phys = Float64[PhysConst.C0, PhysConst.γ, PhysConst.g, PhysConst.Rair, PhysConst.cp, PhysConst.cv,
                   PhysConst.pref, PhysConst.γm1, 0.0, 0.0, 0.0]
    # engine options (include/jexrhs.h): inputs[:b200_dss] = :gather (deterministic, reference summation order, default)
    # or :atomics (throughput mode); the element kernel is chosen by the library (JX_ELEM_AUTO) unless
    # inputs[:b200_elem_kernel] names a variant; the shared jx_pow keeps the equation of state reproducible
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 1, get(inputs, :b200_dss, :gather) == :atomics ? 1 : 0))
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 2, get(inputs, :b200_pow, 1)))
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 3, get(inputs, :b200_elem_kernel, 0)))
    # inputs[:b200_overlap] = n > 0 (atomics mode, several ranks): interface elements first, the NCCL exchange beside the
    # interior launch with n SMs left to it; set ENV["NCCL_MAX_CTAS"] = ENV["NCCL_MAX_P2P_NCHANNELS"] = string(n) before
    # jx_init so that the send/recv kernels fit on those SMs (JX_OPT_OVERLAP)
    check(c, ccall((:jx_set_option, LIB), Cint, (Ctx, Cint, Int64), c, 5, get(inputs, :b200_overlap, 0)))
    μ = Float64.(params.visc_coeff)                                          # inputs[:μ], params_setup.jl:307-315
    lpert = inputs[:SOL_VARS_TYPE] == PERT() ? 1 : 0
    check(c, ccall((:jx_set_problem, LIB), Cint,
                   (Ctx, Cint, Cint, Cint, Int64, Int64, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Cint),
                   c, nsd, ngl, neqs, mesh.nelem, mesh.npoin, EQ_IDS[inputs[:equations]], lpert,
                   inputs[:lsource] ? 1 : 0, inputs[:lvisc] ? 1 : 0, μ, phys, length(phys)))
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
    cache = params.g_dss_cache                                               # AssemblerCache, mpi_communications.jl:48-73
    if cache !== nothing && nranks > 1
        csr(v) = (Int64[0; cumsum(length.(v))], Int64.(reduce(vcat, v; init = Int64[])))
        sp, sv = csr(cache.send_i); rp, rv = csr(cache.recv_idx_buffers); bp, bv = csr(cache.recvback_idx_buffers)
        check(c, ccall((:jx_upload_halo, LIB), Cint, (Ctx, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                       c, sp, sv, rp, rv, bp, bv))
    end
    return c
end

"Same signature and semantics as rhs!(du,u,params,time) (rhs.jl:121-134): in place, returns nothing, mutates u."
function rhs_b200!(du, u, params, time)
    c = params.b200ctx
    check(c, ccall((:jx_rhs, LIB), Cint, (Ctx, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c, time, u, du, u))
    return nothing
end

"Replaces solve(prob, alg; dt, adaptive=false) for the fixed-step explicit schemes: all stages on the device."
function step_b200!(u, params, inputs, t, nsteps)
    c = params.b200ctx
    dt = Float64(Float32(inputs[:Δt]))                                        # TimeIntegrators.jl:464-465
    check(c, ccall((:jx_set_state, LIB), Cint, (Ctx, Ptr{Float64}), c, u))
    check(c, ccall((:jx_step, LIB), Cint, (Ctx, Cint, Float64, Float64, Cint), c, SCHEMES[string(nameof(typeof(inputs[:ode_solver])))], t, dt, nsteps))
    check(c, ccall((:jx_get_state, LIB), Cint, (Ctx, Ptr{Float64}), c, u))
    return t + nsteps * dt
end

b200_free(c::Ctx) = ccall((:jx_destroy, LIB), Cvoid, (Ctx,), c)

end # module
