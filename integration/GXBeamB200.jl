# GXBeamB200.jl — Julia glue a GXBeam maintainer would add next to src/analyses.jl.
#
# It leaves the public API intact (Assembly, PrescribedConditions, DistributedLoads, PointMass, static_analysis,
# AssemblyState) and adds one batched entry point, `static_analysis_batch`, that packs the inputs of many independent
# assemblies sharing a topology, `ccall`s libgxbeam_b200.so (include/gxbeam_b200.h), and rebuilds one AssemblyState per
# instance with the *existing* constructor (src/output.jl:106).
#
# NOTE: Julia is not available in the build image of this repository, so this file has not been executed here.  The same
# C symbols are exercised by the Python ctypes binding (gxbeam_b200/api.py) in tests/ and bench.py.
module GXBeamB200

using GXBeam, StaticArrays

const LIB = get(ENV, "GXBEAM_B200_LIB", "libgxbeam_b200.so")

struct Opts            # mirrors gxb_opts
    linear::Int32
    two_dimensional::Int32
    ftol::Float64
    iterations::Int64
    maxstep::Float64
    c1::Float64
    rho_hi::Float64
    rho_lo::Float64
    profile::Int32
    reserved::Int32
end

check(rc) = rc == 0 || error("gxbeam_b200: " * unsafe_string(ccall((:gxb_last_error, LIB), Cstring, ())))

# Element{Float64} is isbits with the layout L, x[3], compliance[36], mass[36], Cab[9], mu[6] (src/assembly.jl:14-21), so a
# Vector{Element{Float64}} can be handed over as is (91 doubles per element, column-major matrices).
pack_elements(assemblies) = reduce(vcat, [reinterpret(Float64, a.elements) for a in assemblies])

function pack_conditions(np, pcs::AbstractDict)
    pd = zeros(UInt8, 6, np); pl = ones(UInt8, 6, np); bc = zeros(18, np); bc[1:6, :] .= NaN
    for (ip, pc) in pcs
        pd[:, ip] .= pc.pd; pl[:, ip] .= pc.pl
        bc[:, ip] .= vcat(pc.u, pc.theta, pc.F, pc.M, pc.Ff, pc.Mf)
    end
    return pd, pl, bc
end

"""
    static_analysis_batch(assemblies; prescribed_conditions, distributed_loads, point_masses, gravity, linear,
        two_dimensional, ftol, iterations, device) -> (states::Vector{AssemblyState}, converged::Vector{Bool}, iters)

Batched counterpart of `static_analysis` (src/analyses.jl:66-179).  `prescribed_conditions[i]` etc. are the per-instance
dictionaries; all instances must share `start`, `stop` and the pd/pl flags.
"""
function static_analysis_batch(assemblies::Vector{<:Assembly};
        prescribed_conditions, distributed_loads=nothing, point_masses=nothing,
        gravity=[zeros(3) for _ in assemblies], linear=false, two_dimensional=false, ftol=1e-9, iterations=1000, device=0)
    nb = length(assemblies); a1 = assemblies[1]
    np, ne = length(a1.points), length(a1.elements)
    start = collect(Int64, a1.start); stop = collect(Int64, a1.stop)
    pd, pl, _ = pack_conditions(np, prescribed_conditions[1])
    bc = Array{Float64}(undef, 18, np, nb)
    for i in 1:nb
        pdi, pli, bci = pack_conditions(np, prescribed_conditions[i])
        (pdi == pd && pli == pl) || error("all instances of a batch must share the pd/pl pattern")
        bc[:, :, i] .= bci
    end
    elems = pack_elements(assemblies)
    fs = [GXBeam.default_force_scaling(a) for a in assemblies]
    grav = reduce(hcat, gravity)

    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gxb_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), ctx, device))
    try
        nstates = Ref{Int64}(0)
        irp = zeros(Int64, np); ire = zeros(Int64, ne); icp = zeros(Int64, np); ice = zeros(Int64, ne)
        check(ccall((:gxb_topology, LIB), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
            ctx[], 0, np, ne, start, stop, pd, pl, nstates, irp, ire, icp, ice))
        # the library's indices are GXBeam's SystemIndices bit for bit
        sys = StaticSystem(a1); @assert irp == sys.indices.irow_point && ice == sys.indices.icol_elem && nstates[] == sys.indices.nstates
        check(ccall((:gxb_set_batch, LIB), Cint, (Ptr{Cvoid}, Int64), ctx[], nb))
        check(ccall((:gxb_set_elements, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], elems))
        check(ccall((:gxb_set_bc, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], bc))
        if distributed_loads !== nothing
            dl = zeros(24, ne, nb)
            for i in 1:nb, (ie, d) in distributed_loads[i]
                dl[:, ie, i] .= vcat(d.f1, d.f2, d.m1, d.m2, d.f1_follower, d.f2_follower, d.m1_follower, d.m2_follower)
            end
            check(ccall((:gxb_set_dloads, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], dl))
        end
        if point_masses !== nothing
            pm = zeros(36, np, nb)
            for i in 1:nb, (ip, m) in point_masses[i]
                pm[:, ip, i] .= vec(m.mass)
            end
            check(ccall((:gxb_set_pmass, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], pm))
        end
        check(ccall((:gxb_set_body, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx[], grav, fs))
        opts = Ref(Opts(linear, two_dimensional, ftol, iterations, 1e6, 1e-4, 0.5, 0.1, 0, 0))
        x = Array{Float64}(undef, nstates[], nb); conv = zeros(Int32, nb); iters = zeros(Int32, nb); rinf = zeros(nb)
        GC.@preserve elems bc grav fs x conv iters rinf begin
            check(ccall((:gxb_static_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Opts}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                ctx[], opts, C_NULL, x, conv, iters, rinf))
        end
        states = map(1:nb) do i
            system = StaticSystem(assemblies[i]; force_scaling=fs[i])
            AssemblyState(view(x, :, i), system, assemblies[i]; prescribed_conditions=prescribed_conditions[i])
        end
        return states, conv .!= 0, iters
    finally
        ccall((:gxb_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    end
end

end # module
