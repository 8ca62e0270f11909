# GXBeamB200.jl — Julia glue a GXBeam maintainer would add next to src/analyses.jl.
#
# It leaves the public API intact (Assembly, PrescribedConditions, DistributedLoads, PointMass, static_analysis,
# AssemblyState) and adds batched entry points (`static_analysis_batch`, `steady_state_analysis_batch`, `eigenvalue_analysis_batch`,
# `time_domain_analysis_batch`, `static_analysis_multi_gpu`) that pack the inputs of many independent
# assemblies sharing a topology, `ccall`s libgxbeam_b200.so (include/gxbeam_b200.h), and rebuilds one AssemblyState per
# instance with the *existing* constructor (src/output.jl:106).
#
# NOTE: Julia is not available in the build image of this repository, so this file has not been executed here.  The same
# C symbols are exercised by the Python ctypes binding (gxbeam_b200/api.py) in tests/ and bench.py.
module GXBeamB200

using GXBeam, StaticArrays, LinearAlgebra

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
    structural_damping::Int32
    persistent::Int32
    update_linearization::Int32
    lu_parts::Int32
    exact_dphi0::Int32
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
        opts = Ref(Opts(linear, two_dimensional, ftol, iterations, 1e6, 1e-4, 0.5, 0.1, 0, 0, 0, 0, 0, 0))
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

"""
    time_domain_analysis_batch(assemblies, tvec; prescribed_conditions, linear_velocity, angular_velocity,
        linear_acceleration, angular_acceleration, u0, theta0, V0, Omega0, Vdot0, Omegadot0, bc_series, ...)
        -> (x_hist, dx_hist, converged)

Batched counterpart of `time_domain_analysis` (src/analyses.jl:2510-2790) for chains: `initial_condition_analysis!`
(:1695-1945) for every instance with `gxb_initial_condition`, then the Newmark march with `gxb_newmark_run` from the state left
in HBM.  `bc_series = (points, fields, values[K, nt, nb])` holds the time-varying prescribed values evaluated by the caller from
its `prescribed_conditions(t)` closures (field 0..17 = u, theta, F, M, Ff, Mf components).  `u0 ... Omegadot0` are `3 x np x nb`
arrays (zeros by default, like the reference).  Returns the saved state vectors `x_hist[nstates, nsave, nb]` and point rates
`dx_hist[12, np, nsave, nb]` (udot, thetadot, Vdot, Omegadot: the `dx` of `newmark_output`); each pair goes into
`AssemblyState(dx, x, system, assembly; prescribed_conditions)` exactly as at src/analyses.jl:3502.
"""
function time_domain_analysis_batch(assemblies::Vector{<:Assembly}, tvec;
        prescribed_conditions, linear_velocity=[zeros(3) for _ in assemblies], angular_velocity=[zeros(3) for _ in assemblies],
        linear_acceleration=[zeros(3) for _ in assemblies], angular_acceleration=[zeros(3) for _ in assemblies],
        gravity=[zeros(3) for _ in assemblies], structural_damping=true, two_dimensional=false, ftol=1e-9, iterations=1000,
        u0=nothing, theta0=nothing, V0=nothing, Omega0=nothing, Vdot0=nothing, Omegadot0=nothing, bc_series=nothing,
        save_every=1, device=0)
    nb = length(assemblies); a1 = assemblies[1]
    np, ne = length(a1.points), length(a1.elements)
    start = collect(Int64, a1.start); stop = collect(Int64, a1.stop)
    pd, pl, _ = pack_conditions(np, prescribed_conditions[1])
    bc = Array{Float64}(undef, 18, np, nb)
    for i in 1:nb
        bc[:, :, i] .= pack_conditions(np, prescribed_conditions[i])[3]
    end
    elems = pack_elements(assemblies)
    pts = reduce(vcat, [reinterpret(Float64, a.points) for a in assemblies])
    fs = [GXBeam.default_force_scaling(a) for a in assemblies]
    motion = reduce(hcat, [vcat(linear_velocity[i], angular_velocity[i], linear_acceleration[i], angular_acceleration[i]) for i in 1:nb])
    ic = zeros(18, np, nb)
    for (k, a) in enumerate((u0, theta0, V0, Omega0, Vdot0, Omegadot0))
        isnothing(a) || (ic[3k-2:3k, :, :] .= a)
    end
    opts = Ref(Opts(0, two_dimensional, ftol, iterations, 1e6, 1e-4, 0.5, 0.1, 0, structural_damping, 0, 0, 0, 0))
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gxb_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), ctx, device))
    try
        nstates = Ref{Int64}(0)
        check(ccall((:gxb_topology, LIB), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
            ctx[], 1, np, ne, start, stop, pd, pl, nstates, C_NULL, C_NULL, C_NULL, C_NULL))
        check(ccall((:gxb_set_batch, LIB), Cint, (Ptr{Cvoid}, Int64), ctx[], nb))
        check(ccall((:gxb_set_elements, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], elems))
        check(ccall((:gxb_set_points, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], pts))
        check(ccall((:gxb_set_bc, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], bc))
        check(ccall((:gxb_set_body, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx[], reduce(hcat, gravity), fs))
        check(ccall((:gxb_set_motion, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], motion))
        n = nstates[]
        x0 = zeros(n, nb); rates0 = zeros(12, np, nb); conv0 = zeros(Int32, nb); it0 = zeros(Int32, nb); rinf = zeros(nb)
        check(ccall((:gxb_initial_condition, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Opts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}),
            ctx[], opts, ic, C_NULL, x0, rates0, conv0, it0, rinf, C_NULL))
        if !isnothing(bc_series)
            points, fields, values = bc_series
            check(ccall((:gxb_set_bc_series, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                ctx[], length(tvec), length(points), Int32.(points), Int32.(fields), values))
        end
        nt = length(tvec); nsave = (nt - 1) ÷ save_every + 1
        xh = zeros(n, nsave, nb); dxh = zeros(12, np, nsave, nb); steps = zeros(Int32, nb); itot = zeros(Int32, nb); conv = zeros(Int32, nb)
        check(ccall((:gxb_newmark_run, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Opts}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
            ctx[], opts, nt, collect(Float64, tvec), C_NULL, C_NULL, save_every, xh, steps, itot, conv, dxh))
        return xh, dxh, (conv0 .!= 0) .& (conv .!= 0)
    finally
        ccall((:gxb_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    end
end

# ---- shared set-up of a kind = 1 (DynamicSystem) context: topology, batch, elements, points, conditions, body motion ----
function dynamic_context(assemblies, prescribed_conditions, point_masses, gravity, linear_velocity, angular_velocity,
        linear_acceleration, angular_acceleration, device)
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
    pts = reduce(vcat, [reinterpret(Float64, a.points) for a in assemblies])
    fs = [GXBeam.default_force_scaling(a) for a in assemblies]
    motion = reduce(hcat, [vcat(linear_velocity[i], angular_velocity[i], linear_acceleration[i], angular_acceleration[i]) for i in 1:nb])
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gxb_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), ctx, device))
    nstates = Ref{Int64}(0)
    check(ccall((:gxb_topology, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        ctx[], 1, np, ne, start, stop, pd, pl, nstates, C_NULL, C_NULL, C_NULL, C_NULL))
    check(ccall((:gxb_set_batch, LIB), Cint, (Ptr{Cvoid}, Int64), ctx[], nb))
    check(ccall((:gxb_set_elements, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], elems))
    check(ccall((:gxb_set_points, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], pts))
    check(ccall((:gxb_set_bc, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], bc))
    if point_masses !== nothing
        pm = zeros(36, np, nb)
        for i in 1:nb, (ip, m) in point_masses[i]
            pm[:, ip, i] .= vec(m.mass)
        end
        check(ccall((:gxb_set_pmass, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], pm))
    end
    check(ccall((:gxb_set_body, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx[], reduce(hcat, gravity), fs))
    check(ccall((:gxb_set_motion, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], motion))
    return ctx, nstates[], fs
end

"""
    steady_state_analysis_batch(assemblies; prescribed_conditions, point_masses, gravity, linear_velocity, angular_velocity,
        linear_acceleration, angular_acceleration, structural_damping, two_dimensional, ftol, iterations, device, keep_context)
        -> (states::Vector{AssemblyState}, converged::Vector{Bool}, iters, x)

Batched counterpart of `steady_state_analysis` (src/analyses.jl:300-520): `gxb_steady_solve` runs `nlsolve!` (:3556) on the
`steady_system_residual!/jacobian!` pair (src/system.jl:427-455, 693-720) of every instance.  `x[:, i]` is instance i's DynamicSystem
state vector, fed to the EXISTING `AssemblyState(x, system, assembly; prescribed_conditions)` (src/output.jl:106).
"""
function steady_state_analysis_batch(assemblies::Vector{<:Assembly};
        prescribed_conditions, point_masses=nothing, gravity=[zeros(3) for _ in assemblies],
        linear_velocity=[zeros(3) for _ in assemblies], angular_velocity=[zeros(3) for _ in assemblies],
        linear_acceleration=[zeros(3) for _ in assemblies], angular_acceleration=[zeros(3) for _ in assemblies],
        structural_damping=false, two_dimensional=false, ftol=1e-9, iterations=1000, device=0, keep_context=false)
    nb = length(assemblies)
    ctx, n, fs = dynamic_context(assemblies, prescribed_conditions, point_masses, gravity, linear_velocity, angular_velocity,
        linear_acceleration, angular_acceleration, device)
    try
        opts = Ref(Opts(0, two_dimensional, ftol, iterations, 1e6, 1e-4, 0.5, 0.1, 0, structural_damping, 0, 0, 0, 0))
        x = Array{Float64}(undef, n, nb); conv = zeros(Int32, nb); iters = zeros(Int32, nb); rinf = zeros(nb)
        check(ccall((:gxb_steady_solve, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Opts}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
            ctx[], opts, C_NULL, x, conv, iters, rinf))
        states = map(1:nb) do i
            system = DynamicSystem(assemblies[i]; force_scaling=fs[i])
            AssemblyState(view(x, :, i), system, assemblies[i]; prescribed_conditions=prescribed_conditions[i])
        end
        keep_context && return states, conv .!= 0, iters, x, ctx, opts
        return states, conv .!= 0, iters, x
    finally
        keep_context || ccall((:gxb_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    end
end

"""
    eigenvalue_analysis_batch(assemblies; nev = 6, m = max(3nev + 10, 30), tol = 1e-9, kwargs...) -> (λ, V, converged)

Batched counterpart of `eigenvalue_analysis` (src/analyses.jl:1180-1380): a steady-state analysis of every instance, then
`solve_eigensystem` (:943-965) about each steady state.  The device builds K (steady Jacobian) and M (`system_mass_matrix!`,
src/system.jl:961-995), factorises K once and runs m Arnoldi steps of v -> K^-1 (M v) (`gxb_eigen_arnoldi`); Julia finishes like
the reference does -- `eigen` of the m x m Hessenberg matrices, sort by (abs, imag) descending (:956), λ = -1/μ -- and
`gxb_eigen_ritz_vectors` forms the eigenvectors `Vk * F.vectors` (:959) on the device.  As with `partialschur(...; tol)` (:954), the
Krylov space is rebuilt larger until the nev leading Ritz pairs of every instance have residual estimates under `tol`.
λ: nev x nb, V: nstates x nev x nb (complex), converged: steady-state convergence flags.
"""
function eigenvalue_analysis_batch(assemblies::Vector{<:Assembly}; nev=6, m=max(3nev + 10, 30), tol=1e-9, kwargs...)
    nb = length(assemblies)
    _, conv, _, x, ctx, opts = steady_state_analysis_batch(assemblies; kwargs..., keep_context=true)
    try
        n = size(x, 1)
        m = min(m, n); mmax = min(n, 8m)
        while true
            H = Array{Float64}(undef, m, m + 1, nb)            # row-major (m+1) x m per instance = column-major m x (m+1)
            check(ccall((:gxb_eigen_arnoldi, LIB), Cint, (Ptr{Cvoid}, Ptr{Opts}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}),
                ctx[], opts, x, m, H, C_NULL))
            singular = zeros(Int32, nb)
            check(ccall((:gxb_eigen_status, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx[], singular))
            any(!=(0), singular) && throw(LinearAlgebra.SingularException(findfirst(!=(0), singular)))   # what `lu(K)` does at :948
            λ = Array{ComplexF64}(undef, nev, nb); Y = Array{ComplexF64}(undef, nev, m, nb); worst = 0.0
            for i in 1:nb
                Hi = permutedims(view(H, :, :, i))             # (m+1) x m
                F = eigen(Hi[1:m, :])
                perm = sortperm(F.values, by=(μ) -> (abs(μ), imag(μ)), rev=true)[1:nev]
                μ = F.values[perm]
                λ[:, i] .= -1 ./ μ
                Y[:, :, i] .= transpose(F.vectors[:, perm])    # Y[j, k, i]: the device wants nb x m x nev, nev fastest
                worst = max(worst, maximum(abs.(Hi[m + 1, m] .* F.vectors[m, perm]) ./ abs.(μ)))
            end
            if worst <= tol || m >= mmax
                V = Array{ComplexF64}(undef, nev, n, nb)       # nb x nstates x nev row-major
                check(ccall((:gxb_eigen_ritz_vectors, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx[], nev, Y, V))
                worst <= tol || @warn "eigenvalue_analysis_batch: tol = $tol not reached with a Krylov space of $m (worst Ritz residual $worst)"
                return λ, permutedims(V, (2, 1, 3)), conv
            end
            m = min(2m, mmax)
        end
    finally
        ccall((:gxb_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    end
end

"""
    static_analysis_multi_gpu(assemblies, devices; prescribed_conditions, ...) -> (x, converged, iters)

`static_analysis_batch` over several GPUs of one box through the library's own multi-GPU entry (`gxb_multi_*`): one context and one
host worker thread per GPU live INSIDE the library, the batch is cut into contiguous shards, every worker reads its slice of these
host arrays and writes its slice of `x` -- no Julia threads, no collective.
"""
function static_analysis_multi_gpu(assemblies::Vector{<:Assembly}, devices::Vector{<:Integer}; prescribed_conditions,
        gravity=[zeros(3) for _ in assemblies], linear=false, two_dimensional=false, ftol=1e-9, iterations=1000)
    nb = length(assemblies); a1 = assemblies[1]
    np, ne = length(a1.points), length(a1.elements)
    start = collect(Int64, a1.start); stop = collect(Int64, a1.stop)
    pd, pl, _ = pack_conditions(np, prescribed_conditions[1])
    bc = Array{Float64}(undef, 18, np, nb)
    for i in 1:nb
        bc[:, :, i] .= pack_conditions(np, prescribed_conditions[i])[3]
    end
    elems = pack_elements(assemblies)
    fs = [GXBeam.default_force_scaling(a) for a in assemblies]
    m = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gxb_multi_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Int32, Ptr{Int32}), m, length(devices), Int32.(devices)))
    try
        nstates = Ref{Int64}(0)
        check(ccall((:gxb_multi_topology, LIB), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
            m[], 0, np, ne, start, stop, pd, pl, nstates, C_NULL, C_NULL, C_NULL, C_NULL))
        check(ccall((:gxb_multi_set_batch, LIB), Cint, (Ptr{Cvoid}, Int64), m[], nb))
        opts = Ref(Opts(linear, two_dimensional, ftol, iterations, 1e6, 1e-4, 0.5, 0.1, 0, 0, 0, 0, 0, 0))
        x = Array{Float64}(undef, nstates[], nb); conv = zeros(Int32, nb); iters = zeros(Int32, nb); rinf = zeros(nb)
        check(ccall((:gxb_multi_solve, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Opts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
            m[], opts, elems, C_NULL, bc, C_NULL, C_NULL, reduce(hcat, gravity), fs, C_NULL, C_NULL, x, conv, iters, rinf, C_NULL))
        return x, conv .!= 0, iters
    finally
        ccall((:gxb_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), m[])
    end
end

"""
    compact_conditions(bc) -> (base, points, fields, values)

Split packed prescribed conditions `bc[18, np, nb]` into the record all instances share and the entries that differ between instances
(`gxb_set_bc_compact`): `points` / `fields` are 1-based point numbers and 0-based field numbers (u, theta, F, M, Ff, Mf components),
`values[K, nb]` the per-instance entries.  NaN (unused) entries compare equal.
"""
function compact_conditions(bc::Array{Float64,3})
    np, nb = size(bc, 2), size(bc, 3)
    points = Int32[]; fields = Int32[]
    for p in 1:np, f in 1:18
        any(i -> !isequal(bc[f, p, i], bc[f, p, 1]), 2:nb) && (push!(points, p); push!(fields, f - 1))
    end
    values = Matrix{Float64}(undef, length(points), nb)
    for i in 1:nb, k in eachindex(points)
        values[k, i] = bc[fields[k] + 1, points[k], i]
    end
    return bc[:, :, 1], points, fields, values
end

"""
    static_analysis_pipelined(assemblies; prescribed_conditions, lanes = 4, device = 0, ...) -> (x, converged, iters)

`static_analysis_batch` for a design sweep without gravity, through ONE call of `gxb_multi_solve_compact`: `lanes` shards share GPU `device`
as pipeline lanes (shard k solves while shard k + 1 crosses PCIe and shard k - 1 is copied back), only the fields of `Element` that the
static path reads are fetched (`gxb_multi_hint_no_gravity`), and the prescribed conditions travel as one shared record plus the entries
that differ between instances.  This is the call `bench.py` times as its end-to-end figure.
"""
function static_analysis_pipelined(assemblies::Vector{<:Assembly}; prescribed_conditions, lanes=4, device=0,
        linear=false, two_dimensional=false, ftol=1e-9, iterations=1000)
    nb = length(assemblies); a1 = assemblies[1]
    np, ne = length(a1.points), length(a1.elements)
    start = collect(Int64, a1.start); stop = collect(Int64, a1.stop)
    pd, pl, _ = pack_conditions(np, prescribed_conditions[1])
    bc = Array{Float64}(undef, 18, np, nb)
    for i in 1:nb
        bc[:, :, i] .= pack_conditions(np, prescribed_conditions[i])[3]
    end
    base, points, fields, values = compact_conditions(bc)
    elems = pack_elements(assemblies)
    fs = [GXBeam.default_force_scaling(a) for a in assemblies]
    m = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gxb_multi_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Int32, Ptr{Int32}), m, lanes, fill(Int32(device), lanes)))
    try
        nstates = Ref{Int64}(0)
        check(ccall((:gxb_multi_topology, LIB), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
            m[], 0, np, ne, start, stop, pd, pl, nstates, C_NULL, C_NULL, C_NULL, C_NULL))
        check(ccall((:gxb_multi_set_batch, LIB), Cint, (Ptr{Cvoid}, Int64), m[], nb))
        check(ccall((:gxb_multi_hint_no_gravity, LIB), Cint, (Ptr{Cvoid}, Cint), m[], 1))
        opts = Ref(Opts(linear, two_dimensional, ftol, iterations, 1e6, 1e-4, 0.5, 0.1, 0, 0, 0, 0, 0, 0))
        x = Array{Float64}(undef, nstates[], nb); conv = zeros(Int32, nb); iters = zeros(Int32, nb); rinf = zeros(nb)
        check(ccall((:gxb_multi_solve_compact, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Opts}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
            m[], opts, elems, base, length(points), points, fields, values, C_NULL, fs, C_NULL, x, conv, iters, rinf, C_NULL))
        return x, conv .!= 0, iters
    finally
        ccall((:gxb_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), m[])
    end
end

"""
    section_properties_batch(nodes, elements; gxbeam_order = true, shear_center = true, device = 0) -> (S, sc, tc, M, mc)

Batched counterpart of `compliance_matrix(nodes, elements)` and `mass_matrix(nodes, elements)` (src/section.jl:641-783, 840-890) for
layups that share one mesh topology: `nodes[i]::Vector{Node{Float64}}` and `elements[i]::Vector{MeshElement{Float64}}` are instance
i's mesh (same `nodenum` everywhere; coordinates, materials and ply angles may differ).  `S[:, :, i]`, `M[:, :, i]` are returned in
Julia's column-major order (the library writes row-major 6 x 6 blocks: transposed on the way out).
"""
function section_properties_batch(nodes::Vector{<:Vector}, elements::Vector{<:Vector}; gxbeam_order=true, shear_center=true, device=0)
    nb = length(nodes); nn = length(nodes[1]); ne = length(elements[1])
    conn = Matrix{Int64}(undef, 4, ne)
    for (e, el) in enumerate(elements[1]); conn[:, e] .= el.nodenum; end
    xy = Array{Float64}(undef, 2, nn, nb); mats = Array{Float64}(undef, 10, ne, nb); theta = Array{Float64}(undef, ne, nb)
    for i in 1:nb
        for (k, n) in enumerate(nodes[i]); xy[1, k, i] = n.x; xy[2, k, i] = n.y; end
        for (e, el) in enumerate(elements[i])
            m = el.material
            mats[:, e, i] .= (m.E1, m.E2, m.E3, m.G12, m.G13, m.G23, m.nu12, m.nu13, m.nu23, m.rho)
            theta[e, i] = el.theta
        end
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gxb_section_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Int64, Int64, Ptr{Int64}), h, device, nn, ne, conn))
    try
        S = Array{Float64}(undef, 6, 6, nb); M = similar(S); sc = zeros(2, nb); tc = zeros(2, nb); mc = zeros(2, nb); ok = zeros(Int32, nb)
        check(ccall((:gxb_section_properties, LIB), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}),
            h[], nb, xy, mats, theta, gxbeam_order, shear_center, S, sc, tc, M, mc, ok, C_NULL))
        all(!=(0), ok) || error("section_properties_batch: degenerate mesh in instance $(findfirst(==(0), ok))")
        return permutedims(S, (2, 1, 3)), sc, tc, permutedims(M, (2, 1, 3)), mc
    finally
        ccall((:gxb_section_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

"""
    strain_recovery_batch(F, M, nodes, elements; gxbeam_order = true, failure = true, device = 0)
        -> (strain_b, stress_b, strain_p, stress_p, tsai_wu)

Batched counterpart of `strain_recovery(F, M, nodes, elements, cache; gxbeam_order)` (src/section.jl:952-1040) followed by
`tsai_wu(stress_p, elements)` (:1108-1131): `F[i]`, `M[i]` are the section loads of layup i.  The compliance solve that fills the
reference's `SectionCache` runs on the device and stays there.  Every output is `6 x ne x nb` (`ne x nb` for the Tsai-Wu index), i.e.
`strain_b[:, :, i]` is what the reference returns for layup i.
"""
function strain_recovery_batch(F::Vector, M::Vector, nodes::Vector{<:Vector}, elements::Vector{<:Vector}; gxbeam_order=true, failure=true, device=0)
    nb = length(nodes); nn = length(nodes[1]); ne = length(elements[1])
    conn = Matrix{Int64}(undef, 4, ne)
    for (e, el) in enumerate(elements[1]); conn[:, e] .= el.nodenum; end
    xy = Array{Float64}(undef, 2, nn, nb); mats = Array{Float64}(undef, 10, ne, nb); theta = Array{Float64}(undef, ne, nb)
    str = Array{Float64}(undef, 9, ne, nb); Fm = Array{Float64}(undef, 3, nb); Mm = similar(Fm)
    for i in 1:nb
        Fm[:, i] .= F[i]; Mm[:, i] .= M[i]
        for (k, n) in enumerate(nodes[i]); xy[1, k, i] = n.x; xy[2, k, i] = n.y; end
        for (e, el) in enumerate(elements[i])
            m = el.material
            mats[:, e, i] .= (m.E1, m.E2, m.E3, m.G12, m.G13, m.G23, m.nu12, m.nu13, m.nu23, m.rho)
            str[:, e, i] .= (m.S1t, m.S1c, m.S2t, m.S2c, m.S3t, m.S3c, m.S12, m.S13, m.S23)
            theta[e, i] = el.theta
        end
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gxb_section_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Int64, Int64, Ptr{Int64}), h, device, nn, ne, conn))
    try
        eb = Array{Float64}(undef, 6, ne, nb); sb = similar(eb); ep = similar(eb); sp = similar(eb); tw = Array{Float64}(undef, ne, nb)
        ok = zeros(Int32, nb)
        check(ccall((:gxb_section_strain_recovery, LIB), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}),
            h[], nb, xy, mats, theta, failure ? pointer(str) : C_NULL, Fm, Mm, gxbeam_order, eb, sb, ep, sp, failure ? pointer(tw) : C_NULL, C_NULL, ok, C_NULL))
        all(!=(0), ok) || error("strain_recovery_batch: degenerate mesh in instance $(findfirst(==(0), ok))")
        return eb, sb, ep, sp, (failure ? tw : nothing)
    finally
        ccall((:gxb_section_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

end # module
