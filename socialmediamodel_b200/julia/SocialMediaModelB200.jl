# SocialMediaModelB200.jl -- the reference-side binding: `ccall` shim that keeps the exported API of
# LuzieH/SocialMediaModel (src/SocialMediaModel.jl:12-16) and forwards the hot path to libsmm_b200.so
# (C ABI: include/smm_b200.h).  A maintainer adds this file to src/ and `include`s it AFTER abm.jl,
# pde.jl and ensemble.jl; it overrides ABMsolve, ABMsolveensemble, PDEsolve, PDEsolveensemble and `f`.
#
# NOTE: Julia is not installed in the build image, so this file has not been executed there; the
# same entry points, argument order and memory layouts are exercised through the ctypes mirror
# socialmediamodel_b200/_lib.py by the test-suite (Julia arrays are column-major Float64 = NumPy
# arrays in Fortran order, which is what the tests pass).
module SocialMediaModelB200

using Libdl
using RecursiveArrayTools: ArrayPartition
# the parameter constructors and initial conditions stay the reference's own (setting.jl, pde.jl:90-180)
import ..SocialMediaModel: ABMconstruct, PDEconstruct, parameters, constructinitial

const LIB = get(ENV, "SMM_B200_LIB", joinpath(@__DIR__, "..", "libsmm_b200.so"))

# ---- mirror of the C structs (include/smm_b200.h) -----------------------------------------------
struct SmmParams
    n::Int32; L::Int32; M::Int32
    eta::Float64; a::Float64; b::Float64; c::Float64; sigma::Float64; sigmahat::Float64; sigmatilde::Float64
    frictionI::Float64; frictionM::Float64; dt::Float64
    dom_lo::NTuple{2,Float64}; dom_hi::NTuple{2,Float64}
    controlspeed1::Float64; controltarget1::NTuple{2,Float64}
    controlspeed2::Float64; controltarget2::NTuple{2,Float64}
end

struct SmmNoise
    mode::Int32; seed::UInt64
    xi::Ptr{Float64}; u1::Ptr{Float64}; u2::Ptr{Float64}; xi_inf::Ptr{Float64}; xi_med::Ptr{Float64}
end

struct SmmGrid
    Nx::Int32; Ny::Int32; x0::Float64; y0::Float64; dx::Float64; dy::Float64
end

const PREC_F32 = Int32(0); const PREC_F64 = Int32(1)
const RK4 = Int32(3)

function params(q, p; dt = hasproperty(p, :dt) ? p.dt : 0.01)
    d = p.domain
    SmmParams(q.n, q.L, q.M, q.eta, q.a, q.b, q.c, q.sigma, q.sigmahat, q.sigmatilde, q.frictionI, q.frictionM, dt,
              (d[1, 1], d[2, 1]), (d[1, 2], d[2, 2]),
              q.controlspeed1, (q.controltarget1[1], q.controltarget1[2]),
              q.controlspeed2, (q.controltarget2[1], q.controltarget2[2]))
end

grid(p) = SmmGrid(size(p.X, 1), size(p.X, 2), p.domain[1, 1], p.domain[2, 1], p.dx, p.dy)

mutable struct Ctx
    h::Ptr{Cvoid}
    devices::Any
end
const CTX = Ref{Union{Nothing,Ctx}}(nothing)

"""
    use_devices!(devices)

Select the GPUs behind the library context of this Julia process: `nothing` = one GPU (`LOCAL_RANK`, default 0),
`:all` = every visible GPU, or a vector of device ids.  With several GPUs the ensemble calls below shard their
realizations / instances over them INSIDE the one `ccall` (smm_ctx_create_multi: one host thread per GPU, NCCL
all-reduce of the ensemble sums), so the unchanged `ABMsolveensemble(200, 1024)` uses the whole box from one task.
The environment variable SMM_B200_DEVICES ("all" or "0,1,2,3") sets the default.
"""
function use_devices!(devices)
    if CTX[] !== nothing
        ccall((:smm_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), CTX[].h)
        CTX[] = nothing
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if devices === nothing
        dev = parse(Int, get(ENV, "LOCAL_RANK", "0"))
        rc = ccall((:smm_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), dev, h)
    elseif devices === :all
        rc = ccall((:smm_ctx_create_multi, LIB), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), 0, C_NULL, h)
    else
        ids = Cint.(collect(devices))
        rc = ccall((:smm_ctx_create_multi, LIB), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), length(ids), ids, h)
    end
    rc == 0 || error("smm_ctx_create failed [$rc]: ", unsafe_string(ccall((:smm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    CTX[] = Ctx(h[], devices)
    return nothing
end

function ctx(; devices = :default)
    if CTX[] === nothing || (devices !== :default && devices != CTX[].devices)
        if devices === :default
            e = get(ENV, "SMM_B200_DEVICES", "")
            devices = e == "" ? nothing : (e == "all" ? :all : parse.(Int, split(e, ",")))
        end
        use_devices!(devices)
    end
    return CTX[].h
end

gpu_count() = Int(ccall((:smm_ctx_gpu_count, LIB), Cint, (Ptr{Cvoid},), ctx()))

check(rc, what) = rc == 0 || error("$what failed [$rc]: ", unsafe_string(ccall((:smm_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx())))

# ---- ABMsolve (abm.jl:154-217) ---------------------------------------------------------------------
# Same signature and return value; `precision` is an added keyword (reference behaviour = Float64).
# Keyword extensions (reference defaults reproduce the reference call): `network = (rowptr, colidx)` a 0-based CSR follower
# network (row j = the agents i with IndNet[j,i] == 1, abm.jl:51; list j itself to keep the self-loop of abm.jl:42) instead of
# the all-to-all IndNet; L may be any square number g^2 (influencers initialised on a g x g partition of the domain,
# generalising abm.jl:21-24); `save_stride` thins the returned history (every save_stride-th step and the last).
function ABMsolve(NT = 100; p = ABMconstruct(), q = parameters(), init = "4inf", chosenseed = 0, precision = PREC_F32,
                  network = nothing, save_stride = 1)
    n, L, M = q.n, q.L, q.M
    sp = params(q, p)
    x0 = zeros(n, 2, 1); state = zeros(Int32, n, 1); follow0 = zeros(Int32, n, 1); inf0 = zeros(L, 2, 1); media0 = zeros(M, 2, 1)
    check(ccall((:smm_abm_init, LIB), Cint,
                (Ptr{Cvoid}, Ref{SmmParams}, UInt64, Int64, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                ctx(), sp, chosenseed, 0, 1, x0, state, follow0, inf0, media0), "smm_abm_init")
    savesteps = Int32.(unique(vcat(collect(0:save_stride:NT), NT)))
    ns = length(savesteps)
    xs = zeros(n, 2, ns, 1); fol = zeros(Int32, n, ns, 1); infs = zeros(L, 2, NT + 1, 1); meds = zeros(M, 2, NT + 1, 1)
    rowptr = network === nothing ? Ptr{Int32}(C_NULL) : Int32.(network[1])
    colidx = network === nothing ? Ptr{Int32}(C_NULL) : Int32.(network[2])
    noise = SmmNoise(0, chosenseed, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)
    check(ccall((:smm_abm_solve, LIB), Cint,
                (Ptr{Cvoid}, Ref{SmmParams}, Int32, Int32, Int32, Int64,
                 Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ref{SmmNoise},
                 Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                ctx(), sp, precision, NT, 1, 0, x0, state, follow0, inf0, media0, rowptr, colidx, noise,
                ns, savesteps, xs, fol, infs, meds, C_NULL), "smm_abm_solve")
    xs_l = [xs[:, :, k, 1] for k in 1:ns]
    stateinfs = [Float64.(fol[:, k, 1]) .+ 1.0 for k in 1:ns]       # abm.jl:164,212: Float64 ids 1.0..L
    infs_l = [infs[:, :, Int(k) + 1, 1] for k in savesteps]
    meds_l = [meds[:, :, Int(k) + 1, 1] for k in savesteps]
    return xs_l, stateinfs, infs_l, meds_l, Int.(state[:, 1]), (p, q)
end

# ---- ABMsolveensemble (ensemble.jl:11-55) -------------------------------------------------------------
# Keyword extensions (defaults = the reference call): `devices` (see use_devices!), `us_type = Float32` returns the
# densities as Float32 (smm_abm_ensemble_f32: half the PCIe and host bytes), `want_us = false` returns the ensemble MEANS
# (av_u, av_z, av_y of plotting.jl:361-363) instead of the per-realization arrays: 1 MB instead of 2 GB.
function ABMsolveensemble(NT = 200, N = 10; savepoints = 4, init = "4inf", p = ABMconstruct(), q = parameters(),
                          sigma = 0.1, save = true, precision = PREC_F32, devices = :default, us_type = Float64, want_us = true)
    L = q.L
    g = grid(p)
    h = ctx(; devices = devices)
    savesteps = Int32.(Int.(round.(LinRange(1, NT, savepoints))) .- 1)   # ensemble.jl:21, 0-based
    noise = SmmNoise(0, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)
    zs = Array{Float64}(undef, 2, 2, savepoints, N); ys = Array{Float64}(undef, 2, L, savepoints, N)
    if !want_us
        su = Array{Float64}(undef, g.Nx, g.Ny, 2, L, savepoints); sz = Array{Float64}(undef, 2, 2, savepoints); sy = Array{Float64}(undef, 2, L, savepoints)
        check(ccall((:smm_abm_ensemble, LIB), Cint,
                    (Ptr{Cvoid}, Ref{SmmParams}, Int32, Int32, Int32, Int64,
                     Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ref{SmmNoise},
                     Int32, Ptr{Int32}, Float64, Ref{SmmGrid},
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                    h, params(q, p), precision, NT, N, 1, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, noise,
                    savepoints, savesteps, sigma, g, C_NULL, C_NULL, C_NULL, su, sz, sy, C_NULL), "smm_abm_ensemble")
        return su ./ N, sz ./ N, sy ./ N, (p, q)
    end
    us = Array{us_type}(undef, g.Nx, g.Ny, 2, L, savepoints, N)          # every element is written by the library
    if us_type === Float32
        check(ccall((:smm_abm_ensemble_f32, LIB), Cint,
                    (Ptr{Cvoid}, Ref{SmmParams}, Int32, Int32, Int64,
                     Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ref{SmmNoise},
                     Int32, Ptr{Int32}, Float64, Ref{SmmGrid},
                     Ptr{Float32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                    h, params(q, p), NT, N, 1, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, noise,
                    savepoints, savesteps, sigma, g, us, zs, ys, C_NULL, C_NULL, C_NULL, C_NULL), "smm_abm_ensemble_f32")
    else
        check(ccall((:smm_abm_ensemble, LIB), Cint,
                    (Ptr{Cvoid}, Ref{SmmParams}, Int32, Int32, Int32, Int64,
                     Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ref{SmmNoise},
                     Int32, Ptr{Int32}, Float64, Ref{SmmGrid},
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                    h, params(q, p), precision, NT, N, 1,                      # realizations k = 1..N (ensemble.jl:25)
                    C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, noise,
                    savepoints, savesteps, sigma, g, us, zs, ys, C_NULL, C_NULL, C_NULL, C_NULL), "smm_abm_ensemble")
    end
    if save
        us2 = us; zs2 = zs; ys2 = ys; p2 = p; q2 = q
        @eval Main using JLD2
        Main.JLD2.jldsave(string("data/abm_ensemble_", init, ".jld2"); us2, zs2, ys2, p2, q2)   # ensemble.jl:46-53
    end
    return us, zs, ys, (p, q)
end

# ---- f (pde.jl:183-268): in-place right-hand side, usable as ODEProblem(f, ...) -------------------------
function f(duzy, uzy, (p, q), t; precision = PREC_F64)
    u, z, y2 = uzy.x
    du, dz, dy2 = duzy.x
    L = q.L
    ci = Float64.(collect(q.controlled_inf)); cm = Float64.(collect(q.controlled_med))
    qq = [params(q, p)]
    check(ccall((:smm_pde_rhs, LIB), Cint,
                (Ptr{Cvoid}, Ref{SmmGrid}, Ptr{SmmParams}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx(), grid(p), qq, 1, L, precision, ci, cm, u, z, y2, du, dz, dy2), "smm_pde_rhs")
    return nothing
end

# ---- PDEsolve (pde.jl:285-294) with the fixed-step RK4 graph instead of DifferentialEquations.solve ----
struct FixedStepSolution
    t::Vector{Float64}
    u::Vector{Any}          # ArrayPartition(u, z, y) per saved time
end
function (sol::FixedStepSolution)(t)
    k = searchsortedfirst(sol.t, t)
    (k <= length(sol.t) && isapprox(sol.t[k], t; atol = 1e-12)) && return sol.u[k]
    w = (t - sol.t[k-1]) / (sol.t[k] - sol.t[k-1])
    return (1 - w) * sol.u[k-1] + w * sol.u[k]
end

function PDEsolve(tmax = 0.1; alg = nothing, init = "4inf", p = PDEconstruct(), q = parameters(), dt = nothing,
                  nsaves = 101, precision = PREC_F32)
    uzy0, q = constructinitial(init, (p, q))
    D = 0.5 * q.sigma^2
    dt === nothing && (dt = min(1e-3, 0.16 * p.dx^2 / D))
    nsteps = max(1, ceil(Int, tmax / dt - 1e-9)); dt = tmax / nsteps
    savesteps = unique(Int32.(round.(LinRange(0, nsteps, min(nsaves, nsteps + 1)))))
    ns = length(savesteps); L = q.L
    u0, z0, y0 = uzy0.x
    us = zeros(p.N_x, p.N_y, 2, L, ns, 1); zs = zeros(2, 2, ns, 1); ys = zeros(2, L, ns, 1)
    ci = Float64.(collect(q.controlled_inf)); cm = Float64.(collect(q.controlled_med))
    check(ccall((:smm_pde_solve, LIB), Cint,
                (Ptr{Cvoid}, Ref{SmmGrid}, Ptr{SmmParams}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Int32, Int32, Int32, Ptr{Int32},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx(), grid(p), [params(q, p)], 1, L, precision, ci, cm, u0, z0, y0, dt, nsteps, RK4, ns, savesteps,
                us, zs, ys, C_NULL, C_NULL, C_NULL), "smm_pde_solve")
    states = Any[ArrayPartition(us[:, :, :, :, m, 1], zs[:, :, m, 1], ys[:, :, m, 1]) for m in 1:ns]
    return FixedStepSolution(Float64.(savesteps) .* dt, states), (p, q)
end

# ---- PDEsolveensemble (ensemble.jl:58-85): all N instances advance as ONE device-resident batch -------
# Keyword extensions: `qs` = a vector of N parameter sets (the reference runs N identical solves unless init = "random",
# ensemble.jl:70; with `qs` the call is a parameter sweep, instance k integrates with qs[k]); `devices` as above.
function PDEsolveensemble(tmax = 0.1, N = 10; savepoints = 4, alg = nothing, init = "4inf", p = PDEconstruct(),
                          q = parameters(), save = true, dt = nothing, precision = PREC_F32, qs = nothing, devices = :default)
    qs !== nothing && (N = length(qs))
    inits = [constructinitial(init, (p, qs === nothing ? q : qs[k])) for k in 1:N]
    qa = inits[1][2]; L = qa.L
    D = maximum(0.5 * i[2].sigma^2 for i in inits)
    dt === nothing && (dt = min(1e-3, 0.16 * p.dx^2 / D))
    per = max(1, ceil(Int, tmax / (savepoints - 1) / dt - 1e-9)); nsteps = per * (savepoints - 1); dt = tmax / nsteps
    savesteps = Int32.((0:savepoints-1) .* per)                       # LinRange(0, tmax, savepoints) (ensemble.jl:65)
    u0 = cat([i[1].x[1] for i in inits]...; dims = 5); z0 = cat([i[1].x[2] for i in inits]...; dims = 3)
    y0 = cat([i[1].x[3] for i in inits]...; dims = 3)
    us = zeros(p.N_x, p.N_y, 2, L, savepoints, N); zs = zeros(2, 2, savepoints, N); ys = zeros(2, L, savepoints, N)
    ci = hcat([Float64.(collect(i[2].controlled_inf)) for i in inits]...); cm = hcat([Float64.(collect(i[2].controlled_med)) for i in inits]...)
    check(ccall((:smm_pde_solve, LIB), Cint,
                (Ptr{Cvoid}, Ref{SmmGrid}, Ptr{SmmParams}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Int32, Int32, Int32, Ptr{Int32},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx(; devices = devices), grid(p), [params(i[2], p) for i in inits], N, L, precision, ci, cm, u0, z0, y0, dt, nsteps, RK4,
                savepoints, savesteps, us, zs, ys, C_NULL, C_NULL, C_NULL), "smm_pde_solve")
    P = (p, qa)
    if save
        (p, q) = P
        @eval Main using JLD2
        Main.JLD2.jldsave(string("data/pde_ensemble_", init, ".jld2"); us, zs, ys, p, q)          # ensemble.jl:80-83
    end
    return us, zs, ys, P
end

# ---- adaptive integration (smm_pde_solve_adaptive): the keywords of the reference's solve calls
#      (pde.jl:291; controlstrategies.jl:10,96) on the device-side embedded 3(2) pair ----------------------
struct SmmAdapt
    abstol::Float64; reltol::Float64; dt0::Float64; dtmin::Float64; dtmax::Float64
    force_dtmin::Int32; max_attempts::Int32
end

function PDEsolveadaptive(tmax = 0.1; init = "4inf", p = PDEconstruct(), q = parameters(), saveat = LinRange(0, tmax, 101),
                          abstol = 1e-6, reltol = 1e-3, dtmin = 0.0, force_dtmin = false, precision = PREC_F32)
    uzy0, qa = constructinitial(init, (p, q)); L = qa.L
    times = collect(Float64, saveat); ns = length(times)
    us = zeros(p.N_x, p.N_y, 2, L, ns, 1); zs = zeros(2, 2, ns, 1); ys = zeros(2, L, ns, 1)
    nacc = zeros(Int64, 1); nrej = zeros(Int64, 1)
    opt = SmmAdapt(abstol, reltol, 0.0, dtmin, 0.0, force_dtmin ? 1 : 0, 0)
    check(ccall((:smm_pde_solve_adaptive, LIB), Cint,
                (Ptr{Cvoid}, Ref{SmmGrid}, Ptr{SmmParams}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{SmmAdapt}, Int32, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                ctx(), grid(p), [params(qa, p)], 1, L, precision, Float64.(collect(qa.controlled_inf)),
                Float64.(collect(qa.controlled_med)), uzy0.x[1], uzy0.x[2], uzy0.x[3], opt, ns, times,
                us, zs, ys, C_NULL, C_NULL, C_NULL, nacc, nrej), "smm_pde_solve_adaptive")
    states = [ArrayPartition(us[:, :, :, :, m, 1], zs[:, :, m, 1], ys[:, :, m, 1]) for m in 1:ns]
    return FixedStepSolution(times, states), (p, qa)
end

# ---- one large population kept on the GPU between calls (config C5: n = 1e7 agents, CSR follower network, L = 64) -----
# Not part of the reference's API (abm.jl:42 builds a dense n x n IndNet and abm.jl:12-13 asserts L == 4): a handle
# around smm_abm_pop_*.  x0 [n,2], state [n] (+-1), follow0 [n] (0-based), inf0 [L,2], media0 [2,2], network = (rowptr, colidx).
mutable struct ABMPopulation
    h::Ptr{Cvoid}
    n::Int; L::Int
end
function ABMPopulation(q, p, x0, state, follow0, inf0, media0, network; seed = 0, precision = PREC_F32)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:smm_abm_pop_create, LIB), Cint,
                (Ptr{Cvoid}, Ref{SmmParams}, Int32, UInt64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                ctx(), params(q, p), precision, seed, Float64.(x0), Int32.(state), Int32.(follow0), Float64.(inf0), Float64.(media0),
                Int32.(network[1]), Int32.(network[2]), out), "smm_abm_pop_create")
    pop = ABMPopulation(out[], q.n, q.L)
    finalizer(x -> ccall((:smm_abm_pop_destroy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), x.h), pop)
    return pop
end
step!(pop::ABMPopulation, nsteps = 1) = check(ccall((:smm_abm_pop_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int32), ctx(), pop.h, nsteps), "smm_abm_pop_step")
function Base.read(pop::ABMPopulation; positions = true)
    x = positions ? zeros(pop.n, 2) : zeros(0, 2); follow = zeros(Int32, pop.n); inf = zeros(pop.L, 2); media = zeros(2, 2)
    numbers = zeros(Int64, 2, pop.L)
    check(ccall((:smm_abm_pop_read, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                ctx(), pop.h, positions ? x : C_NULL, follow, inf, media, numbers), "smm_abm_pop_read")
    return (x = x, follow = follow .+ 1, inf = inf, media = media, numbers = numbers)
end

end # module
