#=
julia_fixtures.jl -- pins the CPU oracle (oracle/*.py, oracle/abm_oracle.c) to the REFERENCE ITSELF.

This image has no Julia, so this script cannot run here; it is written to be run ONCE by anybody who has the
reference's environment (Julia 1.8.5, `julia --project=<reference checkout>`):

    julia --project=/path/to/SocialMediaModel baseline/julia_fixtures.jl /path/to/SocialMediaModel tests/golden/ref

It writes plain .npy files + manifest.json into tests/golden/ref/.  `python -m pytest tests/test_golden_cpu.py` then
compares the oracle against them (and the -m gpu suite compares the CUDA path); without the directory those tests
report "julia absent: parity unpinned".

How the reference is driven with known random draws WITHOUT editing it: the reference's own source files
(src/setting.jl, src/abm.jl, src/ensemble.jl, src/pde.jl) are include()d into a fresh module in which `rand` and
`randn` are module-local functions defined BEFORE the include.  Every draw the reference makes (abm.jl:16,39,122,126,
173,181-196) therefore comes from a seeded generator of this script and is logged in consumption order.  The normal
draws have a fixed layout (randn(n,2) per step, then L + M times randn(2,1)); the uniform draws of changeinfluencer
(one per agent, a second one only when the agent switches) are replayed on the Python side in the same order.
Nothing of the reference is copied or modified.
=#
using Random, LinearAlgebra

const REF = length(ARGS) >= 1 ? ARGS[1] : error("usage: julia_fixtures.jl <reference checkout> <output dir>")
const OUT = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden", "ref")
mkpath(OUT)

# ---- minimal .npy writer (format 1.0, little endian, Fortran order = Julia's memory order) ---------------------
npy_descr(::Type{Float64}) = "<f8"
npy_descr(::Type{Int64}) = "<i8"
function write_npy(name::String, A::AbstractArray{T}) where {T<:Union{Float64,Int64}}
    A = Array(A)
    shape = join(string.(size(A)), ", ") * (ndims(A) == 1 ? "," : "")
    hdr = "{'descr': '$(npy_descr(T))', 'fortran_order': True, 'shape': ($shape), }"
    pad = 64 - ((10 + length(hdr) + 1) % 64)
    hdr = hdr * (" "^(pad % 64)) * "\n"
    open(joinpath(OUT, name * ".npy"), "w") do io
        write(io, UInt8[0x93, 0x4e, 0x55, 0x4d, 0x50, 0x59, 0x01, 0x00])
        write(io, UInt16(length(hdr)))
        write(io, hdr)
        write(io, A)
    end
end
write_npy(name::String, A::AbstractArray{<:Integer}) = write_npy(name, Int64.(A))
write_npy(name::String, A::AbstractArray{<:Real}) = write_npy(name, Float64.(A))

# ---- the reference with logged draws ------------------------------------------------------------------------------
module RefLogged
    import Random
    const GEN = Random.MersenneTwister(20231020)
    const UNI = Float64[]      # every uniform draw, in consumption order
    const NRM = Float64[]      # every normal draw, in consumption order (column-major within one randn call)
    reset!(seed) = (Random.seed!(GEN, seed); empty!(UNI); empty!(NRM); nothing)
    # module-local replacements: the reference's calls rand(), rand(n), rand(n,2), randn(n,2), randn(2,1) resolve here
    function rand(dims::Integer...)
        A = Random.rand(GEN, Float64, dims...)
        append!(UNI, vec(A))
        return A
    end
    function rand()
        v = Random.rand(GEN, Float64)
        push!(UNI, v)
        return v
    end
    function randn(dims::Integer...)
        A = Random.randn(GEN, Float64, dims...)
        append!(NRM, vec(A))
        return A
    end
    const REF = Main.REF
    include(joinpath(REF, "src", "setting.jl"))
    include(joinpath(REF, "src", "abm.jl"))
    include(joinpath(REF, "src", "ensemble.jl"))
end

manifest = Dict{String,Any}("julia" => string(VERSION), "reference" => REF)

# (1) ABMsolve: full history for three seeds, default parameters, NT = 60 ------------------------------------------
for (tag, NT, seed) in (("abm_a", 60, 1), ("abm_b", 60, 2), ("abm_c", 200, 3))
    RefLogged.reset!(seed)
    xs, stateinfs, infs, meds, state, (p, q) = RefLogged.ABMsolve(NT; chosenseed = seed)
    n, L = q.n, q.L
    write_npy(tag * "_xs", cat(xs...; dims = 3))                  # [n,2,NT+1]
    write_npy(tag * "_stateinfs", hcat(stateinfs...))             # [n,NT+1], values 1.0..L (abm.jl:164,212)
    write_npy(tag * "_infs", cat(infs...; dims = 3))              # [L,2,NT+1]
    write_npy(tag * "_meds", cat(meds...; dims = 3))              # [M,2,NT+1]
    write_npy(tag * "_state", state)
    write_npy(tag * "_uniform_log", copy(RefLogged.UNI))          # rand(n,2) (2n), rand(n) (n), then per step: u1 (and u2 on a switch) per agent
    write_npy(tag * "_normal_log", copy(RefLogged.NRM))           # per step: randn(n,2) (2n), L x randn(2,1), M x randn(2,1)
    manifest[tag] = Dict("NT" => NT, "n" => n, "L" => L, "M" => q.M, "dt" => p.dt, "seed" => seed)
end

# (2) single calls of the force routines on a fixed state (no draws involved) ---------------------------------------
let
    RefLogged.reset!(11)
    p, q = RefLogged.ABMconstruct(), RefLogged.parameters()
    x, media, inf, FolInfNet, state, IndNet = RefLogged.ABMinit((p, q))
    write_npy("force_x", x); write_npy("force_state", state); write_npy("force_inf", inf); write_npy("force_media", media)
    write_npy("force_FolInfNet", FolInfNet)
    write_npy("force_attraction", RefLogged.attraction(x, IndNet))
    write_npy("force_influence", RefLogged.influence(x, media, inf, FolInfNet, state, (p, q)))
end

# (3) ABMsolveensemble: KDE snapshots of 3 realizations on a coarse KDE grid -------------------------------------------
let
    RefLogged.reset!(5)
    p = RefLogged.ABMconstruct(dx = 0.2)
    q = RefLogged.parameters()
    us, zs, ys, _ = RefLogged.ABMsolveensemble(40, 3; savepoints = 3, p = p, q = q, sigma = 0.1, save = false)
    write_npy("ens_us", us); write_npy("ens_zs", zs); write_npy("ens_ys", ys)
    write_npy("ens_uniform_log", copy(RefLogged.UNI)); write_npy("ens_normal_log", copy(RefLogged.NRM))
    manifest["ens"] = Dict("NT" => 40, "N" => 3, "savepoints" => 3, "dx" => 0.2, "sigma" => 0.1, "n" => q.n, "L" => q.L)
end

# (4) PDE right-hand side f at uniforminit and at a perturbed state (coarse and default grid).  pde.jl needs the
#     reference's heavier dependencies (DifferentialEquations, Distances, ...): load the package itself for this part.
try
    push!(LOAD_PATH, REF)
    @eval using SocialMediaModel
    SM = SocialMediaModel
    for (tag, p) in (("pde_coarse", SM.PDEconstructcoarse()), ("pde_default", SM.PDEconstruct()))
        q = SM.parameters()
        uzy0, q2 = SM.constructinitial("4inf", (p, q))
        rng = MersenneTwister(7)
        for (sub, scale) in (("init", 0.0), ("pert", 0.3))
            uzy = deepcopy(uzy0)
            uzy.x[1] .*= (1 .+ scale .* rand(rng, size(uzy.x[1])...))
            uzy.x[3] .+= scale .* 0.2 .* randn(rng, size(uzy.x[3])...)
            duzy = zero(uzy)
            SM.f(duzy, uzy, (p, q2), 0.0)
            for (k, nm) in enumerate(("u", "z", "y"))
                write_npy("$(tag)_$(sub)_$(nm)", uzy.x[k]); write_npy("$(tag)_$(sub)_d$(nm)", duzy.x[k])
            end
        end
        manifest[tag] = Dict("N_x" => p.N_x, "dx" => p.dx)
    end
    # the reference's own time integration, for context (adaptive, default tolerances): state at t = 0.1
    sol, (p, q2) = SM.PDEsolve(0.1; p = SM.PDEconstructcoarse())
    u, z, y = SM.sol2uyz(sol, 0.1)
    write_npy("pde_coarse_sol01_u", u); write_npy("pde_coarse_sol01_z", z); write_npy("pde_coarse_sol01_y", y)
    manifest["pde_solve"] = Dict("tmax" => 0.1, "steps" => length(sol.t))
catch err
    @warn "PDE fixtures skipped (the SocialMediaModel package did not load)" err
    manifest["pde_error"] = string(err)
end

open(joinpath(OUT, "manifest.json"), "w") do io
    # tiny JSON writer (no JSON.jl dependency)
    tojson(x::AbstractString) = "\"" * escape_string(x) * "\""
    tojson(x::Number) = string(x)
    tojson(d::Dict) = "{" * join([tojson(string(k)) * ": " * tojson(v) for (k, v) in d], ", ") * "}"
    write(io, tojson(manifest))
end
println("wrote reference fixtures to ", OUT)
