#=
julia_bench.jl -- times the REFERENCE's own CPU path for the same quantities bench.py reports (SURVEY 8d, tier 1).
Cannot run in this image (no Julia); bench.py runs it automatically when a `julia` binary and the reference checkout
are present (env SMM_REFERENCE_DIR), otherwise the C restatement of the reference's loops is the CPU baseline.

    julia -t auto --project=/path/to/SocialMediaModel baseline/julia_bench.jl /path/to/SocialMediaModel [N_realizations]

Prints ONE JSON line: agent-steps/s of ABMsolve and ABMsolveensemble, seconds per evaluation of f, cell-steps/s of
PDEsolve.  BLAS threads as in the reference's dev.jl:7.
=#
using LinearAlgebra
const REF = length(ARGS) >= 1 ? ARGS[1] : error("usage: julia_bench.jl <reference checkout> [N]")
const NREAL = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 16
push!(LOAD_PATH, REF)
using SocialMediaModel
const SM = SocialMediaModel
BLAS.set_num_threads(haskey(ENV, "SLURM_JOB_CPUS_PER_NODE") ? parse(Int, ENV["SLURM_JOB_CPUS_PER_NODE"]) : Sys.CPU_THREADS)  # dev.jl:7

SM.ABMsolve(5)                                              # compile
t_abm = @elapsed SM.ABMsolve(100)                           # C1: n = 250, NT = 100
SM.ABMsolveensemble(5, 2; save = false)
t_ens = @elapsed SM.ABMsolveensemble(200, NREAL; save = false)   # C2 sample: NT = 200, KDE at 4 save points
p = SM.PDEconstruct(); q = SM.parameters()
uzy0, q2 = SM.constructinitial("4inf", (p, q))
duzy = zero(uzy0)
SM.f(duzy, uzy0, (p, q2), 0.0)
t_f = @elapsed for _ in 1:5 SM.f(duzy, uzy0, (p, q2), 0.0) end
SM.PDEsolve(0.01)
t_pde = @elapsed (sol = SM.PDEsolve(0.1)[1])
cells = p.N_x * p.N_y * 2 * 4
println("{\"impl\": \"reference-julia\", \"threads\": $(Threads.nthreads()), \"blas_threads\": $(BLAS.get_num_threads()), ",
        "\"abm_C1_agent_steps_per_s\": $(250 * 100 / t_abm), \"abm_C2_agent_steps_per_s\": $(NREAL * 200 * 250 / t_ens), ",
        "\"abm_C2_realizations\": $NREAL, \"pde_f_seconds_per_eval\": $(t_f / 5), ",
        "\"pde_solve_0p1_seconds\": $t_pde, \"pde_solver_steps\": $(length(sol.t)), \"pde_cells\": $cells}")
