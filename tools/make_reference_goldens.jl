# Pins the repo's oracle fixtures against the REAL reference.  Not runnable in the build container (no Julia
# there): for a maintainer with Julia >= 1.10 and a checkout of matthieugomez/EconPDEs.jl v3.
#
#   python tools/export_golden_inputs.py                      # tests/golden/*.npz -> tests/golden/reference_io/<case>/*.csv
#   julia --project=/path/to/EconPDEs.jl tools/make_reference_goldens.jl /path/to/EconPDEs.jl
#   python tools/export_golden_inputs.py --compare            # reference outputs vs the committed oracle fixtures
#
# For every exported case it reads the EXACT grids and guess the oracle fixture was made with (17 significant
# digits; rebuilding them from quantiles would differ in the last bits between Distributions.jl and scipy), solves
# with stock `pdesolve` (default keywords unless the case's options.csv says otherwise), and writes next to the inputs
#   ref_solution_<unknown>.csv   one value per line, column-major
#   ref_log.txt                  the table pdesolve prints (iteration, time step, residual / rejection marks)
#   ref_summary.csv              residual_norm, converged, wall seconds
# The model bodies come from the reference's own example files (their top-level demo solves run once at include
# time; they take a few seconds); the synthetic 3-state HJB of config C5 is defined below as in SURVEY.md 8(d).
using EconPDEs, DelimitedFiles

refdir = length(ARGS) >= 1 ? ARGS[1] : error("usage: julia make_reference_goldens.jl /path/to/EconPDEs.jl [io-dir]")
iodir = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden", "reference_io")

# each example file defines its model struct and makes it callable as pde(state, u)
module RefExamples
    using EconPDEs
    ex(refdir, rel) = Base.include(@__MODULE__, joinpath(refdir, "examples", rel))
end
for rel in ("asset_pricing/campbell_cochrane.jl", "asset_pricing/bansal_yaron.jl",
            "consumption_saving/consumption_saving_diffusion_income.jl", "neoclassical_growth.jl")
    try
        RefExamples.ex(refdir, rel)
    catch err   # Plots or another docs-only dependency missing: the struct and functor are defined before that point
        @warn "include of $rel stopped early" err
    end
end

# config C5 (SURVEY.md 8(d)): vt = -(x+y+z + sum mu_d v_d(upwind) + 1/2 sum sigma_d^2 v_dd + corr sx sy v_xy_up - rho v - theta v^2)
Base.@kwdef struct Hjb3dSynth
    kx::Float64 = 0.2; ky::Float64 = 0.3; kz::Float64 = 0.4
    sx::Float64 = 0.1; sy::Float64 = 0.1; sz::Float64 = 0.1
    corr::Float64 = 0.5; rho::Float64 = 0.05; theta::Float64 = 1e-3
end
function (m::Hjb3dSynth)(state::NamedTuple, u::NamedTuple)
    (; x, y, z) = state
    mx, my, mz = -m.kx * (x - 0.5), -m.ky * (y - 0.5), -m.kz * (z - 0.5)
    vx = mx >= 0 ? u.vx_up : u.vx_down
    vy = my >= 0 ? u.vy_up : u.vy_down
    vz = mz >= 0 ? u.vz_up : u.vz_down
    drift = mx * vx + my * vy + mz * vz
    diff = 0.5 * m.sx^2 * u.vxx + 0.5 * m.sy^2 * u.vyy + 0.5 * m.sz^2 * u.vzz
    cross = m.corr * m.sx * m.sy * u.vxy_up
    (; vt = -((x + y + z) + drift + diff + cross - m.rho * u.v - m.theta * (u.v * u.v)))
end

# fixture prefix -> (constructor taking the parameter vector of params.csv, state names, unknown names)
models = Dict(
    "c1_readme" => (p -> ((state, u) -> (μ = -p[2] * (state.x - 0.5); (; vt = -(state.x + μ * (μ >= 0 ? u.vx_up : u.vx_down) - p[1] * u.v)))), (:x,), (:v,)),
    "c4_" => (p -> RefExamples.CampbellCochraneModel(μ = p[1], σ = p[2], γ = p[3], ρ = p[4], κs = p[5], b = p[6]), (:s,), (:p,)),
    "c3_bansal_yaron" => (p -> RefExamples.BansalYaronModel(μbar = p[1], vbar = p[2], κμ = p[3], νμ = p[4], κv = p[5], νv = p[6], ρ = p[7], γ = p[8], ψ = p[9]), (:μ, :v), (:p,)),
    "c2_ahlm" => (p -> RefExamples.AchdouHanLasryLionsMollModel_Diffusion(κy = p[1], ybar = p[2], σy = p[3], r = p[4], ρ = p[5], γ = p[6], amin = p[7], amax = p[8]), (:y, :a), (:v,)),
    "c5_hjb3d" => (p -> Hjb3dSynth(p...), (:x, :y, :z), (:v,)),
)

for case in sort(readdir(iodir))
    dir = joinpath(iodir, case)
    isdir(dir) || continue
    key = findfirst(k -> startswith(case, k), collect(keys(models)))
    key === nothing && (@info "skipping $case (no reference model wired)"; continue)
    make, snames, unames = models[collect(keys(models))[key]]
    params = vec(readdlm(joinpath(dir, "params.csv"), ',', Float64))
    grid = NamedTuple{snames}(Tuple(vec(readdlm(joinpath(dir, "grid_$(i).csv"), ',', Float64)) for i in 1:length(snames)))
    S = Tuple(length(g) for g in grid)
    guess = NamedTuple{unames}(Tuple(reshape(vec(readdlm(joinpath(dir, "guess_$(i).csv"), ',', Float64)), S) for i in 1:length(unames)))
    pde = make(params)
    logfile = joinpath(dir, "ref_log.txt")
    t0 = time()
    result = open(logfile, "w") do io
        redirect_stdout(io) do
            pdesolve(pde, grid, guess)          # stock defaults: Δ = 1, Newton, FiniteDiff colored Jacobian, UMFPACK
        end
    end
    wall = time() - t0
    for (i, k) in enumerate(unames)
        writedlm(joinpath(dir, "ref_solution_$(i).csv"), vec(result.solution[k]), ',')
    end
    writedlm(joinpath(dir, "ref_summary.csv"), Any[result.residual_norm Int(result.converged) wall], ',')
    @info "$case: residual $(result.residual_norm) converged $(result.converged) in $(round(wall, digits = 2)) s"
end
