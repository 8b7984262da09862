# EconPDEsCUDA.jl -- the binding a maintainer of EconPDEs.jl would add to route registered models
# to libeconpdes_cuda.so.  NOT exercised in this repository's CI: the build image has no Julia
# (see DESIGN.md section 2); the same C ABI is driven from Python ctypes in econpdes.jl_b200/.
#
# Usage:   using EconPDEs, EconPDEsCUDA
#          result = pdesolve(CUDAModel(BansalYaronModel()), stategrid, guess)   # -> EconPDEResult
module EconPDEsCUDA

using EconPDEs
using EconPDEs: EconPDEResult, StateGrid, _asnamedtuple, _check_generated_names, _fill_is_algebraic,
                _bc_array, _bound_array, _delta_array

const lib = get(ENV, "ECONPDES_CUDA_LIB", "libeconpdes_cuda")

# the struct mirrors below were written against this ECONPDES_ABI_VERSION (4: ECONPDES_MAX_DIMS = 4, so
# `size` / `grid` of CProblem are 4-tuples); a library of another version would mis-read them
const ABI_VERSION = 4
function __init__()
    v = ccall((:econpdes_abi_version, lib), Cint, ())
    v == ABI_VERSION || error("libeconpdes_cuda has ABI version $v, this binding was written against $ABI_VERSION")
end

# mirrors `econpdes_problem`, `econpdes_options`, `econpdes_log_row`, `econpdes_result`
# (include/econpdes_cuda.h); field order and types must match exactly.
struct CProblem
    model_id::Int32; ndims::Int32; nfuncs::Int32; nparams::Int32
    size::NTuple{4, Int64}
    grid::NTuple{4, Ptr{Float64}}
    params::Ptr{Float64}; guess::Ptr{Float64}; bc::Ptr{Float64}; is_algebraic::Ptr{UInt8}
    lower_bound::Ptr{Float64}; upper_bound::Ptr{Float64}; delta_vec::Ptr{Float64}
end
mutable struct COptions
    abstol::Float64; inner_abstol::Float64; delta::Float64; scale::Float64; min_delta::Float64
    max_delta::Float64; max_residual_growth::Float64
    maxiters::Int32; inner_maxiters::Int32; eps_policy::Int32; linear_solver::Int32
    lin_rtol::Float64; lin_maxiters::Int32; lin_restart::Int32
    check_monotonicity::Int32; monotonicity_tol::Float64
    COptions() = new()
end
struct CLogRow
    iter::Int32; status::Int32; newton_iters::Int32; linear_iters::Int32; delta::Float64; residual::Float64
end
mutable struct CResult
    solution::Ptr{Float64}; saved::Ptr{Float64}; log::Ptr{CLogRow}
    log_capacity::Int32; log_len::Int32
    residual_norm::Float64
    converged::Int32; exit_reason::Int32; n_outer::Int32; n_newton::Int32
    n_linear_iters::Int64; kernel_launches::Int64; device_ms::Float64
    sign_flags::Int32; nonmonotone_entries::Int32; nonmonotone_max::Float64   # src/monotonicity.jl diagnostics
    n_linear_solves::Int32; n_linear_loose::Int32; lin_max_relres::Float64    # quality of the Krylov solves (ABI 3)
    CResult() = new()
end

"Wraps one of the example model structs; `name` selects the registered device functor."
struct CUDAModel{M}
    model::M
    name::String
    params::Vector{Float64}
    saved::Tuple{Vararg{Symbol}}
end
# e.g. CUDAModel(m::BansalYaronModel) = CUDAModel(m, "bansal_yaron",
#          Float64[m.μbar, m.vbar, m.κμ, m.νμ, m.κv, m.νv, m.ρ, m.γ, m.ψ], ())
#      CUDAModel(m::DiTellaModel) = CUDAModel(m, "di_tella", Float64[m.γ, m.ψ, m.ρ, m.τ, m.A, m.δ, m.B, m.σ, m.ϕ, m.νbar, m.κν, m.σνbar],
#          (:σx, :Q, :r, :κ, :σp, :growth, :investment))
#      CUDAModel(m::GarleanuPanageasModel) = CUDAModel(m, "garleanu_panageas",
#          Float64[m.γA, m.ψA, m.γB, m.ψB, m.ρ, m.δ, m.νA, m.μ, m.σ, m.B1, m.δ1, m.B2, m.δ2, m.ω], (:r, :κ))
#      CUDAModel(m::AchdouHanLasryLionsMollModel_TwoStates) = CUDAModel(m, "ahlm_two_states",
#          Float64[m.yl, m.yh, m.λlh, m.λhl, m.r, m.ρ, m.γ, m.amin, m.amax], (:cl, :ch, :μla, :μha))
#      CUDAModel(m::TuckmanVilaModel) = CUDAModel(m, "tuckman_vila", Float64[m.c, m.r, m.ρ, m.σ, m.a, m.T], ())
#          (a three-argument body pde(state, u, τ): only the method with a time grid applies, as in the reference)
#      CUDAModel(m::WachterModel) = CUDAModel(m, "wachter", Float64[m.μ, m.σ, m.λbar, m.κλ, m.νλ, mean(m.ZDistribution),
#          std(m.ZDistribution), m.ρ, m.γ, m.ψ, m.ϕ, mgf(m.ZDistribution, 1), mgf(m.ZDistribution, -m.γ),
#          mgf(m.ZDistribution, 1 - m.γ)], ())      (the three mgf values ride along as parameters: any ZDistribution works)
#      CUDAModel(m::GomezModel) = CUDAModel(m, "gomez", Float64[m.γ, m.ψ, m.ρ, m.ρE, m.αE, m.ν, m.η, m.δ, m.ϕ, m.πE, m.g, m.σ,
#          m.λ, m.τ], (:x, :μx, :σx, :κ, :r, :σWH, :μWH, :σp, :p, :px, :σR, :μR, :μRλ, :σRλ, :μErelative, :σErelative,
#          :μHrelative, :σHrelative, :μCE, :μxW, :σxW, :pH, :μCH, :wedge, :xW, :σWE, :σCH, :αE, :pE, :μp))
#      likewise "he_krishnamurthy", "haddad", "bolton_chen_wang", "ahlm_risky_asset", "leland", "wang_wang_yang":
#      parameters in the field order of the example's struct, saved names in the order of the body's second NamedTuple.

last_error() = unsafe_string(ccall((:econpdes_last_error, lib), Cstring, ()))

function EconPDEs.pdesolve(m::CUDAModel, grid, guess; is_algebraic = nothing, bc = nothing,
                           lower_bound = nothing, upper_bound = nothing, abstol = sqrt(eps()),
                           verbose = true, maxiters = 100, Δ = 1.0, kwargs...)
    # identical normalisation to src/pdesolve.jl:97-123
    grid = _asnamedtuple(grid); guess = _asnamedtuple(guess)
    _check_generated_names(collect(keys(grid)), collect(keys(guess)))
    stategrid = StateGrid(grid); S = size(stategrid); F = length(guess)
    guess_array = cat(values(guess)...; dims = length(S) + 1)
    alg = UInt8.(cat(values(_fill_is_algebraic(is_algebraic, guess, S))...; dims = length(S) + 1))
    bcs = _bc_array(bc, guess, grid)
    lb = _bound_array(lower_bound, guess, -Inf, :lower_bound)
    ub = _bound_array(upper_bound, guess, Inf, :upper_bound)
    Δvec = Δ isa Union{NamedTuple, AbstractDict} ? vec(_delta_array(Δ, guess, S)) : nothing
    grids = [collect(Float64, g) for g in values(grid)]
    opts = COptions(); ccall((:econpdes_default_options, lib), Cvoid, (Ref{COptions},), opts)
    opts.abstol = abstol; opts.maxiters = maxiters
    Δvec === nothing && (opts.delta = Δ)
    for (k, v) in kwargs   # scale, minΔ, maxΔ, max_residual_growth, inner_maxiters, inner_abstol
        k === :minΔ && (opts.min_delta = v; continue)
        k === :maxΔ && (opts.max_delta = v; continue)
        hasfield(COptions, k) || throw(ArgumentError("unsupported keyword `$k` for the CUDA path"))
        setfield!(opts, k, convert(fieldtype(COptions, k), v))
    end
    mid = ccall((:econpdes_model_id, lib), Cint, (Cstring,), m.name)
    mid < 0 && throw(ArgumentError(last_error()))
    sol = similar(guess_array); saved = Array{Float64}(undef, S..., length(m.saved)); log = Vector{CLogRow}(undef, maxiters)
    res = CResult(); res.solution = pointer(sol); res.saved = isempty(m.saved) ? C_NULL : pointer(saved)
    res.log = pointer(log); res.log_capacity = maxiters
    GC.@preserve grids guess_array alg bcs lb ub Δvec sol saved log m begin
        gp = ntuple(d -> d <= length(S) ? pointer(grids[d]) : Ptr{Float64}(C_NULL), 4)
        prob = CProblem(mid, length(S), F, length(m.params), ntuple(d -> d <= length(S) ? S[d] : 1, 4), gp,
                        pointer(m.params), pointer(guess_array), pointer(bcs), pointer(alg), pointer(lb), pointer(ub),
                        Δvec === nothing ? C_NULL : pointer(Δvec))
        rc = ccall((:econpdes_solve_stationary, lib), Cint, (Ref{CProblem}, Ref{COptions}, Ref{CResult}), prob, opts, res)
    end
    rc < 0 && throw(ArgumentError(last_error()))    # e.g. "G! returns NaN with the initial value"
    ynames = tuple(keys(guess)...)
    solution = NamedTuple{ynames}(ntuple(k -> collect(selectdim(sol, ndims(sol), k)), F))
    savednt = isempty(m.saved) ? NamedTuple() :
        merge(solution, NamedTuple{m.saved}(ntuple(k -> collect(selectdim(saved, ndims(saved), k)), length(m.saved))))
    verbose && for r in log[1:min(res.log_len, maxiters)]   # same table as src/finiteschemesolve.jl:421, 433-435, 441-445
        if r.status == 0
            Printf.@printf("%4d %8.2e %8.2e\n", r.iter, r.delta, r.residual)
        elseif r.status == 1      # growth rejection: the engine stores residual/old_residual in `residual`
            growth_text = isfinite(r.residual) ? Printf.@sprintf("%.1f×", r.residual) : "∞"
            Printf.@printf("%4d %8.2e        ✗ (residual grew %s > max_residual_growth)\n", r.iter, r.delta, growth_text)
        else
            reason = r.status == 3 ? " (model returned NaN)" : r.status == 4 ? " (singular Jacobian)" : ""
            Printf.@printf("%4d %8.2e        ✗%s\n", r.iter, r.delta, reason)
        end
    end
    return EconPDEResult(solution, res.residual_norm, savednt, abstol)
end

# Time-dependent problems: pdesolve(m, grid, guess, τs) -> _solve_backward (src/pdesolve.jl:158-224).
# The guess is the terminal value; every unknown gains a trailing time dimension.
function EconPDEs.pdesolve(m::CUDAModel, grid, guess, τs::AbstractVector; bc = nothing, is_algebraic = nothing,
                           lower_bound = nothing, upper_bound = nothing, abstol = sqrt(eps()),
                           maxiters = 100, inner_maxiters = maxiters, inner_abstol = abstol, kwargs...)
    EconPDEs._reject_time_dependent_stationary_keywords(kwargs); EconPDEs._check_time_grid(τs)
    # every keyword the device path does not implement is an error, not a silently different solve
    isempty(kwargs) || throw(ArgumentError("unsupported keyword(s) for the CUDA time-dependent path: $(collect(keys(kwargs)))"))
    grid = _asnamedtuple(grid); guess = _asnamedtuple(guess)
    stategrid = StateGrid(grid); S = size(stategrid); F = length(guess); nt = length(τs)
    guess_array = cat(values(guess)...; dims = length(S) + 1)
    bcs = _bc_array(bc, guess, grid); grids = [collect(Float64, g) for g in values(grid)]; ts = collect(Float64, τs)
    opts = COptions(); ccall((:econpdes_default_options, lib), Cvoid, (Ref{COptions},), opts)
    opts.abstol = abstol; opts.maxiters = maxiters; opts.inner_maxiters = inner_maxiters; opts.inner_abstol = inner_abstol
    mid = ccall((:econpdes_model_id, lib), Cint, (Cstring,), m.name)
    sol = Array{Float64}(undef, S..., F, nt); saved = Array{Float64}(undef, S..., length(m.saved), nt)
    rn = zeros(nt); nit = zeros(Int32, nt); ms = Ref(0.0); nl = Ref(Int64(0))
    # is_algebraic / bounds normalised like the stationary method (src/pdesolve.jl:112-115): _solve_backward honours them
    alg = UInt8.(cat(values(_fill_is_algebraic(is_algebraic, guess, S))...; dims = length(S) + 1))
    lb = _bound_array(lower_bound, guess, -Inf, :lower_bound)
    ub = _bound_array(upper_bound, guess, Inf, :upper_bound)
    GC.@preserve grids guess_array bcs alg lb ub sol saved rn nit ts m begin
        gp = ntuple(d -> d <= length(S) ? pointer(grids[d]) : Ptr{Float64}(C_NULL), 4)
        prob = CProblem(mid, length(S), F, length(m.params), ntuple(d -> d <= length(S) ? S[d] : 1, 4), gp,
                        pointer(m.params), pointer(guess_array), pointer(bcs), pointer(alg), pointer(lb), pointer(ub), C_NULL)
        rc = ccall((:econpdes_solve_backward, lib), Cint,
                   (Ref{CProblem}, Ref{COptions}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Int32}, Ref{Float64}, Ref{Int64}),
                   prob, opts, ts, nt, sol, isempty(m.saved) ? C_NULL : pointer(saved), rn, nit, ms, nl)
    end
    rc < 0 && throw(ArgumentError(last_error()))
    ynames = tuple(keys(guess)...)
    solution = NamedTuple{ynames}(ntuple(k -> collect(selectdim(sol, length(S) + 1, k)), F))
    savednt = isempty(m.saved) ? NamedTuple() :
        merge(solution, NamedTuple{m.saved}(ntuple(k -> collect(selectdim(saved, length(S) + 1, k)), length(m.saved))))
    return EconPDEResult(solution, rn, savednt, abstol)
end

# A loop of `pdesolve` calls over parameter draws (calibration / estimation sweeps of a 1-D model) in one
# call: params is nparams x B, grids and guess are S x B (one column per problem, column-major = the
# [B][S] row-major layout of the C ABI).  The library pipelines slices of the batch over three
# copy/compute streams; returns (solution S x B, residual_norm, exit_reason, n_outer, n_newton).
function pdesolve_batched(name::String, params::Matrix{Float64}, grids::Matrix{Float64}, guess::Matrix{Float64};
                          abstol = sqrt(eps()), maxiters = 100, Δ = 1.0)
    S, B = size(grids); size(guess) == (S, B) || throw(ArgumentError("guess must be S x B like grids"))
    opts = COptions(); ccall((:econpdes_default_options, lib), Cvoid, (Ref{COptions},), opts)
    opts.abstol = abstol; opts.maxiters = maxiters; opts.delta = Δ
    mid = ccall((:econpdes_model_id, lib), Cint, (Cstring,), name)
    sol = Matrix{Float64}(undef, S, B); rn = Vector{Float64}(undef, B)
    er = Vector{Int32}(undef, B); no = Vector{Int32}(undef, B); nn = Vector{Int32}(undef, B)
    rc = ccall((:econpdes_solve_batched, lib), Cint,
               (Cint, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{COptions}, Ptr{Float64}, Ptr{Float64},
                Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
               mid, B, S, params, grids, guess, opts, sol, rn, er, no, nn)
    rc < 0 && throw(ArgumentError(last_error()))
    return sol, rn, er, no, nn
end

# The same sweep with COMPACT inputs (econpdes_batch_inputs): the first `n_range` points of problem b are
# range(range_lo[b], points[1, b], length = n_range + 1)[1:n_range], built on the device exactly as `range` would,
# followed by points[:, b]; every unknown starts at `guess_fill`.  For the Campbell-Cochrane sweep
# (initialize_stategrid, examples/asset_pricing/campbell_cochrane.jl:40-48): range_lo = -300, points = shigh[2:end],
# n_range = sn - div(sn, 10) - 1, guess_fill = 1.
struct CBatchInputs
    params::Ptr{Float64}; grids::Ptr{Float64}; n_range::Int32; range_lo::Ptr{Float64}; points::Ptr{Float64}
    guess::Ptr{Float64}; guess_fill::Float64
end
function pdesolve_batched(name::String, params::Matrix{Float64}, range_lo::Vector{Float64}, points::Matrix{Float64},
                          n_range::Integer, guess_fill::Float64; abstol = sqrt(eps()), maxiters = 100, Δ = 1.0)
    np, B = size(points); S = n_range + np
    opts = COptions(); ccall((:econpdes_default_options, lib), Cvoid, (Ref{COptions},), opts)
    opts.abstol = abstol; opts.maxiters = maxiters; opts.delta = Δ
    mid = ccall((:econpdes_model_id, lib), Cint, (Cstring,), name)
    sol = Matrix{Float64}(undef, S, B); rn = Vector{Float64}(undef, B)
    er = Vector{Int32}(undef, B); no = Vector{Int32}(undef, B); nn = Vector{Int32}(undef, B)
    GC.@preserve params range_lo points begin
        inp = CBatchInputs(pointer(params), C_NULL, n_range, pointer(range_lo), pointer(points), C_NULL, guess_fill)
        rc = ccall((:econpdes_solve_batched_inputs, lib), Cint,
                   (Cint, Int64, Int64, Ref{CBatchInputs}, Ref{COptions}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32},
                    Ptr{Int32}, Ptr{Int32}), mid, B, S, inp, opts, sol, rn, er, no, nn)
    end
    rc < 0 && throw(ArgumentError(last_error()))
    return sol, rn, er, no, nn
end

end # module
