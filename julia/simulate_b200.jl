# simulate_b200.jl -- `simulate` with the period loop on the GPU (drop-in for src/simulate.jl:68-143 and the
# `_solve_ensemble` hook of ext/ClosedLoopReachabilityOrdinaryDiffEqExt.jl:22-45).
#
# NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image).  The array-backed solution type below is the Julia twin of
# closedloopreachability.jl_b200/simulate.py (executed: tests/test_gpu_rollout.py::test_simulate_mirror).
#
# The reference computes, per control period, the controls of all trajectories in a serial loop (src/simulate.jl:101-106),
# then one `EnsembleProblem` solve (Tsit5, EnsembleThreads), then projects the end states (src/simulate.jl:132-136).
# Here ONE `clr_rollout` call runs all periods of all trajectories (controller evaluation + Tsit5 with OrdinaryDiffEq's
# controller, or RK4); sampling stays in Julia, in the reference's RNG order (x0 once, then W₀ once per period,
# src/simulate.jl:89, :111), so a seeded run draws the same numbers.
#
# Usage:  include("ClrB200.jl"); include("simulate_b200.jl")
#         sim = simulate_b200(prob, ClrB200.CLR_PLANT_UNICYCLE; T=10.0, trajectories=10^6, ctx=ClrB200.Context(0))
using ClosedLoopReachability
using ClosedLoopReachability: control_preprocessing, control_postprocessing, states, disturbances, controls, period,
                              initial_state, _get_tspan
const RA_ = ClosedLoopReachability.ReachabilityAnalysis
using LazySets

const CLR_PLANT_UNICYCLE = 1; const CLR_PLANT_TORA = 2; const CLR_PLANT_ACC = 3; const CLR_PLANT_INVPEND = 4

"""
Array-backed counterpart of `EnsembleSimulationSolution` (src/simulate.jl:17-47): the same accessors over four dense
arrays instead of trajectories x periods small vectors.  `sol[j]`, `trajectory`, `controls`, `disturbances`, `length`
behave like the reference's; a trajectory piece is a lazy `(t, u)` view of the step log (what `plot_simulation!` reads,
ext/ClosedLoopReachabilityOrdinaryDiffEqPlotsExt.jl:57-76).
"""
struct EnsembleSimulationArrays
    x0::Matrix{Float64}        # n_state x N
    Xend::Array{Float64,3}     # n_state x N x iters   end state of every period
    U::Array{Float64,3}        # n_ctrl x N x iters    control of every period
    W::Union{Nothing,Array{Float64,3}}   # n_dist x N x iters
    log_n::Union{Nothing,Matrix{Int32}}  # accepted steps per (trajectory, period) when the step log was requested
    log_t::Union{Nothing,Array{Float64,3}}
    log_u::Union{Nothing,Array{Float64,4}}
end
Base.length(s::EnsembleSimulationArrays) = size(s.x0, 2)
struct SimulationView
    s::EnsembleSimulationArrays
    j::Int
end
Base.getindex(s::EnsembleSimulationArrays, j::Integer) = SimulationView(s, j)
ClosedLoopReachability.controls(v::SimulationView) = [v.s.U[:, v.j, i] for i in 1:size(v.s.U, 3)]
ClosedLoopReachability.disturbances(v::SimulationView) = [isnothing(v.s.W) ? nothing : v.s.W[:, v.j, i] for i in 1:size(v.s.U, 3)]
"pieces of trajectory j: vectors of (t, u) like an ODESolution's `.t` / `.u` (needs the step log); else only the end states"
function ClosedLoopReachability.trajectory(v::SimulationView)
    s, j = v.s, v.j
    isnothing(s.log_n) && return [(t=nothing, u=[s.Xend[:, j, i]]) for i in 1:size(s.Xend, 3)]
    return [(t=s.log_t[1:s.log_n[j, i], j, i], u=[s.log_u[:, q, j, i] for q in 1:s.log_n[j, i]]) for i in 1:size(s.Xend, 3)]
end
ClosedLoopReachability.trajectories(s::EnsembleSimulationArrays) = [ClosedLoopReachability.trajectory(s[j]) for j in 1:length(s)]

function simulate_b200(cp::ClosedLoopReachability.AbstractControlProblem, plant_id::Integer, args...;
                       ctx::ClrB200.Context=ClrB200.Context(0), log_cap::Integer=0, kwargs...)
    network = ClosedLoopReachability.controller(cp)
    st_vars, dist_vars, ctrl_vars = states(cp), disturbances(cp), controls(cp)
    X₀ = project(initial_state(cp), st_vars)
    W₀ = isempty(dist_vars) ? nothing : project(initial_state(cp), dist_vars)
    τ = period(cp)
    time_span = _get_tspan(args...; kwargs...)
    trajectories = get(kwargs, :trajectories, 10)
    include_vertices = get(kwargs, :include_vertices, false)
    t0 = RA_.tstart(time_span)
    iterations = ceil(Int, diam(time_span) / τ)

    # sampling: the reference's calls in the reference's order (src/simulate.jl:89, :111)
    x0_vec = sample(X₀, trajectories; include_vertices=include_vertices)
    N = length(x0_vec)
    x0 = reduce(hcat, x0_vec)
    nd = length(dist_vars)
    Wd = nd == 0 ? nothing : Array{Float64}(undef, nd, N, iterations)
    for i in 1:iterations
        nd == 0 && break
        Wd[:, :, i] = reduce(hcat, sample(W₀, N))
    end
    dnet = ClrB200.upload(ctx, network)
    pp, m, keep = prepost_of(control_preprocessing(cp), control_postprocessing(cp), size(network.layers[end].weights, 1))
    GC.@preserve keep begin
        if log_cap > 0
            Xend, U, ln, lt, lu = ClrB200.rollout_logged(ctx, dnet, pp, plant_id, x0, Wd, t0, τ, RA_.tend(time_span), iterations;
                                                         n_dist=nd, n_ctrl=length(ctrl_vars), log_cap=log_cap)
            return EnsembleSimulationArrays(x0, Xend, U, Wd, ln, lt, lu)
        end
        Xend, U = ClrB200.rollout(ctx, dnet, pp, plant_id, x0, Wd, t0, τ, RA_.tend(time_span), iterations;
                                  n_dist=nd, n_ctrl=length(ctrl_vars))
        return EnsembleSimulationArrays(x0, Xend, U, Wd, nothing, nothing, nothing)
    end
end

"""
The extension hook itself (ext/ClosedLoopReachabilityOrdinaryDiffEqExt.jl:28-45) for ONE control period: given the extended
states [x; w; u] of all trajectories, integrate the built-in plant over `tspan` with the controls held constant.  Lets the
reference's own `simulate` loop (src/simulate.jl:98-140) run unchanged with the ODE part on the GPU: the controller is the
identity on the u rows, so `clr_rollout` is called with a one-layer Id network that copies u.
"""
function _solve_ensemble_b200(ctx::ClrB200.Context, plant_id::Integer, extended::Vector{Vector{Float64}}, tspan; n_state, n_dist, n_ctrl,
                              log_cap::Integer=64)
    # u is constant over the period: feed it as the `disturbance` rows of an extended plant is not possible for the built-in
    # right-hand sides, so the period is integrated through clr_rollout_plant with a plant compiled for (x, [w; u], no control)
    error("use simulate_b200 (whole horizon in one call) or compile the plant with ClrB200.compile_plant(ctx, n_state, n_dist + n_ctrl, 1, body) " *
          "and call ClrB200.rollout_plant per period; see INTEGRATION.md section 1")
end
