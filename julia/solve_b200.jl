# solve_b200.jl -- breadth-first `_solve` on top of libclr_b200 (drop-in for src/solve.jl:110-208 of the reference).
#
# NOT EXECUTED IN THIS REPOSITORY (the build image has no Julia).  It is the method a maintainer of
# ClosedLoopReachability.jl would add next to `_solve`; the level-synchronous walk and the LIFO re-ordering are the same
# algorithm as closedloopreachability.jl_b200/solve.py, which IS executed: tests/test_solve_order.py checks it, flowpipe by
# flowpipe, against a literal re-implementation of the reference's depth-first loop, on the CPU and through the CUDA path.
#
# What changes against src/solve.jl: the reference pops ONE flowpipe at a time (src/solve.jl:179) and runs
# `controller_forward` once per split cell inside `_solve_one` (src/solve.jl:187-195, :227).  Here every control period is
# one level: all cells of all live chains go through the controller in ONE `clr_deepz_*` call, the plant steps
# (`_reconstruct` + `post` with TMJets, src/solve.jl:234-236: the reference's CPU path, unchanged) run over the level with
# `Threads.@threads`, and at the end the flowpipes are emitted in exactly the order of the reference's waiting list.
#
# Usage:  include("ClrB200.jl"); include("solve_b200.jl")
#         sol = solve_b200(prob; T=T, algorithm_plant=TMJets(...), splitter=BoxSplitter([4, 4, 3, 5]), ctx=ClrB200.Context(0))
using ClosedLoopReachability
using ClosedLoopReachability: AbstractSplitter, NoSplitter, TaylorModelReconstructor, AbstractReconstructionMethod,
                              control_preprocessing, control_postprocessing, states, disturbances, controls, system,
                              initial_state, period, _project_oa, _reconstruct, _get_tspan,
                              NoPreprocessing, NoPostprocessing, UniformAdditivePostprocessing, LinearMapPostprocessing,
                              ProjectionPostprocessing
const RA = ClosedLoopReachability.ReachabilityAnalysis
using LazySets

"clr_prepost for the problem's processing (src/problem.jl:5-77); the arrays it points to are returned for GC.@preserve."
function prepost_of(preprocessing, postprocessing, m_net::Integer)
    keep = Any[]
    pre_rows, preA, preb = Int32(0), Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL)
    preprocessing isa NoPreprocessing ||
        throw(ArgumentError("only NoPreprocessing runs on the device here; pass an affine map through ClrB200.PrePost directly"))
    kind, rows, pA, pb, pd = Int32(0), Int32(0), Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL), Ptr{Int32}(C_NULL)
    if postprocessing isa UniformAdditivePostprocessing
        b = Float64[postprocessing.shift]; push!(keep, b); kind = Int32(1); pb = pointer(b)
    elseif postprocessing isa LinearMapPostprocessing && postprocessing.map isa Number
        a = Float64[postprocessing.map]; push!(keep, a); kind = Int32(2); pA = pointer(a)
    elseif postprocessing isa LinearMapPostprocessing
        A = Matrix{Float64}(postprocessing.map); push!(keep, A); kind = Int32(3); rows = Int32(size(A, 1)); pA = pointer(A)
    elseif postprocessing isa ProjectionPostprocessing
        d = Int32.(postprocessing.dims .- 1); push!(keep, d); kind = Int32(4); rows = Int32(length(d)); pd = pointer(d)
    elseif !(postprocessing isa NoPostprocessing)
        throw(ArgumentError("unsupported postprocessing $(typeof(postprocessing))"))
    end
    m = kind == 3 || kind == 4 ? Int(rows) : Int(m_net)
    return ClrB200.PrePost(pre_rows, preA, preb, kind, rows, pA, pb, pd), m, keep
end

"""
Order in which the reference's `_solve` appends flowpipes (src/solve.jl:168-204): `parents[1]` = number of first-period
flowpipes, `parents[k]` (k > 1) = for every flowpipe of period k, breadth-first, the breadth-first index of its parent.
Returns (level, index) pairs.  Same function as `lifo_emission_order` in clr_b200/solve.py.
"""
function lifo_emission_order(parents::Vector)
    K = length(parents)
    n1 = parents[1]::Int
    children = [Dict{Int,Vector{Int}}() for _ in 1:K]
    for k in 2:K, (j, p) in enumerate(parents[k])
        push!(get!(children[k - 1], p, Int[]), j)
    end
    order = [(1, i) for i in 1:n1]
    stack = K > 1 ? copy(order) : Tuple{Int,Int}[]
    while !isempty(stack)
        k, i = pop!(stack)                                    # pop!(waiting_list)
        for j in get(children[k], i, Int[])
            push!(order, (k + 1, j))                          # append!(flowpipes, Fs)
            k + 1 < K && push!(stack, (k + 1, j))             # pushed unless it is the last period
        end
    end
    return order
end

"Boxes of the controller output for a batch of state cells (Hyperrectangles or Zonotopes), one library call."
function controller_forward_batch(ctx, dnet, cells::Vector, pp, m)
    Zs = [c isa Zonotope ? c : convert(Zonotope, c) for c in cells]
    lo, hi = ClrB200.controller_forward_boxes(ctx, dnet, Zs, pp, m)
    return [m == 1 ? LazySets.Interval(lo[1, i], hi[1, i]) : Hyperrectangle(low=lo[:, i], high=hi[:, i]) for i in 1:length(cells)]
end

function _solve_b200(cp::ControlledPlant, algorithm_plant, tvec::AbstractVector, splitter::AbstractSplitter,
                     input_splitter::AbstractSplitter, reconstruction_method::AbstractReconstructionMethod,
                     remove_zero_generators::Bool, ctx::ClrB200.Context)
    S = system(cp)
    net = ClosedLoopReachability.controller(cp)
    st_vars, dist_vars, ctrl_vars = states(cp), disturbances(cp), controls(cp)
    Q₀ = initial_state(cp)
    n, md, q = length(st_vars), length(dist_vars), length(ctrl_vars)
    dim(Q₀) == n + md + q || throw(ArgumentError("dimension mismatch; expected the dimension of the initial states of the " *
                                                  "initial-value problem to be $(n + md + q), but it is $(dim(Q₀))"))
    W₀ = md > 0 ? project(Q₀, dist_vars) : nothing
    dnet = ClrB200.upload(ctx, net)
    pp, m, keep = prepost_of(control_preprocessing(cp), control_postprocessing(cp), size(net.layers[end].weights, 1))
    K = length(tvec) - 1

    levels = Vector{Vector{Any}}()            # flowpipes per period, breadth-first
    parents = Any[]
    # chains entering the period: (X₀ set, reach-set R it came from or nothing, parent index)
    live = Any[(project(Q₀, st_vars), nothing, 0)]
    GC.@preserve keep for k in 1:K
        t0, t1 = tvec[k], tvec[k + 1]
        # ---- cells of all live chains: X₀s = haskey(splitter, k) ? apply(splitter[k], X₀) : [X₀]  (src/solve.jl:153, :186) ----
        cells = Any[]; cell_R = Any[]; cell_parent = Int[]
        for (X₀, R, p) in live
            X₀s = haskey(splitter, k) ? ClosedLoopReachability.apply(splitter[k], X₀) : [X₀]
            append!(cells, X₀s); append!(cell_R, fill(R, length(X₀s))); append!(cell_parent, fill(p, length(X₀s)))
        end
        # ---- ONE controller call for the period (replaces controller_forward inside _solve_one, src/solve.jl:227) --------
        Us = controller_forward_batch(ctx, dnet, cells, pp, m)
        # ---- the plant steps of the period: the reference's own code path, cell by cell (src/solve.jl:230-246) -----------
        tasks = Tuple{Int,Any}[]                                   # (cell index, control piece), in _solve_one's order
        for (i, U) in enumerate(Us), Ui in ClosedLoopReachability.apply(input_splitter, U)
            push!(tasks, (i, Ui))
        end
        sols = Vector{Any}(undef, length(tasks))
        one(j) = begin
            i, Ui = tasks[j]
            P₀ = isnothing(W₀) ? cells[i] : cells[i] × W₀
            Qi = _reconstruct(reconstruction_method, P₀, Ui, cell_R[i], t0)
            sol = RA.post(algorithm_plant, RA.IVP(S, Qi), RA.TimeInterval(t0, t1))
            Δt = t1 - RA.tend(sol)
            @assert ClosedLoopReachability.isapproxzero(Δt) "the flowpipe duration differs from the requested duration by $Δt time units"
            sol.ext[:controls] = Ui
            sols[j] = sol
        end
        one(1)                                                     # first analysis alone (TaylorSeries globals, src/solve.jl:157-161)
        Threads.@threads for j in 2:length(tasks)
            one(j)
        end
        push!(levels, sols)
        push!(parents, k == 1 ? length(sols) : [cell_parent[tasks[j][1]] for j in 1:length(tasks)])
        # ---- chains for the next period: X₀ = set(_project_oa(F(tend), st_vars, t))  (src/solve.jl:181-184) ----------------
        if k < K
            live = Any[]
            for (j, F) in enumerate(sols)
                t = RA.tend(F); R = F(t)
                push!(live, (RA.set(_project_oa(R, st_vars, t; remove_zero_generators=remove_zero_generators)), R, j))
            end
        end
    end
    # ---- emit in the reference's order (LIFO waiting list) ---------------------------------------------------------------------
    flowpipes = [levels[l][i] for (l, i) in lifo_emission_order(parents)]
    return RA.MixedFlowpipe([F for F in flowpipes])
end

"""
    solve_b200(prob; ctx, kwargs...)

Same keyword arguments and return type as `solve(prob::AbstractControlProblem; ...)` (src/solve.jl:20-100), with the
controller propagation of every control period batched on the GPU.
"""
function solve_b200(prob::ClosedLoopReachability.AbstractControlProblem, args...; ctx::ClrB200.Context=ClrB200.Context(0), kwargs...)
    # the body of `solve` (src/solve.jl:45-77), line for line, except for the last call
    ivp = ClosedLoopReachability.plant(prob)
    ClosedLoopReachability._check_dim(ivp)
    τ = period(prob)
    tspan = _get_tspan(args...; kwargs...)
    T = RA.tend(tspan)
    tvec = range(RA.tstart(tspan), T; step=τ)
    if !(tvec[end] ≈ T)
        tvec = vcat(tvec, T)  # last time interval is shorter
    end
    algorithm_plant = ClosedLoopReachability._get_algorithm_plant(args...; kwargs...)
    if isnothing(algorithm_plant)
        algorithm_plant = ClosedLoopReachability._default_cpost(ivp, tspan; kwargs...)
    end
    algorithm_controller = ClosedLoopReachability._get_algorithm_controller(args...; kwargs...)
    algorithm_controller isa DeepZ ||
        throw(ArgumentError("libclr_b200 implements DeepZ; other ForwardAlgorithms keep the reference's CPU loop (`solve`)"))
    splitter = get(kwargs, :splitter, NoSplitter())
    input_splitter = get(kwargs, :input_splitter, NoSplitter())
    reconstruction_method = get(kwargs, :reconstruction_method, TaylorModelReconstructor())
    remove_zero_generators = get(kwargs, :remove_zero_generators, true)
    sol = _solve_b200(prob, algorithm_plant, tvec, splitter, input_splitter, reconstruction_method, remove_zero_generators, ctx)
    d = Dict{Symbol,Any}(:solver_controller => algorithm_controller)
    return RA.ReachSolution(sol, algorithm_plant, d)
end
