# dump_reference.jl -- produce reference-pinned golden vectors from the REAL ClosedLoopReachability.jl.
#
# STATUS: never executed in the build image (no `julia`, no Julia depot, no network).  It exists so that the
# day a Julia installation with ClosedLoopReachability.jl is at hand, ONE command turns "parity: unpinned" into
# "parity: pinned":
#
#     julia --project=<env with ClosedLoopReachability, OrdinaryDiffEq> tests/golden/dump_reference.jl \
#           /path/to/ClosedLoopReachability.jl  tests/golden/reference
#
# It writes plain JSON files (Float64 printed with Julia's shortest round-trip representation, so that Python's
# `json` reads back the identical bits; no JSON.jl / NPZ.jl dependency).  tests/test_reference_goldens.py consumes
# every file it finds there (CPU oracle in the `not gpu` suite, the CUDA path in the `gpu` suite) and reports
# "parity unpinned" while the directory is empty.
#
# What is dumped (SURVEY.md section 8c item 6, VERDICT r1 "next" item 1):
#   nets.json        weights as read by read_POLAR (+ their eltype: settles the Float32/Float64 parsing question)
#   splits.json      LazySets.split cell tables (centres, radii, in the order apply(BoxSplitter(p), X) returns them)
#                    and generator splits (ZonotopeSplitter), src/splitters.jl:20-43
#   deepz.json       per cell and per network prefix (layers 1..l): centre and generator matrix of
#                    forward(X, net[1:l], DeepZ()), src/solve.jl:212; and the box of controller_forward, src/solve.jl:210-218
#   simulate.json    simulate(prob; T, trajectories) with Random.seed!: sampled x0, disturbances, controls and the accepted
#                    (t, u) steps of every trajectory piece (src/simulate.jl:68-143, ext/...OrdinaryDiffEqExt.jl:28-45)
#   setops.json      _project_oa / _reduce_order on the first flowpipes of a short TORA solve (src/setops.jl:9-12, :84-101)
#   order.json       emission order of solve() with a splitter at k = 1 and an IndexedSplitter at k = 2 and a SignSplitter on
#                    the inputs: (k, control interval) of every flowpipe in the order of the returned MixedFlowpipe
#                    (src/solve.jl:168-204: LIFO waiting list)
using Random

const REF = length(ARGS) >= 1 ? ARGS[1] : error("usage: dump_reference.jl <ClosedLoopReachability.jl checkout> <out dir>")
const OUT = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "reference")
mkpath(OUT)

using ClosedLoopReachability
import OrdinaryDiffEq
const CLR = ClosedLoopReachability
using ClosedLoopReachability: UniformAdditivePostprocessing, LinearMapPostprocessing, NoSplitter, BoxSplitter,
                              ZonotopeSplitter, IndexedSplitter, SignSplitter, Splitter, controller_forward,
                              FeedforwardNetwork, DenseLayerOp
const RA = ClosedLoopReachability.ReachabilityAnalysis
using ClosedLoopReachability.ReachabilityAnalysis: TMJets, tend
using LazySets

# ---- minimal JSON writer -----------------------------------------------------------------------------------------
js(x::AbstractFloat) = isfinite(x) ? repr(Float64(x)) : (isnan(x) ? "\"nan\"" : (x > 0 ? "\"inf\"" : "\"-inf\""))
js(x::Integer) = string(x)
js(x::Bool) = x ? "true" : "false"
js(x::AbstractString) = "\"" * escape_string(x) * "\""
js(x::Symbol) = js(string(x))
js(::Nothing) = "null"
js(x::AbstractVector) = "[" * join((js(v) for v in x), ",") * "]"
js(x::Tuple) = js(collect(x))
js(x::AbstractMatrix) = "[" * join((js(collect(x[i, :])) for i in 1:size(x, 1)), ",") * "]"   # row-major list of rows
js(x::AbstractDict) = "{" * join((js(string(k)) * ":" * js(v) for (k, v) in x), ",") * "}"
function dump(name, obj)
    open(joinpath(OUT, name), "w") do io
        write(io, js(obj))
    end
    println("wrote ", joinpath(OUT, name))
end

model(parts...) = joinpath(REF, "models", parts...)
actname(a) = string(nameof(typeof(a)))
zono(Z) = Dict("c" => collect(center(Z)), "G" => Matrix(genmat(Z)))
hbox(H) = Dict("lo" => collect(low(H)), "hi" => collect(high(H)))

# ---- networks ----------------------------------------------------------------------------------------------------
nets = Dict{String,Any}()
netfiles = Dict("TORA_ReLU" => ("TORA", "TORA_ReLU_controller.polar"),
                "TORA_sigmoid" => ("TORA", "TORA_sigmoid_controller.polar"),
                "TORA_ReLUtanh" => ("TORA", "TORA_ReLUtanh_controller.polar"),
                "ACC_relu" => ("ACC", "ACC_controller_relu.polar"),
                "ACC_tanh" => ("ACC", "ACC_controller_tanh.polar"),
                "Quadrotor" => ("Quadrotor", "Quadrotor_controller.polar"),
                "Unicycle" => ("Unicycle", "Unicycle_controller.polar"),
                "InvertedPendulum" => ("InvertedPendulum", "InvertedPendulum_controller.polar"))
loaded = Dict{String,Any}()
for (name, parts) in netfiles
    path = model(parts...)
    isfile(path) || continue
    net = read_POLAR(path)
    loaded[name] = net
    nets[name] = Dict("layers" => [Dict("W" => Matrix{Float64}(L.weights), "b" => Vector{Float64}(L.bias),
                                        "act" => actname(L.activation), "eltype" => string(eltype(L.weights)))
                                   for L in net.layers])
end
dump("nets.json", nets)

# ---- splits --------------------------------------------------------------------------------------------------------
splits = Dict{String,Any}()
function boxsplit(name, lo, hi, partition)
    cells = CLR.apply(BoxSplitter(partition), Hyperrectangle(low=lo, high=hi))
    splits[name] = Dict("lo" => lo, "hi" => hi, "partition" => partition,
                        "centers" => [collect(center(c)) for c in cells],
                        "radii" => [collect(radius_hyperrectangle(c)) for c in cells])
    return cells
end
tora_lo, tora_hi = [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]
cells_tora = boxsplit("tora_4435", tora_lo, tora_hi, [4, 4, 3, 5])
boxsplit("tora_1010101", tora_lo, tora_hi, [10, 10, 10, 1])
boxsplit("unicycle_3171", [9.5, -4.5, 2.1, 1.5], [9.55, -4.45, 2.11, 1.51], [3, 1, 7, 1])
boxsplit("thirds", [0.0, -1.0, 0.1], [1.0, 2.0, 0.1 + 1e-9], [3, 7, 13])          # ranges whose steps are not dyadic
boxsplit("flat_dim", [1.0, 0.1], [1.2, 0.1], [4, 1])                                # zero-radius dimension
boxsplit("c4_stable", fill(0.29, 6), fill(0.31, 6), [10, 10, 10, 1, 1, 1])
let Z = Zonotope([0.65, -0.65, -0.35, 0.55], [0.05 0.0 0.0 0.0 1e-3; 0.0 0.05 0.0 0.0 -2e-3; 0.0 0.0 0.05 0.0 5e-4; 0.0 0.0 0.0 0.05 1e-3])
    cells = CLR.apply(ZonotopeSplitter([1, 2, 3, 5], [2, 2, 1, 1]), Z)
    splits["zonotope_split"] = Dict("c" => collect(center(Z)), "G" => Matrix(genmat(Z)), "gens" => [1, 2, 3, 5],
                                    "nparts" => [2, 2, 1, 1], "cells" => [zono(c) for c in cells])
    cells = CLR.apply(ZonotopeSplitter(), Z)
    splits["zonotope_split_default"] = Dict("cells" => [zono(c) for c in cells])
end
splits["sign_split"] = [[[low(I, 1), high(I, 1)] for I in CLR.apply(SignSplitter(), LazySets.Interval(a, b))]
                        for (a, b) in ((-1.0, 2.0), (0.0, 2.0), (-3.0, -1.0), (-1.0, 0.0))]
dump("splits.json", splits)

# ---- DeepZ forward, per network prefix -----------------------------------------------------------------------------
prefix(net, l) = FeedforwardNetwork(net.layers[1:l])
function deepz_case(net, X; pre=CLR.NoPreprocessing(), post=CLR.NoPostprocessing())
    Xz = X isa Zonotope ? X : convert(Zonotope, X)
    out = Dict{String,Any}("input" => zono(Xz), "layers" => Any[])
    for l in 1:length(net.layers)
        Z = forward(Xz, prefix(net, l), DeepZ())
        push!(out["layers"], zono(Z))
    end
    U = controller_forward(DeepZ(), net, X, pre, post)
    out["U_lo"] = collect(low(U)); out["U_hi"] = collect(high(U))
    return out
end
deepz = Dict{String,Any}()
if haskey(loaded, "TORA_ReLU")
    post = UniformAdditivePostprocessing(-10.0)
    deepz["TORA_ReLU_4435"] = [deepz_case(loaded["TORA_ReLU"], c; post=post) for c in cells_tora[1:8:end]]
    deepz["TORA_ReLU_X0"] = [deepz_case(loaded["TORA_ReLU"], Hyperrectangle(low=tora_lo, high=tora_hi); post=post)]
end
for name in ("TORA_sigmoid", "TORA_ReLUtanh")
    haskey(loaded, name) || continue
    X2 = Hyperrectangle(low=[-0.77, -0.45, 0.51, -0.3], high=[-0.75, -0.43, 0.54, -0.28])
    deepz[name * "_X0"] = [deepz_case(loaded[name], X2; post=LinearMapPostprocessing(11.0))]
end
if haskey(loaded, "Quadrotor")
    r = 0.01 .* vcat(fill(0.4, 6), zeros(6))
    cells = CLR.apply(BoxSplitter(vcat(fill(2, 6), fill(1, 6))), Hyperrectangle(zeros(12), r))
    deepz["Quadrotor_2p6"] = [deepz_case(loaded["Quadrotor"], c) for c in cells[1:9:end]]
end
if haskey(loaded, "Unicycle")
    cells = CLR.apply(BoxSplitter([3, 1, 7, 1]), Hyperrectangle(low=[9.5, -4.5, 2.1, 1.5], high=[9.55, -4.45, 2.11, 1.51]))
    deepz["Unicycle_3171"] = [deepz_case(loaded["Unicycle"], c; post=UniformAdditivePostprocessing(-20.0)) for c in cells]
end
if haskey(loaded, "ACC_relu")
    # the affine preprocessing of models/ACC/ACC.jl:81-99 on a zonotope with dense generators
    M = zeros(3, 6); M[1, 5] = 1.0; M[2, 1] = 1.0; M[2, 4] = -1.0; M[3, 2] = 1.0; M[3, 5] = -1.0
    A = vcat(zeros(2, 6), M); b = vcat([30.0, 1.4], zeros(3))
    pre = CLR.FunctionPreprocessing(X -> affine_map(A, X, b))
    rng = MersenneTwister(5)
    Xs = [Zonotope([90.0, 32.0, 0.0, 10.0, 30.0, 0.0] .+ rand(rng, 6), 0.05 .* (2 .* rand(rng, 6, 9) .- 1)) for _ in 1:6]
    deepz["ACC_relu_pre"] = [begin
                                 d = deepz_case(loaded["ACC_relu"], affine_map(A, X, b))
                                 U = controller_forward(DeepZ(), loaded["ACC_relu"], X, pre, CLR.NoPostprocessing())
                                 d["raw_input"] = zono(X); d["U_lo"] = collect(low(U)); d["U_hi"] = collect(high(U)); d
                             end for X in Xs]
end
let rng = MersenneTwister(11)
    # random zonotopes through every network (the layout tests/test_gpu_deepz.py::test_random_zonotopes_vs_oracle uses)
    for (name, net) in loaded
        n = size(net.layers[1].weights, 2)
        Xs = [Zonotope(1.2 .* rand(rng, n) .- 0.6, 0.02 .* (2 .* rand(rng, n, rand(rng, 0:14)) .- 1)) for _ in 1:12]
        deepz["random_" * name] = [deepz_case(net, X) for X in Xs]
    end
end
dump("deepz.json", deepz)

# ---- simulate() ------------------------------------------------------------------------------------------------------
sims = Dict{String,Any}()
function dump_sim(name, prob; T, trajectories, seed)
    Random.seed!(seed)
    sim = simulate(prob; T=T, trajectories=trajectories)
    out = Any[]
    for j in 1:length(sim)
        sol = sim[j]
        pieces = [Dict("t" => collect(p.t), "u" => [collect(u) for u in p.u]) for p in CLR.trajectory(sol)]
        push!(out, Dict("pieces" => pieces, "controls" => [collect(c) for c in CLR.controls(sol)],
                        "disturbances" => [d === nothing ? nothing : collect(d) for d in CLR.disturbances(sol)]))
    end
    sims[name] = Dict("T" => T, "seed" => seed, "period" => CLR.period(prob), "trajectories" => out,
                      "OrdinaryDiffEq" => string(pkgversion(OrdinaryDiffEq)))
end
# The model scripts define their problems inside modules that also run the (expensive) analyses when included, so the two
# plants used here are restated from models/Unicycle/Unicycle.jl:40-75 and models/TORA/TORA.jl:44-100.
@taylorize function Unicycle!(dx, x, p, t)
    x₁, x₂, x₃, x₄, w, u₁, u₂ = x
    dx[1] = x₄ * cos(x₃)
    dx[2] = x₄ * sin(x₃)
    dx[3] = u₂
    dx[4] = u₁ + w
    dx[5] = zero(x₁)
    dx[6] = zero(x₁)
    dx[7] = zero(x₁)
    return dx
end
@taylorize function TORA!(dx, x, p, t)
    x₁, x₂, x₃, x₄, u = x
    dx[1] = x₂
    dx[2] = -x₁ + (0.1 * sin(x₃))
    dx[3] = x₄
    dx[4] = u
    dx[5] = zero(u)
    return dx
end
if haskey(loaded, "Unicycle")
    X0 = Hyperrectangle(low=[9.5, -4.5, 2.1, 1.5, -1e-4], high=[9.55, -4.45, 2.11, 1.51, 1e-4])
    ivp = @ivp(x' = Unicycle!(x), dim: 7, x(0) ∈ X0 × ZeroSet(2))
    prob = ControlledPlant(ivp, loaded["Unicycle"], Dict(:states => 1:4, :disturbances => [5], :controls => 6:7), 0.2;
                           postprocessing=UniformAdditivePostprocessing(-20.0))
    dump_sim("Unicycle", prob; T=2.0, trajectories=8, seed=5)
end
if haskey(loaded, "TORA_ReLU")
    ivp = @ivp(x' = TORA!(x), dim: 5, x(0) ∈ Hyperrectangle(low=tora_lo, high=tora_hi) × ZeroSet(1))
    prob_tora = ControlledPlant(ivp, loaded["TORA_ReLU"], Dict(:states => 1:4, :controls => 5), 1.0;
                                postprocessing=UniformAdditivePostprocessing(-10.0))
    dump_sim("TORA_ReLU", prob_tora; T=5.0, trajectories=6, seed=7)
end
dump("simulate.json", sims)

# ---- solve(): emission order, _project_oa, _reduce_order ------------------------------------------------------------
if haskey(loaded, "TORA_ReLU")
    alg = TMJets(abstol=3e-2, orderT=3, orderQ=1)
    splitter = IndexedSplitter(Dict(1 => BoxSplitter([2, 2, 1, 1]), 2 => BoxSplitter([1, 1, 2, 1])))
    sol = solve(prob_tora; T=3.0, algorithm_controller=DeepZ(), algorithm_plant=alg, splitter=splitter,
                input_splitter=SignSplitter())
    order = Any[]
    for F in sol
        U = F.ext[:controls]
        push!(order, Dict("t0" => RA.tstart(F), "t1" => tend(F),
                          "U_lo" => collect(low(U)), "U_hi" => collect(high(U)),
                          "X_end" => hbox(overapproximate(RA.set(overapproximate(F(tend(F)), Zonotope, tend(F))),
                                                          Hyperrectangle))))
    end
    dump("order.json", Dict("flowpipes" => order, "splitter" => "k=1: [2,2,1,1]; k=2: [1,1,2,1]", "input_splitter" => "SignSplitter"))

    setops = Any[]
    for F in sol
        t = tend(F)
        R = F(t)
        Z = RA.set(CLR._project_oa(R, 1:4, t; remove_zero_generators=true))
        Zr = RA._reduce_order(Z, 2; force_reduction=true)   # src/setops.jl:85
        push!(setops, Dict("projected" => zono(Z), "reduced_order2" => zono(Zr)))
        length(setops) >= 6 && break
    end
    dump("setops.json", Dict("cases" => setops))
end
println("done: golden vectors of the reference in ", OUT)
