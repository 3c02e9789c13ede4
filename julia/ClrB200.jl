# ClrB200.jl -- the reference-side binding of libclr_b200.so (see INTEGRATION.md).
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia.  It is the shim a maintainer of
# ClosedLoopReachability.jl would add; every `ccall` below mirrors one prototype of include/clr_b200.h and the
# Python ctypes mirror (closedloopreachability.jl_b200/_lib.py, runtime.py) exercises exactly the same call
# sequences in the tests.
module ClrB200

using LazySets: LazySets, Hyperrectangle, Zonotope, Interval, center, genmat, ngens, low, high,
                box_approximation, radius_hyperrectangle, dim
using ControllerFormats: FeedforwardNetwork, DenseLayerOp, Id, ReLU, Sigmoid, Tanh

const LIB = get(ENV, "CLR_B200_LIB", "libclr_b200.so")

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:clr_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, out)
        rc == 0 || error(unsafe_string(ccall((:clr_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(out[])
        finalizer(close, ctx)
        return ctx
    end
end
function Base.close(c::Context)
    c.h == C_NULL || ccall((:clr_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h)
    c.h = C_NULL
    return nothing
end

function check(ctx::Context, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:clr_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))
    # CLR_ERR_ARG (-1) / CLR_ERR_DIM (-3) are what src/solve.jl:88-99, :130-134 raise as ArgumentError
    (rc == -1 || rc == -3) ? throw(ArgumentError(msg)) : error("libclr_b200 ($rc): $msg")
end

act_id(::Id) = Int32(0); act_id(::ReLU) = Int32(1); act_id(::Sigmoid) = Int32(2); act_id(::Tanh) = Int32(3)

# clr_net: freed by its own finalizer; it holds its Context so that the context outlives the network (clr_net_free is also
# safe after clr_destroy: the library keeps the device id in the clr_net, not a pointer into the context)
mutable struct DeviceNetwork
    h::Ptr{Cvoid}
    net::FeedforwardNetwork
    ctx::Context
    function DeviceNetwork(h, net, ctx)
        d = new(h, net, ctx)
        finalizer(x -> (x.h == C_NULL || ccall((:clr_net_free, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), d)
        return d
    end
end

function upload(ctx::Context, net::FeedforwardNetwork)
    Ws = [Matrix{Float64}(L.weights) for L in net.layers]      # column-major, widened exactly
    bs = [Vector{Float64}(L.bias) for L in net.layers]
    dims = Int32[size(Ws[1], 2); [size(W, 1) for W in Ws]]
    acts = Int32[act_id(L.activation) for L in net.layers]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Ws bs begin
        Wp = [pointer(W) for W in Ws]; bp = [pointer(b) for b in bs]
        check(ctx, ccall((:clr_net_upload, LIB), Int32,
                         (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ref{Ptr{Cvoid}}),
                         ctx.h, length(Ws), dims, acts, Wp, bp, out))
    end
    return DeviceNetwork(out[], net, ctx)
end

# clr_prepost, field for field
struct PrePost
    pre_rows::Int32; preA::Ptr{Float64}; preb::Ptr{Float64}
    post_kind::Int32; post_rows::Int32; postA::Ptr{Float64}; postb::Ptr{Float64}; post_dims::Ptr{Int32}
end
PrePost() = PrePost(0, C_NULL, C_NULL, 0, 0, C_NULL, C_NULL, C_NULL)

"""
Array-backed replacement of `apply(BoxSplitter(partition), X)` (src/splitters.jl:20-28): the tables of
`split(box_approximation(X), partition)`; cells stay implicit (cell id -> multi-index, dimension 1 fastest).
"""
struct BoxGrid
    partition::Vector{Int32}
    centers::Vector{Float64}      # concatenated per-dimension tables
    radii::Vector{Float64}
end

function BoxGrid(X, partition::AbstractVector{<:Integer})
    H = box_approximation(X); c = center(H); R = radius_hyperrectangle(H)
    cs = Float64[]; rs = Float64[]
    for (i, m) in enumerate(partition)
        r = R[i] / m
        lo = c[i] - R[i]
        # LazySets' split: block centres are range(lo + r; step=2r, length=m); a single block has the centre lo + R (not c:
        # the two can differ by one ulp).  The Python mirror (splitters.py) emulates Julia's range with exact rationals; it
        # has not been checked against a real Julia (tests/golden/dump_reference.jl produces the tables that would).
        append!(cs, m == 1 ? [lo + R[i]] : collect(range(lo + r; step=2r, length=m)))
        append!(rs, fill(m == 1 ? R[i] : r, m))
    end
    return BoxGrid(Int32.(partition), cs, rs)
end
Base.length(g::BoxGrid) = prod(Int.(g.partition))

"""
Batched `controller_forward` (src/solve.jl:210-218) over all cells of a split: replaces the loops at
src/solve.jl:162-166 and :190-195.  Returns the m x N matrices of `low` / `high` of U per cell; with the default
`TaylorModelReconstructor(box=true)` (src/setops.jl:76-83) that is everything `_reconstruct` consumes.
"""
function controller_forward_boxes(ctx::Context, dnet::DeviceNetwork, g::BoxGrid, pp::PrePost, m::Integer;
                                  cell_begin::Integer=0, cell_count::Integer=length(g) - cell_begin)
    lo = Matrix{Float64}(undef, m, cell_count); hi = similar(lo)
    check(ctx, ccall((:clr_deepz_boxgrid, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ref{PrePost},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int32, Int32),
                     ctx.h, dnet.h, length(g.partition), g.partition, g.centers, g.radii, cell_begin, cell_count, pp,
                     lo, hi, C_NULL, C_NULL, C_NULL, 0, 2 #= CLR_FLAG_NO_STATS =#))
    return lo, hi
end

"Same for arbitrary zonotope cells (control steps k > 1: X0s = [set(_project_oa(R, st_vars, t))], src/solve.jl:183-187)."
function controller_forward_boxes(ctx::Context, dnet::DeviceNetwork, Zs::AbstractVector{<:Zonotope}, pp::PrePost, m::Integer)
    n = dim(first(Zs)); N = length(Zs)
    C = reduce(hcat, center.(Zs)); off = Int64[0; cumsum(ngens.(Zs))]; G = reduce(hcat, genmat.(Zs))
    lo = Matrix{Float64}(undef, m, N); hi = similar(lo)
    check(ctx, ccall((:clr_deepz_zonotopes, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ref{PrePost},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int32, Int32),
                     ctx.h, dnet.h, n, N, C, off, G, pp, lo, hi, C_NULL, C_NULL, C_NULL, 0, 2))
    return lo, hi
end

"""
Same for cells in the padded layout `project_oa_batch` returns (centres n x N, generators n x cap x N, counts): the cells of
control period k + 1 without building a `Zonotope` per chain (src/solve.jl:183-195).
"""
function controller_forward_boxes(ctx::Context, dnet::DeviceNetwork, C::Matrix{Float64}, G::Array{Float64,3}, ncols::Vector{Int32},
                                  pp::PrePost, m::Integer)
    n, N = size(C); cap = size(G, 2)
    lo = Matrix{Float64}(undef, m, N); hi = similar(lo)
    check(ctx, ccall((:clr_deepz_zonotopes_padded, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Ref{PrePost},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int32, Int32),
                     ctx.h, dnet.h, n, N, C, G, ncols, cap, pp, lo, hi, C_NULL, C_NULL, C_NULL, 0, 2))
    return lo, hi
end

"""
`TaylorModelReconstructor(box=false)` (src/setops.jl:84-101) for all cells of the last `keep_cap > 0` run at once:
`Z0 = _reduce_order(U0, 2; force_reduction=true)`.  Returns the centres (m x N) and the generators (m x 2m x N);
cell i's `M = G[:, 1:m, i]` and diagonal `D = G[:, m+1:2m, i]` are what the loop at src/setops.jl:96-101 reads.
"""
function reduced_controls(ctx::Context, m::Integer, N::Integer; order::Integer=2)
    c = Matrix{Float64}(undef, m, N); G = Array{Float64}(undef, m, m * order, N)
    check(ctx, ccall((:clr_deepz_fetch_reduced, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int32),
                     ctx.h, order, c, G, 0))
    return c, G
end

"""
Batched `_project_oa(R::AbstractTaylorModelReachSet, vars, t)` (src/setops.jl:9-12) for the reach-sets `Rs` that end the
live chains of a control period (src/solve.jl:181-184): one call instead of one `overapproximate(R, Zonotope, t)` +
`project` per chain.  The Taylor models are flattened on the Julia side: component v of reach-set i is a
`TaylorModel1{TaylorN{Float64},Float64}`; its time coefficient k contributes `vcat((c.coeffs for c in p[k].coeffs)...)`
(all homogeneous polynomials, degree 0 first -- TaylorSeries' coefficient order), the exponent table is
`reduce(hcat, vcat(TaylorSeries.coeff_table...))`, and dt = t - tstart(R).  Returns the centres (n_keep x N), the
generators (n_keep x 2n x N) and the generator counts: the arguments `clr_deepz_zonotopes_padded` takes for the next period
(with CLR_FLAG_DEVICE_IO the two calls chain on the device).
"""
function project_oa_batch(ctx::Context, exps::Matrix{Int32}, coeffs::Array{Float64,4}, rem::Array{Float64,3}, dt::Vector{Float64},
                          vars::Vector{Int32}; remove_zero_generators::Bool=true)
    n, M = size(exps); KT = size(coeffs, 2); N = size(coeffs, 4)          # coeffs: M x KT x n x N, rem: 2 x n x N
    nk = length(vars)
    c = Matrix{Float64}(undef, nk, N); G = Array{Float64}(undef, nk, 2n, N); ncols = Vector{Int32}(undef, N)
    check(ctx, ccall((:clr_project_oa, LIB), Int32,
                     (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32,
                      Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32),
                     ctx.h, n, KT - 1, M, exps, N, coeffs, rem, dt, vars .- Int32(1), nk, remove_zero_generators, c, G, ncols, 0))
    return c, G, ncols
end

"""
`simulate`'s period loop (src/simulate.jl:98-140) for a built-in plant: x0 is n_state x N, Wd n_dist x N x iters
(drawn on the Julia side in the reference's RNG order: x0 once, then W0 once per period, src/simulate.jl:89, :111).
"""
function rollout(ctx::Context, dnet::DeviceNetwork, pp::PrePost, plant::Integer, x0::Matrix{Float64}, Wd, t0, τ, T, iters;
                 n_dist::Integer, n_ctrl::Integer, abstol=1e-6, reltol=1e-3)
    ns, N = size(x0)
    Xend = Array{Float64}(undef, ns, N, iters); U = Array{Float64}(undef, n_ctrl, N, iters)
    check(ctx, ccall((:clr_rollout, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ref{PrePost}, Int32, Int32, Int32, Int32, Int64, Int32, Ptr{Float64}, Ptr{Float64},
                      Float64, Float64, Float64, Int32, Float64, Float64, Int32, Int32, Ptr{Float64}, Ptr{Float64},
                      Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int32),
                     ctx.h, dnet.h, pp, plant, ns, n_dist, n_ctrl, N, iters, x0, n_dist > 0 ? Wd : C_NULL,
                     t0, τ, T, 0 #= Tsit5 =#, abstol, reltol, 4, 1 #= fastpow =#, Xend, U, C_NULL, C_NULL, 0, C_NULL, C_NULL, C_NULL, 0))
    return Xend, U
end

"""
A plant that is not built in: `rhs_body` is the CUDA C body of `rhs(const double* x, const double* q, double* dx)` (x: the
dynamic states, q = [w; u]) -- the device-side twin of the user's `f!(dx, x, p, t)` (RA.inplace_field!(ivp),
ext/ClosedLoopReachabilityOrdinaryDiffEqExt.jl:33-37).  Returns the handle `clr_rollout_plant` takes in place of the plant id.
"""
function compile_plant(ctx::Context, n_state::Integer, n_dist::Integer, n_ctrl::Integer, rhs_body::AbstractString)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:clr_plant_compile, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Cstring, Ref{Ptr{Cvoid}}),
                     ctx.h, n_state, n_dist, n_ctrl, rhs_body, out))
    return out[]
end

"`forward(x, network)` + pre/post-processing for a batch of points (src/simulate.jl:101-106): X is nx x N, returns m x N."
function forward_points(ctx::Context, dnet::DeviceNetwork, pp::PrePost, X::Matrix{Float64}, m::Integer)
    nx, N = size(X)
    U = Matrix{Float64}(undef, m, N)
    check(ctx, ccall((:clr_forward_points, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ref{PrePost}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Int32),
                     ctx.h, dnet.h, pp, nx, N, X, U, 0))
    return U
end

"""
The output sets U of the last `keep_cap > 0` run (src/solve.jl:213 before the box of :214-216): generator counts (N),
centres (m x N), generators (m x keep_cap x N, reference column order, zero padded).
"""
function fetch_zonotopes(ctx::Context, m::Integer, N::Integer, keep_cap::Integer)
    ncols = Vector{Int32}(undef, N); c = Matrix{Float64}(undef, m, N); G = Array{Float64}(undef, m, keep_cap, N)
    check(ctx, ccall((:clr_deepz_fetch_zonotopes, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}), ctx.h, ncols, c, G))
    return ncols, c, G
end

"Full output zonotopes of a batch of zonotope cells: `[Zonotope(c[:, i], G[:, 1:ncols[i], i]) for i]` is `forward(X_i, net, DeepZ())`."
function controller_forward_zonotopes(ctx::Context, dnet::DeviceNetwork, Zs::AbstractVector{<:Zonotope}, pp::PrePost, m::Integer,
                                      keep_cap::Integer)
    n = dim(first(Zs)); N = length(Zs)
    C = reduce(hcat, center.(Zs)); off = Int64[0; cumsum(ngens.(Zs))]; G = reduce(hcat, genmat.(Zs))
    lo = Matrix{Float64}(undef, m, N); hi = similar(lo)
    check(ctx, ccall((:clr_deepz_zonotopes, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ref{PrePost},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int32, Int32),
                     ctx.h, dnet.h, n, N, C, off, G, pp, lo, hi, C_NULL, C_NULL, C_NULL, keep_cap, 2))
    ncols, c, Gk = fetch_zonotopes(ctx, m, N, keep_cap)
    return [Zonotope(c[:, i], Gk[:, 1:ncols[i], i]) for i in 1:N]
end

"`clr_rollout_plant`: the period loop of `simulate` for a plant compiled with `compile_plant` (same arguments as `rollout`)."
function rollout_plant(ctx::Context, dnet::DeviceNetwork, pp::PrePost, plant::Ptr{Cvoid}, x0::Matrix{Float64}, Wd, t0, τ, T, iters;
                       n_ctrl::Integer, abstol=1e-6, reltol=1e-3, counts::Bool=false)
    ns, N = size(x0)
    Xend = Array{Float64}(undef, ns, N, iters); U = Array{Float64}(undef, n_ctrl, N, iters)
    na = counts ? Matrix{Int32}(undef, N, iters) : C_NULL
    check(ctx, ccall((:clr_rollout_plant, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ref{PrePost}, Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Float64},
                      Float64, Float64, Float64, Int32, Float64, Float64, Int32, Int32, Ptr{Float64}, Ptr{Float64},
                      Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int32),
                     ctx.h, dnet.h, pp, plant, N, iters, x0, Wd === nothing ? C_NULL : Wd,
                     t0, τ, T, 0, abstol, reltol, 4, 1, Xend, U, na, C_NULL, 0, C_NULL, C_NULL, C_NULL, 0))
    return Xend, U, na
end
free_plant(p::Ptr{Cvoid}) = ccall((:clr_plant_free, LIB), Cvoid, (Ptr{Cvoid},), p)

"""
`rollout` with the accepted steps of every period logged (the `ODESolution.t` / `.u` of src/simulate.jl:129-136, which
`plot_simulation!` reads): returns Xend, U, the step counts (N x iters), times (log_cap x N x iters) and extended states
(n_ext x log_cap x N x iters).
"""
function rollout_logged(ctx::Context, dnet::DeviceNetwork, pp::PrePost, plant::Integer, x0::Matrix{Float64}, Wd, t0, τ, T, iters;
                        n_dist::Integer, n_ctrl::Integer, abstol=1e-6, reltol=1e-3, log_cap::Integer=64)
    ns, N = size(x0); ne = ns + n_dist + n_ctrl
    Xend = Array{Float64}(undef, ns, N, iters); U = Array{Float64}(undef, n_ctrl, N, iters)
    ln = Matrix{Int32}(undef, N, iters); lt = Array{Float64}(undef, log_cap, N, iters); lu = Array{Float64}(undef, ne, log_cap, N, iters)
    check(ctx, ccall((:clr_rollout, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ref{PrePost}, Int32, Int32, Int32, Int32, Int64, Int32, Ptr{Float64}, Ptr{Float64},
                      Float64, Float64, Float64, Int32, Float64, Float64, Int32, Int32, Ptr{Float64}, Ptr{Float64},
                      Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int32),
                     ctx.h, dnet.h, pp, plant, ns, n_dist, n_ctrl, N, iters, x0, n_dist > 0 ? Wd : C_NULL,
                     t0, τ, T, 0, abstol, reltol, 4, 1, Xend, U, C_NULL, C_NULL, log_cap, ln, lt, lu, 0))
    return Xend, U, ln, lt, lu
end

"""
Block-cyclic shard of a split (`clr_deepz_boxgrid_cyclic`): local cell i is cell `cell_begin + (i ÷ block)·stride + i mod block`
(0-based ids).  Device r of G takes `cell_begin = r·block`, `stride = G·block`.
"""
function controller_forward_boxes_cyclic(ctx::Context, dnet::DeviceNetwork, g::BoxGrid, pp::PrePost, m::Integer,
                                         cell_begin::Integer, cell_count::Integer, block::Integer, stride::Integer)
    lo = Matrix{Float64}(undef, m, cell_count); hi = similar(lo)
    check(ctx, ccall((:clr_deepz_boxgrid_cyclic, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int64, Ref{PrePost},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int32, Int32),
                     ctx.h, dnet.h, length(g.partition), g.partition, g.centers, g.radii, cell_begin, cell_count, block, stride, pp,
                     lo, hi, C_NULL, C_NULL, C_NULL, 0, 2))
    return lo, hi
end

"""
`apply(ZonotopeSplitter(gens, nparts), Z)` (src/splitters.jl:30-43) for a batch in the padded layout, on the device
(`clr_split_zonotopes`; `gens` 1-based like in Julia).  Returns the children (2^sum(nparts) per input, child-major) in the same layout.
"""
function split_zonotopes(ctx::Context, C::Matrix{Float64}, G::Array{Float64,3}, ncols::Vector{Int32}, gens::Vector{<:Integer},
                         nparts::Vector{<:Integer})
    n, N = size(C); cap = size(G, 2); H = sum(nparts)
    oC = Matrix{Float64}(undef, n, N << H); oG = Array{Float64}(undef, n, cap, N << H); oN = Vector{Int32}(undef, N << H)
    check(ctx, ccall((:clr_split_zonotopes, LIB), Int32,
                     (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Int32, Ptr{Int32}, Ptr{Int32},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32),
                     ctx.h, n, N, C, G, ncols, cap, length(gens), Int32.(gens .- 1), Int32.(nparts), oC, oG, oN, 0))
    return oC, oG, oN
end

"`apply(SignSplitter(), U)` (src/splitters.jl:60-71) for a batch of control intervals: pieces (lo, hi) and their parent index (1-based)."
function sign_split(ctx::Context, lo::Vector{Float64}, hi::Vector{Float64})
    N = length(lo)
    ol = Vector{Float64}(undef, 2N); oh = similar(ol); op = Vector{Int64}(undef, 2N); cnt = Ref{Int64}(0)
    check(ctx, ccall((:clr_sign_split, LIB), Int32,
                     (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ref{Int64}, Int32),
                     ctx.h, N, lo, hi, ol, oh, op, cnt, 0))
    k = cnt[]
    return ol[1:k], oh[1:k], op[1:k] .+ 1
end

function last_device_error(ctx::Context)
    code = Ref{Int32}(0)
    check(ctx, ccall((:clr_last_device_error, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), ctx.h, code))
    return code[]
end

# ---- tuning / diagnostics ------------------------------------------------------------------------------------------
synchronize(ctx::Context) = check(ctx, ccall((:clr_synchronize, LIB), Int32, (Ptr{Cvoid},), ctx.h))
stream(ctx::Context) = ccall((:clr_stream, LIB), Ptr{Cvoid}, (Ptr{Cvoid},), ctx.h)
launch_count(ctx::Context) = ccall((:clr_launch_count, LIB), Int64, (Ptr{Cvoid},), ctx.h)
set_tile_cells!(ctx::Context, cells::Integer) = check(ctx, ccall((:clr_set_tile_cells, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, cells))
set_reorder!(ctx::Context, on::Bool) = check(ctx, ccall((:clr_set_reorder, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, on))
set_tiny!(ctx::Context, on::Bool) = check(ctx, ccall((:clr_set_tiny, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, on))
set_profiling!(ctx::Context, on::Bool) = check(ctx, ccall((:clr_set_profiling, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, on))
function last_pass_info(ctx::Context)
    ms = zeros(3); retried = Ref{Int64}(0); big = Ref{Int64}(0)
    check(ctx, ccall((:clr_last_pass_info, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}, Ref{Int64}), ctx.h, ms, retried, big))
    return ms, retried[], big[]
end
function measure_fp64_peak(ctx::Context, ms::Real=20.0)
    out = Ref{Float64}(0.0)
    check(ctx, ccall((:clr_measure_fp64_peak, LIB), Int32, (Ptr{Cvoid}, Float64, Ref{Float64}), ctx.h, ms, out))
    return out[]
end
function device_alloc(ctx::Context, bytes::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:clr_device_alloc, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx.h, bytes, out))
    return out[]
end
device_free(ctx::Context, p::Ptr{Cvoid}) = check(ctx, ccall((:clr_device_free, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, p))
memcpy_h2d(ctx::Context, dst::Ptr{Cvoid}, src::Array) =
    GC.@preserve src check(ctx, ccall((:clr_memcpy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx.h, dst, pointer(src), sizeof(src)))
memcpy_d2h(ctx::Context, dst::Array, src::Ptr{Cvoid}) =
    GC.@preserve dst check(ctx, ccall((:clr_memcpy_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx.h, pointer(dst), src, sizeof(dst)))
# Page-lock a Julia array in place (cudaHostRegister): results are then written into it by one DMA at PCIe speed instead of going
# through the context's pinned staging buffers.  Register once per buffer, unregister before the array can be collected.
host_register(ctx::Context, a::Array) =
    GC.@preserve a check(ctx, ccall((:clr_host_register, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx.h, pointer(a), sizeof(a)))
host_unregister(ctx::Context, a::Array) =
    GC.@preserve a check(ctx, ccall((:clr_host_unregister, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, pointer(a)))

# ---- several GPUs behind one handle (clr_create_multi) ------------------------------------------------------------
mutable struct MultiContext
    h::Ptr{Cvoid}
    function MultiContext(devices::AbstractVector{<:Integer})
        out = Ref{Ptr{Cvoid}}(C_NULL)
        devs = Int32.(devices)
        rc = ccall((:clr_create_multi, LIB), Int32, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devs, length(devs), out)
        rc == 0 || error(unsafe_string(ccall((:clr_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        m = new(out[])
        finalizer(x -> (x.h == C_NULL || ccall((:clr_destroy_multi, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), m)
        return m
    end
end
Base.length(mc::MultiContext) = Int(ccall((:clr_multi_size, LIB), Int32, (Ptr{Cvoid},), mc.h))
device_context(mc::MultiContext, i::Integer) = ccall((:clr_multi_ctx, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Int32), mc.h, i - 1)   # raw clr_ctx* of device i
function mcheck(mc::MultiContext, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:clr_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), mc.h))
    (rc == -1 || rc == -3) ? throw(ArgumentError(msg)) : error("libclr_b200 ($rc): $msg")
end
mutable struct MultiNetwork
    h::Ptr{Cvoid}
    mc::MultiContext
end
function upload(mc::MultiContext, net::FeedforwardNetwork)
    Ws = [Matrix{Float64}(L.weights) for L in net.layers]; bs = [Vector{Float64}(L.bias) for L in net.layers]
    dims = Int32[size(Ws[1], 2); [size(W, 1) for W in Ws]]; acts = Int32[act_id(L.activation) for L in net.layers]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Ws bs begin
        Wp = [pointer(W) for W in Ws]; bp = [pointer(b) for b in bs]
        mcheck(mc, ccall((:clr_multi_net_upload, LIB), Int32,
                         (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ref{Ptr{Cvoid}}),
                         mc.h, length(Ws), dims, acts, Wp, bp, out))
    end
    mn = MultiNetwork(out[], mc)
    finalizer(x -> (x.h == C_NULL || ccall((:clr_multi_net_free, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), mn)
    return mn
end
"All cells of a split over every device of `mc` (block-cyclic shards by default), boxes in cell order, hull of all cells."
function controller_forward_boxes(mc::MultiContext, mn::MultiNetwork, g::BoxGrid, pp::PrePost, m::Integer; balance::Integer=2)
    N = length(g)
    lo = Matrix{Float64}(undef, m, N); hi = similar(lo); hl = Vector{Float64}(undef, m); hh = similar(hl)
    mcheck(mc, ccall((:clr_multi_deepz_boxgrid, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ref{PrePost},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Float64}),
                     mc.h, mn.h, length(g.partition), g.partition, g.centers, g.radii, 0, N, pp, lo, hi, hl, hh, C_NULL, balance,
                     C_NULL, C_NULL))
    return lo, hi, hl, hh
end
"`rollout` with the trajectories sharded over every device of `mc`."
function rollout(mc::MultiContext, mn::MultiNetwork, pp::PrePost, plant::Integer, x0::Matrix{Float64}, Wd, t0, τ, T, iters;
                 n_dist::Integer, n_ctrl::Integer, abstol=1e-6, reltol=1e-3)
    ns, N = size(x0)
    Xend = Array{Float64}(undef, ns, N, iters); U = Array{Float64}(undef, n_ctrl, N, iters)
    mcheck(mc, ccall((:clr_multi_rollout, LIB), Int32,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ref{PrePost}, Int32, Int32, Int32, Int32, Int64, Int32, Ptr{Float64}, Ptr{Float64},
                      Float64, Float64, Float64, Int32, Float64, Float64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                     mc.h, mn.h, pp, plant, ns, n_dist, n_ctrl, N, iters, x0, n_dist > 0 ? Wd : C_NULL, t0, τ, T, 0, abstol, reltol,
                     4, 1, Xend, U, C_NULL, C_NULL))
    return Xend, U
end

end # module
