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
        finalizer(c -> ccall((:clr_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

function check(ctx::Context, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:clr_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))
    # CLR_ERR_ARG (-1) / CLR_ERR_DIM (-3) are what src/solve.jl:88-99, :130-134 raise as ArgumentError
    (rc == -1 || rc == -3) ? throw(ArgumentError(msg)) : error("libclr_b200 ($rc): $msg")
end

act_id(::Id) = Int32(0); act_id(::ReLU) = Int32(1); act_id(::Sigmoid) = Int32(2); act_id(::Tanh) = Int32(3)

struct DeviceNetwork
    h::Ptr{Cvoid}
    net::FeedforwardNetwork
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
    return DeviceNetwork(out[], net)
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
        append!(cs, m == 1 ? [c[i]] : collect(range(c[i] - R[i] + r; step=2r, length=m)))   # LazySets' split
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

end # module
