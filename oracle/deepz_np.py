"""NumPy restatement of the reference's set-propagation hot path (TEST INFRASTRUCTURE).

Follows, per function, the reference call site in /root/reference and the
upstream (un-vendored) algorithm it dispatches to.  See ``oracle/__init__.py``
for the "parity unpinned" statement.

Reference call chain restated here
  src/solve.jl:210-218      controller_forward: pre -> forward(DeepZ) -> post -> 1-D collapse
  src/splitters.jl:20-28    BoxSplitter -> LazySets.split(box_approximation(X), partition)
  src/problem.jl:10-58      apply(postprocessing, X::LazySet)
  src/setops.jl:76-83       box_approximation(U0) -- all the default plant step consumes
  src/setops.jl:84-86       _reduce_order(Z0, 2; force_reduction=true) -- TaylorModelReconstructor(box=false)

Upstream algorithms (recalled; [UPSTREAM-RECALL] in SURVEY.md 8c)
  LazySets  AbstractZonotope: linear_map (incl. _linear_map_zonotope_1D), low/high
  LazySets  overapproximate(::Rectification{<:AbstractZonotope}, Zonotope)
  NeuralNetworkReachability  ForwardAlgorithms/DeepZ.jl  (_overapproximate_zonotope)
  ReachabilityBase.Comparison  _leq/_isapprox/isapproxzero (rtol=sqrt(eps), ztol=10eps, atol=0)
  ReachabilityBase.Arrays      remove_zero_columns
  LazySets  AbstractHyperrectangle: split(H, num_blocks), genmat (flat dims dropped)
  LazySets  reduce_order(Z, r, GIR05()): _weighted_gens!, _interval_hull, _hcat_KLred

A network is a list of layers ``(W, b, act)`` with ``W`` of shape (m, n), ``b`` of
shape (m,) and ``act`` in {"Id", "ReLU", "Sigmoid", "Tanh"}.
A zonotope is ``(c, G)`` with ``c`` of shape (n,), ``G`` of shape (n, p).
"""
from __future__ import annotations

import itertools
import math
from fractions import Fraction

import numpy as np

EPS = np.finfo(np.float64).eps
RTOL = math.sqrt(EPS)          # Base.rtoldefault(Float64)
ZTOL = 10.0 * EPS              # ReachabilityBase.Comparison ABSZTOL(Float64)

ACT_IDS = {"Id": 0, "ReLU": 1, "Sigmoid": 2, "Tanh": 3}


# --------------------------------------------------------------------------------------
# ReachabilityBase.Comparison
# --------------------------------------------------------------------------------------
def isapproxzero(x: float) -> bool:
    return abs(x) <= ZTOL


def _isapprox(x: float, y: float) -> bool:
    if isapproxzero(x) and isapproxzero(y):
        return True
    # Base.isapprox with atol = 0
    return x == y or abs(x - y) <= RTOL * max(abs(x), abs(y))


def _leq(x: float, y: float) -> bool:
    return x <= y or _isapprox(x, y)


# --------------------------------------------------------------------------------------
# LazySets zonotope primitives
# --------------------------------------------------------------------------------------
def low_high(c: np.ndarray, G: np.ndarray):
    """low(Z, i) / high(Z, i): sequential sums over the generators starting at the centre."""
    lo = c.astype(np.float64).copy()
    hi = c.astype(np.float64).copy()
    for j in range(G.shape[1]):
        a = np.abs(G[:, j])
        lo -= a
        hi += a
    return lo, hi


def box_of(c: np.ndarray, G: np.ndarray):
    """box_approximation(Z): radius = sum(abs, G; dims=2); returns (low, high)."""
    r = np.zeros_like(c, dtype=np.float64)
    for j in range(G.shape[1]):
        r += np.abs(G[:, j])
    return c - r, c + r


def remove_zero_columns(G: np.ndarray) -> np.ndarray:
    if G.shape[1] == 0:
        return G
    keep = np.any(G != 0.0, axis=0)
    return G if keep.all() else G[:, keep]


def hyperrectangle_to_zonotope(center: np.ndarray, radius: np.ndarray):
    """convert(Zonotope, H): generators = radius_i * e_i for the non-flat dimensions."""
    center = np.asarray(center, dtype=np.float64)
    radius = np.asarray(radius, dtype=np.float64)
    idx = [i for i in range(len(radius)) if radius[i] != 0.0]
    G = np.zeros((len(center), len(idx)))
    for j, i in enumerate(idx):
        G[i, j] = radius[i]
    return center.copy(), G


def linear_map(W: np.ndarray, c: np.ndarray, G: np.ndarray):
    """LazySets.linear_map(M, Z::AbstractZonotope) incl. the one-row special case."""
    W = np.asarray(W, dtype=np.float64)
    if W.shape[0] == 1:
        # _linear_map_zonotope_1D: a single generator sum_j |W[1,:] . g_j|
        c2 = W @ c
        gi = 0.0
        for j in range(G.shape[1]):
            s = 0.0
            for i in range(G.shape[0]):
                s += W[0, i] * G[i, j]
            gi += abs(s)
        return c2, np.array([[gi]])
    return W @ c, W @ G


def affine_map(W, b, c, G):
    c2, G2 = linear_map(W, c, G)
    return c2 + np.asarray(b, dtype=np.float64), G2


# --------------------------------------------------------------------------------------
# activation relaxations
# --------------------------------------------------------------------------------------
def relu_zonotope(c: np.ndarray, G: np.ndarray, stats=None):
    """overapproximate(Rectification(Z), Zonotope) (LazySets)."""
    c = c.copy()
    G = G.copy()
    n, m = G.shape
    lo, hi = low_high(c, G)
    rows, mus = [], []
    for i in range(n):
        lb, ub = lo[i], hi[i]
        if stats is not None and (abs(lb) < 1e-12 or abs(ub) < 1e-12):
            stats["near_threshold"] = stats.get("near_threshold", 0) + 1
        if not _leq(lb, 0.0):
            continue
        if _leq(ub, 0.0):
            c[i] = 0.0
            G[i, :] = 0.0
        else:
            lam = ub / (ub - lb)
            mu = -lam * lb / 2
            c[i] = c[i] * lam + mu
            G[i, :] = G[i, :] * lam
            rows.append(i)
            mus.append(mu)
    if rows:
        Gnew = np.zeros((n, len(rows)))
        for j, i in enumerate(rows):
            Gnew[i, j] = mus[j]
        G = np.hstack([G, Gnew])
    if stats is not None:
        stats.setdefault("unstable", []).append(len(rows))
    return c, remove_zero_columns(G)


def _sigmoid(x: float) -> float:
    return 1.0 / (1.0 + math.exp(-x))


def _sigmoid_prime(x: float) -> float:
    ex = math.exp(x)
    return ex / (1.0 + ex) ** 2


def _tanh_prime(x: float) -> float:
    return 1.0 - math.tanh(x) ** 2


def sigtanh_zonotope(c, G, act: str, stats=None):
    """NeuralNetworkReachability _overapproximate_zonotope(Z, act, act')."""
    f, fp = (_sigmoid, _sigmoid_prime) if act == "Sigmoid" else (math.tanh, _tanh_prime)
    c = c.copy()
    G = G.copy()
    n, m = G.shape
    lo, hi = low_high(c, G)
    rows, mus = [], []
    for i in range(n):
        lx, ux = float(lo[i]), float(hi[i])
        ly, uy = f(lx), f(ux)
        if _isapprox(lx, ux):
            c[i] = uy
            G[i, :] = 0.0
        else:
            lam = min(fp(lx), fp(ux))
            mu1 = (uy + ly - lam * (ux + lx)) / 2
            mu2 = (uy - ly - lam * (ux - lx)) / 2
            c[i] = c[i] * lam + mu1
            G[i, :] = G[i, :] * lam
            rows.append(i)
            mus.append(mu2)
    if rows:
        Gnew = np.zeros((n, len(rows)))
        for j, i in enumerate(rows):
            Gnew[i, j] = mus[j]
        G = np.hstack([G, Gnew])
    if stats is not None:
        stats.setdefault("unstable", []).append(len(rows))
    return c, remove_zero_columns(G)


def activation_zonotope(c, G, act, stats=None):
    if act == "Id":
        if stats is not None:
            stats.setdefault("unstable", []).append(0)
        return c, G
    if act == "ReLU":
        return relu_zonotope(c, G, stats)
    if act in ("Sigmoid", "Tanh"):
        return sigtanh_zonotope(c, G, act, stats)
    raise ValueError(f"unknown activation {act}")


# --------------------------------------------------------------------------------------
# forward(X, net, DeepZ())      (src/solve.jl:212)
# --------------------------------------------------------------------------------------
def deepz_forward(net, c, G, stats=None, layers_out=None):
    """Propagate the zonotope (c, G) through all layers; returns the output zonotope."""
    c = np.asarray(c, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64).reshape(len(c), -1)
    for (W, b, act) in net:
        if stats is not None:
            stats.setdefault("p_in", []).append(G.shape[1])
        c, G = affine_map(W, b, c, G)
        c, G = activation_zonotope(c, G, act, stats)
        if layers_out is not None:
            layers_out.append((c.copy(), G.copy()))
    return c, G


def apply_post_set(post, c, G):
    """apply(postprocessing, X::LazySet)  (src/problem.jl:10, 28-30, 40-42, 52-58).

    post = None | ("shift", s) | ("scale", a) | ("matrix", M) | ("project", dims0)
    """
    if post is None:
        return c, G
    kind, val = post
    if kind == "shift":
        return c + float(val), G
    if kind == "scale":
        return float(val) * c, float(val) * G
    if kind == "matrix":
        return linear_map(np.asarray(val, dtype=np.float64), c, G)
    if kind == "project":
        d = list(val)
        return c[d], remove_zero_columns(G[d, :])
    raise ValueError(kind)


def apply_post_point(post, u):
    if post is None:
        return u
    kind, val = post
    if kind == "shift":
        return u + float(val)
    if kind == "scale":
        return float(val) * u
    if kind == "matrix":
        return np.asarray(val, dtype=np.float64) @ u
    if kind == "project":
        return u[list(val)]
    raise ValueError(kind)


def apply_pre_set(pre, c, G):
    """Affine preprocessing x -> A x + d (ACC: models/ACC/ACC.jl:81-99) on a zonotope."""
    if pre is None:
        return c, G
    A, d = pre
    A = np.asarray(A, dtype=np.float64)
    return A @ c + np.asarray(d, dtype=np.float64), A @ G


def apply_pre_point(pre, x):
    if pre is None:
        return x
    A, d = pre
    return np.asarray(A, dtype=np.float64) @ x + np.asarray(d, dtype=np.float64)


def controller_forward(net, c, G, pre=None, post=None, stats=None):
    """controller_forward (src/solve.jl:210-218) -> output zonotope and its box (lo, hi)."""
    c, G = apply_pre_set(pre, c, G)
    c, G = deepz_forward(net, c, G, stats)
    c, G = apply_post_set(post, c, G)
    lo, hi = box_of(c, G)
    return c, G, lo, hi


# --------------------------------------------------------------------------------------
# pointwise forward(x, net)      (src/simulate.jl:104; ControllerFormats)
# --------------------------------------------------------------------------------------
def forward_point(net, x):
    x = np.asarray(x, dtype=np.float64)
    for (W, b, act) in net:
        y = np.asarray(W, dtype=np.float64) @ x + np.asarray(b, dtype=np.float64)
        if act == "ReLU":
            y = np.maximum(y, 0.0)
        elif act == "Sigmoid":
            y = 1.0 / (1.0 + np.exp(-y))
        elif act == "Tanh":
            y = np.tanh(y)
        x = y
    return x


# --------------------------------------------------------------------------------------
# BoxSplitter -> LazySets.split(H, num_blocks)       (src/splitters.jl:20-28)
# --------------------------------------------------------------------------------------
def _rat(x: float):
    """Julia Base.rat (twiceprecision.jl): small rational approximation of a Float64."""
    y = x
    a = d = 1
    b = c = 0
    m = 16777216  # maxintfloat(Float32)
    while abs(y) <= m:
        f = int(y)  # trunc
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > m:
            return c, d
        if b != 0 and a / b == x:
            break
        if y == 0.0:
            break
        y = 1.0 / y
    return a, b


def julia_range(start: float, step: float, length: int) -> np.ndarray:
    """range(start; step=step, length=length) for Float64 -- best-effort emulation.

    Julia rationalises start and step (Base.range_start_step_length); when that
    succeeds every element is the correctly rounded value of the exact rational
    (start_n + j*step_n)/den, otherwise elements are start + j*step evaluated in
    twice precision, which we model with exact rational arithmetic as well.
    """
    sn, sd = _rat(start)
    tn, td = _rat(step)
    if sd != 0 and td != 0 and sn / sd == start and tn / td == step:
        fs, ft = Fraction(sn, sd), Fraction(tn, td)
    else:
        fs, ft = Fraction(start), Fraction(step)
    return np.array([float(fs + j * ft) for j in range(length)], dtype=np.float64)


def split_tables(lo, hi, partition):
    """Per-dimension centre tables and the shared radius of split(H, num_blocks).

    H = Hyperrectangle(low=lo, high=hi): center = (lo+hi)/2, radius = (hi-lo)/2.
    Returns (centers: list of 1-D arrays, radius: array).
    """
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    c = (lo + hi) / 2
    R = (hi - lo) / 2
    lo_h = c - R  # low(H)
    centers, radius = [], np.empty_like(R)
    for i, m in enumerate(partition):
        m = int(m)
        if m <= 0:
            raise ValueError("each dimension needs at least one block")
        if m == 1:
            centers.append(np.array([lo_h[i] + R[i]]))
            radius[i] = R[i]
        else:
            radius[i] = R[i] / m
            centers.append(julia_range(lo_h[i] + radius[i], 2 * radius[i], m))
    return centers, radius


def split_box(lo, hi, partition):
    """All cells in Iterators.product order (dimension 1 fastest): (centers (N, n), radius (n,))."""
    centers, radius = split_tables(lo, hi, partition)
    n = len(centers)
    # itertools.product varies the LAST factor fastest; Julia's product the FIRST
    cells = np.array([tuple(reversed(t)) for t in itertools.product(*reversed(centers))],
                     dtype=np.float64).reshape(-1, n)
    return cells, radius


def cell_multi_index(cell_id: int, partition):
    """cell id (0-based, product order) -> per-dimension block index (0-based)."""
    idx = []
    for m in partition:
        idx.append(cell_id % int(m))
        cell_id //= int(m)
    return idx


def deepz_boxgrid(net, lo, hi, partition, pre=None, post=None, cell_range=None, stats=None):
    """The batched hot loop (src/solve.jl:153-166): every cell through controller_forward.

    Returns out_lo, out_hi of shape (N, m).
    """
    cells, radius = split_box(lo, hi, partition)
    if cell_range is not None:
        cells = cells[cell_range[0]:cell_range[1]]
    out_lo, out_hi = [], []
    for cc in cells:
        c, G = hyperrectangle_to_zonotope(cc, radius)
        st = {} if stats is not None else None
        _, _, l, h = controller_forward(net, c, G, pre, post, st)
        if stats is not None:
            stats.setdefault("cells", []).append(st)
        out_lo.append(l)
        out_hi.append(h)
    return np.array(out_lo), np.array(out_hi)


def algorithmic_flops(net, p_in):
    """F = sum_l 2*m_l*n_l*(p_l + 1), p_l = generator columns entering layer l."""
    return sum(2 * W.shape[0] * W.shape[1] * (p + 1) for (W, _, _), p in zip(net, p_in))


# --------------------------------------------------------------------------------------
# ZonotopeSplitter -> LazySets.split(Z, gens, nparts)       (src/splitters.jl:30-43) [UPSTREAM-RECALL]
# --------------------------------------------------------------------------------------
def split_zonotope_once(c, G, j):
    """split(Z, j): (Zonotope(c - g_j/2, G'), Zonotope(c + g_j/2, G')) with G' = G, column j halved."""
    G2 = G.copy()
    G2[:, j] = G[:, j] / 2
    return (c - G2[:, j], G2), (c + G2[:, j], G2)


def split_zonotope(c, G, gens, nparts):
    """Recursive restatement: every round replaces each zonotope by its two halves, in order."""
    Zs = [(np.asarray(c, dtype=np.float64), np.asarray(G, dtype=np.float64))]
    for g, times in zip(gens, nparts):
        for _ in range(times):
            nxt = []
            for (cc, GG) in Zs:
                a, b = split_zonotope_once(cc, GG, g)
                nxt += [a, b]
            Zs = nxt
    return Zs


# --------------------------------------------------------------------------------------
# TaylorModelReconstructor(box=false): order reduction of U0      (src/setops.jl:84-101) [UPSTREAM-RECALL]
# --------------------------------------------------------------------------------------
def reduce_order_gir05(c, G, r=2):
    """LazySets reduce_order(Z, r, GIR05()) with ReachabilityAnalysis' force_reduction=true.

    weights_j = norm(g_j, 1) - norm(g_j, Inf); indices = sortperm(weights; rev=true) (stable); the first
    m*(r-1) generators are kept, the others are replaced by their interval hull, whose diagonal entry i is
    the sum of |G[i, j]| over the boxed generators in SORTED order (the loop of _interval_hull).
    Returns (c, [G[:, kept] | diag(d)]) with exactly m*r generators; r == 1 boxes everything unsorted.
    A zonotope with fewer than m*(r-1) generators raises IndexError (the reference: BoundsError).
    """
    c = np.asarray(c, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64)
    m, p = G.shape
    kept = m * (r - 1)
    if p < kept:
        raise IndexError(f"order-{r} reduction of a zonotope with {p} < {kept} generators")
    if kept == 0:
        order = list(range(p))
    else:
        w = []
        for j in range(p):
            s1 = 0.0
            mx = 0.0
            for i in range(m):
                a = abs(float(G[i, j]))
                s1 += a
                mx = max(mx, a)
            w.append(s1 - mx)
        order = sorted(range(p), key=lambda j: -w[j])      # Python's sort is stable, like Julia's sortperm
    d = np.zeros(m)
    for i in range(m):
        acc = 0.0
        for j in order[kept:]:
            acc += abs(float(G[i, j]))
        d[i] = acc
    return c.copy(), np.hstack([G[:, order[:kept]], np.diag(d)])


# --------------------------------------------------------------------------------------
# _project_oa(R::AbstractTaylorModelReachSet, vars, t)            (src/setops.jl:9-12) [UPSTREAM-RECALL]
#   = project(set(overapproximate(R, Zonotope, t)), vars; remove_zero_generators)
# ReachabilityAnalysis overapproximate(::TaylorModelReachSet, Zonotope, t): evaluate the TaylorModel1 components at
# t - t0 (TaylorSeries Horner, plain Float64), TaylorModelN over the symmetric box, fp_rpa, then LazySets
# overapproximate(::Vector{<:TaylorModelN}, Zonotope): constant + linear terms -> centre / generators, the nonlinear
# terms and the remainder -> an interval whose midpoint moves the centre and whose radius is a generator r_v e_v.
# Interval additions round outwards exactly (directed rounding via TwoSum); multiplications by 0 / +-1 are exact.
# LazySets' overapproximate(::Vector{TaylorModel1}, Zonotope; remove_redundant_generators=true) then merges parallel
# generators (remove_redundant_generators below) before the projection.
# --------------------------------------------------------------------------------------
def _two_sum(a: float, b: float):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def _add_dn(a: float, b: float) -> float:
    """a + b rounded towards -inf (exact directed rounding through the error-free TwoSum)."""
    s, e = _two_sum(a, b)
    return math.nextafter(s, -math.inf) if e < 0.0 else s


def _add_up(a: float, b: float) -> float:
    s, e = _two_sum(a, b)
    return math.nextafter(s, math.inf) if e > 0.0 else s


def monomial_table(n: int, order: int):
    """Exponent table in TaylorSeries' coefficient order: by total degree, and inside a degree the order of
    TaylorSeries.coeff_table (descending lexicographic: x1^d first).  Shape (M, n)."""
    out = []
    for d in range(order + 1):
        def rec(prefix, left, k):
            if k == n - 1:
                out.append(prefix + [left]); return
            for e in range(left, -1, -1):
                rec(prefix + [e], left - e, k + 1)
        rec([], d, 0)
    return np.array(out, dtype=np.int32)


def ismultiple(u, v):
    """ReachabilityBase.Arrays.ismultiple(u, v) [UPSTREAM-RECALL]: (True, factor) if u = factor * v entry by entry (entries
    compared with isapproxzero / _isapprox), (True, 0) if both are zero, else (False, 0)."""
    no_factor, factor = True, 0.0
    for ui, vi in zip(u, v):
        if isapproxzero(ui):
            if not isapproxzero(vi):
                return False, 0.0
            continue
        if isapproxzero(vi):
            return False, 0.0
        if no_factor:
            no_factor, factor = False, ui / vi
        elif not _isapprox(factor, ui / vi):
            return False, 0.0
    return True, (0.0 if no_factor else factor)


def remove_redundant_generators(G):
    """LazySets.remove_redundant_generators(Z) on the generator matrix [UPSTREAM-RECALL]: zero columns are dropped; then, column
    by column, every later column that is a multiple of the (already accumulated) column is added to it (subtracted when the
    factor is negative) and dropped."""
    G = np.asarray(G, dtype=np.float64)
    if G.shape[1] <= 1:
        return G
    G = remove_zero_columns(G)
    p = G.shape[1]
    done = [False] * p
    cols = []
    for j1 in range(p):
        if done[j1]:
            continue
        g1 = G[:, j1].copy()
        for j2 in range(j1 + 1, p):
            if done[j2]:
                continue
            ok, f = ismultiple(g1, G[:, j2])
            if ok:
                g1 = g1 + G[:, j2] if f > 0.0 else g1 - G[:, j2]
                done[j2] = True
        cols.append(g1)
    return np.stack(cols, axis=1) if cols else np.zeros((G.shape[0], 0))


def project_oa_taylor_model(exps, coeffs, rem, dt, vars_keep, remove_zero_generators=True, remove_redundant=False):
    """One reach-set.  exps (M, n); coeffs (n, KT, M); rem (n, 2); dt scalar; vars_keep 0-based.
    Returns (c (n_keep,), G (n_keep, p))."""
    exps = np.asarray(exps); n = exps.shape[1]
    coeffs = np.asarray(coeffs, dtype=np.float64); rem = np.asarray(rem, dtype=np.float64)
    KT, M = coeffs.shape[1], coeffs.shape[2]
    deg = exps.sum(axis=1)
    c = np.zeros(n); Glin = np.zeros((n, n)); rad = np.zeros(n)
    for v in range(n):
        lo = hi = 0.0
        c0 = 0.0
        for m in range(M):
            s = float(coeffs[v, KT - 1, m])
            for k in range(KT - 2, -1, -1):
                s = s * float(dt) + float(coeffs[v, k, m])          # TaylorSeries evaluate: suma = suma*dt + a[k]
            if deg[m] == 0:
                c0 = s
            elif deg[m] == 1:
                Glin[v, int(np.argmax(exps[m]))] = s
            elif np.all(exps[m] % 2 == 0):                           # s * [0, 1]
                lo = _add_dn(lo, min(s, 0.0)); hi = _add_up(hi, max(s, 0.0))
            else:                                                    # s * [-1, 1]
                lo = _add_dn(lo, -abs(s)); hi = _add_up(hi, abs(s))
        i0lo, i0hi = _add_dn(c0, float(rem[v, 0])), _add_up(c0, float(rem[v, 1]))
        c0m = 0.5 * (i0lo + i0hi)                                    # fp_rpa
        rlo, rhi = _add_dn(i0lo, -c0m), _add_up(i0hi, -c0m)
        lo, hi = _add_dn(lo, rlo), _add_up(hi, rhi)
        c[v] = c0m + 0.5 * (lo + hi)
        rad[v] = 0.5 * _add_up(hi, -lo)
    G = np.hstack([Glin, np.diag(rad)])
    if remove_redundant:                                             # default of LazySets' overapproximate: on the FULL zonotope
        G = remove_redundant_generators(G)
    vk = list(vars_keep)
    c, G = c[vk], G[vk, :]                                           # LazySets.project
    if remove_zero_generators:
        G = remove_zero_columns(G)
    return c, G


def eval_taylor_model_point(exps, coeffs, dt, xi):
    """Value of the polynomial part at the normalised point xi (test helper for the soundness property)."""
    exps = np.asarray(exps); coeffs = np.asarray(coeffs, dtype=np.float64)
    n, KT, M = coeffs.shape
    mono = np.prod(np.power(np.asarray(xi, dtype=np.float64)[None, :], exps), axis=1)       # (M,)
    tpow = float(dt) ** np.arange(KT)
    return np.einsum("vkm,k,m->v", coeffs, tpow, mono)
