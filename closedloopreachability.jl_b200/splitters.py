"""Host side of the cell generators -- mirror of /root/reference/src/splitters.jl.

``BoxSplitter(partition)`` is ``split(box_approximation(X), partition)`` (src/splitters.jl:20-28;
LazySets ``split(::AbstractHyperrectangle, num_blocks)``).  The reference materialises one
``Hyperrectangle`` object per cell; here a split is an array-backed ``BoxGrid``: per-dimension
centre / radius tables plus the partition.  The cells themselves are generated on the device from
the cell index (``Iterators.product`` order, dimension 1 fastest), so 10^6 cells cost 10^6 indices,
not 10^6 host objects.

Two partition forms, as in LazySets:
  * integers ``[m_1, ..., m_n]``: m_i equal blocks; the block centres are the Julia range
    ``range(lo_i + r_i; step = 2 r_i, length = m_i)`` with r_i = R_i / m_i, whose elements Julia
    evaluates in twice precision -- emulated below with exact rationals;
  * cut points ``[[c_11, c_12, ...], ...]`` (models/InvertedPendulum/InvertedPendulum.jl:210):
    dimension i is cut at the given coordinates.
"""
from __future__ import annotations

from dataclasses import dataclass
from fractions import Fraction
from typing import List, Optional, Sequence

import numpy as np


def _rat(x: float):
    """Base.rat (twiceprecision.jl): small-denominator rational approximation used by range()."""
    y = x
    a = d = 1
    b = c = 0
    m = 16777216  # maxintfloat(Float32)
    while abs(y) <= m:
        f = int(y)
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
    """Elements of ``range(start; step, length)`` for Float64 arguments (correctly rounded rationals)."""
    sn, sd = _rat(start)
    tn, td = _rat(step)
    if sd != 0 and td != 0 and sn / sd == start and tn / td == step:
        fs, ft = Fraction(sn, sd), Fraction(tn, td)
    else:
        fs, ft = Fraction(start), Fraction(step)
    return np.array([float(fs + j * ft) for j in range(length)], dtype=np.float64)


@dataclass
class BoxGrid:
    """The cells of one box split, array backed."""
    partition: np.ndarray          # (n,) int32 blocks per dimension
    centers: List[np.ndarray]      # per dimension: block centres
    radii: List[np.ndarray]        # per dimension: block radii

    @property
    def dim(self) -> int:
        return len(self.partition)

    def __len__(self) -> int:
        n = 1
        for m in self.partition:
            n *= int(m)
        return n

    def multi_index(self, cell: int) -> List[int]:
        idx = []
        for m in self.partition:
            idx.append(cell % int(m))
            cell //= int(m)
        return idx

    def cell(self, i: int):
        """(center, radius) of cell i -- what ``apply(splitter, X)[i+1]`` is in the reference."""
        k = self.multi_index(i)
        return (np.array([self.centers[d][k[d]] for d in range(self.dim)]),
                np.array([self.radii[d][k[d]] for d in range(self.dim)]))

    def shard(self, rank: int, world: int):
        """Contiguous cell range of one GPU: [begin, begin+count)."""
        n = len(self)
        b = (n * rank) // world
        e = (n * (rank + 1)) // world
        return b, e - b


def split_box(lo, hi, partition) -> BoxGrid:
    """LazySets ``split(Hyperrectangle(low=lo, high=hi), partition)`` as a BoxGrid."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    if lo.shape != hi.shape or lo.ndim != 1:
        raise ValueError("lo and hi must be vectors of equal length")
    if len(partition) != len(lo):
        raise ValueError("the partition needs one entry per dimension")
    c = (lo + hi) / 2
    R = (hi - lo) / 2
    lo_h = c - R
    hi_h = c + R
    centers, radii, part = [], [], []
    for i, m in enumerate(partition):
        if np.ndim(m) == 0:
            m = int(m)
            if m <= 0:
                raise ValueError("each dimension needs at least one block")
            if m == 1:
                centers.append(np.array([lo_h[i] + R[i]]))
                radii.append(np.array([R[i]]))
            else:
                r = R[i] / m
                centers.append(julia_range(lo_h[i] + r, 2 * r, m))
                radii.append(np.full(m, r))
            part.append(m)
        else:
            cuts = [float(x) for x in m]
            pts = [lo_h[i]] + cuts + [hi_h[i]]
            if any(b < a for a, b in zip(pts[:-1], pts[1:])):
                raise ValueError("cut points must be increasing and inside the box")
            centers.append(np.array([(a + b) / 2 for a, b in zip(pts[:-1], pts[1:])]))
            radii.append(np.array([(b - a) / 2 for a, b in zip(pts[:-1], pts[1:])]))
            part.append(len(pts) - 1)
    return BoxGrid(np.array(part, dtype=np.int32), centers, radii)


class Splitter:
    """``Splitter`` of the reference: fires at control step k == 1 only (src/splitters.jl:11-13)."""

    def __init__(self, split_fun):
        self.split_fun = split_fun

    def apply(self, lo, hi):
        return self.split_fun(lo, hi)

    def haskey(self, k: int) -> bool:
        return k == 1

    def __getitem__(self, k: int):
        if k != 1:
            raise KeyError(f"key {k} not found")
        return self


def NoSplitter() -> Splitter:
    return Splitter(lambda lo, hi: split_box(lo, hi, [1] * len(lo)))


def BoxSplitter(partition: Optional[Sequence] = None) -> Splitter:
    if partition is None:   # default: one split per dimension (src/splitters.jl:21-23)
        return Splitter(lambda lo, hi: split_box(lo, hi, [2] * len(lo)))
    return Splitter(lambda lo, hi: split_box(lo, hi, partition))


class IndexedSplitter:
    """Per-step splitters (src/splitters.jl:49-54): ``IndexedSplitter({k: splitter})``."""

    def __init__(self, index2splitter):
        self.index2splitter = dict(index2splitter)

    def haskey(self, k: int) -> bool:
        return k in self.index2splitter

    def __getitem__(self, k: int):
        return self.index2splitter[k]


@dataclass
class ZonotopeCells:
    """A batch of zonotope cells in the layout clr_deepz_zonotopes takes: centres (N, n), generator offsets (N+1,),
    generators (sum p, n) -- row j is generator j."""
    C: np.ndarray
    col_off: np.ndarray
    G: np.ndarray

    def __len__(self) -> int:
        return self.C.shape[0]

    def cell(self, i: int):
        return self.C[i], self.G[self.col_off[i]:self.col_off[i + 1]].T


def split_zonotope(c, G, gens, nparts) -> ZonotopeCells:
    """LazySets ``split(Z, gens, nparts)``: generator ``gens[i]`` (0-based here) is halved ``nparts[i]`` times; every
    halving replaces each zonotope Z by (Z - g/2, Z + g/2) with that generator halved, in this order, so the result
    has 2^sum(nparts) cells (src/splitters.jl:30-43 passes the user's arguments straight through)."""
    c = np.asarray(c, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64).reshape(len(c), -1)
    if len(gens) != len(nparts):
        raise ValueError("gens and nparts must have the same length")
    cs = c[None, :].copy()
    Gc = G.copy()                      # all cells share the generator matrix, only the centres differ
    for g, times in zip(gens, nparts):
        if not (0 <= g < Gc.shape[1]):
            raise ValueError(f"generator index {g} out of range")
        for _ in range(int(times)):
            half = Gc[:, g] / 2
            cs = np.stack([cs - half, cs + half], axis=1).reshape(-1, len(c))   # (Z1a, Z1b, Z2a, Z2b, ...)
            Gc[:, g] = half
    N, p = cs.shape[0], Gc.shape[1]
    return ZonotopeCells(np.ascontiguousarray(cs), np.arange(N + 1, dtype=np.int64) * p,
                         np.ascontiguousarray(np.tile(Gc.T, (N, 1))))


class ZonotopeSplitterFun:
    def __init__(self, generators=None, splits=None):
        self.generators, self.splits = generators, splits

    def __call__(self, c, G):
        G = np.asarray(G, dtype=np.float64).reshape(len(c), -1)
        if self.generators is None and self.splits is None:       # default: one split per generator
            return split_zonotope(c, G, list(range(G.shape[1])), [1] * G.shape[1])
        return split_zonotope(c, G, self.generators, self.splits)


def ZonotopeSplitter(generators=None, splits=None) -> Splitter:
    """``ZonotopeSplitter`` of the reference (src/splitters.jl:30-43); ``apply(c, G)`` returns ZonotopeCells."""
    return Splitter(ZonotopeSplitterFun(generators, splits))


class SignSplitter:
    """``SignSplitter`` (src/splitters.jl:60-71): an interval containing 0 in its interior is cut at 0."""

    def apply(self, lo: float, hi: float):
        if lo < 0.0 and hi > 0.0:
            return [(lo, 0.0), (0.0, hi)]
        return [(lo, hi)]
