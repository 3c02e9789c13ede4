"""Breadth-first ``_solve`` -- host mirror of /root/reference/src/solve.jl:110-208 for the controller part.

The reference walks the control periods depth-first: every flowpipe of period k is pushed on a waiting list and the
LAST one is popped first (``pop!``, src/solve.jl:179), its successors of period k + 1 are appended to the output and pushed
(src/solve.jl:197-204).  Each pop costs one ``controller_forward`` per split cell (src/solve.jl:187-195) -- the per-cell
loop this library replaces.  Here all live chains of a period go through the controller in ONE ``clr_deepz_*`` call
(breadth-first, level synchronous), and ``lifo_emission_order`` then puts the flowpipes back into exactly the order the
reference's ``MixedFlowpipe`` has, so that downstream code (``sol[i]``, plotting, predicates over ``sol``) sees the same
sequence.  The plant step (``post`` with TMJets + ``_project_oa``, src/solve.jl:181-184, :236) stays on the reference's CPU
path: it is a callback here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .splitters import BoxGrid, SignSplitter, ZonotopeCells, split_box


def lifo_emission_order(parents: Sequence) -> List[Tuple[int, int]]:
    """Order in which the reference's ``_solve`` appends flowpipes (src/solve.jl:168-204).

    ``parents[0]`` is the number of first-period flowpipes (in cell order); ``parents[k]`` (k >= 1) holds, for every
    flowpipe of period k + 1 in breadth-first order (parents in breadth-first order, children in the order
    ``_solve_one`` returns them), the breadth-first index of its parent in period k.  Returns ``(level, index)`` pairs,
    level 0-based.  Flowpipes of the last period are appended but never pushed (``k < length(tvec) - 1``).
    """
    n_levels = len(parents)
    n0 = int(parents[0])
    children: List[List[List[int]]] = []
    sizes = [n0]
    for k in range(1, n_levels):
        par = np.asarray(parents[k], dtype=np.int64)
        if par.size and (par.min() < 0 or par.max() >= sizes[-1]):
            raise ValueError(f"parent index out of range at level {k}")
        if par.size > 1 and np.any(np.diff(par) < 0):
            raise ValueError("children must be listed parent by parent (breadth-first order)")
        ch: List[List[int]] = [[] for _ in range(sizes[-1])]
        for j, p in enumerate(par):
            ch[int(p)].append(j)
        children.append(ch)
        sizes.append(int(par.size))
    order = [(0, i) for i in range(n0)]
    stack = [(0, i) for i in range(n0)] if n_levels > 1 else []
    while stack:
        k, i = stack.pop()                      # pop!(waiting_list): last in, first out
        for j in children[k][i]:
            order.append((k + 1, j))            # append!(flowpipes, Fs)
            if k + 1 < n_levels - 1:
                stack.append((k + 1, j))        # push!(waiting_list, WaitingListElement(F, k))
    return order


@dataclass
class Level:
    """All flowpipes of one control period, breadth-first."""
    k: int                     # control period, 1-based like the reference's k
    parent: np.ndarray         # breadth-first index of the parent flowpipe in period k - 1 (-1 in the first period)
    cell_c: np.ndarray         # (N, n) centre of the split cell the flowpipe started from
    cell_G: List[np.ndarray]   # per flowpipe: (n, p) generators of that cell
    U_lo: np.ndarray           # (N, m) control piece attached to the flowpipe (``sol.ext[:controls]``)
    U_hi: np.ndarray
    end_c: Optional[np.ndarray] = None      # (N, n) end-of-period state set (the stand-in for _project_oa(F(tend)))
    end_G: Optional[List[np.ndarray]] = None


def _box_of(c, G):
    r = np.abs(G).sum(axis=1) if G.size else np.zeros_like(c)
    return c - r, c + r


def _cells_of_box(lo, hi, splitter_k):
    """apply(splitter[k], X) for a box-producing splitter: a BoxGrid."""
    return splitter_k.apply(lo, hi)


def _grid_to_zonotopes(g: BoxGrid):
    """Cells of a BoxGrid as explicit zonotopes (Hyperrectangle -> Zonotope: diag(r) with flat dimensions dropped)."""
    N = len(g)
    C = np.empty((N, g.dim))
    Gs = []
    for i in range(N):
        c, r = g.cell(i)
        C[i] = c
        Gs.append(np.diag(r)[:, r != 0.0])
    return C, Gs


def _pack(C, Gs) -> ZonotopeCells:
    off = np.zeros(len(Gs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([G.shape[1] for G in Gs])
    n = C.shape[1]
    G = np.concatenate([G.T for G in Gs]) if off[-1] else np.zeros((0, n))
    return ZonotopeCells(np.ascontiguousarray(C), off, np.ascontiguousarray(G))


def solve_controller_bfs(controller_batch: Callable, X0_lo, X0_hi, periods: int, splitter, plant_step: Callable,
                         input_splitter=None) -> Tuple[List[Level], List[Tuple[int, int]]]:
    """Level-synchronous version of ``_solve`` (src/solve.jl:110-208).

    ``controller_batch(cells)`` propagates a whole batch of cells (``BoxGrid`` or ``ZonotopeCells``) through the
    controller -- one ``clr_deepz_boxgrid`` / ``clr_deepz_zonotopes`` call -- and returns ``(lo, hi)`` of shape (N, m).
    ``splitter`` follows the reference's protocol (``haskey(k)`` / ``[k].apply(lo, hi)``; a plain splitter fires at k = 1
    only, an ``IndexedSplitter`` where it has a key; a splitter always sees ``box_approximation(X0)``, src/splitters.jl:25).
    ``input_splitter`` (``None`` or ``SignSplitter``) is applied to EVERY control box of EVERY period (src/solve.jl:232).
    ``plant_step(k, cell_c, cell_G, U_lo, U_hi)`` maps the flowpipes of a period to their end-of-period state sets
    ``(end_c (N, n), end_G list)``: the reference's ``post`` + ``_project_oa``, which stay on its CPU path.

    Returns the levels (breadth-first) and the emission order of the reference (``lifo_emission_order``).
    """
    lo0, hi0 = np.asarray(X0_lo, dtype=np.float64), np.asarray(X0_hi, dtype=np.float64)
    levels: List[Level] = []
    parents: List = []
    for k in range(1, periods + 1):
        # ---- the cells of this period: X0s = haskey(splitter, k) ? apply(splitter[k], X0) : [X0] for every live chain ----
        par_cells: List[int] = []
        if k == 1:
            if splitter is not None and splitter.haskey(1):
                grid = _cells_of_box(lo0, hi0, splitter[1])
                cells = grid                                     # the device decodes the cells from the cell id
                C, Gs = _grid_to_zonotopes(grid)
            else:
                C, Gs = _grid_to_zonotopes(split_box(lo0, hi0, [1] * len(lo0)))
                cells = _pack(C, Gs)
            par_cells = [-1] * len(Gs)
        else:
            prev = levels[-1]
            Cl, Gs = [], []
            for i in range(prev.end_c.shape[0]):
                c, G = prev.end_c[i], prev.end_G[i]
                if splitter is not None and splitter.haskey(k):
                    blo, bhi = _box_of(c, G)
                    Ci, Gi = _grid_to_zonotopes(_cells_of_box(blo, bhi, splitter[k]))
                    Cl.extend(Ci); Gs.extend(Gi); par_cells.extend([i] * len(Gi))
                else:
                    Cl.append(c); Gs.append(G); par_cells.append(i)
            C = np.array(Cl).reshape(len(Gs), -1)
            cells = _pack(C, Gs)
        # ---- ONE controller call for all cells of the period (replaces the loops at src/solve.jl:153-166, :187-195) ----
        lo, hi = controller_batch(cells)
        # ---- input splitter on every control box, flowpipe records in _solve_one's order (src/solve.jl:230-246) ----
        rec_par, rec_c, rec_G, rec_lo, rec_hi = [], [], [], [], []
        for i in range(len(Gs)):
            if input_splitter is None:
                pieces = [(lo[i], hi[i])]
            elif isinstance(input_splitter, SignSplitter):
                if lo.shape[1] != 1:
                    raise ValueError("SignSplitter is defined for one-dimensional controls (src/splitters.jl:63)")
                pieces = [(np.array([a]), np.array([b])) for a, b in input_splitter.apply(float(lo[i, 0]), float(hi[i, 0]))]
            else:
                pieces = input_splitter(lo[i], hi[i])
            for a, b in pieces:
                rec_par.append(par_cells[i]); rec_c.append(C[i]); rec_G.append(Gs[i]); rec_lo.append(a); rec_hi.append(b)
        lvl = Level(k, np.array(rec_par, dtype=np.int64), np.array(rec_c).reshape(len(rec_G), -1), rec_G,
                    np.array(rec_lo).reshape(len(rec_G), -1), np.array(rec_hi).reshape(len(rec_G), -1))
        if k < periods:
            lvl.end_c, lvl.end_G = plant_step(k, lvl.cell_c, lvl.cell_G, lvl.U_lo, lvl.U_hi)
        levels.append(lvl)
        parents.append(len(rec_G) if k == 1 else lvl.parent)
    return levels, lifo_emission_order(parents)


def emitted_sequence(levels: List[Level], order: List[Tuple[int, int]]):
    """The flowpipe records in the reference's order: (k, cell centre, cell generators, U_lo, U_hi)."""
    return [(levels[l].k, levels[l].cell_c[i], levels[l].cell_G[i], levels[l].U_lo[i], levels[l].U_hi[i]) for l, i in order]
