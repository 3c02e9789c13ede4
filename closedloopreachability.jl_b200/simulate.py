"""Host mirror of /root/reference/src/simulate.jl: ``simulate`` and its result containers.

``simulate(ctx, prob, T=..., trajectories=N)`` keeps the reference's structure -- sample x0 in X0 once, then per
control period: controller on every trajectory, fresh disturbances, extended state, ODE solve, projection
(src/simulate.jl:89-137) -- but the whole period loop is ONE ``clr_rollout`` call; the disturbances are pre-drawn in
the reference's order (x0 first, then W0 once per period: nothing else consumes the generator).  The solver is the
reference's: adaptive Tsit5 with OrdinaryDiffEq's defaults (ext/ClosedLoopReachabilityOrdinaryDiffEqExt.jl:28-45).
Results are array backed; ``EnsembleSimulationSolution`` exposes the accessors of src/simulate.jl:1-47 lazily.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .problem import ControlledPlant


def sample_box(rng, lo, hi, n, include_vertices=False):
    """LazySets ``sample(H, n; include_vertices)`` for a hyperrectangle: uniform per dimension; the vertices
    (``vertices_list``) are appended when asked for, which makes the batch longer (src/simulate.jl:89-90)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    x = lo + (hi - lo) * rng.random((n, len(lo)))
    if include_vertices:
        verts = np.array(list(itertools.product(*[(l,) if l == h else (h, l) for l, h in zip(lo, hi)])), dtype=np.float64)
        x = np.vstack([x, verts.reshape(-1, len(lo))])
    return x


@dataclass
class SimulationSolution:
    """One trajectory: pieces per control cycle (src/simulate.jl:1-15)."""
    t: list            # per piece: times of the accepted steps
    u: list            # per piece: extended states at those times (steps x n_ext)
    controls_: np.ndarray      # (pieces, n_ctrl)
    disturbances_: Optional[np.ndarray]   # (pieces, n_dist) or None

    def __len__(self):
        return len(self.controls_)

    def trajectory(self):
        return list(zip(self.t, self.u))

    def controls(self):
        return self.controls_

    def disturbances(self):
        return self.disturbances_


class EnsembleSimulationSolution:
    """Array-backed ensemble (src/simulate.jl:17-47): nothing per trajectory is materialised until asked for."""

    def __init__(self, res, x0, Wd, t0, tau, T, n_state, n_dist, n_ctrl):
        self.res, self.x0, self.Wd = res, x0, Wd
        self.t0, self.tau, self.T = t0, tau, T
        self.n_state, self.n_dist, self.n_ctrl = n_state, n_dist, n_ctrl

    def __len__(self):
        return self.x0.shape[0]

    @property
    def X_end(self):
        """(pieces, N, n_state): states at the end of every control cycle."""
        return self.res["X_end"]

    @property
    def U(self):
        return self.res["U"]

    def __getitem__(self, j) -> SimulationSolution:
        iters = self.res["U"].shape[0]
        ts, us = [], []
        if "log_n" in self.res:
            for i in range(iters):
                k = int(self.res["log_n"][i, j])
                ts.append(self.res["log_t"][i, j, :k].copy())
                us.append(self.res["log_u"][i, j, :k].copy())
        else:                      # end points only
            for i in range(iters):
                a = self.t0 + i * self.tau
                b = self.T if i == iters - 1 else a + self.tau
                start = self.x0[j] if i == 0 else self.res["X_end"][i - 1, j]
                ext = lambda x: np.concatenate([x] + ([self.Wd[i, j]] if self.n_dist else []) + [self.res["U"][i, j]])
                ts.append(np.array([a, b]))
                us.append(np.vstack([ext(start), ext(self.res["X_end"][i, j])]))
        return SimulationSolution(ts, us, self.res["U"][:, j], self.Wd[:, j] if self.n_dist else None)

    def solutions(self, i):
        """Piece i of every trajectory (src/simulate.jl:39-41)."""
        return [self[j].trajectory()[i] for j in range(len(self))]

    def trajectory(self, j):
        return self[j].trajectory()

    def trajectories(self):
        return [self[j].trajectory() for j in range(len(self))]

    def controls(self, j=None):
        return self.res["U"].transpose(1, 0, 2) if j is None else self.res["U"][:, j]

    def disturbances(self, j=None):
        if not self.n_dist:
            return None
        return self.Wd.transpose(1, 0, 2) if j is None else self.Wd[:, j]


def simulate(ctx, prob: ControlledPlant, T=None, tspan=None, trajectories=10, include_vertices=False, rng=None,
             save_steps=False, dnet=None, alg="Tsit5", abstol=1e-6, reltol=1e-3, log_cap=64):
    """``simulate(cp; T | tspan, trajectories, include_vertices)`` (src/simulate.jl:68-143)."""
    if tspan is None:
        if T is None:
            raise ValueError("either T or tspan must be given")   # _get_tspan
        tspan = (0.0, float(T))
    t0, tend = float(tspan[0]), float(tspan[1])
    tau = float(prob.period)
    iterations = int(math.ceil((tend - t0) / tau))
    rng = rng if rng is not None else np.random.default_rng()
    xlo, xhi = prob.project(prob.states())
    x0 = sample_box(rng, xlo, xhi, int(trajectories), include_vertices)
    N = x0.shape[0]
    nd = len(prob.disturbances())
    Wd = None
    if nd:
        wlo, whi = prob.project(prob.disturbances())
        Wd = np.stack([sample_box(rng, wlo, whi, N) for _ in range(iterations)])
    dnet = dnet if dnet is not None else ctx.upload(prob.controller)
    res = ctx.rollout(dnet, prob.plant, x0, Wd, t0, tau, tend, iterations, alg=alg, abstol=abstol, reltol=reltol,
                      pre=prob.preprocessing.as_arg(), post=prob.postprocessing.as_arg(),
                      log_cap=log_cap if save_steps else 0)
    return EnsembleSimulationSolution(res, x0, Wd, t0, tau, tend, x0.shape[1], nd, len(prob.controls()))
