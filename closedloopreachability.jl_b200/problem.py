"""Host mirror of /root/reference/src/problem.jl: ControlledPlant and the control pre/post-processing types.

Only what the hot path touches is mirrored: the struct, its accessors and the ``apply`` semantics of
the processing types.  The plant is named by a built-in right-hand side (``plant="Unicycle"`` ...;
the reference's plants are arbitrary Julia functions) and the initial states by a box
``(lo, hi)`` over all variables (states, disturbances, controls) as in the model scripts, e.g.
models/Unicycle/Unicycle.jl:72-80.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .network import FeedforwardNetwork


# ---- control postprocessing (src/problem.jl:5-58) ------------------------------------------------
class ControlPostprocessing:
    def apply(self, x):
        raise NotImplementedError

    def as_arg(self):
        """The ``post=`` argument of the runtime calls (clr_prepost)."""
        raise NotImplementedError


class NoPostprocessing(ControlPostprocessing):
    def apply(self, x):
        return x

    def as_arg(self):
        return None


class UniformAdditivePostprocessing(ControlPostprocessing):
    def __init__(self, shift: float):
        if shift == 0:
            warnings.warn("zero additive control postprocessing is discouraged; use `NoPostprocessing` instead")
        self.shift = float(shift)

    def apply(self, x):
        return np.asarray(x, dtype=np.float64) + self.shift

    def as_arg(self):
        return ("shift", self.shift)


class ProjectionPostprocessing(ControlPostprocessing):
    """``dims`` are 1-based like in the reference."""

    def __init__(self, dims: Sequence[int]):
        self.dims = [int(d) for d in dims]

    def apply(self, x):
        return np.asarray(x, dtype=np.float64)[[d - 1 for d in self.dims]]

    def as_arg(self):
        return ("project", [d - 1 for d in self.dims])


class LinearMapPostprocessing(ControlPostprocessing):
    def __init__(self, map):
        self.map = map

    def apply(self, x):
        return np.asarray(self.map, dtype=np.float64) @ np.asarray(x, dtype=np.float64) if np.ndim(self.map) == 2 \
            else float(self.map) * np.asarray(x, dtype=np.float64)

    def as_arg(self):
        return ("matrix", np.asarray(self.map, dtype=np.float64)) if np.ndim(self.map) == 2 else ("scale", float(self.map))


# ---- control preprocessing (src/problem.jl:64-77) ------------------------------------------------
class ControlPreprocessing:
    def apply(self, x):
        raise NotImplementedError

    def as_arg(self):
        raise NotImplementedError


class NoPreprocessing(ControlPreprocessing):
    def apply(self, x):
        return x

    def as_arg(self):
        return None


class AffinePreprocessing(ControlPreprocessing):
    """``FunctionPreprocessing`` restricted to affine maps x -> A x + d (the only kind the shipped models use:
    models/ACC/ACC.jl:81-99), which is what can run on the device."""

    def __init__(self, A, d):
        self.A = np.asarray(A, dtype=np.float64)
        self.d = np.asarray(d, dtype=np.float64)

    def apply(self, x):
        return self.A @ np.asarray(x, dtype=np.float64) + self.d

    def as_arg(self):
        return (self.A, self.d)


# ---- ControlledPlant (src/problem.jl:115-145) -------------------------------------------------------
@dataclass
class ControlledPlant:
    plant: str                                  # built-in right-hand side on the extended state [x; w; u]
    initial_lo: np.ndarray                      # box of initial values of all variables
    initial_hi: np.ndarray
    controller: FeedforwardNetwork
    vars: Dict[str, List[int]]                  # 1-based indices: {"states": [...], "disturbances": [...], "controls": [...]}
    period: float
    preprocessing: ControlPreprocessing = field(default_factory=NoPreprocessing)
    postprocessing: ControlPostprocessing = field(default_factory=NoPostprocessing)

    def __post_init__(self):
        self.initial_lo = np.asarray(self.initial_lo, dtype=np.float64)
        self.initial_hi = np.asarray(self.initial_hi, dtype=np.float64)
        n = len(self.states()) + len(self.disturbances()) + len(self.controls())
        if self.initial_lo.shape != (n,) or self.initial_hi.shape != (n,):
            # src/solve.jl:130-134: dimension mismatch is an ArgumentError
            raise ValueError(f"the initial set has dimension {len(self.initial_lo)} but the variables sum up to {n}")

    def states(self):
        return self.vars.get("states", [])

    def disturbances(self):
        return self.vars.get("disturbances", [])

    def controls(self):
        return self.vars.get("controls", [])

    def project(self, idx: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
        i = [k - 1 for k in idx]
        return self.initial_lo[i], self.initial_hi[i]


def controller_forward(ctx, dnet, prob: ControlledPlant, cells, **kw):
    """Batched ``controller_forward`` (src/solve.jl:210-218) for all cells of a split: ``cells`` is a BoxGrid or
    ZonotopeCells.  Returns the runtime's result dict (boxes per cell in the reference's cell order)."""
    from .splitters import BoxGrid
    pre, post = prob.preprocessing.as_arg(), prob.postprocessing.as_arg()
    if isinstance(cells, BoxGrid):
        return ctx.deepz_boxgrid(dnet, cells.partition, cells.centers, cells.radii, pre=pre, post=post, **kw)
    return ctx.deepz_zonotopes(dnet, cells.C, cells.col_off, cells.G, pre=pre, post=post, **kw)
