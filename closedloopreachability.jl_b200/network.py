"""FeedforwardNetwork / DenseLayerOp host containers.

Mirror of the ControllerFormats types the reference re-exports
(/root/reference/src/init.jl:7-10) and builds in src/utils.jl:45-86.
Activations are named as in ControllerFormats: ``Id``, ``ReLU``, ``Sigmoid``, ``Tanh``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

ACT_IDS = {"Id": 0, "ReLU": 1, "Sigmoid": 2, "Tanh": 3}


@dataclass
class DenseLayerOp:
    weights: np.ndarray  # (m, n) Float64
    bias: np.ndarray     # (m,)
    activation: str      # key of ACT_IDS

    def __post_init__(self):
        self.weights = np.ascontiguousarray(self.weights, dtype=np.float64)
        self.bias = np.ascontiguousarray(self.bias, dtype=np.float64)
        if self.weights.ndim != 2 or self.bias.shape != (self.weights.shape[0],):
            raise ValueError("inconsistent layer dimensions")
        if self.activation not in ACT_IDS:
            raise ValueError(f"unknown activation {self.activation!r}")

    @property
    def dim_in(self) -> int:
        return self.weights.shape[1]

    @property
    def dim_out(self) -> int:
        return self.weights.shape[0]


@dataclass
class FeedforwardNetwork:
    layers: List[DenseLayerOp]

    def __post_init__(self):
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            if a.dim_out != b.dim_in:
                raise ValueError("layer dimensions do not chain")

    @property
    def dim_in(self) -> int:
        return self.layers[0].dim_in

    @property
    def dim_out(self) -> int:
        return self.layers[-1].dim_out

    @property
    def dims(self):
        return [self.dim_in] + [l.dim_out for l in self.layers]

    def as_tuples(self):
        """[(W, b, act)] -- the plain form the test oracle consumes."""
        return [(l.weights, l.bias, l.activation) for l in self.layers]

    @staticmethod
    def from_tuples(layers):
        return FeedforwardNetwork([DenseLayerOp(W, b, a) for (W, b, a) in layers])

    def save_npz(self, path: str):
        d = {"acts": np.array([l.activation for l in self.layers])}
        for i, l in enumerate(self.layers):
            d[f"W{i}"] = l.weights
            d[f"b{i}"] = l.bias
        np.savez_compressed(path, **d)

    @staticmethod
    def load_npz(path: str) -> "FeedforwardNetwork":
        z = np.load(path, allow_pickle=False)
        acts = [str(a) for a in z["acts"]]
        return FeedforwardNetwork([DenseLayerOp(z[f"W{i}"], z[f"b{i}"], a)
                                   for i, a in enumerate(acts)])
