"""read_POLAR: controller networks in the POLAR/Sherlock text format.

Mirrors ControllerFormats.read_POLAR, the call every reference model script makes
(e.g. /root/reference/models/TORA/TORA.jl:69, models/Unicycle/Unicycle.jl:55).

Layout (verified on all 22 bundled files): n_in, n_out, number of
hidden layers, the hidden sizes, one activation name per layer (``ReLU``,
``sigmoid``, ``tanh``, ``Affine``), then for each layer and each output neuron its
weights followed by its bias, then two trailing lines (output offset and scale).
"""
from __future__ import annotations

import numpy as np

from .network import DenseLayerOp, FeedforwardNetwork

_ACT = {"relu": "ReLU", "sigmoid": "Sigmoid", "tanh": "Tanh", "affine": "Id", "id": "Id",
        "linear": "Id"}


def read_POLAR(path: str, parse_dtype=np.float64) -> FeedforwardNetwork:
    """Parse a ``.polar`` file.

    ``parse_dtype=np.float32`` reproduces a reader that parses numbers as Float32
    before widening (weight eltype ambiguity noted in SURVEY.md section 8c); the default
    parses as Float64.
    """
    with open(path, "r") as fh:
        tok = fh.read().split()
    pos = 0

    def nxt():
        nonlocal pos
        t = tok[pos]
        pos += 1
        return t

    n_in = int(nxt())
    n_out = int(nxt())
    n_hidden = int(nxt())
    hidden = [int(nxt()) for _ in range(n_hidden)]
    acts = []
    for _ in range(n_hidden + 1):
        name = nxt()
        key = name.lower()
        if key not in _ACT:
            raise ValueError(f"unknown activation {name!r} in {path}")
        acts.append(_ACT[key])
    dims = [n_in] + hidden + [n_out]
    layers = []
    for l in range(n_hidden + 1):
        m, n = dims[l + 1], dims[l]
        flat = np.array(tok[pos:pos + m * (n + 1)], dtype=parse_dtype).astype(np.float64)
        if flat.size != m * (n + 1):
            raise ValueError(f"truncated POLAR file {path}")
        pos += m * (n + 1)
        blk = flat.reshape(m, n + 1)
        layers.append(DenseLayerOp(np.ascontiguousarray(blk[:, :n]), blk[:, n].copy(), acts[l]))
    rest = tok[pos:]
    if len(rest) != 2:
        raise ValueError(f"unexpected trailer in POLAR file {path}: {len(rest)} tokens")
    return FeedforwardNetwork(layers)
