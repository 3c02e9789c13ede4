"""Synthetic workloads of the shapes BASELINE.json names.

Only host-side input construction (NumPy, seeded); no compute.
"""
from __future__ import annotations

import numpy as np

from .network import FeedforwardNetwork
from .splitters import split_box

TORA_X0 = ([0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6])            # models/TORA/TORA.jl:98
UNICYCLE_X0 = ([9.5, -4.5, 2.1, 1.5], [9.55, -4.45, 2.11, 1.51])       # models/Unicycle/Unicycle.jl:72-73
UNICYCLE_W = (-1e-4, 1e-4)


def synthetic_relu_net(seed=4, n_in=6, hidden=100, n_hidden=4, n_out=1) -> FeedforwardNetwork:
    """C4: 6 -> 100 x 4 -> 1, ReLU hidden, Id output; W ~ U(-1,1) sqrt(3/n_in), b ~ U(-0.1, 0.1)."""
    rng = np.random.default_rng(seed)
    dims = [n_in] + [hidden] * n_hidden + [n_out]
    layers = []
    for l in range(len(dims) - 1):
        W = rng.uniform(-1, 1, (dims[l + 1], dims[l])) * np.sqrt(3 / dims[l])
        b = rng.uniform(-0.1, 0.1, dims[l + 1])
        layers.append((W, b, "ReLU" if l < len(dims) - 2 else "Id"))
    return FeedforwardNetwork.from_tuples(layers)


def c4_grid(regime="stable", per_dim=10, n_in=6, last_dim_blocks=None):
    """Box split of the C4 scaling config.  regime 'dense': [-1,1]^6 as is (every neuron unstable,
    p grows to ~360); 'mixed': [0.2,0.4]^6 (p ~ 80); 'stable': [0.29,0.31]^6 (p ~ 11)."""
    c0, s = {"dense": (0.0, 1.0), "mixed": (0.3, 0.1), "stable": (0.3, 0.01)}[regime]
    part = [per_dim] * n_in
    if last_dim_blocks is not None:
        part[-1] = last_dim_blocks
    return split_box([c0 - s] * n_in, [c0 + s] * n_in, part)


def random_zonotopes(rng, N, n, p_lo, p_hi, center_lo, center_hi, gen_scale):
    """N zonotopes with p in [p_lo, p_hi] generators: (C (N,n), col_off (N+1,), G (sum p, n))."""
    p = rng.integers(p_lo, p_hi + 1, N)
    col_off = np.zeros(N + 1, dtype=np.int64)
    col_off[1:] = np.cumsum(p)
    C = rng.uniform(center_lo, center_hi, (N, n))
    G = gen_scale * rng.uniform(-1, 1, (int(col_off[-1]), n))
    return C, col_off, G


def unicycle_samples(rng, N, iters):
    lo, hi = UNICYCLE_X0
    x0 = rng.uniform(lo, hi, (N, 4))
    Wd = rng.uniform(UNICYCLE_W[0], UNICYCLE_W[1], (iters, N, 1))
    return x0, Wd
