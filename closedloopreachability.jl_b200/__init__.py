"""clr-b200: B200-native hot path of ClosedLoopReachability.jl (host-side mirror).

Import as ``import clr_b200`` (the directory name contains a dot, so the repo root
carries a tiny loader module ``clr_b200.py`` that registers this package).
"""
from .network import DenseLayerOp, FeedforwardNetwork, ACT_IDS  # noqa: F401
from .polar import read_POLAR  # noqa: F401
