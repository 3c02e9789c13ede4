"""clr-b200: B200-native hot path of ClosedLoopReachability.jl (host-side mirror).

Import as ``import clr_b200`` (the directory name contains a dot, so the repo root
carries a tiny loader module ``clr_b200.py`` that registers this package).
The compute lives in ``libclr_b200.so`` (CUDA, sm_100a); see ``include/clr_b200.h``.
"""
from .network import DenseLayerOp, FeedforwardNetwork, ACT_IDS  # noqa: F401
from .polar import read_POLAR  # noqa: F401
from .splitters import (BoxGrid, BoxSplitter, IndexedSplitter, NoSplitter, SignSplitter, Splitter,  # noqa: F401
                        ZonotopeCells, ZonotopeSplitter, julia_range, split_box, split_zonotope)

from .problem import (AffinePreprocessing, ControlledPlant, LinearMapPostprocessing, NoPostprocessing,  # noqa: F401
                      NoPreprocessing, ProjectionPostprocessing, UniformAdditivePostprocessing, controller_forward)
from .simulate import EnsembleSimulationSolution, SimulationSolution, sample_box, simulate  # noqa: F401


def __getattr__(name):
    # the CUDA-backed parts are imported lazily so that pure-host helpers work before build()
    if name in ("Context", "DeviceNetwork", "CompiledPlant", "ClrError", "ClrArgumentError", "PrePostArg", "PLANTS"):
        from . import runtime
        return getattr(runtime, name)
    raise AttributeError(name)
