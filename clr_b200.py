"""Loader: registers the package directory ``closedloopreachability.jl_b200/`` as ``clr_b200``.

The directory name (fixed by the repo layout) contains a dot and so cannot be
imported by name; importing this module swaps itself for the real package.
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "closedloopreachability.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "clr_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["clr_b200"] = _mod
_spec.loader.exec_module(_mod)
