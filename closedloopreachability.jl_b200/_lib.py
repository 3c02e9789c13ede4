"""ctypes binding of libclr_b200.so (the C ABI in include/clr_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared object has not
been built (``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C csrc``) importing
this module raises, and on a machine without a CUDA device ``clr_create`` fails with CLR_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLR_B200_LIB") or os.path.join(_HERE, "libclr_b200.so")
MAX_LAYERS = 32

# every symbol include/clr_b200.h declares (tests check the export list against the header)
SYMBOLS = [
    "clr_create", "clr_destroy", "clr_last_error", "clr_stream", "clr_synchronize", "clr_launch_count",
    "clr_device_alloc", "clr_device_free", "clr_memcpy_h2d", "clr_memcpy_d2h", "clr_host_register", "clr_host_unregister", "clr_set_tile_cells",
    "clr_set_profiling", "clr_last_pass_info", "clr_set_reorder", "clr_measure_fp64_peak", "clr_last_device_error", "clr_set_tiny",
    "clr_net_upload", "clr_net_free", "clr_deepz_boxgrid", "clr_deepz_boxgrid_cyclic", "clr_deepz_zonotopes", "clr_deepz_zonotopes_padded",
    "clr_deepz_fetch_zonotopes", "clr_deepz_fetch_reduced", "clr_project_oa", "clr_forward_points", "clr_rollout",
    "clr_plant_compile", "clr_plant_free", "clr_rollout_plant",
    "clr_split_zonotopes", "clr_sign_split", "clr_create_multi", "clr_destroy_multi", "clr_multi_size", "clr_multi_ctx", "clr_multi_last_error",
    "clr_multi_net_upload", "clr_multi_net_free", "clr_multi_deepz_boxgrid", "clr_multi_rollout",
]


class Stats(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("sum_p_in", C.c_int64 * MAX_LAYERS),
                ("sum_unstable", C.c_int64 * MAX_LAYERS), ("near_threshold", C.c_int64),
                ("max_p_out", C.c_int64), ("flops", C.c_double)]


class PrePost(C.Structure):
    _fields_ = [("pre_rows", C.c_int32), ("preA", C.c_void_p), ("preb", C.c_void_p),
                ("post_kind", C.c_int32), ("post_rows", C.c_int32), ("postA", C.c_void_p),
                ("postb", C.c_void_p), ("post_dims", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build() or "
            "`make -C closedloopreachability.jl_b200/csrc`). clr_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.clr_create.argtypes = [i32, C.POINTER(vp)]
    L.clr_create.restype = i32
    L.clr_destroy.argtypes = [vp]
    L.clr_destroy.restype = None
    L.clr_last_error.argtypes = [vp]
    L.clr_last_error.restype = C.c_char_p
    L.clr_stream.argtypes = [vp]
    L.clr_stream.restype = vp
    L.clr_synchronize.argtypes = [vp]
    L.clr_synchronize.restype = i32
    L.clr_launch_count.argtypes = [vp]
    L.clr_launch_count.restype = i64
    L.clr_device_alloc.argtypes = [vp, i64, C.POINTER(vp)]
    L.clr_device_alloc.restype = i32
    L.clr_device_free.argtypes = [vp, vp]
    L.clr_device_free.restype = i32
    L.clr_memcpy_h2d.argtypes = [vp, vp, vp, i64]
    L.clr_memcpy_h2d.restype = i32
    L.clr_memcpy_d2h.argtypes = [vp, vp, vp, i64]
    L.clr_memcpy_d2h.restype = i32
    L.clr_host_register.argtypes = [vp, vp, i64]
    L.clr_host_register.restype = i32
    L.clr_host_unregister.argtypes = [vp, vp]
    L.clr_host_unregister.restype = i32
    L.clr_set_tile_cells.argtypes = [vp, i32]
    L.clr_set_tile_cells.restype = i32
    L.clr_set_reorder.argtypes = [vp, i32]
    L.clr_set_reorder.restype = i32
    L.clr_set_tiny.argtypes = [vp, i32]
    L.clr_set_tiny.restype = i32
    L.clr_set_profiling.argtypes = [vp, i32]
    L.clr_set_profiling.restype = i32
    L.clr_last_pass_info.argtypes = [vp, vp, vp, vp]
    L.clr_last_pass_info.restype = i32
    L.clr_last_device_error.argtypes = [vp, C.POINTER(i32)]
    L.clr_last_device_error.restype = i32
    L.clr_measure_fp64_peak.argtypes = [vp, dbl, C.POINTER(dbl)]
    L.clr_measure_fp64_peak.restype = i32
    L.clr_net_upload.argtypes = [vp, i32, vp, vp, vp, vp, C.POINTER(vp)]
    L.clr_net_upload.restype = i32
    L.clr_net_free.argtypes = [vp]
    L.clr_net_free.restype = None
    L.clr_deepz_boxgrid.argtypes = [vp, vp, i32, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, i32, i32]
    L.clr_deepz_boxgrid.restype = i32
    L.clr_deepz_boxgrid_cyclic.argtypes = [vp, vp, i32, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp, i32, i32]
    L.clr_deepz_boxgrid_cyclic.restype = i32
    L.clr_deepz_zonotopes.argtypes = [vp, vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32]
    L.clr_deepz_zonotopes.restype = i32
    L.clr_deepz_zonotopes_padded.argtypes = [vp, vp, i32, i64, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32]
    L.clr_deepz_zonotopes_padded.restype = i32
    L.clr_deepz_fetch_zonotopes.argtypes = [vp, vp, vp, vp]
    L.clr_deepz_fetch_zonotopes.restype = i32
    L.clr_deepz_fetch_reduced.argtypes = [vp, i32, vp, vp, i32]
    L.clr_deepz_fetch_reduced.restype = i32
    L.clr_project_oa.argtypes = [vp, i32, i32, i32, vp, i64, vp, vp, vp, vp, i32, i32, vp, vp, vp, i32]
    L.clr_project_oa.restype = i32
    L.clr_forward_points.argtypes = [vp, vp, vp, i32, i64, vp, vp, i32]
    L.clr_forward_points.restype = i32
    L.clr_rollout.argtypes = [vp, vp, vp, i32, i32, i32, i32, i64, i32, vp, vp, dbl, dbl, dbl, i32, dbl, dbl,
                              i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32]
    L.clr_rollout.restype = i32
    L.clr_plant_compile.argtypes = [vp, i32, i32, i32, C.c_char_p, C.POINTER(vp)]
    L.clr_plant_compile.restype = i32
    L.clr_plant_free.argtypes = [vp]
    L.clr_plant_free.restype = None
    L.clr_rollout_plant.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, dbl, dbl, dbl, i32, dbl, dbl, i32, i32, vp, vp, vp, vp, i32,
                                    vp, vp, vp, i32]
    L.clr_rollout_plant.restype = i32
    L.clr_split_zonotopes.argtypes = [vp, i32, i64, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, i32]
    L.clr_split_zonotopes.restype = i32
    L.clr_sign_split.argtypes = [vp, i64, vp, vp, vp, vp, vp, C.POINTER(i64), i32]
    L.clr_sign_split.restype = i32
    L.clr_create_multi.argtypes = [vp, i32, C.POINTER(vp)]
    L.clr_create_multi.restype = i32
    L.clr_destroy_multi.argtypes = [vp]
    L.clr_destroy_multi.restype = None
    L.clr_multi_size.argtypes = [vp]
    L.clr_multi_size.restype = i32
    L.clr_multi_ctx.argtypes = [vp, i32]
    L.clr_multi_ctx.restype = vp
    L.clr_multi_last_error.argtypes = [vp]
    L.clr_multi_last_error.restype = C.c_char_p
    L.clr_multi_net_upload.argtypes = [vp, i32, vp, vp, vp, vp, C.POINTER(vp)]
    L.clr_multi_net_upload.restype = i32
    L.clr_multi_net_free.argtypes = [vp]
    L.clr_multi_net_free.restype = None
    L.clr_multi_deepz_boxgrid.argtypes = [vp, vp, i32, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.clr_multi_deepz_boxgrid.restype = i32
    L.clr_multi_rollout.argtypes = [vp, vp, vp, i32, i32, i32, i32, i64, i32, vp, vp, dbl, dbl, dbl, i32, dbl, dbl, i32, i32,
                                    vp, vp, vp, vp]
    L.clr_multi_rollout.restype = i32
    _lib = L
    return L
