"""NumPy-facing wrapper of the C ABI (include/clr_b200.h): one ``Context`` per CUDA device.

Array conventions: a batch of vectors is a C-ordered ``(N, dim)`` array, which is the
``dim x N`` column-major layout the C ABI (and Julia) uses.  All compute happens in
libclr_b200.so; nothing here falls back to NumPy.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .network import ACT_IDS, FeedforwardNetwork

PLANTS = {"Unicycle": 1, "TORA": 2, "ACC": 3, "InvertedPendulum": 4}
PLANT_DIMS = {1: (4, 1, 2), 2: (4, 0, 1), 3: (6, 0, 1), 4: (2, 0, 1)}   # n_state, n_dist, n_ctrl
FLAG_DEVICE_IO = 1
FLAG_NO_STATS = 2
FLAG_MERGE_REDUNDANT = 4
ERR_ARG, ERR_CAPACITY, ERR_DIM, ERR_CUDA, ERR_ALLOC, ERR_STATE = -1, -2, -3, -4, -5, -6


class ClrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libclr_b200 error {code}: {msg}")
        self.code = code


class ClrArgumentError(ClrError, ValueError):
    """CLR_ERR_ARG / CLR_ERR_DIM -- what the Julia host turns into ArgumentError (src/solve.jl:88-99, :130-134)."""


def _raise(code, msg):
    if code in (ERR_ARG, ERR_DIM):
        raise ClrArgumentError(code, msg)
    raise ClrError(code, msg)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):            # torch tensor (device or pinned host memory)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class PrePostArg:
    """Builds a clr_prepost; ``pre = (A, d)`` (affine x -> A x + d) or None;
    ``post = None | ("shift", s) | ("scale", a) | ("matrix", M) | ("project", dims0)``
    mirroring the reference's post-processing types (src/problem.jl:5-58)."""

    def __init__(self, pre=None, post=None):
        self.keep = []
        s = _lib.PrePost()
        if pre is not None:
            A, d = pre
            A = np.asfortranarray(A, dtype=np.float64)
            d = np.ascontiguousarray(d, dtype=np.float64)
            if A.ndim != 2 or d.shape != (A.shape[0],):
                raise ClrArgumentError(ERR_DIM, "preprocessing must be (A (r, n), d (r,))")
            self.keep += [A, d]
            s.pre_rows, s.preA, s.preb = A.shape[0], A.ctypes.data, d.ctypes.data
        if post is not None:
            kind, val = post
            if kind == "shift":
                v = np.array([val], dtype=np.float64)
                self.keep.append(v)
                s.post_kind, s.postb = 1, v.ctypes.data
            elif kind == "scale":
                v = np.array([val], dtype=np.float64)
                self.keep.append(v)
                s.post_kind, s.postA = 2, v.ctypes.data
            elif kind == "matrix":
                M = np.asfortranarray(val, dtype=np.float64)
                self.keep.append(M)
                s.post_kind, s.post_rows, s.postA = 3, M.shape[0], M.ctypes.data
            elif kind == "project":
                d = np.ascontiguousarray(val, dtype=np.int32)
                self.keep.append(d)
                s.post_kind, s.post_rows, s.post_dims = 4, len(d), d.ctypes.data
            else:
                raise ClrArgumentError(ERR_ARG, f"unknown postprocessing {kind!r}")
        self.s = s

    def out_dim(self, net_out):
        return self.s.post_rows if self.s.post_kind in (3, 4) else net_out


def stats_dict(st, L):
    return {"n_cells": st.n_cells, "sum_p_in": list(st.sum_p_in)[:L],
            "sum_unstable": list(st.sum_unstable)[:L], "near_threshold": st.near_threshold,
            "max_p_out": st.max_p_out, "flops": st.flops}


class DeviceNetwork:
    """A FeedforwardNetwork uploaded to one Context (clr_net)."""

    def __init__(self, ctx: "Context", net: FeedforwardNetwork):
        self.ctx = ctx
        self.net = net
        self.dims = list(net.dims)
        L = len(net.layers)
        self._W = [np.asfortranarray(l.weights, dtype=np.float64) for l in net.layers]
        self._b = [np.ascontiguousarray(l.bias, dtype=np.float64) for l in net.layers]
        dims = (C.c_int32 * (L + 1))(*self.dims)
        act = (C.c_int32 * L)(*[ACT_IDS[l.activation] for l in net.layers])
        Wp = (C.c_void_p * L)(*[w.ctypes.data for w in self._W])
        bp = (C.c_void_p * L)(*[b.ctypes.data for b in self._b])
        h = C.c_void_p()
        ctx._check(ctx.L.clr_net_upload(ctx.h, L, dims, act, Wp, bp, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.L.clr_net_free(self.h)
        except Exception:
            pass
        self.h = None


class CompiledPlant:
    """A plant right-hand side compiled at run time (clr_plant_compile): ``rhs_body`` is the CUDA C body of
    ``rhs(const double* x, const double* q, double* dx)`` with x the dynamic states and q = [w; u].
    ``ctx=None`` only checks that the source compiles (no device needed)."""

    def __init__(self, ctx, n_state: int, n_dist: int, n_ctrl: int, rhs_body: str):
        self.ctx = ctx
        self.dims = (int(n_state), int(n_dist), int(n_ctrl))
        L = ctx.L if ctx is not None else _lib.lib()
        h = C.c_void_p()
        rc = L.clr_plant_compile(ctx.h if ctx is not None else None, n_state, n_dist, n_ctrl, rhs_body.encode(), C.byref(h))
        if rc != 0:
            _raise(rc, (L.clr_last_error(ctx.h if ctx is not None else None) or b"").decode())
        self.L, self.h = L, h

    def __del__(self):
        try:
            if self.h:
                self.L.clr_plant_free(self.h)
        except Exception:
            pass
        self.h = None


class Context:
    """clr_ctx: bound to one CUDA device; owns a stream and all device scratch memory."""

    def __init__(self, device: int = 0):
        self.L = _lib.lib()
        h = C.c_void_p()
        rc = self.L.clr_create(device, C.byref(h))
        if rc != 0:
            _raise(rc, (self.L.clr_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.clr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            _raise(rc, (self.L.clr_last_error(self.h) or b"").decode())

    # ---- plumbing -------------------------------------------------------------------------------
    @property
    def stream(self) -> int:
        return int(self.L.clr_stream(self.h) or 0)

    def synchronize(self):
        self._check(self.L.clr_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.L.clr_launch_count(self.h))

    def set_tile_cells(self, cells: int):
        self._check(self.L.clr_set_tile_cells(self.h, int(cells)))

    def set_reorder(self, on: bool):
        self._check(self.L.clr_set_reorder(self.h, 1 if on else 0))

    def set_tiny(self, on: bool):
        self._check(self.L.clr_set_tiny(self.h, 1 if on else 0))

    def set_profiling(self, on: bool):
        self._check(self.L.clr_set_profiling(self.h, 1 if on else 0))

    def last_pass_info(self):
        ms = (C.c_double * 3)()
        retried, big = C.c_int64(), C.c_int64()
        self._check(self.L.clr_last_pass_info(self.h, ms, C.byref(retried), C.byref(big)))
        return {"pass_ms": list(ms), "retried": retried.value, "big": big.value}

    def host_register(self, a: np.ndarray) -> None:
        """Page-lock a NumPy array in place (clr_host_register): results are then written into it by one DMA."""
        self._check(self.L.clr_host_register(self.h, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def host_unregister(self, a: np.ndarray) -> None:
        self._check(self.L.clr_host_unregister(self.h, a.ctypes.data_as(C.c_void_p)))

    def last_device_error(self) -> int:
        """Device-side status word of the last CLR_FLAG_DEVICE_IO call, after synchronising (clr_last_device_error)."""
        out = C.c_int32()
        self._check(self.L.clr_last_device_error(self.h, C.byref(out)))
        return out.value

    def measure_fp64_peak(self, ms: float = 20.0) -> float:
        """FP64 mma.sync (DMMA) issue peak of this device in TFLOP/s (clr_measure_fp64_peak)."""
        out = C.c_double()
        self._check(self.L.clr_measure_fp64_peak(self.h, float(ms), C.byref(out)))
        return out.value

    def upload(self, net: FeedforwardNetwork) -> DeviceNetwork:
        return DeviceNetwork(self, net)

    def compile_plant(self, n_state: int, n_dist: int, n_ctrl: int, rhs_body: str) -> CompiledPlant:
        return CompiledPlant(self, n_state, n_dist, n_ctrl, rhs_body)

    # ---- DeepZ ------------------------------------------------------------------------------------
    def deepz_boxgrid(self, dnet: DeviceNetwork, partition, centers, radii, cell_begin=0, cell_count=None,
                      pre=None, post=None, hull=False, stats=True, keep_cap=0, boxes=True, out_lo=None, out_hi=None,
                      block=0, stride=0):
        """All cells [cell_begin, cell_begin+cell_count) of a box split through controller_forward.

        ``centers`` / ``radii``: per-dimension tables (lists of 1-D arrays; ``radii`` per block).
        Returns ``{"lo", "hi"}`` of shape (cell_count, m) plus optional ``hull_lo/hull_hi``, ``stats``.
        """
        part = np.ascontiguousarray(partition, dtype=np.int32)
        n = len(part)
        total = int(np.prod(part.astype(object))) if n else 0
        if cell_count is None:
            cell_count = total - cell_begin
        cc = np.ascontiguousarray(np.concatenate([np.atleast_1d(c) for c in centers]), dtype=np.float64)
        rr = np.ascontiguousarray(np.concatenate([np.atleast_1d(r) for r in radii]), dtype=np.float64)
        if n and (len(cc) != int(part.sum()) or len(rr) != int(part.sum())):
            raise ClrArgumentError(ERR_DIM, "centre / radius tables do not match the partition")
        pp = PrePostArg(pre, post)
        m = pp.out_dim(dnet.dims[-1])
        cnt = max(int(cell_count), 0)
        lo = (out_lo if out_lo is not None else np.empty((cnt, m))) if boxes else None
        hi = (out_hi if out_hi is not None else np.empty((cnt, m))) if boxes else None
        hl = np.empty(m) if hull else None
        hh = np.empty(m) if hull else None
        st = _lib.Stats() if stats else None
        rc = self.L.clr_deepz_boxgrid_cyclic(self.h, dnet.h, n, _ptr(part), _ptr(cc), _ptr(rr), int(cell_begin),
                                             int(cell_count), int(block), int(stride), C.byref(pp.s), _ptr(lo), _ptr(hi),
                                             _ptr(hl), _ptr(hh), C.byref(st) if stats else None, int(keep_cap),
                                             0 if stats else FLAG_NO_STATS)
        self._check(rc)
        out = {"lo": lo, "hi": hi}
        if hull:
            out.update(hull_lo=hl, hull_hi=hh)
        if stats:
            out["stats"] = stats_dict(st, len(dnet.net.layers))
        if keep_cap:
            out.update(self.fetch_zonotopes(cnt, m, keep_cap))
        return out

    def deepz_zonotopes(self, dnet: DeviceNetwork, Cc, col_off, G, pre=None, post=None, hull=False, stats=True,
                        keep_cap=0, boxes=True):
        """Cc: (N, n) centres; G: (total_cols, n), row j = generator j; col_off: (N+1,) int64."""
        Cc = np.ascontiguousarray(Cc, dtype=np.float64)
        if Cc.ndim != 2:
            raise ClrArgumentError(ERR_DIM, "centres must be (N, n)")
        N, n = Cc.shape
        G = np.ascontiguousarray(G, dtype=np.float64).reshape(-1, n)
        col_off = np.ascontiguousarray(col_off, dtype=np.int64)
        if col_off.shape != (N + 1,) or (N > 0 and col_off[-1] != G.shape[0]):
            raise ClrArgumentError(ERR_DIM, "col_off must have N+1 entries and end at the number of generators")
        pp = PrePostArg(pre, post)
        m = pp.out_dim(dnet.dims[-1])
        lo = np.empty((N, m)) if boxes else None
        hi = np.empty((N, m)) if boxes else None
        hl = np.empty(m) if hull else None
        hh = np.empty(m) if hull else None
        st = _lib.Stats() if stats else None
        rc = self.L.clr_deepz_zonotopes(self.h, dnet.h, n, N, _ptr(Cc), _ptr(col_off), _ptr(G), C.byref(pp.s),
                                        _ptr(lo), _ptr(hi), _ptr(hl), _ptr(hh), C.byref(st) if stats else None,
                                        int(keep_cap), 0 if stats else FLAG_NO_STATS)
        self._check(rc)
        out = {"lo": lo, "hi": hi}
        if hull:
            out.update(hull_lo=hl, hull_hi=hh)
        if stats:
            out["stats"] = stats_dict(st, len(dnet.net.layers))
        if keep_cap:
            out.update(self.fetch_zonotopes(N, m, keep_cap))
        return out

    def deepz_zonotopes_padded(self, dnet: DeviceNetwork, Cc, G, ncols, pre=None, post=None, hull=False, stats=True, keep_cap=0):
        """Zonotope cells in the padded layout of ``project_oa`` / ``fetch_zonotopes``: Cc (N, n), G (N, cap, n) with the first
        ncols[i] rows of G[i] the generators of cell i."""
        Cc = np.ascontiguousarray(Cc, dtype=np.float64)
        G = np.ascontiguousarray(G, dtype=np.float64)
        ncols = np.ascontiguousarray(ncols, dtype=np.int32)
        if Cc.ndim != 2 or G.ndim != 3 or G.shape[0] != Cc.shape[0] or G.shape[2] != Cc.shape[1] or ncols.shape != (Cc.shape[0],):
            raise ClrArgumentError(ERR_DIM, "padded zonotopes: Cc (N, n), G (N, cap, n), ncols (N,)")
        N, n = Cc.shape
        pp = PrePostArg(pre, post)
        m = pp.out_dim(dnet.dims[-1])
        lo = np.empty((N, m)); hi = np.empty((N, m))
        hl = np.empty(m) if hull else None
        hh = np.empty(m) if hull else None
        st = _lib.Stats() if stats else None
        self._check(self.L.clr_deepz_zonotopes_padded(self.h, dnet.h, n, N, _ptr(Cc), _ptr(G), _ptr(ncols), G.shape[1], C.byref(pp.s),
                                                      _ptr(lo), _ptr(hi), _ptr(hl), _ptr(hh), C.byref(st) if stats else None,
                                                      int(keep_cap), 0 if stats else FLAG_NO_STATS))
        out = {"lo": lo, "hi": hi}
        if hull:
            out.update(hull_lo=hl, hull_hi=hh)
        if stats:
            out["stats"] = stats_dict(st, len(dnet.net.layers))
        if keep_cap:
            out.update(self.fetch_zonotopes(N, m, keep_cap))
        return out

    def deepz_zonotopes_padded_dev(self, dnet, n, N, d_C, d_G, d_ncols, cap, pp: PrePostArg, d_lo, d_hi, d_hull_lo=None,
                                   d_hull_hi=None, stats=None, keep_cap=0):
        """Device-resident variant: chains the outputs of ``project_oa_dev`` into the controller propagation."""
        flags = FLAG_DEVICE_IO | (0 if stats is not None else FLAG_NO_STATS)
        self._check(self.L.clr_deepz_zonotopes_padded(self.h, dnet.h, int(n), int(N), _ptr(d_C), _ptr(d_G), _ptr(d_ncols), int(cap),
                                                      C.byref(pp.s), _ptr(d_lo), _ptr(d_hi), _ptr(d_hull_lo), _ptr(d_hull_hi),
                                                      C.byref(stats) if stats is not None else None, int(keep_cap), flags))

    def fetch_zonotopes(self, N, m, keep_cap):
        nc = np.zeros(N, dtype=np.int32)
        c = np.zeros((N, m))
        G = np.zeros((N, keep_cap, m))
        if N > 0:
            self._check(self.L.clr_deepz_fetch_zonotopes(self.h, _ptr(nc), _ptr(c), _ptr(G)))
        return {"ncols": nc, "c": c, "G": G}

    def fetch_reduced(self, N, m, order=2):
        """Order-`order` Girard reduction of the zonotopes kept by the last keep_cap > 0 run: the
        Z0 = _reduce_order(U0, 2; force_reduction=true) of TaylorModelReconstructor(box=false)
        (reference src/setops.jl:84-86) for every cell.  -> c (N, m), G (N, m*order, m): row j = generator j."""
        c = np.zeros((N, m))
        G = np.zeros((N, m * order, m))
        if N > 0:
            self._check(self.L.clr_deepz_fetch_reduced(self.h, int(order), _ptr(c), _ptr(G), 0))
        return {"c": c, "G": G}

    def project_oa(self, exps, coeffs, rem, dt, vars_keep, remove_zero_generators=True, remove_redundant_generators=False):
        """Batched ``_project_oa(R::AbstractTaylorModelReachSet, vars, t)`` (reference src/setops.jl:9-12) for the N reach-sets
        that end the chains of a control period (src/solve.jl:181-184).

        exps (M, n): exponents of the monomials in TaylorSeries' coefficient order; coeffs (N, n, order_t + 1, M): time
        coefficient k of component v, as a polynomial in the n normalised space variables; rem (N, n, 2): remainder
        intervals; dt (N,): evaluation time minus the expansion point; vars_keep: 0-based components kept.
        -> c (N, n_keep), G (N, 2n, n_keep) (row j = generator j), ncols (N,): ready for ``deepz_zonotopes``."""
        exps = np.ascontiguousarray(exps, dtype=np.int32)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        if exps.ndim != 2 or coeffs.ndim != 4 or coeffs.shape[1] != exps.shape[1] or coeffs.shape[3] != exps.shape[0]:
            raise ClrArgumentError(ERR_DIM, "coeffs must be (N, n, order_t + 1, M) with exps (M, n)")
        N, n, KT, M = coeffs.shape
        rem = np.ascontiguousarray(rem, dtype=np.float64)
        dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (N,)))
        if rem.shape != (N, n, 2):
            raise ClrArgumentError(ERR_DIM, "rem must be (N, n, 2)")
        vk = np.ascontiguousarray(vars_keep, dtype=np.int32)
        c = np.zeros((N, len(vk))); G = np.zeros((N, 2 * n, len(vk))); nc = np.zeros(N, dtype=np.int32)
        self._check(self.L.clr_project_oa(self.h, n, KT - 1, M, _ptr(exps), N, _ptr(coeffs), _ptr(rem), _ptr(dt), _ptr(vk),
                                          len(vk), int(bool(remove_zero_generators)), _ptr(c), _ptr(G), _ptr(nc),
                                          FLAG_MERGE_REDUNDANT if remove_redundant_generators else 0))
        return {"c": c, "G": G, "ncols": nc}

    def project_oa_dev(self, n, order_t, exps, N, d_coeffs, d_rem, d_dt, vars_keep, remove_zero_generators, d_c, d_G, d_ncols,
                       remove_redundant_generators=False):
        """Same with device-resident arrays (CLR_FLAG_DEVICE_IO); exps / vars_keep stay host arrays."""
        exps = np.ascontiguousarray(exps, dtype=np.int32)
        vk = np.ascontiguousarray(vars_keep, dtype=np.int32)
        self._check(self.L.clr_project_oa(self.h, int(n), int(order_t), exps.shape[0], _ptr(exps), int(N), _ptr(d_coeffs), _ptr(d_rem),
                                          _ptr(d_dt), _ptr(vk), len(vk), int(bool(remove_zero_generators)), _ptr(d_c), _ptr(d_G),
                                          _ptr(d_ncols), FLAG_DEVICE_IO | (FLAG_MERGE_REDUNDANT if remove_redundant_generators else 0)))

    def split_zonotopes(self, Cc, G, ncols, gens, nparts):
        """apply(ZonotopeSplitter(gens, nparts), Z) for a batch in the padded layout: Cc (N, n), G (N, cap, n), ncols (N,); gens
        0-based.  Returns the children (N * 2^H of them, child-major) in the same layout (clr_split_zonotopes)."""
        Cc = np.ascontiguousarray(Cc, dtype=np.float64)
        G = np.ascontiguousarray(G, dtype=np.float64)
        N, n = Cc.shape
        cap = G.shape[1] if G.ndim == 3 else 0
        nc = np.ascontiguousarray(ncols, dtype=np.int32)
        g = np.ascontiguousarray(gens, dtype=np.int32)
        t = np.ascontiguousarray(nparts, dtype=np.int32)
        H = int(t.sum())
        oc = np.empty((N << H, n)); oG = np.empty((N << H, cap, n)); on = np.empty(N << H, dtype=np.int32)
        self._check(self.L.clr_split_zonotopes(self.h, n, N, _ptr(Cc), _ptr(G), _ptr(nc), cap, len(g), _ptr(g), _ptr(t), _ptr(oc), _ptr(oG),
                                               _ptr(on), 0))
        return {"c": oc, "G": oG, "ncols": on}

    def sign_split(self, lo, hi):
        """apply(SignSplitter(), U) for a batch of intervals: (lo, hi, parent index) of the pieces, order kept (clr_sign_split)."""
        lo = np.ascontiguousarray(lo, dtype=np.float64).ravel(); hi = np.ascontiguousarray(hi, dtype=np.float64).ravel()
        N = len(lo)
        ol = np.empty(2 * N); oh = np.empty(2 * N); op = np.empty(2 * N, dtype=np.int64)
        cnt = C.c_int64(0)
        self._check(self.L.clr_sign_split(self.h, N, _ptr(lo), _ptr(hi), _ptr(ol), _ptr(oh), _ptr(op), C.byref(cnt), 0))
        k = cnt.value
        return ol[:k], oh[:k], op[:k]

    def fetch_reduced_dev(self, order, d_c, d_G):
        """Same with device-resident outputs (CLR_FLAG_DEVICE_IO): d_c (N, m), d_G (N, m*order, m)."""
        self._check(self.L.clr_deepz_fetch_reduced(self.h, int(order), _ptr(d_c), _ptr(d_G), FLAG_DEVICE_IO))

    # device-resident variant: every data pointer is a device pointer / torch CUDA tensor
    def deepz_boxgrid_dev(self, dnet, n, d_partition_host, d_centers, d_radii, cell_begin, cell_count, pp: PrePostArg,
                          d_lo, d_hi, d_hull_lo=None, d_hull_hi=None, stats=None, keep_cap=0, block=0, stride=0):
        part = np.ascontiguousarray(d_partition_host, dtype=np.int32)
        flags = FLAG_DEVICE_IO | (0 if stats is not None else FLAG_NO_STATS)
        rc = self.L.clr_deepz_boxgrid_cyclic(self.h, dnet.h, n, _ptr(part), _ptr(d_centers), _ptr(d_radii),
                                      int(cell_begin), int(cell_count), int(block), int(stride), C.byref(pp.s), _ptr(d_lo), _ptr(d_hi),
                                      _ptr(d_hull_lo), _ptr(d_hull_hi), C.byref(stats) if stats is not None else None,
                                      int(keep_cap), flags)
        self._check(rc)

    def deepz_zonotopes_dev(self, dnet, n, N, d_C, d_col_off, d_G, pp: PrePostArg, d_lo, d_hi, d_hull_lo=None,
                            d_hull_hi=None, stats=None, keep_cap=0):
        flags = FLAG_DEVICE_IO | (0 if stats is not None else FLAG_NO_STATS)
        rc = self.L.clr_deepz_zonotopes(self.h, dnet.h, n, int(N), _ptr(d_C), _ptr(d_col_off), _ptr(d_G),
                                        C.byref(pp.s), _ptr(d_lo), _ptr(d_hi), _ptr(d_hull_lo), _ptr(d_hull_hi),
                                        C.byref(stats) if stats is not None else None, int(keep_cap), flags)
        self._check(rc)

    # ---- pointwise controller / rollouts ---------------------------------------------------------------
    def forward_points(self, dnet: DeviceNetwork, X, pre=None, post=None):
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ClrArgumentError(ERR_DIM, "X must be (N, n)")
        N, nx = X.shape
        pp = PrePostArg(pre, post)
        U = np.empty((N, pp.out_dim(dnet.dims[-1])))
        self._check(self.L.clr_forward_points(self.h, dnet.h, C.byref(pp.s), nx, N, _ptr(X), _ptr(U), 0))
        return U

    def rollout(self, dnet: DeviceNetwork, plant, x0, Wd, t0, tau, T, iters, alg="Tsit5", abstol=1e-6, reltol=1e-3,
                substeps=4, pow_mode=1, pre=None, post=None, log_cap=0, counts=True, out=None):
        """x0: (N, n_state); Wd: (iters, N, n_dist) or None.  Returns X_end (iters, N, n_state), U (iters, N, n_ctrl).
        `out` = {"X_end": array, "U": array}: results go into the caller's (e.g. page-locked, reused) arrays."""
        user = plant if isinstance(plant, CompiledPlant) else None
        pid = 0 if user else (PLANTS[plant] if isinstance(plant, str) else int(plant))
        ns, nd, nc = user.dims if user else PLANT_DIMS.get(pid, (0, 0, 0))
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        if x0.ndim != 2:
            raise ClrArgumentError(ERR_DIM, "x0 must be (N, n_state)")
        N = x0.shape[0]
        if user and x0.shape[1] != ns:
            raise ClrArgumentError(ERR_DIM, f"the plant has {ns} states, x0 has {x0.shape[1]} columns")
        if nd:
            if Wd is None:
                raise ClrArgumentError(ERR_ARG, "the plant has disturbances: Wd is required")
            Wd = np.ascontiguousarray(Wd, dtype=np.float64).reshape(iters, N, nd)
        pp = PrePostArg(pre, post)
        X_end = out["X_end"] if out else np.empty((iters, N, ns))
        U = out["U"] if out else np.empty((iters, N, nc))
        if X_end.shape != (iters, N, ns) or U.shape != (iters, N, nc) or X_end.dtype != np.float64 or U.dtype != np.float64 \
                or not X_end.flags.c_contiguous or not U.flags.c_contiguous:
            raise ClrArgumentError(ERR_DIM, "out arrays must be C-contiguous float64 of shapes (iters, N, n_state) / (iters, N, n_ctrl)")
        na = np.zeros((iters, N), dtype=np.int32) if counts else None
        nr = np.zeros((iters, N), dtype=np.int32) if counts else None
        ln = lt = lu = None
        if log_cap:
            ln = np.zeros((iters, N), dtype=np.int32)
            lt = np.zeros((iters, N, log_cap))
            lu = np.zeros((iters, N, log_cap, ns + nd + nc))
        tail = (float(t0), float(tau), float(T), 0 if alg == "Tsit5" else 1, float(abstol), float(reltol), int(substeps),
                int(pow_mode), _ptr(X_end), _ptr(U), _ptr(na), _ptr(nr), int(log_cap), _ptr(ln), _ptr(lt), _ptr(lu), 0)
        if user:
            rc = self.L.clr_rollout_plant(self.h, dnet.h, C.byref(pp.s), user.h, N, int(iters), _ptr(x0),
                                          _ptr(Wd) if nd else None, *tail)
        else:
            rc = self.L.clr_rollout(self.h, dnet.h, C.byref(pp.s), pid, x0.shape[1], nd, nc, N, int(iters), _ptr(x0),
                                    _ptr(Wd) if nd else None, *tail)
        self._check(rc)
        res = {"X_end": X_end, "U": U}
        if counts:
            res.update(naccept=na, nreject=nr)
        if log_cap:
            res.update(log_n=ln, log_t=lt, log_u=lu)
        return res

    def rollout_dev(self, dnet, plant, N, iters, d_x0, d_Wd, t0, tau, T, alg, abstol, reltol, substeps, pow_mode,
                    pp: PrePostArg, d_X_end, d_U, d_naccept=None, d_nreject=None):
        pid = PLANTS[plant] if isinstance(plant, str) else int(plant)
        ns, nd, nc = PLANT_DIMS[pid]
        rc = self.L.clr_rollout(self.h, dnet.h, C.byref(pp.s), pid, ns, nd, nc, int(N), int(iters), _ptr(d_x0),
                                _ptr(d_Wd), float(t0), float(tau), float(T), 0 if alg == "Tsit5" else 1,
                                float(abstol), float(reltol), int(substeps), int(pow_mode), _ptr(d_X_end), _ptr(d_U),
                                _ptr(d_naccept), _ptr(d_nreject), 0, None, None, None, FLAG_DEVICE_IO)
        self._check(rc)


class MultiContext:
    """clr_multi: several GPUs of one box behind one handle (one host thread and stream per device inside the library; cells
    and trajectories sharded by contiguous id range, no exchange step, hulls gathered peer to peer).  Host buffers only."""

    def __init__(self, devices):
        self.L = _lib.lib()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.L.clr_create_multi(devs, len(devices), C.byref(h))
        if rc != 0:
            _raise(rc, (self.L.clr_last_error(None) or b"").decode())
        self.h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "h", None):
            self.L.clr_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            _raise(rc, (self.L.clr_multi_last_error(self.h) or b"").decode())

    def __len__(self):
        return int(self.L.clr_multi_size(self.h))

    def set_reorder(self, on: bool):
        for i in range(len(self)):
            self.L.clr_set_reorder(self.L.clr_multi_ctx(self.h, i), 1 if on else 0)

    def launch_count(self):
        return [int(self.L.clr_launch_count(self.L.clr_multi_ctx(self.h, i))) for i in range(len(self))]

    def upload(self, net: FeedforwardNetwork):
        return MultiNetwork(self, net)

    def deepz_boxgrid(self, mnet, partition, centers, radii, cell_begin=0, cell_count=None, pre=None, post=None, hull=False,
                      stats=True, balance=False):
        part = np.ascontiguousarray(partition, dtype=np.int32)
        n = len(part)
        total = int(np.prod(part.astype(object)))
        if cell_count is None:
            cell_count = total - cell_begin
        cc = np.ascontiguousarray(np.concatenate([np.atleast_1d(c) for c in centers]), dtype=np.float64)
        rr = np.ascontiguousarray(np.concatenate([np.atleast_1d(r) for r in radii]), dtype=np.float64)
        pp = PrePostArg(pre, post)
        m = pp.out_dim(mnet.dims[-1])
        cnt = max(int(cell_count), 0)
        lo, hi = np.empty((cnt, m)), np.empty((cnt, m))
        hl = np.empty(m) if hull else None
        hh = np.empty(m) if hull else None
        st = _lib.Stats() if stats else None
        G = len(self)
        bounds = np.zeros(G + 1, dtype=np.int64)
        ms = np.zeros(G, dtype=np.float64)
        rc = self.L.clr_multi_deepz_boxgrid(self.h, mnet.h, n, _ptr(part), _ptr(cc), _ptr(rr), int(cell_begin), int(cell_count),
                                            C.byref(pp.s), _ptr(lo), _ptr(hi), _ptr(hl), _ptr(hh), C.byref(st) if stats else None,
                                            {False: 0, True: 1, "cost": 1, "cyclic": 2, 0: 0, 1: 1, 2: 2}[balance], _ptr(bounds), _ptr(ms))
        self._check(rc)
        out = {"lo": lo, "hi": hi, "shard_bounds": bounds, "shard_ms": ms}
        if hull:
            out.update(hull_lo=hl, hull_hi=hh)
        if stats:
            out["stats"] = stats_dict(st, len(mnet.net.layers))
        return out

    def rollout(self, mnet, plant, x0, Wd, t0, tau, T, iters, alg="Tsit5", abstol=1e-6, reltol=1e-3, substeps=4, pow_mode=1,
                pre=None, post=None):
        pid = PLANTS[plant] if isinstance(plant, str) else int(plant)
        ns, nd, nc = PLANT_DIMS[pid]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        N = x0.shape[0]
        Wd = np.ascontiguousarray(Wd, dtype=np.float64).reshape(iters, N, nd) if nd else None
        pp = PrePostArg(pre, post)
        X = np.empty((iters, N, ns))
        U = np.empty((iters, N, nc))
        na = np.empty((iters, N), dtype=np.int32)
        nr = np.empty((iters, N), dtype=np.int32)
        rc = self.L.clr_multi_rollout(self.h, mnet.h, C.byref(pp.s), pid, ns, nd, nc, N, int(iters), _ptr(x0), _ptr(Wd), float(t0),
                                      float(tau), float(T), 0 if alg == "Tsit5" else 1, float(abstol), float(reltol), int(substeps), int(pow_mode),
                                      _ptr(X), _ptr(U), _ptr(na), _ptr(nr))
        self._check(rc)
        return {"X_end": X, "U": U, "naccept": na, "nreject": nr}


class MultiNetwork:
    """A FeedforwardNetwork replicated on every device of a MultiContext (clr_multi_net)."""

    def __init__(self, mctx: MultiContext, net: FeedforwardNetwork):
        self.mctx, self.net = mctx, net
        self.dims = list(net.dims)
        L = len(net.layers)
        W = [np.asfortranarray(l.weights, dtype=np.float64) for l in net.layers]
        b = [np.ascontiguousarray(l.bias, dtype=np.float64) for l in net.layers]
        dims = (C.c_int32 * (L + 1))(*self.dims)
        act = (C.c_int32 * L)(*[ACT_IDS[l.activation] for l in net.layers])
        Wp = (C.c_void_p * L)(*[w.ctypes.data for w in W])
        bp = (C.c_void_p * L)(*[x.ctypes.data for x in b])
        h = C.c_void_p()
        mctx._check(mctx.L.clr_multi_net_upload(mctx.h, L, dims, act, Wp, bp, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.mctx.h:
                self.mctx.L.clr_multi_net_free(self.h)
        except Exception:
            pass
        self.h = None
