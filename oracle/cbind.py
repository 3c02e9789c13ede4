"""ctypes binding of oracle/libclr_oracle.so (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAX_LAYERS = 32
ACT_IDS = {"Id": 0, "ReLU": 1, "Sigmoid": 2, "Tanh": 3}
PLANTS = {"Unicycle": 1, "TORA": 2, "ACC": 3, "InvertedPendulum": 4, "AttitudeControl": 5}


class Stats(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("sum_p_in", C.c_int64 * MAX_LAYERS),
                ("sum_unstable", C.c_int64 * MAX_LAYERS), ("near_threshold", C.c_int64),
                ("max_p_out", C.c_int64), ("flops", C.c_double)]


class PrePost(C.Structure):
    _fields_ = [("pre_rows", C.c_int), ("preA", C.c_void_p), ("preb", C.c_void_p),
                ("post_kind", C.c_int), ("post_rows", C.c_int), ("postA", C.c_void_p),
                ("postb", C.c_void_p), ("post_dims", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libclr_oracle.so")
    src = os.path.join(_HERE, "clr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libclr_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.clro_net_create.restype = C.c_void_p
        _LIB.clro_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Net:
    """Network handle; ``layers`` = [(W (m,n), b (m,), act)]."""

    def __init__(self, layers):
        self.layers = layers
        L = len(layers)
        self.dims = [layers[0][0].shape[1]] + [W.shape[0] for (W, _, _) in layers]
        self._W = [np.asfortranarray(W, dtype=np.float64) for (W, _, _) in layers]
        self._b = [np.ascontiguousarray(b, dtype=np.float64) for (_, b, _) in layers]
        dims = (C.c_int * (L + 1))(*self.dims)
        act = (C.c_int * L)(*[ACT_IDS[a] for (_, _, a) in layers])
        Wp = (C.c_void_p * L)(*[w.ctypes.data for w in self._W])
        bp = (C.c_void_p * L)(*[b.ctypes.data for b in self._b])
        self.h = C.c_void_p(lib().clro_net_create(L, dims, act, Wp, bp))

    def __del__(self):
        try:
            lib().clro_net_free(self.h)
        except Exception:
            pass


class _PP:
    """Keeps the arrays behind a PrePost struct alive."""

    def __init__(self, pre=None, post=None):
        self.keep = []
        s = PrePost()
        if pre is not None:
            A, d = pre
            A = np.asfortranarray(A, dtype=np.float64)
            d = np.ascontiguousarray(d, dtype=np.float64)
            self.keep += [A, d]
            s.pre_rows, s.preA, s.preb = A.shape[0], A.ctypes.data, d.ctypes.data
        if post is not None:
            kind, val = post
            if kind == "shift":
                v = np.array([val], dtype=np.float64); self.keep.append(v)
                s.post_kind, s.postb = 1, v.ctypes.data
            elif kind == "scale":
                v = np.array([val], dtype=np.float64); self.keep.append(v)
                s.post_kind, s.postA = 2, v.ctypes.data
            elif kind == "matrix":
                M = np.asfortranarray(val, dtype=np.float64); self.keep.append(M)
                s.post_kind, s.post_rows, s.postA = 3, M.shape[0], M.ctypes.data
            elif kind == "project":
                d = np.ascontiguousarray(val, dtype=np.int32); self.keep.append(d)
                s.post_kind, s.post_rows, s.post_dims = 4, len(d), d.ctypes.data
            else:
                raise ValueError(kind)
        self.s = s

    def out_dim(self, net):
        return self.s.post_rows if self.s.post_kind in (3, 4) else net.dims[-1]


def stats_dict(st, L):
    return {"n_cells": st.n_cells, "sum_p_in": list(st.sum_p_in)[:L],
            "sum_unstable": list(st.sum_unstable)[:L], "near_threshold": st.near_threshold,
            "max_p_out": st.max_p_out, "flops": st.flops}


def deepz_boxgrid(net: Net, partition, centers, radii, cell_begin=0, cell_count=None, pre=None,
                  post=None, keep_cap=0):
    """centers / radii: per-dimension tables (lists of 1-D arrays, radii per block)."""
    part = np.ascontiguousarray(partition, dtype=np.int32)
    n = len(part)
    total = int(np.prod(part.astype(np.int64)))
    if cell_count is None:
        cell_count = total - cell_begin
    cc = np.ascontiguousarray(np.concatenate(centers), dtype=np.float64)
    rr = np.ascontiguousarray(np.concatenate(radii), dtype=np.float64)
    pp = _PP(pre, post)
    m = pp.out_dim(net)
    lo = np.empty((cell_count, m)); hi = np.empty((cell_count, m))
    st = Stats()
    nc = cg = cG = None
    if keep_cap:
        nc = np.zeros(cell_count, dtype=np.int32); cg = np.zeros((cell_count, m))
        cG = np.zeros((cell_count, keep_cap, m))
    rc = lib().clro_deepz_boxgrid(net.h, n, _p(part), _p(cc), _p(rr), C.c_int64(cell_begin),
                                  C.c_int64(cell_count), C.byref(pp.s), _p(lo), _p(hi), C.byref(st),
                                  keep_cap, _p(nc), _p(cg), _p(cG))
    if rc != 0:
        raise RuntimeError(f"clro_deepz_boxgrid rc={rc}")
    out = {"lo": lo, "hi": hi, "stats": stats_dict(st, len(net.layers))}
    if keep_cap:
        out.update(ncols=nc, c=cg, G=cG)
    return out


def deepz_zonotopes(net: Net, Cc, col_off, G, pre=None, post=None, keep_cap=0):
    """Cc: (N, n) centres; G: (total_cols, n) generators (row j = generator j); col_off: (N+1,)."""
    Cc = np.ascontiguousarray(Cc, dtype=np.float64)
    G = np.ascontiguousarray(G, dtype=np.float64).reshape(-1, Cc.shape[1])
    col_off = np.ascontiguousarray(col_off, dtype=np.int64)
    N, n = Cc.shape
    pp = _PP(pre, post)
    m = pp.out_dim(net)
    lo = np.empty((N, m)); hi = np.empty((N, m))
    st = Stats()
    nc = cg = cG = None
    if keep_cap:
        nc = np.zeros(N, dtype=np.int32); cg = np.zeros((N, m)); cG = np.zeros((N, keep_cap, m))
    rc = lib().clro_deepz_zonotopes(net.h, n, C.c_int64(N), _p(Cc), _p(col_off), _p(G), C.byref(pp.s),
                                    _p(lo), _p(hi), C.byref(st), keep_cap, _p(nc), _p(cg), _p(cG))
    if rc != 0:
        raise RuntimeError(f"clro_deepz_zonotopes rc={rc}")
    out = {"lo": lo, "hi": hi, "stats": stats_dict(st, len(net.layers))}
    if keep_cap:
        out.update(ncols=nc, c=cg, G=cG)
    return out


def reduce_order(ncols, c, G, order=2):
    """ncols: (N,), c: (N, m), G: (N, cap, m) as returned with keep_cap; -> (c, G_red (N, m*order, m))."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    G = np.ascontiguousarray(G, dtype=np.float64)
    ncols = np.ascontiguousarray(ncols, dtype=np.int32)
    N, cap, m = G.shape
    oc = np.empty((N, m)); oG = np.empty((N, m * order, m))
    L = lib()
    L.clro_reduce_order.restype = C.c_int
    rc = L.clro_reduce_order(m, cap, int(order), C.c_int64(N), _p(ncols), _p(c), _p(G), _p(oc), _p(oG))
    if rc != 0:
        raise IndexError(f"clro_reduce_order rc={rc}")
    return oc, oG


def forward_points(net: Net, X, pre=None, post=None):
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, nx = X.shape
    pp = _PP(pre, post)
    U = np.empty((N, pp.out_dim(net)))
    lib().clro_forward_points(net.h, C.byref(pp.s), nx, C.c_int64(N), _p(X), _p(U))
    return U


def rollout(net: Net, plant, n_state, n_dist, n_ctrl, x0, Wd, t0, tau, T, iters, alg="Tsit5",
            abstol=1e-6, reltol=1e-3, substeps=4, pow_mode=1, pre=None, post=None, log_cap=0):
    """x0: (N, n_state); Wd: (iters, N, n_dist) or None.  Returns dict of arrays."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    N = x0.shape[0]
    if n_dist:
        Wd = np.ascontiguousarray(Wd, dtype=np.float64).reshape(iters, N, n_dist)
    pp = _PP(pre, post)
    X_end = np.empty((iters, N, n_state)); U = np.empty((iters, N, n_ctrl))
    ns = np.zeros((iters, N), dtype=np.int32); nr = np.zeros((iters, N), dtype=np.int32)
    n_ext = n_state + n_dist + n_ctrl
    ln = lt = lu = None
    if log_cap:
        ln = np.zeros((iters, N), dtype=np.int32); lt = np.zeros((iters, N, log_cap))
        lu = np.zeros((iters, N, log_cap, n_ext))
    pid = PLANTS[plant] if isinstance(plant, str) else int(plant)
    rc = lib().clro_rollout(net.h, C.byref(pp.s), pid, n_state, n_dist, n_ctrl, C.c_int64(N), iters,
                            _p(x0), _p(Wd) if n_dist else None, C.c_double(t0), C.c_double(tau),
                            C.c_double(T), 0 if alg == "Tsit5" else 1, C.c_double(abstol),
                            C.c_double(reltol), substeps, pow_mode, _p(X_end), _p(U), _p(ns), _p(nr),
                            log_cap, _p(ln), _p(lt), _p(lu))
    if rc != 0:
        raise RuntimeError(f"clro_rollout rc={rc}")
    out = {"X_end": X_end, "U": U, "naccept": ns, "nreject": nr}
    if log_cap:
        out.update(log_n=ln, log_t=lt, log_u=lu)
    return out


def num_threads():
    return lib().clro_num_threads()
