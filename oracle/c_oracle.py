"""ctypes wrapper over oracle/gnx_oracle.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgnx_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gnx_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "gnx_math.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr) if os.path.exists(p))
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-s", "-C", _HERE], env=env)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_exp.restype = C.c_double
        _lib.orc_exp.argtypes = [C.c_double]
        _lib.orc_expf_cr.restype = C.c_float
        _lib.orc_expf_cr.argtypes = [C.c_float]
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline should use every host core."""
    n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(n))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return num_threads()


def pack_windows(arrs):
    """list of 2-D arrays -> (flat concat, int64 offsets)"""
    offs = np.zeros(len(arrs), dtype=np.int64)
    tot = 0
    for i, a in enumerate(arrs):
        offs[i] = tot
        tot += a.size
    flat = np.concatenate([np.ascontiguousarray(a).ravel() for a in arrs])
    return flat, offs


def lr_f64(X, coefs, intercepts, C_, M, ctx, A):
    X = np.ascontiguousarray(X, dtype=np.int8)
    N = X.shape[0]
    W = C_ // M
    flat, offs = pack_windows([np.asarray(c, dtype=np.float64) for c in coefs])
    b = np.ascontiguousarray(np.asarray(intercepts, dtype=np.float64))
    out = np.zeros((N, W, A), dtype=np.float64)
    lib().orc_lr_f64(_p(X), C.c_int64(N), C.c_int64(X.strides[0]), C.c_int64(C_), C.c_int64(M),
                     C.c_int64(ctx), C.c_int(A), _p(flat), _p(offs), _p(b), _p(out))
    return out


def lr_fixed(X, qfold, intercepts, C_, M, ctx, A, s, want_f64=False):
    X = np.ascontiguousarray(X, dtype=np.int8)
    N = X.shape[0]
    W = C_ // M
    flat, offs = pack_windows([np.asarray(q, dtype=np.int64) for q in qfold])
    b = np.ascontiguousarray(np.asarray(intercepts, dtype=np.float64))
    Bf = np.zeros((N, W, A), dtype=np.float32)
    Bd = np.zeros((N, W, A), dtype=np.float64) if want_f64 else None
    lib().orc_lr_fixed(_p(X), C.c_int64(N), C.c_int64(X.strides[0]), C.c_int64(C_), C.c_int64(M),
                       C.c_int64(ctx), C.c_int(A), _p(flat), _p(offs), C.c_int(s), _p(b), _p(Bf), _p(Bd))
    return (Bf, Bd) if want_f64 else Bf


def slide_window(B, S):
    B = np.ascontiguousarray(B, dtype=np.float32)
    N, W, A = B.shape
    out = np.zeros((N * W, S * A), dtype=np.float32)
    lib().orc_slide_window(_p(B), C.c_int64(N), C.c_int64(W), C.c_int(A), C.c_int(S), _p(out))
    return out


def _gbt_args(m):
    return (C.c_int(m.n_trees), _p(m.feat), _p(m.thr), _p(m.left), _p(m.right), _p(m.default_left),
            _p(m.leaf), _p(m.tree_offsets), _p(m.base_margin))


def gbt_rows(m, rows):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    k, F = rows.shape
    out = np.zeros((k, m.A), dtype=np.float32)
    lib().orc_gbt_rows(C.c_int(m.A), *_gbt_args(m), _p(rows), C.c_int64(k), C.c_int64(F), _p(out))
    return out


def gbt_smooth(m, B, S, want_proba=True):
    B = np.ascontiguousarray(B, dtype=np.float32)
    N, W, A = B.shape
    proba = np.zeros((N, W, A), dtype=np.float32) if want_proba else None
    label = np.zeros((N, W), dtype=np.int32)
    lib().orc_gbt_smooth(C.c_int(A), C.c_int(S), *_gbt_args(m), _p(B), C.c_int64(N), C.c_int64(W),
                         _p(proba), _p(label))
    return proba, label


def crf_smooth(B, state_w, trans_w):
    B = np.ascontiguousarray(B, dtype=np.float64)
    N, W, A = B.shape
    sw = np.ascontiguousarray(state_w, dtype=np.float64)
    tw = np.ascontiguousarray(trans_w, dtype=np.float64)
    L = sw.shape[1]
    proba = np.zeros((N, W, L), dtype=np.float64)
    label = np.zeros((N, W), dtype=np.int32)
    lib().orc_crf_smooth(_p(B), C.c_int64(N), C.c_int64(W), C.c_int(A), C.c_int(L), _p(sw), _p(tw),
                         _p(proba), _p(label))
    return proba, label


def covrsk(X, Y, Ms):
    X = np.ascontiguousarray(X, dtype=np.int8)
    Y = np.ascontiguousarray(Y, dtype=np.int8)
    Mlen = X.shape[1]
    ohe = np.zeros(Mlen + 2, dtype=np.uint8)
    for m in Ms:
        if m <= Mlen:
            ohe[m] = 1
    K = np.zeros((X.shape[0], Y.shape[0]), dtype=np.int64)
    lib().orc_covrsk(_p(X), C.c_int64(X.shape[0]), _p(Y), C.c_int64(Y.shape[0]), C.c_int64(Mlen),
                     _p(ohe), _p(K))
    return K


def svc_proba(K, n_support, dual_coef, intercept, probA, probB):
    K = np.ascontiguousarray(K, dtype=np.int64)
    n, nSV = K.shape
    ns = np.ascontiguousarray(n_support, dtype=np.int32)
    k = len(ns)
    dc = np.ascontiguousarray(dual_coef, dtype=np.float64)
    ic = np.ascontiguousarray(intercept, dtype=np.float64)
    pa = np.ascontiguousarray(probA, dtype=np.float64)
    pb = np.ascontiguousarray(probB, dtype=np.float64)
    out = np.zeros((n, k), dtype=np.float64)
    lib().orc_svc_proba(_p(K), C.c_int64(n), C.c_int64(nSV), C.c_int(k), _p(ns), _p(dc), _p(ic),
                        _p(pa), _p(pb), _p(out))
    return out
