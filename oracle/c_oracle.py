"""ctypes front end of oracle/c/sem_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header).
Only tests/ and bench.py's CPU-baseline legs import this."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsem_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", _HERE])
        h = C.CDLL(_SO)
        h.so_num_threads.restype = C.c_int
        h.so_ax.restype = None
        h.so_gs.restype = None
        h.so_cg.restype = C.c_int
        _lib = h
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().so_num_threads()


def elliptic_ax(geom9, p, N, D):
    n = N + 1
    geom9 = np.ascontiguousarray(geom9, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    D = np.ascontiguousarray(D, dtype=np.float64)
    out = np.empty_like(p)
    lib().so_ax(C.c_int64(geom9.shape[0]), C.c_int(n), _p(D), _p(geom9), _p(p), _p(out))
    return out


def gather_scatter(x, g2l, gstart):
    x = np.array(x, dtype=np.float64, copy=True)
    g2l = np.ascontiguousarray(g2l, dtype=np.int64)
    gstart = np.ascontiguousarray(gstart, dtype=np.int64)
    lib().so_gs(C.c_int64(gstart.size - 1), _p(g2l), _p(gstart), _p(x))
    return x


def cg(geom9, N, D, mask, g2l, gstart, rmult, b, tol, maxit, dinv=None):
    n = N + 1
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (geom9, D, mask, rmult, b)]
    geom9, D, mask, rmult, b = arrs
    g2l = np.ascontiguousarray(g2l, dtype=np.int64)
    gstart = np.ascontiguousarray(gstart, dtype=np.int64)
    x = np.empty_like(b)
    work = np.empty(4 * b.size)
    dv = np.ascontiguousarray(dinv, dtype=np.float64) if dinv is not None else None
    it = lib().so_cg(C.c_int64(geom9.shape[0]), C.c_int(n), _p(D), _p(geom9), _p(mask), C.c_int64(gstart.size - 1),
                     _p(g2l), _p(gstart), _p(rmult), _p(dv) if dv is not None else None, C.c_int(1 if dv is not None else 0),
                     _p(b), _p(x), C.c_double(tol), C.c_int(maxit), _p(work))
    return x, it
