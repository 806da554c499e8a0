"""cg / pcg for any operator, vector work on the GPU.

Drop-in for sempy/iterative.py:4-85 (same signatures, stopping rules, verbose
print format, the ``assert b.ndim == 1`` of pcg).  The loop, the axpys and the
dots run in libsemb200 (semb_cg_generic); ``A`` and ``Minv`` are either

* objects with ``apply_device(d_in, d_out, n)`` (e.g. ``mesh.operator()``):
  applied on device pointers, nothing leaves the GPU, or
* plain Python callables on NumPy arrays, as in the reference's tests: the
  iterate is copied to the host, the callable applied, the result copied back.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from sempy_b200 import _capi


def _wrap(op, device):
    lib = _capi.lib()
    if op is None:
        return C.cast(None, _capi.APPLY_FN), None
    err = {}

    if hasattr(op, "apply_device"):

        def call(_user, d_in, d_out, n):
            try:
                op.apply_device(d_in, d_out, n)
                return 0
            except Exception as exc:  # surfaced after the C call returns
                err["exc"] = exc
                return 1

    else:

        def call(_user, d_in, d_out, n):
            try:
                h_in = np.empty(n, dtype=np.float64)
                _capi.check(lib.semb_memcpy(device, h_in.ctypes.data, d_in, 8 * n, 1))
                h_out = _capi.as_f64(op(h_in))
                if h_out.shape != (n,):
                    raise ValueError(f"operator returned shape {h_out.shape}, expected ({n},)")
                _capi.check(lib.semb_memcpy(device, d_out, h_out.ctypes.data, 8 * n, 0))
                return 0
            except Exception as exc:
                err["exc"] = exc
                return 1

    return _capi.APPLY_FN(call), err


def _solve(A, Minv, b, tol, maxit, verbose):
    lib = _capi.lib()
    _capi.require_gpu()
    device = _capi.current_device()
    b = _capi.as_f64(b)
    n = b.size
    x = np.empty(n, dtype=np.float64)
    hist = np.zeros((max(maxit, 1), 5), dtype=np.float64)
    niter = C.c_int(0)
    fa, ea = _wrap(A, device)
    fm, em = _wrap(Minv, device)
    rc = lib.semb_cg_generic(device, n, fa, None, fm, None, b.ctypes.data, x.ctypes.data, float(tol), int(maxit),
                             _capi.HOST, C.byref(niter), hist.ctypes.data)
    for e in (ea, em):
        if e and "exc" in e:
            raise e["exc"]
    _capi.check(rc)
    if verbose:
        if Minv is None:
            print("Initial rnorm={}".format(hist[0, 0] if niter.value else float(np.dot(b, b))))
        for it in range(niter.value):
            r0, r1, alpha, beta, pap = hist[it]
            print("niter={} r0={} r1={} alpha={} beta={} pap={}".format(it, r0, r1, alpha, beta, pap))
    return x, niter.value


def cg(A, b, tol=1e-12, maxit=100, verbose=0):
    return _solve(A, None, np.asarray(b).reshape(-1) if np.asarray(b).ndim != 1 else b, tol, maxit, verbose)


def pcg(A, Minv, b, tol=1e-8, maxit=100, verbose=0):
    assert b.ndim == 1
    return _solve(A, Minv, b, tol, maxit, verbose)
