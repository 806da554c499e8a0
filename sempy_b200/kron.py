"""Tensor-product application (Sz (x) Sy (x) Sx) u on the GPU.

Drop-in for sempy/kron.py:4-43 (same names, argument order and return shape).
``order=None`` means C order (the reference's default crashes on numpy >= 2,
sempy/kron.py:24-26; every caller assumes C order).  ``U`` may hold a batch:
shape (nelem, mz*my*mx) is contracted element by element in one launch.
"""

from __future__ import annotations

import numpy as np

from sempy_b200 import _capi


def _dense(S):
    if hasattr(S, "toarray"):
        S = S.toarray()
    return _capi.as_f64(S)


def _kron_c(Sz, Sy, Sx, U2d):
    lib = _capi.lib()
    _capi.require_gpu()
    Sz, Sy, Sx = _dense(Sz), _dense(Sy), _dense(Sx)
    nz, mz = Sz.shape
    ny, my = Sy.shape
    nx, mx = Sx.shape
    U2d = _capi.as_f64(U2d)
    nelem = U2d.shape[0]
    if U2d.shape[1] != mz * my * mx:
        raise ValueError(f"kron: U has {U2d.shape[1]} entries per element, factors expect {mz * my * mx}")
    out = np.empty((nelem, nz * ny * nx), dtype=np.float64)
    _capi.check(
        lib.semb_kron(_capi.current_device(), Sz.ctypes.data, nz, mz, Sy.ctypes.data, ny, my, Sx.ctypes.data, nx, mx,
                      nelem, U2d.ctypes.data, out.ctypes.data, _capi.HOST)
    )
    return out


def kron(Sz, Sy, Sx, U, order=None):
    nx, mx = Sx.shape
    ny, my = Sy.shape
    nz, mz = Sz.shape
    U = np.asarray(U)
    batched = U.ndim == 2
    Ub = U if batched else U.reshape(1, -1)
    if order in (None, "C"):
        out = _kron_c(Sz, Sy, Sx, Ub)
    elif order == "F":
        # Fortran-ordered data is the C-ordered problem with the factor roles reversed
        out = _kron_c(Sx, Sy, Sz, Ub)
    else:
        raise ValueError("kron: order must be None, 'C' or 'F'")
    return out if batched else out.reshape((nx * ny * nz,))


def kron_2d(Sy, Sx, U, order=None):
    one = np.ones((1, 1))
    nx, _ = Sx.shape
    ny, _ = Sy.shape
    U = np.asarray(U)
    batched = U.ndim == 2
    Ub = U if batched else U.reshape(1, -1)
    if order in (None, "C"):
        out = _kron_c(one, Sy, Sx, Ub)
    elif order == "F":
        out = _kron_c(one, Sx, Sy, Ub)
    else:
        raise ValueError("kron_2d: order must be None, 'C' or 'F'")
    return out if batched else out.reshape((nx * ny,))
