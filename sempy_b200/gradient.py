"""r/s/t derivative contractions of one element (or a batch) on the GPU.

Drop-in for sempy/gradient.py:6-70: ``gradient(U, n)`` returns (Ur, Us, Ut) and
``gradient_transpose(Wx, Wy, Wz, n)`` their transposed sum, with D built once
per order (the reference rebuilds it on each call, gradient.py:7,40).  Inputs
of shape (nelem, n^3) are processed as a batch in one launch.
"""

from __future__ import annotations

import numpy as np

from sempy_b200 import _capi
from sempy_b200.derivative import reference_derivative_matrix
from sempy_b200.kron import kron_2d


def gradient(U, n, D=None):
    lib = _capi.lib()
    _capi.require_gpu()
    D = _capi.as_f64(reference_derivative_matrix(n - 1) if D is None else D)
    U = np.asarray(U)
    batched = U.ndim == 2
    Ub = _capi.as_f64(U if batched else U.reshape(1, -1))
    if Ub.shape[1] != n * n * n:
        raise ValueError(f"gradient: expected {n * n * n} values per element, got {Ub.shape[1]}")
    out = [np.empty_like(Ub) for _ in range(3)]
    _capi.check(
        lib.semb_gradient(_capi.current_device(), n, D.ctypes.data, Ub.shape[0], Ub.ctypes.data, out[0].ctypes.data,
                          out[1].ctypes.data, out[2].ctypes.data, _capi.HOST)
    )
    if batched:
        return tuple(out)
    return tuple(o.reshape((n * n * n,)) for o in out)


def gradient_transpose(Wx, Wy, Wz, n, D=None):
    lib = _capi.lib()
    _capi.require_gpu()
    D = _capi.as_f64(reference_derivative_matrix(n - 1) if D is None else D)
    Wx = np.asarray(Wx)
    batched = Wx.ndim == 2
    W = [_capi.as_f64(np.asarray(w) if batched else np.asarray(w).reshape(1, -1)) for w in (Wx, Wy, Wz)]
    if any(w.shape != W[0].shape for w in W) or W[0].shape[1] != n * n * n:
        raise ValueError("gradient_transpose: the three inputs must hold n^3 values per element")
    out = np.empty_like(W[0])
    _capi.check(
        lib.semb_gradient_transpose(_capi.current_device(), n, D.ctypes.data, W[0].shape[0], W[0].ctypes.data,
                                    W[1].ctypes.data, W[2].ctypes.data, out.ctypes.data, _capi.HOST)
    )
    return out if batched else out.reshape((n * n * n,))


def gradient_2d(U, n):
    D = reference_derivative_matrix(n - 1)
    eye = np.eye(n)
    return kron_2d(eye, D, U), kron_2d(D, eye, U)


def gradient_transpose_2d(Wx, Wy, n):
    D = reference_derivative_matrix(n - 1)
    eye = np.eye(n)
    return kron_2d(eye, D.T, Wx) + kron_2d(D.T, eye, Wy)
