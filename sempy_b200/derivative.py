"""Lagrange differentiation matrix on GLL nodes (host input of the hot path).

Follows sempy/derivative.py:6-37 operation for operation; the matrix is
handed to the device once per mesh instead of being rebuilt per element
(sempy/gradient.py:7,40).
"""

import functools

import numpy as np

from sempy_b200.quadrature import gauss_lobatto


def lagrange_derivative_matrix(x):
    assert x.ndim == 1
    n = x.size
    bary = np.ones((n,), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            if j != i:
                bary[i] = bary[i] * (x[i] - x[j])
    bary = 1.0 / bary
    D = np.zeros((n, n))
    for i in range(n):
        D[i, :] = bary[i] * (x[i] - x)
        D[i, i] = 1
    for j in range(n):
        D[:, j] = D[:, j] / bary[j]
    D = 1.0 / D
    for i in range(n):
        D[i, i] = 0
        D[i, i] = -sum(D[i, :])
    return D


@functools.lru_cache(maxsize=None)
def _cached(p):
    z, _ = gauss_lobatto(p)
    D = lagrange_derivative_matrix(z)
    D.setflags(write=False)
    return D


def reference_derivative_matrix(p):
    return _cached(int(p)).copy()
