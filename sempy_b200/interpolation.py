"""Lagrange interpolation matrix in extended precision (sempy/interpolation.py:4-29)."""

import numpy as np


def lagrange(x_out, x_in, dtype=np.float64):
    assert x_out.ndim == x_in.ndim == 1
    n_in = x_in.size
    bary = np.ones((n_in,), dtype=np.longdouble)
    for i in range(n_in):
        for j in range(n_in):
            if j != i:
                bary[i] = bary[i] * (x_in[i] - x_in[j])
    bary = 1.0 / bary
    J = np.zeros((x_out.size, n_in), dtype=dtype)
    left = np.ones((n_in,), dtype=np.longdouble)
    right = np.ones((n_in,), dtype=np.longdouble)
    for row, x in enumerate(x_out):
        for j in range(1, n_in):
            left[j] = left[j - 1] * (x - x_in[j - 1])
            right[n_in - j - 1] = right[n_in - j] * (x - x_in[n_in - j])
        J[row, :] = bary * left * right
    return J
