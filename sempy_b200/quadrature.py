"""Gauss-Lobatto-Legendre nodes and weights (host input of the hot path).

Same algorithm and operation order as the reference (sempy/quadrature.py:8-40)
so that both paths see the same bits: interior nodes from the eigenvalues of
the Jacobi matrix (np.linalg.eig), weights from the Legendre recurrence.
"""

import numpy as np


def gauss_lobatto(p):
    n = p + 1
    z = np.zeros((n,))
    w = np.zeros((n,))
    z[0] = -1
    z[-1] = 1
    if p == 2:
        z[1] = 0
    elif p > 2:
        size = p - 1
        jac = np.zeros((size, size))
        idx = np.arange(1, size, dtype=np.float64)
        off = 0.5 * np.sqrt((idx * (idx + 2.0)) / ((idx + 0.5) * (idx + 1.5)))
        jac[np.arange(size - 1), np.arange(1, size)] = off
        jac[np.arange(1, size), np.arange(size - 1)] = off
        z[1:p] = np.sort(np.linalg.eig(jac)[0])
    w[0] = w[n - 1] = 2.0 / (p * n)
    for q in range(1, p):
        x = z[q]
        prev, cur = 1, x
        for deg in range(1, p):
            prev, cur = cur, x * cur * (2.0 * deg + 1.0) / (deg + 1.0) - prev * deg / (deg + 1.0)
        w[q] = 2.0 / (p * n * cur * cur)
    return z, w
