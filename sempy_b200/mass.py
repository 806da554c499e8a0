"""Diagonal reference mass matrices (sempy/mass.py:6-33)."""

import numpy as np

from sempy_b200.quadrature import gauss_lobatto


def reference_mass_matrix_1d(p):
    _, w = gauss_lobatto(p)
    return np.diag(w)


def reference_mass_matrix_2d(p):
    _, w = gauss_lobatto(p)
    return (w[:, None] * w[None, :]).reshape(-1)


def reference_mass_matrix_3d(p):
    _, w = gauss_lobatto(p)
    # B[k,j,i] = (w_i w_j) w_k, i fastest
    return ((w[None, None, :] * w[None, :, None]) * w[:, None, None]).reshape(-1)
