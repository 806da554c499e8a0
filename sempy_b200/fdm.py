"""Fast-diagonalisation (FDM) preconditioner: host set-up for ``semb_set_fdm`` / SEMB_PRECOND_FDM.

Multi-element form of the single-element sketch in the reference's example.py:

* ``fdm_1d(N)`` follows ``fdm_d_inv`` (example.py:71-114): the 1-D stiffness ``Ah = Dh^T Bh Dh``
  (symmetrised, :75-76) and mass ``Bh`` on the GLL nodes, restricted (``R A R^T``, :82-90) to the nodes that
  carry unknowns, and the generalised eigenproblem ``A s = lambda B s`` (``sla.eigh(Ax, Bx)``, :92-100).
  Here the restriction drops BOTH end nodes: the block that is inverted exactly is the element interior.
* the device kernel (csrc/fdm_kernels.cuh) applies, per element,
  ``(S(x)S(x)S) diag(scale / (a_r l_i + a_s l_j + a_t l_k)) (S(x)S(x)S)^T`` -- example.py:155-171 (``fast_kron``)
  and :186-190 (``precon_fdm``: R, S^T.., dinv, S.., R^T) -- to the interior nodes and the assembled Jacobi
  diagonal to the element surfaces.

The per-element factors carry the element's edge lengths: for a box element L_r x L_s x L_t the operator is
J [a_r B(x)B(x)A + a_s B(x)A(x)B + a_t A(x)B(x)B] with a = (2/L)^2, J = L_r L_s L_t / 8; deformed elements
use the mean length of their four edges per direction (a separable approximation of the true block).
"""

from __future__ import annotations

import numpy as np

from sempy_b200 import _capi
from sempy_b200.derivative import reference_derivative_matrix
from sempy_b200.mass import reference_mass_matrix_1d

_GMSH_TO_LEX = [0, 1, 3, 2, 4, 5, 7, 6]


def fdm_1d(N):
    """(S, lam): eigenvectors (columns, S^T B S = I) and eigenvalues of the interior 1-D problem."""
    Bh = reference_mass_matrix_1d(N)
    Dh = reference_derivative_matrix(N)
    Ah = Dh.T @ Bh @ Dh
    Ah = 0.5 * (Ah + Ah.T)
    A, bdiag = Ah[1:-1, 1:-1], np.diag(Bh)[1:-1]
    if A.shape[0] == 0:
        return np.zeros((0, 0)), np.zeros(0)
    # B is diagonal: the generalised problem is the standard one for B^-1/2 A B^-1/2
    scale = 1.0 / np.sqrt(bdiag)
    lam, V = np.linalg.eigh(scale[:, None] * A * scale[None, :])
    return np.ascontiguousarray(scale[:, None] * V), np.ascontiguousarray(lam)


def element_scales(vx, vy, vz):
    """(E,4): a_r, a_s, a_t, scale from the vertex coordinates (E,8) in gmsh order."""
    v = np.stack([vx[:, _GMSH_TO_LEX], vy[:, _GMSH_TO_LEX], vz[:, _GMSH_TO_LEX]], axis=2).reshape(-1, 2, 2, 2, 3)
    Lr = np.linalg.norm(v[:, :, :, 1] - v[:, :, :, 0], axis=-1).mean(axis=(1, 2))
    Ls = np.linalg.norm(v[:, :, 1, :] - v[:, :, 0, :], axis=-1).mean(axis=(1, 2))
    Lt = np.linalg.norm(v[:, 1, :, :] - v[:, 0, :, :], axis=-1).mean(axis=(1, 2))
    return np.ascontiguousarray(np.stack([(2.0 / Lr) ** 2, (2.0 / Ls) ** 2, (2.0 / Lt) ** 2, 8.0 / (Lr * Ls * Lt)], axis=1))


def setup_fdm(mesh):
    """Hand S, lambda and the element factors of ``mesh`` to its device context (idempotent)."""
    ctx = mesh._context(need=("D", "geom", "gs"))
    if getattr(ctx, "fdm_ready", False):
        return ctx
    S, lam = fdm_1d(mesh.N)
    esc = element_scales(mesh.x, mesh.y, mesh.z)
    if S.size == 0:  # N = 1: no interior nodes, the preconditioner is point Jacobi
        S, lam = np.zeros(1), np.ones(1)
    _capi.check(ctx.lib.semb_set_fdm(ctx.handle, _capi.as_f64(S).ctypes.data, _capi.as_f64(lam).ctypes.data, esc.ctypes.data))
    ctx.fdm_ready = True
    return ctx
