"""Stand-ins that let the UNMODIFIED reference (``/root/reference/sempy``) be
imported in the build container -- TEST INFRASTRUCTURE ONLY.

The reference imports two packages that are neither vendored nor installable
here (no network): ``meshio`` (sempy/mesh.py:4,105) and ``gslib_wrapper``
(sempy/mesh.py:6,437-445).  ``install()`` puts minimal replacements into
``sys.modules`` and rebinds ``sempy.mesh.kron`` to C order, because
``kron(..., order=None)`` crashes inside ``np.einsum(optimize=True)`` on
numpy >= 2 (sempy/kron.py:24-26).  Used only by tests/golden/make_golden.py
(which runs where /root/reference exists); never on the GPU box.
"""

from __future__ import annotations

import functools
import sys
import types

import numpy as np

from . import sem_oracle


class _MeshioResult:
    def __init__(self, points, hexes, quads):
        self.points = points
        self.cells = {"hexahedron": hexes, "quad": quads}


def _meshio_read(fname):
    pts, hexes, quads = sem_oracle.read_gmsh22(fname)
    return _MeshioResult(pts, hexes, quads)


class _GS:
    """gslib handle as SemPy drives it: GS(0).setup(int32 ids); gs(x, dtype, op)
    replaces x in place by the per-id sums (ids are 1-based, never 0)."""

    def __init__(self, comm):
        self._order = None

    def setup(self, ids):
        ids = np.asarray(ids)
        self._order = np.argsort(ids, kind="stable")
        sorted_ids = ids[self._order]
        brk = np.nonzero(np.diff(sorted_ids))[0] + 1
        self._start = np.concatenate(([0], brk, [ids.size]))

    def gs(self, x, dtype, op):
        x[:] = sem_oracle.gather_scatter(x, self._order, self._start)


def install(reference_root="/root/reference", gslib=None):
    """Make ``import sempy.mesh`` work against the reference checkout (or its copy oracle/_ref).
    ``gslib``: a module to serve as ``gslib_wrapper`` instead of the NumPy stand-in (the GPU tests pass
    ``sempy_b200.gslib_wrapper`` to run the unmodified reference mesh code on the CUDA gather-scatter)."""
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    meshio = types.ModuleType("meshio")
    meshio.read = _meshio_read
    sys.modules["meshio"] = meshio
    gsw = types.ModuleType("gslib_wrapper")
    gsw.GS = _GS
    gsw.gs_add = "gs_add"
    gsw.gs_double = "gs_double"
    sys.modules["gslib_wrapper"] = gslib if gslib is not None else gsw
    import sempy.kron
    import sempy.mesh

    if gslib is not None:  # sempy.mesh may have been imported before with the stand-in
        sempy.mesh.GS, sempy.mesh.gs_add, sempy.mesh.gs_double = gslib.GS, gslib.gs_add, gslib.gs_double
    else:
        sempy.mesh.GS, sempy.mesh.gs_add, sempy.mesh.gs_double = _GS, "gs_add", "gs_double"

    sempy.mesh.kron = functools.partial(sempy.kron.kron, order="C")
    return sempy
