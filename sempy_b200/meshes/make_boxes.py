"""Writes box001/002/004/008.msh (gmsh 2.2 ASCII) with the vertices and hexahedra of the reference's
fixtures of the same names (sempy/meshes/box00x.msh): same vertex coordinates bit for bit (including the
~1e-12 offsets of box002) and same element orientations (box008's rotated elements), taken from the
golden fixtures tests/golden/box00x_N7.npz that tests/golden/make_golden.py recorded from the reference's
own reader.  Only what SemPy reads is written: $Nodes and the 8-node hexahedra (no physical names, no
boundary quads).

    python sempy_b200/meshes/make_boxes.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from sempy_b200.mesh import write_gmsh22  # noqa: E402

if __name__ == "__main__":
    for name in ("box001", "box002", "box004", "box008"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + "_N7.npz"))
        write_gmsh22(os.path.join(HERE, name + ".msh"), g["points"], g["hexes"])
        print(name, g["points"].shape, g["hexes"].shape)
