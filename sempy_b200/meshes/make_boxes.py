"""Regenerate the bundled unit-cube meshes (1, 2, 4, 8 hexahedra) with the package's own generator."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sempy_b200.mesh import box_mesh_arrays, write_gmsh22  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
for name, dims in (("box001", (1, 1, 1)), ("box002", (2, 1, 1)), ("box004", (2, 2, 1)), ("box008", (2, 2, 2))):
    pts, hexes = box_mesh_arrays(*dims)
    write_gmsh22(os.path.join(here, name + ".msh"), pts, hexes)
    print(name, len(hexes))
