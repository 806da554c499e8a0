"""Recipe for oracle/_ref: an UNMODIFIED copy of the reference's Python package -- TEST INFRASTRUCTURE ONLY.

The reference (SealUofI/SemPy, /root/reference) is pure Python: there is nothing to compile.  To time
its own NumPy path on the GPU box's host cores (bench.py, ``cpu_baseline_reference``), the files of the
hot path are copied verbatim, where they lie, into ``oracle/_ref/sempy`` -- a directory that is
git-ignored (never part of the history) but travels to the GPU box with the working tree, exactly like
a built ``.so``.  Nothing under oracle/_ref is imported by the product (``sempy_b200``); bench.py and the
tests import it only through ``oracle/ref_stubs.install`` (stand-ins for the reference's two
un-vendored imports, meshio and gslib_wrapper).

    python oracle/make_ref.py            # run where /root/reference exists (the build container)
"""

from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DEST = os.path.join(HERE, "_ref")
# the modules on the Poisson path and the mesh fixtures
FILES = ["elliptic.py", "gradient.py", "kron.py", "mesh.py", "iterative.py", "derivative.py", "quadrature.py", "mass.py",
         "interpolation.py", "types.py"]


def make(ref=REF, dest=DEST):
    src = os.path.join(ref, "sempy")
    if not os.path.isdir(src):
        return False
    out = os.path.join(dest, "sempy")
    os.makedirs(os.path.join(out, "meshes"), exist_ok=True)
    for fn in FILES:
        shutil.copyfile(os.path.join(src, fn), os.path.join(out, fn))
    # the reference's driver (BASELINE config 1), run unmodified against the product by a GPU test
    shutil.copyfile(os.path.join(ref, "poisson.py"), os.path.join(dest, "poisson.py"))
    for fn in os.listdir(os.path.join(src, "meshes")):
        if fn.endswith(".msh"):
            shutil.copyfile(os.path.join(src, "meshes", fn), os.path.join(out, "meshes", fn))
    return True


def available(dest=DEST):
    return os.path.isfile(os.path.join(dest, "sempy", "elliptic.py"))


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref:", "copied from " + REF if ok else "reference not present, nothing done")
    sys.exit(0)
