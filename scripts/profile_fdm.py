#!/usr/bin/env python
"""Short ncu target: a few FDM-PCG iterations (NEL^3 elements, order ORDER)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sempy_b200.elliptic import elliptic_pcg_fdm  # noqa: E402

bench.N_ORDER = int(os.environ.get("ORDER", "7"))
mesh, bd, _ = bench.build_problem(int(os.environ.get("NEL", "32")), 0.3)
for _ in range(2):
    x, it = elliptic_pcg_fdm(mesh, bd, tol=1e-30, maxit=4)
torch.cuda.synchronize()
print("iterations", it)
