#!/usr/bin/env python
"""Short profiling target: the bench.py workload (NEL^3 elements, order ORDER, Jacobi-PCG) for a few
iterations plus a few plain operator / gather-scatter calls, without the end-to-end / CPU-baseline legs,
so that an ncu pass over it stays short.  Environment: NEL (32), ORDER (7), ITERS (8)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sempy_b200 import _capi  # noqa: E402
from sempy_b200.elliptic import _ax, elliptic_pcg  # noqa: E402

bench.N_ORDER = int(os.environ.get("ORDER", "7"))
mesh, bd, _ = bench.build_problem(int(os.environ.get("NEL", "32")), 0.0)
for _ in range(2):
    x, it = elliptic_pcg(mesh, bd, tol=1e-30, maxit=int(os.environ.get("ITERS", "8")))
w = _ax(mesh, bd, 0)
w = _ax(mesh, bd, _capi.AX_MASK | _capi.AX_GS)
torch.cuda.synchronize()
print("iterations per solve", it)
