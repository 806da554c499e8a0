#!/usr/bin/env python
"""Short profiling target: the bench.py workload (32^3 elements, N=7, Jacobi-PCG) for a few iterations,
without the end-to-end / CPU-baseline legs, so that an ncu pass over it stays short."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_problem  # noqa: E402
from sempy_b200.elliptic import elliptic_pcg  # noqa: E402

mesh, b = build_problem(int(os.environ.get("NEL", "32")), 0.0)
bd = torch.from_numpy(b).cuda()
for _ in range(3):
    x, it = elliptic_pcg(mesh, bd, tol=1e-30, maxit=int(os.environ.get("ITERS", "8")))
torch.cuda.synchronize()
print("iterations per solve", it)
