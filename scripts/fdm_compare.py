#!/usr/bin/env python
"""Jacobi-PCG against FDM-PCG (and plain CG) on straight and deformed boxes: iterations and time to tol 1e-8.

    python scripts/fdm_compare.py [--nel 32] [--order 7]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sempy_b200.elliptic import elliptic_cg, elliptic_pcg, elliptic_pcg_fdm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nel", type=int, default=32)
ap.add_argument("--order", type=int, default=7)
args = ap.parse_args()
bench.N_ORDER = args.order
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for perturb in (0.0, 0.3):
    mesh, b, exact = bench.build_problem(args.nel, perturb)
    for name, solve in (("cg", elliptic_cg), ("jacobi", elliptic_pcg), ("fdm", elliptic_pcg_fdm)):
        x, it = solve(mesh, b, tol=1e-8, maxit=10000)
        torch.cuda.synchronize()
        ev0.record()
        x, it = solve(mesh, b, tol=1e-8, maxit=10000)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        err = (x - exact).abs().max().item()
        print(f"nel {args.nel} N {args.order} perturb {perturb}: {name:7s} iterations {it:5d}  solve {ms:9.2f} ms  "
              f"{ms / max(it, 1) * 1e3:7.1f} us/iter  max error vs exact {err:.2e}", flush=True)
    del mesh, b, exact
    torch.cuda.empty_cache()
