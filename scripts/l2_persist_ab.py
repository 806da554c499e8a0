#!/usr/bin/env python
"""Dev tool: A/B of the L2 persisting window for Ap inside semb_cg (32^3 elements, N=7, Jacobi-PCG)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_problem  # noqa: E402
from sempy_b200 import _capi  # noqa: E402
from sempy_b200.elliptic import elliptic_pcg  # noqa: E402

mesh, b = build_problem(32, 0.0)
bd = torch.from_numpy(b).cuda()
lib = _capi.lib()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for mb in [int(v) for v in sys.argv[1:]] or [0, 32, 64, 96, 128, 0]:
    _capi.check(lib.semb_set_tuning(b"l2_persist_mb", mb))
    x, it = elliptic_pcg(mesh, bd, tol=1e-8, maxit=1000)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        ev0.record()
        for _ in range(5):
            x, it = elliptic_pcg(mesh, bd, tol=1e-8, maxit=1000)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) / 5)
    if ref is None:
        ref = x.clone()
    print(f"l2_persist_mb={mb:4d}: {best:7.3f} ms per solve ({it} iterations, {best/it*1e3:6.1f} us/iter), "
          f"{b.size*it/best/1e6:6.2f} GDOF/s, max |x - x0| = {(x-ref).abs().max().item():.1e}", flush=True)
