#!/usr/bin/env python
"""Dev tool: A/B of the chunked operator / gather-scatter pipeline inside semb_cg (knob "gs_chunks")."""
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
for nch in [int(v) for v in sys.argv[1:]] or [1, 2, 4, 8, 16, 1]:
    _capi.check(lib.semb_set_tuning(b"gs_chunks", nch))
    mesh._dirty["gs"] = True          # rebuild the gather-scatter schedule with the new chunking
    mesh._context()
    x, it = elliptic_pcg(mesh, bd, tol=1e-8, maxit=1000)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        ev0.record()
        for _ in range(4):
            x, it = elliptic_pcg(mesh, bd, tol=1e-8, maxit=1000)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) / 4)
    if ref is None:
        ref = x.clone()
    print(f"gs_chunks={nch:3d}: {best:7.3f} ms per solve ({it} iterations, {best/it*1e3:6.1f} us/iter), "
          f"{b.size*it/best/1e6:6.2f} GDOF/s, max |x - x_ref| / max |x_ref| = {((x-ref).abs().max()/ref.abs().max()).item():.1e}", flush=True)
