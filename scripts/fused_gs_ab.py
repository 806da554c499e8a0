#!/usr/bin/env python
"""Dev tool: A/B of the experimental persistent operator + gather-scatter kernel (knob "ax_fused_gs")."""
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
for spec in sys.argv[1:] or ["0", "1:2368", "1:592", "1:4736", "0"]:
    on = int(spec.split(":")[0])
    lag = int(spec.split(":")[1]) if ":" in spec else 2368
    _capi.check(lib.semb_set_tuning(b"ax_fused_gs", on))
    _capi.check(lib.semb_set_tuning(b"ax_fused_lag", lag))
    mesh._dirty["gs"] = True   # the closure tables are built with the gather-scatter schedule
    mesh._context()
    x, it = elliptic_pcg(mesh, bd, tol=1e-8, maxit=1000)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(2):
        ev0.record()
        for _ in range(4):
            x, it = elliptic_pcg(mesh, bd, tol=1e-8, maxit=1000)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) / 4)
    if ref is None:
        ref = x.clone()
    print(f"ax_fused_gs={spec:>7s}: {best:7.3f} ms per solve ({it} iterations, {best/it*1e3:6.1f} us/iter), "
          f"{b.size*it/best/1e6:6.2f} GDOF/s, max |x - x_ref| / max |x_ref| = {((x-ref).abs().max()/ref.abs().max()).item():.1e}", flush=True)
