#!/usr/bin/env python
"""Dev tool: A/B of the L2 bulk prefetch in the Nq=8 Ax kernel on the 32^3, N=7 box.
(The wider variant sweep of round 1 is kept under profiles/experiments/.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_problem  # noqa: E402
from sempy_b200 import _capi  # noqa: E402
from sempy_b200.elliptic import _ax  # noqa: E402

nel = int(os.environ.get("NEL", "32"))
mesh, b = build_problem(nel, 0.3)
lib = _capi.lib()
n = b.size
p = torch.from_numpy(b).cuda()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for pf in (0, 1, 0, 1):
    _capi.check(lib.semb_set_tuning(b"ax_l2_prefetch", pf))
    w = _ax(mesh, p, 0)
    torch.cuda.synchronize()
    times = []
    for rep in range(5):
        ev0.record()
        for _ in range(20):
            w = _ax(mesh, p, 0)
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1) / 20)
    t = min(times)
    gbs = 64.0 * n / (t * 1e-3) / 1e9
    print(f"l2_prefetch={pf}: Ax {t*1e3:7.1f} us  {gbs:7.1f} GB/s  ({gbs/6543.4*100:5.1f}% of measured HBM)")
