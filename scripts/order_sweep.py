#!/usr/bin/env python
"""Dev tool: Ax and PCG-iteration throughput for several polynomial orders (same ~17 M local nodes)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sempy_b200 import _capi  # noqa: E402
from sempy_b200.elliptic import _ax, elliptic_pcg  # noqa: E402
from sempy_b200.mesh import box_mesh  # noqa: E402

if os.environ.get("SEMB_LIB"):  # a library variant built by scripts/build_variant.sh
    _capi._LIB_PATH = os.path.abspath(os.environ["SEMB_LIB"])
for kv in os.environ.get("KNOBS", "").split(","):
    if kv:
        _capi.check(_capi.lib().semb_set_tuning(kv.split("=")[0].encode(), int(kv.split("=")[1])))
cases = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]] or [(7, 32), (15, 16), (10, 23), (3, 64)]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for N, nel in cases:
    t0 = time.perf_counter()
    m = box_mesh(nel, perturb=0.3, seed=0)
    m.find_physical_coordinates(N)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    n = m.get_num_elems() * m.Np
    X, Y, Z = m.get_x(), m.get_y(), m.get_z()
    b = 3 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z) * m.get_mass() * m.get_jaco()
    b = m.apply_mask(m.dssum(b))
    tset = time.perf_counter() - t0
    p = torch.from_numpy(b).cuda()
    w = _ax(m, p, 0)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        ev0.record()
        for _ in range(10):
            w = _ax(m, p, 0)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) / 10)
    x, it = elliptic_pcg(m, p, tol=1e-30, maxit=30)
    torch.cuda.synchronize()
    pcg = 1e9
    for rep in range(4):
        ev0.record()
        x, it = elliptic_pcg(m, p, tol=1e-30, maxit=30)
        ev1.record()
        torch.cuda.synchronize()
        pcg = min(pcg, ev0.elapsed_time(ev1) / it)
    fs = 1 - ((N - 1) / (N + 1)) ** 3
    print(f"N={N:2d} nel={nel:3d} nodes={n/1e6:6.2f}M setup {tset:5.1f}s | Ax {best*1e3:8.1f} us {n/best/1e6:7.2f} GDOF/s "
          f"{64*n/best/1e6/6543.4*100:5.1f}% HBM | PCG iter {pcg*1e3:8.1f} us {n/pcg/1e6:6.2f} GDOF/s "
          f"{(160+20*fs)*n/pcg/1e6/6543.4*100:5.1f}% of PCG roofline", flush=True)
    del m, p, w, x
    torch.cuda.empty_cache()
