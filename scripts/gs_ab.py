import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sempy_b200 import _capi
from sempy_b200.elliptic import _ax
for nel in (32, 48):
    mesh, b, _ = bench.build_problem(nel, 0.0)
    n = b.numel()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x = b.clone()
    for _ in range(3): mesh.dssum(x)
    torch.cuda.synchronize(); ev0.record()
    for _ in range(20): mesh.dssum(x)
    ev1.record(); torch.cuda.synchronize()
    gs = ev0.elapsed_time(ev1) / 20
    for _ in range(3): w = _ax(mesh, b, _capi.AX_MASK | _capi.AX_GS)
    torch.cuda.synchronize(); ev0.record()
    for _ in range(20): w = _ax(mesh, b, _capi.AX_MASK | _capi.AX_GS)
    ev1.record(); torch.cuda.synchronize()
    axgs = ev0.elapsed_time(ev1) / 20
    print(f"nel {nel}: dssum {gs*1e3:.1f} us  ax+mask+gs {axgs*1e3:.1f} us  frac {75.5625*n/axgs/1e6/6543.4:.3f}", flush=True)
    del mesh, b, x, w
    torch.cuda.empty_cache()
