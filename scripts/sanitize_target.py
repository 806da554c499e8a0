#!/usr/bin/env python
"""Small target for compute-sanitizer (racecheck / memcheck): every kernel family of the hot path on a
deformed 5x4x3 box at N=7 and N=10 plus a 3x3x2 box at N=3 -- operator (plain and CG variants),
gather-scatter (push and pull forms), Jacobi and FDM PCG, mask, dot, the device set-up kernels.
With WORLD_SIZE > 1 (torchrun) the mesh is split over the ranks: halo kernels and peer-memory all-reduce."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sempy_b200 import _capi, distributed  # noqa: E402
from sempy_b200.elliptic import _ax, elliptic_cg, elliptic_pcg, elliptic_pcg_fdm  # noqa: E402
from sempy_b200.mesh import box_mesh  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
comm = None
if world > 1:
    import torch.distributed as dist

    os.environ.setdefault("SEMB_DEVICE", os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")
    comm = distributed.TorchComm()

for dims, N in (((5, 4, 3 * world), 7), ((3, 2, 2 * world), 10), ((3, 3, 2 * world), 3)):
    m = box_mesh(*dims, perturb=0.3, seed=1, comm=comm)
    m.find_physical_coordinates(N)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    n = m.get_num_elems() * m.Np
    rng = np.random.default_rng(3)
    b = m.apply_mask(m.dssum(rng.standard_normal(n)))
    w = _ax(m, b, 0)
    w = _ax(m, b, _capi.AX_MASK | _capi.AX_GS)
    for pull in (1, 0):
        _capi.check(_capi.lib().semb_set_tuning(b"gs_pull", pull))
        x, it = elliptic_pcg(m, b, tol=1e-8, maxit=12)
        x, it2 = elliptic_cg(m, b, tol=1e-8, maxit=6)
    _capi.check(_capi.lib().semb_set_tuning(b"gs_pull", 1))
    x, it3 = elliptic_pcg_fdm(m, b, tol=1e-8, maxit=8)
    print(f"dims {dims} N {N}: pcg {it} cg {it2} fdm {it3} |x| {np.abs(x).max():.3e}", flush=True)
if world > 1:
    dist.destroy_process_group()
print("sanitize target done")
