"""A/B of tuning knobs on the BASELINE config-2 PCG solve: per-kernel times (CUDA events) and GDOF/s.

    python scripts/pcg_ab.py [--nel 32] [--order 7] knob=value[,knob=value] ...

Each positional argument is one variant (e.g. ``gs_pull=0``); "base" means the defaults."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nel", type=int, default=32)
    ap.add_argument("--order", type=int, default=7)
    ap.add_argument("--perturb", type=float, default=0.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("variants", nargs="*", default=["base"])
    args = ap.parse_args()
    import torch

    from sempy_b200 import _capi
    from sempy_b200.elliptic import elliptic_pcg

    bench.N_ORDER = args.order
    mesh, b, _ = bench.build_problem(args.nel, args.perturb)
    b = b if _capi.is_torch_cuda(b) else torch.from_numpy(b).cuda()
    ctx = mesh._context()
    lib = ctx.lib
    dof = mesh.get_num_elems() * mesh.Np
    names = [("ax", _capi.K_AX), ("gs", _capi.K_GS), ("update", _capi.K_UPDATE), ("precond", _capi.K_DIRECTION)]
    for var in args.variants:
        knobs = [] if var == "base" else [kv.split("=") for kv in var.split(",")]
        for k, v in knobs:
            _capi.check(lib.semb_set_tuning(k.encode(), int(v)))
        for _ in range(2):
            x, it = elliptic_pcg(mesh, b, tol=1e-8, maxit=1000)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 0
        for _ in range(args.reps):
            x, it = elliptic_pcg(mesh, b, tol=1e-8, maxit=1000)
            iters += it
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        _capi.check(lib.semb_profile_enable(ctx.handle, 1))
        _capi.check(lib.semb_profile_reset(ctx.handle))
        x, it = elliptic_pcg(mesh, b, tol=1e-8, maxit=1000)
        parts = []
        for nm, kid in names:
            tot, cnt = C.c_double(0), C.c_int64(0)
            _capi.check(lib.semb_profile_get(ctx.handle, kid, C.byref(tot), C.byref(cnt)))
            parts.append(f"{nm} {tot.value / max(it, 1) * 1e3:.1f}us")
        _capi.check(lib.semb_profile_enable(ctx.handle, 0))
        print(f"{var:28s} iters {it:3d}  {ms / iters * 1e3:7.1f} us/iter  {dof * iters / ms / 1e6:6.2f} GDOF/s   " + "  ".join(parts),
              flush=True)
        for k, v in knobs:  # back to defaults
            _capi.check(lib.semb_set_tuning(k.encode(), 1))


if __name__ == "__main__":
    main()
