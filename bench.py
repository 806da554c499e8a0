#!/usr/bin/env python
"""Benchmark of the SemPy Poisson hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one Jacobi-PCG solve (tol 1e-8) of the 3-D Poisson problem on a
box mesh of 32^3 elements at N=7 (BASELINE config 2; 16.8 M local
nodes per GPU), i.e. ~50 passes of the hot path (fused Ax + mask + local pAp,
gather-scatter, fused vector updates).  ``value`` is local DOF x iterations per
second with everything resident in HBM; ``e2e`` is the same solve called through
the drop-in API with HOST (pinned) NumPy buffers, host<->device copies timed.
``roofline`` describes the dominant kernel (the fused Ax), timed live with CUDA
events inside the timed region.  ``--impl reference`` times the CPU oracle (a
C/OpenMP restatement of the reference's algorithm, all host cores) on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ = 8
N_ORDER = 7
F_SHARED = 1.0 - ((N_ORDER - 1) / (N_ORDER + 1)) ** 3
BYTES_AX = 64.0                      # u 8 + six G 48 + w 8 per local node
BYTES_GS = 20.0 * F_SHARED           # shared nodes: read 8 + index 4 + write 8
BYTES_UPDATE_PCG = 49.0 + 8.0        # x,p,r,Ap reads 32 + 1-byte multiplicity + x,r writes 16 (+ dinv 8)
BYTES_DIR_PCG = 24.0 + 8.0           # r,p reads + p write (+ dinv)
TOL = 1e-8


def ncu_traffic():
    """DRAM bytes per launch of the fused Ax kernel from the committed ncu capture
    (profiles/ax_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    path = os.path.join(ROOT, "profiles", "ax_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons; samples are kept with their
    timestamps and only those inside the timed window are summarised."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin, t_end):
        """Summarise the samples taken between the two time.time() stamps."""
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx, pw, all_sm = [], set(), None, [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    clk, mxv, pwv = float(f[1]), float(f[2]), float(f[3])
                except ValueError:
                    continue
                mx = mxv
                all_sm.append(clk)
                if t_begin - 0.02 <= ts <= t_end + 0.02:
                    sm.append(clk)
                    pw.append(pwv)
                    for nm, val in zip(names, f[4:8]):
                        if val.lower().startswith("active"):
                            reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["power_w_max"] = max(pw)
            out["samples"] = len(sm)
        elif all_sm:
            out["sm_mhz"] = float(np.median(all_sm))
            out["samples"] = 0
            out["note"] = "no sample fell inside the timed window; median over the whole run"
        out["sm_max_mhz"] = mx
        out["reasons"] = sorted(reasons)
        return out


def build_problem(nel, perturb=0.0, comm=None, strong=0):
    """Box mesh of nel^3 elements (deformed if perturb > 0) at N=7 plus the manufactured RHS of
    the reference's poisson.py:27-40.  Returns (mesh, b_host).  strong > 0: a fixed strong^3-element box
    shared by all ranks (BASELINE config 4: --strong 64 --perturb 0.3)."""
    from sempy_b200.mesh import box_mesh

    if strong > 0:
        m = box_mesh(strong, perturb=perturb, seed=0, comm=comm)
    elif comm is not None and comm.size > 1:
        # weak scaling: nel^3 elements per GPU as z-slabs of a (2 nel) x (2 nel) x (nel/4 * ranks) box,
        # which ends at the (2 nel)^3 cube on 8 GPUs (64^3 for nel = 32)
        m = box_mesh(2 * nel, 2 * nel, (nel // 4) * comm.size, perturb=perturb, seed=0, comm=comm)
    else:
        m = box_mesh(nel, perturb=perturb, seed=0)
    m.find_physical_coordinates(N_ORDER)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    X, Y, Z = m.get_x(), m.get_y(), m.get_z()
    b = 3 * np.pi * np.pi * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
    b = b * m.get_mass() * m.get_jaco()
    b = m.dssum(b)
    b = m.apply_mask(b)
    return m, b


def oracle_problem(mesh, b):
    """Arrays of the SAME problem for the CPU oracle (the mesh set-up is input preparation, done once
    through the product's mesh class; the timed CPU path is the oracle's)."""
    from sempy_b200.derivative import reference_derivative_matrix

    D = np.ascontiguousarray(reference_derivative_matrix(N_ORDER))
    g2l, gs = mesh.get_global_to_local_map()
    return dict(geom=np.ascontiguousarray(mesh.get_geom()), D=D, g2l=g2l, gs=gs, rmult=mesh.get_rmult().copy(),
                mask=mesh.mask.copy(), b=np.array(b, copy=True), dof=mesh.get_num_elems() * mesh.Np)


CPU_ITERS = 10  # PCG iterations per CPU step: a bounded sample of the ~45-iteration solve


def time_oracle(prob, iters, steps, warmup):
    """C/OpenMP oracle (oracle/c/sem_oracle.c), Jacobi-PCG with a fixed iteration count per step on all
    host threads; returns (GDOF/s, ms per step, threads)."""
    from oracle import c_oracle as co
    from oracle import sem_oracle as so

    geom, D = prob["geom"], prob["D"]
    diag = so.gather_scatter(so.jacobi_diagonal(geom, N_ORDER, D), prob["g2l"], prob["gs"])
    dinv = prob["mask"] / diag
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        _, it = co.cg(geom, N_ORDER, D, prob["mask"], prob["g2l"], prob["gs"], prob["rmult"], prob["b"], 1e-30, iters, dinv=dinv)
        t1 = time.perf_counter()
        assert it == iters
        if s >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return prob["dof"] * iters * steps / total / 1e9, total / steps * 1e3, co.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh, b = build_problem(args.nel, args.perturb)
    prob = oracle_problem(mesh, b)
    warm = min(args.warmup, 1)
    val, ms, threads = time_oracle(prob, CPU_ITERS, args.steps, warm)
    sample = (f"C/OpenMP restatement of the reference algorithm (oracle/c/sem_oracle.c), {threads} threads, Jacobi-PCG on the "
              f"full {args.nel}^3-element N={N_ORDER} mesh, {CPU_ITERS} iterations per step x {args.steps} steps")
    line = {
        "impl": "reference", "metric": f"FP64 Jacobi-PCG Poisson throughput, N={N_ORDER} (local DOF x iterations / s)",
        "value": val, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"3D Poisson, {'deformed' if args.perturb else 'straight'} box {args.nel}^3 elements, N={N_ORDER}, Jacobi-PCG tol 1e-8 (BASELINE config 2)",
                   "sample": f"{CPU_ITERS} PCG iterations per step on the full mesh",
                   "note": "the reference is pure Python (~1 MDOF/s, one core) and cannot be shipped to the GPU box; "
                           "this arm times its C/OpenMP restatement on all host cores"},
        "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch

    from sempy_b200 import _capi, distributed
    from sempy_b200.elliptic import elliptic_ax, elliptic_pcg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    comm = None
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = distributed.TorchComm()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def over_ranks(v, op):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    peak, peak_kind = measured_peaks()
    nel = args.nel
    t_setup = time.perf_counter()
    mesh, b_host = build_problem(nel, args.perturb, comm, args.strong)
    t_setup = time.perf_counter() - t_setup
    dof = mesh.get_num_elems() * mesh.Np
    ctx = mesh._context()
    lib = ctx.lib

    b_dev = torch.from_numpy(b_host).cuda()
    # pinned host buffers for the end-to-end arm
    b_pin = torch.empty(dof, dtype=torch.float64).pin_memory()
    b_pin.copy_(torch.from_numpy(b_host))
    x_pin = torch.empty(dof, dtype=torch.float64).pin_memory()
    b_pin_np, x_pin_np = b_pin.numpy(), x_pin.numpy()

    def solve_dev():
        return elliptic_pcg(mesh, b_dev, tol=TOL, maxit=1000)

    def solve_host():
        return elliptic_pcg(mesh, b_pin_np, tol=TOL, maxit=1000, out=x_pin_np)

    # ---- device-resident arm -------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        x_dev, niter = solve_dev()
    sync_all()
    # inside the timed region only the dominant kernel (the fused Ax) is bracketed by events;
    # the other kernels are timed in a second, untimed pass below (every bracket costs two event records)
    _capi.check(lib.semb_profile_enable(ctx.handle, 1 | ((1 << _capi.K_AX) << 1)))
    _capi.check(lib.semb_profile_reset(ctx.handle))
    launches0 = lib.semb_launch_count(ctx.handle)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    wall0 = time.time()
    ev0.record()
    iters_total = 0
    for _ in range(args.steps):
        x_dev, niter = solve_dev()
        iters_total += niter
    ev1.record()
    sync_all()
    clocks = sampler.stop(wall0, time.time())
    ms_total = over_ranks(ev0.elapsed_time(ev1), "MAX")
    dof_all = over_ranks(dof, "SUM")
    launches = lib.semb_launch_count(ctx.handle) - launches0
    import ctypes as C

    def read_profile(names):
        out = {}
        for name, kid in names:
            tot, cnt = C.c_double(0), C.c_int64(0)
            _capi.check(lib.semb_profile_get(ctx.handle, kid, C.byref(tot), C.byref(cnt)))
            out[name] = (tot.value, cnt.value)
        return out

    kern = read_profile([("ax", _capi.K_AX)])
    # second pass (not part of `value`): all four kernel kinds bracketed
    _capi.check(lib.semb_profile_enable(ctx.handle, 1))
    _capi.check(lib.semb_profile_reset(ctx.handle))
    ev0.record()
    _, iters_pass2 = solve_dev()
    ev1.record()
    sync_all()
    ms_pass2 = ev0.elapsed_time(ev1)
    kern2 = read_profile([("ax", _capi.K_AX), ("gs", _capi.K_GS), ("update", _capi.K_UPDATE), ("direction", _capi.K_DIRECTION)])
    _capi.check(lib.semb_profile_enable(ctx.handle, 0))
    value = dof_all * iters_total / (ms_total * 1e-3) / 1e9
    # per CG iteration: with a halo the Ax of one iteration is two launches (interface elements, then the rest)
    ax_ms = kern["ax"][0] / max(iters_total, 1)
    ax_gbs = BYTES_AX * dof / (ax_ms * 1e-3) / 1e9

    def kernel_entry(name, bytes_per_dof):
        tot, cnt = kern2[name]
        ms = tot / max(iters_pass2, 1)
        gbs = bytes_per_dof * dof / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"ms_per_iteration": ms, "launches": cnt, "bytes_per_dof": bytes_per_dof, "achieved_gbs": gbs, "frac": gbs / peak,
                "share_of_step": tot / ms_pass2}

    # ---- fused Ax + mask + gather-scatter alone (the north-star 70 % target) ----
    from sempy_b200.elliptic import _ax

    p_dev = x_dev.clone()
    for _ in range(3):
        w = _ax(mesh, p_dev, _capi.AX_MASK | _capi.AX_GS)
    sync_all()
    reps = 20
    ev0.record()
    for _ in range(reps):
        w = _ax(mesh, p_dev, _capi.AX_MASK | _capi.AX_GS)
    ev1.record()
    sync_all()
    axgs_ms = over_ranks(ev0.elapsed_time(ev1), "MAX") / reps
    axgs_gdofs = dof_all / (axgs_ms * 1e-3) / 1e9
    axgs_gbs = (BYTES_AX + BYTES_GS) * dof / (axgs_ms * 1e-3) / 1e9

    # ---- end-to-end arm: host buffers through the public API ----------------
    for _ in range(2):
        solve_host()
    sync_all()
    t0 = time.perf_counter()
    e_iters = 0
    for _ in range(args.steps):
        _, it = solve_host()
        e_iters += it
    sync_all()
    e_s = over_ranks(time.perf_counter() - t0, "MAX")
    e2e_val = dof_all * e_iters / e_s / 1e9
    # sanity: same answer on both arms and a solved problem
    assert np.allclose(x_pin_np, x_dev.cpu().numpy(), rtol=0, atol=1e-12)

    if world > 1:
        all_clocks = comm.allgather(clocks)
        reasons = sorted({r for c in all_clocks for r in c.get("reasons", [])})
        sm = [c["sm_mhz"] for c in all_clocks if c.get("sm_mhz")]
        clocks = {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": reasons,
                  "per_rank_sm_mhz": sm}
    unique_local = int(mesh.global_start.size - 1)
    if rank != 0:
        dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle on a bounded sample of the same workload (N = 1 only) ---
    cpu_baseline = None
    if world == 1 and dof <= 20e6:  # the CPU arm copies G as (E,3,3,Np): keep it to the default size
        t_cpu0 = time.perf_counter()
        prob = oracle_problem(mesh, b_host)
        cpu_val, _, threads = time_oracle(prob, CPU_ITERS, 2, 1)
        cpu_secs = time.perf_counter() - t_cpu0
        cpu_baseline = {"value": cpu_val, "unit": "GDOF/s", "cores": threads, "kind": "port",
                        "sample": "C/OpenMP restatement of the reference algorithm (oracle/c/sem_oracle.c), %d threads, "
                                  "Jacobi-PCG on the same full mesh, %d iterations x 2 steps (%.0f s); the Python "
                                  "reference itself runs ~1 MDOF/s on one core (BASELINE.md)" % (threads, CPU_ITERS, cpu_secs)}
        del prob
    if args.strong > 0:
        workload = (f"3D Poisson, {'deformed' if args.perturb else 'straight'} box {args.strong}^3 elements shared by "
                    f"{world} GPU(s) in z-slabs, N={N_ORDER}, Jacobi-PCG tol 1e-8 (strong scaling, BASELINE config 4)")
    elif world == 1:
        workload = (f"3D Poisson, {'deformed' if args.perturb else 'straight'} box {nel}^3 elements, N={N_ORDER}, "
                    "Jacobi-PCG tol 1e-8 (BASELINE config 2)")
    else:
        workload = (f"3D Poisson, {'deformed' if args.perturb else 'straight'} box {2 * nel}x{2 * nel}x{(nel // 4) * world} "
                    f"elements in z-slabs of {nel ** 3} elements per GPU, N={N_ORDER}, Jacobi-PCG tol 1e-8 "
                    "(weak scaling of BASELINE config 2; the 8-GPU mesh is the 64^3 box of config 4)")

    line = {
        "metric": f"FP64 Jacobi-PCG Poisson throughput, N={N_ORDER} (local DOF x iterations / s)",
        "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong > 0 else "weak",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload,
                   "elements_per_gpu": mesh.get_num_elems(), "local_dof_per_gpu": dof, "local_dof_total": int(dof_all),
                   "unique_dof_rank0": unique_local,
                   "iterations_per_step": iters_total / args.steps, "cache": "inputs larger than L2 (G alone is %.0f MB per GPU)" % (48.0 * dof / 1e6),
                   "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, one z-slab of elements per GPU, NCCL halo exchange overlapped with interior Ax + all-reduced dots",
                   "setup_s": t_setup},
        "cg_iters_per_s": iters_total / (ms_total * 1e-3),
        "ax_gs": {"gdofs": axgs_gdofs, "ms": axgs_ms, "bytes_per_dof": BYTES_AX + BYTES_GS, "achieved_gbs": axgs_gbs,
                  "frac_of_hbm_roofline": axgs_gbs / peak, "what": "fused Ax+mask kernel followed by gather-scatter, device resident"},
        "roofline": {"bound": "hbm", "kernel": ("ax8_kernel" if N_ORDER == 7 else "axl_kernel" if N_ORDER >= 8 else "ax_generic_kernel") + " (fused Ax + mask + local pAp)", "achieved": ax_gbs, "peak": peak,
                     "peak_source": peak_kind, "unit": "GB/s", "frac": ax_gbs / peak,
                     "traffic": (ncu_traffic() or {}).get("dram_bytes_per_launch") if world == 1 and nel == 32 and N_ORDER == 7 else None,
                     "traffic_source": (ncu_traffic() or {}).get("source"), "algorithmic_bytes_per_launch": BYTES_AX * dof,
                     "bytes_per_dof": BYTES_AX, "avg_ms": ax_ms, "avg_ms_is": "per CG iteration", "launches_timed": kern["ax"][1],
                     "share_of_step": kern["ax"][0] / ms_total, "gdofs": dof / (ax_ms * 1e-3) / 1e9},
        "kernels": {"ax": kernel_entry("ax", BYTES_AX), "gs": kernel_entry("gs", BYTES_GS),
                    "update": kernel_entry("update", BYTES_UPDATE_PCG), "direction": kernel_entry("direction", BYTES_DIR_PCG)},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * int(dof_all), "d2h_bytes_per_step": 8 * int(dof_all),
                "ms_per_step": e_s / args.steps * 1e3, "api": "sempy_b200.elliptic.elliptic_pcg(mesh, b_host) -> x_host"},
        "gpu_launches": int(launches),
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nel", type=int, default=32, help="elements per side per GPU (default: BASELINE config 2)")
    ap.add_argument("--perturb", type=float, default=0.0, help="vertex displacement in units of h (0.3 = the deformed mesh of config 4)")
    ap.add_argument("--strong", type=int, default=0, help="strong scaling: one box of STRONG^3 elements split over all GPUs (config 4: 64)")
    ap.add_argument("--order", type=int, default=7, help="polynomial order N (BASELINE config 3: --order 15 --nel 16 or 32)")
    args = ap.parse_args()
    global N_ORDER, NQ, F_SHARED, BYTES_GS
    N_ORDER, NQ = args.order, args.order + 1
    F_SHARED = 1.0 - ((N_ORDER - 1) / (N_ORDER + 1)) ** 3
    BYTES_GS = 20.0 * F_SHARED
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
