#!/usr/bin/env python
"""Benchmark of the SemPy Poisson hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--nel n] [--order N] [--perturb d] [--strong n]

A *step* is one Jacobi-PCG solve (tol 1e-8) of the 3-D Poisson problem of the reference's poisson.py on
a box mesh: by default 32^3 elements at N=7 (BASELINE config 2; 16.8 M local nodes per GPU), i.e. ~45
passes of the hot path (fused direction update + Ax + mask + local pAp, gather-scatter, fused vector
update with the dot products).  ``value`` is local DOF x iterations per second with everything resident
in HBM; ``e2e`` is the same solve through the reference-facing call ``elliptic_pcg(mesh, b) -> x`` with
HOST NumPy arrays, host<->device copies timed.  ``roofline`` describes the dominant kernel, timed live
with CUDA events inside the timed region (around each of its launches in every fourth timed step).  Other BASELINE configs: ``--order 15 --nel 32`` (config 3),
``--strong 64 --perturb 0.3`` (config 4), ``--nel 48`` on 1/2/4/8 GPUs (config 5).

``--impl reference`` times the reference's algorithm on the host cores: its C/OpenMP restatement
(oracle/c/sem_oracle.c) with all threads on the same mesh, set up by the oracle's own NumPy code --
nothing of sempy_b200 is imported on that arm.
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

N_ORDER = 7
TOL = 1e-8
MAXIT = 10000  # as the reference driver (poisson.py:42)
CPU_ITERS = 10  # PCG iterations per CPU step: a bounded sample of the ~45-iteration solve


def f_shared(N):
    return 1.0 - ((N - 1) / (N + 1)) ** 3


# Algorithmic bytes per local node (Ax 64, gather-scatter 20 per shared node, CG vector work)
BYTES_AX = 64.0            # u 8 + six G 48 + w 8
BYTES_AX_DIR = 88.0        # + the folded direction update p <- dinv r + beta p: r 8, dinv 8, p written 8
BYTES_UPDATE = 49.0        # r, Ap reads 16 + dinv 8 + 1-byte multiplicity + r write 8, and every SECOND iteration
                           # p_{k-1}, p_k, x reads 24 + x write 8 (the solution advances two terms at a time)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(kernel)
    except Exception:
        return None


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


def mesh_dims(nel, world, strong):
    """Element counts of the whole box.  Weak scaling: nel^3 elements per GPU as z-slabs of a
    (2 nel) x (2 nel) x (nel/4 * ranks) box, which ends at the (2 nel)^3 cube on 8 GPUs (BASELINE config 5:
    nel = 48 -> 96^3).  Strong scaling: one strong^3 box for every rank count (config 4: 64)."""
    if strong > 0:
        return strong, strong, strong
    if world > 1:
        return 2 * nel, 2 * nel, (nel // 4) * world
    return nel, nel, nel


def workload_name(args, world):
    nx, ny, nz = mesh_dims(args.nel, world, args.strong)
    kind = "deformed" if args.perturb else "straight"
    cfg = ("BASELINE config 4, strong scaling" if args.strong > 0 else
           "BASELINE config 3" if N_ORDER == 15 else
           "BASELINE config 5, weak scaling" if args.nel == 48 else
           "BASELINE config 2" if (world == 1 and args.nel == 32 and N_ORDER == 7) else
           "weak scaling of BASELINE config 2" if N_ORDER == 7 and args.nel == 32 else "custom size")
    return f"3D Poisson, {kind} box {nx}x{ny}x{nz} elements, N={N_ORDER}, Jacobi-PCG tol 1e-8 ({cfg})"


def build_problem(nel, perturb=0.0, comm=None, strong=0):
    """Box mesh at order N_ORDER plus the manufactured RHS of the reference's poisson.py:27-40, everything
    built on the device.  Returns (mesh, b, exact) with b and exact as CUDA tensors."""
    import torch

    from sempy_b200.mesh import box_mesh

    world = comm.size if comm is not None else 1
    nx, ny, nz = mesh_dims(nel, world, strong)
    stamps = [("start", time.perf_counter())]

    def stamp(name):
        torch.cuda.synchronize()
        stamps.append((name, time.perf_counter()))

    m = box_mesh(nx, ny, nz, perturb=perturb, seed=0, comm=comm)
    stamp("vertex_mesh")
    m.find_physical_coordinates(N_ORDER)
    stamp("coordinates")
    m.establish_global_numbering()
    stamp("numbering_gs_halo")
    m.calc_geometric_factors()
    stamp("geometric_factors")
    m.setup_mask()
    stamp("mask")
    m.setup_breakdown = {b[0]: round(b[1] - a[1], 3) for a, b in zip(stamps[:-1], stamps[1:])}
    if getattr(m, "halo_timings", None):
        m.setup_breakdown["halo"] = m.halo_timings
    dev = torch.device("cuda", torch.cuda.current_device())
    X, Y, Z, J = (torch.as_tensor(m.device_array(k), device=dev).reshape(-1) for k in ("x", "y", "z", "jaco"))
    exact = torch.sin(np.pi * X) * torch.sin(np.pi * Y) * torch.sin(np.pi * Z)
    B = torch.from_numpy(np.array(m.mass[0], copy=True)).to(dev)
    b = (3 * np.pi * np.pi * exact).reshape(-1, m.Np) * B * J.reshape(-1, m.Np)
    b = b.reshape(-1).contiguous()
    m.dssum(b)
    m.apply_mask(b)
    return m, b, exact


# ---------------------------------------------------------------------------
# CPU arms (oracle only; nothing of sempy_b200 is imported here)
# ---------------------------------------------------------------------------
def oracle_problem(nx, ny, nz, perturb, N):
    """The same problem set up by the oracle's own NumPy restatement of the reference's mesh code."""
    from oracle import sem_oracle as so

    pts, hexes = so.box_mesh(nx, ny, nz, perturb=perturb, seed=0)
    xe, ye, ze = so.physical_coordinates(*so.element_vertices(pts, hexes), N)
    _, g2l, gs = so.global_numbering(xe, ye, ze)
    geom, jaco, mass = so.geometric_factors(xe, ye, ze, N)
    rmult, mask = so.multiplicity_inverse(g2l, gs), so.dirichlet_mask(xe, ye, ze)
    b = 3 * np.pi ** 2 * np.sin(np.pi * xe) * np.sin(np.pi * ye) * np.sin(np.pi * ze) * mass * jaco
    b = so.gather_scatter(b.reshape(-1), g2l, gs) * mask
    D = np.ascontiguousarray(so.reference_derivative_matrix(N))
    return dict(geom=np.ascontiguousarray(geom), D=D, g2l=g2l, gs=gs, rmult=rmult, mask=mask, b=b, dof=xe.size)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


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


def time_numpy_reference(n=8, N=7, cg_iters=3):
    """The UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py; meshio / gslib_wrapper replaced
    by oracle/ref_stubs.py) on an n^3-element box: Ax + mask + dssum and elliptic_cg, one host core."""
    from oracle import make_ref, ref_stubs
    from oracle import sem_oracle as so

    if not make_ref.available():
        return None
    ref_stubs.install(make_ref.DEST)
    from sempy.elliptic import elliptic_ax, elliptic_cg
    from sempy.mesh import Mesh

    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "box.msh")
        so.write_gmsh22(path, *so.box_mesh(n, n, n))
        t0 = time.perf_counter()
        mesh = Mesh(path)
        mesh.find_physical_coordinates(N)
        mesh.establish_global_numbering()
        mesh.calc_geometric_factors()
        mesh.setup_mask()
        t_setup = time.perf_counter() - t0
    dof = mesh.get_num_elems() * mesh.Np
    X, Y, Z = mesh.get_x(), mesh.get_y(), mesh.get_z()
    b = 3 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z) * mesh.get_mass() * mesh.get_jaco()
    b = mesh.apply_mask(mesh.dssum(b))
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        mesh.dssum(mesh.apply_mask(elliptic_ax(mesh, b)))
    t_ax = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    _, it = elliptic_cg(mesh, b, tol=1e-30, maxit=cg_iters)
    t_cg = (time.perf_counter() - t0) / max(it, 1)
    return {"value": dof / t_cg / 1e9, "unit": "GDOF/s", "cores": 1, "kind": "reference",
            "ax_gs_gdofs": dof / t_ax / 1e9, "setup_s": t_setup,
            "sample": f"unmodified sempy (NumPy) from oracle/_ref on a {n}^3-element N={N} box ({dof} local DOF): "
                      f"elliptic_cg, {it} iterations (unpreconditioned; the reference has no mesh-aware PCG), "
                      f"and elliptic_ax + apply_mask + dssum x {reps}; one core (per-element Python loop); "
                      "dssum is the stub of oracle/ref_stubs.py, not gslib"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    use_all_host_threads()
    # bounded sample: the N-GPU mesh capped at one GPU's share (the CPU arm has one host to run on)
    nx, ny, nz = mesh_dims(args.nel, 1, 0) if args.strong == 0 else (args.strong,) * 3
    if args.strong > 0 or N_ORDER > 7:
        # 134 M nodes need > 10 GB for G alone: sample an eighth of the box (same element shape)
        nz = max(nz // 8, 1)
    t0 = time.perf_counter()
    prob = oracle_problem(nx, ny, nz, args.perturb, N_ORDER)
    t_setup = time.perf_counter() - t0
    warm = min(args.warmup, 1)
    val, ms, threads = time_oracle(prob, CPU_ITERS, args.steps, warm)
    sample = (f"C/OpenMP restatement of the reference algorithm (oracle/c/sem_oracle.c), {threads} threads, Jacobi-PCG on a "
              f"{nx}x{ny}x{nz}-element N={N_ORDER} box ({prob['dof']} local DOF) set up by the oracle's NumPy code "
              f"({t_setup:.0f} s), {CPU_ITERS} iterations per step x {args.steps} steps")
    line = {
        "impl": "reference", "metric": f"FP64 Jacobi-PCG Poisson throughput, N={N_ORDER} (local DOF x iterations / s)",
        "value": val, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong > 0 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "sample": sample,
                   "note": "throughput per host, independent of N: the CPU arm runs on rank 0's host cores only"},
        "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        ref = time_numpy_reference()
        if ref is not None:
            line["cpu_baseline_reference"] = ref
    except Exception as exc:  # the NumPy reference is an extra; never lose the line over it
        line["cpu_baseline_reference"] = {"unavailable": repr(exc)}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def small_vs_oracle(comm, world, rank):
    """A deformed 4 x 4 x (2 ranks) box at N=7 solved by all ranks and compared with the CPU oracle on the
    WHOLE mesh (rank 0): iteration count +-1, solution 1e-10 relative (the BASELINE tolerances)."""
    from oracle import sem_oracle as so
    from sempy_b200.elliptic import elliptic_pcg
    from sempy_b200.mesh import box_mesh_arrays, Mesh

    N = 7
    dims = (4, 4, 2 * world)
    pts, hexes = box_mesh_arrays(*dims, perturb=0.3, seed=2)
    m = Mesh.from_arrays(pts, hexes, comm=comm)
    m.find_physical_coordinates(N)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    E, Np = len(hexes), m.Np
    gx, gy, gz = so.physical_coordinates(*so.element_vertices(pts, hexes), N)
    _, g2l, gs = so.global_numbering(gx, gy, gz)
    geom, jaco, mass = so.geometric_factors(gx, gy, gz, N)
    rmult, mask = so.multiplicity_inverse(g2l, gs), so.dirichlet_mask(gx, gy, gz)
    b = 3 * np.pi ** 2 * np.sin(np.pi * gx) * np.sin(np.pi * gy) * np.sin(np.pi * gz) * mass * jaco
    b = so.gather_scatter(b.reshape(-1), g2l, gs) * mask
    take = lambda v: np.ascontiguousarray(v.reshape(E, Np)[m.elem_global].reshape(-1))
    D = np.ascontiguousarray(so.reference_derivative_matrix(N))
    axo = lambda v: so.elliptic_ax_batched(geom, v, N, D)
    dinv = mask / so.gather_scatter(so.jacobi_diagonal(geom, N, D), g2l, gs)
    xo, ito = so.elliptic_pcg_jacobi(axo, dinv, mask, g2l, gs, rmult, b, tol=1e-10, maxit=3000)
    xg, itg = elliptic_pcg(m, take(b), tol=1e-10, maxit=3000)
    if itg != ito:  # every rank sees the same counts: compare the iterates after the same number of steps
        xg, _ = elliptic_pcg(m, take(b), tol=0.0, maxit=ito)
    err = float(np.abs(xg - take(xo)).max() / np.abs(xo).max())
    errs = comm.allgather(err) if world > 1 else [err]
    return {"mesh": "deformed %dx%dx%d, N=7" % dims, "iters": int(itg), "oracle_iters": int(ito), "rel_err_max": max(errs),
            "ok": bool(abs(itg - ito) <= 1 and max(errs) < 1e-10)}


def run_ours(args):
    import ctypes as C

    import torch

    from sempy_b200 import _capi, distributed
    from sempy_b200.elliptic import _ax, elliptic_pcg

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
    sampler = ClockSampler(local_rank)  # started early: nvidia-smi needs about a second before its first sample
    sampler.start()
    t_setup = time.perf_counter()
    mesh, b_dev, exact = build_problem(nel, args.perturb, comm, args.strong)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    dof = mesh.get_num_elems() * mesh.Np
    ctx = mesh._context()
    lib = ctx.lib
    p2p_on = world > 1 and os.environ.get("SEMB_P2P", "1") != "0"

    def solve_dev():
        return elliptic_pcg(mesh, b_dev, tol=TOL, maxit=MAXIT)

    # ---- device-resident arm -------------------------------------------------
    for _ in range(args.warmup):
        x_dev, niter = solve_dev()
    sync_all()
    # Inside the timed region only the dominant kernel (the fused operator) is bracketed by events, and only in
    # every fourth step (spread over the region, so that the sample sees the same clocks as the rest): a bracket
    # costs two event records and takes the kernel out of the programmatic-dependent-launch chain on both sides.
    # Switching the brackets off reads the events back at a step boundary, where the stream is idle anyway.
    # The other kernels are timed in a second, untimed pass.
    k1_only = 1 | ((1 << _capi.K_AX) << 1)
    _capi.check(lib.semb_profile_enable(ctx.handle, k1_only))
    _capi.check(lib.semb_profile_reset(ctx.handle))
    _capi.check(lib.semb_profile_enable(ctx.handle, 0))
    pick = 1 if args.steps > 1 else 0
    launches0 = lib.semb_launch_count(ctx.handle)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    wall0 = time.time()
    ev0.record()
    iters_total = 0
    iters_bracketed = 0
    bracketed_steps = 0
    for step in range(args.steps):
        bracket = step % 4 == pick
        if bracket:
            _capi.check(lib.semb_profile_enable(ctx.handle, k1_only))
        x_dev, niter = solve_dev()
        iters_total += niter
        if bracket:
            _capi.check(lib.semb_profile_enable(ctx.handle, 0))
            iters_bracketed += niter
            bracketed_steps += 1
    ev1.record()
    sync_all()
    clocks = sampler.stop(wall0, time.time())
    ms_total = over_ranks(ev0.elapsed_time(ev1), "MAX")
    dof_all = over_ranks(dof, "SUM")
    launches = lib.semb_launch_count(ctx.handle) - launches0

    def read_profile(names):
        out = {}
        for name, kid in names:
            tot, cnt = C.c_double(0), C.c_int64(0)
            _capi.check(lib.semb_profile_get(ctx.handle, kid, C.byref(tot), C.byref(cnt)))
            out[name] = (tot.value, cnt.value)
        return out

    kern = read_profile([("ax", _capi.K_AX)])
    # second pass (not part of `value`): every kernel kind bracketed
    _capi.check(lib.semb_profile_enable(ctx.handle, 1))
    _capi.check(lib.semb_profile_reset(ctx.handle))
    ev0.record()
    _, iters_pass2 = solve_dev()
    ev1.record()
    sync_all()
    ms_pass2 = ev0.elapsed_time(ev1)
    kern2 = read_profile([("ax", _capi.K_AX), ("gs", _capi.K_GS), ("update", _capi.K_UPDATE)])
    _capi.check(lib.semb_profile_enable(ctx.handle, 0))
    value = dof_all * iters_total / (ms_total * 1e-3) / 1e9
    # per CG iteration: with a halo the operator of one iteration is two launches (interface elements, then the rest)
    ax_ms = kern["ax"][0] / max(iters_bracketed, 1)
    ax_gbs = BYTES_AX_DIR * dof / (ax_ms * 1e-3) / 1e9
    fs = f_shared(N_ORDER)
    # the in-place pass only covers the segments with more than two copies; the pairs are summed by the update kernel
    np1 = N_ORDER + 1
    f_long = (12 * (N_ORDER - 1) + 8) / np1 ** 3
    bytes_gs_long = 20.0 * f_long

    def kernel_entry(name, bytes_per_dof, what):
        tot, cnt = kern2[name]
        ms = tot / max(iters_pass2, 1)
        gbs = bytes_per_dof * dof / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"what": what, "ms_per_iteration": ms, "launches": cnt, "bytes_per_dof": bytes_per_dof, "achieved_gbs": gbs,
                "frac": gbs / peak, "share_of_step": tot / ms_pass2}

    # ---- fused Ax + mask + gather-scatter alone (the north-star 70 % target) ----
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
    axgs_gbs = (BYTES_AX + 20.0 * fs) * dof / (axgs_ms * 1e-3) / 1e9
    del w, p_dev

    # ---- end-to-end arm: host arrays through the reference-facing call -------------
    # (a) the drop-in call exactly as the reference's callers make it: a pageable NumPy b, a fresh NumPy x
    b_host = b_dev.cpu().numpy()
    for _ in range(2):  # two results alive at a time: both pinned blocks of the result pool exist after this
        x_host, _ = elliptic_pcg(mesh, b_host, tol=TOL, maxit=MAXIT)
    sync_all()
    t0 = time.perf_counter()
    e_iters = 0
    for _ in range(args.steps):
        x_host, it = elliptic_pcg(mesh, b_host, tol=TOL, maxit=MAXIT)
        e_iters += it
    sync_all()
    e_s = over_ranks(time.perf_counter() - t0, "MAX")
    e2e_val = dof_all * e_iters / e_s / 1e9
    same_answer = bool(np.allclose(x_host, x_dev.cpu().numpy(), rtol=0, atol=1e-12))
    # (b) the same call with caller-provided pinned buffers (out=, not in the reference's signature)
    b_pin = torch.empty(dof, dtype=torch.float64).pin_memory()
    b_pin.copy_(torch.from_numpy(b_host))
    x_pin = torch.empty(dof, dtype=torch.float64).pin_memory()
    b_pin_np, x_pin_np = b_pin.numpy(), x_pin.numpy()
    for _ in range(1):
        elliptic_pcg(mesh, b_pin_np, tol=TOL, maxit=MAXIT, out=x_pin_np)
    sync_all()
    t0 = time.perf_counter()
    p_iters = 0
    for _ in range(args.steps):
        _, it = elliptic_pcg(mesh, b_pin_np, tol=TOL, maxit=MAXIT, out=x_pin_np)
        p_iters += it
    sync_all()
    p_s = over_ranks(time.perf_counter() - t0, "MAX")
    e2e_pinned = dof_all * p_iters / p_s / 1e9
    del b_pin, x_pin, b_pin_np, x_pin_np, x_host, b_host

    # ---- correctness of what was just timed: facts that do not depend on the rank count -------------
    def maxabs(t):
        return over_ranks(t.abs().max().item(), "MAX")

    rm = torch.empty(dof, dtype=torch.float64, device="cuda")
    _capi.check(lib.semb_get_rmult(ctx.handle, rm.data_ptr(), _capi.DEVICE))
    err_exact = maxabs(x_dev - exact)
    xc = x_dev.clone()
    mesh.dssum(xc)
    err_cont = maxabs(xc * rm - x_dev) / max(maxabs(x_dev), 1e-300)
    xm = x_dev.clone()
    mesh.apply_mask(xm)
    boundary_zero = over_ranks(float(torch.equal(xm, x_dev)), "MIN") == 1.0
    res = b_dev - _ax(mesh, x_dev, _capi.AX_MASK | _capi.AX_GS)
    rnorm = np.sqrt(over_ranks(torch.dot(rm * res, res).item(), "SUM"))
    bnorm = np.sqrt(over_ranks(torch.dot(rm * b_dev, b_dev).item(), "SUM"))
    del xc, xm, res
    check = {"manufactured_solution_max_err": err_exact, "continuity_rel": err_cont, "boundary_zero": boundary_zero,
             "true_residual_rel": rnorm / bnorm, "host_and_device_arms_agree": same_answer,
             "iterations": int(niter)}
    check["ok"] = bool(err_exact < 1e-5 and err_cont < 1e-12 and boundary_zero and rnorm / bnorm < 1e-5 and same_answer)
    if not args.no_oracle_check:
        check["n_rank_solve_vs_oracle"] = small_vs_oracle(comm if world > 1 else distributed.SerialComm(), world, rank)
        check["ok"] = bool(check["ok"] and check["n_rank_solve_vs_oracle"]["ok"])

    if world > 1:
        all_clocks = comm.allgather(clocks)
        reasons = sorted({r for c in all_clocks for r in c.get("reasons", [])})
        sm = [c["sm_mhz"] for c in all_clocks if c.get("sm_mhz")]
        clocks = {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": reasons,
                  "per_rank_sm_mhz": sm}
    unique_local = int(mesh.num_unique)
    if rank != 0:
        dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle on a bounded sample of the same workload (N = 1 only) ---
    cpu_baseline = cpu_ref = None
    if world == 1 and not args.no_cpu:
        threads = use_all_host_threads()
        t_cpu0 = time.perf_counter()
        nx, ny, nz = mesh_dims(nel, 1, args.strong)
        if dof > 20e6:
            nz = max(nz // 8, 1)
        prob = oracle_problem(nx, ny, nz, args.perturb, N_ORDER)
        cpu_val, _, threads = time_oracle(prob, CPU_ITERS, 2, 1)
        cpu_secs = time.perf_counter() - t_cpu0
        cpu_baseline = {"value": cpu_val, "unit": "GDOF/s", "cores": threads, "kind": "port",
                        "sample": "C/OpenMP restatement of the reference algorithm (oracle/c/sem_oracle.c), %d threads, "
                                  "Jacobi-PCG on a %dx%dx%d-element N=%d box set up by the oracle's NumPy code, %d iterations x 2 "
                                  "steps (%.0f s with set-up)" % (threads, nx, ny, nz, N_ORDER, CPU_ITERS, cpu_secs)}
        del prob
        try:
            cpu_ref = time_numpy_reference()
        except Exception as exc:
            cpu_ref = {"unavailable": repr(exc)}

    if world == 1:
        parallelism = "1 GPU"
    elif p2p_on:
        parallelism = (f"{world} GPUs, one z-slab of elements per GPU (no data-path NCCL call): interface sums written straight "
                       "into the neighbours' memory over NVLink (peer memory, LL words) while the interior elements are "
                       "computed; the two dot products per iteration are summed over the ranks inside the kernels that "
                       "produce them (peer-memory all-reduce)")
    else:
        parallelism = (f"{world} GPUs, one z-slab of elements per GPU, NCCL send/recv halo exchange on a second stream "
                       "overlapped with the interior elements + ncclAllReduce of the dot products")
    k1 = "ax8_kernel" if N_ORDER == 7 else "axl_kernel" if N_ORDER >= 8 else "ax_generic_kernel"
    traffic = ncu_traffic(k1 + "_dir") if (world == 1 and nel == 32 and N_ORDER == 7) else None
    line = {
        "metric": f"FP64 Jacobi-PCG Poisson throughput, N={N_ORDER} (local DOF x iterations / s)",
        "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong > 0 else "weak",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world),
                   "elements_per_gpu": mesh.get_num_elems(), "local_dof_per_gpu": dof, "local_dof_total": int(dof_all),
                   "unique_dof_rank0": unique_local,
                   "iterations_per_step": iters_total / args.steps,
                   "cache": "inputs larger than L2 (G alone is %.0f MB per GPU)" % (48.0 * dof / 1e6),
                   "parallelism": parallelism, "setup_s": t_setup, "setup_breakdown_s": getattr(mesh, "setup_breakdown", None)},
        "cg_iters_per_s": iters_total / (ms_total * 1e-3),
        "us_per_iteration": ms_total / max(iters_total, 1) * 1e3,
        "ax_gs": {"gdofs": axgs_gdofs, "ms": axgs_ms, "bytes_per_dof": BYTES_AX + 20.0 * fs, "achieved_gbs": axgs_gbs,
                  "frac_of_hbm_roofline": axgs_gbs / peak,
                  "what": "semb_ax(MASK | GS): fused Ax + mask kernel followed by the in-place gather-scatter, device resident"},
        "roofline": {"bound": "hbm", "kernel": k1 + "<MASK, DOT, DIR=2> (direction update + Ax + mask + local pAp in one kernel)",
                     "achieved": ax_gbs, "peak": peak, "peak_source": peak_kind, "unit": "GB/s", "frac": ax_gbs / peak,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                     "algorithmic_bytes_per_launch": BYTES_AX_DIR * dof, "bytes_per_dof": BYTES_AX_DIR,
                     "bytes_per_dof_is": "p 8 + r 8 + dinv 8 + six G 48 + p written 8 + Ap written 8",
                     "avg_ms": ax_ms,
                     "avg_ms_is": "per CG iteration, CUDA events around every launch of this kernel in "
                                  f"{bracketed_steps} of the {args.steps} timed steps (every fourth)",
                     "launches_timed": kern["ax"][1],
                     "share_of_step": ax_ms * iters_total / ms_total, "gdofs": dof / (ax_ms * 1e-3) / 1e9},
        "kernels": {"ax": kernel_entry("ax", BYTES_AX_DIR, "direction update + Ax + mask + pAp"),
                    "gs": kernel_entry("gs", bytes_gs_long, "in-place gather-scatter of the segments with more than two copies"),
                    "update": kernel_entry("update", BYTES_UPDATE + 4.0, "r update (x every second iteration) + gather-scatter through 4-byte links + r.z + scalar step")},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * int(dof_all), "d2h_bytes_per_step": 8 * int(dof_all),
                "ms_per_step": e_s / args.steps * 1e3,
                "api": "x, niter = sempy_b200.elliptic.elliptic_pcg(mesh, b) with a pageable NumPy b and a fresh NumPy x "
                       "(the reference's calling convention)",
                "with_pinned_buffers": {"value": e2e_pinned, "ms_per_step": p_s / args.steps * 1e3,
                                        "api": "elliptic_pcg(mesh, b_pinned, out=x_pinned)"}},
        "gpu_launches": int(launches),
        "check": check,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if cpu_ref is not None:
        line["cpu_baseline_reference"] = cpu_ref
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nel", type=int, default=32, help="elements per side per GPU (default: BASELINE config 2; config 5: 48)")
    ap.add_argument("--perturb", type=float, default=0.0, help="vertex displacement in units of h (0.3 = the deformed mesh of config 4)")
    ap.add_argument("--strong", type=int, default=0, help="strong scaling: one box of STRONG^3 elements split over all GPUs (config 4: 64)")
    ap.add_argument("--order", type=int, default=7, help="polynomial order N (BASELINE config 3: --order 15 --nel 32)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the small N-rank solve against the CPU oracle")
    args = ap.parse_args()
    global N_ORDER
    N_ORDER = args.order
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
