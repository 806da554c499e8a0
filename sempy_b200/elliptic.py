"""Matrix-free Poisson operator and the mesh-aware CG on the GPU.

Drop-in for sempy/elliptic.py: ``elliptic_ax(mesh, p)`` (lines 9-27) and
``elliptic_cg(mesh, b, tol=1e-12, maxit=100, verbose=0)`` (lines 30-73) keep
their signatures, return values and verbose print format.  The whole CG loop
runs in libsemb200 (``semb_cg``): fused Ax + mask + local pAp, atomic-free
gather-scatter, fused vector updates with device-resident scalars.

``p`` / ``b`` may be NumPy arrays (result is a new NumPy array, one host->device
and one device->host copy per call) or torch CUDA tensors (result is a new CUDA
tensor, nothing crosses PCIe).  ``elliptic_pcg`` adds the Jacobi-preconditioned
variant of BASELINE configs 2-5; ``elliptic_cg_loopy`` (the reference's OpenCL
prototype, elliptic.py:76-172) is routed to the same CUDA path.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from sempy_b200 import _capi


def _new_like(mesh, x, name):
    size = mesh.get_num_elems() * mesh.Np
    if _capi.is_torch_cuda(x):
        import torch

        if x.dtype != torch.float64 or x.numel() != size:
            raise ValueError(f"{name}: need a float64 CUDA tensor of {size} entries")
        xin = x.contiguous().reshape(-1)
        return xin, torch.empty_like(xin), _capi.DEVICE
    xin = _capi.as_f64(np.asarray(x)).reshape(-1)
    if xin.size != size:
        raise ValueError(f"{name}: expected {size} entries, got {xin.size}")
    # the result lives in pinned host memory (direct DMA from the device); to the caller it is a fresh ndarray
    return xin, _capi.pinned_empty(xin.size), _capi.HOST


def _ax(mesh, p, flags):
    need = ["D", "geom"] + (["mask"] if flags & _capi.AX_MASK else []) + (["gs"] if flags & _capi.AX_GS else [])
    ctx = mesh._context(need=tuple(need))
    pin, out, mem = _new_like(mesh, p, "elliptic_ax")
    mesh._bind_stream(ctx, mem)
    _capi.check(ctx.lib.semb_ax(ctx.handle, _capi.ptr(pin), _capi.ptr(out), flags, mem))
    return out


def elliptic_ax(mesh, p):
    """Unassembled, unmasked A p on local vectors (sempy/elliptic.py:9-27)."""
    return _ax(mesh, p, 0)


def _cg(mesh, b, tol, maxit, verbose, precond, out=None):
    ctx = mesh._context(need=("D", "geom", "gs"))  # the mask is optional: applied when setup_mask() has run
    if precond == _capi.PRECOND_FDM:
        from sempy_b200.fdm import setup_fdm

        setup_fdm(mesh)
    bin_, x, mem = _new_like(mesh, b, "elliptic_cg")
    if out is not None:
        # caller-provided result buffer (e.g. pinned host memory); same kind as b
        _p, omem, _keep = mesh._vector_args(out, "out")
        if omem != mem:
            raise ValueError("out must live where b lives (both NumPy or both CUDA tensors)")
        x = out
    mesh._bind_stream(ctx, mem)
    niter = C.c_int(0)
    hist = np.zeros((max(int(maxit), 1), 5), dtype=np.float64) if verbose else None
    _capi.check(
        ctx.lib.semb_cg(ctx.handle, _capi.ptr(bin_), _capi.ptr(x), float(tol), int(maxit), precond, mem,
                        C.byref(niter), hist.ctypes.data if verbose else None)
    )
    if verbose:
        if niter.value > 0:
            print("Initial rnorm={}".format(hist[0, 0]))
        else:
            w = np.empty(1)
            _capi.check(ctx.lib.semb_dot(ctx.handle, _capi.ptr(bin_), _capi.ptr(bin_), 1, w.ctypes.data, mem))
            print("Initial rnorm={}".format(w[0]))
        for it in range(niter.value):
            r0, r1, alpha, beta, pap = hist[it]
            print("niter={} r0={} r1={} alpha={} beta={} pap={}".format(it, r0, r1, alpha, beta, pap))
    return x, niter.value


def elliptic_cg(mesh, b, tol=1e-12, maxit=100, verbose=0, out=None):
    """CG on local vectors (sempy/elliptic.py:30-73); returns (x, niter).
    ``out`` (optional, not in the reference) receives x instead of a new array."""
    return _cg(mesh, b, tol, maxit, verbose, _capi.PRECOND_NONE, out)


def elliptic_pcg(mesh, b, tol=1e-8, maxit=100, verbose=0, out=None):
    """Jacobi-preconditioned CG on local vectors: sempy/iterative.py:45-85 with
    Minv = mask / dssum(diag A), dots weighted by rmult as in elliptic_cg."""
    return _cg(mesh, b, tol, maxit, verbose, _capi.PRECOND_JACOBI, out)


def elliptic_pcg_fdm(mesh, b, tol=1e-8, maxit=100, verbose=0, out=None):
    """PCG with the fast-diagonalisation preconditioner (sempy_b200/fdm.py; the reference's sketch is
    example.py:71-114,155-171,186-190): exact solves on the element interiors, Jacobi on the element surfaces."""
    return _cg(mesh, b, tol, maxit, verbose, _capi.PRECOND_FDM, out)


def elliptic_cg_loopy(mesh, b, tol=1e-12, maxit=100, verbose=0):
    """Signature of the reference's OpenCL/loopy variant (elliptic.py:76-172);
    the device path here is the CUDA one."""
    return elliptic_cg(mesh, b, tol=tol, maxit=maxit, verbose=verbose)
