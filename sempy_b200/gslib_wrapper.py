"""``gslib_wrapper``-compatible gather-scatter handle backed by the CUDA kernels of libsemb200.

The reference drives gslib (Nek5000's gather-scatter library) through three names,
``from gslib_wrapper import GS, gs_add, gs_double`` (sempy/mesh.py:6):

    self.gs = GS(0)                                   # mesh.py:437
    self.gs.setup(self.global_id.reshape((size,)))    # mesh.py:438, int32 ids, 1-based
    self.gs.gs(x, gs_double, gs_add)                  # mesh.py:441,445, in place

This module provides exactly those, so the UNMODIFIED ``sempy/mesh.py`` runs on the GPU gather-scatter
with ``sys.modules["gslib_wrapper"] = sempy_b200.gslib_wrapper`` (or by putting this file on the path
under that name).  ``x`` may be a NumPy float64 array (copied to the device and back) or a torch CUDA
tensor (updated in place on the device).  Ids are grouped on the device (semb_gs_setup); entries with id
0 take no part, as in gslib.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from sempy_b200 import _capi

# the two constants SemPy passes through (gslib has more types / operations; the reference uses these)
gs_double = "gs_double"
gs_add = "gs_add"


class GS:
    def __init__(self, comm=0):
        # the reference passes comm = 0 (single process, sempy/mesh.py:437); multi-GPU meshes go through
        # sempy_b200.mesh.Mesh(comm=...), which knows the element partition
        if comm not in (0, None):
            raise ValueError("GS: only the single-process handle GS(0) of the reference is provided here")
        self._lib = _capi.lib()
        self._handle = None
        self._n = 0

    def setup(self, ids):
        _capi.require_gpu()
        ids = np.ascontiguousarray(np.asarray(ids).reshape(-1), dtype=np.int32)
        self._free()
        handle = C.c_void_p()
        _capi.check(self._lib.semb_gs_create(C.byref(handle), _capi.current_device(), ids.size))
        self._handle, self._n = handle, ids.size
        _capi.check(self._lib.semb_gs_setup(handle, ids.ctypes.data, _capi.HOST))

    def gs(self, x, dtype=gs_double, op=gs_add):
        if self._handle is None:
            raise RuntimeError("GS.gs: call setup(ids) first")
        if dtype != gs_double or op != gs_add:
            raise NotImplementedError("GS.gs: the reference only uses gs_double with gs_add")
        if _capi.is_torch_cuda(x):
            import torch

            if x.dtype != torch.float64 or not x.is_contiguous() or x.numel() != self._n:
                raise ValueError(f"GS.gs: need a contiguous float64 CUDA tensor of {self._n} entries")
            stream = torch.cuda.current_stream().cuda_stream
            _capi.check(self._lib.semb_set_stream(self._handle, stream if stream else 1))
            _capi.check(self._lib.semb_gs(self._handle, x.data_ptr(), _capi.DEVICE))
            return
        if not isinstance(x, np.ndarray) or x.dtype != np.float64 or not x.flags.c_contiguous or x.size != self._n:
            raise ValueError(f"GS.gs: need a C-contiguous float64 NumPy array of {self._n} entries")
        _capi.check(self._lib.semb_set_stream(self._handle, None))
        _capi.check(self._lib.semb_gs(self._handle, x.ctypes.data, _capi.HOST))

    def _free(self):
        if self._handle is not None:
            self._lib.semb_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass
