"""ctypes binding of libsemb200.so (include/semb200.h).

The CUDA library is the product: if it is missing or cannot be loaded this
module raises -- there is no CPU fallback anywhere in the package.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

HOST, DEVICE = 0, 1
G9, G6 = 9, 6
AX_MASK, AX_GS = 1, 2
PRECOND_NONE, PRECOND_JACOBI, PRECOND_FDM = 0, 1, 2
K_AX, K_GS, K_UPDATE, K_DIRECTION = 0, 1, 2, 3

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsemb200.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)

# name -> (restype, argtypes); every symbol include/semb200.h declares
SIGNATURES = {
    "semb_version": (C.c_int, []),
    "semb_last_error": (C.c_char_p, []),
    "semb_device_count": (C.c_int, []),
    "semb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int]),
    "semb_destroy": (C.c_int, [C.c_void_p]),
    "semb_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "semb_synchronize": (C.c_int, [C.c_void_p]),
    "semb_set_derivative": (C.c_int, [C.c_void_p, C.c_void_p]),
    "semb_set_geom": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "semb_set_gather_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "semb_gs_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64]),
    "semb_gs_setup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "semb_set_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "semb_setup_mask_box": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]),
    "semb_get_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "semb_minmax": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]),
    "semb_setup_geom": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "semb_get_geom": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "semb_set_fdm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "semb_get_rmult": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "semb_ax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "semb_gs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "semb_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "semb_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "semb_diag": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "semb_gradient": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int]),
    "semb_gradient_transpose": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int]),
    "semb_kron": (
        C.c_int,
        [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
         C.c_int64, C.c_void_p, C.c_void_p, C.c_int],
    ),
    "semb_numbering": (
        C.c_int,
        [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
         C.POINTER(C.c_int64), C.c_int],
    ),
    "semb_geom_factors": (
        C.c_int,
        [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int],
    ),
    "semb_cg": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p],
    ),
    "semb_cg_generic": (
        C.c_int,
        [C.c_int, C.c_int64, APPLY_FN, C.c_void_p, APPLY_FN, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
         C.c_int, C.POINTER(C.c_int), C.c_void_p],
    ),
    "semb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "semb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "semb_set_halo": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_int64],
    ),
    "semb_p2p_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.c_char_p]),
    "semb_p2p_attach": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p]),
    "semb_malloc": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int64]),
    "semb_free": (C.c_int, [C.c_int, C.c_void_p]),
    "semb_memcpy": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "semb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "semb_host_free": (C.c_int, [C.c_void_p]),
    "semb_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "semb_profile_reset": (C.c_int, [C.c_void_p]),
    "semb_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "semb_launch_count": (C.c_int64, [C.c_void_p]),
    "semb_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
}


def lib_path():
    return _LIB_PATH


def lib():
    """Load the library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or sempy_b200/csrc/build.sh (needs nvcc). sempy_b200 has no CPU fallback."
            )
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().semb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"semb200 error {rc}: {msg}")


def require_gpu():
    n = lib().semb_device_count()
    if n <= 0:
        raise RuntimeError("sempy_b200 needs a CUDA device (B200); none is visible and there is no CPU fallback")
    return n


def is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


def as_f64(x):
    """C-contiguous float64 view/copy of a numpy input."""
    return np.ascontiguousarray(x, dtype=np.float64)


def ptr(a):
    """Raw address of a numpy array or torch tensor."""
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def current_device():
    dev = os.environ.get("SEMB_DEVICE")
    if dev is not None:
        return int(dev)
    return int(os.environ.get("LOCAL_RANK", "0"))


class DevBuf:
    """Device array owned through semb_malloc / semb_free: the mesh keeps its node arrays (coordinates,
    numbering, ...) here so that set-up never round-trips through the host.  ``__cuda_array_interface__``
    lets torch / cupy wrap it without a copy (``torch.as_tensor(buf, device="cuda")``)."""

    def __init__(self, shape, dtype, device=None):
        self.shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.device = current_device() if device is None else device
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        handle = C.c_void_p()
        check(lib().semb_malloc(self.device, C.byref(handle), max(self.nbytes, 1)))
        self.ptr = handle.value

    @classmethod
    def from_host(cls, arr, dtype=None, device=None):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        buf = cls(arr.shape, arr.dtype, device)
        if buf.nbytes:
            check(lib().semb_memcpy(buf.device, buf.ptr, arr.ctypes.data, buf.nbytes, 0))
        return buf

    def to_host(self, count=None):
        out = np.empty(self.shape, dtype=self.dtype)
        nbytes = self.nbytes if count is None else int(count) * self.dtype.itemsize
        if nbytes:
            check(lib().semb_memcpy(self.device, out.ctypes.data, self.ptr, nbytes, 1))
        return out if count is None else out.reshape(-1)[: int(count)]

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 2}

    def __del__(self):
        try:
            if self.ptr:
                lib().semb_free(self.device, self.ptr)
                self.ptr = None
        except Exception:
            pass


class _PinnedPool:
    """Result vectors in pinned host memory.  ``empty(n)`` returns a fresh float64 ndarray whose memory
    comes from cudaMallocHost (the device -> host copy of a result is then one direct DMA, ~2.5x faster
    than into pageable memory); when the array (and every view of it) has been garbage collected the
    block goes back to the pool, so a solve loop allocates once.  At most ``keep`` idle blocks are kept."""

    def __init__(self, keep=4):
        self.keep = keep
        self.idle = {}  # nbytes -> [address]

    def empty(self, n):
        import weakref

        nbytes = max(int(n) * 8, 8)
        free = self.idle.get(nbytes)
        if free:
            addr = free.pop()
        else:
            handle = C.c_void_p()
            check(lib().semb_host_alloc(C.byref(handle), nbytes))
            addr = handle.value
        buf = (C.c_double * int(n)).from_address(addr)
        weakref.finalize(buf, self._release, addr, nbytes)
        return np.frombuffer(buf, dtype=np.float64, count=int(n))

    def _release(self, addr, nbytes):
        try:
            free = self.idle.setdefault(nbytes, [])
            if sum(len(v) for v in self.idle.values()) < self.keep:
                free.append(addr)
            else:
                lib().semb_host_free(addr)
        except Exception:
            pass


_pinned = _PinnedPool()


def pinned_empty(n):
    return _pinned.empty(n)
