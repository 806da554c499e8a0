"""Hex mesh with GLL nodes: set-up on the host, hot-path state on the GPU.

Drop-in for the ``Mesh`` of sempy/mesh.py (same constructor, method names and
attributes): ``find_physical_coordinates``, ``calc_geometric_factors``,
``establish_global_numbering``, ``setup_mask``, ``dssum``, ``apply_mask`` and
the getters.  Differences that matter:

* no ``meshio`` / ``gslib_wrapper``: a gmsh-2.2 ASCII reader is built in and the
  gather-scatter is the atomic-free CUDA kernel of libsemb200 (``semb_gs``);
* set-up is vectorised (no per-node Python loops) and the tensor contractions
  in it run on the GPU (``semb_kron``, ``semb_gradient``);
* ``dssum`` / ``apply_mask`` / ``elliptic_ax`` / ``elliptic_cg`` use a device
  context (``semb_ctx``) that holds D, the six geometric factors, the
  gather-scatter schedule and the mask bits; it is built on first use.

Vectors handed to ``dssum``/``apply_mask`` may be NumPy arrays (updated in
place through a host<->device copy, as the reference's callers expect) or
torch CUDA tensors (updated in place on the device, no copies).
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from sempy_b200 import _capi, distributed
from sempy_b200.derivative import reference_derivative_matrix
from sempy_b200.gradient import gradient
from sempy_b200.interpolation import lagrange
from sempy_b200.kron import kron
from sempy_b200.mass import reference_mass_matrix_3d
from sempy_b200.quadrature import gauss_lobatto

# gmsh hexahedron vertex order -> lexicographic (i fastest); sempy/mesh.py:272-274
_GMSH_TO_LEX = [0, 1, 3, 2, 4, 5, 7, 6]


def read_gmsh22(fname):
    """gmsh 2.2 ASCII: vertices (nv,3) and 8-node hexahedra (E,8), 0-based.
    Takes the place of meshio.read in sempy/mesh.py:102-110."""
    with open(fname) as fh:
        lines = fh.read().splitlines()
    pos = 0
    points = None
    hexes = []
    nquads = 0
    while pos < len(lines):
        tag = lines[pos].strip()
        pos += 1
        if tag == "$MeshFormat":
            version = lines[pos].split()
            if not version[0].startswith("2") or version[1] != "0":
                raise Exception("Only gmsh 2.x ASCII meshes are supported")
        elif tag == "$Nodes":
            count = int(lines[pos])
            block = np.array([ln.split() for ln in lines[pos + 1: pos + 1 + count]], dtype=np.float64)
            points = np.zeros((count, 3))
            points[block[:, 0].astype(np.int64) - 1] = block[:, 1:4]
            pos += 1 + count
        elif tag == "$Elements":
            count = int(lines[pos])
            for ln in lines[pos + 1: pos + 1 + count]:
                f = ln.split()
                etype, ntags = int(f[1]), int(f[2])
                if etype == 5:
                    hexes.append([int(v) - 1 for v in f[3 + ntags: 3 + ntags + 8]])
                elif etype == 3:
                    nquads += 1
            pos += 1 + count
    if points is None:
        raise Exception(f"{fname}: no $Nodes section")
    return points, np.array(hexes, dtype=np.int64).reshape(-1, 8), nquads


def write_gmsh22(fname, points, hexes):
    """Write vertices and hexahedra as gmsh 2.2 ASCII (full double precision)."""
    with open(fname, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(points))
        for q, pt in enumerate(points):
            fh.write("%d %r %r %r\n" % (q + 1, float(pt[0]), float(pt[1]), float(pt[2])))
        fh.write("$EndNodes\n$Elements\n%d\n" % len(hexes))
        for q, h in enumerate(hexes):
            fh.write("%d 5 2 2 1 %s\n" % (q + 1, " ".join(str(int(v) + 1) for v in h)))
        fh.write("$EndElements\n")


def box_mesh_arrays(nx, ny=None, nz=None, perturb=0.0, seed=0):
    """Vertices/hexes of an nx*ny*nz box on [0,1]^3, elements i-fastest and
    k-slowest (z-slabs are contiguous element ranges); ``perturb`` displaces the
    interior vertices by perturb*h*(U(0,1)-0.5) per axis (full 3x3 G)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    gx, gy, gz = (np.linspace(0.0, 1.0, m + 1) for m in (nx, ny, nz))
    Z, Y, X = np.meshgrid(gz, gy, gx, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if perturb:
        rng = np.random.default_rng(seed)
        h = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
        shift = perturb * h * (rng.random(pts.shape) - 0.5)
        kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
        inner = ((ii > 0) & (ii < nx) & (jj > 0) & (jj < ny) & (kk > 0) & (kk < nz)).ravel()
        pts[inner] += shift[inner]
    k, j, i = (a.ravel() for a in np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij"))
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    base = k * sy + j * sx + i
    hexes = np.stack([base, base + 1, base + 1 + sx, base + sx,
                      base + sy, base + 1 + sy, base + 1 + sx + sy, base + sx + sy], axis=1).astype(np.int64)
    return pts, hexes


def global_numbering_from_coordinates(xe, ye, ze, tol=1e-12):
    """The reference's numbering rule (sempy/mesh.py:389-435), vectorised:
    stable lexicographic sort of all local nodes by exact (x, y, z), new id
    whenever the squared distance to the next sorted point exceeds ``tol``,
    ids from 1.  Returns (global_id int32 (E,Np), global_to_local int64,
    global_start int64) -- bit-identical to the reference's Python-object sort
    for the same coordinates."""
    E, Np = xe.shape
    size = E * Np
    x, y, z = xe.reshape(-1), ye.reshape(-1), ze.reshape(-1)
    order = np.lexsort((z, y, x)).astype(np.int64)  # last key is primary; lexsort is stable
    sx, sy, sz = x[order], y[order], z[order]
    gap = (sx[:-1] - sx[1:]) ** 2 + (sy[:-1] - sy[1:]) ** 2 + (sz[:-1] - sz[1:]) ** 2 > tol
    ids = np.empty(size, dtype=np.int32)
    ids[order] = np.concatenate(([1], 1 + np.cumsum(gap, dtype=np.int64))).astype(np.int32)
    start = np.concatenate(([0], np.flatnonzero(gap) + 1, [size])).astype(np.int64)
    return ids.reshape(E, Np), order, start


def global_numbering_device(xe, ye, ze, tol=1e-12):
    """Same rule, same outputs, computed on the GPU (semb_numbering: three stable radix-sort passes,
    gap flags, prefix sum).  Bit-identical to ``global_numbering_from_coordinates``."""
    lib = _capi.lib()
    _capi.require_gpu()
    E, Np = xe.shape
    n = E * Np
    x, y, z = (_capi.as_f64(c).reshape(-1) for c in (xe, ye, ze))
    gid = np.empty(n, dtype=np.int32)
    g2l = np.empty(n, dtype=np.int64)
    start = np.empty(n + 1, dtype=np.int64)
    ng = C.c_int64(0)
    _capi.check(lib.semb_numbering(_capi.current_device(), n, x.ctypes.data, y.ctypes.data, z.ctypes.data, float(tol),
                                   gid.ctypes.data, g2l.ctypes.data, start.ctypes.data, C.byref(ng), _capi.HOST))
    return gid.reshape(E, Np), g2l, start[: ng.value + 1].copy()


def device_setup_enabled():
    """Numbering and geometric factors run on the GPU (semb_numbering, semb_geom_factors) unless
    SEMB_DEVICE_SETUP=0 selects the vectorised host NumPy forms (same results: ids bit-identical,
    factors to rounding)."""
    return os.environ.get("SEMB_DEVICE_SETUP", "1") != "0"


def bounding_box_mask(xe, ye, ze, tol=1e-8, bbox=None):
    """0 where any coordinate is within ``tol`` of the global bounding box
    (homogeneous Dirichlet), else 1; sempy/mesh.py:449-477, flat (E*Np,).
    ``bbox`` = ((xmin, xmax), (ymin, ymax), (zmin, zmax)) overrides the local
    extrema (multi-GPU: the box of the whole mesh)."""
    mask = np.ones(xe.shape)
    for axis, c in enumerate((xe, ye, ze)):
        lo, hi = (np.min(c), np.max(c)) if bbox is None else bbox[axis]
        mask[np.abs(c - lo) < tol] = 0
        mask[np.abs(c - hi) < tol] = 0
    return mask.reshape(-1)


class _Context:
    """Owner of one semb_ctx; frees it with the mesh."""

    def __init__(self, device, nelem, Nq):
        self.lib = _capi.lib()
        _capi.require_gpu()
        self.device = device
        handle = C.c_void_p()
        _capi.check(self.lib.semb_create(C.byref(handle), device, nelem, Nq))
        self.handle = handle
        self.have = set()  # which parts of the operator state are set: "D", "geom", "gs", "mask"

    def __del__(self):
        try:
            if self.handle:
                self.lib.semb_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _NodeArray:
    """A mesh array that lives on the device, on the host, or on both (mirrored on first request).
    Set-up keeps everything on the device; the host copy only exists if somebody asks for it
    (``mesh.xe``, ``mesh.get_x()``, ``mesh.global_id`` ... as the reference's callers do)."""

    def __init__(self, host=None, dev=None, count=None, shape=None):
        self._host, self._dev, self._count, self._shape = host, dev, count, shape

    def host(self):
        if self._host is None:
            h = self._dev.to_host(self._count)
            self._host = h.reshape(self._shape) if self._shape is not None else h
        return self._host

    def dev(self):
        if self._dev is None:
            self._dev = _capi.DevBuf.from_host(self._host)
        return self._dev

    @property
    def ptr(self):
        return self.dev().ptr


class _TimedComm:
    """Communicator proxy that adds up the time spent in collectives (set-up breakdown of bench.py)."""

    def __init__(self, comm):
        self._comm, self.rank, self.size, self.seconds = comm, comm.rank, comm.size, 0.0

    def allgather(self, obj):
        import time

        t0 = time.perf_counter()
        out = self._comm.allgather(obj)
        self.seconds += time.perf_counter() - t0
        return out


class Mesh:
    def __init__(self, fname, debug=0, comm=None):
        if not isinstance(fname, str):
            raise Exception("Only strings are allowed for file name")
        self.debug = debug
        self.comm = comm if comm is not None else distributed.SerialComm()
        self._reset_device_state()
        self.read(fname)

    # -- construction -------------------------------------------------------
    @classmethod
    def from_arrays(cls, points, hexes, debug=0, comm=None):
        """Build from vertex coordinates (nv,3) and hexahedra (E,8) in gmsh order.
        With a multi-rank ``comm`` every rank passes the WHOLE (vertex-level) mesh and keeps
        its own contiguous slab of elements (see sempy_b200/distributed.py)."""
        self = cls.__new__(cls)
        self.debug = debug
        self.comm = comm if comm is not None else distributed.SerialComm()
        self._reset_device_state()
        self._set_topology(np.asarray(points, dtype=np.float64), np.asarray(hexes, dtype=np.int64))
        return self

    def _reset_device_state(self):
        self._ctx = None
        self._coords = [None, None, None]
        self._gid = self._g2l = self._gst = None
        self._geom6_host = self._mask_host = self._rmult_host = None
        self._jaco = None
        self.num_unique = None

    def read(self, fname):
        points, hexes, nquads = read_gmsh22(fname)
        if len(hexes) == 0:
            if nquads:
                raise Exception("2D not supported yet.")
            raise Exception(f"{fname}: no hexahedra")
        self._set_topology(points, hexes)

    def _set_topology(self, points, hexes):
        assert len(hexes) > 0 and len(points) > 0
        self.num_global_elements = hexes.shape[0]
        self._halo = None
        self.halo_timings = {}
        self._shared_vertex = None
        self._n_if_elems = 0
        if self.comm.size > 1:
            # this rank's slab, interface elements first so that their Ax can be launched ahead
            ranges = distributed.slab_ranges(hexes.shape[0], self.comm.size)
            lo, hi = ranges[self.comm.rank]
            if hi <= lo:
                raise Exception("more ranks than elements")
            self._shared_vertex = distributed.shared_vertex_flags(hexes, ranges, len(points))
            order, self._n_if_elems = distributed.interface_first_order(hexes[lo:hi], self._shared_vertex)
            self.elem_global = lo + order
            hexes = hexes[self.elem_global]
        else:
            self.elem_global = np.arange(hexes.shape[0])
        self.elem_to_vert_map = hexes
        self.num_elements = hexes.shape[0]
        self.num_verts = hexes.shape[1]
        self.ndim = points.shape[1]
        if self.ndim != 3:
            raise Exception("2D not supported yet.")
        self.num_faces = 2 * self.ndim
        self.nface_verts = self.ndim + 1
        self.face_to_vert_map = np.array(
            [[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [0, 1, 2, 3], [4, 5, 6, 7]]
        )
        self.x = points[hexes, 0]
        self.y = points[hexes, 1]
        self.z = points[hexes, 2]

    # -- simple getters (sempy/mesh.py:167-186) -------------------------------
    def get_num_elems(self):
        return self.num_elements

    def get_ndim(self):
        return self.ndim

    def get_num_faces(self):
        return self.num_faces

    def get_num_verts(self):
        return self.num_verts

    def get_face_to_vert_map(self):
        return self.face_to_vert_map

    def get_elem_to_vert_map(self):
        return self.elem_to_vert_map

    def get_num_face_verts(self):
        return self.nface_verts

    # -- node arrays: device resident, host copies on request ---------------------
    def _coord_get(self, axis):
        a = self._coords[axis]
        if a is None:
            raise AttributeError("call find_physical_coordinates() first")
        return a.host()

    def _coord_set(self, axis, value):
        value = _capi.as_f64(value)
        self._coords[axis] = _NodeArray(host=value.reshape(self.num_elements, -1))

    xe = property(lambda self: self._coord_get(0), lambda self, v: self._coord_set(0, v))
    ye = property(lambda self: self._coord_get(1), lambda self, v: self._coord_set(1, v))
    ze = property(lambda self: self._coord_get(2), lambda self, v: self._coord_set(2, v))

    def _numbering_get(self, which):
        a = getattr(self, which)
        if a is None:
            raise AttributeError("call establish_global_numbering() first")
        return a.host()

    def _numbering_set(self, which, value, dtype):
        setattr(self, which, None if value is None else _NodeArray(host=np.ascontiguousarray(value, dtype=dtype)))

    global_id = property(lambda self: self._numbering_get("_gid"), lambda self, v: self._numbering_set("_gid", v, np.int32))
    global_to_local = property(lambda self: self._numbering_get("_g2l"),
                               lambda self, v: self._numbering_set("_g2l", v, np.int64))
    global_start = property(lambda self: self._numbering_get("_gst"), lambda self, v: self._numbering_set("_gst", v, np.int64))

    @property
    def geom6(self):
        """(E,6,Np): the six distinct geometric factors rr, rs, rt, ss, st, tt (host copy of the device array)."""
        if self._geom6_host is None:
            ctx = self._context(need=("geom",))
            g6 = np.empty((self.num_elements, 6, self.Np))
            _capi.check(ctx.lib.semb_get_geom(ctx.handle, g6.ctypes.data, _capi.HOST))
            self._geom6_host = g6
        return self._geom6_host

    @geom6.setter
    def geom6(self, value):
        g6 = _capi.as_f64(value).reshape(self.num_elements, 6, self.Np)
        ctx = self._context(need=())
        _capi.check(ctx.lib.semb_set_geom(ctx.handle, g6.ctypes.data, _capi.G6, _capi.HOST))
        ctx.have.add("geom")
        ctx.fdm_ready = False
        self._geom6_host = g6

    @property
    def geom(self):
        """(E,3,3,Np) view of the geometric factors as the reference stores them."""
        return self.geom6[:, [[0, 1, 2], [1, 3, 4], [2, 4, 5]], :]

    @property
    def jaco(self):
        if self._jaco is None:
            raise AttributeError("call calc_geometric_factors() first")
        return self._jaco.host()

    @property
    def mass(self):
        return np.broadcast_to(self._B, (self.num_elements, self.Np))

    @property
    def mask(self):
        """0.0 at Dirichlet nodes, 1.0 elsewhere (sempy/mesh.py:459-475); None before setup_mask()."""
        if self._mask_host is None and self._ctx is not None and "mask" in self._ctx.have:
            ctx = self._ctx
            m = np.empty(self.num_elements * self.Np)
            _capi.check(ctx.lib.semb_get_mask(ctx.handle, m.ctypes.data, _capi.HOST))
            self._mask_host = m
        return self._mask_host

    @mask.setter
    def mask(self, value):
        self._mask_host = None if value is None else _capi.as_f64(value).reshape(-1)
        if value is not None and self._ctx is not None:
            ids = np.ascontiguousarray(np.flatnonzero(self._mask_host == 0), dtype=np.int64)
            _capi.check(self._ctx.lib.semb_set_mask(self._ctx.handle, ids.ctypes.data, ids.size))
            self._ctx.have.add("mask")

    @property
    def rmult(self):
        if self._rmult_host is None:
            ctx = self._context(need=("gs",))
            r = np.empty(self.num_elements * self.Np)
            _capi.check(ctx.lib.semb_get_rmult(ctx.handle, r.ctypes.data, _capi.HOST))
            self._rmult_host = r
        return self._rmult_host

    def device_array(self, name):
        """Device-resident array as an object with ``__cuda_array_interface__`` (wrap it with
        ``torch.as_tensor(a, device="cuda")``; no copy): "x", "y", "z", "jaco", "global_id"."""
        src = {"x": self._coords[0], "y": self._coords[1], "z": self._coords[2], "jaco": self._jaco, "global_id": self._gid}[name]
        if src is None:
            raise RuntimeError(f"{name} is not set up yet")
        return src.dev()

    # -- set-up ---------------------------------------------------------------
    def find_physical_coordinates(self, N):
        """GLL coordinates by trilinear interpolation of the 8 vertices
        (sempy/mesh.py:247-302): one batched kron launch per coordinate, results stay on the device."""
        self.N = N
        self.Nq = N + 1
        self.Nfp = (N + 1) * (N + 1)
        self.Np = (N + 1) ** 3
        self._reset_device_state()
        z_1, _ = gauss_lobatto(1)
        z_N, _ = gauss_lobatto(N)
        J = _capi.as_f64(lagrange(z_N, z_1))
        ctx = self._context(need=())
        if not device_setup_enabled():
            for axis, v in enumerate((self.x, self.y, self.z)):
                self._coord_set(axis, kron(J, J, J, v[:, _GMSH_TO_LEX]))
            return
        E, Nq = self.num_elements, self.Nq
        for axis, v in enumerate((self.x, self.y, self.z)):
            verts = _capi.DevBuf.from_host(_capi.as_f64(v[:, _GMSH_TO_LEX]), device=ctx.device)
            out = _capi.DevBuf((E, self.Np), np.float64, device=ctx.device)
            _capi.check(ctx.lib.semb_kron(ctx.device, J.ctypes.data, Nq, 2, J.ctypes.data, Nq, 2, J.ctypes.data, Nq, 2, E,
                                          verts.ptr, out.ptr, _capi.DEVICE))
            self._coords[axis] = _NodeArray(dev=out)

    def calc_geometric_factors(self):
        """G = (grad a . grad b) B J, Jacobian and mass per node (sempy/mesh.py:304-356): one fused kernel
        (nine coordinate derivatives + the metric formulas per node) writing straight into the operator's
        device array; nothing of G crosses PCIe."""
        n = self.Nq
        D = _capi.as_f64(reference_derivative_matrix(self.N))
        self._B = reference_mass_matrix_3d(self.N)
        self._geom6_host = None
        E, Np = self.num_elements, self.Np
        if self._ctx is not None and device_setup_enabled():
            ctx = self._context(need=())
            _, w = gauss_lobatto(self.N)
            w = _capi.as_f64(w)
            jac = _capi.DevBuf((E, Np), np.float64, device=ctx.device)
            _capi.check(ctx.lib.semb_setup_geom(ctx.handle, w.ctypes.data, self._coords[0].ptr, self._coords[1].ptr,
                                                self._coords[2].ptr, jac.ptr, _capi.DEVICE))
            self._jaco = _NodeArray(dev=jac)
            ctx.have.add("geom")
            ctx.fdm_ready = False
            return
        xr, xs, xt = gradient(self.xe, n, D)
        yr, ys, yt = gradient(self.ye, n, D)
        zr, zs, zt = gradient(self.ze, n, D)
        J = xr * (ys * zt - yt * zs) - yr * (xs * zt - xt * zs) + zr * (xs * yt - ys * xt)
        rx = (ys * zt - yt * zs) / J
        sx = (yt * zr - yr * zt) / J
        tx = (yr * zs - ys * zr) / J
        ry = -(zt * xs - zs * xt) / J
        sy = -(zr * xt - zt * xr) / J
        ty = -(zs * xr - zr * xs) / J
        rz = (xs * yt - xt * ys) / J
        sz = -(xr * yt - xt * yr) / J
        tz = (xr * ys - xs * yr) / J
        del xr, xs, xt, yr, ys, yt, zr, zs, zt
        B = self._B
        g6 = np.empty((E, 6, Np))
        g6[:, 0] = (rx * rx + ry * ry + rz * rz) * B * J
        g6[:, 1] = (rx * sx + ry * sy + rz * sz) * B * J
        g6[:, 2] = (rx * tx + ry * ty + rz * tz) * B * J
        g6[:, 3] = (sx * sx + sy * sy + sz * sz) * B * J
        g6[:, 4] = (sx * tx + sy * ty + sz * tz) * B * J
        g6[:, 5] = (tx * tx + ty * ty + tz * tz) * B * J
        self._jaco = _NodeArray(host=J)
        self.geom6 = g6

    def establish_global_numbering(self):
        """Global ids by coordinate sort (sempy/mesh.py:389-442) and, taking the place of the reference's
        gslib handle, the device gather-scatter schedule -- numbering, schedule and multiplicities are all
        built on the device from the device-resident coordinates."""
        ctx = self._context(need=())
        lib = ctx.lib
        n = self.num_elements * self.Np
        self._rmult_host = None
        if device_setup_enabled():
            gid = _capi.DevBuf((self.num_elements, self.Np), np.int32, device=ctx.device)
            g2l = _capi.DevBuf(n, np.int64, device=ctx.device)
            gst = _capi.DevBuf(n + 1, np.int64, device=ctx.device)
            ng = C.c_int64(0)
            _capi.check(lib.semb_numbering(ctx.device, n, self._coords[0].ptr, self._coords[1].ptr, self._coords[2].ptr, 1e-12,
                                           gid.ptr, g2l.ptr, gst.ptr, C.byref(ng), _capi.DEVICE))
            self.num_unique = int(ng.value)
            self._gid, self._g2l = _NodeArray(dev=gid), _NodeArray(dev=g2l)
            self._gst = _NodeArray(dev=gst, count=self.num_unique + 1)
            _capi.check(lib.semb_set_gather_scatter(ctx.handle, g2l.ptr, gst.ptr, self.num_unique, _capi.DEVICE))
        else:
            self.global_id, self.global_to_local, self.global_start = global_numbering_from_coordinates(self.xe, self.ye, self.ze)
            self._upload_numbering(ctx)
        ctx.have.add("gs")
        if self.comm.size > 1:
            import time

            t0 = time.perf_counter()
            self._build_halo()
            t1 = time.perf_counter()
            self._attach_halo(ctx)
            self.halo_timings = dict(self.halo_timings, build_halo=round(t1 - t0, 3),
                                     attach_total=round(time.perf_counter() - t1, 3))

    def _upload_numbering(self, ctx):
        g2l, gst = self.global_to_local, self.global_start
        self.num_unique = int(gst.size - 1)
        _capi.check(ctx.lib.semb_set_gather_scatter(ctx.handle, g2l.ctypes.data, gst.ctypes.data, gst.size - 1, _capi.HOST))

    def _build_halo(self):
        """Interface tables: only the leading ``_n_if_elems`` elements (those touching a vertex shared with
        another rank) can hold interface nodes, so only their coordinates and ids come to the host."""
        ne, Np = self._n_if_elems, self.Np
        cand = distributed.candidate_node_mask(self.elem_to_vert_map[:ne], self._shared_vertex, self.N)
        lead = []
        for a in (*self._coords, self._gid):
            if a._host is not None:
                lead.append(a._host.reshape(self.num_elements, Np)[:ne])
            else:
                lead.append(a._dev.to_host(ne * Np).reshape(ne, Np))
        # the union of all ranks' candidates is numbered with the same rule -- on the device when set-up runs there
        number = global_numbering_device if device_setup_enabled() else global_numbering_from_coordinates
        timed = _TimedComm(self.comm)
        self._halo = distributed.build_halo(timed, lead[0], lead[1], lead[2], cand, lead[3], number)
        self.halo_timings = dict(self.halo_timings, build_halo_allgather=round(timed.seconds, 3))

    def setup_mask(self):
        """Homogeneous Dirichlet mask of the bounding box (sempy/mesh.py:449-477); with several ranks the box
        of the whole mesh."""
        self._mask_host = None
        if self._ctx is None or not device_setup_enabled():
            bbox = None
            if self.comm.size > 1:
                ext = self.comm.allgather([(float(c.min()), float(c.max())) for c in (self.xe, self.ye, self.ze)])
                bbox = [(min(e[a][0] for e in ext), max(e[a][1] for e in ext)) for a in range(3)]
            self.mask = bounding_box_mask(self.xe, self.ye, self.ze, bbox=bbox)
            return
        ctx = self._ctx
        n = self.num_elements * self.Np
        ext = []
        for a in self._coords:
            lo_hi = np.empty(2)
            _capi.check(ctx.lib.semb_minmax(ctx.device, n, a.ptr, lo_hi.ctypes.data, _capi.DEVICE))
            ext.append((float(lo_hi[0]), float(lo_hi[1])))
        if self.comm.size > 1:
            every = self.comm.allgather(ext)
            ext = [(min(e[a][0] for e in every), max(e[a][1] for e in every)) for a in range(3)]
        bbox6 = _capi.as_f64([v for pair in ext for v in pair])
        _capi.check(ctx.lib.semb_setup_mask_box(ctx.handle, self._coords[0].ptr, self._coords[1].ptr, self._coords[2].ptr,
                                                bbox6.ctypes.data, 1e-8, _capi.DEVICE))
        ctx.have.add("mask")

    # -- device context ---------------------------------------------------------
    _NEED_HINT = {"geom": "calc_geometric_factors()", "gs": "establish_global_numbering()", "mask": "setup_mask()"}

    def _context(self, need=("D", "geom", "gs", "mask")):
        """The semb_ctx of this mesh (created with the order N known); raises if a part of the operator
        state that the caller needs has not been set up."""
        if self._ctx is None:
            if not hasattr(self, "Nq"):
                raise RuntimeError("call find_physical_coordinates(N) first")
            self._ctx = _Context(_capi.current_device(), self.num_elements, self.Nq)
            D = _capi.as_f64(reference_derivative_matrix(self.N))
            _capi.check(self._ctx.lib.semb_set_derivative(self._ctx.handle, D.ctypes.data))
            self._ctx.have.add("D")
            # host-side state set before the context existed (CPU-prepared meshes)
            if self._mask_host is not None:
                self.mask = self._mask_host
        for key in need:
            if key not in self._ctx.have:
                raise RuntimeError(f"call {self._NEED_HINT.get(key, key)} first")
        return self._ctx

    def _attach_halo(self, ctx):
        """Communicator (once per context) and this rank's interface tables."""
        import time

        lib = ctx.lib
        tick = [time.perf_counter()]
        marks = {}

        def mark(name):
            now = time.perf_counter()
            marks[name] = round(now - tick[0], 3)
            tick[0] = now

        p2p = os.environ.get("SEMB_P2P", "1") != "0"
        if not getattr(ctx, "comm_ready", False):
            if p2p:
                # peer-memory transport: no NCCL communicator is created at all
                _capi.check(lib.semb_comm_init(ctx.handle, self.comm.size, self.comm.rank, None))
            else:
                uid = C.create_string_buffer(128)
                if self.comm.rank == 0:
                    _capi.check(lib.semb_comm_unique_id(uid))
                raw = self.comm.bcast(bytes(uid.raw), root=0)
                _capi.check(lib.semb_comm_init(ctx.handle, self.comm.size, self.comm.rank, raw))
            ctx.comm_ready = True
        mark("comm_init")
        h = self._halo
        keep = [np.ascontiguousarray(h[k]) for k in ("nb_rank", "nb_count", "send_start", "send_ids", "src_start",
                                                      "src_off", "copy_start", "copy_ids")]
        _capi.check(lib.semb_set_halo(ctx.handle, keep[0].size, keep[0].ctypes.data, keep[1].ctypes.data,
                                      keep[2].ctypes.data, keep[3].ctypes.data, h["n_if"], keep[4].ctypes.data,
                                      keep[5].ctypes.data, keep[6].ctypes.data, keep[7].ctypes.data, self._n_if_elems))
        mark("set_halo")
        if p2p:
            # peer-memory transport: publish this rank's receive region, map everybody else's
            handle = C.create_string_buffer(64)
            cap = max(self.comm.allgather(int(keep[3].size and h["send_start"].size - 1)))
            _capi.check(lib.semb_p2p_alloc(ctx.handle, cap, handle))
            everyone = self.comm.allgather((bytes(handle.raw), [int(q) for q in h["nb_rank"]],
                                            [int(n) for n in h["nb_count"]]))
            peer_off = distributed.peer_receive_offsets([(e[1], e[2]) for e in everyone], self.comm.rank, h["nb_rank"])
            blob = b"".join(e[0] for e in everyone)
            mark("p2p_alloc_allgather")
            _capi.check(lib.semb_p2p_attach(ctx.handle, blob, peer_off.ctypes.data))
            mark("p2p_attach")
        self.halo_timings = dict(self.halo_timings, **marks)

    def _vector_args(self, x, name):
        """(pointer, mem flag, keep-alive) for an in-place vector argument."""
        size = self.num_elements * self.Np
        if _capi.is_torch_cuda(x):
            import torch

            if x.dtype != torch.float64 or not x.is_contiguous() or x.numel() != size:
                raise ValueError(f"{name}: need a contiguous float64 CUDA tensor of {size} entries")
            if x.data_ptr() % 16:
                raise ValueError(f"{name}: CUDA tensors must be 16-byte aligned (a view with an odd element offset is not)")
            return x.data_ptr(), _capi.DEVICE, x
        if not isinstance(x, np.ndarray) or x.dtype != np.float64 or not x.flags.c_contiguous or x.size != size:
            raise ValueError(f"{name}: need a C-contiguous float64 NumPy array of {size} entries")
        return x.ctypes.data, _capi.HOST, x

    def _bind_stream(self, ctx, mem):
        if mem == _capi.DEVICE:
            import torch

            # torch's default stream has handle 0, which semb_set_stream reads as "the context's own
            # stream": name the legacy default stream explicitly (cudaStreamLegacy == 0x1).
            handle = torch.cuda.current_stream().cuda_stream
            _capi.check(ctx.lib.semb_set_stream(ctx.handle, handle if handle else 1))
        else:
            _capi.check(ctx.lib.semb_set_stream(ctx.handle, None))

    # -- hot-path operations ------------------------------------------------------
    def dssum(self, x):
        """In-place direct-stiffness sum, returns x (sempy/mesh.py:444-447)."""
        ctx = self._context(need=("gs",))
        p, mem, _keep = self._vector_args(x, "dssum")
        self._bind_stream(ctx, mem)
        _capi.check(ctx.lib.semb_gs(ctx.handle, p, mem))
        return x

    def apply_mask(self, x):
        """In-place x[mask == 0] = 0, returns x (sempy/mesh.py:487-494)."""
        ctx = self._context(need=("mask",))
        p, mem, _keep = self._vector_args(x, "apply_mask")
        self._bind_stream(ctx, mem)
        _capi.check(ctx.lib.semb_mask(ctx.handle, p, mem))
        return x

    def operator(self, mask=True, assemble=True):
        """Device operator p -> [dssum][mask] A p for sempy_b200.iterative.cg/pcg."""
        return _DeviceOperator(self, mask, assemble)

    # -- getters (sempy/mesh.py:479-527) --------------------------------------------
    def get_global_to_local_map(self):
        return self.global_to_local, self.global_start

    def get_mask_ids(self):
        return np.flatnonzero(self.mask == 0)

    def get_x(self):
        return self.xe.reshape((self.num_elements * self.Np,))

    def get_y(self):
        return self.ye.reshape((self.num_elements * self.Np,))

    def get_z(self):
        return self.ze.reshape((self.num_elements * self.Np,))

    def get_rmult(self):
        return self.rmult

    def get_geom(self):
        return np.array(self.geom).reshape((self.num_elements, self.ndim, self.ndim, self.Np))

    def get_jaco(self):
        return np.array(self.jaco).reshape((self.num_elements * self.Np,))

    def get_mass(self):
        return np.array(self.mass).reshape((self.num_elements * self.Np,))


class _DeviceOperator:
    """A as a device callback: used by iterative.cg/pcg without host copies."""

    def __init__(self, mesh, mask, assemble):
        self.mesh = mesh
        self.flags = (_capi.AX_MASK if mask else 0) | (_capi.AX_GS if assemble else 0)

    def apply_device(self, d_in, d_out, n):
        need = ["D", "geom"] + (["mask"] if self.flags & _capi.AX_MASK else []) + (["gs"] if self.flags & _capi.AX_GS else [])
        ctx = self.mesh._context(need=tuple(need))
        _capi.check(ctx.lib.semb_set_stream(ctx.handle, None))
        _capi.check(ctx.lib.semb_ax(ctx.handle, d_in, d_out, self.flags, _capi.DEVICE))
        _capi.check(ctx.lib.semb_synchronize(ctx.handle))

    def __call__(self, p):
        from sempy_b200.elliptic import _ax

        return _ax(self.mesh, p, self.flags)


def load_mesh(fname):
    """Mesh from the package's ``meshes`` directory (sempy/mesh.py:530-533)."""
    dir_path = os.path.dirname(os.path.realpath(__file__))
    return Mesh(os.path.join(dir_path, "meshes", fname))


def box_mesh(nx, ny=None, nz=None, perturb=0.0, seed=0, comm=None):
    """Synthetic box mesh without a file round trip."""
    return Mesh.from_arrays(*box_mesh_arrays(nx, ny, nz, perturb, seed), comm=comm)
