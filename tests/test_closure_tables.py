"""CPU test of the schedule of the experimental fused operator + gather-scatter kernel
(sempy_b200/csrc/closure_tables.h through semb_host_closure_tables): segments grouped by the element that
closes them.  The kernel's algorithm is replayed sequentially in NumPy and compared with the oracle's
gather-scatter."""

import ctypes as C

import numpy as np
import pytest

from conftest import MESH_CASES, load_golden
from oracle import sem_oracle as so
from sempy_b200 import _capi


def closure_tables(E, Np, g2l, gstart):
    lib = _capi.lib()
    g2l = np.ascontiguousarray(g2l, dtype=np.int64)
    gstart = np.ascontiguousarray(gstart, dtype=np.int64)
    sizes = np.zeros(3, dtype=np.int64)
    _capi.check(lib.semb_host_closure_tables(E, Np, g2l.ctypes.data, gstart.ctypes.data, gstart.size - 1,
                                             sizes.ctypes.data, None, None, None, None, None, None, None))
    t = dict(base=np.zeros(E + 1, np.int32), n2=np.zeros(E, np.uint16), n4=np.zeros(E, np.uint16),
             n8=np.zeros(E, np.uint16), ids=np.zeros(max(int(sizes[0]), 1), np.int32),
             dep_start=np.zeros(E + 1, np.int32), dep_elem=np.zeros(max(int(sizes[1]), 1), np.int32))
    _capi.check(lib.semb_host_closure_tables(E, Np, g2l.ctypes.data, gstart.ctypes.data, gstart.size - 1,
                                             sizes.ctypes.data, t["base"].ctypes.data, t["n2"].ctypes.data,
                                             t["n4"].ctypes.data, t["n8"].ctypes.data, t["ids"].ctypes.data,
                                             t["dep_start"].ctypes.data, t["dep_elem"].ctypes.data))
    t["ids"] = t["ids"][: int(sizes[0])]
    t["dep_elem"] = t["dep_elem"][: int(sizes[1])]
    t["other"] = int(sizes[2])
    return t


def replay(t, x, E, Np, lag):
    """What ax8_fused_kernel does, one work item at a time: item e publishes element e, item e >= lag closes
    element e - lag after checking that its dependencies are published."""
    w = np.full_like(x, np.nan)
    published = np.zeros(E, dtype=bool)
    for item in range(E + lag):
        if item < E:
            w[item * Np:(item + 1) * Np] = x[item * Np:(item + 1) * Np]
            published[item] = True
        ec = item - lag
        if ec < 0:
            continue
        deps = t["dep_elem"][t["dep_start"][ec]:t["dep_start"][ec + 1]]
        assert published[ec] and published[deps].all() and (deps < ec).all()
        ids = t["ids"][t["base"][ec]:t["base"][ec + 1]]
        n2, n4, n8 = int(t["n2"][ec]), int(t["n4"][ec]), int(t["n8"][ec])
        assert ids.size == 2 * n2 + 4 * n4 + 8 * n8
        off = 0
        for length, count in ((2, n2), (4, n4), (8, n8)):
            for s in range(count):
                seg = ids[off:off + length]
                assert (np.diff(seg) > 0).all() and seg[-1] // Np == ec
                w[seg] = w[seg].sum()
                off += length
    return w


@pytest.mark.parametrize("case", MESH_CASES)
def test_closure_tables_on_reference_meshes(case):
    g = load_golden(case)
    E, Np = g["xe"].shape
    t = closure_tables(E, Np, g["global_to_local"], g["global_start"])
    x = np.random.default_rng(3).standard_normal(E * Np)
    ref = so.gather_scatter(x, g["global_to_local"], g["global_start"])
    for lag in (1, 3):
        got = replay(t, x, E, Np, lag)
        if t["other"] == 0:
            assert np.max(np.abs(got - ref)) <= 1e-14 * np.max(np.abs(ref))
    # every segment of length 2, 4 or 8 appears exactly once
    lens = np.diff(g["global_start"])
    assert t["ids"].size == int(lens[np.isin(lens, (2, 4, 8))].sum())
    assert t["other"] == int(((lens > 1) & ~np.isin(lens, (2, 4, 8))).sum())
    assert np.unique(t["ids"]).size == t["ids"].size


def test_closure_tables_generated_box():
    from sempy_b200.mesh import box_mesh_arrays, global_numbering_from_coordinates

    pts, hexes = box_mesh_arrays(4, 3, 5, perturb=0.3, seed=1)
    N = 3
    xe, ye, ze = so.physical_coordinates(*so.element_vertices(pts, hexes), N)
    _, g2l, gs = global_numbering_from_coordinates(xe, ye, ze)
    E, Np = xe.shape
    t = closure_tables(E, Np, g2l, gs)
    assert t["other"] == 0
    x = np.random.default_rng(4).standard_normal(E * Np)
    got = replay(t, x, E, Np, lag=7)
    ref = so.gather_scatter(x, g2l, gs)
    assert np.max(np.abs(got - ref)) <= 1e-14 * np.max(np.abs(ref))
    # dependencies of an interior element of a structured box: at most 7 lower neighbours
    ndep = np.diff(t["dep_start"])
    assert ndep.max() <= 7 and ndep[0] == 0
