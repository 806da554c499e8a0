"""Multi-rank host logic on CPU (gloo, world sizes 2 and 3): slab partition,
interface discovery with the reference's numbering rule, halo tables.  The
device semantics of the tables are restated in NumPy
(distributed.emulate_gather_scatter) and checked against the oracle's global
gather-scatter on the unpartitioned mesh."""

import os
import socket

import numpy as np
import pytest

from oracle import sem_oracle as so
from sempy_b200 import distributed as dd
from sempy_b200 import mesh as pm


def test_slab_ranges():
    assert dd.slab_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert dd.slab_ranges(8, 8)[-1] == (7, 8)
    r = dd.slab_ranges(64 * 64 * 64, 8)
    assert all(hi - lo == 32768 for lo, hi in r)


def test_candidates_cover_the_interface_nodes():
    pts, hexes = pm.box_mesh_arrays(3, 2, 4)
    ranges = dd.slab_ranges(len(hexes), 2)
    shared = dd.shared_vertex_flags(hexes, ranges, len(pts))
    # vertices on the plane z = 0.5 (k = 2 of 4) are the shared ones
    assert np.array_equal(shared, np.isclose(pts[:, 2], 0.5))
    N = 3
    cand = dd.candidate_node_mask(hexes[:12], shared, N)
    xe, ye, ze = so.physical_coordinates(*so.element_vertices(pts, hexes[:12]), N)
    assert np.array_equal(cand, np.isclose(ze, 0.5))
    order, n_if = dd.interface_first_order(hexes[:12], shared)
    assert n_if == 6 and sorted(order[:6]) == list(range(6, 12))


def test_peer_receive_offsets():
    # four slabs: 0 | 1 | 2 | 3 with interface sizes 10, 20, 30
    layouts = [([1], [10]), ([0, 2], [10, 20]), ([1, 3], [20, 30]), ([2], [30])]
    assert list(dd.peer_receive_offsets(layouts, 0, [1])) == [0]        # rank 1 receives rank 0's block first
    assert list(dd.peer_receive_offsets(layouts, 1, [0, 2])) == [0, 0]
    assert list(dd.peer_receive_offsets(layouts, 2, [1, 3])) == [10, 0]  # after rank 0's 10 entries at rank 1
    assert list(dd.peer_receive_offsets(layouts, 3, [2])) == [20]
    assert dd.peer_receive_offsets(layouts, 3, [2]).dtype == np.int64


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, dims, perturb, N, out, side_group=False):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        if side_group:
            # as under torchrun with the NCCL backend: the control plane must move to a gloo group of its own
            dist.get_backend = lambda *a, **k: "nccl"
        comm = dd.TorchComm()
        assert (comm._group is not None) == side_group
        pts, hexes = pm.box_mesh_arrays(*dims, perturb=perturb, seed=2)
        m = pm.Mesh.from_arrays(pts, hexes, comm=comm)
        # coordinates from the oracle (the product computes them on the GPU); numbering rule is the product's
        m.N, m.Nq, m.Np = N, N + 1, (N + 1) ** 3
        m.xe, m.ye, m.ze = so.physical_coordinates(m.x, m.y, m.z, N)
        m.global_id, m.global_to_local, m.global_start = pm.global_numbering_from_coordinates(m.xe, m.ye, m.ze)
        cand = dd.candidate_node_mask(m.elem_to_vert_map, m._shared_vertex, N)
        ne = m._n_if_elems  # only the leading interface elements are handed over, as Mesh._build_halo does
        halo = dd.build_halo(comm, m.xe[:ne], m.ye[:ne], m.ze[:ne], cand[:ne], m.global_id[:ne],
                             pm.global_numbering_from_coordinates)
        # global reference: the whole mesh numbered at once, seeded vector
        gxe, gye, gze = so.physical_coordinates(*so.element_vertices(pts, hexes), N)
        _, g2l, gs = so.global_numbering(gxe, gye, gze)
        xg = np.random.default_rng(11).standard_normal(gxe.size)
        ref = so.gather_scatter(xg, g2l, gs).reshape(len(hexes), -1)[m.elem_global].reshape(-1)
        mine = xg.reshape(len(hexes), -1)[m.elem_global].reshape(-1)
        got = dd.emulate_gather_scatter(comm, halo, mine, m.global_to_local, m.global_start)
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        # multiplicities across ranks
        rm = 1.0 / dd.emulate_gather_scatter(comm, halo, np.ones(mine.size), m.global_to_local, m.global_start)
        rm_ref = so.multiplicity_inverse(g2l, gs).reshape(len(hexes), -1)[m.elem_global].reshape(-1)
        # interface elements come first
        el_has = np.zeros(m.get_num_elems(), bool)
        el_has[halo["copy_ids"] // m.Np] = True
        ok_order = bool(el_has[m._n_if_elems:].sum() == 0)
        # global Dirichlet mask from the global bounding box
        m.setup_mask()
        mask_ref = so.dirichlet_mask(gxe, gye, gze).reshape(len(hexes), -1)[m.elem_global].reshape(-1)
        res = comm.allgather((err, bool(np.array_equal(rm, rm_ref)), ok_order, bool(np.array_equal(m.mask, mask_ref)),
                              int(halo["n_if"]), [int(q) for q in halo["nb_rank"]]))
        if rank == 0:
            np.save(out, np.array([[r[0], r[1], r[2], r[3], r[4], len(r[5])] for r in res], dtype=np.float64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,dims,perturb,N", [(2, (3, 3, 4), 0.3, 3), (3, (2, 3, 6), 0.0, 4), (2, (2, 2, 2), 0.25, 7)])
def test_halo_tables_reproduce_global_gather_scatter(size, dims, perturb, N, tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(size, _free_port(), dims, perturb, N, out), nprocs=size, join=True)
    res = np.load(out)
    assert res.shape[0] == size
    assert res[:, 0].max() < 1e-15          # distributed dssum == global dssum
    assert res[:, 1].all() and res[:, 2].all() and res[:, 3].all()
    plane = (dims[0] * N + 1) * (dims[1] * N + 1)
    expect_if = [plane * (1 if r in (0, size - 1) else 2) for r in range(size)]
    assert list(res[:, 4].astype(int)) == expect_if
    assert list(res[:, 5].astype(int)) == [1 if r in (0, size - 1) else 2 for r in range(size)]


def test_control_plane_on_a_side_group(tmp_path):
    """With NCCL as the default backend TorchComm opens a gloo group for its pickled host objects (no
    host -> device -> host round trip, no NCCL connection set-up inside the mesh set-up); same tables."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), (2, 3, 4), 0.2, 3, out, True), nprocs=2, join=True)
    res = np.load(out)
    assert res[:, 0].max() < 1e-15 and res[:, 1].all() and res[:, 2].all() and res[:, 3].all()


@pytest.mark.parametrize("size,dims,perturb,N", [(2, (3, 3, 3), 0.3, 3), (3, (2, 2, 5), 0.2, 2), (4, (3, 3, 3), 0.0, 2)])
def test_non_planar_interfaces(size, dims, perturb, N, tmp_path):
    """Element counts that do not divide into whole layers: the slab boundary cuts through a layer, so a
    rank's interface is L-shaped and ranks can share nodes along edges and at corners only."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(size, _free_port(), dims, perturb, N, out), nprocs=size, join=True)
    res = np.load(out)
    assert res.shape[0] == size
    assert res[:, 0].max() < 1e-15 and res[:, 1].all() and res[:, 2].all() and res[:, 3].all()
    assert (res[:, 4] > 0).all()


def test_first_occurrence_by_reversed_assignment():
    """build_halo finds the first send entry of every interface node with a fancy assignment written back to
    front (the smallest index wins because NumPy keeps the last value assigned to a repeated index); pinned
    here against np.minimum.at so that a change of that behaviour is noticed."""
    rng = np.random.default_rng(0)
    for _ in range(50):
        n_if, n_send = int(rng.integers(1, 60)), int(rng.integers(1, 300))
        node_of = rng.integers(0, n_if, n_send)
        want = np.full(n_if, n_send, dtype=np.int64)
        np.minimum.at(want, node_of, np.arange(n_send))
        got = np.full(n_if, n_send, dtype=np.int64)
        got[node_of[::-1]] = np.arange(n_send - 1, -1, -1)
        assert np.array_equal(got, want)
