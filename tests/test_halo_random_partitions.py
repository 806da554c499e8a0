"""Interface ("halo") tables on ARBITRARY partitions, all ranks in one process (threads + a barrier-based
all-gather): elements are shuffled before being cut into ranges, so ranks touch along faces, edges and
single vertices, and nodes are shared by up to eight ranks.  The NumPy statement of the device kernels
(distributed.emulate_gather_scatter) must reproduce the oracle's global gather-scatter on every rank."""

import threading

import numpy as np
import pytest

from oracle import sem_oracle as so
from sempy_b200 import distributed as dd
from sempy_b200 import mesh as pm


class ThreadComm:
    """In-process communicator for `size` threads."""

    def __init__(self, rank, size, shared):
        self.rank, self.size, self._s = rank, size, shared

    def allgather(self, obj):
        s = self._s
        s["slots"][self.rank] = obj
        s["barrier"].wait()
        out = list(s["slots"])
        s["barrier"].wait()
        return out

    def bcast(self, obj, root=0):
        return self.allgather(obj)[root]


def run_partition(dims, N, size, perturb, seed):
    pts, hexes = pm.box_mesh_arrays(*dims, perturb=perturb, seed=seed)
    rng = np.random.default_rng(seed)
    hexes = hexes[rng.permutation(len(hexes))]  # arbitrary partition = contiguous ranges of a shuffled list
    E = len(hexes)
    gxe, gye, gze = so.physical_coordinates(*so.element_vertices(pts, hexes), N)
    Np = gxe.shape[1]
    _, g2l, gs = so.global_numbering(gxe, gye, gze)
    xg = rng.standard_normal(E * Np)
    ref = so.gather_scatter(xg, g2l, gs).reshape(E, Np)
    ranges = dd.slab_ranges(E, size)
    shared_v = dd.shared_vertex_flags(hexes, ranges, len(pts))
    shared = {"slots": [None] * size, "barrier": threading.Barrier(size)}
    errs, info = [None] * size, [None] * size

    def worker(rank):
        try:
            comm = ThreadComm(rank, size, shared)
            lo, hi = ranges[rank]
            xe, ye, ze = gxe[lo:hi], gye[lo:hi], gze[lo:hi]
            gid, l2l, ls = pm.global_numbering_from_coordinates(xe, ye, ze)
            cand = dd.candidate_node_mask(hexes[lo:hi], shared_v, N)
            halo = dd.build_halo(comm, xe, ye, ze, cand, gid, pm.global_numbering_from_coordinates)
            got = dd.emulate_gather_scatter(comm, halo, xg.reshape(E, Np)[lo:hi].reshape(-1), l2l, ls)
            errs[rank] = float(np.abs(got - ref[lo:hi].reshape(-1)).max() / np.abs(ref).max())
            info[rank] = (int(halo["n_if"]), len(halo["nb_rank"]), int(np.diff(halo["src_start"]).max()) if halo["n_if"] else 0)
        except BaseException as exc:  # surface the failure, do not leave the others at the barrier
            errs[rank] = exc
            shared["barrier"].abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errs:
        if isinstance(e, BaseException):
            raise e
    return errs, info


@pytest.mark.parametrize("dims,N,size,perturb,seed", [
    ((3, 3, 3), 2, 4, 0.0, 0), ((3, 3, 3), 3, 8, 0.3, 1), ((4, 2, 3), 2, 5, 0.2, 2), ((2, 2, 2), 4, 8, 0.3, 3),
    ((5, 3, 2), 1, 6, 0.0, 4), ((3, 3, 3), 2, 27, 0.1, 5),
])
def test_random_partitions(dims, N, size, perturb, seed):
    errs, info = run_partition(dims, N, size, perturb, seed)
    assert max(errs) < 1e-14
    # shuffled partitions make ranks share nodes with many neighbours; a vertex node can have up to 8 sources
    assert max(i[1] for i in info) >= 2
    assert max(i[2] for i in info) >= 3


def test_one_element_per_rank_all_copies_remote():
    """27 ranks, one element each: every shared node has all its other copies on other ranks."""
    errs, info = run_partition((3, 3, 3), 2, 27, 0.1, 5)
    assert max(errs) < 1e-14
    assert max(i[2] for i in info) == 8  # the centre vertex of the 2x2x2 blocks is shared by eight ranks
