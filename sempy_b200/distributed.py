"""Host side of the multi-GPU path: element slabs and interface ("halo") tables.

The reference is single-process (it hands gslib ``comm = 0``, sempy/mesh.py:437,
and has no partitioner), so there is no reference code to mirror here; the
semantics to preserve are those of its gather-scatter: after ``dssum`` every
copy of a node -- on whichever GPU it lives -- holds the sum over all copies.

Layout: one process per GPU; rank r owns a contiguous range of the global
element list (a z-slab for the synthetic boxes, whose elements are ordered
k-slowest).  Every rank numbers its own nodes with the reference's rule
(``global_numbering_from_coordinates``); nodes shared between ranks are found by
applying the *same rule* to the union of all ranks' boundary candidates, which
every rank evaluates identically, so both sides of an interface list their
shared nodes in the same order without any further negotiation.

Everything in this file is NumPy + a tiny communicator interface, so it runs
(and is tested) on CPU with the gloo backend; the device work is in
libsemb200 (semb_comm_init / semb_set_halo, include/semb200.h).
"""

from __future__ import annotations

import numpy as np

_GMSH_TO_LEX = [0, 1, 3, 2, 4, 5, 7, 6]


class SerialComm:
    """Single-rank stand-in."""

    rank = 0
    size = 1

    def allgather(self, obj):
        return [obj]

    def bcast(self, obj, root=0):
        return obj


class TorchComm:
    """torch.distributed as the control-plane communicator (gloo or nccl backend)."""

    def __init__(self, group=None):
        """Collective when the default backend is NCCL and no group is given: every rank must construct it.

        The control plane moves pickled host objects (candidate coordinates, IPC handles).  Over NCCL that
        is a host -> device -> host round trip and the first collective pays NCCL's lazy connection set-up
        inside the mesh set-up (0.6-2.5 s of `build_halo` on 2-8 ranks); a gloo group next to the NCCL one
        keeps it on the host."""
        import torch.distributed as dist

        self._dist = dist
        if group is None and dist.get_backend() == "nccl":
            try:
                group = dist.new_group(backend="gloo")
            except Exception:  # no usable interface for gloo: stay on the default group
                group = None
        self._group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def allgather(self, obj):
        out = [None] * self.size
        self._dist.all_gather_object(out, obj, group=self._group)
        return out

    def bcast(self, obj, root=0):
        box = [obj]
        self._dist.broadcast_object_list(box, src=root, group=self._group)
        return box[0]


def slab_ranges(num_elems, nranks):
    """Contiguous, balanced element ranges [(lo, hi), ...]."""
    base, extra = divmod(num_elems, nranks)
    out, lo = [], 0
    for r in range(nranks):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shared_vertex_flags(hexes, ranges, num_points):
    """True for mesh vertices used by elements of more than one rank."""
    owner_lo = np.full(num_points, np.iinfo(np.int64).max, dtype=np.int64)
    owner_hi = np.full(num_points, -1, dtype=np.int64)
    for r, (lo, hi) in enumerate(ranges):
        v = np.unique(hexes[lo:hi])
        owner_lo[v] = np.minimum(owner_lo[v], r)
        owner_hi[v] = np.maximum(owner_hi[v], r)
    return owner_hi > owner_lo


def candidate_node_mask(local_hexes, shared_vertex, N):
    """(E, Np) bool: GLL nodes lying on an element face, edge or corner all of
    whose mesh vertices are shared with another rank -- a superset of the nodes
    that really have copies elsewhere."""
    n = N + 1
    sv = shared_vertex[local_hexes[:, _GMSH_TO_LEX]].reshape(-1, 2, 2, 2)  # [e, kk, jj, ii]
    cat = np.full(n, 2)
    cat[0], cat[-1] = 0, 1
    sel = [slice(0, 1), slice(1, 2), slice(0, 2)]
    table = np.zeros((sv.shape[0], 3, 3, 3), dtype=bool)  # [e, ck, cj, ci]
    for ck in range(3):
        for cj in range(3):
            for ci in range(3):
                if ck == cj == ci == 2:
                    continue
                table[:, ck, cj, ci] = sv[:, sel[ck], sel[cj], sel[ci]].all(axis=(1, 2, 3))
    K, J, I = np.meshgrid(cat, cat, cat, indexing="ij")
    return table[:, K.ravel(), J.ravel(), I.ravel()]


def interface_first_order(local_hexes, shared_vertex):
    """Permutation putting the elements that touch a shared vertex first
    (stable), and their count."""
    touches = shared_vertex[local_hexes].any(axis=1)
    order = np.concatenate((np.flatnonzero(touches), np.flatnonzero(~touches)))
    return order, int(touches.sum())


def build_halo(comm, xe, ye, ze, cand, global_id, numbering):
    """Interface tables of this rank (arguments of semb_set_halo).

    xe, ye, ze, cand, global_id: (E', Np) arrays of the LEADING E' elements of this rank -- all of them,
    or only those that touch a vertex shared with another rank (``interface_first_order`` puts them
    first): every copy of an interface node lies in such an element and is itself a candidate, so the
    interior of a slab never has to leave the device.  cand: candidate mask; global_id: this rank's
    LOCAL numbering; numbering: the coordinate-numbering function (the reference's rule).
    Collective: one allgather of the candidates' coordinates ((3, n) arrays).
    """
    rank, size = comm.rank, comm.size
    empty = dict(nb_rank=np.zeros(0, np.int32), nb_count=np.zeros(0, np.int64), send_start=np.zeros(1, np.int64),
                 send_ids=np.zeros(0, np.int32), n_if=0, src_start=np.zeros(1, np.int64), src_off=np.zeros(0, np.int32),
                 copy_start=np.zeros(1, np.int64), copy_ids=np.zeros(0, np.int32))
    if size == 1:
        return empty
    x, y, z = xe.reshape(-1), ye.reshape(-1), ze.reshape(-1)
    # candidate copies grouped by local id, ascending local index inside a group (fixed summation order)
    cpos = np.flatnonzero(cand.reshape(-1))
    cgid = global_id.reshape(-1)[cpos]
    by_id = np.argsort(cgid, kind="stable")
    cpos, cgid = cpos[by_id], cgid[by_id]
    first = np.flatnonzero(np.concatenate(([True], cgid[1:] != cgid[:-1]))) if cgid.size else np.zeros(0, np.int64)
    ncopy_all = np.diff(np.concatenate((first, [cgid.size])))
    rep = cpos[first]                                                         # one local copy of each
    mine = np.stack([x[rep], y[rep], z[rep]])                                  # (3, n): coordinate-major
    parts = comm.allgather(mine)
    if mine.shape[1] == 0:
        return empty
    # Only points inside the bounding box of MY candidates can share a node with me (the matching
    # tolerance is 1e-6 in distance): drop the rest before sorting -- whole ranks first (with slabs all but
    # the two neighbours), then point by point.  Every member of each of my groups survives on every rank
    # that owns one of them, and the order of two groups is decided by their own members' coordinates, so
    # both sides of an interface still list their shared nodes in the same order.
    lo, hi = mine.min(axis=1) - 1e-5, mine.max(axis=1) + 1e-5
    kept, owner, pos = [], [], []
    for q, part in enumerate(parts):
        if part.shape[1] == 0:
            continue
        if q == rank:
            sel = np.arange(part.shape[1])
        else:
            if np.any(part.min(axis=1) > hi) or np.any(part.max(axis=1) < lo):
                continue
            near = (part[0] >= lo[0]) & (part[0] <= hi[0])
            for a in (1, 2):
                near &= part[a] >= lo[a]
                near &= part[a] <= hi[a]
            sel = np.flatnonzero(near)
            part = part[:, sel]
        kept.append(part)
        owner.append(np.full(sel.size, q, dtype=np.int64))
        pos.append(sel)
    allc, owner, pos = np.concatenate(kept, axis=1), np.concatenate(owner), np.concatenate(pos)
    # the reference's rule on the union
    gid_all, _, gs_all = numbering(allc[0:1], allc[1:2], allc[2:3])
    grp = gid_all.reshape(-1).astype(np.int64) - 1
    ngrp = int(gs_all.size - 1)
    multi = np.diff(gs_all)[grp] >= 2
    grp, owner, pos = grp[multi], owner[multi], pos[multi]
    # a group must not hold two entries of one rank (the local numbering already merged those)
    if grp.size and np.bincount(grp * size + owner).max() > 1:
        raise RuntimeError("interface matching: two local nodes of one rank fell into the same global node")
    me = owner == rank
    order = np.argsort(grp[me], kind="stable")
    g_mine = grp[me][order]                     # my interface nodes, ascending group id (canonical order)
    idx_mine = pos[me][order]                   # their rows in first / ncopy_all
    n_if = g_mine.size
    if n_if == 0:
        return empty
    # the other members of my groups: (neighbour rank, group)
    node_of_grp = np.full(ngrp, -1, dtype=np.int64)
    node_of_grp[g_mine] = np.arange(n_if)
    other = (~me) & (node_of_grp[grp] >= 0)
    m_o, g_o = owner[other], grp[other]
    lay = np.argsort(m_o * ngrp + g_o, kind="stable")   # send layout: by neighbour, then by group id
    m_o, g_o = m_o[lay], g_o[lay]
    n_send = m_o.size
    node_of = node_of_grp[g_o]                  # my node behind each send entry
    nb_rank, nb_count = np.unique(m_o, return_counts=True)
    # local copies per interface node (ascending local index)
    lo = first[idx_mine]
    ncopy = ncopy_all[idx_mine].astype(np.int64)
    copy_start = np.concatenate(([0], np.cumsum(ncopy)))
    flat = np.repeat(lo - copy_start[:-1], ncopy) + np.arange(copy_start[-1])
    copy_ids = cpos[flat]
    # send entries -> the copies of their node
    ns = ncopy[node_of]
    send_start = np.concatenate(([0], np.cumsum(ns)))
    sflat = np.repeat(copy_start[node_of] - send_start[:-1], ns) + np.arange(send_start[-1])
    send_ids = copy_ids[sflat]
    # sources per node in ascending rank order: own partial (first send entry of the node) + received ones
    own_off = np.full(n_if, n_send, dtype=np.int64)
    own_off[node_of[::-1]] = np.arange(n_send - 1, -1, -1)   # written back to front: the smallest entry wins
    s_node = np.concatenate((np.arange(n_if), node_of))
    s_rank = np.concatenate((np.full(n_if, rank), m_o))
    s_off = np.concatenate((own_off, n_send + np.arange(n_send)))
    so = np.argsort(s_node * size + s_rank, kind="stable")
    src_off = s_off[so]
    src_start = np.concatenate(([0], np.cumsum(np.bincount(s_node, minlength=n_if))))
    return dict(nb_rank=nb_rank.astype(np.int32), nb_count=nb_count.astype(np.int64),
                send_start=send_start.astype(np.int64), send_ids=send_ids.astype(np.int32), n_if=int(n_if),
                src_start=src_start.astype(np.int64), src_off=src_off.astype(np.int32),
                copy_start=copy_start.astype(np.int64), copy_ids=copy_ids.astype(np.int32))


def peer_receive_offsets(layouts, me, nb_rank):
    """Where this rank's block starts in each neighbour's receive layout (argument of semb_p2p_attach).

    layouts[q] = (neighbour ranks of q, entry counts of q) as q passed them to semb_set_halo; a rank's
    receive region mirrors its send region, block by block in that order, so my block for q starts after
    the blocks of q's neighbours listed before me."""
    out = []
    for q in nb_rank:
        ranks_q, counts_q = layouts[int(q)]
        at = list(ranks_q).index(me)
        out.append(int(sum(counts_q[:at])))
    return np.ascontiguousarray(out, dtype=np.int64)


def emulate_gather_scatter(comm, halo, x, g2l, gstart):
    """NumPy statement of what the device does with the tables (used by the CPU
    tests of this module and as executable documentation): returns dssum(x)."""
    x = x.copy()
    n_send = int(halo["send_start"].size - 1)
    send = np.add.reduceat(x[halo["send_ids"]], halo["send_start"][:-1]) if n_send else np.zeros(0)
    # exchange: neighbour q receives my block for q; blocks are ordered by neighbour rank
    blocks = {}
    off = 0
    for q, cnt in zip(halo["nb_rank"], halo["nb_count"]):
        blocks[int(q)] = send[off: off + cnt]
        off += cnt
    everyone = comm.allgather(blocks)
    recv = np.concatenate([everyone[int(q)][comm.rank] for q in halo["nb_rank"]]) if n_send else np.zeros(0)
    hbuf = np.concatenate((send, recv))
    # local sum of every node, written to all its copies
    seg = np.add.reduceat(x[g2l], gstart[:-1])
    x[g2l] = np.repeat(seg, np.diff(gstart))
    if halo["n_if"]:
        tot = np.add.reduceat(hbuf[halo["src_off"]], halo["src_start"][:-1])
        x[halo["copy_ids"]] = np.repeat(tot, np.diff(halo["copy_start"]))
    return x
