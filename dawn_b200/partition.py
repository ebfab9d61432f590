"""Partitioning of an unstructured mesh over the GPUs of one box (one process per GPU).

The reference has no distribution for unstructured stencils; what it offers is the index-space
convention of ICON: every location type is ordered [interior | halo] and the generated code takes the
boundaries as `UnstructuredSubdomain` splitter indices (driver-includes/unstructured_domain.hpp:23-34,
used by CudaIcoCodeGen.cpp:327-343).  This module produces exactly that from a (renumbered,
dawn_b200/reorder.py) mesh:

  * cells are owned in contiguous id ranges; an edge or a vertex belongs to the LOWEST part among the
    owners of its adjacent cells;
  * a part's local mesh holds its owned elements first and then the halo: every element of the cells
    within `rings` vertex-adjacency rings of its owned cells, numbered by ascending global id;
  * the stencil runs on the whole local mesh; results are valid on owned elements (their complete
    neighbourhood is local, in the original neighbour order, so they are bit-identical to the
    undivided run), halo results are redundant work;
  * only INPUT fields travel: before a step the halo copies are refreshed from their owners
    (`Part.exchange`: gather at the owner, point-to-point send / receive, scatter into the halo rows -
    torch.distributed isend/irecv, i.e. NCCL between GPUs and gloo in the CPU tests).

Everything here is integer work and deterministic: every rank computes the same global partition and
cuts its own part out of it.
"""
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .reorder import TableMesh
from .trimesh import CELL, EDGE, VERTEX, NAMES

LOCS = (CELL, EDGE, VERTEX)
PAIRS = [(CELL, EDGE), (CELL, VERTEX), (EDGE, CELL), (EDGE, VERTEX), (VERTEX, CELL), (VERTEX, EDGE)]
# ::dawn::UnstructuredSubdomain (driver-includes/unstructured_domain.hpp:23)
LATERAL_BOUNDARY, NUDGING, INTERIOR, HALO, END = 0, 1000, 2000, 3000, 4000


def owners(mesh, nparts):
    """{loc: int32 owner per element} - cells in contiguous ranges, edges and vertices by the
    lowest-owner rule; elements without cells (invalid edge slots) are owned by nobody (-1)."""
    nc = mesh.num_cells
    base, rem = divmod(nc, nparts)
    bounds = np.cumsum([0] + [base + (1 if p < rem else 0) for p in range(nparts)])
    own_c = (np.searchsorted(bounds, np.arange(nc), side="right") - 1).astype(np.int32)
    out = {CELL: own_c}
    for loc in (EDGE, VERTEX):
        t = mesh.nbh_table([loc, CELL])
        o = np.where(t >= 0, own_c[np.maximum(t, 0)], nparts).min(axis=1)
        out[loc] = np.where(o == nparts, -1, o).astype(np.int32)
    return out


@dataclass
class Part:
    rank: int
    nparts: int
    mesh: TableMesh                       # local numbering: owned elements first, then the halo
    n_owned: Dict[int, int]
    global_ids: Dict[int, np.ndarray]     # local id -> global id
    send: Dict[int, Dict[int, np.ndarray]] = field(default_factory=dict)  # loc -> peer -> local ids (owned here)
    recv: Dict[int, Dict[int, np.ndarray]] = field(default_factory=dict)  # loc -> peer -> local ids (halo here)

    def splitters(self):
        """splitter indices for `Module.on_mesh(..., splitters=)`: [Interior, Halo) is what this part owns"""
        s = {}
        for loc in LOCS:
            s[(loc, INTERIOR, 0)] = 0
            s[(loc, HALO, 0)] = self.n_owned[loc]
            s[(loc, END, 0)] = self.mesh.num(loc)
        return s

    def to_local(self, loc, a, axis=-1):
        """global field -> this part's local field (element axis `axis`)"""
        return np.take(a, self.global_ids[NAMES[loc]], axis=axis)

    def peers(self):
        return sorted({p for loc in LOCS for p in list(self.send.get(loc, {})) + list(self.recv.get(loc, {}))})

    def exchange(self, loc, tensor, dist, axis=-1):
        """Refreshes the halo entries of `tensor` (a torch tensor, element axis `axis`) from their owners.
        Collective over the neighbouring parts: every part calls it for the same fields in the same order."""
        import torch
        loc = NAMES[loc]
        ops, landing = [], []
        dev = tensor.device
        cache = self.__dict__.setdefault("_dev_ids", {})   # index lists live on the device after the first call

        def on_dev(kind, peer, ids):
            key = (kind, loc, peer, str(dev))
            if key not in cache:
                cache[key] = torch.as_tensor(ids, device=dev)
            return cache[key]

        for p in self.peers():
            ids = self.send.get(loc, {}).get(p)
            if ids is not None and len(ids):
                buf = tensor.index_select(axis, on_dev("send", p, ids)).contiguous()
                ops.append(dist.P2POp(dist.isend, buf, p))
            ids = self.recv.get(loc, {}).get(p)
            if ids is not None and len(ids):
                shape = list(tensor.shape)
                shape[axis] = len(ids)
                buf = torch.empty(shape, dtype=tensor.dtype, device=dev)
                ops.append(dist.P2POp(dist.irecv, buf, p))
                landing.append((on_dev("recv", p, ids), buf))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for ids, buf in landing:
            tensor.index_copy_(axis, ids, buf)
        return tensor


def partition(mesh, nparts, rings=1) -> List[Part]:
    """All parts of `mesh` (every rank computes them all - index work - and keeps its own).  `rings` =
    vertex-adjacency rings of halo cells around the owned cells: 1 covers stencils whose owned results
    reach inputs through at most (element -> adjacent element -> adjacent element), e.g. ICON's nabla2."""
    own = owners(mesh, nparts)
    base = {p: mesh.nbh_table(list(p)) for p in PAIRS}
    c2v, v2c = base[(CELL, VERTEX)], base[(VERTEX, CELL)]
    parts = []
    for r in range(nparts):
        # ---- cells: owned, then `rings` rings of vertex-adjacent cells ----------------------------------
        in_set = own[CELL] == r
        for _ in range(rings):
            verts = np.unique(c2v[in_set][c2v[in_set] >= 0])
            around = v2c[verts]
            grown = in_set.copy()
            grown[around[around >= 0]] = True
            in_set = grown
        local = {CELL: np.nonzero(in_set)[0]}
        # ---- edges / vertices: everything the local cells touch, plus whatever this part owns ---------------
        for loc in (EDGE, VERTEX):
            t = base[(CELL, loc)][local[CELL]]
            mask = np.zeros(mesh.num(loc), dtype=bool)
            mask[t[t >= 0]] = True
            mask |= own[loc] == r
            local[loc] = np.nonzero(mask)[0]
        gids, n_owned, g2l = {}, {}, {}
        for loc in LOCS:
            ids = local[loc]
            mine = own[loc][ids] == r
            order = np.concatenate([ids[mine], ids[~mine]])       # both halves ascending in global id
            gids[loc], n_owned[loc] = order.astype(np.int64), int(mine.sum())
            m = np.full(mesh.num(loc), -1, dtype=np.int64)
            m[order] = np.arange(order.shape[0])
            g2l[loc] = m
        tables = {}
        for (a, b) in PAIRS:
            t = base[(a, b)][gids[a]]
            tables[(a, b)] = np.where(t >= 0, g2l[b][np.maximum(t, 0)], -1).astype(np.int32)
        cxy = getattr(mesh, "centroids", lambda: None)()
        lm = TableMesh(len(gids[CELL]), len(gids[EDGE]), len(gids[VERTEX]), tables,
                       mesh.valid(EDGE)[gids[EDGE]],
                       None if cxy is None else (cxy[0][gids[CELL]], cxy[1][gids[CELL]]))
        parts.append(Part(r, nparts, lm, n_owned, gids))
    # ---- exchange lists: a halo element is received from its owner, in ascending global id on both sides ---
    lookup = []
    for p in parts:
        d = {}
        for loc in LOCS:
            m = np.full(mesh.num(loc), -1, dtype=np.int64)
            m[p.global_ids[loc]] = np.arange(p.global_ids[loc].shape[0])
            d[loc] = m
        lookup.append(d)
    for p in parts:
        for loc in LOCS:
            halo_local = np.arange(p.n_owned[loc], p.mesh.num(loc))
            halo_global = p.global_ids[loc][halo_local]
            who = own[loc][halo_global]
            for q in np.unique(who):
                if q < 0:
                    continue            # nobody owns it (invalid edge slot): nothing to refresh
                sel = who == q
                p.recv.setdefault(loc, {})[int(q)] = halo_local[sel]
                src = lookup[int(q)][loc][halo_global[sel]]
                assert (src >= 0).all() and (src < parts[int(q)].n_owned[loc]).all()
                parts[int(q)].send.setdefault(loc, {})[p.rank] = src
    return parts
