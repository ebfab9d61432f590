"""Structured triangle mesh with toylib-compatible numbering, and neighbour tables for the b200-ico
backend.

The reference's generated unstructured code is mesh-library agnostic: it asks an adapter for
`getNeighbors(mesh, chain, element)` and (cuda-ico) flattens the answers into int32 tables
(`dawn/src/driver-includes/cuda_utils.hpp:212-263`).  This module is that adapter for the regular
triangulated quad mesh of the reference's own toy library (element numbering, neighbour ORDER and the
edge-slot id space of `dawn/src/toylib/toylib.hpp:129-359`), built with vectorised index arithmetic
instead of pointer-linked objects so that ~20 M-edge meshes are generated in seconds.

Numbering (non-periodic nx x ny quads, each split by the diagonal (i,j)-(i+1,j+1)):
  vertices  id = j*(nx+1) + i                         (nx+1)*(ny+1)
  cells     id = 2*(j*nx + i) + c    c=0 upward, 1 downward        2*nx*ny
  edges     id = 3*(j*(nx+1) + i) + c   c=0 horizontal, 1 diagonal, 2 vertical
            (slot id space of size 3*(nx+1)*(ny+1); slots that do not exist on the boundary are
             invalid: their table rows are all -1 and `edge_valid` is False)
Missing neighbours are -1 (DEVICE_MISSING_VALUE of the reference).  Tables are row-major
[element, neighbour]; `dawn_b200.mesh.Mesh` uploads them transposed (SoA) to the device.
"""
import numpy as np

CELL, EDGE, VERTEX = 0, 1, 2  # ::dawn::LocationType (driver-includes/unstructured_interface.hpp)
NAMES = {"Cell": CELL, "Edge": EDGE, "Vertex": VERTEX, "Cells": CELL, "Edges": EDGE,
         "Vertices": VERTEX, CELL: CELL, EDGE: EDGE, VERTEX: VERTEX}


class TriMesh:
    def __init__(self, nx, ny, lx=0.0, ly=0.0, equilat=False):
        self.nx, self.ny = int(nx), int(ny)
        nx, ny = self.nx, self.ny
        self.num_vertices = (nx + 1) * (ny + 1)
        self.num_cells = 2 * nx * ny
        self.num_edges = 3 * (nx + 1) * (ny + 1)  # slot id space
        self._base = {}
        self._build(lx, ly, equilat)

    # ---- id helpers (arrays) -------------------------------------------------------------------
    def _vid(self, i, j):
        return j * (self.nx + 1) + i

    def _cid(self, i, j, c):
        return 2 * (j * self.nx + i) + c

    def _eid(self, i, j, c):
        return 3 * (j * (self.nx + 1) + i) + c

    def _build(self, lx, ly, equilat):
        nx, ny = self.nx, self.ny
        # vertex coordinates (toylib.hpp:166-183)
        jj, ii = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
        px = ii.astype(float) if lx == 0.0 else ii / nx * lx
        py = jj.astype(float) if ly == 0.0 else jj / ny * ly
        if equilat:
            px = px - 0.5 * py
            py = py * np.sqrt(3) / 2.0
        self.vertex_x, self.vertex_y = px.reshape(-1), py.reshape(-1)

        # ---- edge validity ------------------------------------------------------------------
        valid = np.zeros((ny + 1, nx + 1, 3), dtype=bool)
        valid[:, :nx, 0] = True       # horizontal: i < nx, all j
        valid[:ny, :nx, 1] = True     # diagonal
        valid[:ny, :, 2] = True       # vertical: j < ny, all i
        self.edge_valid = valid.reshape(-1)

        # ---- cell -> edge / vertex (toylib.hpp:185-206) ------------------------------------
        cj, ci = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        ci, cj = ci.reshape(-1), cj.reshape(-1)
        c2e = np.empty((self.num_cells, 3), dtype=np.int32)
        c2v = np.empty((self.num_cells, 3), dtype=np.int32)
        up, dn = self._cid(ci, cj, 0), self._cid(ci, cj, 1)
        c2e[up] = np.stack([self._eid(ci, cj, 2), self._eid(ci, cj, 1), self._eid(ci, cj + 1, 0)], 1)
        c2v[up] = np.stack([self._vid(ci, cj), self._vid(ci + 1, cj + 1), self._vid(ci, cj + 1)], 1)
        c2e[dn] = np.stack([self._eid(ci, cj, 1), self._eid(ci, cj, 0), self._eid(ci + 1, cj, 2)], 1)
        c2v[dn] = np.stack([self._vid(ci, cj), self._vid(ci + 1, cj), self._vid(ci + 1, cj + 1)], 1)

        # ---- edge -> vertex / cell (toylib.hpp:208-244, diagonal faces swapped at :296-300) --
        e2v = np.full((self.num_edges, 2), -1, dtype=np.int32)
        e2c = np.full((self.num_edges, 2), -1, dtype=np.int32)
        ej, ei = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
        ei, ej = ei.reshape(-1), ej.reshape(-1)
        # horizontal
        m = ei < nx
        i, j = ei[m], ej[m]
        e = self._eid(i, j, 0)
        e2v[e] = np.stack([self._vid(i, j), self._vid(i + 1, j)], 1)
        below = np.where(j > 0, self._cid(i, np.maximum(j - 1, 0), 0), -1)     # upward face of row j-1
        above = np.where(j < ny, self._cid(i, np.minimum(j, ny - 1), 1), -1)   # downward face of row j
        first = np.where(below >= 0, below, above)
        second = np.where(below >= 0, above, -1)
        e2c[e] = np.stack([first, second], 1)
        # diagonal (faces added downward, upward and then swapped -> upward, downward)
        m = (ei < nx) & (ej < ny)
        i, j = ei[m], ej[m]
        e = self._eid(i, j, 1)
        e2v[e] = np.stack([self._vid(i, j), self._vid(i + 1, j + 1)], 1)
        e2c[e] = np.stack([self._cid(i, j, 0), self._cid(i, j, 1)], 1)
        # vertical
        m = ej < ny
        i, j = ei[m], ej[m]
        e = self._eid(i, j, 2)
        e2v[e] = np.stack([self._vid(i, j), self._vid(i, j + 1)], 1)
        right = np.where(i < nx, self._cid(np.minimum(i, nx - 1), j, 0), -1)   # upward face of column i
        left = np.where(i > 0, self._cid(np.maximum(i - 1, 0), j, 1), -1)      # downward face of col i-1
        first = np.where(right >= 0, right, left)
        second = np.where(right >= 0, left, -1)
        e2c[e] = np.stack([first, second], 1)

        # ---- vertex -> edge / cell (toylib.hpp:246-287), conditional appends => compact rows --
        vj, vi = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
        vi, vj = vi.reshape(-1), vj.reshape(-1)
        im, jm = np.maximum(vi - 1, 0), np.maximum(vj - 1, 0)
        cand_e = np.stack([
            np.where(vi > 0, self._eid(im, vj, 0), -1),
            np.where((vi > 0) & (vj > 0), self._eid(im, jm, 1), -1),
            np.where(vj > 0, self._eid(vi, jm, 2), -1),
            np.where(vi < nx, self._eid(vi, vj, 0), -1),
            np.where((vi < nx) & (vj < ny), self._eid(vi, vj, 1), -1),
            np.where(vj < ny, self._eid(vi, vj, 2), -1)], 1).astype(np.int32)
        ic, jc = np.minimum(vi, nx - 1), np.minimum(vj, ny - 1)
        a = (vi > 0) & (vj > 0)
        b = (vi < nx) & (vj > 0)
        c = (vi < nx) & (vj < ny)
        d = (vi > 0) & (vj < ny)
        cand_c = np.stack([
            np.where(a, self._cid(im, jm, 0), -1),
            np.where(a, self._cid(im, jm, 1), -1),
            np.where(b, self._cid(ic, jm, 0), -1),
            np.where(c, self._cid(ic, jc, 1), -1),
            np.where(c, self._cid(ic, jc, 0), -1),
            np.where(d, self._cid(im, jc, 1), -1)], 1).astype(np.int32)
        self._base = {
            (CELL, EDGE): c2e, (CELL, VERTEX): c2v, (EDGE, VERTEX): e2v, (EDGE, CELL): e2c,
            (VERTEX, EDGE): _compact(cand_e), (VERTEX, CELL): _compact(cand_c),
        }

    def num(self, loc):
        return (self.num_cells, self.num_edges, self.num_vertices)[NAMES[loc]]

    def valid(self, loc):
        if NAMES[loc] == EDGE:
            return self.edge_valid
        return np.ones(self.num(loc), dtype=bool)

    # ---- neighbour tables ----------------------------------------------------------------------
    def nbh_table(self, chain, include_center=False, nbh_size=None):
        """Row-major int32 table [num(chain[0]), nbh_size], -1 padded, neighbours in the order the
        reference's toylib adapter returns them (`toylib_interface.hpp:95-274`): walking the chain
        front by front, every element of the TARGET type adjacent to the current front is collected,
        then duplicates (and the origin, if the chain starts and ends on the same type) are dropped
        keeping first occurrences."""
        chain = [NAMES[c] for c in chain]
        if len(chain) < 2:
            raise ValueError("neighbour chain needs at least two locations")
        if len(chain) == 2 and not include_center:
            t = self._base[(chain[0], chain[1])]
            if chain[0] == EDGE:
                t = np.where(self.edge_valid[:, None], t, -1).astype(np.int32)
            if nbh_size is not None and nbh_size != t.shape[1]:
                out = np.full((t.shape[0], nbh_size), -1, dtype=np.int32)
                out[:, :t.shape[1]] = t
                t = out
            return np.ascontiguousarray(t)
        rows = self._chain_rows(chain, include_center)
        width = nbh_size if nbh_size is not None else max((len(r) for r in rows), default=0)
        out = np.full((len(rows), width), -1, dtype=np.int32)
        for n, r in enumerate(rows):
            out[n, :len(r)] = r
        return out

    def _chain_rows(self, chain, include_center):
        return chain_rows(self._base, chain, include_center, self.num(chain[0]), self.valid(chain[0]))

    def centroids(self):
        """cell centroids (x, y) from the vertex coordinates"""
        c2v = self._base[(CELL, VERTEX)]
        return self.vertex_x[c2v].mean(axis=1), self.vertex_y[c2v].mean(axis=1)


def chain_rows(base, chain, include_center, n0, valid0):
    """Neighbour lists of every element of chain[0] over `chain`, composed from the two-element tables
    `base[(from, to)]` in the reference's order (`toylib_interface.hpp:95-274`): walking the chain front
    by front, every element of the TARGET type adjacent to the current front is collected, then
    duplicates (and the origin, if the chain starts and ends on the same type) are dropped keeping
    first occurrences."""
    target = chain[-1]
    same = chain[0] == chain[-1]
    rows = []
    for elem in range(n0):
        if not valid0[elem]:
            rows.append([])
            continue
        front = [elem]
        result = []
        for hop in range(len(chain) - 1):
            frm, to = chain[hop], chain[hop + 1]
            new_front = []
            for idx in front:
                new_front.extend(int(x) for x in base[(frm, to)][idx] if x >= 0)
                if (frm, target) in base and frm != target:
                    result.extend(int(x) for x in base[(frm, target)][idx] if x >= 0)
            front = new_front
        seen = set()
        uniq = []
        for x in result:
            if same and x == elem:
                continue
            if x not in seen:
                seen.add(x)
                uniq.append(x)
        if include_center:
            uniq = [elem] + uniq
        rows.append(uniq)
    return rows


def _compact(cand):
    """Moves the valid (>= 0) entries of each row to the front, keeping their order."""
    order = np.argsort(cand < 0, axis=1, kind="stable")
    return np.take_along_axis(cand, order, axis=1).astype(np.int32)
