"""Mesh renumbering for the b200-ico backend: locality-preserving permutations of cells, edges and
vertices, recorded so that every result can be mapped back bit for bit.

The reference runs ICON's unstructured stencils on whatever numbering the host model hands it
(`GlobalGpuTriMesh`, driver-includes/cuda_utils.hpp:43-57); the order of the elements decides how many
32-byte sectors a warp of 32 consecutive elements touches when it gathers through a neighbour table.
Renumbering is index work only and must be exact:
  * a permutation is a bijection new id -> old id per location type (`Renumbering`);
  * `ReorderedMesh` serves the renumbered neighbour tables: rows are permuted like their source
    elements, entries are relabelled, -1 stays -1 and the ORDER of the neighbours inside a row is kept,
    so every reduction adds the same numbers in the same order and results are bit-identical after the
    inverse permutation;
  * `TableMesh` is the adapter for meshes that exist only as connectivity tables (what an ICON or atlas
    host can provide): the six two-element tables, from which longer chains are composed exactly like
    the reference's mesh interface does (`toylib_interface.hpp:95-274`).

Orderings: `rcm_cells` (reverse Cuthill-McKee over the cell-edge-cell graph: topology only) and
`morton_cells` (space-filling curve over cell centroids, when vertex coordinates exist); edges and
vertices follow the cells (`follow_cells`: sorted by the smallest new id among their adjacent cells),
invalid edge slots go last.
"""
import numpy as np

from .trimesh import CELL, EDGE, VERTEX, NAMES, chain_rows

LOCS = (CELL, EDGE, VERTEX)


class Renumbering:
    """new id -> old id for every location type (`new_to_old[loc]`), and its inverse."""

    def __init__(self, new_to_old):
        self.new_to_old, self.old_to_new = {}, {}
        for loc in LOCS:
            p = np.ascontiguousarray(new_to_old[loc], dtype=np.int64)
            n = p.shape[0]
            inv = np.full(n, -1, dtype=np.int64)
            inv[p] = np.arange(n)
            if p.ndim != 1 or (inv < 0).any():
                raise ValueError("renumbering of location %d is not a bijection" % loc)
            self.new_to_old[loc], self.old_to_new[loc] = p, inv

    @staticmethod
    def identity(mesh):
        return Renumbering({loc: np.arange(mesh.num(loc)) for loc in LOCS})

    @staticmethod
    def random(mesh, seed):
        rng = np.random.default_rng(seed)
        return Renumbering({loc: rng.permutation(mesh.num(loc)) for loc in LOCS})

    def to_new(self, loc, a, axis=-1):
        """field in the old numbering -> field in the new numbering (element axis `axis`)"""
        return np.take(a, self.new_to_old[NAMES[loc]], axis=axis)

    def to_old(self, loc, a, axis=-1):
        return np.take(a, self.old_to_new[NAMES[loc]], axis=axis)

    def then(self, later):
        """the renumbering that applies `self` first and `later` (defined on the result) second"""
        return Renumbering({loc: self.new_to_old[loc][later.new_to_old[loc]] for loc in LOCS})


class ReorderedMesh:
    """`base` seen through a `Renumbering`; same adapter interface as `TriMesh`."""

    def __init__(self, base, renumbering):
        self.base, self.renumbering = base, renumbering
        self.num_cells, self.num_edges, self.num_vertices = base.num_cells, base.num_edges, base.num_vertices
        for loc in LOCS:
            if renumbering.new_to_old[loc].shape[0] != base.num(loc):
                raise ValueError("renumbering does not fit the mesh")

    def num(self, loc):
        return self.base.num(loc)

    def valid(self, loc):
        return self.base.valid(loc)[self.renumbering.new_to_old[NAMES[loc]]]

    @property
    def edge_valid(self):
        return self.valid(EDGE)

    def nbh_table(self, chain, include_center=False, nbh_size=None):
        chain = [NAMES[c] for c in chain]
        t = self.base.nbh_table(chain, include_center, nbh_size)
        t = t[self.renumbering.new_to_old[chain[0]]]
        relabel = self.renumbering.old_to_new[chain[-1]]
        return np.where(t >= 0, relabel[np.maximum(t, 0)], -1).astype(np.int32)

    def centroids(self):
        c = getattr(self.base, "centroids", None)
        if c is None:
            return None
        x, y = c()
        p = self.renumbering.new_to_old[CELL]
        return x[p], y[p]


class TableMesh:
    """A mesh given by its two-element connectivity tables (row-major int32, -1 = missing):
    base[(CELL, EDGE)], (CELL, VERTEX), (EDGE, CELL), (EDGE, VERTEX), (VERTEX, CELL), (VERTEX, EDGE).
    Longer chains are composed from them in the reference's neighbour order."""

    def __init__(self, num_cells, num_edges, num_vertices, base, edge_valid=None, cell_xy=None):
        self.num_cells, self.num_edges, self.num_vertices = int(num_cells), int(num_edges), int(num_vertices)
        self._base = {(NAMES[a], NAMES[b]): np.ascontiguousarray(t, dtype=np.int32) for (a, b), t in base.items()}
        for (a, b), t in self._base.items():
            if t.shape[0] != self.num(a) or (t >= self.num(b)).any() or (t < -1).any():
                raise ValueError("table (%d, %d) does not fit the element counts" % (a, b))
        self.edge_valid = (np.ones(self.num_edges, dtype=bool) if edge_valid is None
                           else np.asarray(edge_valid, dtype=bool))
        self._cell_xy = cell_xy

    @staticmethod
    def from_mesh(mesh):
        """the two-element tables of any adapter, e.g. to renumber once and keep plain tables"""
        pairs = [(CELL, EDGE), (CELL, VERTEX), (EDGE, CELL), (EDGE, VERTEX), (VERTEX, CELL), (VERTEX, EDGE)]
        c = getattr(mesh, "centroids", None)
        return TableMesh(mesh.num_cells, mesh.num_edges, mesh.num_vertices,
                         {p: mesh.nbh_table(list(p)) for p in pairs}, mesh.valid(EDGE),
                         None if c is None else c())

    def num(self, loc):
        return (self.num_cells, self.num_edges, self.num_vertices)[NAMES[loc]]

    def valid(self, loc):
        return self.edge_valid if NAMES[loc] == EDGE else np.ones(self.num(loc), dtype=bool)

    def centroids(self):
        return self._cell_xy

    def nbh_table(self, chain, include_center=False, nbh_size=None):
        chain = [NAMES[c] for c in chain]
        if len(chain) < 2:
            raise ValueError("neighbour chain needs at least two locations")
        if len(chain) == 2 and not include_center:
            t = self._base[(chain[0], chain[1])]
            if chain[0] == EDGE:
                t = np.where(self.edge_valid[:, None], t, -1).astype(np.int32)
            rows = None
        else:
            rows = chain_rows(self._base, chain, include_center, self.num(chain[0]), self.valid(chain[0]))
            width = max((len(r) for r in rows), default=0)
            t = np.full((len(rows), width), -1, dtype=np.int32)
            for n, r in enumerate(rows):
                t[n, :len(r)] = r
        if nbh_size is not None and nbh_size != t.shape[1]:
            if nbh_size < t.shape[1] and (t[:, nbh_size:] >= 0).any():
                raise ValueError("neighbour table is wider than nbh_size")
            out = np.full((t.shape[0], nbh_size), -1, dtype=np.int32)
            w = min(nbh_size, t.shape[1])
            out[:, :w] = t[:, :w]
            t = out
        return np.ascontiguousarray(t)


# ---- orderings ---------------------------------------------------------------------------------------
def rcm_cells(mesh):
    """new -> old cell ids by reverse Cuthill-McKee over the cell-edge-cell adjacency (topology only)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    e2c = mesh.nbh_table([EDGE, CELL])          # two cells share an edge: that is the whole graph
    both = (e2c[:, 0] >= 0) & (e2c[:, 1] >= 0)
    a, b = e2c[both, 0].astype(np.int64), e2c[both, 1].astype(np.int64)
    n = mesh.num_cells
    g = coo_matrix((np.ones(2 * a.shape[0], dtype=np.int8), (np.concatenate([a, b]), np.concatenate([b, a]))),
                   shape=(n, n)).tocsr()
    return np.asarray(reverse_cuthill_mckee(g, symmetric_mode=True), dtype=np.int64)


def _interleave16(v):
    v = v.astype(np.uint64) & np.uint64(0xFFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
    return v


def morton_cells(mesh):
    """new -> old cell ids along a Z-order curve over the cell centroids (16 bits per axis); ties are
    broken by the old id, so the order is deterministic."""
    c = mesh.centroids() if hasattr(mesh, "centroids") else None
    if c is None:
        raise ValueError("morton_cells needs cell centroids (mesh.centroids())")
    x, y = (np.asarray(v, dtype=float) for v in c)

    def quant(v):
        lo, hi = v.min(), v.max()
        return np.zeros(v.shape, dtype=np.uint64) if hi == lo else np.minimum(
            ((v - lo) / (hi - lo) * 65535.0), 65535.0).astype(np.uint64)
    code = _interleave16(quant(x)) | (_interleave16(quant(y)) << np.uint64(1))
    return np.lexsort((np.arange(x.shape[0]), code)).astype(np.int64)


def follow_cells(mesh, cells_new_to_old):
    """Complete renumbering from a cell order: an edge / vertex is placed by the smallest NEW id among
    its adjacent cells (ties: old id); elements without cells (invalid edge slots) go last."""
    n = mesh.num_cells
    inv = np.empty(n, dtype=np.int64)
    inv[np.asarray(cells_new_to_old, dtype=np.int64)] = np.arange(n)
    out = {CELL: np.asarray(cells_new_to_old, dtype=np.int64)}
    for loc in (EDGE, VERTEX):
        t = mesh.nbh_table([loc, CELL])
        key = np.where(t >= 0, inv[np.maximum(t, 0)], n).min(axis=1)
        out[loc] = np.lexsort((np.arange(t.shape[0]), key)).astype(np.int64)
    return Renumbering(out)


def reorder(mesh, method="rcm"):
    """`ReorderedMesh` of `mesh` in a locality-preserving numbering ("rcm" or "morton")."""
    cells = {"rcm": rcm_cells, "morton": morton_cells}[method](mesh)
    return ReorderedMesh(mesh, follow_cells(mesh, cells))


# ---- what the numbering costs a gather ------------------------------------------------------------------
def sectors_per_warp(mesh, chain, include_center=False, elem_bytes=8):
    """Mean number of distinct 32-byte sectors one warp (32 consecutive source elements) touches per
    neighbour slot when it gathers a dense field of `elem_bytes` elements through the chain's table:
    32*elem_bytes/32 (= 8 for doubles) when the neighbours of consecutive elements are consecutive,
    32 when they are scattered."""
    t = mesh.nbh_table(chain, include_center)
    per = 32 // elem_bytes
    n = (t.shape[0] // 32) * 32
    if n == 0 or t.shape[1] == 0:
        return 0.0
    total, slots = 0, 0
    for c in range(t.shape[1]):
        col = t[:n, c].reshape(-1, 32)
        sec = np.where(col >= 0, col // per, -1)
        sec.sort(axis=1)
        distinct = (np.diff(sec, axis=1) != 0).sum(axis=1) + 1 - (sec[:, 0] < 0)  # -1 is not a sector
        has = (col >= 0).any(axis=1)
        total += int(distinct[has].sum())
        slots += int(has.sum())
    return total / max(slots, 1)
