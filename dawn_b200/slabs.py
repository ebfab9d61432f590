"""j-slab decomposition of a Cartesian domain over the GPUs of one box, and the halo-exchange plan.

The reference has no communication layer; its only distribution hook is the
`(rank, xcols, ycols)` constructor argument of generated stencils with `computeGlobalOffsets`
(dawn/src/dawn/CodeGen/CodeGen.cpp:461-470).  This module is the host logic on top of that hook:
rank r of P owns the global rows [r*J/P, (r+1)*J/P) plus `halo` rows each side; before a stencil
runs, every *input* field with a non-zero accumulated j extent (`<field>_extent`, exported by the
generated module) receives `jminus`/`jplus` rows from its slab neighbours.

The data path on GPUs is b200rt_halo_exchange (pack kernel -> ncclSend/ncclRecv -> unpack kernel,
dawn_b200/csrc/runtime/rt.cu); `exchange_torch` below is the same plan over torch.distributed
point-to-point calls, used by the CPU (gloo) tests and usable with any backend.
"""
from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass
class Slab:
    rank: int
    nranks: int
    j0: int          # first global interior row owned
    nj_local: int    # interior rows owned
    halo: int

    @property
    def has_lower(self):
        return self.rank > 0

    @property
    def has_upper(self):
        return self.rank + 1 < self.nranks


def decompose(J: int, nranks: int, rank: int, halo: int = 3) -> Slab:
    """Balanced contiguous split of J interior rows."""
    base, rem = divmod(J, nranks)
    nj = base + (1 if rank < rem else 0)
    j0 = rank * base + min(rank, rem)
    return Slab(rank, nranks, j0, nj, halo)


def exchange_widths(fields: List[dict], written: set) -> Dict[str, Tuple[int, int]]:
    """{field: (rows needed below, rows needed above)} for the inputs of a generated module
    (module.info["fields"] entries carry extent = [i-, i+, j-, j+, k-, k+])."""
    out = {}
    for f in fields:
        jm, jp = -f["extent"][2], f["extent"][3]
        if f["name"] not in written and "j" in f["dims"] and "i" in f["dims"] and (jm > 0 or jp > 0):
            out[f["name"]] = (jm, jp)
    return out


def row_ranges(slab: Slab, nj_total: int, width_minus: int, width_plus: int):
    """Row index ranges (in the local storage incl. halos) of one exchange:
    send_down/send_up are interior rows handed to rank-1 / rank+1, recv_low/recv_high the halo rows
    they fill here.  Mirrors b200rt_halo_exchange."""
    h = slab.halo
    j_int0, j_int1 = h, nj_total - h
    return {
        "send_down": (j_int0, j_int0 + width_plus),        # -> rank-1's high halo
        "send_up": (j_int1 - width_minus, j_int1),         # -> rank+1's low halo
        "recv_low": (j_int0 - width_minus, j_int0),        # <- rank-1
        "recv_high": (j_int1, j_int1 + width_plus),        # <- rank+1
    }


def exchange_torch(t, slab: Slab, width_minus: int, width_plus: int, dist):
    """Halo exchange of a [k, j, i] tensor (local slab incl. halos) with torch.distributed P2P."""
    r = row_ranges(slab, t.shape[1], width_minus, width_plus)
    ops, recvs = [], []
    if slab.has_lower:
        if width_plus:
            ops.append(dist.P2POp(dist.isend, t[:, r["send_down"][0]:r["send_down"][1], :].contiguous(),
                                  slab.rank - 1))
        if width_minus:
            buf = t.new_empty((t.shape[0], width_minus, t.shape[2]))
            ops.append(dist.P2POp(dist.irecv, buf, slab.rank - 1))
            recvs.append((buf, r["recv_low"]))
    if slab.has_upper:
        if width_minus:
            ops.append(dist.P2POp(dist.isend, t[:, r["send_up"][0]:r["send_up"][1], :].contiguous(),
                                  slab.rank + 1))
        if width_plus:
            buf = t.new_empty((t.shape[0], width_plus, t.shape[2]))
            ops.append(dist.P2POp(dist.irecv, buf, slab.rank + 1))
            recvs.append((buf, r["recv_high"]))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, (a, b) in recvs:
        t[:, a:b, :] = buf
    return t


# ------------------------------------------------------------------------------------------------
# Unstructured: row slabs of a TriMesh
# ------------------------------------------------------------------------------------------------
@dataclass
class MeshSlab:
    """Rank-local piece of a global TriMesh(nx, ny): the owned quad rows [j0, j0+nrows) plus one halo
    quad row on every side that has a neighbour.  Element ids are row-major in j for all three
    location types, so a quad row of edge slots / cells / vertices is a contiguous id range and the
    Cartesian halo plan applies with rows = slot rows."""
    rank: int
    nranks: int
    nx: int
    j0: int        # first owned global quad row
    nrows: int     # owned quad rows
    halo_lo: int   # halo quad rows below (0 on the first rank)
    halo_hi: int   # halo quad rows above (0 on the last rank)

    @property
    def local_ny(self):
        return self.nrows + self.halo_lo + self.halo_hi

    @property
    def row_offset(self):
        """global quad row of local quad row 0"""
        return self.j0 - self.halo_lo

    def slots_per_row(self, loc):
        """elements per row of slots: 0 cells, 1 edges, 2 vertices"""
        return (2 * self.nx, 3 * (self.nx + 1), self.nx + 1)[loc]

    def owned_slice(self, loc):
        """local id range of the elements this rank owns (edges/vertices of the row above the last
        owned quad row belong to the next rank, except on the last rank)"""
        n = self.slots_per_row(loc)
        rows = self.nrows + (1 if (loc != 0 and self.halo_hi == 0) else 0)
        return slice(self.halo_lo * n, (self.halo_lo + rows) * n)

    def global_slice(self, loc):
        n = self.slots_per_row(loc)
        rows = self.nrows + (1 if (loc != 0 and self.halo_hi == 0) else 0)
        return slice(self.j0 * n, (self.j0 + rows) * n)


def mesh_slab(nx: int, ny: int, nranks: int, rank: int) -> MeshSlab:
    sl = decompose(ny, nranks, rank, 1)
    return MeshSlab(rank, nranks, nx, sl.j0, sl.nj_local, 1 if rank > 0 else 0,
                    1 if rank + 1 < nranks else 0)
