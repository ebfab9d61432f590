"""Device-side mesh for generated b200-ico stencils: element counts + SoA neighbour tables
(b200_mesh_t of include/dawn_b200_rt.h; the reference's dawn::GlobalGpuTriMesh,
driver-includes/cuda_utils.hpp:43-57,212-263)."""
import ctypes

import numpy as np

from . import runtime as _rt


class NbhTable(ctypes.Structure):
    _fields_ = [("chain", ctypes.c_int * 8), ("chain_len", ctypes.c_int),
                ("include_center", ctypes.c_int), ("nbh_size", ctypes.c_int),
                ("table", ctypes.c_void_p)]


class Splitter(ctypes.Structure):
    _fields_ = [("location", ctypes.c_int), ("subdomain", ctypes.c_int), ("offset", ctypes.c_int),
                ("index", ctypes.c_int)]


class MeshStruct(ctypes.Structure):
    _fields_ = [("num_cells", ctypes.c_int), ("num_edges", ctypes.c_int),
                ("num_vertices", ctypes.c_int), ("num_tables", ctypes.c_int),
                ("tables", ctypes.POINTER(NbhTable)), ("num_splitters", ctypes.c_int),
                ("splitters", ctypes.POINTER(Splitter))]


class Mesh:
    """Uploads the neighbour tables a stencil module asks for (module.info["tables"]) from any mesh
    adapter offering num_cells/num_edges/num_vertices and nbh_table(chain, include_center, nbh_size)
    (e.g. dawn_b200.trimesh.TriMesh)."""

    def __init__(self, adapter, tables, device="cuda", splitters=None):
        import torch
        self.adapter = adapter
        self._dev = []
        entries = []
        for t in tables:
            chain, center, size = list(t["chain"]), int(t.get("include_center", 0)), int(t["size"])
            host = np.ascontiguousarray(adapter.nbh_table(chain, bool(center), nbh_size=size),
                                        dtype=np.int32)
            n_elem = host.shape[0]
            dev = torch.empty(size * n_elem, dtype=torch.int32, device=device)
            _rt.check(_rt.lib().b200rt_upload_nbh_table(host.ctypes.data, n_elem, size, dev.data_ptr(),
                                                        None))
            self._dev.append(dev)
            e = NbhTable()
            for n, c in enumerate(chain):
                e.chain[n] = c
            e.chain_len, e.include_center, e.nbh_size, e.table = len(chain), center, size, dev.data_ptr()
            entries.append(e)
        self._entries = (NbhTable * max(1, len(entries)))(*entries)
        spl = [Splitter(int(l), int(s), int(o), int(i)) for (l, s, o), i in (splitters or {}).items()]
        self._splitters = (Splitter * max(1, len(spl)))(*spl)
        self.struct = MeshStruct(adapter.num_cells, adapter.num_edges, adapter.num_vertices,
                                 len(entries), self._entries, len(spl), self._splitters)

    def soa_table(self, index):
        """device table `index` back on the host, shaped [nbh, element]"""
        e = self._entries[index]
        n = {0: self.adapter.num_cells, 1: self.adapter.num_edges, 2: self.adapter.num_vertices}[e.chain[0]]
        return self._dev[index].cpu().numpy().reshape(e.nbh_size, n)
