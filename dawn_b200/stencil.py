"""Host-side mirror of the reference's generated-stencil interface for Python drivers.

The reference's user code is C++ (`domain`, storages, `dawn_generated::<ns>::<S> s(dom); s.run(...)`,
gtclang/test/integration-test/CodeGen/hori_diff_type2_stencil_benchmark.cpp:44-73).  The same objects
exist here, over the C ABI every generated module exports (include/dawn_b200_stencil.h):
`Domain`, `Storage` (HBM-resident, reference layout rules of driver-includes/storage.hpp in this
repo), and `Stencil(domain).run(...)`.  torch provides device memory and streams only.
"""
import ctypes
import json
import os

import numpy as np

from . import build as _build
from . import codegen as _codegen
from . import runtime as _rt

HALO = 3  # GRIDTOOLS_DAWN_HALO_EXTENT


class Domain:
    """gridtools::dawn::domain (sizes include halos; driver-includes/domain.hpp)."""

    def __init__(self, isize, jsize, ksize):
        self.dims = (int(isize), int(jsize), int(ksize))
        self.halos = [HALO, HALO, HALO, HALO, 0, 0]

    def set_halos(self, iminus=0, iplus=0, jminus=0, jplus=0, kminus=0, kplus=0):
        self.halos = [iminus, iplus, jminus, jplus, kminus, kplus]

    isize = property(lambda s: s.dims[0])
    jsize = property(lambda s: s.dims[1])
    ksize = property(lambda s: s.dims[2])

    def interior(self):
        h = self.halos
        return (self.dims[0] - h[0] - h[1], self.dims[1] - h[2] - h[3], self.dims[2] - h[4] - h[5])


def _round_up(x, m):
    return (x + m - 1) // m * m


class Storage:
    """Device-resident storage with the layout of driver-includes/storage.hpp: i fastest, rows padded
    to 16 doubles, first interior point of each row 128-byte aligned.  `dims` in {"ijk","ij","i","j",
    "k"}.  Backed by a torch CUDA tensor; dtype "float64" (::dawn::float_type = double) or "float32" (modules
    generated with single_precision=1): rows are padded to 128 bytes either way."""

    def __init__(self, ni, nj, nk, dims="ijk", device="cuda", halo=HALO, dtype="float64"):
        import torch
        self.dims = dims
        self.dtype = getattr(torch, dtype) if isinstance(dtype, str) else dtype
        self.itemsize = 4 if self.dtype == torch.float32 else 8
        line = 128 // self.itemsize   # elements per 128-byte line
        mi, mj, mk = "i" in dims, "j" in dims, "k" in dims
        self.shape = (ni if mi else 1, nj if mj else 1, nk if mk else 1)
        s = 1
        self.lead = 0
        self.stride_i = self.stride_j = self.stride_k = 0
        if mi:
            self.stride_i = 1
            self.lead = (line - halo % line) % line
            s = _round_up(self.shape[0], line)
        if mj:
            self.stride_j = s
            s *= self.shape[1]
        if mk:
            self.stride_k = s
            s *= self.shape[2]
        self.elems = s + self.lead + line
        self.buf = torch.zeros(self.elems, dtype=self.dtype, device=device)

    # -- views --------------------------------------------------------------------------------
    def _strided(self, t):
        import torch
        ni, nj, nk = self.shape
        return torch.as_strided(t, (nk, nj, ni), (max(self.stride_k, 0), max(self.stride_j, 0),
                                                  max(self.stride_i, 0)), self.lead)

    def view(self):
        """torch view indexed [k, j, i] over the logical (unpadded) extent."""
        return self._strided(self.buf)

    def from_numpy(self, arr):
        """arr indexed [k, j, i] (length 1 along masked dimensions)."""
        import torch
        arr = np.ascontiguousarray(arr, dtype=np.float32 if self.itemsize == 4 else np.float64).reshape(
            self.shape[2], self.shape[1], self.shape[0])
        self.view().copy_(torch.from_numpy(arr))
        return self

    def to_numpy(self):
        return self.view().cpu().numpy().copy()

    def fill(self, v):
        self.buf.fill_(v)
        return self

    def descriptor(self):
        return _rt.Field(self.buf.data_ptr() + self.itemsize * self.lead, self.stride_j, self.stride_k)


def dense_descriptor(ptr, ni, nj, dims="ijk"):
    """Descriptor of a dense [k][j][i] device buffer (no padding)."""
    mi, mj = "i" in dims, "j" in dims
    sj = ni if mi else 1
    sk = (ni if mi else 1) * (nj if mj else 1)
    return _rt.Field(ptr, sj if mj else 0, sk if "k" in dims else 0)


class StencilModule:
    """A compiled generated translation unit (one .so per stencil instantiation)."""

    def __init__(self, name, so_path):
        if not os.path.exists(so_path):
            raise FileNotFoundError(
                "generated stencil module %s not built (run __graft_entry__.build()); there is no "
                "CPU fallback" % so_path)
        _rt.lib()  # the module links against the runtime
        self.name = name
        self.path = so_path
        self.lib = ctypes.CDLL(so_path)
        info = getattr(self.lib, "b200_module_info_" + name)
        info.restype = ctypes.c_char_p
        self.info = json.loads(info().decode())
        self.fields = self.info["fields"]
        self.unstructured = self.info.get("grid") == "Unstructured"
        # ::dawn::float_type of the translation unit (codegen option single_precision)
        self.dtype = "float32" if self.info.get("float_type") == "float" else "float64"
        self._setup = getattr(self.lib, "setup_" + name)
        if self.unstructured:
            self._setup.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p]
        else:
            self._setup.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
        self._run = getattr(self.lib, "run_" + name)
        self._run.argtypes = [ctypes.c_void_p, ctypes.POINTER(_rt.Field), ctypes.c_int,
                              ctypes.c_void_p]
        if not self.unstructured:
            self._run_rows = getattr(self.lib, "run_rows_" + name)
            self._run_rows.argtypes = [ctypes.c_void_p, ctypes.POINTER(_rt.Field), ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        self._free = getattr(self.lib, "free_" + name)
        self._free.argtypes = [ctypes.c_void_p]
        self._launches = getattr(self.lib, "launches_" + name)
        self._launches.argtypes = [ctypes.c_void_p]
        self._launches.restype = ctypes.c_long
        if not self.unstructured:
            self._last_launch = getattr(self.lib, "last_launch_" + name)
            self._last_launch.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
            self._last_launch.restype = None

    def __call__(self, domain, stream=None):
        return Stencil(self, domain, stream)

    def on_mesh(self, adapter, k_size, stream=None, splitters=None):
        """Unstructured modules: bind to a mesh adapter (e.g. trimesh.TriMesh) and a level count,
        like the reference's setup_<N>(mesh, k_size, stream) (CudaIcoCodeGen.cpp:1198-1219).
        splitters: {(location, subdomain, offset): index} for unstructured iteration spaces."""
        return UnstructuredStencil(self, adapter, k_size, stream, splitters)


class Stencil:
    """Instance of a generated stencil bound to a domain: constructor from domain, fields bound at
    run(), like `dawn_generated::cuda::<S>` (CudaCodeGen.cpp:144-174,412-452)."""

    def __init__(self, module, domain, stream=None):
        self.module = module
        self.domain = domain
        self.handle = ctypes.c_void_p()
        halos = (ctypes.c_int * 6)(*domain.halos)
        _rt.check(module._setup(ctypes.byref(self.handle), domain.isize, domain.jsize, domain.ksize,
                                halos, stream))
        self._host_storages = None

    def set_global(self, name, value):
        fn = getattr(self.module.lib, "set_%s_%s" % (self.module.name, name))
        fn.argtypes = [ctypes.c_void_p, ctypes.c_double]
        fn(self.handle, float(value))

    def get_global(self, name):
        fn = getattr(self.module.lib, "get_%s_%s" % (self.module.name, name))
        fn.argtypes = [ctypes.c_void_p]
        fn.restype = ctypes.c_double
        return fn(self.handle)

    def run(self, *fields, stream=None, rows=None):
        """fields: Storage objects or runtime.Field descriptors in API order (device-resident).
        rows=(j_begin, j_end) restricts the computation to those interior rows."""
        n = len(self.module.fields)
        if len(fields) != n:
            raise TypeError("%s.run expects %d fields (%s)" % (
                self.module.name, n, ", ".join(f["name"] for f in self.module.fields)))
        arr = (_rt.Field * n)(*[f.descriptor() if isinstance(f, Storage) else f for f in fields])
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        if rows is None:
            _rt.check(self.module._run(self.handle, arr, n, ctypes.c_void_p(stream)))
        else:
            _rt.check(self.module._run_rows(self.handle, arr, n, ctypes.c_void_p(stream),
                                            int(rows[0]), int(rows[1])))

    def run_from_host(self, *arrays, stream=None, upload_outputs=True, k_chunk=0):
        """End-to-end call with HOST buffers in API order, indexed [k, j, i] (length 1 along masked
        dimensions): numpy arrays, or torch CPU tensors (pinned ones are copied from directly).
        Copies the inputs host->device, runs, and copies every field the stencil writes back into
        the given buffers.  Returns (h2d_bytes, d2h_bytes).

        upload_outputs=False skips the upload of fields the stencil only writes (their halo on the
        device is whatever an earlier call uploaded).  k_chunk > 0 streams a level-independent stencil
        (module.info["k_parallel"]) through the device in chunks of k_chunk levels on three streams,
        so that H2D copies, kernels and D2H copies of different chunks overlap (PCIe is full duplex)."""
        import torch
        dom = self.domain
        if self._host_storages is None:
            self._host_storages = [Storage(dom.isize, dom.jsize, a.shape[0], f["dims"], dtype=self.module.dtype)
                                   for f, a in zip(self.module.fields, arrays)]
            self._pinned = [None] * len(arrays)
            self._uploaded_once = False
        written = self.written_fields()
        staged = []
        for n, a in enumerate(arrays):
            if isinstance(a, torch.Tensor) and a.is_pinned():
                pin = a
            else:
                if self._pinned[n] is None:
                    self._pinned[n] = torch.empty(tuple(a.shape), dtype=getattr(torch, self.module.dtype)).pin_memory()
                pin = self._pinned[n]
                pin.copy_(a if isinstance(a, torch.Tensor) else torch.from_numpy(a))
            staged.append(pin)

        def needs_upload(f):
            if f["name"] not in written:
                return True
            return upload_outputs or not self._uploaded_once or not f.get("write_only", False)

        if k_chunk and self.module.info.get("k_parallel") and all(
                f["dims"] in ("ijk", "ij", "i", "j") for f in self.module.fields):
            h2d, d2h = self._run_chunked(staged, k_chunk, needs_upload, written)
        else:
            h2d = d2h = 0
            for f, st, pin in zip(self.module.fields, self._host_storages, staged):
                if needs_upload(f):
                    st.view().copy_(pin.reshape(st.shape[2], st.shape[1], st.shape[0]),
                                    non_blocking=True)
                    h2d += pin.numel() * pin.element_size()
            self.run(*self._host_storages, stream=stream)
            for f, st, pin in zip(self.module.fields, self._host_storages, staged):
                if f["name"] in written:
                    pin.reshape(st.shape[2], st.shape[1], st.shape[0]).copy_(st.view(),
                                                                             non_blocking=True)
                    d2h += pin.numel() * pin.element_size()
            torch.cuda.current_stream().synchronize()
        self._uploaded_once = True
        for f, a, pin in zip(self.module.fields, arrays, staged):
            if f["name"] in written and pin is not a:
                if isinstance(a, torch.Tensor):
                    a.copy_(pin)
                else:
                    np.copyto(a, pin.numpy())
        return h2d, d2h

    def _run_chunked(self, staged, k_chunk, needs_upload, written):
        """H2D / compute / D2H software pipeline over chunks of levels."""
        import torch
        dom = self.domain
        nz = dom.ksize
        if not hasattr(self, "_chunk_streams"):
            self._chunk_streams = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
            self._chunk_instances = {}
        s_up, s_run, s_down = self._chunk_streams
        main = torch.cuda.current_stream()
        for s in self._chunk_streams:
            s.wait_stream(main)
        h2d = d2h = 0
        fields = self.module.fields
        # fields without a k dimension go up once
        with torch.cuda.stream(s_up):
            for f, st, pin in zip(fields, self._host_storages, staged):
                if "k" not in f["dims"] and needs_upload(f):
                    st.view().copy_(pin.reshape(st.shape[2], st.shape[1], st.shape[0]),
                                    non_blocking=True)
                    h2d += pin.numel() * pin.element_size()
        for k0 in range(0, nz, k_chunk):
            kc = min(k_chunk, nz - k0)
            ev_up, ev_run = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(s_up):
                for f, st, pin in zip(fields, self._host_storages, staged):
                    if "k" in f["dims"] and needs_upload(f):
                        st.view()[k0:k0 + kc].copy_(pin[k0:k0 + kc], non_blocking=True)
                        h2d += pin[k0:k0 + kc].numel() * pin.element_size()
                ev_up.record(s_up)
            inst = self._chunk_instances.get(kc)
            if inst is None:
                d = Domain(dom.isize, dom.jsize, kc)
                d.halos = list(dom.halos)
                inst = self._chunk_instances[kc] = Stencil(self.module, d)
            descs = []
            for f, st in zip(fields, self._host_storages):
                fd = st.descriptor()
                if "k" in f["dims"]:
                    fd = _rt.Field(fd.data + st.itemsize * k0 * st.stride_k, fd.stride_j, fd.stride_k)
                descs.append(fd)
            s_run.wait_event(ev_up)
            inst.run(*descs, stream=s_run.cuda_stream)
            ev_run.record(s_run)
            s_down.wait_event(ev_run)
            with torch.cuda.stream(s_down):
                for f, st, pin in zip(fields, self._host_storages, staged):
                    if f["name"] in written and "k" in f["dims"]:
                        pin[k0:k0 + kc].copy_(st.view()[k0:k0 + kc], non_blocking=True)
                        d2h += pin[k0:k0 + kc].numel() * pin.element_size()
        for s in self._chunk_streams:
            main.wait_stream(s)
        main.synchronize()
        return h2d, d2h

    def written_fields(self):
        return set(self.module.info.get("written", [f["name"] for f in self.module.fields]))

    def launches(self):
        return int(self.module._launches(self.handle))

    def last_launch(self):
        """Shape of the most recent kernel launch (last_launch_<N> of the C ABI)."""
        out = (ctypes.c_int * 5)()
        self.module._last_launch(self.handle, out)
        return {"grid": (out[0], out[1], out[2]), "k_chunk": out[3],
                "kernel": ("plain", "tma", "ring", "column_cache")[out[4]]}

    def close(self):
        if self.handle:
            self.module._free(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class UnstructuredStencil:
    """Instance of a generated b200-ico stencil on a mesh.  Fields are torch CUDA tensors (float64):
    dense [k, element], sparse [k, nbh, element], in the module's API order."""

    def __init__(self, module, adapter, k_size, stream=None, splitters=None):
        from .mesh import Mesh
        self.module = module
        self.mesh = Mesh(adapter, module.info["tables"], splitters=splitters)
        self.k_size = int(k_size)
        self.handle = ctypes.c_void_p()
        _rt.check(module._setup(ctypes.byref(self.handle), ctypes.addressof(self.mesh.struct),
                                self.k_size, stream))

    def descriptors(self, tensors):
        out = []
        for f, t in zip(self.module.fields, tensors):
            has_k = bool(f.get("k", 1))
            if f["location"] == "none":      # vertical [k]
                out.append(_rt.Field(t.data_ptr(), 0, 1))
            elif f["sparse"]:                # [k, nbh, n] or horizontal-only [nbh, n]
                n = t.shape[-1]
                out.append(_rt.Field(t.data_ptr(), n, t.shape[-2] * n if has_k else 0))
            else:                            # [k, n] or horizontal-only [n]
                out.append(_rt.Field(t.data_ptr(), 0, t.shape[-1]))
        return out

    def set_splitter_index(self, location, subdomain, offset, index):
        """dawn::GlobalGpuTriMesh::set_splitter_index (driver-includes/cuda_utils.hpp:53-56)."""
        fn = getattr(self.module.lib, "set_splitter_index_" + self.module.name)
        fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4
        fn.restype = None
        fn(self.handle, int(location), int(subdomain), int(offset), int(index))

    def run(self, *tensors, stream=None):
        n = len(self.module.fields)
        if len(tensors) != n:
            raise TypeError("%s.run expects %d fields (%s)" % (
                self.module.name, n, ", ".join(f["name"] for f in self.module.fields)))
        arr = (_rt.Field * n)(*self.descriptors(tensors))
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        _rt.check(self.module._run(self.handle, arr, n, ctypes.c_void_p(stream)))

    def launches(self):
        return int(self.module._launches(self.handle))

    def close(self):
        if self.handle:
            self.module._free(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_stencil(name, iir_path=None, tag="", **options) -> StencilModule:
    """Loads the module generated from stencils/<name>.iir; with codegen options (or a tag) a variant
    is emitted and compiled in-tree first (needs nvcc)."""
    if options or tag:
        path = iir_path or os.path.join(_build.ROOT, "stencils", name + ".iir")
        with open(path) as f:
            iir = f.read()
        backend = "b200-ico" if '"gridType": "Unstructured"' in iir else "b200"
        code = _codegen.compile_iir(iir, backend=backend, **options)
        tag = tag or _build.variant_tag(options)
        so = _build.compile_module(name, code, tag=tag)
    else:
        so = _build.module_path(name)
        if not os.path.exists(so):
            path = iir_path or os.path.join(_build.ROOT, "stencils", name + ".iir")
            with open(path) as f:
                iir = f.read()
            backend = "b200-ico" if '"gridType": "Unstructured"' in iir else "b200"
            so = _build.compile_module(name, _codegen.compile_iir(iir, backend=backend))
    return StencilModule(name, so)
