"""TEST INFRASTRUCTURE: compiles a translation unit emitted by the b200-ico backend for the CPU emulation
(tests/cpu_emu/emu_cuda.hpp) and runs it on host memory through the module's own C ABI."""
import ctypes
import hashlib
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")

LAUNCH = re.compile(r"(\w+)<<<\s*([^,>]+?)\s*,\s*([^,>]+?)\s*,\s*([^,>]+?)\s*,[^>]*>>>\(")
DYN_SMEM = "extern __shared__ __align__(128) unsigned char b200_smem[];"


def emulated_source(code):
    """`kernel<<<grid, block, smem, stream>>>(args)` -> `EMU_LAUNCH(kernel, grid, block, smem, args)`; the
    dynamic shared memory of a kernel becomes the per-block buffer the launcher provides"""
    out, n = LAUNCH.subn(r"EMU_LAUNCH(\1, \2, \3, \4, ", code)
    out = out.replace(DYN_SMEM, "unsigned char* const b200_smem = emu_dynamic_smem;")
    out = out.replace("extern __shared__ ::dawn::float_type cc_smem[];",
                      "::dawn::float_type* const cc_smem = reinterpret_cast<::dawn::float_type*>(emu_dynamic_smem);")
    if "extern __shared__" in out:
        raise RuntimeError("a dynamic shared memory declaration was not recognised")
    if n != code.count("<<<"):
        raise RuntimeError("a kernel launch was not recognised")
    return out


def build(name, code, keep_k_chunks=True):
    os.makedirs(OUT, exist_ok=True)
    src = emulated_source(code)
    if keep_k_chunks:
        # generated launchers split k finer while a launch has fewer than 2 blocks per SM (296 blocks): on the
        # small domains an emulation can afford that leaves 2 levels per block and the steady state of the
        # TMA ring (refill while consuming) would never run - keep the planned chunk instead
        src = src.replace(" < 296) kchunk = (kchunk + 1) / 2;", " < 0) kchunk = (kchunk + 1) / 2;")
    # kernels with block barriers run their threads as host threads (EMU_BLOCK_THREADS), the others serially
    threads = ["-DEMU_BLOCK_THREADS", "-pthread"] if ("__syncthreads" in src or "tma::" in src) else []
    with open(os.path.join(HERE, "emu_rt.cpp")) as f, open(os.path.join(HERE, "emu_cuda.hpp")) as g:
        tag = hashlib.sha256((src + f.read() + g.read() + "Bsymbolic" + "".join(threads)).encode()).hexdigest()[:12]
    so = os.path.join(OUT, "%s_%s.so" % (name, tag))
    if not os.path.exists(so):
        cpp = os.path.join(OUT, "%s_%s.cpp" % (name, tag))
        with open(cpp, "w") as f:
            f.write(src)
        # one IEEE operation per node, like nvcc -fmad=false
        # -Bsymbolic: the module must call ITS stand-in runtime even when the real libdawn_b200rt.so has
        # been loaded into the process with RTLD_GLOBAL by another test
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                               "-Wl,-Bsymbolic"] + threads + [
                               "-include", os.path.join(HERE, "emu_cuda.hpp"), "-I", os.path.join(ROOT, "dawn_b200"),
                               "-o", so, cpp, os.path.join(HERE, "emu_rt.cpp")])
    return ctypes.CDLL(so)


class NbhTable(ctypes.Structure):
    _fields_ = [("chain", ctypes.c_int * 8), ("chain_len", ctypes.c_int), ("include_center", ctypes.c_int),
                ("nbh_size", ctypes.c_int), ("table", ctypes.c_void_p)]


class Splitter(ctypes.Structure):
    _fields_ = [("location", ctypes.c_int), ("subdomain", ctypes.c_int), ("offset", ctypes.c_int),
                ("index", ctypes.c_int)]


class MeshStruct(ctypes.Structure):
    _fields_ = [("num_cells", ctypes.c_int), ("num_edges", ctypes.c_int), ("num_vertices", ctypes.c_int),
                ("num_tables", ctypes.c_int), ("tables", ctypes.POINTER(NbhTable)), ("num_splitters", ctypes.c_int),
                ("splitters", ctypes.POINTER(Splitter))]


class Field(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("stride_j", ctypes.c_int), ("stride_k", ctypes.c_int)]


def run(name, code, info, mesh, k, fields, globals_=None, splitters=None):
    """Runs module `name` (emitted `code`, module info dict `info`) on `mesh` with `fields` in the oracle's
    layout (dense [k, n], sparse [k, n, nbh], vertical [k], horizontal [n]); returns the fields afterwards
    and the number of kernel launches."""
    lib = build(name, code)
    keep, entries = [], []
    for t in info["tables"]:
        host = np.ascontiguousarray(mesh.nbh_table(list(t["chain"]), bool(t["include_center"]), nbh_size=int(t["size"])),
                                    dtype=np.int32)
        soa = np.ascontiguousarray(host.T)                      # [nbh][element], like b200rt_upload_nbh_table
        keep.append(soa)
        e = NbhTable()
        for n, c in enumerate(t["chain"]):
            e.chain[n] = c
        e.chain_len, e.include_center, e.nbh_size, e.table = len(t["chain"]), int(t["include_center"]), int(t["size"]), soa.ctypes.data
        entries.append(e)
    tabs = (NbhTable * max(1, len(entries)))(*entries)
    spl = [Splitter(int(l), int(s), int(o), int(i)) for (l, s, o), i in (splitters or {}).items()]
    spls = (Splitter * max(1, len(spl)))(*spl)
    ms = MeshStruct(mesh.num_cells, mesh.num_edges, mesh.num_vertices, len(entries), tabs, len(spl), spls)
    handle = ctypes.c_void_p()
    lib.b200rt_last_error.restype = ctypes.c_char_p
    setup = getattr(lib, "setup_" + name)
    setup.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(MeshStruct), ctypes.c_int, ctypes.c_void_p]
    if setup(ctypes.byref(handle), ctypes.byref(ms), k, None):
        raise RuntimeError(lib.b200rt_last_error().decode())
    for g, v in (globals_ or {}).items():
        fn = getattr(lib, "set_%s_%s" % (name, g))
        fn.argtypes = [ctypes.c_void_p, ctypes.c_double]
        fn(handle, float(v))
    n_of = {"cells": mesh.num_cells, "edges": mesh.num_edges, "vertices": mesh.num_vertices, "none": 1}
    dev, fs = [], []
    for fd in info["fields"]:
        a = np.array(fields[fd["name"]], dtype=np.float32 if info.get("float_type") == "float" else np.float64)
        if fd["sparse"]:
            a = np.swapaxes(a, -1, -2)                           # [k, n, nbh] -> [k, nbh, n]
        a = np.ascontiguousarray(a)
        dev.append(a)
        n = n_of[fd["location"]]
        sp = fd["sparse"] or 1
        fs.append(Field(a.ctypes.data, n if fd["sparse"] else 0, n * sp))
    arr = (Field * len(fs))(*fs)
    runf = getattr(lib, "run_" + name)
    runf.argtypes = [ctypes.c_void_p, ctypes.POINTER(Field), ctypes.c_int, ctypes.c_void_p]
    if runf(handle, arr, len(fs), None):
        raise RuntimeError(lib.b200rt_last_error().decode())
    launches = getattr(lib, "launches_" + name)
    launches.restype = ctypes.c_long
    launches.argtypes = [ctypes.c_void_p]
    nl = launches(handle)
    status = getattr(lib, "chain_status_" + name)   # chained launches: every consumer block saw its producers
    status.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    if status(handle, None):
        raise RuntimeError(lib.b200rt_last_error().decode())
    freef = getattr(lib, "free_" + name)
    freef.argtypes = [ctypes.c_void_p]
    freef(handle)
    lib.emu_guard_violations.restype = ctypes.c_long
    if lib.emu_guard_violations():
        raise RuntimeError("a kernel of %s wrote outside a scratch allocation" % name)
    lib.emu_bulk_copies_issued.restype = ctypes.c_long
    run.bulk_copies = lib.emu_bulk_copies_issued()   # > 0: blocks took the staged (cp.async.bulk) variant
    out = {}
    for fd, a in zip(info["fields"], dev):
        out[fd["name"]] = np.swapaxes(a, -1, -2) if fd["sparse"] else a
    return out, nl


def run_cartesian(name, code, info, dims, halos, fields, globals_=None, keep_k_chunks=True):
    """Runs an emitted Cartesian module (plain kernels: no TMA, no barriers) on dense host arrays
    [k][j][i] (masked dimensions have length 1) and returns them with the launch count."""
    lib = build(name, code, keep_k_chunks)
    ni, nj, nk = dims
    handle = ctypes.c_void_p()
    lib.b200rt_last_error.restype = ctypes.c_char_p
    setup = getattr(lib, "setup_" + name)
    setup.argtypes = [ctypes.POINTER(ctypes.c_void_p)] + [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    if setup(ctypes.byref(handle), ni, nj, nk, (ctypes.c_int * 6)(*halos), None):
        raise RuntimeError(lib.b200rt_last_error().decode())
    for g, v in (globals_ or {}).items():
        fn = getattr(lib, "set_%s_%s" % (name, g))
        fn.argtypes = [ctypes.c_void_p, ctypes.c_double]
        fn(handle, float(v))
    dev, fs, keep, views = [], [], [], []
    ft = np.float32 if info.get("float_type") == "float" else np.float64   # ::dawn::float_type of the module
    es = np.dtype(ft).itemsize
    line = 128 // es
    for fd in info["fields"]:
        src = np.ascontiguousarray(np.array(fields[fd["name"]], dtype=ft))
        # the layout of dawn_b200.Storage / driver-includes storage_info: rows padded to 16 elements, a lead pad
        # that puts the first interior point (i = halo) on a 128-byte line; wrapper-allocated fields of a
        # generated class get exactly this layout and the launchers insist that all ijk fields share strides
        d = fd["dims"]
        lk, lj, li = src.shape
        h = halos[0]
        s_, lead, sj, sk = 1, 0, 0, 0
        if "i" in d:
            lead = (line - h % line) % line
            s_ = (li + line - 1) // line * line
        if "j" in d:
            sj = s_
            s_ *= lj
        if "k" in d:
            sk = s_
            s_ *= lk
        raw = np.zeros(s_ + lead + line + 2 * line, dtype=ft)
        off = 0
        while (raw.ctypes.data + es * off) % 256:
            off += 1
        base = raw[off + lead:]
        view = np.lib.stride_tricks.as_strided(base, shape=(lk, lj, li), strides=(es * sk, es * sj, es))
        view[...] = src
        keep.append(raw)
        views.append(view)
        fs.append(Field(base.ctypes.data, sj, sk))
    arr = (Field * len(fs))(*fs)
    runf = getattr(lib, "run_" + name)
    runf.argtypes = [ctypes.c_void_p, ctypes.POINTER(Field), ctypes.c_int, ctypes.c_void_p]
    if runf(handle, arr, len(fs), None):
        raise RuntimeError(lib.b200rt_last_error().decode())
    launches = getattr(lib, "launches_" + name)
    launches.restype = ctypes.c_long
    launches.argtypes = [ctypes.c_void_p]
    nl = launches(handle)
    freef = getattr(lib, "free_" + name)
    freef.argtypes = [ctypes.c_void_p]
    freef(handle)
    lib.emu_guard_violations.restype = ctypes.c_long
    if lib.emu_guard_violations():
        raise RuntimeError("a kernel of %s wrote outside a scratch allocation" % name)
    lib.emu_tensor_maps_encoded.restype = ctypes.c_long
    run_cartesian.tensor_maps = lib.emu_tensor_maps_encoded()   # > 0: the TMA variant of a kernel was chosen
    return {fd["name"]: np.array(v) for fd, v in zip(info["fields"], views)}, nl
