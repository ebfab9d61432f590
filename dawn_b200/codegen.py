"""Python binding of the emitter's C ABI (include/dawn_b200_codegen.h).

Mirrors the compile entry point Python users of the reference have, `dawn4py.compile(sir, groups,
backend=CodeGenBackend.CUDA, **kwargs)` (dawn/src/dawn4py/__init__.py:88-129; enum bound in
dawn/src/dawn4py/_dawn4py.cpp:61-67), for the code-generation half: optimized IIR in, source out.
"""
import ctypes
import enum
import os

from . import build as _build


class CodeGenBackend(enum.IntEnum):
    """dawn::codegen::Backend (CodeGen/Options.h:22) + the two backends of this project."""
    GridTools = 0
    CXXNaive = 1
    CXXNaiveIco = 2
    CUDAIco = 3
    CUDA = 4
    CXXOpt = 5
    B200 = 6
    B200Ico = 7


class CodeGenError(RuntimeError):
    pass


class _Options(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "max_halo_points", "use_parallel_ep", "max_blocks_per_sm", "nsms", "domain_size_i",
        "domain_size_j", "domain_size_k", "run_with_sync", "block_size_i", "block_size_j",
        "rows_per_thread", "block_size_k", "inline_temporaries", "fuse_columns", "unroll_k",
        "levels_per_thread", "block_size_ico", "use_tma", "tma_stages", "column_cache", "column_ring", "ring_levels", "prefetch_k", "cols_per_thread",
        "shared_temporaries", "co_schedule", "ring_drain", "ring_early", "ring_stagger",
        "ico_full_blocks", "pair_stores", "ico_cache_hints", "ico_chain", "ico_chain_l1", "ico_chain_lag", "ring_producer", "single_precision", "ico_stage", "ico_hoist", "persistent_ring")]


_lib = None


def _load():
    global _lib
    if _lib is None:
        so = _build.build_emitter()
        _lib = ctypes.CDLL(so)
        _lib.dawn_b200_codegen_last_error.restype = ctypes.c_char_p
        _lib.dawn_b200_parse_backend.argtypes = [ctypes.c_char_p]
        _lib.dawn_b200_free.argtypes = [ctypes.c_void_p]
    return _lib


def parse_backend_string(s: str) -> CodeGenBackend:
    """dawn::codegen::parseBackendString (CodeGen/Driver.cpp:47-64)."""
    v = _load().dawn_b200_parse_backend(s.encode())
    if v < 0:
        raise ValueError("Backend not supported")
    return CodeGenBackend(v)


def default_options(**kw) -> _Options:
    o = _Options()
    _load().dawn_b200_default_options(ctypes.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError("unknown codegen option %r" % k)
        setattr(o, k, int(v))
    return o


def compile_iir(iir_json: str, backend="b200", name="restoredIIR", **options) -> str:
    """Optimized IIR (proto3 JSON text) -> CUDA translation unit (string)."""
    lib = _load()
    be = parse_backend_string(backend) if isinstance(backend, str) else CodeGenBackend(backend)
    opts = default_options(**options)
    data = iir_json.encode() if isinstance(iir_json, str) else iir_json
    names = (ctypes.c_char_p * 1)(name.encode())
    bufs = (ctypes.c_char_p * 1)(data)
    lens = (ctypes.c_size_t * 1)(len(data))
    out = ctypes.c_void_p()
    out_len = ctypes.c_size_t()
    rc = lib.dawn_b200_codegen_run(names, bufs, lens, 1, int(be), ctypes.byref(opts),
                                   ctypes.byref(out), ctypes.byref(out_len))
    if rc != 0:
        raise CodeGenError(lib.dawn_b200_codegen_last_error().decode())
    try:
        return ctypes.string_at(out.value, out_len.value).decode()
    finally:
        lib.dawn_b200_free(out)


def ico_chain_size(chain, include_center=False) -> int:
    """ICOChainSize (dawn/src/dawn/CodeGen/IcoChainSizes.cpp:199-264). chain: 0 cell,1 edge,2 vertex."""
    arr = (ctypes.c_int * len(chain))(*chain)
    return _load().dawn_b200_ico_chain_size(arr, len(chain), int(include_center))


def interface_text(iir_json: str, kind="c") -> str:
    """C header (kind="c") or Fortran interface module (kind="f90") of an unstructured stencil's C ABI
    (the reference's --output-c-header / --output-f90-interface)."""
    lib = _load()
    data = iir_json.encode() if isinstance(iir_json, str) else iir_json
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    rc = lib.dawn_b200_codegen_interface(data, ctypes.c_size_t(len(data)), 0 if kind == "c" else 1,
                                         ctypes.byref(out), ctypes.byref(n))
    if rc != 0:
        raise CodeGenError(lib.dawn_b200_codegen_last_error().decode())
    try:
        return ctypes.string_at(out.value, n.value).decode()
    finally:
        lib.dawn_b200_free(out)


def iir_convert(data, to="byte") -> bytes:
    """Serialized StencilInstantiation -> the other IIRSerializer format ("json" text or "byte" wire
    format).  Input format is detected from the data."""
    lib = _load()
    raw = data.encode() if isinstance(data, str) else bytes(data)
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    rc = lib.dawn_b200_iir_convert(raw, ctypes.c_size_t(len(raw)), 0 if to == "json" else 1,
                                   ctypes.byref(out), ctypes.byref(n))
    if rc != 0:
        raise CodeGenError(lib.dawn_b200_codegen_last_error().decode())
    try:
        return ctypes.string_at(out.value, n.value)
    finally:
        lib.dawn_b200_free(out)


def iir_schema() -> str:
    """The wire schema the Byte reader implements (one line per field / enumerator)."""
    lib = _load()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    if lib.dawn_b200_iir_schema(ctypes.byref(out), ctypes.byref(n)) != 0:
        raise CodeGenError(lib.dawn_b200_codegen_last_error().decode())
    try:
        return ctypes.string_at(out.value, n.value).decode()
    finally:
        lib.dawn_b200_free(out)
