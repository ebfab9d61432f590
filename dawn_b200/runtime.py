"""ctypes binding of the C-ABI runtime libdawn_b200rt.so (include/dawn_b200_rt.h)."""
import ctypes
import os

from . import build as _build


class RuntimeError_(RuntimeError):
    pass


class Field(ctypes.Structure):
    """b200_field_t"""
    _fields_ = [("data", ctypes.c_void_p), ("stride_j", ctypes.c_int), ("stride_k", ctypes.c_int)]


class VerificationMetrics(ctypes.Structure):
    """b200_verification_metrics (twin of driver-includes/verification_metrics.hpp:3-10)"""
    _fields_ = [("iteration", ctypes.c_int), ("is_valid", ctypes.c_int),
                ("max_rel_err", ctypes.c_double), ("min_rel_err", ctypes.c_double),
                ("max_abs_err", ctypes.c_double), ("min_abs_err", ctypes.c_double),
                ("n_failed", ctypes.c_longlong)]


_lib = None

# every symbol include/dawn_b200_rt.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "b200rt_last_error", "b200rt_set_error", "b200rt_device_count", "b200rt_set_device",
    "b200rt_device_synchronize", "b200rt_malloc", "b200rt_free", "b200rt_malloc_host",
    "b200rt_free_host", "b200rt_memset", "b200rt_memcpy_h2d", "b200rt_memcpy_d2h",
    "b200rt_memcpy_d2d", "b200rt_memcpy2d", "b200rt_stream_create", "b200rt_stream_destroy",
    "b200rt_stream_synchronize", "b200rt_event_create", "b200rt_event_destroy",
    "b200rt_event_record", "b200rt_event_synchronize", "b200rt_event_elapsed_ms",
    "b200rt_check_last_launch", "b200rt_tensor_map_3d", "b200rt_tensor_map_3d_es", "b200rt_verify_field", "b200rt_verify_field_f32",
    "b200rt_reshape", "b200rt_reshape_back", "b200rt_reshape_f32", "b200rt_reshape_back_f32",
    "b200rt_stream_wait_event", "b200rt_chain_plan_create", "b200rt_chain_plan_view", "b200rt_chain_plan_next_epoch", "b200rt_chain_plan_status",
    "b200rt_chain_plan_destroy",
    "b200rt_upload_nbh_table", "b200rt_nccl_unique_id", "b200rt_comm_init_rank",
    "b200rt_comm_destroy", "b200rt_halo_plan_create", "b200rt_halo_plan_create2",
    "b200rt_halo_exchange", "b200rt_halo_plan_ipc_handle", "b200rt_halo_plan_connect",
    "b200rt_halo_exchange_p2p", "b200rt_halo_plan_p2p_status",
    "b200rt_halo_plan_destroy", "b200rt_metrics_record_create", "b200rt_metrics_record_add",
    "b200rt_metrics_record_json", "b200rt_metrics_record_write", "b200rt_metrics_record_destroy",
    "b200rt_graph_begin", "b200rt_graph_end", "b200rt_graph_launch", "b200rt_graph_destroy",
]


def lib():
    """Loads (building in-tree if needed) the runtime.  There is no fallback: without the CUDA
    runtime library the product cannot run and this raises."""
    global _lib
    if _lib is None:
        so = os.path.join(_build.LIB, "libdawn_b200rt.so")
        if not os.path.exists(so):
            so = _build.build_runtime()
        _lib = ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
        _lib.b200rt_last_error.restype = ctypes.c_char_p
        for name in ("b200rt_malloc", "b200rt_malloc_host"):
            getattr(_lib, name).argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
        for name in ("b200rt_free", "b200rt_free_host"):
            getattr(_lib, name).argtypes = [ctypes.c_void_p]
        for name in ("b200rt_memcpy_h2d", "b200rt_memcpy_d2h", "b200rt_memcpy_d2d"):
            getattr(_lib, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                            ctypes.c_void_p]
        _lib.b200rt_verify_field.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_double, ctypes.c_double,
            ctypes.c_int, ctypes.POINTER(VerificationMetrics), ctypes.c_void_p]
        for name in ("b200rt_reshape", "b200rt_reshape_back"):
            getattr(_lib, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.b200rt_upload_nbh_table.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_void_p]
        _lib.b200rt_comm_init_rank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_int]
        _lib.b200rt_comm_destroy.argtypes = [ctypes.c_void_p]
        _lib.b200rt_nccl_unique_id.argtypes = [ctypes.c_void_p]
        _lib.b200rt_halo_plan_create.argtypes = [
            ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p] + [ctypes.c_int] * 10
        _lib.b200rt_halo_plan_create2.argtypes = [
            ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p] + [ctypes.c_int] * 11
        _lib.b200rt_halo_exchange.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.b200rt_halo_plan_destroy.argtypes = [ctypes.c_void_p]
        _lib.b200rt_halo_plan_ipc_handle.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.b200rt_halo_plan_connect.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        _lib.b200rt_halo_exchange_p2p.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.b200rt_halo_plan_p2p_status.argtypes = [ctypes.c_void_p]
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError_("dawn_b200 runtime error %d: %s" %
                            (rc, lib().b200rt_last_error().decode(errors="replace")))


def verify_field(dsl_ptr, ref_ptr, n, rel_tol=1e-12, abs_tol=0.0, iteration=0, stream=None):
    """GPU-side single-pass verification (replaces cuda_verify.hpp:84-196)."""
    m = VerificationMetrics()
    check(lib().b200rt_verify_field(dsl_ptr, ref_ptr, n, rel_tol, abs_tol, iteration,
                                    ctypes.byref(m), stream))
    return m


class MetricsRecord:
    """Verification record with the JSON shape of the reference's MetricsSerialiser
    (driver-includes/to_json.cpp): {"<stencil>": [{"<field>": {...}} per iteration]}."""

    def __init__(self):
        self.handle = ctypes.c_void_p()
        check(lib().b200rt_metrics_record_create(ctypes.byref(self.handle)))

    def add(self, stencil, field, metrics):
        fn = lib().b200rt_metrics_record_add
        fn.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p,
                       ctypes.POINTER(VerificationMetrics)]
        check(fn(self.handle, stencil.encode(), field.encode(), ctypes.byref(metrics)))

    def json(self) -> str:
        fn = lib().b200rt_metrics_record_json
        fn.argtypes = [ctypes.c_void_p]
        fn.restype = ctypes.c_char_p
        return fn(self.handle).decode()

    def write(self, path):
        fn = lib().b200rt_metrics_record_write
        fn.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        check(fn(self.handle, str(path).encode()))

    def close(self):
        if self.handle:
            fn = lib().b200rt_metrics_record_destroy
            fn.argtypes = [ctypes.c_void_p]
            fn(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def connect_halo_plan(plan, rank, world, all_gather):
    """Peer-memory set-up of a halo plan (include/dawn_b200_rt.h, b200rt_halo_exchange_p2p): every rank
    publishes the CUDA IPC handle of its landing buffers and maps those of its slab neighbours.
    `all_gather(bytes) -> [bytes of rank 0, ...]` is the host's collective, e.g.
    torch.distributed.all_gather_object wrapped in a lambda.  Collective: every rank must call it."""
    buf = ctypes.create_string_buffer(64)
    check(lib().b200rt_halo_plan_ipc_handle(plan, buf))
    handles = all_gather(buf.raw)
    lo = handles[rank - 1] if rank > 0 else None
    hi = handles[rank + 1] if rank + 1 < world else None
    check(lib().b200rt_halo_plan_connect(plan, lo, hi))


def device_numa_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off (PCI bus id -> sysfs), intersected with what this
    process may run on; None when the platform does not say (no NUMA information, container mask)."""
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None


class near_device:
    """Context manager: runs the enclosed host code on the CPUs next to the GPU, so that host buffers
    allocated (first touched, pinned) inside are local to the socket the GPU's PCIe link belongs to.
    With one process per GPU on a two-socket box this keeps every rank's staging traffic off the
    inter-socket link.  A no-op where the platform exposes no NUMA topology."""

    def __init__(self, device_index):
        self.cpus = device_numa_cpus(device_index)
        self.saved = None

    def __enter__(self):
        if self.cpus:
            self.saved = os.sched_getaffinity(0)
            os.sched_setaffinity(0, self.cpus)
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False
