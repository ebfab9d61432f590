"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes access to oracle/liboracle.so (C restatements,
oracle/naive_stencils.c) and oracle/_ref/libdawn_ref.so (the reference's own code, see
oracle/ref_build/)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "naive_stencils.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if os.path.isdir("/root/reference/dawn/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def ref():
    """The reference's own code (None if oracle/_ref was never built, e.g. a bare checkout)."""
    global _ref
    if _ref is None:
        p = os.path.join(HERE, "_ref", "libdawn_ref.so")
        if not os.path.exists(p):
            return None
        _ref = ctypes.CDLL(p)
    return _ref


MASKS = {"ijk": 7, "ij": 3, "i": 1, "j": 2, "k": 4}


def fillmath(a, b, c, d, e, f, ni, nj, nk, dims="ijk", use_ref=False):
    """verifier::fillMath (driver-includes/verify.hpp:48-64) on a (ni,nj,nk) storage; returns an
    array shaped [k, j, i] with length 1 along masked dimensions."""
    mask = MASKS[dims]
    shape = (nk if mask & 4 else 1, nj if mask & 2 else 1, ni if mask & 1 else 1)
    out = np.zeros(shape)
    L = ref() if use_ref else lib()
    fn = L.ref_fillmath if use_ref else L.oracle_fillmath
    fn.restype = ctypes.c_int if use_ref else None
    fn(*(ctypes.c_double(x) for x in (a, b, c, d, e, f)), ni, nj, nk, mask, _ptr(out))
    return out


def _strides(arr):
    nk, nj, ni = arr.shape
    return ctypes.c_size_t(ni), ctypes.c_size_t(ni * nj)


def hori_diff_type2(out, inn, crlato, crlatu, hdmask, halo=3, nk=None, threads=1):
    nka, nj, ni = inn.shape
    nk = nka if nk is None else nk
    lap = np.zeros_like(inn)
    sj, sk = _strides(inn)
    if threads <= 1:
        lib().oracle_hori_diff_type2(ni, nj, nk, halo, sj, sk, _ptr(out), _ptr(inn), _ptr(crlato),
                                     _ptr(crlatu), _ptr(hdmask), _ptr(lap))
    else:
        arr = (_dp * 6)(_ptr(out), _ptr(inn), _ptr(crlato), _ptr(crlatu), _ptr(hdmask), _ptr(lap))
        lib().oracle_run_threaded(0, threads, ni, nj, nk, halo, sj, sk, arr, 6)
    return out


def hori_diff_01(u, out, halo=3, nk=None):
    nka, nj, ni = u.shape
    nk = nka if nk is None else nk
    lap = np.zeros_like(u)
    sj, sk = _strides(u)
    lib().oracle_hori_diff_01_krange(ni, nj, nk, halo, sj, sk, _ptr(u), _ptr(out), _ptr(lap), 0, nk - 1)
    return out


def hori_diff_dawn4py(inn, out, coeff, halo=3, nk=None):
    nka, nj, ni = inn.shape
    nk = nka if nk is None else nk
    lap = np.zeros_like(inn)
    sj, sk = _strides(inn)
    lib().oracle_hori_diff_dawn4py_krange(ni, nj, nk, halo, sj, sk, _ptr(inn), _ptr(out), _ptr(coeff),
                                          _ptr(lap), 0, nk - 1)
    return out


def tridiagonal_solve(d, a, b, c, halo=3, nk=None, threads=1):
    nka, nj, ni = d.shape
    nk = nka if nk is None else nk
    sj, sk = _strides(d)
    if threads <= 1:
        lib().oracle_tridiagonal_solve(ni, nj, nk, halo, sj, sk, _ptr(d), _ptr(a), _ptr(b), _ptr(c))
    else:
        arr = (_dp * 4)(_ptr(d), _ptr(a), _ptr(b), _ptr(c))
        lib().oracle_run_threaded(1, threads, ni, nj, nk, halo, sj, sk, arr, 4)
    return d, c


def tutorial_laplacian(out, inn, dx, nk, halo=3):
    _, nj, ni = inn.shape
    lib().oracle_tutorial_laplacian(ni, nj, nk, halo, ctypes.c_size_t(ni), ctypes.c_double(dx),
                                    _ptr(out), _ptr(inn))
    return out


def icon_laplacian_reference_timed(nx, ny, k_size, reps):
    """Wall time of `reps` calls of run() of the reference's generated naive-ico nabla2
    (ICON_laplacian_stencil_ref.cpp on toylib, oracle/_ref); returns (valid edges, seconds[reps]) or
    None when oracle/_ref was never built."""
    R = ref()
    if R is None or not hasattr(R, "ref_icon_laplacian_timed"):
        return None
    t = np.zeros(reps)
    ne = R.ref_icon_laplacian_timed(nx, ny, k_size, reps, _ptr(t))
    return (ne, t) if ne > 0 else None


_refcuda = None


def refcuda():
    """The reference's own pre-generated CUDA code recompiled for sm_100a (oracle/ref_build/ref_cuda.cu),
    or None when oracle/_ref was never built.  Entry points take device pointers."""
    global _refcuda
    if _refcuda is None:
        p = os.path.join(HERE, "_ref", "libdawn_refcuda.so")
        if not os.path.exists(p):
            return None
        _refcuda = ctypes.CDLL(p)
        _refcuda.refcuda_hori_diff.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [
            ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        _refcuda.refcuda_tridiagonal_solve.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4 + [
            ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    return _refcuda


_refcuda_small = {}


def refcuda_small(name):
    """One of the reference's small pre-generated CUDA texts recompiled for sm_100a
    (oracle/ref_build/ref_cuda_small.cu): conditional_stencil, global_indexing, nonoverlapping_stencil,
    copy_stencil (`refcuda_run(ni, nj, nk, halo, stride_j, stride_k, in_dev, out_dev, var1)`) and update_dz_c
    (`refcuda_update_dz_c`, ref_cuda_dzc.cu); None when
    oracle/_ref was never built."""
    if name not in _refcuda_small:
        p = os.path.join(HERE, "_ref", "libdawn_refcuda_%s.so" % name)
        if not os.path.exists(p):
            return None
        lib = ctypes.CDLL(p)
        if name == "update_dz_c":   # refcuda_update_dz_c(ni, nj, nk, halo, dt, double* const fields[9]): dense buffers
            lib.refcuda_update_dz_c.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_void_p]
        else:
            lib.refcuda_run.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2 + [ctypes.c_int]
        _refcuda_small[name] = lib
    return _refcuda_small[name]
