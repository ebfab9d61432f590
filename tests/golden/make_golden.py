"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libdawn_ref.so, built from
/root/reference by oracle/ref_build/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The committed vectors pin the oracle on machines where neither /root/reference nor oracle/_ref exist.
Inputs are regenerated from the seeds recorded here, outputs are stored.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import capi  # noqa: E402

dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp)  # noqa: E731
H = 3


def cart_inputs(seed, names, ni, nj, nk, lo=0.5, hi=1.5):
    rng = np.random.default_rng(seed)
    return {n: rng.uniform(lo, hi, (nk + 1, nj, ni)) for n in names}


def main():
    R = capi.ref()
    assert R is not None, "oracle/_ref not built"
    out = {}
    # fillMath, reference driver constants (hori_diff_type2_stencil_benchmark.cpp:59-64)
    for key, c, dims in (("u", (8.0, 2.0, 1.5, 1.5, 2.0, 4.0), "ijk"),
                         ("crlato", (6.5, 1.2, 1.7, 0.9, 2.1, 2.0), "j"),
                         ("hdmask_ij", (5.0, 2.2, 1.7, 1.9, 2.0, 4.0), "ij")):
        out["fillmath_" + key] = capi.fillmath(*c, 18, 18, 11, dims, use_ref=True)
    np.savez_compressed(os.path.join(HERE, "fillmath.npz"), **out)

    I, J, K = 12, 14, 20
    ni, nj, nk = I + 2 * H, J + 2 * H, K
    # conditional_stencil (reference/conditional_stencil.cpp), both branches of the global
    res = {}
    for var1 in (1, 0):
        f = cart_inputs(11, ["in"], ni, nj, nk, -1, 1)
        o = np.full_like(f["in"], -1.0)
        R.ref_conditional_stencil(ni, nj, nk, nk + 1, H, var1, P(f["in"]), P(o))
        res["out_var1_%d" % var1] = o
    np.savez_compressed(os.path.join(HERE, "conditional_stencil.npz"), seed=11, dims=(ni, nj, nk), **res)
    # update_dz_c (reference/update_dz_c.cpp), dt = 0.3
    names = ["dp_ref", "zs", "area", "ut", "vt", "gz", "gz_x", "gz_y", "ws3"]
    f = cart_inputs(12, names, ni, nj, nk)
    arr = (dp * 9)(*[P(f[n]) for n in names])
    R.ref_update_dz_c(ni, nj, nk, nk + 1, H, ctypes.c_double(0.3), arr)
    np.savez_compressed(os.path.join(HERE, "update_dz_c.npz"), seed=12, dims=(ni, nj, nk), dt=0.3,
                        gz=f["gz"], ws3=f["ws3"])
    # global_indexing / nonoverlapping_stencil
    f = cart_inputs(13, ["in"], ni, nj, nk, -1, 1)
    o = np.full_like(f["in"], -1.0)
    R.ref_global_indexing(ni, nj, nk, nk + 1, H, P(f["in"]), P(o))
    o2 = np.full_like(f["in"], -1.0)
    R.ref_nonoverlapping_stencil(ni, nj, nk, nk + 1, H, P(f["in"]), P(o2))
    np.savez_compressed(os.path.join(HERE, "indexing.npz"), seed=13, dims=(ni, nj, nk),
                        global_indexing=o, nonoverlapping_from_k11=o2[11:])
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
