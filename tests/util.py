"""Shared helpers for the parity tests: reference-style inputs and oracle drivers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import capi, naive_interp  # noqa: E402

HALO = 3
STENCILS = os.path.join(ROOT, "stencils")


def iir_path(name):
    return os.path.join(STENCILS, name + ".iir")


# fillMath constants of the reference drivers
#   gtclang/test/integration-test/CodeGen/hori_diff_type2_stencil_benchmark.cpp:59-64
#   gtclang/test/integration-test/CodeGen/tridiagonal_solve_benchmark.cpp:53-58
FILL = {
    "u": (8.0, 2.0, 1.5, 1.5, 2.0, 4.0),
    "hdmask": (5.0, 2.2, 1.7, 1.9, 2.0, 4.0),
    "crlato": (6.5, 1.2, 1.7, 0.9, 2.1, 2.0),
    "crlatu": (5.0, 2.2, 1.7, 0.9, 2.0, 1.0),
    "d": (8.0, 2.0, 1.5, 1.5, 2.0, 4.0),
    "a": (7.4, 2.0, 1.5, 1.3, 2.1, 3.0),
    "b": (8.0, 2.0, 1.4, 1.2, 2.3, 3.0),
    "c": (7.8, 2.0, 1.1, 1.7, 1.9, 4.1),
}


def fill(key, ni, nj, nk, dims="ijk"):
    return capi.fillmath(*FILL[key], ni, nj, nk, dims)


def hori_diff_type2_inputs(I, J, K):
    ni, nj = I + 2 * HALO, J + 2 * HALO
    return {
        "out": np.full((K + 1, nj, ni), -1.0),
        "in": fill("u", ni, nj, K + 1),
        "crlato": fill("crlato", ni, nj, K + 1, "j"),
        "crlatu": fill("crlatu", ni, nj, K + 1, "j"),
        "hdmask": fill("hdmask", ni, nj, K + 1),
    }


def tridiagonal_inputs(I, J, K, dominant=True):
    ni, nj = I + 2 * HALO, J + 2 * HALO
    f = {k: fill(k, ni, nj, K + 1) for k in ("d", "a", "b", "c")}
    if dominant:  # keep the system well conditioned at every size
        f["b"] = f["b"] + 12.0
    return f


def run_oracle(name, dom, fields, globals_=None):
    out = {k: v.copy() for k, v in fields.items()}
    naive_interp.run_cartesian(iir_path(name), dom, out, halos=(HALO,) * 4 + (0, 0), globals_=globals_)
    return out


def bit_equal(a, b):
    return np.array_equal(a, b, equal_nan=True)


def unstructured_oracle_on_levels(name, mesh, module_fields, device_tensors, levels):
    """Oracle run of a level-independent unstructured stencil (no vertical offsets: ICON nabla2) on a SAMPLE
    of levels of device-resident fields: copies the sampled levels to the host (sparse fields come as
    [k, nbh, n] from the device and go to the oracle as [k, n, nbh]), runs the interpreter with
    k_size = len(levels) and returns {field: array[len(levels), ...]} in the oracle's layout.  Lets the
    benchmark-scale runs (65 levels, millions of elements) be checked in seconds."""
    import json
    import torch
    idx = torch.as_tensor(list(levels), device=device_tensors[0].device)
    f = {}
    for fd, t in zip(module_fields, device_tensors):
        a = t.index_select(0, idx).cpu().numpy()
        f[fd["name"]] = np.ascontiguousarray(np.swapaxes(a, -1, -2)) if fd["sparse"] else a
    d = json.load(open(iir_path(name)))["metadata"]
    for aid in d.get("temporaryFieldIDs", []):     # the interpreter wants storage for temporaries too
        dims = d["fieldIDtoDimensions"][str(aid)]
        chain = dims["unstructured_horizontal_dimension"]["iter_space"]["chain"]
        assert len(chain) == 1
        f.setdefault(d["accessIDToName"][str(aid)], np.zeros((len(levels), mesh.num(chain[0]))))
    return naive_interp.run_unstructured(iir_path(name), mesh, len(levels), f)
