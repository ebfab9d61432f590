"""Goldens for EVERY optimized-IIR fixture the reference tree ships (dawn/test/**/*.iir, ~90 files written
by its own optimizer unit tests and serializer tests).  For each file the b200 emitter accepts, seeded
random inputs are pushed through the oracle (oracle/naive_interp.py, the pinned naive-backend
restatement) HERE, where /root/reference exists, and the fields the stencil changes are stored in
tests/golden/reftree.npz; tests/test_reference_tree_gpu.py regenerates the inputs from the same seed on
the GPU box and compares the generated kernels' results bit for bit.  The IIR files themselves are not
copied: the modules are compiled from where they lie by dawn_b200.build.build_reference_tree().

    python tests/golden/make_reftree_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import naive_interp  # noqa: E402

HALO = 3
CART_DOM = (11, 9, 6)     # interior I, J, K
MESH = (7, 5, 4)          # TriMesh nx, ny, levels


def seed_of(rel):
    return zlib.crc32(rel.encode()) & 0x7FFFFFFF


def cart_inputs(iir, rel):
    meta = iir["metadata"]
    names = meta["accessIDToName"]
    dims = meta["fieldIDtoDimensions"]
    I, J, K = CART_DOM
    ni, nj = I + 2 * HALO, J + 2 * HALO
    rng = np.random.default_rng(seed_of(rel))
    f = {}
    for aid in meta.get("APIFieldIDs", []):
        d = dims[str(aid)]
        c = d.get("cartesian_horizontal_dimension", {"mask_cart_i": 1, "mask_cart_j": 1})
        f[names[str(aid)]] = rng.uniform(0.5, 1.5, (K + 1 if d.get("mask_k", 0) else 1,
                                                    nj if c.get("mask_cart_j", 0) else 1,
                                                    ni if c.get("mask_cart_i", 0) else 1))
    return f, (ni, nj, K)


def allocated_fields(iir, dom):
    """Storages the wrapper owns (field versions, inter-stencil temporaries): zero-initialised."""
    meta = iir["metadata"]
    ni, nj, K = dom
    return {meta["accessIDToName"][str(a)]: np.zeros((K + 1, nj, ni)) for a in meta.get("allocatedFieldIDs", [])}


def ico_inputs(iir, rel, mesh):
    from dawn_b200.trimesh import TriMesh  # noqa: F401
    meta = iir["metadata"]
    names = meta["accessIDToName"]
    dims = meta["fieldIDtoDimensions"]
    loc = {"Cell": 0, "Edge": 1, "Vertex": 2}
    k = MESH[2]
    rng = np.random.default_rng(seed_of(rel))
    f = {}
    tmp = {}
    for aid in list(meta.get("APIFieldIDs", [])) + list(meta.get("temporaryFieldIDs", [])):
        d = dims[str(aid)]
        u = d.get("unstructured_horizontal_dimension")
        shape = [k] if d.get("mask_k", 0) else []
        if u:
            chain = u["iter_space"]["chain"] if "iter_space" in u else (
                u.get("sparse_part") or [u.get("dense_location_type")])
            shape.append(mesh.num(loc[chain[0]]))
            if len(chain) > 1:
                shape.append(mesh.nbh_table(chain, u.get("iter_space", {}).get("include_center", False)).shape[1])
        if aid in meta.get("APIFieldIDs", []):
            f[names[str(aid)]] = rng.uniform(0.5, 1.5, shape)
        else:
            tmp[names[str(aid)]] = np.zeros(shape)
    return f, k, tmp


def reads_outside_levels(iir, K):
    """True if some access k + shift leaves [0, K] (storages carry K + 1 levels) on a level of its
    DoMethod: such fixtures are optimizer-test inputs that were never meant to execute."""
    bad = []

    def visit(node, k0, k1):
        if isinstance(node, dict):
            fa = node.get("field_access_expr")
            if fa is not None:
                dk = fa.get("vertical_shift", 0)
                if k0 + dk < 0 or k1 + dk > K:
                    bad.append(fa.get("name"))
            for key, val in node.items():
                if key not in ("data", "loc"):
                    visit(val, k0, k1)
        elif isinstance(node, list):
            for x in node:
                visit(x, k0, k1)
    for st in iir["internalIR"]["stencils"]:
        for ms in st["multiStages"]:
            for sg in ms["stages"]:
                for dm in sg["doMethods"]:
                    lo, hi = naive_interp.interval_bounds(dm["interval"])
                    k0, k1 = naive_interp.resolve(lo, 0, K - 1), naive_interp.resolve(hi, 0, K - 1)
                    if k0 <= k1:
                        visit(dm["ast"], k0, k1)
    return bad


def main():
    from dawn_b200 import build
    from dawn_b200.trimesh import TriMesh
    manifest = build.build_reference_tree()
    out = {}
    status = {}
    mesh = TriMesh(MESH[0], MESH[1])
    for e in manifest:
        rel = e["file"]
        iir = naive_interp.load_iir(os.path.join(build.REFERENCE_ROOT, rel))
        outside = reads_outside_levels(iir, CART_DOM[2] if e["backend"] == "b200" else MESH[2])
        if outside:
            status[rel] = "skipped: reads %s outside the allocated levels" % sorted(set(outside))
            continue
        try:
            if e["backend"] == "b200":
                f, dom = cart_inputs(iir, rel)
                start = {k: v.copy() for k, v in f.items()}
                start.update(allocated_fields(iir, dom))
                res = naive_interp.run_cartesian(iir, dom, start, halos=(HALO,) * 4 + (0, 0))
            else:
                f, k, tmp = ico_inputs(iir, rel, mesh)
                start = {n: v.copy() for n, v in f.items()}
                start.update(tmp)
                res = naive_interp.run_unstructured(iir, mesh, k, start)
        except Exception as exc:  # noqa: BLE001
            status[rel] = "oracle: %s" % repr(exc)[:120]
            continue
        changed = [n for n in f if not np.array_equal(res[n], f[n], equal_nan=True)]
        for n in changed:
            out["%s::%s" % (rel, n)] = res[n]
        status[rel] = "ok (%d field(s) change)" % len(changed)
    np.savez_compressed(os.path.join(HERE, "reftree.npz"), **out)
    with open(os.path.join(HERE, "reftree_status.json"), "w") as fjs:
        json.dump(status, fjs, indent=1, sort_keys=True)
    ok = sum(1 for v in status.values() if v.startswith("ok"))
    print("%d fixtures with goldens, %d skipped" % (ok, len(status) - ok))
    for k, v in sorted(status.items()):
        if not v.startswith("ok"):
            print("  ", k, v)


if __name__ == "__main__":
    main()
