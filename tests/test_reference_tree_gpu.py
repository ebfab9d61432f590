"""Every optimized-IIR fixture the reference tree ships (dawn/test/**/*.iir), compiled by the b200
emitters where it lies (dawn_b200.build.build_reference_tree) and run on the GPU against goldens the
oracle produced in the build container (tests/golden/make_reftree_golden.py).  The IIR files are not
copied into the repo; inputs are regenerated from the per-file seed."""
import json
import os
import sys
import zlib

import numpy as np
import pytest

from util import HALO, bit_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MANIFEST = os.path.join(ROOT, "dawn_b200", "lib", "stencils", "reftree_manifest.json")
GOLD = os.path.join(ROOT, "tests", "golden", "reftree.npz")
STATUS = os.path.join(ROOT, "tests", "golden", "reftree_status.json")
CART_DOM = (11, 9, 6)
MESH = (7, 5, 4)


def _entries():
    if not (os.path.exists(MANIFEST) and os.path.exists(STATUS)):
        return []
    status = json.load(open(STATUS))
    return [e for e in json.load(open(MANIFEST))
            if "module" in e and status.get(e["file"], "").startswith("ok")]


ENTRIES = _entries()


def test_fixture_inventory():
    """67 of the 91 fixtures run; the rest are rejected with a diagnostic (inputs of the optimizer's own
    unit tests that have not been through stage splitting / field versioning yet, stages without a
    location type, racy halo writes) or use levels outside the storages the drivers allocate."""
    status = json.load(open(STATUS))
    assert sum(1 for v in status.values() if v.startswith("ok")) >= 67


@pytest.mark.gpu
@pytest.mark.parametrize("entry", ENTRIES, ids=[os.path.basename(e["file"]) for e in ENTRIES])
def test_reference_tree_fixture(entry):
    import torch
    import dawn_b200
    from dawn_b200.stencil import StencilModule
    gold = np.load(GOLD)
    rel = entry["file"]
    mod = StencilModule(entry["stencil"], os.path.join(ROOT, "dawn_b200", "lib", "stencils",
                                                       entry["module"] + ".so"))
    rng = np.random.default_rng(zlib.crc32(rel.encode()) & 0x7FFFFFFF)
    if entry["backend"] == "b200":
        I, J, K = CART_DOM
        ni, nj = I + 2 * HALO, J + 2 * HALO
        dom = dawn_b200.Domain(ni, nj, K)
        dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
        st = mod(dom)
        f, stor = {}, []
        for fd in mod.fields:
            d = fd["dims"]
            a = rng.uniform(0.5, 1.5, (K + 1 if "k" in d else 1, nj if "j" in d else 1, ni if "i" in d else 1))
            f[fd["name"]] = a
            stor.append(dawn_b200.Storage(ni, nj, a.shape[0], d).from_numpy(a))
        st.run(*stor)
        torch.cuda.synchronize()
        got = {fd["name"]: s.to_numpy() for fd, s in zip(mod.fields, stor)}
    else:
        from dawn_b200.trimesh import TriMesh
        loc = {"cells": 0, "edges": 1, "vertices": 2}
        m = TriMesh(MESH[0], MESH[1])
        k = MESH[2]
        st = mod.on_mesh(m, k)
        f, tens = {}, []
        for fd in mod.fields:
            shape = [k] if fd.get("k", 1) else []
            if fd["location"] != "none":
                shape.append(m.num(loc[fd["location"]]))
            if fd["sparse"]:
                shape.append(fd["sparse"])
            a = rng.uniform(0.5, 1.5, shape)
            f[fd["name"]] = a
            dev = np.ascontiguousarray(np.swapaxes(a, -1, -2)) if fd["sparse"] else a
            tens.append(torch.from_numpy(np.ascontiguousarray(dev)).cuda())
        st.run(*tens)
        torch.cuda.synchronize()
        got = {}
        for fd, t in zip(mod.fields, tens):
            a = t.cpu().numpy()
            a = np.swapaxes(a, -1, -2) if fd["sparse"] else a
            v = m.valid(loc[fd["location"]]) if fd["location"] != "none" else None
            got[fd["name"]] = (a, v, 1 if fd.get("k", 1) else 0)
    changed = [k2.split("::", 1)[1] for k2 in gold.files if k2.startswith(rel + "::")]
    for fd in mod.fields:
        n = fd["name"]
        want = gold[rel + "::" + n] if n in changed else f[n]
        if entry["backend"] == "b200":
            assert bit_equal(got[n], want), (rel, n)
        else:
            a, v, axis = got[n]
            if v is None:
                assert bit_equal(a, want), (rel, n)
            else:
                assert bit_equal(np.compress(v, a, axis=axis), np.compress(v, want, axis=axis)), (rel, n)
