"""GPU sweep of b200-ico planner knobs on ICON nabla2 (kernel-only timing)."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dawn_b200
from dawn_b200.trimesh import TriMesh
n = int(sys.argv[1]); K = 65
axes = []
combos = None
if len(sys.argv) > 2 and sys.argv[2] == "list":   # explicit variants 'k=v,k=v' ('' = default), cf. tools/precompile.py
    combos = [[(kv.split("=")[0], int(kv.split("=")[1])) for kv in spec.split(",") if kv] for spec in sys.argv[3:]]
else:
    for a in sys.argv[2:]:
        k, vs = a.split("="); axes.append([(k, int(v)) for v in vs.split(",")])
mesh = TriMesh(n, n)
E, C, V = mesh.num_edges, mesh.num_cells, mesh.num_vertices
nof = {"cells": C, "edges": E, "vertices": V}
base = dawn_b200.load_stencil("ICON_laplacian_stencil")
tens = []
for fd in base.fields:
    shape = (K, fd["sparse"], nof[fd["location"]]) if fd["sparse"] else (K, nof[fd["location"]])
    tens.append(torch.rand(shape, dtype=torch.float64, device="cuda") + 0.5)
bytes_run = K * (40 * E + 56 * V + 32 * C) + 4 * (6 * V + 3 * C + 4 * E)
written = [i for i, fd in enumerate(base.fields) if fd["name"] in base.info["written"]]
first = None   # results of the first variant (run '' first): every other variant must reproduce them bit for bit
for combo in (combos if combos is not None else itertools.product(*axes)):
    opts = dict(combo)
    mod = dawn_b200.load_stencil("ICON_laplacian_stencil", **opts)
    st = mod.on_mesh(mesh, K)
    for i in written: tens[i].fill_(-1.0)
    for _ in range(3): st.run(*tens)
    torch.cuda.synchronize()
    if first is None:
        first = [tens[i].clone() for i in written]
    same = all(torch.equal(tens[i], r) for i, r in zip(written, first))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): st.run(*tens)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-60s %.3f ms  %.0f GB/s  %.1f%%  %s" % (opts, ms, bytes_run / ms / 1e6, bytes_run / ms / 1e6 / 65.415,
                                                  "same bits as the first variant" if same else "RESULTS DIFFER"), flush=True)
