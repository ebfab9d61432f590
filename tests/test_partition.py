"""Partitioned unstructured meshes (dawn_b200/partition.py): ownership is a partition of the valid
elements, local meshes reproduce the undivided result on what they own bit for bit, and the halo
refresh moves exactly the owners' values (in-process emulation and two gloo ranks)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import ROOT, bit_equal, iir_path
from oracle import naive_interp
from dawn_b200 import partition as pt, reorder as ro
from dawn_b200.trimesh import CELL, EDGE, VERTEX, TriMesh

LOC_OF = {"vec": EDGE, "div_vec": CELL, "rot_vec": VERTEX, "nabla2_vec": EDGE, "primal_edge_length": EDGE,
          "dual_edge_length": EDGE, "tangent_orientation": EDGE, "__tmp_nab_60": EDGE, "__tmp_nab_61": EDGE,
          "geofac_rot": VERTEX, "geofac_div": CELL}


def nabla2_fields(m, k, seed):
    rng = np.random.default_rng(seed)
    f = {n: rng.uniform(0.5, 1.5, (k, m.num(loc))) for n, loc in LOC_OF.items() if not n.startswith("geofac")}
    f["geofac_rot"] = rng.uniform(-1, 1, (k, m.num_vertices, 6))
    f["geofac_div"] = rng.uniform(-1, 1, (k, m.num_cells, 3))
    return f


@pytest.mark.parametrize("nparts", [2, 3, 5])
def test_every_valid_element_has_one_owner(nparts):
    m = ro.reorder(TriMesh(9, 7), "rcm")
    own = pt.owners(m, nparts)
    assert (np.diff(own[CELL]) >= 0).all() and set(own[CELL]) == set(range(nparts))   # contiguous ranges
    sizes = np.bincount(own[CELL])
    assert sizes.max() - sizes.min() <= 1
    e2c = m.nbh_table([EDGE, CELL])
    for e in range(m.num_edges):
        cells = [c for c in e2c[e] if c >= 0]
        assert own[EDGE][e] == (min(own[CELL][c] for c in cells) if cells else -1)       # lowest-owner rule
    assert ((own[EDGE] >= 0) == m.valid(EDGE)).all() and (own[VERTEX] >= 0).all()
    parts = pt.partition(m, nparts)
    for loc in (CELL, EDGE, VERTEX):
        owned = np.concatenate([p.global_ids[loc][:p.n_owned[loc]] for p in parts])
        assert np.array_equal(np.sort(owned), np.nonzero(own[loc] >= 0)[0])              # once each


@pytest.mark.parametrize("numbering", ["structured", "rcm"])
@pytest.mark.parametrize("nparts", [2, 3])
def test_owned_results_equal_the_undivided_run(nparts, numbering):
    m = TriMesh(10, 8)
    if numbering == "rcm":
        m = ro.reorder(m, "rcm")
    k = 3
    f = nabla2_fields(m, k, 21)
    want = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k, {n: v.copy() for n, v in f.items()})
    for p in pt.partition(m, nparts, rings=1):
        g = {n: p.to_local(LOC_OF[n], v, axis=1).copy() for n, v in f.items()}
        got = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), p.mesh, k, g)
        for n in ("rot_vec", "div_vec", "nabla2_vec"):
            loc = LOC_OF[n]
            no = p.n_owned[loc]
            v = p.mesh.valid(loc)[:no]
            assert bit_equal(got[n][:, :no][:, v], want[n][:, p.global_ids[loc][:no]][:, v]), (p.rank, n)
        s = p.splitters()
        assert s[(EDGE, pt.INTERIOR, 0)] == 0 and s[(EDGE, pt.HALO, 0)] == p.n_owned[EDGE]
        assert s[(EDGE, pt.END, 0)] == p.mesh.num_edges


def test_halo_refresh_moves_the_owners_values():
    m, k = TriMesh(8, 9), 2
    parts = pt.partition(m, 3)
    glob = np.random.default_rng(22).uniform(0.5, 1.5, (k, m.num_edges))
    local = []
    for p in parts:
        a = p.to_local(EDGE, glob).copy()
        a[:, p.n_owned[EDGE]:] = np.nan                      # halo copies are stale
        local.append(a)
    for p in parts:                                          # what Part.exchange does, without processes
        for q, ids in p.recv.get(EDGE, {}).items():
            src = parts[q].send[EDGE][p.rank]
            assert len(src) == len(ids)
            local[p.rank][:, ids] = local[q][:, src]
    for p in parts:
        refreshed = np.isfinite(local[p.rank])
        expected = np.ones(p.mesh.num_edges, dtype=bool)
        expected[p.n_owned[EDGE]:] = p.mesh.valid(EDGE)[p.n_owned[EDGE]:]   # invalid slots have no owner
        assert (refreshed.all(axis=0) == expected).all()
        ok = refreshed.all(axis=0)
        assert bit_equal(local[p.rank][:, ok], glob[:, p.global_ids[EDGE][ok]])


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"]); sys.path.insert(0, os.path.join(os.environ["REPO"], "tests"))
from util import iir_path, bit_equal
from oracle import naive_interp
from dawn_b200 import partition as pt, reorder as ro
from dawn_b200.trimesh import CELL, EDGE, VERTEX, TriMesh
from test_partition import LOC_OF, nabla2_fields
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
m, k = ro.reorder(TriMesh(9, 8), "rcm"), 3
f = nabla2_fields(m, k, 23)
want = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k, {n: v.copy() for n, v in f.items()})
p = pt.partition(m, world)[rank]
g = {n: p.to_local(LOC_OF[n], v, axis=1).copy() for n, v in f.items()}
vec = torch.from_numpy(g["vec"])
vec[:, p.n_owned[EDGE]:] = float("nan")                      # stale halo, refreshed from the owners
p.exchange(EDGE, vec, dist)
got = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), p.mesh, k, g)
ok = True
for n in ("rot_vec", "div_vec", "nabla2_vec"):
    loc = LOC_OF[n]; no = p.n_owned[loc]; v = p.mesh.valid(loc)[:no]
    ok = ok and bit_equal(got[n][:, :no][:, v], want[n][:, p.global_ids[loc][:no]][:, v])
res = torch.tensor([1 if ok else 0]); dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0: print("PARTS_OK" if int(res) == 1 else "PARTS_MISMATCH")
dist.destroy_process_group()
"""


def test_two_rank_partitioned_nabla2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert "PARTS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nparts", [2, 3])
def test_nabla4_runs_on_partitions_with_two_halo_rings(nparts):
    """ICON_nabla4_stencil = nabla2(nabla2(vec)) reaches two neighbourhoods deep: with rings=2 the halo holds
    everything the owned results depend on, so the whole stencil runs on the local mesh with only the INPUT
    `vec` refreshed - no exchange between its two passes ('or recompute halo redundantly to
    need only the input').  One ring is not enough."""
    from test_unstructured_oracle import _nabla_inputs, nabla4_oracle
    m, k = ro.reorder(TriMesh(12, 10), "rcm"), 3
    f = _nabla_inputs(m, k, 43)
    f["nabla2_vec"] = np.full((k, m.num_edges), -1.0)
    f["nabla4_vec"] = np.full((k, m.num_edges), -1.0)
    want = nabla4_oracle(m, k, f)
    loc = {"geofac_rot": VERTEX, "geofac_div": CELL, "__tmp_rot_0": VERTEX, "__tmp_rot_1": VERTEX,
           "__tmp_div_0": CELL, "__tmp_div_1": CELL}
    for rings, expect in ((2, True), (1, False)):
        same = True
        for p in pt.partition(m, nparts, rings=rings):
            g = {n: p.to_local(loc.get(n, EDGE), f.get(n, np.zeros_like(v)), axis=1).copy() for n, v in want.items()}
            got = naive_interp.run_unstructured(iir_path("ICON_nabla4_stencil"), p.mesh, k, g)
            no = p.n_owned[EDGE]
            v = p.mesh.valid(EDGE)[:no]
            for n in ("nabla2_vec", "nabla4_vec"):
                same = same and bit_equal(got[n][:, :no][:, v], want[n][:, p.global_ids[EDGE][:no]][:, v])
        assert same == expect, rings
