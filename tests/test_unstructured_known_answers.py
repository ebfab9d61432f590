"""Known-answer tests of the reference's toylib integration suite
(dawn/test/integration-test/unstructured/ToylibIntegrationTestCompareOutput.cpp:62-760) run through
the unstructured oracle: they pin its reduction nesting, sparse indexing, neighbour-loop and
vertical-indirection semantics to the values the reference's own tests expect."""
import numpy as np
import pytest

from util import iir_path
from oracle import naive_interp
from dawn_b200.trimesh import CELL, EDGE, VERTEX, TriMesh

EPS = 1e3 * np.finfo(float).eps


def full(m, loc, k, value):
    return np.full((k, m.num(loc)), float(value))


def sparse(m, chain, k, value):
    return np.full((k, m.num(chain[0]), m.nbh_table(chain).shape[1]), float(value))


def run(name, m, k, fields, **kw):
    return naive_interp.run_unstructured(iir_path(name), m, k, fields, **kw)


def test_copy_cell_and_edge():  # :62-100
    m = TriMesh(32, 32)
    for name, loc in (("copyCell", CELL), ("copyEdge", EDGE)):
        f = run(name, m, 1, {"in_field": full(m, loc, 1, 1.0), "out_field": full(m, loc, 1, -1.0)})
        assert np.all(f["out_field"][:, m.valid(loc)] == 1.0)


def test_vertical_sum():  # :102-124
    m = TriMesh(32, 32)
    k = 10
    f = run("verticalSum", m, k, {"in_field": full(m, CELL, k, 10.0), "out_field": full(m, CELL, k, -1.0)})
    assert np.all(f["out_field"][1:k - 1] == 20.0)
    assert np.all(f["out_field"][[0, k - 1]] == -1.0)


def test_accumulate_edge_to_cell():  # :126-145
    m = TriMesh(32, 32)
    f = run("accumulateEdgeToCell", m, 1, {"in_field": full(m, EDGE, 1, 1.0), "out_field": full(m, CELL, 1, 0.0)})
    assert np.all(np.abs(f["out_field"] - 3.0) < EPS)


def test_vertical_solver():  # :343-376
    m = TriMesh(5, 5)
    k = 5
    lev = np.arange(k, dtype=float)[:, None]
    a = full(m, CELL, k, 0) + lev + 1
    b = full(m, CELL, k, 0) + lev + 1
    c = full(m, CELL, k, 0) + lev + 2
    d = full(m, CELL, k, 0) + np.array([5.0, 15, 31, 53, 45])[:, None]
    f = run("tridiagonalSolve", m, k, {"a": a, "b": b, "c": c, "d": d})
    assert np.all(np.abs(f["d"] - (lev + 1)) < EPS)


def test_nested_reductions():  # :381-429
    m = TriMesh(10, 10)
    f = {"cell_field": full(m, CELL, 1, 0), "edge_field": full(m, EDGE, 1, 0), "vertex_field": full(m, VERTEX, 1, 1)}
    assert np.all(np.abs(run("nestedSimple", m, 1, f)["cell_field"] - 6.0) < EPS)
    f = {"cell_field": full(m, CELL, 1, 0), "edge_field": full(m, EDGE, 1, 200), "vertex_field": full(m, VERTEX, 1, 1)}
    assert np.all(np.abs(run("nestedWithField", m, 1, f)["cell_field"] - 606.0) < EPS)


@pytest.mark.parametrize("name", ["sparseDimension", "sparseDimensionTwice"])
def test_sparse_dimensions(name):  # :431-455, :696-720
    m = TriMesh(10, 10)
    f = {"cell_field": full(m, CELL, 1, 0), "edge_field": full(m, EDGE, 1, 1),
         "sparse_dim": sparse(m, [CELL, EDGE], 1, 200)}
    assert np.all(np.abs(run(name, m, 1, f)["cell_field"] - 600.0) < EPS)


def test_nested_reduce_sparse_dimensions():  # :457-491
    m = TriMesh(10, 10)
    f = {"cell_field": full(m, CELL, 1, 0), "edge_field": full(m, EDGE, 1, 1), "vertex_field": full(m, VERTEX, 1, 2),
         "ce_sparse": sparse(m, [CELL, EDGE], 1, 200), "ev_sparse": sparse(m, [EDGE, VERTEX], 1, 300)}
    assert np.all(np.abs(run("nestedWithSparse", m, 1, f)["cell_field"] - 4200.0) < EPS)


def _valid_entries(m, chain):
    t = m.nbh_table(chain)
    return (t >= 0) & m.valid(chain[0])[:, None]


def test_sparse_assignment_0_1_2():  # :493-594
    m = TriMesh(10, 10)
    ecv = [EDGE, CELL, VERTEX]
    ok = _valid_entries(m, ecv)
    f = {"vn": sparse(m, ecv, 1, -1), "uVert": full(m, VERTEX, 1, 1), "vVert": full(m, VERTEX, 1, 2),
         "nx": full(m, VERTEX, 1, 3), "ny": full(m, VERTEX, 1, 4)}
    assert np.all(np.abs(run("sparseAssignment0", m, 1, f)["vn"][0][ok] - 11.0) < EPS)
    f = {"vn": sparse(m, ecv, 1, -1), "uVert": full(m, VERTEX, 1, 3), "vVert": full(m, VERTEX, 1, 4),
         "nx": sparse(m, ecv, 1, 1), "ny": sparse(m, ecv, 1, 2)}
    assert np.all(np.abs(run("sparseAssignment1", m, 1, f)["vn"][0][ok] - 11.0) < EPS)
    f = {"sparse": sparse(m, ecv, 1, -7), "e": full(m, EDGE, 1, 1), "v": full(m, VERTEX, 1, 2)}
    got = run("sparseAssignment2", m, 1, f)["sparse"][0]
    assert np.all(np.abs(got[ok] - (-2.0)) < EPS)
    assert np.all(got[~ok & m.valid(EDGE)[:, None]] == -7.0)  # missing neighbours are never written
    assert ok.sum(axis=1).max() == 4 and ok[m.valid(EDGE)].sum(axis=1).min() == 3


def test_sparse_assignment_3():  # :596-633
    m = TriMesh(10, 10)
    chain = [CELL, EDGE, CELL, EDGE, CELL]
    ok = _valid_entries(m, chain)
    B = np.arange(m.num_cells, dtype=float)[None, :]
    f = {"sparse": sparse(m, chain, 1, 1), "A": full(m, CELL, 1, 1), "B": B.copy()}
    got = run("sparseAssignment3", m, 1, f)["sparse"][0]
    want = np.broadcast_to(1.0 - B[0][:, None], got.shape)
    assert np.all(np.abs(got[ok] - want[ok]) < EPS)


def test_sparse_assignment_4_5():  # :635-694
    m = TriMesh(10, 10)
    f = {"sparse": sparse(m, [CELL, EDGE], 1, 0), "v": full(m, VERTEX, 1, 1)}
    assert np.all(np.abs(run("sparseAssignment4", m, 1, f)["sparse"] - 2.0) < EPS)
    f = {"sparse": sparse(m, [CELL, EDGE], 1, 0), "v": full(m, VERTEX, 1, 3), "c": full(m, CELL, 1, 2)}
    got = run("sparseAssignment5", m, 1, f)["sparse"][0]
    ce = m.nbh_table([CELL, EDGE])
    ev = m.nbh_table([EDGE, VERTEX])
    cells_of_vertex = (m.nbh_table([VERTEX, CELL]) >= 0).sum(axis=1)
    for c in range(m.num_cells):
        for s in range(3):
            e = ce[c, s]
            want = 6.0 * cells_of_vertex[ev[e, 0]] + 6.0 * cells_of_vertex[ev[e, 1]]
            assert abs(got[c, s] - want) < EPS * 100


def test_vertical_indirection():  # :722-748
    m = TriMesh(10, 10)
    k = 10
    lev = np.arange(k, dtype=float)[:, None]
    f = {"in": full(m, CELL, k, 0) + lev, "out": full(m, CELL, k, -1), "kidx": full(m, CELL, k, 0) + lev + 1}
    f = run("verticalIndirecion", m, k, f)
    assert np.all(f["kidx"] == lev + 1)
    assert np.all(f["out"][:k - 1] == lev[:k - 1] + 1)
    assert np.all(f["out"][k - 1] == -1.0)


def test_horizontal_and_vertical_only_fields():  # GenerateUnstructuredStencils.cpp:781-817
    m = TriMesh(9, 7)
    k = 4
    rng = np.random.default_rng(5)
    ec = [EDGE, CELL]
    f = {"horizontal": rng.uniform(0, 1, m.num_edges), "vertical": rng.uniform(0, 1, k),
         "horizontal_sparse": rng.uniform(0, 1, (m.num_edges, 2)),
         "full": rng.uniform(0, 1, (k, m.num_edges)), "out1": np.zeros((k, m.num_edges)),
         "out2": np.zeros((k, m.num_edges))}
    want1 = (f["full"] + f["horizontal"][None, :]) + f["vertical"][:, None]
    t = m.nbh_table(ec)
    want2 = np.zeros(m.num_edges)
    for s in range(2):
        want2 = np.where(t[:, s] >= 0, want2 + f["horizontal_sparse"][:, s], want2)
    got = run("horizontalVertical", m, k, f)
    v = m.valid(EDGE)
    assert np.array_equal(got["out1"][:, v], want1[:, v])
    assert np.array_equal(got["out2"][:, v], np.broadcast_to(want2, (k, m.num_edges))[:, v])


def test_unstructured_iteration_space():  # GenerateUnstructuredStencils.cpp:847-872
    m = TriMesh(8, 8)
    k = 2
    n = m.num_cells
    f = {"out": full(m, CELL, k, 0), "in_1": full(m, CELL, k, 1), "in_2": full(m, CELL, k, 2)}
    # Interior [10, 60), Halo [60, 100), End = 100 (driver-includes/unstructured_domain.hpp)
    spl = {(CELL, 2000, 0): 10, (CELL, 3000, 0): 60, (CELL, 4000, 0): 100}
    got = run("iterationSpaceUnstructured", m, k, f, splitters=spl)["out"]
    want = np.zeros(n)
    want[60:100] = 1
    want[10:60] = 2
    assert np.array_equal(got, np.broadcast_to(want, (k, n)))


# ---- the stencils only the reference's Atlas suite runs (AtlasIntegrationTestCompareOutput.cpp:1081-1370):
# their expected values are closed forms in the number of interior edges of a cell, so they hold on the
# toylib-numbered TriMesh as well ------------------------------------------------------------------
def _interior_edges_per_cell(m):
    ce, ec = m.nbh_table([CELL, EDGE]), m.nbh_table([EDGE, CELL])
    boundary = (ec < 0).any(axis=1)
    return boundary, np.array([sum(1 for e in row if e >= 0 and not boundary[e]) for row in ce], dtype=float)


def test_global_var():  # :1082-1105 (dt = 2 set at run time over the default 0.5)
    m, k = TriMesh(10, 10), 10
    f = run("globalVar", m, k, {"in_field": full(m, CELL, k, 1.0), "out_field": full(m, CELL, k, -1.0)},
            globals_={"dt": 2.0})
    assert np.all(f["out_field"] == 2.0)
    f = run("globalVar", m, k, {"in_field": full(m, CELL, k, 1.0), "out_field": full(m, CELL, k, -1.0)})
    assert np.all(f["out_field"] == 0.5)


def test_temp_field_allocation():  # :1111-1131 and :1137-1158
    m, k = TriMesh(10, 10), 10
    f = run("tempFieldAllocation", m, k, {"in_field": full(m, CELL, k, 1.0), "out_field": full(m, CELL, k, -1.0),
                                          "tmp_field": full(m, CELL, k, 0.0)})
    assert np.all(f["out_field"] == 2.0)
    f = run("sparseTempFieldAllocation", m, k,
            {"in_field": full(m, CELL, k, 1.0), "out_field": full(m, CELL, k, -1.0),
             "tmp_field": sparse(m, [CELL, EDGE], k, 0.0)})
    assert np.all(f["out_field"] == 3.0)


def test_reduction_in_if_conditional():  # :1165-1216
    m, k = TriMesh(10, 10), 1
    boundary, n_int = _interior_edges_per_cell(m)
    e = np.where(boundary, 0.0, 1.0)[None, :]
    f = run("reductionInIfConditional", m, k,
            {"c_field": full(m, CELL, k, 1.0), "sparse_field": sparse(m, [CELL, EDGE], k, 1.0),
             "e_field": e.copy(), "out_field": full(m, CELL, k, -1.0)})
    assert np.array_equal(f["out_field"][0], np.where(n_int < 3, 6.0, 12.0))
    assert set(np.unique(f["out_field"])) == {6.0, 12.0}   # both branches are taken on this mesh


@pytest.mark.parametrize("name,scale", [("reductionWithCenter", 1.0), ("reductionWithCenterSparse", 2.0),
                                        ("reductionAndFillWithCenterSparse", 2.0)])
def test_reductions_with_center(name, scale):  # :1222-1370: value = scale * (#cell neighbours + 1)
    m, k = TriMesh(10, 10), 1
    _, n_int = _interior_edges_per_cell(m)
    f = {"cin_field": full(m, CELL, k, 1.0), "cout_field": full(m, CELL, k, 0.0)}
    if name != "reductionWithCenter":
        size = m.nbh_table([CELL, EDGE, CELL], True).shape[1]
        assert size == 4                                    # CEC_SIZE + 1
        f["sparse"] = np.full((k, m.num_cells, size), 2.0 if name == "reductionWithCenterSparse" else 0.0)
    f = run(name, m, k, f)
    assert np.array_equal(f["cout_field"][0], scale * n_int + scale)


def test_icon_gradient():  # dawn4py-tests/ICON_gradient.py: p_grad = sum over (self + cell neighbours) of 2 * p_ccpr
    m, k = TriMesh(9, 6), 2
    rng = np.random.default_rng(9)
    p = rng.uniform(0.5, 1.5, (k, m.num_cells))
    f = run("ICON_gradient", m, k, {"p_grad": full(m, CELL, k, -1.0), "p_ccpr": p.copy(),
                                    "geofac_grg": np.zeros((k, m.num_cells, 4))})
    t = m.nbh_table([CELL, EDGE, CELL], True)
    want = np.zeros((k, m.num_cells))
    for i in range(4):                                       # chain order, centre first
        want = want + np.where(t[:, i] >= 0, 2.0 * p[:, np.maximum(t[:, i], 0)], 0.0)
    assert np.array_equal(f["p_grad"], want)
    assert np.all(f["geofac_grg"][:, t >= 0] == 2.0) and np.all(f["geofac_grg"][:, t < 0] == 0.0)
