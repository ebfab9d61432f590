"""GPU parity of the seven stencils only the reference's Atlas suite runs (globalVar,
tempFieldAllocation, sparseTempFieldAllocation, reductionInIfConditional, reductionWithCenter,
reductionWithCenterSparse, reductionAndFillWithCenterSparse; GenerateUnstructuredStencils.cpp:874-1134).
Their oracle side is pinned on the reference's expected values (tests/test_unstructured_known_answers.py).

These were authored after the round's GPU budget was spent, so their first GPU run is still pending:
they are non-strict xfail until a run has been seen (an XPASS is the expected outcome), and the file is
collected last so that nothing else runs behind them."""
import numpy as np
import pytest

from util import bit_equal, iir_path
from oracle import naive_interp
from dawn_b200.trimesh import CELL, EDGE, TriMesh
from test_unstructured_gpu import _run

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first GPU run pending (authored without GPU budget)")]


def _fields(name, m, k, rng):
    nc, ne = m.num_cells, m.num_edges
    cell = lambda: rng.uniform(0.5, 1.5, (k, nc))  # noqa: E731
    if name == "globalVar":
        return {"in_field": cell(), "out_field": cell()}, {"out_field": CELL}, {}
    if name == "tempFieldAllocation":
        return {"in_field": cell(), "out_field": cell()}, {"out_field": CELL}, {"tmp_field": np.zeros((k, nc))}
    if name == "sparseTempFieldAllocation":
        return ({"in_field": cell(), "out_field": cell()}, {"out_field": CELL},
                {"tmp_field": np.zeros((k, nc, 3))})
    if name == "reductionInIfConditional":
        return ({"c_field": cell(), "sparse_field": rng.uniform(0.5, 1.5, (k, nc, 3)),
                 "e_field": rng.uniform(0.5, 1.5, (k, ne)), "out_field": cell()}, {"out_field": CELL}, {})
    f = {"cin_field": cell(), "cout_field": cell()}
    if name == "reductionWithCenterSparse":
        f["sparse"] = rng.uniform(0.5, 1.5, (k, nc, 4))
    extra = {"sparse": np.zeros((k, nc, 4))} if name == "reductionAndFillWithCenterSparse" else {}
    return f, {"cout_field": CELL}, extra


@pytest.mark.parametrize("nx,ny,k", [(10, 10, 3), (13, 7, 11)])
@pytest.mark.parametrize("name", ["globalVar", "tempFieldAllocation", "sparseTempFieldAllocation",
                                  "reductionInIfConditional", "reductionWithCenter",
                                  "reductionWithCenterSparse", "reductionAndFillWithCenterSparse"])
def test_atlas_suite_stencils(name, nx, ny, k):
    m = TriMesh(nx, ny)
    f, outs, temporaries = _fields(name, m, k, np.random.default_rng(51))
    g = {"dt": 2.0} if name == "globalVar" else None
    want = {n: v.copy() for n, v in f.items()}
    want.update(temporaries)
    naive_interp.run_unstructured(iir_path(name), m, k, want, globals_=g)
    got, _ = _run(name, m, k, f, globals_=g)
    for n, loc in outs.items():
        v = m.valid(loc)
        assert bit_equal(got[n][:, v], want[n][:, v]), n
