"""CPU tests: the unstructured oracle (IIR interpreter + TriMesh) against the reference's own toylib
and naive-ico stencils compiled from /root/reference (oracle/_ref), bit for bit."""
import ctypes
import os

import numpy as np
import pytest

from util import bit_equal, iir_path
from oracle import capi, naive_interp
from dawn_b200.trimesh import CELL, EDGE, VERTEX, TriMesh

R = capi.ref()
needs_ref = pytest.mark.skipif(R is None, reason="oracle/_ref not built")
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)
P = lambda a: a.ctypes.data_as(dp)  # noqa: E731

CHAINS = [[CELL, EDGE], [CELL, VERTEX], [EDGE, CELL], [EDGE, VERTEX], [VERTEX, CELL], [VERTEX, EDGE],
          [EDGE, CELL, VERTEX], [CELL, EDGE, CELL], [EDGE, CELL, EDGE], [VERTEX, EDGE, CELL],
          [CELL, EDGE, CELL, EDGE, CELL], [EDGE, VERTEX, EDGE], [CELL, VERTEX, CELL],
          [VERTEX, EDGE, VERTEX], [EDGE, CELL, VERTEX, CELL]]


@needs_ref
@pytest.mark.parametrize("nx,ny", [(4, 3), (1, 1), (2, 7), (6, 6)])
def test_trimesh_tables_are_bit_exact_with_toylib(nx, ny):
    m = TriMesh(nx, ny)
    nv, ne, nc, nvalid = (ctypes.c_int() for _ in range(4))
    R.ref_toylib_sizes(nx, ny, ctypes.byref(nv), ctypes.byref(ne), ctypes.byref(nc), ctypes.byref(nvalid))
    assert (m.num_vertices, m.num_edges, m.num_cells) == (nv.value, ne.value, nc.value)
    assert int(m.edge_valid.sum()) == nvalid.value
    ids = np.zeros(ne.value, dtype=np.int32)
    n = R.ref_toylib_valid_edges(nx, ny, ids.ctypes.data_as(ip))
    assert np.array_equal(ids[:n], np.nonzero(m.edge_valid)[0])
    for chain in CHAINS:
        want = np.zeros((m.num(chain[0]), 40), dtype=np.int32)
        ch = (ctypes.c_int * len(chain))(*chain)
        assert R.ref_toylib_nbh_table(nx, ny, ch, len(chain), 40, want.ctypes.data_as(ip)) >= 0
        assert np.array_equal(m.nbh_table(chain, nbh_size=40), want), chain


@needs_ref
def test_trimesh_vertex_coordinates():
    m = TriMesh(5, 4, 2.0, 2.0, True)
    x, y = np.zeros(m.num_vertices), np.zeros(m.num_vertices)
    R.ref_toylib_vertex_xy(5, 4, ctypes.c_double(2.0), ctypes.c_double(2.0), 1, P(x), P(y))
    assert bit_equal(x, m.vertex_x) and bit_equal(y, m.vertex_y)


def test_neighbour_set_known_answers():
    # dawn/test/unit-test/interface/TestToylibInterface.cpp:53-186: diamond = 4 vertices, star/fan = 12,
    # intp = 9 on interior elements
    m = TriMesh(10, 10)
    def count(chain): return int((m.nbh_table(chain) >= 0).sum(axis=1).max())
    assert count([EDGE, CELL, VERTEX]) == 4
    assert count([CELL, EDGE, CELL, EDGE, CELL]) == 9
    assert count([VERTEX, CELL, EDGE]) == 12       # star
    assert count([VERTEX, EDGE, CELL]) == 6


def _fields(rng, m, k, spec):
    out = {}
    for name, loc in spec.items():
        out[name] = rng.uniform(0.5, 1.5, (k, m.num(loc)))
    return out


@needs_ref
def test_interpreter_matches_reference_unstructured_stencils():
    nx, ny, k = 7, 5, 3
    m = TriMesh(nx, ny)
    rng = np.random.default_rng(21)
    ev = m.edge_valid
    # diffusion
    f = _fields(rng, m, k, {"in_field": CELL, "out_field": CELL})
    want = f["out_field"].copy()
    R.ref_unstr_diffusion(nx, ny, k, P(f["in_field"]), P(want))
    got = naive_interp.run_unstructured(iir_path("diffusion"), m, k, {n: v.copy() for n, v in f.items()})
    assert bit_equal(got["out_field"], want)
    # gradient (edge field then cell field, both in/out)
    f = _fields(rng, m, k, {"cell_field": CELL, "edge_field": EDGE})
    wc, we = f["cell_field"].copy(), f["edge_field"].copy()
    R.ref_unstr_gradient(nx, ny, k, P(wc), P(we))
    got = naive_interp.run_unstructured(iir_path("gradient"), m, k, {n: v.copy() for n, v in f.items()})
    assert bit_equal(got["cell_field"], wc) and bit_equal(got["edge_field"][:, ev], we[:, ev])
    # diamond
    f = _fields(rng, m, k, {"edge_field": EDGE, "vertex_field": VERTEX})
    we = f["edge_field"].copy()
    R.ref_unstr_diamond(nx, ny, k, P(we), P(f["vertex_field"]))
    got = naive_interp.run_unstructured(iir_path("diamond"), m, k, {n: v.copy() for n, v in f.items()})
    assert bit_equal(got["edge_field"][:, ev], we[:, ev])
    # diamondWeights
    f = _fields(rng, m, k, {"out": EDGE, "inv_edge_length": EDGE, "inv_vert_length": EDGE, "in": VERTEX})
    wo = f["out"].copy()
    R.ref_unstr_diamond_weights(nx, ny, k, P(wo), P(f["inv_edge_length"]), P(f["inv_vert_length"]), P(f["in"]))
    got = naive_interp.run_unstructured(iir_path("diamondWeights"), m, k, {n: v.copy() for n, v in f.items()})
    assert bit_equal(got["out"][:, ev], wo[:, ev])
    # intp
    f = _fields(rng, m, k, {"in": CELL, "out": CELL})
    wo = f["out"].copy()
    R.ref_unstr_intp(nx, ny, k, P(f["in"]), P(wo))
    got = naive_interp.run_unstructured(iir_path("intp"), m, k, {n: v.copy() for n, v in f.items()})
    assert bit_equal(got["out"], wo)


def icon_fields(rng, m, k):
    f = _fields(rng, m, k, {"vec": EDGE, "div_vec": CELL, "rot_vec": VERTEX, "nabla2_vec": EDGE,
                            "primal_edge_length": EDGE, "dual_edge_length": EDGE,
                            "tangent_orientation": EDGE})
    f["geofac_rot"] = rng.uniform(-1, 1, (k, m.num_vertices, 6))
    f["geofac_div"] = rng.uniform(-1, 1, (k, m.num_cells, 3))
    f["__tmp_nab_60"] = np.zeros((k, m.num_edges))
    f["__tmp_nab_61"] = np.zeros((k, m.num_edges))
    return f


@needs_ref
def test_interpreter_matches_reference_icon_laplacian():
    nx, ny, k = 8, 6, 4
    m = TriMesh(nx, ny)
    rng = np.random.default_rng(22)
    f = icon_fields(rng, m, k)
    w = {n: f[n].copy() for n in ("div_vec", "rot_vec", "nabla2_vec")}
    R.ref_icon_laplacian(nx, ny, k, P(f["vec"]), P(w["div_vec"]), P(w["rot_vec"]), P(w["nabla2_vec"]),
                         P(f["primal_edge_length"]), P(f["dual_edge_length"]),
                         P(f["tangent_orientation"]), P(f["geofac_rot"]), P(f["geofac_div"]))
    got = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k,
                                        {n: v.copy() for n, v in f.items()})
    ev = m.edge_valid
    assert bit_equal(got["div_vec"], w["div_vec"]) and bit_equal(got["rot_vec"], w["rot_vec"])
    assert bit_equal(got["nabla2_vec"][:, ev], w["nabla2_vec"][:, ev])
    # golden for machines without oracle/_ref
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "icon_laplacian.npz")
    if not os.path.exists(path):
        np.savez_compressed(path, seed=22, dims=(nx, ny, k), **w)


def test_interpreter_matches_icon_laplacian_golden():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "icon_laplacian.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    g = np.load(path)
    nx, ny, k = (int(x) for x in g["dims"])
    m = TriMesh(nx, ny)
    f = icon_fields(np.random.default_rng(int(g["seed"])), m, k)
    got = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k, f)
    ev = m.edge_valid
    assert bit_equal(got["div_vec"], g["div_vec"]) and bit_equal(got["rot_vec"], g["rot_vec"])
    assert bit_equal(got["nabla2_vec"][:, ev], g["nabla2_vec"][:, ev])


def _nabla_inputs(m, k, seed):
    rng = np.random.default_rng(seed)
    f = {n: rng.uniform(0.5, 1.5, (k, m.num(loc))) for n, loc in (
        ("vec", EDGE), ("primal_edge_length", EDGE), ("dual_edge_length", EDGE),
        ("tangent_orientation", EDGE))}
    f["geofac_rot"] = rng.uniform(-1, 1, (k, m.num_vertices, 6))
    f["geofac_div"] = rng.uniform(-1, 1, (k, m.num_cells, 3))
    return f


def nabla4_by_two_nabla2_passes(m, k, f):
    """nabla4 = nabla2(nabla2(vec)) evaluated with the reference-pinned ICON_laplacian_stencil oracle
    applied twice: the anchor for ICON_nabla4_stencil, which the reference itself does not have."""
    def one_pass(vec):
        w = {n: v.copy() for n, v in f.items()}
        w["vec"] = vec.copy()
        w["rot_vec"] = np.zeros((k, m.num_vertices))
        w["div_vec"] = np.zeros((k, m.num_cells))
        for n in ("nabla2_vec", "__tmp_nab_60", "__tmp_nab_61"):
            w[n] = np.zeros((k, m.num_edges))
        naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k, w)
        return w["nabla2_vec"]
    n2 = one_pass(f["vec"])
    # edge slots without a valid edge hold whatever the caller left there; the second pass only ever
    # gathers valid edges, so zeros are as good as anything
    return n2, one_pass(n2)


def nabla4_oracle(m, k, f):
    w = {n: v.copy() for n, v in f.items()}
    for n in ("nabla2_vec", "nabla4_vec", "__tmp_nab_60", "__tmp_nab_61", "__tmp_nab_62", "__tmp_nab_63"):
        w[n] = np.zeros((k, m.num_edges))
    for n in ("__tmp_rot_0", "__tmp_rot_1"):
        w[n] = np.zeros((k, m.num_vertices))
    for n in ("__tmp_div_0", "__tmp_div_1"):
        w[n] = np.zeros((k, m.num_cells))
    naive_interp.run_unstructured(iir_path("ICON_nabla4_stencil"), m, k, w)
    return w


@pytest.mark.parametrize("nx,ny,k", [(7, 5, 3), (16, 12, 4)])
def test_nabla4_is_nabla2_applied_twice(nx, ny, k):
    m = TriMesh(nx, ny)
    f = _nabla_inputs(m, k, 41)
    n2, n4 = nabla4_by_two_nabla2_passes(m, k, f)
    w = nabla4_oracle(m, k, f)
    v = m.valid(EDGE)
    assert bit_equal(w["nabla2_vec"][:, v], n2[:, v])
    assert bit_equal(w["nabla4_vec"][:, v], n4[:, v])


@needs_ref
def test_timed_reference_entry_counts_valid_edges():
    # bench.py's ICON cpu_baseline / reference arm: unit of work = valid edges x levels
    r = capi.icon_laplacian_reference_timed(8, 6, 3, 2)
    assert r is not None
    ne, t = r
    assert ne == int(TriMesh(8, 6).edge_valid.sum()) and len(t) == 2 and (t > 0).all()


def diamond_fields(m, k, seed):
    rng = np.random.default_rng(seed)
    f = {n: rng.uniform(0.5, 1.5, (k, m.num_edges)) for n in (
        "diff_multfac_smag", "tangent_orientation", "inv_primal_edge_length", "inv_vert_vert_length", "vn",
        "dvt_tang", "dvt_norm", "kh_smag_1", "kh_smag_2", "kh_smag", "nabla2")}
    f.update({n: rng.uniform(0.5, 1.5, (k, m.num_vertices)) for n in ("u_vert", "v_vert")})
    f.update({n: rng.uniform(-1, 1, (k, m.num_edges, 4)) for n in (
        "primal_normal_x", "primal_normal_y", "dual_normal_x", "dual_normal_y", "vn_vert")})
    return f


def test_icon_diamond_laplacian_against_a_direct_restatement():
    # ICON_laplacian_diamond_stencil.py:37-245 restated with whole-array numpy on the edges whose diamond
    # is complete (4 vertices); weighted reductions accumulate init + w0*x0 + w1*x1 + ... in chain order
    m, k = TriMesh(11, 8), 3
    f = diamond_fields(m, k, 44)
    w = naive_interp.run_unstructured(iir_path("ICON_laplacian_diamond_stencil"), m, k,
                                      {n: v.copy() for n, v in f.items()})
    t = m.nbh_table([EDGE, CELL, VERTEX])
    full = (t >= 0).all(axis=1)
    nb = t[full]                                                     # [edges, 4]
    g = lambda n: f[n][:, full]                                      # noqa: E731
    uv = lambda n: f[n][:, nb]                                       # noqa: E731  [k, edges, 4]

    def reduce(x, weights):
        acc = np.zeros(x.shape[:2])
        for i in range(4):
            acc = acc + weights[i] * x[:, :, i]
        return acc
    vn_vert = uv("u_vert") * g("primal_normal_x") + uv("v_vert") * g("primal_normal_y")
    dual = uv("u_vert") * g("dual_normal_x") + uv("v_vert") * g("dual_normal_y")
    first, second = [-1.0, 1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 1.0]
    tang, ipel, ivvl = g("tangent_orientation"), g("inv_primal_edge_length"), g("inv_vert_vert_length")
    dvt_tang = reduce(dual, first) * tang
    dvt_norm = reduce(dual, second)
    k1 = ((reduce(vn_vert, first) * tang) * ipel) + (dvt_norm * ivvl)
    k1 = k1 * k1
    k2 = (reduce(vn_vert, second) * ivvl) + (dvt_tang * ipel)
    k2 = k2 * k2
    kh = g("diff_multfac_smag") * np.sqrt(k1 + k2)
    nabla2 = reduce(4.0 * vn_vert, [ipel * ipel, ipel * ipel, ivvl * ivvl, ivvl * ivvl])
    nabla2 = nabla2 - (((8.0 * g("vn")) * (ipel * ipel)) + ((8.0 * g("vn")) * (ivvl * ivvl)))
    for name, want in (("vn_vert", vn_vert), ("dvt_tang", dvt_tang), ("dvt_norm", dvt_norm), ("kh_smag_1", k1),
                       ("kh_smag_2", k2), ("kh_smag", kh), ("nabla2", nabla2)):
        assert bit_equal(w[name][:, full], want), name


def versioned_fields(m, k, seed):
    rng = np.random.default_rng(seed)
    f = {n: rng.uniform(0.5, 1.5, (k, m.num_edges)) for n in "abc"}
    f["d"] = (rng.uniform(0, 1, (k, m.num_edges)) < 0.4).astype(float)        # conditions: zero / non-zero
    f["e"] = (rng.uniform(0, 1, (k, m.num_edges)) < 0.5).astype(float) * 3.0
    f["c_0"] = np.zeros((k, m.num_edges))
    return f


@needs_ref
def test_interpreter_matches_reference_versioned_field():
    # the reference's own generated text data/generate_versioned_field_ref.cpp, compiled unmodified
    nx, ny, k = 7, 5, 4
    m = TriMesh(nx, ny)
    f = versioned_fields(m, k, 3)
    w = {n: f[n].copy() for n in "ac"}
    R.ref_generate_versioned_field(nx, ny, k, P(w["a"]), P(f["b"]), P(w["c"]), P(f["d"]), P(f["e"]))
    got = naive_interp.run_unstructured(iir_path("generate_versioned_field"), m, k, {n: v.copy() for n, v in f.items()})
    ev = m.edge_valid
    assert bit_equal(got["a"][:, ev], w["a"][:, ev]) and bit_equal(got["c"][:, ev], w["c"][:, ev])
    changed = (got["c"][:, ev] != f["c"][:, ev]).mean()
    assert 0.1 < changed < 0.6 and 0.2 < (got["a"][:, ev] == f["b"][:, ev]).mean() < 0.7   # every branch taken
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "versioned_field.npz")
    if not os.path.exists(path):
        np.savez_compressed(path, seed=3, dims=(nx, ny, k), a=w["a"], c=w["c"])


def test_interpreter_matches_versioned_field_golden():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "versioned_field.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    g = np.load(path)
    nx, ny, k = (int(x) for x in g["dims"])
    m = TriMesh(nx, ny)
    got = naive_interp.run_unstructured(iir_path("generate_versioned_field"), m, k,
                                        versioned_fields(m, k, int(g["seed"])))
    ev = m.edge_valid
    assert bit_equal(got["a"][:, ev], g["a"][:, ev]) and bit_equal(got["c"][:, ev], g["c"][:, ev])
