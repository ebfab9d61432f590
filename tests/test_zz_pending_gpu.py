"""GPU parity cases whose FIRST GPU run is still pending: they were written after round 1's GPU minutes were
spent.  Their logic is already checked on the CPU (oracle known answers; kernel emulation,
tests/test_kernel_emulation.py):
  * the seven stencils only the reference's Atlas suite runs (GenerateUnstructuredStencils.cpp:874-1134),
    ICON's diamond nabla2 + Smagorinsky, ICON_gradient, generate_versioned_field (dawn4py-tests);
  * ICON nabla2 on renumbered meshes (dawn_b200/reorder.py);
  * the reference's own pre-generated CUDA texts recompiled for sm_100a (oracle/_ref/libdawn_refcuda*.so)
    against ours, the oracle and the goldens of the reference's naive code.
The cases run in a child process started by one non-strict xfail test (an XPASS is the expected outcome):
a device fault in a kernel nobody has run yet cannot reach the main test process; the file is collected
last so that nothing else runs behind it."""
import numpy as np
import pytest

from util import bit_equal, iir_path
from oracle import naive_interp
from dawn_b200.trimesh import CELL, EDGE, TriMesh
from test_unstructured_gpu import _run

import os
import subprocess
import sys

INPROC = os.environ.get("B200_PENDING_INPROC") == "1"
pytestmark = [pytest.mark.gpu]
# the cases below only run inside the child process that test_pending_cases_in_a_child_process starts:
# a device fault in a kernel nobody has run yet must not take the CUDA context of the main test
# process (or the process itself) with it
inproc = pytest.mark.skipif(not INPROC, reason="runs inside the child process of test_pending_cases_...")


@pytest.mark.xfail(strict=False, reason="first GPU run pending (authored without GPU budget)")
@pytest.mark.skipif(INPROC, reason="this is the child")
def test_pending_cases_in_a_child_process():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"], env=dict(os.environ, B200_PENDING_INPROC="1"),
                       capture_output=True, text=True, timeout=180)
    print(r.stdout[-3000:], r.stderr[-1500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1500:]


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


@inproc
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


@inproc
@pytest.mark.parametrize("method", ["random", "rcm", "morton"])
def test_icon_nabla2_on_a_renumbered_mesh(method):
    """Same kernels, renumbered neighbour tables (dawn_b200/reorder.py): the result mapped back to the
    original numbering equals the oracle's on the original mesh bit for bit."""
    from dawn_b200 import reorder as ro
    from dawn_b200.trimesh import VERTEX
    m, k = TriMesh(21, 13), 9
    rng = np.random.default_rng(52)
    loc_of = {"vec": EDGE, "div_vec": CELL, "rot_vec": VERTEX, "nabla2_vec": EDGE, "primal_edge_length": EDGE,
              "dual_edge_length": EDGE, "tangent_orientation": EDGE, "geofac_rot": VERTEX, "geofac_div": CELL}
    f = {n: rng.uniform(0.5, 1.5, (k, m.num(loc))) for n, loc in loc_of.items() if not n.startswith("geofac")}
    f["geofac_rot"] = rng.uniform(-1, 1, (k, m.num_vertices, 6))
    f["geofac_div"] = rng.uniform(-1, 1, (k, m.num_cells, 3))
    want = {n: v.copy() for n, v in f.items()}
    want["__tmp_nab_60"] = np.zeros((k, m.num_edges))
    want["__tmp_nab_61"] = np.zeros((k, m.num_edges))
    naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k, want)
    rm = ro.ReorderedMesh(m, ro.Renumbering.random(m, 53)) if method == "random" else ro.reorder(m, method)
    r = rm.renumbering
    got, _ = _run("ICON_laplacian_stencil", rm, k, {n: np.ascontiguousarray(r.to_new(loc_of[n], v, axis=1))
                                                    for n, v in f.items()})
    for n, loc in (("rot_vec", VERTEX), ("div_vec", CELL), ("nabla2_vec", EDGE)):
        v = m.valid(loc)
        assert bit_equal(r.to_old(loc, got[n], axis=1)[:, v], want[n][:, v]), n


@inproc
@pytest.mark.parametrize("nx,ny,k", [(9, 7, 3), (33, 20, 9)])
def test_icon_diamond_laplacian(nx, ny, k):
    """ICON_laplacian_diamond_stencil.py (diamond nabla2 + Smagorinsky): 13 fused edge stages, one launch."""
    from test_unstructured_oracle import diamond_fields
    m = TriMesh(nx, ny)
    f = diamond_fields(m, k, 54)
    want = naive_interp.run_unstructured(iir_path("ICON_laplacian_diamond_stencil"), m, k,
                                         {n: v.copy() for n, v in f.items()})
    got, st = _run("ICON_laplacian_diamond_stencil", m, k, f)
    v = m.valid(EDGE)
    for n in ("vn_vert", "dvt_tang", "dvt_norm", "kh_smag_1", "kh_smag_2", "kh_smag", "nabla2"):
        assert bit_equal(got[n][:, v], want[n][:, v]), n
    assert st.launches() == 1


@inproc
def test_icon_gradient():
    """dawn4py-tests/ICON_gradient.py: sparse API field with centre filled by a loop, then reduced."""
    m, k = TriMesh(14, 9), 5
    rng = np.random.default_rng(55)
    f = {"p_grad": rng.uniform(0.5, 1.5, (k, m.num_cells)), "p_ccpr": rng.uniform(0.5, 1.5, (k, m.num_cells)),
         "geofac_grg": np.zeros((k, m.num_cells, 4))}
    want = naive_interp.run_unstructured(iir_path("ICON_gradient"), m, k, {n: v.copy() for n, v in f.items()})
    got, _ = _run("ICON_gradient", m, k, f)
    assert bit_equal(got["p_grad"], want["p_grad"])
    t = m.nbh_table([CELL, EDGE, CELL], True)
    assert bit_equal(got["geofac_grg"][:, t >= 0], want["geofac_grg"][:, t >= 0])


@inproc
def test_versioned_field_against_the_reference_golden():
    """generate_versioned_field: golden made by the reference's own generated text (oracle/_ref build)."""
    import os
    from test_unstructured_oracle import versioned_fields
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "versioned_field.npz"))
    nx, ny, k = (int(x) for x in g["dims"])
    m = TriMesh(nx, ny)
    f = versioned_fields(m, k, int(g["seed"]))
    f.pop("c_0")
    got, _ = _run("generate_versioned_field", m, k, f)
    ev = m.edge_valid
    assert bit_equal(got["a"][:, ev], g["a"][:, ev]) and bit_equal(got["c"][:, ev], g["c"][:, ev])


@inproc
@pytest.mark.parametrize("I,J,K", [(12, 12, 10), (70, 41, 9), (128, 64, 20)])
def test_reference_cuda_kernel_and_ours_agree_bit_for_bit(I, J, K):
    """hori_diff (dawn4py flavour: in, out, coeff): the reference's own pre-generated CUDA code, recompiled
    unmodified for sm_100a (oracle/_ref/libdawn_refcuda.so), against the b200 kernel and the oracle."""
    import ctypes
    import torch
    import dawn_b200
    from oracle import capi
    from util import HALO, run_oracle
    lib = capi.refcuda()
    if lib is None:
        pytest.skip("oracle/_ref/libdawn_refcuda.so not built")
    ni, nj = I + 2 * HALO, J + 2 * HALO
    rng = np.random.default_rng(56)
    f = {n: rng.uniform(0.5, 1.5, (K + 1, nj, ni)) for n in ("in", "out", "coeff")}
    f["out"] = np.full((K + 1, nj, ni), -1.0)
    want = run_oracle("hori_diff_stencil", (ni, nj, K), f)["out"]
    mod = dawn_b200.load_stencil("hori_diff_stencil")
    assert [fd["name"] for fd in mod.fields] == ["in", "out", "coeff"]
    dom = dawn_b200.Domain(ni, nj, K)
    dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
    ours = [dawn_b200.Storage(ni, nj, K + 1, "ijk").from_numpy(f[n]) for n in ("in", "out", "coeff")]
    mod(dom).run(*ours)
    theirs = [dawn_b200.Storage(ni, nj, K + 1, "ijk").from_numpy(f[n]) for n in ("in", "out", "coeff")]
    sec = ctypes.c_double()
    rc = lib.refcuda_hori_diff(ni, nj, K, HALO, theirs[0].stride_j, theirs[0].stride_k,
                               *[ctypes.c_void_p(s.buf.data_ptr() + 8 * s.lead) for s in theirs], 1,
                               ctypes.byref(sec))
    torch.cuda.synchronize()
    assert rc == 0
    a, b = ours[1].to_numpy(), theirs[1].to_numpy()
    assert bit_equal(a, want) and bit_equal(b, want)


@inproc
@pytest.mark.parametrize("case", ["conditional_var1_1", "conditional_var1_0", "global_indexing",
                                  "nonoverlapping", "copy"])
def test_reference_cuda_texts_reproduce_the_reference_naive_goldens(case):
    """The reference's pre-generated CUDA texts (CodeGen/reference/*.cu, copy_stencil_ref.cpp), recompiled
    unmodified for sm_100a, against the goldens its own generated NAIVE code produced (tests/golden):
    reference naive == reference CUDA on a B200 == ours (tests/test_cartesian_gpu.py)."""
    import ctypes
    import os
    import torch
    import dawn_b200
    from oracle import capi
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    name = {"conditional_var1_1": "conditional_stencil", "conditional_var1_0": "conditional_stencil",
            "nonoverlapping": "nonoverlapping_stencil", "copy": "copy_stencil"}.get(case, case)
    lib = capi.refcuda_small(name)
    if lib is None:
        pytest.skip("oracle/_ref/libdawn_refcuda_%s.so not built" % name)
    g = np.load(os.path.join(gold, "conditional_stencil.npz" if name == "conditional_stencil" else "indexing.npz"))
    ni, nj, nk = (int(x) for x in g["dims"])
    inn = np.random.default_rng(int(g["seed"])).uniform(-1, 1, (nk + 1, nj, ni))
    s_in = dawn_b200.Storage(ni, nj, nk + 1).from_numpy(inn)
    s_out = dawn_b200.Storage(ni, nj, nk + 1).fill(-1.0)
    var1 = 0 if case.endswith("_0") else 1
    rc = lib.refcuda_run(ni, nj, nk, 3, s_in.stride_j, s_in.stride_k,
                         ctypes.c_void_p(s_in.buf.data_ptr() + 8 * s_in.lead),
                         ctypes.c_void_p(s_out.buf.data_ptr() + 8 * s_out.lead), var1)
    torch.cuda.synchronize()
    assert rc == 0
    out = s_out.to_numpy()
    if name == "conditional_stencil":
        assert bit_equal(out, g["out_var1_%d" % var1])
    elif name == "global_indexing":
        assert bit_equal(out, g["global_indexing"])
    elif name == "nonoverlapping_stencil":
        assert bit_equal(out[11:], g["nonoverlapping_from_k11"])   # below: an uninitialised local in the text
    else:
        h = 3
        want = np.full_like(inn, -1.0)
        want[:nk, h:-h, h:-h] = inn[:nk, h:-h, h:-h]
        assert bit_equal(out, want)


@inproc
def test_reference_cuda_update_dz_c_reproduces_the_reference_naive_golden():
    """update_dz_c.cu (7 kernels, block-private temporaries in global memory, a versioned field), recompiled
    unmodified for sm_100a, against the golden of the reference's generated naive code."""
    import ctypes
    import os
    import torch
    from oracle import capi
    lib = capi.refcuda_small("update_dz_c")
    if lib is None:
        pytest.skip("oracle/_ref/libdawn_refcuda_update_dz_c.so not built")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "update_dz_c.npz"))
    ni, nj, nk = (int(x) for x in g["dims"])
    names = ["dp_ref", "zs", "area", "ut", "vt", "gz", "gz_x", "gz_y", "ws3"]
    rng = np.random.default_rng(int(g["seed"]))
    f = {n: rng.uniform(0.5, 1.5, (nk + 1, nj, ni)) for n in names}
    dev = [torch.from_numpy(f[n]).cuda().contiguous() for n in names]
    ptrs = (ctypes.c_void_p * 9)(*[t.data_ptr() for t in dev])
    rc = lib.refcuda_update_dz_c(ni, nj, nk, 3, float(g["dt"]), ptrs)
    torch.cuda.synchronize()
    assert rc == 0
    got = {n: t.cpu().numpy() for n, t in zip(names, dev)}
    assert bit_equal(got["gz"], g["gz"]) and bit_equal(got["ws3"], g["ws3"])
    for n in names:
        if n not in ("gz", "ws3"):
            assert bit_equal(got[n], f[n]), n
