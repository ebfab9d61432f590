"""CPU tests of the emitter: backend selection contract, planner decisions, derived information,
error behaviour (the reference's CodeGen tests are golden-string tests without execution,
dawn/test/unit-test/dawn/CodeGen/Stencils.cpp:112-123; structural assertions play that role here)."""
import json
import os
import re

import pytest

from util import iir_path
import dawn_b200
from dawn_b200 import codegen

REF = "/root/reference/dawn/test/"


def _code(name, **opts):
    return codegen.compile_iir(open(iir_path(name)).read(), **opts)


def test_parse_backend_string_contract():
    # dawn::codegen::parseBackendString, CodeGen/Driver.cpp:47-64 + the two new values
    P = codegen.parse_backend_string
    B = codegen.CodeGenBackend
    assert P("cuda") == B.CUDA and P("CUDA") == B.CUDA
    assert P("cuda-ico") == B.CUDAIco and P("CUDAIco") == B.CUDAIco
    assert P("naive") == B.CXXNaive and P("c++-naive") == B.CXXNaive and P("cxxnaive") == B.CXXNaive
    assert P("ico") == B.CXXNaiveIco and P("naive-ico") == B.CXXNaiveIco
    assert P("gt") == B.GridTools and P("gridtools") == B.GridTools
    assert P("b200") == B.B200 and P("B200") == B.B200
    assert P("b200-ico") == B.B200Ico and P("B200Ico") == B.B200Ico
    with pytest.raises(ValueError, match="Backend not supported"):
        P("opencl")


def test_only_b200_backends_are_provided():
    with pytest.raises(codegen.CodeGenError, match="not provided"):
        _code("copy_stencil", backend="cuda")


def test_hori_diff_plan_inlines_lap_and_stages_inputs_with_tma():
    code = _code("hori_diff_type2_stencil")
    kernels = re.findall(r"^// kernel (\S+): (\w+) kernel", code, flags=re.M)
    assert [k[1] for k in kernels] == ["tile", "tile"]
    assert kernels[1][0].endswith("_tma_kernel")
    assert "lap_ijcache" not in code and "__shared__ ::dawn::float_type lap" not in code
    assert "const ::dawn::float_type lap_p0_p0" in code       # lap re-evaluated in registers
    assert "CUtensorMap tm_in" in code and "CUtensorMap tm_hdmask" in code
    assert "tma::load_3d" in code and "tma::wait" in code
    # no temporary storage, one launch per run
    assert "scratch_field" not in code
    # the reference's class API is kept
    for s in ("class hori_diff_type2_stencil", "void run(storage_ijk_t out, storage_ijk_t in, "
              "storage_j_t crlato, storage_j_t crlatu, storage_ijk_t hdmask)", "get_total_time",
              "reset_meters", "get_name", "in_extent = {-2, 2, -2, 2, 0, 0}"):
        assert s in code, s


def test_temporaries_fall_back_to_scratch_when_inlining_is_off():
    code = _code("hori_diff_type2_stencil", inline_temporaries=0)
    assert "scratch_field m_lap" in code and "__syncthreads();" in code
    assert "stage" in code and "extent i[-1,1] j[-1,1]" in code  # redundant halo computation


def test_tridiagonal_plan_fuses_sweeps_with_register_kcaches():
    code = _code("tridiagonal_solve")
    kernels = re.findall(r"^// kernel (\S+): (\w+) kernel over (\d+) multistage", code, flags=re.M)
    assert all(k[1] == "column" and k[2] == "2" for k in kernels)
    assert any(k[0].endswith("_ring_kernel") for k in kernels)
    assert "c_kc0" in code and "c_kc1" in code and "d_kc1" in code
    assert "fence_proxy_async" in code  # backward sweep streams what the forward sweep stored
    unfused = _code("tridiagonal_solve", fuse_columns=0, column_ring=0)
    assert len(re.findall(r"^// kernel ", unfused, flags=re.M)) == 2
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).replace('\\"', '"'))
    assert info["written"] == ["d", "c"]
    assert [f["extent"] for f in info["fields"]] == [[0, 0, 0, 0, -1, 1], [0] * 6, [0] * 6,
                                                     [0, 0, 0, 0, -1, 0]]


def test_planner_knobs_change_the_tile():
    code = _code("hori_diff_type2_stencil", block_size_i=32, block_size_j=8, rows_per_thread=4)
    assert "block 32x8 threads, 4 row(s)/thread" in code and "__launch_bounds__(128)" in code
    with pytest.raises(codegen.CodeGenError, match="block size"):
        _code("hori_diff_type2_stencil", block_size_i=64, block_size_j=64)
    with pytest.raises(TypeError):
        _code("hori_diff_type2_stencil", no_such_option=1)


def test_globals_and_masked_fields():
    code = _code("laplacian_stencil")
    assert "struct globals" in code and "double dx;" in code and "dx(1)" in code
    assert "double get_dx()" in code and "void set_dx(double dx)" in code
    assert "storage_ij_t out_field, storage_ij_t in_field" in code
    assert "g.dx" in code


def test_malformed_iir_is_an_error_not_a_crash():
    with pytest.raises(codegen.CodeGenError):
        codegen.compile_iir("{not json")
    with pytest.raises(codegen.CodeGenError, match="missing key"):
        codegen.compile_iir("{}")
    iir = json.load(open(iir_path("copy_stencil")))
    stmt = iir["internalIR"]["stencils"][0]["multiStages"][0]["stages"][0]["doMethods"][0]["ast"][
        "block_stmt"]["statements"][0]
    stmt["expr_stmt"]["expr"]["assignment_expr"]["right"] = {"stencil_fun_call_expr": {"callee": "f"}}
    with pytest.raises(codegen.CodeGenError, match="inlined"):
        codegen.compile_iir(json.dumps(iir))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_iir_fixtures_are_accepted():
    fixtures = {
        "unit-test/dawn/CodeGen/input/update_dz_c.iir": 7,
        "unit-test/dawn/CodeGen/input/conditional_stencil.iir": 1,
        "integration-test/serializer/reference_iir/copy_stencil.iir": 1,
        "integration-test/serializer/reference_iir/lap_stencil.iir": 1,
        "unit-test/dawn/Optimizer/input/tridiagonal_solve.iir": 1,
    }
    for rel, nk in fixtures.items():
        code = codegen.compile_iir(open(REF + rel).read(), use_tma=0, column_ring=0)
        assert len(re.findall(r"^// kernel ", code, flags=re.M)) == nk, rel
    code = codegen.compile_iir(open(REF + "unit-test/dawn/CodeGen/input/update_dz_c.iir").read())
    # extents exported like reference/update_dz_c.cpp:80-89
    assert "dp_ref_extent = {0, 1, 0, 1, -2, 1}" in code and "gz_x_extent = {-1, 1, 0, 1, 0, 0}" in code
    assert "gz_0_extent = {0, 0, 0, 0, 0, 1}" in code
    assert "gridtools::dawn::storage_t m_gz_0" in code  # allocated (versioned) field owned by wrapper


def test_cli_matches_library(tmp_path):
    import subprocess
    from dawn_b200 import build
    build.build_emitter()
    exe = os.path.join(build.LIB, "dawn-b200-codegen")
    out = tmp_path / "x.cu"
    subprocess.check_call([exe, "--backend=b200", "-o", str(out), iir_path("copy_stencil")])
    assert out.read_text() == _code("copy_stencil")
    r = subprocess.run([exe, "-b", "nope", iir_path("copy_stencil")], capture_output=True, text=True)
    assert r.returncode == 1 and "Backend not supported" in r.stderr


def test_ico_chain_sizes_known_answers():
    # dawn/test/unit-test/dawn/CodeGen/Cuda-Ico/TestCuda-IcoCodeGen.cpp:31-87 (C=0, E=1, V=2)
    C, E, V = 0, 1, 2
    tests = [([C, E], 3), ([C, V], 3), ([E, C], 2), ([E, V], 2), ([V, C], 6), ([V, E], 6),
             ([E, C, E], 4), ([E, C, V], 4), ([E, C, V, C], 16), ([C, V, C], 12), ([C, V, C, E], 24),
             ([E, V, E], 10), ([E, V, E, C], 10), ([V, E, C], 6), ([V, E, C, V], 6),
             ([V, E, C, V, E], 30), ([V, E, C, V, E, C], 24), ([C, E, C], 3), ([C, E, C, E], 9),
             ([C, E, C, E, C], 9), ([C, E, C, E, C, E], 21)]
    assert len(tests) == 21
    for chain, want in tests:
        assert codegen.ico_chain_size(chain) == want, chain
    assert codegen.ico_chain_size([C, C]) == -1  # invalid chain


def test_icon_laplacian_plan_fuses_edge_stages():
    code = _code("ICON_laplacian_stencil", backend="b200-ico", co_schedule=0)
    kernels = re.findall(r"^// kernel (\S+): (\d+) fused stage\(s\) on (\w+)", code, flags=re.M)
    assert [(k[1], k[2]) for k in kernels] == [("1", "vertices"), ("1", "cells"), ("5", "edges")]
    # the independent vertex and cell groups (both gather `vec`) share one launch
    code = _code("ICON_laplacian_stencil", backend="b200-ico")
    heads = re.findall(r"^// kernel .*$", code, flags=re.M)
    assert len(heads) == 2 and "on vertices co-scheduled with 1 stage(s) on cells" in heads[0]
    # ico_chain=1: the edge stages, which gather rot_vec / div_vec, are chained into that launch - producers
    # publish flags, consumers wait for them and read the handed-over fields with ld.global.cg
    code = _code("ICON_laplacian_stencil", backend="b200-ico", ico_chain=1)
    heads = re.findall(r"^// kernel .*$", code, flags=re.M)
    assert len(heads) == 1 and "chained with 5 stage(s) on edges" in heads[0]
    assert "chain::publish(&flags[bidx], epoch)" in code and "chain::wait(&flags[q], epoch, cv.error)" in code
    assert "__ldcg(&rot_vec_p[" in code and "__ldcg(&div_vec_p[" in code and "__ldg(&rot_vec_p[" not in code
    l1 = _code("ICON_laplacian_stencil", backend="b200-ico", ico_chain=1, ico_chain_l1=1)
    assert "chain::load(&rot_vec_p[" in l1 and "chain::load(&div_vec_p[" in l1
    assert "if(hi0 > lo0) { lo0 = lo0 > 0 ? lo0 - 1 : 0;" in l1      # whole cache lines final: L1 may keep them
    exact = _code("ICON_laplacian_stencil", backend="b200-ico", ico_chain=1, ico_cache_hints=1)
    assert "__ldcg(&rot_vec_p[" in exact and "__ldcg(&div_vec_p[" in exact and "lo0 - 1" not in exact
    assert "__ldcs(&geofac_rot_p[" in exact and "__stcs(&nabla2_vec_p[" in exact and "__ldg(&vec_p[" in exact
    assert "__tmp_nab_60_p" not in code and "scratch_field" not in code  # temporaries stay in registers
    assert "setup_ICON_laplacian_stencil(void** handle, const b200_mesh_t* mesh, int k_size" in code
    unfused = _code("ICON_laplacian_stencil", backend="b200-ico", fuse_columns=0, co_schedule=0)
    assert len(re.findall(r"^// kernel ", unfused, flags=re.M)) == 7
    assert "scratch_field m___tmp_nab_60" in unfused
    with pytest.raises(codegen.CodeGenError, match="unstructured"):
        _code("ICON_laplacian_stencil", backend="b200")
    with pytest.raises(codegen.CodeGenError, match="Cartesian"):
        _code("copy_stencil", backend="b200-ico")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_unstructured_fixtures_are_accepted():
    for rel in ("integration-test/serializer/reference_iir/unstructured_sum_edge_to_cells.iir",
                "integration-test/serializer/reference_iir/unstructured_mixed_copies.iir"):
        code = codegen.compile_iir(open(REF + rel).read(), backend="b200-ico")
        assert "__global__" in code


def test_unstructured_interface_files():
    iir = open(iir_path("ICON_laplacian_stencil")).read()
    h = codegen.interface_text(iir, "c")
    for sym in ("setup_ICON_laplacian_stencil", "run_ICON_laplacian_stencil_from_c_host",
                "verify_ICON_laplacian_stencil", "run_and_verify_ICON_laplacian_stencil",
                "free_ICON_laplacian_stencil"):
        assert sym in h
    # the header is valid C against the runtime header
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "icon.h"), "w").write(h)
        open(os.path.join(d, "t.c"), "w").write('#include "icon.h"\nint main(void){return 0;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(
            os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include"), "-I", d,
            os.path.join(d, "t.c")])
    f = codegen.interface_text(iir, "f90")
    assert "module ICON_laplacian_stencil_b200" in f and "bind(c)" in f
    with pytest.raises(codegen.CodeGenError):
        codegen.interface_text(open(iir_path("copy_stencil")).read(), "c")


def test_unstructured_loops_sparse_and_indirection_are_emitted():
    """Constructs of Cuda-ico/ASTStencilBody.cpp: neighbour loops writing sparse fields, nested
    reductions fetched where they are used, vertical indirection, horizontal-only fields, iteration
    spaces with splitter indices."""
    code = _code("sparseAssignment2", backend="b200-ico")
    assert code.count("if(nb_ecv_") == 4                      # loop over e->c->v unrolled 4 times
    assert "sparse_p[((size_t)k * 4 + 3) * n_edges + pidx]" in code
    nested = _code("sparseAssignment5", backend="b200-ico")
    assert "nb_ce_0" in nested and "nb_ev_0" not in nested    # only the loop's chain is preloaded
    assert "__ldg(&tbl_vc[" in nested
    ind = _code("verticalIndirecion", backend="b200-ico")
    assert "(int)(__ldg(&kidx_p[(size_t)k * n_cells + pidx])) + (0)" in ind
    hv = _code("horizontalVertical", backend="b200-ico")
    assert "horizontal_p[pidx]" in hv and "vertical_p[k]" in hv
    assert "horizontal_sparse_p[(size_t)0 * n_edges + pidx]" in hv
    sp = _code("iterationSpaceUnstructured", backend="b200-ico")
    assert "splitter(0, 3000, 0, &elem_lo)" in sp and "splitter(0, 4000, 0, &elem_hi)" in sp
    assert "set_splitter_index_iterationSpaceUnstructured" in sp
    assert len(re.findall(r"^// kernel ", sp, flags=re.M)) == 2   # different spaces are not fused
    with pytest.raises(codegen.CodeGenError, match="b200-ico"):
        _code("verticalIndirecion", backend="b200")


def test_control_flow_of_the_description_is_emitted():
    code = _code("stencil_desc_ast_04")
    run = code[code.index("  void run(storage_ijk_t in, storage_ijk_t out) {"):]
    run = run[:run.index("sync_storages(")]
    assert run.count("if(") == 3 and run.count(".run(in, out);") == 2
    assert "m_globals.var_compiletime != (int)1" in run
    plain = _code("hori_diff_type2_stencil")
    run = plain[plain.index("  void run(storage_ijk_t out, storage_ijk_t in, storage_j_t crlato"):]
    run = run[:run.index("m_stencil_23.run(")]
    # no control flow of the description; the only branch is the dispatch to the host <-> device pipeline
    assert run.count("if(") == 1 and "if(m_pipe_levels > 0 && (out.host_is_newer()" in run
    # the pipeline exists for stencils whose levels are independent, and only for them
    assert "void set_host_pipeline(int levels)" in plain and "run_pipelined(" in plain
    assert "set_host_pipeline" not in _code("tridiagonal_solve")
    assert "set_host_pipeline" not in code    # stencil_desc_ast_04: two stencil calls


def test_pure_locals_are_substituted_into_inlined_temporaries():
    code = _code("hd_smagorinsky_stencil")
    assert len(re.findall(r"^// kernel .*_kernel:", code, flags=re.M)) == 2   # plain + tma twin, one multistage
    assert "scratch_field" not in code and not re.search(r"\bT_sqr_s_p\b", code)
    staged = _code("hd_smagorinsky_stencil", inline_temporaries=0)
    assert "scratch_field" in staged


def test_racy_halo_accesses_to_stored_fields_are_rejected():
    """A stage with a horizontal extent runs redundantly over tile + extent in every block; a stored
    field it writes (or reads) must not be touched again (written) by a later stage of the same
    multistage - the reference's CUDA backend races silently on this, here it is an error."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from iir_builder import FULL, IIRBuilder
    b = IIRBuilder("racy_write")
    i, o = b.field("in"), b.field("out")
    t = b.tmp("t")
    s1 = b.stage(b.do(FULL, b.assign(t, i(1, 0)), b.assign(o, i())))     # extent from t(-1,0) below
    s2 = b.stage(b.do(FULL, b.assign(o, t(-1, 0) + t(1, 0))))
    b.stencil(b.multistage("Forward", s1, s2))                           # Forward: t cannot be inlined away
    with pytest.raises(codegen.CodeGenError, match="blocks would race"):
        codegen.compile_iir(b.dumps(), inline_temporaries=0)
    b = IIRBuilder("racy_read")
    i, o = b.field("in"), b.field("out")
    t = b.tmp("t")
    s1 = b.stage(b.do(FULL, b.assign(t, o() * 2.0)))                     # reads out in the halo ...
    s2 = b.stage(b.do(FULL, b.assign(o, t(-1, 0) + i())))                # ... which this stage rewrites
    b.stencil(b.multistage("Parallel", s1, s2))
    with pytest.raises(codegen.CodeGenError, match="version the field"):
        codegen.compile_iir(b.dumps(), inline_temporaries=0)


def test_write_after_read_at_an_offset_is_not_fused():
    """A later stage overwriting what an earlier stage read at ANOTHER point (horizontal offset /
    neighbour chain).  The naive backend runs stage after stage over the whole plane, so the earlier
    stage sees the old values everywhere.  b200-ico: the overwrite starts a new launch; Cartesian: the
    multistage is rejected (Dawn's field versioning removes such cycles before code generation)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from iir_builder import FULL, IIRBuilder
    code = codegen.compile_iir(open(iir_path("overwriteGathered")).read(), backend="b200-ico")
    groups = re.findall(r"(\d+) fused stage\(s\) on cells", code)
    assert len(re.findall(r"^__global__", code, flags=re.M)) == 2, groups
    b = IIRBuilder("war_cart")
    g, h, o = b.field("g"), b.field("h"), b.field("o")
    s1 = b.stage(b.do(FULL, b.assign(o, g(1, 0) + g(0, 1))))   # zero stage extent, reads g at offsets
    s2 = b.stage(b.do(FULL, b.assign(g, h() * 2.0)))
    b.stencil(b.multistage("Parallel", s1, s2))
    with pytest.raises(codegen.CodeGenError, match="at a horizontal offset and written by stage"):
        codegen.compile_iir(b.dumps())
    # pointwise read + later overwrite stays legal (every thread touches only its own point)
    b = IIRBuilder("war_ok")
    g, h, o = b.field("g"), b.field("h"), b.field("o")
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, g() + h(1, 0)))),
                           b.stage(b.do(FULL, b.assign(g, h() * 2.0)))))
    codegen.compile_iir(b.dumps())


def test_vertical_dependency_in_a_parallel_multistage_is_rejected():
    """Level blocks of a Parallel multistage run concurrently: reading, at another level, a field the
    multistage writes is a race (found by the random unstructured stencils) - diagnosed for both
    backends."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from iir_builder import FULL, IIRBuilder, Interval
    b = IIRBuilder("kdep", grid="Unstructured")
    i, t, o = (b.field(n, "k", "Edge") for n in ("in", "mid", "out"))
    s1 = b.stage(b.do(FULL, b.assign(t, i.at() * 2.0)), location="Edge")
    s2 = b.stage(b.do(Interval("Start", "End", 0, -1), b.assign(o, t.at(1))), location="Edge")
    b.stencil(b.multistage("Parallel", s1, s2))
    with pytest.raises(codegen.CodeGenError, match="read at a vertical offset"):
        codegen.compile_iir(b.dumps(), backend="b200-ico")
    b = IIRBuilder("kdep_cart")
    i, t, o = b.field("in"), b.field("mid"), b.field("out")
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(t, i() * 2.0))),
                           b.stage(b.do(Interval("Start", "End", 0, -1), b.assign(o, t(0, 0, 1))))))
    with pytest.raises(codegen.CodeGenError, match="read at a vertical offset"):
        codegen.compile_iir(b.dumps())


def test_one_translation_unit_for_several_instantiations_in_either_format():
    """codegen::run takes a map of instantiations (Driver.h:38-40): one TU, one include block, every
    stencil class; inputs may mix the Json and Byte formats."""
    import ctypes
    lib = codegen._load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    datas = [open(os.path.join(root, "stencils", "copy_stencil.iir"), "rb").read(),
             open(os.path.join(root, "tests", "golden", "proto", "lap.iirb"), "rb").read()]
    names = (ctypes.c_char_p * 2)(b"copy_stencil", b"lap")
    bufs = (ctypes.c_char_p * 2)(*datas)
    lens = (ctypes.c_size_t * 2)(*[len(d) for d in datas])
    out, n = ctypes.c_void_p(), ctypes.c_size_t()
    opts = codegen.default_options()
    rc = lib.dawn_b200_codegen_run(names, bufs, lens, 2, int(codegen.CodeGenBackend.B200),
                                   ctypes.byref(opts), ctypes.byref(out), ctypes.byref(n))
    assert rc == 0, lib.dawn_b200_codegen_last_error()
    code = ctypes.string_at(out.value, n.value).decode()
    lib.dawn_b200_free(out)
    assert re.findall(r"^class (\w+) \{", code, flags=re.M) == ["copy_stencil", "lap"]
    assert code.count("#include <driver-includes/gridtools_includes.hpp>") == 1
    assert "setup_copy_stencil" in code and "setup_lap" in code


def test_generated_c_header_is_valid_c(tmp_path):
    """--output-c-header twin: the header of an unstructured stencil must compile as plain C against
    include/dawn_b200_rt.h and declare every entry point the module exports."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = codegen.interface_text(open(iir_path("ICON_laplacian_stencil")).read(), "c")
    hdr = tmp_path / "ICON_laplacian_stencil.h"
    hdr.write_text(text)
    src = tmp_path / "use.c"
    src.write_text('#include "ICON_laplacian_stencil.h"\nint main(void) { return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"),
                    "-I", str(tmp_path), str(src)], check=True)
    for sym in ("setup_", "run_", "free_", "verify_", "run_and_verify_", "set_splitter_index_",
                "set_metrics_record_"):
        assert sym + "ICON_laplacian_stencil" in text, sym
    f90 = codegen.interface_text(open(iir_path("ICON_laplacian_stencil")).read(), "f90")
    assert "module ICON_laplacian_stencil_b200" in f90 and "bind(c)" in f90
    assert "set_splitter_index_ICON_laplacian_stencil" in f90


def test_cli_interface_files_and_conversion(tmp_path):
    """dawn-b200-codegen: --output-c-header / --output-f90-interface (cuda-ico's options) and --convert."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dawn_b200", "lib",
                       "dawn-b200-codegen")
    h, f90, cu = tmp_path / "s.h", tmp_path / "s.f90", tmp_path / "s.cu"
    subprocess.run([exe, "-b", "b200-ico", "--output-c-header=%s" % h, "--output-f90-interface=%s" % f90,
                    "-o", str(cu), iir_path("ICON_laplacian_stencil")], check=True)
    assert "setup_ICON_laplacian_stencil" in h.read_text() and "bind(c)" in f90.read_text()
    assert "__global__" in cu.read_text()
    b = tmp_path / "s.iirb"
    subprocess.run([exe, "--convert=byte", "-o", str(b), iir_path("hori_diff_type2_stencil")], check=True)
    assert b.read_bytes() == codegen.iir_convert(open(iir_path("hori_diff_type2_stencil")).read(), "byte")
    back = subprocess.run([exe, "--convert=json", str(b)], check=True, capture_output=True).stdout
    assert codegen.compile_iir(back.decode()) == codegen.compile_iir(open(iir_path("hori_diff_type2_stencil")).read())


def _ring_kernel(code):
    start = code.index("_ring_kernel(")
    return code[start:code.index("\n}\n", start)]


def test_column_ring_issues_its_first_prologue_at_kernel_start():
    # tridiagonal_solve: level 0 is its own k range (plain loads); the ring range [1, kmax] must have its
    # TMA prologue in flight before that range runs, and must not issue it a second time
    body = _ring_kernel(_code("tridiagonal_solve"))
    early = body.index("TMA prologue of k range [kmin + 1, kmax + 0], ahead of the ranges before it")
    assert early < body.index("{ // k in [kmin + 0, kmin + 0]") < body.index("{ // k in [kmin + 1, kmax + 0]")
    fwd = body[:body.index("(Backward)")]
    assert fwd.count("for(int c = 0; c < 2 && c < nch; ++c)") == 1     # 2 slots of 16 levels (planner default)
    assert "prologue already issued at the top of the sweep" in fwd
    # the backward sweep streams what the forward sweep stored: its prologue stays behind the proxy fence
    bwd = body[body.index("(Backward)"):]
    assert bwd.index("fence_proxy_async") < bwd.index("for(int c = 0; c < 2 && c < nch; ++c)")
    old = _ring_kernel(_code("tridiagonal_solve", ring_early=0))
    assert "ahead of the ranges before it" not in old
    assert old.index("{ // k in [kmin + 0, kmin + 0]") < old.index("for(int c = 0; c < 2 && c < nch; ++c)")


def test_ring_stagger_is_off_by_default():
    assert "stagger" not in _code("tridiagonal_solve")
    code = _code("tridiagonal_solve", ring_stagger=3000)
    assert "const int stagger_ns = 3000;" in code and "stagger_cycles, stagger_phases, stagger_blocks" in code


def test_icon_nabla4_plan():
    # nabla2 applied twice: (vertices | cells) co-scheduled + 5 fused edge stages, both twice; the second
    # vertex/cell launch gathers nabla2_vec, which the first edge kernel must therefore store
    code = _code("ICON_nabla4_stencil", backend="b200-ico", ico_chain=1)
    heads = re.findall(r"^// kernel .*$", code, flags=re.M)
    assert len(heads) == 2 and all("chained with 5 stage(s) on edges" in h for h in heads)
    code = _code("ICON_nabla4_stencil", backend="b200-ico")
    heads = re.findall(r"^// kernel .*$", code, flags=re.M)
    assert len(heads) == 4
    assert "co-scheduled" in heads[0] and "co-scheduled" in heads[2]
    assert "5 fused stage(s) on edges" in heads[1] and "5 fused stage(s) on edges" in heads[3]
    for t in ("__tmp_nab_60", "__tmp_nab_61", "__tmp_nab_62", "__tmp_nab_63"):
        assert t + "_p" not in code  # edge temporaries never leave registers


def test_icon_diamond_laplacian_is_one_kernel():
    # 13 edge stages (ICON_laplacian_diamond_stencil.py): everything but the sparse vn_vert and the API
    # outputs stays in registers, one launch
    code = _code("ICON_laplacian_diamond_stencil", backend="b200-ico")
    heads = re.findall(r"^// kernel .*$", code, flags=re.M)
    assert len(heads) == 1 and "13 fused stage(s) on edges" in heads[0]
    assert code.count("<<<") == 1 and "sqrt(" in code


def test_a_stage_that_gathers_what_it_writes_is_rejected():
    # `tempField` exactly as GenerateUnstructuredStencils.cpp:900-937 builds it (field versioning, no stage
    # splitter): its second stage writes `dense` and gathers it from the neighbours in the next statement - a race
    # for one thread per element (the reference's cuda-ico kernel included), order-dependent in the naive loops
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_stencils
    with pytest.raises(codegen.CodeGenError, match="gathers field 'dense' from other elements and writes it"):
        codegen.compile_iir(make_stencils.ico_temp_field(split=False).dumps(), backend="b200-ico")
    assert "tempField_stencil" in codegen.compile_iir(make_stencils.ico_temp_field().dumps(), backend="b200-ico")
    # the mirror image: a later statement overwrites what an earlier one of the same stage gathered
    b = make_stencils._ico("warInsideAStage")
    c, e = b.field("c", "k", "Cell"), b.field("e", "k", "Cell")
    chain = ["Cell", "Edge", "Cell"]
    b.stencil(b.multistage("Parallel", b.stage(b.do(make_stencils.FULL, b.assign(e, make_stencils.reduce_over(chain, c.nbh)),
                                                      b.assign(c, make_stencils.lit(1.0))), location="Cell")))
    with pytest.raises(codegen.CodeGenError, match="overwrites field 'c' that an earlier statement"):
        codegen.compile_iir(b.dumps(), backend="b200-ico")
