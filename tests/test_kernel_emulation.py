"""CPU emulation of EMITTED kernels (tests/cpu_emu): every unstructured stencil fixture - and, at the end of
this file, the plain (barrier-free) kernels of the Cartesian ones - is generated, compiled with g++ behind a thin CUDA stand-in and executed thread by thread on host memory
through the module's own C ABI, then compared with the oracle bit for bit.  This checks the logic of the
emitted code - indexing, neighbour order, stage fusion and co-scheduling, register forwarding, scratch
temporaries, splitter ranges, centre-including chains - on machines without a GPU; the GPU tests check the
same modules as the hardware runs them."""
import glob
import json
import os
import re
import sys

import numpy as np
import pytest

from util import ROOT, bit_equal, iir_path
from oracle import naive_interp
from dawn_b200 import codegen, reorder as ro
from dawn_b200.trimesh import CELL, EDGE, VERTEX, TriMesh

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_emu"))
import build as emu  # noqa: E402

LOC = {"cells": CELL, "edges": EDGE, "vertices": VERTEX}
UNSTRUCTURED = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "stencils", "*.iir"))
                      if '"gridType": "Unstructured"' in open(p).read())


def emitted(name, **opts):
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200-ico", **opts)
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    return code, info


def random_fields(info, m, k, rng):
    f = {}
    for fd in info["fields"]:
        shape = []
        if fd.get("k", 1):
            shape.append(k)
        if fd["location"] != "none":
            shape.append(m.num(LOC[fd["location"]]))
        if fd["sparse"]:
            shape.append(fd["sparse"])
        f[fd["name"]] = rng.uniform(0.5, 1.5, shape)
    return f


def with_temporaries(name, m, k, fields):
    """the oracle wants storage for the stencil's temporaries as well"""
    d = json.load(open(iir_path(name)))["metadata"]
    out = {n: v.copy() for n, v in fields.items()}
    for aid in d.get("temporaryFieldIDs", []):
        dims = d["fieldIDtoDimensions"][str(aid)]
        chain = dims["unstructured_horizontal_dimension"]["iter_space"]["chain"]
        shape = [k] if dims.get("mask_k", 1) else []
        shape.append(m.num(chain[0]))
        if len(chain) > 1:
            center = dims["unstructured_horizontal_dimension"]["iter_space"].get("include_center", False)
            shape.append(m.nbh_table(chain, center).shape[1])
        out.setdefault(d["accessIDToName"][str(aid)], np.zeros(shape))
    return out


def compare(info, m, got, want):
    for fd in info["fields"]:
        n = fd["name"]
        if fd["location"] == "none":
            assert bit_equal(got[n], want[n]), n
            continue
        v = m.valid(LOC[fd["location"]])
        axis = 1 if fd.get("k", 1) else 0
        assert bit_equal(np.compress(v, got[n], axis=axis), np.compress(v, want[n], axis=axis)), n


def test_the_fixture_list_is_what_we_think_it_is():
    assert len(UNSTRUCTURED) >= 50 and {"ICON_laplacian_stencil", "ICON_nabla4_stencil", "fuzzico_15",
                                        "ICON_laplacian_diamond_stencil", "reductionWithCenterSparse",
                                        "generate_versioned_field", "sparseAssignment5"} <= set(UNSTRUCTURED)


@pytest.mark.parametrize("name", UNSTRUCTURED)
def test_emitted_kernels_match_the_oracle_on_the_cpu(name):
    code, info = emitted(name)
    for (nx, ny, k) in ((9, 7, 5), (13, 10, 9)):
        m = TriMesh(nx, ny)
        rng = np.random.default_rng(300 + len(name))
        f = random_fields(info, m, k, rng)
        g, spl = None, None
        if name == "tridiagonalSolve":
            f["b"] += 4.0
        if name == "verticalIndirecion":
            f["kidx"] = np.zeros((k, m.num_cells)) + np.arange(k, dtype=float)[:, None] + 1
            f["kidx"][: k - 1] = rng.integers(0, k, (k - 1, m.num_cells)).astype(float)
        if name == "iterationSpaceUnstructured":
            spl = {(CELL, 2000, 0): 10, (CELL, 3000, 0): 60, (CELL, 4000, 0): 100}
        if name == "generate_versioned_field":
            f["d"] = (f["d"] < 0.9).astype(float)
            f["e"] = (f["e"] < 1.0).astype(float) * 3.0
        if info.get("globals"):
            g = {n: 0.625 for n in info["globals"]}
        want = naive_interp.run_unstructured(iir_path(name), m, k, with_temporaries(name, m, k, f),
                                             globals_=g, splitters=spl)
        got, launches = emu.run(name, code, info, m, k, f, globals_=g, splitters=spl)
        compare(info, m, got, want)
        assert launches == info["kernels"] or launches <= info["kernels"]


@pytest.mark.parametrize("opts", [dict(fuse_columns=0, levels_per_thread=1), dict(co_schedule=0),
                                  dict(levels_per_thread=3), dict(ico_chain=1), dict(ico_chain=1, ico_chain_lag=1),
                                  dict(ico_chain=1, ico_chain_lag=100000, co_schedule=0),
                                  dict(ico_chain=1, ico_chain_l1=1), dict(ico_chain=1, ico_cache_hints=1)])
@pytest.mark.parametrize("name", ["ICON_laplacian_stencil", "ICON_nabla4_stencil", "ICON_laplacian_diamond_stencil",
                                  "fuzzico_03", "fuzzico_11", "fuzzico_16", "fuzzico_17", "fuzzico_18", "fuzzico_19",
                                  "overwriteGathered", "sparseAssignment5", "nestedWithSparse"])
def test_emitter_variants_on_the_cpu(name, opts):
    code, info = emitted(name, **opts)
    m, k = TriMesh(11, 8), 7
    f = random_fields(info, m, k, np.random.default_rng(17))
    g = {n: 0.625 for n in info["globals"]} if info.get("globals") else None
    want = naive_interp.run_unstructured(iir_path(name), m, k, with_temporaries(name, m, k, f), globals_=g)
    got, _ = emu.run(name, code, info, m, k, f, globals_=g)
    compare(info, m, got, want)


# ---- ico_stage: own-element rows (dense fields and sparse weights of the kernel's location) arrive in shared
# memory by 1-D bulk copies; the stand-in copy insists on the hardware's 16-byte rules, so the lead-in arithmetic
# for rows that start at odd element indices (odd element counts; four phases in float) is what is tested here.
@pytest.mark.parametrize("opts", [dict(ico_stage=1), dict(ico_stage=2, co_schedule=0, ico_hoist=1),
                                  dict(ico_stage=1, single_precision=1),
                                  dict(ico_stage=2, single_precision=1, levels_per_thread=5)])
@pytest.mark.parametrize("name", ["ICON_laplacian_stencil", "ICON_laplacian_diamond_stencil", "nestedWithSparse",
                                  "fuzzico_16"])
def test_staged_rows_on_the_cpu(name, opts):
    code, info = emitted(name, **opts)
    assert "bulk_load_1d" in code and "stage_ok" in code
    for nx, ny in ((20, 14), (21, 15)):     # 315 vertices / 945 edges (odd), 352 / 1056 (even)
        m, k = TriMesh(nx, ny), 9
        f = random_fields(info, m, k, np.random.default_rng(23))
        w = with_temporaries(name, m, k, f)
        if opts.get("single_precision"):
            f = {n: v.astype(np.float32) for n, v in f.items()}
            w = {n: v.astype(np.float32) for n, v in with_temporaries(name, m, k, f).items()}
        want = naive_interp.run_unstructured(iir_path(name), m, k, w)
        got, _ = emu.run(name, code, info, m, k, f)
        assert emu.run.bulk_copies > 0      # full blocks took the staged variant
        compare(info, m, got, want)


# ---- ico_hoist: the read-only loads of a level move to its top (one constant each, guards kept)
@pytest.mark.parametrize("name", ["ICON_laplacian_stencil", "ICON_laplacian_diamond_stencil", "verticalIndirecion",
                                  "reductionInIfConditional", "tridiagonalSolve", "fuzzico_14", "fuzzico_18",
                                  "generate_versioned_field", "iterationSpaceUnstructured"])
def test_hoisted_loads_on_the_cpu(name):
    code, info = emitted(name, ico_hoist=1)
    assert "const ::dawn::float_type hl0 = " in code
    m, k = TriMesh(11, 8), 7
    rng = np.random.default_rng(29)
    f = random_fields(info, m, k, rng)
    spl = None
    if name == "tridiagonalSolve":
        f["b"] += 4.0
    if name == "verticalIndirecion":
        f["kidx"] = np.zeros((k, m.num_cells)) + np.arange(k, dtype=float)[:, None] + 1
        f["kidx"][: k - 1] = rng.integers(0, k, (k - 1, m.num_cells)).astype(float)
    if name == "iterationSpaceUnstructured":
        spl = {(CELL, 2000, 0): 10, (CELL, 3000, 0): 60, (CELL, 4000, 0): 100}
    if name == "generate_versioned_field":
        f["d"] = (f["d"] < 0.9).astype(float)
        f["e"] = (f["e"] < 1.0).astype(float) * 3.0
    want = naive_interp.run_unstructured(iir_path(name), m, k, with_temporaries(name, m, k, f), splitters=spl)
    got, _ = emu.run(name, code, info, m, k, f, splitters=spl)
    compare(info, m, got, want)


def test_hoisting_keeps_guards_and_vertical_offsets():
    code, _ = emitted("ICON_laplacian_stencil", ico_hoist=1)
    # a gather through a neighbour index that may be -1 keeps its guard; own-element loads need none
    assert "= nb_ev_0 >= 0 ? __ldg(&rot_vec_p[(size_t)k * n_vertices + nb_ev_0]) : (::dawn::float_type)0;" in code
    assert re.search(r"const ::dawn::float_type hl\d+ = __ldg\(&dual_edge_length_p\[\(size_t\)k \* n_edges \+ pidx\]\);", code)
    # loads at another level stay where the interval guards protect them
    seq, _ = emitted("verticalSum", ico_hoist=1)
    for line in seq.splitlines():
        if " hl" in line and "= " in line and "__ldg" in line:
            assert "(k + (" not in line


def test_staging_leaves_written_and_gathered_accesses_alone():
    # rows are only what a kernel reads at its own element and no role of the launch writes
    code, _ = emitted("ICON_laplacian_stencil", ico_stage=1)
    staged = set(re.findall(r"bulk_load_1d\(stg \+ [^,]+, (\w+)_p", code))
    assert staged == {"geofac_rot", "geofac_div", "tangent_orientation", "primal_edge_length", "dual_edge_length"}
    assert "ico_stage" not in emitted("ICON_laplacian_stencil")[0] and "b200_smem" not in emitted("ICON_laplacian_stencil")[0]
    # sequential sweeps and chained launches keep their plain form
    assert "bulk_load_1d" not in emitted("tridiagonalSolve", ico_stage=1)[0]
    assert "bulk_load_1d" not in emitted("ICON_laplacian_stencil", ico_stage=1, ico_chain=1)[0].split("_k2_")[0]


@pytest.mark.parametrize("method", ["random", "rcm", "morton"])
def test_renumbered_mesh_through_the_emitted_kernels(method):
    # same kernels, renumbered tables (dawn_b200/reorder.py): maps back to the undivided oracle run
    name = "ICON_laplacian_stencil"
    code, info = emitted(name)
    m, k = TriMesh(12, 9), 4
    f = random_fields(info, m, k, np.random.default_rng(18))
    want = naive_interp.run_unstructured(iir_path(name), m, k, with_temporaries(name, m, k, f))
    rm = ro.ReorderedMesh(m, ro.Renumbering.random(m, 19)) if method == "random" else ro.reorder(m, method)
    r = rm.renumbering
    loc_of = {fd["name"]: LOC[fd["location"]] for fd in info["fields"]}
    got, _ = emu.run(name, code, info, rm, k, {n: np.ascontiguousarray(r.to_new(loc_of[n], v, axis=1))
                                               for n, v in f.items()})
    for n in ("rot_vec", "div_vec", "nabla2_vec"):
        v = m.valid(loc_of[n])
        assert bit_equal(r.to_old(loc_of[n], got[n], axis=1)[:, v], want[n][:, v]), n


def test_the_emulation_catches_an_undersized_scratch():
    # regression for a defect this emulation found: a sparse temporary's scratch was sized like a dense one
    # ([k][element] instead of [k][nbh][element]) - on a GPU a silent out-of-bounds write
    name = "sparseTempFieldAllocation"
    code, info = emitted(name)
    assert "m_tmp_field.ensure(n_cells, n_cells * 3, k_size)" in code
    bad = code.replace("m_tmp_field.ensure(n_cells, n_cells * 3, k_size)", "m_tmp_field.ensure(0, n_cells, k_size)")
    m, k = TriMesh(6, 5), 2
    f = {"in_field": np.ones((k, m.num_cells)), "out_field": np.zeros((k, m.num_cells))}
    with pytest.raises(RuntimeError, match="outside a scratch allocation"):
        emu.run(name, bad, info, m, k, f)


# ---- Cartesian: the PLAIN kernels (no TMA staging) of every Cartesian fixture run on the CPU as well - thread
# by thread, or, where stages hand temporaries over through scratch behind __syncthreads(), with the threads of
# a block as host threads meeting at a barrier; the TMA / shared-memory variants are GPU-only ---------
from util import HALO, run_oracle  # noqa: E402

CARTESIAN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "stencils", "*.iir"))
                   if '"gridType": "Cartesian"' in open(p).read())
CART_GLOBALS = {"p_grad_c": {"hydrostatic": 1, "dt2": 0.5}, "laplacian_stencil": {"dx": 0.25},
                "globals_stencil": {"var_runtime": 1, "var_default": 2.0, "var_compiletime": 1},
                "stencil_desc_ast_03": {"var_runtime": 1, "var_compiletime": 2},
                "stencil_desc_ast_04": {"var_runtime": 1, "var_compiletime": 2},
                "stencil_desc_ast_07": {"var_runtime": 1, "var_compiletime": 2}}


def emitted_cartesian(name):
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200", use_tma=0, column_ring=0)
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    return code, info


@pytest.mark.parametrize("name", CARTESIAN)
def test_plain_cartesian_kernels_match_the_oracle_on_the_cpu(name):
    code, info = emitted_cartesian(name)   # kernels with __syncthreads() run their blocks as host threads
    for (I, J, K) in ((17, 11, 12), (29, 14, 13)):   # the interval fixtures name levels up to k_start + 11
        ni, nj = I + 2 * HALO, J + 2 * HALO
        rng = np.random.default_rng(400 + len(name))
        f = {}
        for fd in info["fields"]:
            d = fd["dims"]
            f[fd["name"]] = rng.uniform(0.5, 1.5, (K + 1 if "k" in d else 1, nj if "j" in d else 1,
                                                   ni if "i" in d else 1))
        if name in ("tridiagonal_solve", "kparallel_solver") and "b" in f:
            f["b"] += 4.0
        g = CART_GLOBALS.get(name)
        if g is None and info.get("globals"):
            g = {n: 0.375 for n in info["globals"]}
        want = run_oracle(name, (ni, nj, K), f, globals_=g)
        got, _ = emu.run_cartesian(name, code, info, (ni, nj, K), (HALO, HALO, HALO, HALO, 0, 0), f, globals_=g)
        for fd in info["fields"]:
            n = fd["name"]
            assert bit_equal(got[n].reshape(want[n].shape), want[n]), (name, n, (I, J, K))


@pytest.mark.parametrize("name", ["hori_diff_type2_stencil", "lap", "hd_smagorinsky_stencil", "hori_diff_stencil_02",
                                  "kcache_epflush", "fuzz_02", "fuzz_07", "fuzz_11"])
def test_staged_temporaries_variant_on_the_cpu(name):
    # inline_temporaries=0: temporaries go through global scratch, computed redundantly on each tile's stage
    # extent, consumers behind a block barrier (the general fallback of the planner)
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200", use_tma=0, column_ring=0,
                               inline_temporaries=0)
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    I, J, K = 37, 21, 12
    ni, nj = I + 2 * HALO, J + 2 * HALO
    rng = np.random.default_rng(23)
    f = {}
    for fd in info["fields"]:
        d = fd["dims"]
        f[fd["name"]] = rng.uniform(0.5, 1.5, (K + 1 if "k" in d else 1, nj if "j" in d else 1, ni if "i" in d else 1))
    g = {n: 0.375 for n in info["globals"]} if info.get("globals") else None
    want = run_oracle(name, (ni, nj, K), f, globals_=g)
    got, _ = emu.run_cartesian(name, code, info, (ni, nj, K), (HALO, HALO, HALO, HALO, 0, 0), f, globals_=g)
    for fd in info["fields"]:
        assert bit_equal(got[fd["name"]].reshape(want[fd["name"]].shape), want[fd["name"]]), fd["name"]


# ---- the DEFAULT kernels: TMA-staged tiles and column rings.  The emulation stands in for the hardware pieces
# (tests/cpu_emu/emu_cuda.hpp): a bulk tensor load copies its box at once with zero fill outside the tensor and
# completes its bytes on an mbarrier word, wait(parity) blocks on the phase, dynamic shared memory is a per-block
# buffer, the threads of a block are host threads.  What is checked is the emitted logic - slot arithmetic,
# phase parities, box origins and extents, micro-tile indexing, drain / refill order, prologue placement. --------
def _cartesian_case(name, code, info, size, seed, globals_=None):
    I, J, K = size
    ni, nj = I + 2 * HALO, J + 2 * HALO
    rng = np.random.default_rng(seed)
    f = {}
    for fd in info["fields"]:
        d = fd["dims"]
        f[fd["name"]] = rng.uniform(0.5, 1.5, (K + 1 if "k" in d else 1, nj if "j" in d else 1, ni if "i" in d else 1))
    if name in ("tridiagonal_solve", "kparallel_solver") and "b" in f:
        f["b"] += 4.0
    g = globals_ if globals_ is not None else CART_GLOBALS.get(name)
    if g is None and info.get("globals"):
        g = {n: 0.375 for n in info["globals"]}
    want = run_oracle(name, (ni, nj, K), f, globals_=g)
    got, _ = emu.run_cartesian(name, code, info, (ni, nj, K), (HALO, HALO, HALO, HALO, 0, 0), f, globals_=g)
    for fd in info["fields"]:
        n = fd["name"]
        assert bit_equal(got[n].reshape(want[n].shape), want[n]), (name, n, size)
    return emu.run_cartesian.tensor_maps


@pytest.mark.parametrize("name", CARTESIAN)
def test_default_cartesian_kernels_match_the_oracle_on_the_cpu(name):
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200")
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    maps = 0
    for size in ((34, 11, 12), (70, 15, 13)):            # even row lengths: TMA needs 16-byte aligned rows
        maps = _cartesian_case(name, code, info, size, 500 + len(name))
    if "tma::" in code:
        assert maps > 0, "the TMA variant was emitted but the plain one ran"


@pytest.mark.parametrize("name,opts", [(n, o) for n in ("hori_diff_type2_stencil", "tridiagonal_solve")
                                       for o in __import__("dawn_b200.build", fromlist=["VARIANTS"]).VARIANTS[n]],
                         ids=lambda v: v if isinstance(v, str) else ",".join("%s=%s" % kv for kv in v.items()))
def test_planner_variants_of_the_headline_kernels_on_the_cpu(name, opts):
    if opts.get("ring_stagger"):
        pytest.skip("reads the environment and sleeps: timing knob, not logic")
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200", **opts)
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    for size in ((70, 26, 12), (130, 13, 37)):
        _cartesian_case(name, code, info, size, 600)


# ---- persistent_ring: a block loops over column sets and requests the first ring slots of its next set while the
# last streamed range of the current one drains.  The emulated device has EMU_SMS resident slots: 1 and 3 make every
# block run several sets (the hand-over between sets, phases continuing across them), 1000 gives one set per block.
@pytest.mark.parametrize("slots", ["1", "3", "1000"])
@pytest.mark.parametrize("name,opts", [("tridiagonal_solve", dict()), ("tridiagonal_solve", dict(ring_levels=8, tma_stages=3)),
                                       ("tridiagonal_solve", dict(ring_early=0)), ("kcache_fill_backward", dict()),
                                       ("intervals03", dict(ring_levels=4, tma_stages=4))],
                         ids=lambda v: v if isinstance(v, str) else ",".join("%s=%s" % kv for kv in v.items()) or "default")
def test_persistent_column_blocks_on_the_cpu(name, opts, slots, monkeypatch):
    monkeypatch.setenv("EMU_SMS", slots)
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200", persistent_ring=1, **opts)
    assert "next column set" in code and "next set, chunk" in code
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    for size in ((70, 9, 12), (130, 5, 37)):
        _cartesian_case(name, code, info, size, 650)


@pytest.mark.parametrize("mutation", ["next set column", "next set sequence"])
def test_the_emulation_notices_a_broken_hand_over_between_sets(mutation, monkeypatch):
    monkeypatch.setenv("EMU_SMS", "2")
    name = "tridiagonal_solve"
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200", persistent_ring=1)
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    bad = {"next set column": code.replace("&tm_a, i0n, j0n, k0", "&tm_a, i0, j0n, k0"),
           "next set sequence": code.replace("const int k0 = pf_kb + (ch + 2 - nch) * 16;", "const int k0 = pf_kb + (ch + 1 - nch) * 16;")}[mutation]
    assert bad != code
    _cartesian_case(name, code, info, (130, 5, 37), 660)
    with pytest.raises(AssertionError):
        _cartesian_case(name, bad, info, (130, 5, 37), 660)


def test_persistent_blocks_are_off_by_default_and_leave_other_kernels_alone():
    plain = codegen.compile_iir(open(iir_path("tridiagonal_solve")).read(), backend="b200")
    assert "next column set" not in plain and "has_next" not in plain
    # tile kernels (horizontal dependencies) have no column ring: the option changes nothing there
    hd = open(iir_path("hori_diff_type2_stencil")).read()
    assert codegen.compile_iir(hd, backend="b200", persistent_ring=1) == codegen.compile_iir(hd, backend="b200")


@pytest.mark.parametrize("mutation", ["prefetch level", "box row", "hdmask slot offset"])
def test_the_emulation_notices_a_broken_tma_pipeline(mutation):
    # the harness must have teeth: a wrong level in the refill, a shifted box or a wrong slot offset in the
    # emitted TMA code has to show up as a mismatch against the oracle
    name = "hori_diff_type2_stencil"
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200")
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    bad = {"prefetch level": code.replace("&tm_in, i0 + (0), j0 + (0), kb + (it + 2),", "&tm_in, i0 + (0), j0 + (0), kb + (it + 1),"),
           "box row": code.replace("&tm_in, i0 + (0), j0 + (0), kb + q,", "&tm_in, i0 + (0), j0 + (1), kb + q,"),
           "hdmask slot offset": code.replace("tma_ring + ts * 2128 + 1232, &tm_hdmask", "tma_ring + ts * 2128 + 1230, &tm_hdmask")}[mutation]
    assert bad != code
    _cartesian_case(name, code, info, (70, 26, 12), 700)
    with pytest.raises(AssertionError):
        _cartesian_case(name, bad, info, (70, 26, 12), 700)


def test_small_launch_k_split_on_the_cpu():
    # the launcher's own rule for small launches (k split finer until 2 blocks per SM exist) stays covered too
    for name in ("hori_diff_type2_stencil", "lap"):
        code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200")
        info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
        I, J, K = 70, 26, 12
        ni, nj = I + 2 * HALO, J + 2 * HALO
        rng = np.random.default_rng(800)
        f = {fd["name"]: rng.uniform(0.5, 1.5, (K + 1 if "k" in fd["dims"] else 1, nj if "j" in fd["dims"] else 1,
                                                ni if "i" in fd["dims"] else 1)) for fd in info["fields"]}
        want = run_oracle(name, (ni, nj, K), f)
        got, _ = emu.run_cartesian(name, code, info, (ni, nj, K), (HALO, HALO, HALO, HALO, 0, 0), f, keep_k_chunks=False)
        for fd in info["fields"]:
            assert bit_equal(got[fd["name"]].reshape(want[fd["name"]].shape), want[fd["name"]]), fd["name"]


# ---- the optimized-IIR fixtures of the reference tree itself (dawn/test/**/*.iir), emitted where they lie and
# emulated against the goldens of tests/golden/reftree.npz (the GPU twin: tests/test_reference_tree_gpu.py) -----
import zlib  # noqa: E402

from test_reference_tree_gpu import CART_DOM, ENTRIES, GOLD, MESH  # noqa: E402

REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present")
@pytest.mark.parametrize("entry", ENTRIES, ids=[os.path.basename(e["file"]) for e in ENTRIES])
def test_reference_tree_fixture_on_the_cpu(entry):
    rel = entry["file"]
    gold = np.load(GOLD)
    code = codegen.compile_iir(open(os.path.join(REFERENCE, rel)).read(), backend=entry["backend"])
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    name = entry["stencil"]
    rng = np.random.default_rng(zlib.crc32(rel.encode()) & 0x7FFFFFFF)
    changed = [k2.split("::", 1)[1] for k2 in gold.files if k2.startswith(rel + "::")]
    f = {}
    if entry["backend"] == "b200":
        I, J, K = CART_DOM
        ni, nj = I + 2 * HALO, J + 2 * HALO
        for fd in info["fields"]:
            d = fd["dims"]
            f[fd["name"]] = rng.uniform(0.5, 1.5, (K + 1 if "k" in d else 1, nj if "j" in d else 1, ni if "i" in d else 1))
        got, _ = emu.run_cartesian(name, code, info, (ni, nj, K), (HALO, HALO, HALO, HALO, 0, 0), f)
        for fd in info["fields"]:
            n = fd["name"]
            want = gold[rel + "::" + n] if n in changed else f[n]
            assert bit_equal(got[n].reshape(want.shape), want), (rel, n)
    else:
        m, k = TriMesh(MESH[0], MESH[1]), MESH[2]
        for fd in info["fields"]:
            shape = [k] if fd.get("k", 1) else []
            if fd["location"] != "none":
                shape.append(m.num(LOC[fd["location"]]))
            if fd["sparse"]:
                shape.append(fd["sparse"])
            f[fd["name"]] = rng.uniform(0.5, 1.5, shape)
        got, _ = emu.run(name, code, info, m, k, f)
        for fd in info["fields"]:
            n = fd["name"]
            want = gold[rel + "::" + n] if n in changed else f[n]
            if fd["location"] == "none":
                assert bit_equal(got[n], want), (rel, n)
            else:
                v = m.valid(LOC[fd["location"]])
                axis = 1 if fd.get("k", 1) else 0
                assert bit_equal(np.compress(v, got[n], axis=axis), np.compress(v, want, axis=axis)), (rel, n)


@pytest.mark.parametrize("name", ["hori_diff_type2_stencil", "tridiagonal_solve", "lap", "hd_smagorinsky_stencil",
                                  "kcache_fill", "fuzz_03", "fuzz_07", "fuzz_12"])
def test_tiny_and_ragged_domains_through_the_default_kernels(name):
    # one point, a sliver, more than four tiles in i with three rows, a single column pair
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200")
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    for size in ((1, 1, 12), (3, 2, 12), (129, 3, 21), (2, 65, 12)):
        _cartesian_case(name, code, info, size, 900)


# ---- single precision (codegen option single_precision=1: ::dawn::float_type = float, the reference's
# DAWN_PRECISION switch, driver-includes/defs.hpp:47-81).  The oracle interprets the same IIR on float32
# storages; with -ffp-contract=off / -fmad=false every node is one IEEE single operation on both sides, so
# the comparison stays bit for bit (math calls excepted: libm and CUDA differ in the last place there) ------
F32_CARTESIAN = ["hori_diff_type2_stencil", "tridiagonal_solve", "hori_diff_stencil", "hori_diff_stencil_02", "lap",
                 "copy_stencil", "kcache_fill_backward", "intervals_stencil", "p_grad_c", "fuzz_03", "fuzz_09"]


def _f32_case(name, code, info, size, seed):
    I, J, K = size
    ni, nj = I + 2 * HALO, J + 2 * HALO
    rng = np.random.default_rng(seed)
    f = {}
    for fd in info["fields"]:
        d = fd["dims"]
        f[fd["name"]] = rng.uniform(0.5, 1.5, (K + 1 if "k" in d else 1, nj if "j" in d else 1,
                                               ni if "i" in d else 1)).astype(np.float32)
    if name in ("tridiagonal_solve", "kparallel_solver") and "b" in f:
        f["b"] += np.float32(4.0)
    g = CART_GLOBALS.get(name)
    if g is None and info.get("globals"):
        g = {n: 0.375 for n in info["globals"]}
    want = run_oracle(name, (ni, nj, K), f, globals_=g)
    got, _ = emu.run_cartesian(name, code, info, (ni, nj, K), (HALO, HALO, HALO, HALO, 0, 0), f, globals_=g)
    for fd in info["fields"]:
        n = fd["name"]
        assert got[n].dtype == np.float32 and want[n].dtype == np.float32
        assert bit_equal(got[n].reshape(want[n].shape), want[n]), (name, n, size)
    return emu.run_cartesian.tensor_maps


@pytest.mark.parametrize("plain", [False, True], ids=["default", "plain"])
@pytest.mark.parametrize("name", F32_CARTESIAN)
def test_single_precision_cartesian_kernels_on_the_cpu(name, plain):
    opts = dict(use_tma=0, column_ring=0) if plain else {}
    code = codegen.compile_iir(open(iir_path(name)).read(), backend="b200", single_precision=1, **opts)
    info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
    assert info["float_type"] == "float" and "sizeof(::dawn::float_type) == 4" in code
    maps = 0
    for size in ((36, 11, 12),) + (() if plain else ((72, 15, 13),)):    # row lengths that are multiples of 4: 16-byte rows of floats
        maps = _f32_case(name, code, info, size, 600 + len(name))
    if "tma::" in code and not plain:
        assert maps > 0, "the TMA variant was emitted but the plain one ran"


@pytest.mark.parametrize("name", ["ICON_laplacian_stencil", "ICON_nabla4_stencil", "diffusion", "gradient", "fuzzico_02",
                                  "fuzzico_07", "sparseAssignment3", "tridiagonalSolve", "verticalSum"])
def test_single_precision_unstructured_kernels_on_the_cpu(name):
    code, info = emitted(name, single_precision=1)
    assert info["float_type"] == "float"
    m, k = TriMesh(11, 8), 7
    f = {n: v.astype(np.float32) for n, v in random_fields(info, m, k, np.random.default_rng(19)).items()}
    if name == "tridiagonalSolve":
        f["b"] += np.float32(4.0)
    g = {n: 0.625 for n in info["globals"]} if info.get("globals") else None
    w = {n: v.astype(np.float32) for n, v in with_temporaries(name, m, k, f).items()}
    want = naive_interp.run_unstructured(iir_path(name), m, k, w, globals_=g)
    got, _ = emu.run(name, code, info, m, k, f, globals_=g)
    for fd in info["fields"]:
        assert got[fd["name"]].dtype == np.float32
    if not info.get("globals"):
        compare(info, m, got, want)
        return
    # a `double` global makes the reference evaluate the expression around it in double; the emitter's
    # value-numbered sub-results are ::dawn::float_type, so the last place may differ: the stated
    # single-precision tolerance (1e-5 relative) applies
    for fd in info["fields"]:
        n = fd["name"]
        v = m.valid(LOC[fd["location"]]) if fd["location"] != "none" else None
        a, b = (got[n], want[n]) if v is None else (np.compress(v, got[n], axis=1 if fd.get("k", 1) else 0),
                                                   np.compress(v, want[n], axis=1 if fd.get("k", 1) else 0))
        assert np.allclose(a, b, rtol=1e-5, atol=0.0, equal_nan=True), n
