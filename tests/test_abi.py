"""CPU tests: the C-ABI libraries load and export every symbol include/*.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from util import ROOT
from dawn_b200 import build, codegen


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rt_\w+|dawn_b200_\w+)\s*\(", text)))


def test_runtime_exports_every_declared_symbol():
    so = build.build_runtime()
    lib = ctypes.CDLL(so)
    names = _declared("dawn_b200_rt.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    from dawn_b200 import runtime
    assert sorted(runtime.SYMBOLS) == names


def test_codegen_exports_every_declared_symbol():
    so = build.build_emitter()
    lib = ctypes.CDLL(so)
    names = [n for n in _declared("dawn_b200_codegen.h")
             if n not in ("dawn_b200_options", "dawn_b200_backend")]
    assert "dawn_b200_codegen_run" in names
    for n in names:
        assert hasattr(lib, n), n


@pytest.mark.parametrize("name", ["hori_diff_type2_stencil", "tridiagonal_solve", "laplacian_stencil"])
def test_generated_module_exports_stencil_abi(name):
    so = build.module_path(name)
    if not os.path.exists(so):
        pytest.skip("generated modules not built (run __graft_entry__.build())")
    ctypes.CDLL(build.build_runtime(), mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(so)
    for sym in ("setup_", "run_", "run_rows_", "free_", "launches_", "last_launch_", "b200_module_info_"):
        assert hasattr(lib, sym + name), sym + name
    if name == "laplacian_stencil":
        assert hasattr(lib, "set_laplacian_stencil_dx") and hasattr(lib, "get_laplacian_stencil_dx")


def test_options_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "dawn_b200_codegen.h")).read()
    body = text[text.index("typedef struct dawn_b200_options {"):text.index("} dawn_b200_options;")]
    fields = re.findall(r"\bint\s+([a-z_0-9, ]+);", body)
    names = [x.strip() for f in fields for x in f.split(",")]
    assert names == [f[0] for f in codegen._Options._fields_]


def test_metrics_record_has_the_reference_json_shape():
    """driver-includes/to_json.cpp:9-53: {"stencil": [{"field": {5 metrics}} per iteration]}, a field
    already recorded for an iteration is kept, a new iteration is appended."""
    import json
    from dawn_b200 import runtime as rt
    rec = rt.MetricsRecord()
    assert rec.json() == "null"

    def m(it, valid, x):
        return rt.VerificationMetrics(it, int(valid), x, x / 2, x * 2, x / 4, 0)
    rec.add("nabla2", "div_vec", m(0, True, 1e-15))
    rec.add("nabla2", "rot_vec", m(0, True, 2e-15))
    rec.add("nabla2", "div_vec", m(0, False, 7.0))      # already there: kept
    rec.add("nabla2", "div_vec", m(1, False, 0.5))      # next iteration: appended
    rec.add("diffusion", "out_field", m(0, True, 0.0))
    d = json.loads(rec.json())
    assert list(d) == ["diffusion", "nabla2"]            # keys sorted like nlohmann::json objects
    assert len(d["nabla2"]) == 2 and set(d["nabla2"][0]) == {"div_vec", "rot_vec"}
    assert d["nabla2"][0]["div_vec"] == {
        "field_is_valid": True, "max_absolute_error": 2e-15, "max_relative_error": 1e-15,
        "min_absolute_error": 2.5e-16, "min_relative_error": 5e-16}
    assert d["nabla2"][1] == {"div_vec": {
        "field_is_valid": False, "max_absolute_error": 1.0, "max_relative_error": 0.5,
        "min_absolute_error": 0.125, "min_relative_error": 0.25}}
    assert d["diffusion"][0]["out_field"]["max_relative_error"] == 0.0
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        rec.write(f.name)
        assert json.load(open(f.name)) == d
    rec.close()


def test_near_device_is_a_no_op_without_numa_information():
    # no GPU here: the topology lookup must give up quietly and leave the affinity mask alone
    import os
    import torch
    from dawn_b200 import runtime as rt
    if torch.cuda.is_available():
        pytest.skip("a GPU box may well know its topology")
    before = os.sched_getaffinity(0)
    assert rt.device_numa_cpus(0) is None
    with rt.near_device(0) as near:
        assert near.cpus is None and os.sched_getaffinity(0) == before
    assert os.sched_getaffinity(0) == before
