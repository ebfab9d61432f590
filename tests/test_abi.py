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
    for sym in ("setup_", "run_", "run_rows_", "free_", "launches_", "b200_module_info_"):
        assert hasattr(lib, sym + name), sym + name
    if name == "laplacian_stencil":
        assert hasattr(lib, "set_laplacian_stencil_dx") and hasattr(lib, "get_laplacian_stencil_dx")


def test_options_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "dawn_b200_codegen.h")).read()
    body = text[text.index("typedef struct dawn_b200_options {"):text.index("} dawn_b200_options;")]
    fields = re.findall(r"\bint\s+([a-z_0-9, ]+);", body)
    names = [x.strip() for f in fields for x in f.split(",")]
    assert names == [f[0] for f in codegen._Options._fields_]
