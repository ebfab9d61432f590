"""The generated class API used from C++ exactly like the reference's drivers use theirs
(gtclang/test/integration-test/CodeGen/hori_diff_type2_stencil_benchmark.cpp)."""
import os
import subprocess
import sys

import pytest

from dawn_b200 import build

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp"))
from build_driver import build_cpp_driver  # noqa: E402


def test_cpp_driver_compiles_against_generated_code():
    exe = build_cpp_driver()
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(12, 12, 10), (70, 33, 7), (512, 512, 40)])
def test_cpp_driver_verifies_on_gpu(size):
    exe = os.path.join(build.LIB, "hori_diff_driver")
    if not os.path.exists(exe):
        exe = build_cpp_driver()
    r = subprocess.run([exe] + [str(x) for x in size], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "VERIFIED" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("args", [(70, 33, 7, 2, 3), (70, 33, 7, 2, 0), (130, 64, 20, 3, 8), (64, 48, 5, 2, 100)])
def test_cpp_class_end_to_end_with_host_pipeline(args):
    """tests/cpp/hori_diff_e2e.cu: host views written every step, run() of the generated class with its host <->
    device pipeline (chunks of `levels` levels on three streams; 0 = whole storages, the reference's shape),
    result read from the host mirror and compared with the oracle bit for bit."""
    import json
    from build_driver import build_e2e_driver
    exe = os.path.join(build.LIB, "hori_diff_e2e")
    if not os.path.exists(exe):
        exe = build_e2e_driver()
    r = subprocess.run([exe] + [str(x) for x in args] + ["2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["differing_values"] == 0 and d["steps"] == args[3] and d["chunk_levels"] == args[4]
