"""Builds tests/cpp/hori_diff_driver.cu: a reference-style C++ driver that #includes the generated
translation unit and checks its result against the oracle's C restatement (test infrastructure: the
only place where generated code and oracle/ are linked into one binary)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from dawn_b200 import build as _b, codegen  # noqa: E402


def build_cpp_driver(force=False):
    src = os.path.join(HERE, "hori_diff_driver.cu")
    exe = os.path.join(_b.LIB, "hori_diff_driver")
    gen = os.path.join(_b.GEN, "hori_diff_type2_stencil.cu")
    if not os.path.exists(gen):
        with open(os.path.join(ROOT, "stencils", "hori_diff_type2_stencil.iir")) as f:
            code = codegen.compile_iir(f.read())
        _b.compile_module("hori_diff_type2_stencil", code)
    oracle_so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(oracle_so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    if force or _b._newer(exe, [src, gen, oracle_so]):
        _b._run([_b.NVCC] + _b.ARCH + ["-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-cudart", "shared",
                                      "-I", _b.PKG, '-DGENERATED_FILE="%s"' % gen, "-o", exe, src,
                                      "-L", _b.LIB, "-ldawn_b200rt", "-Xlinker", "-rpath,$ORIGIN",
                                      oracle_so, "-Xlinker", "-rpath," + os.path.join(ROOT, "oracle")])
    return exe


def build_e2e_driver(force=False):
    """tests/cpp/hori_diff_e2e.cu: the end-to-end benchmark through the generated C++ class (bench.py's e2e leg)."""
    src = os.path.join(HERE, "hori_diff_e2e.cu")
    exe = os.path.join(_b.LIB, "hori_diff_e2e")
    gen = os.path.join(_b.GEN, "hori_diff_type2_stencil.cu")
    build_cpp_driver()      # makes sure the generated file and liboracle.so exist
    oracle_so = os.path.join(ROOT, "oracle", "liboracle.so")
    hdrs = [os.path.join(_b.PKG, "driver-includes", h) for h in os.listdir(os.path.join(_b.PKG, "driver-includes"))]
    if force or _b._newer(exe, [src, gen, oracle_so] + hdrs):
        _b._run([_b.NVCC] + _b.ARCH + ["-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-cudart", "shared",
                                      "-I", _b.PKG, '-DGENERATED_FILE="%s"' % gen, "-o", exe, src,
                                      "-L", _b.LIB, "-ldawn_b200rt", "-Xlinker", "-rpath,$ORIGIN",
                                      oracle_so, "-Xlinker", "-rpath," + os.path.join(ROOT, "oracle")])
    return exe


if __name__ == "__main__":
    print(build_cpp_driver())
    print(build_e2e_driver())
