"""In-tree builds: emitter (g++), C-ABI runtime (nvcc, sm_100a) and generated stencil modules.

Everything lands under dawn_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).  nvcc
cross-compiles for sm_100a without a GPU.  Generated translation units are compiled with
-fmad=false so that every arithmetic node stays one IEEE operation (bit parity with cxxnaive).
"""
import hashlib
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib")
GEN = os.path.join(LIB, "generated")
CSRC = os.path.join(PKG, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                      "-cudart", "shared"]

EMITTER_SRCS = ["codegen.cpp", "iir.cpp", "proto_wire.cpp", "emit_cartesian.cpp", "emit_ico.cpp", "capi.cpp"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_emitter(force=False):
    os.makedirs(LIB, exist_ok=True)
    edir = os.path.join(CSRC, "emitter")
    srcs = [os.path.join(edir, s) for s in EMITTER_SRCS]
    deps = srcs + [os.path.join(edir, h) for h in os.listdir(edir) if h.endswith((".hpp", ".inc"))] + \
        [os.path.join(ROOT, "include", "dawn_b200_codegen.h")]
    so = os.path.join(LIB, "libdawn_b200_codegen.so")
    exe = os.path.join(LIB, "dawn-b200-codegen")
    if force or _newer(so, deps):
        _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-o", so] + srcs)
    if force or _newer(exe, deps + [os.path.join(edir, "main.cpp")]):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(edir, "main.cpp")] +
             srcs[:-1])
    return so


def build_runtime(force=False):
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(CSRC, "runtime", "rt.cu")
    so = os.path.join(LIB, "libdawn_b200rt.so")
    if force or _newer(so, [src, os.path.join(ROOT, "include", "dawn_b200_rt.h")]):
        _run([NVCC] + NVCC_COMMON + ["-o", so, src, "-ldl"])
    return so


def module_path(name, tag=""):
    return os.path.join(LIB, "stencils", name + (("_" + tag) if tag else "") + ".so")


def compile_module(name, code, tag="", force=False, extra_flags=(), keep_ptxas=False):
    """Compiles one generated translation unit into dawn_b200/lib/stencils/<name>[_tag].so."""
    build_runtime()
    os.makedirs(GEN, exist_ok=True)
    os.makedirs(os.path.join(LIB, "stencils"), exist_ok=True)
    cu = os.path.join(GEN, name + (("_" + tag) if tag else "") + ".cu")
    so = module_path(name, tag)
    digest = hashlib.sha256(code.encode()).hexdigest()
    stamp = so + ".sha"
    hdrs = [os.path.join(PKG, "driver-includes", h)
            for h in os.listdir(os.path.join(PKG, "driver-includes"))]
    if (not force and os.path.exists(so) and os.path.exists(stamp)
            and open(stamp).read() == digest and not _newer(so, hdrs)):
        return so
    with open(cu, "w") as f:
        f.write(code)
    # hidden visibility + -Bsymbolic: several variants of one stencil (same kernel and class names)
    # can live in one process without their symbols interposing each other
    cmd = [NVCC] + NVCC_COMMON + ["-fmad=false", "-Xptxas", "-v", "-I", PKG, "-o", so, cu,
                                  "-Xcompiler", "-fvisibility=hidden", "-Xlinker", "-Bsymbolic",
                                  "-L", LIB, "-ldawn_b200rt", "-Xlinker", "-rpath,$ORIGIN/.."] + \
        list(extra_flags)
    out = _run(cmd)
    with open(so + ".ptxas.txt", "w") as f:
        f.write(out)
    with open(stamp, "w") as f:
        f.write(digest)
    return so


# planner variants exercised by tests/ and bench sweeps; prebuilt here so they travel to the GPU box
VARIANTS = {
    "hori_diff_type2_stencil": [
        dict(rows_per_thread=1), dict(rows_per_thread=4),
        dict(block_size_i=32, block_size_j=8, rows_per_thread=2), dict(inline_temporaries=0),
        dict(use_tma=0), dict(tma_stages=2, block_size_k=3), dict(inline_temporaries=0, use_tma=0),
        dict(cols_per_thread=1), dict(cols_per_thread=2, use_tma=0, rows_per_thread=1),
        dict(rows_per_thread=3), dict(ring_drain=2), dict(shared_temporaries=1),
        dict(shared_temporaries=1, cols_per_thread=1), dict(shared_temporaries=1, rows_per_thread=1),
        dict(shared_temporaries=1, tma_stages=2, block_size_j=8),
    ],
    "hori_diff_stencil_01": [dict(shared_temporaries=1)], "hori_diff_stencil": [dict(shared_temporaries=1)],
    "copy_stencil": [dict(shared_temporaries=1)], "nonoverlapping_stencil": [dict(shared_temporaries=1)],
    "ICON_laplacian_stencil": [dict(fuse_columns=0, levels_per_thread=1), dict(levels_per_thread=8),
                               dict(co_schedule=0), dict(ico_chain=1), dict(ico_chain=1, co_schedule=0),
                               dict(ico_chain=1, ico_chain_lag=1), dict(ico_chain=1, ico_chain_lag=100000),
                               dict(ico_chain=1, ico_chain_l1=1), dict(ico_chain=1, ico_cache_hints=1),
                               dict(ico_stage=1), dict(ico_stage=2), dict(ico_stage=2, levels_per_thread=5),
                               dict(ico_hoist=1), dict(ico_hoist=1, ico_stage=2)],
    "ICON_nabla4_stencil": [dict(fuse_columns=0, levels_per_thread=1), dict(co_schedule=0), dict(ico_chain=1),
                            dict(ico_stage=1)],
    "ICON_laplacian_diamond_stencil": [dict(ico_stage=1), dict(ico_hoist=1, ico_stage=2)],
    "overwriteGathered": [dict(ico_chain=1)],
    **{"fuzzico_%02d" % n: [dict(ico_chain=1)] for n in range(16, 20)},
    "kcache_fill": [dict(column_ring=0)], "kcache_flush": [dict(column_ring=0)],
    "kcache_fill_backward": [dict(column_ring=0)], "intervals02": [dict(column_ring=0)],
    "intervals03": [dict(column_ring=0)], "kparallel_solver": [dict(column_ring=0)],
    "asymmetric_stencil": [dict(use_tma=0)], "coriolis_stencil": [dict(use_tma=0)],
    **{n: [dict(use_tma=0, inline_temporaries=0)] for n in (
        "intervals_stencil", "intervals01", "kcache_fill_kparallel", "kcache_epflush", "lap",
        "hori_diff_stencil_02", "hd_smagorinsky_stencil", "p_grad_c")},
    **{"fuzz_%02d" % n: [dict(use_tma=0, inline_temporaries=0)] for n in range(16)},
    "tridiagonal_solve": [dict(ring_early=0), dict(ring_stagger=-1), dict(ring_drain=0), dict(ring_drain=1, tma_stages=2), dict(ring_drain=1, tma_stages=4),
                          dict(ring_drain=1, ring_levels=4, tma_stages=4), dict(fuse_columns=0), dict(column_ring=0), dict(column_ring=0, column_cache=1, prefetch_k=3),
                          dict(ring_levels=2, tma_stages=2), dict(block_size_i=32, block_size_j=2),
                          dict(ring_producer=1), dict(ring_producer=1, ring_levels=4, tma_stages=3, block_size_i=32),
                          dict(persistent_ring=1), dict(persistent_ring=1, ring_levels=4, tma_stages=4)],
}


# single-precision builds (codegen option single_precision=1) the GPU tests load
PRECISION_VARIANTS = {n: [dict(single_precision=1)] for n in (
    "hori_diff_type2_stencil", "tridiagonal_solve", "hori_diff_stencil_02", "ICON_laplacian_stencil", "diffusion",
    "ICON_laplacian_diamond_stencil")}

# optimized IIR fixtures shipped by the reference itself: compiled where they lie when the reference
# tree is present (build container); only the built modules travel to the GPU box
REFERENCE_IIR = {
    "update_dz_c": "dawn/test/unit-test/dawn/CodeGen/input/update_dz_c.iir",
    "conditional_stencil": "dawn/test/unit-test/dawn/CodeGen/input/conditional_stencil.iir",
}
REFERENCE_ROOT = "/root/reference"


# three letters per option: the first two of the name and the first of its last word.  Options whose letters would
# repeat those of an older option get their own (ico_hoist / ico_cache_hints, persistent_ring is free)
_TAG = {"ico_hoist": "iho"}


def variant_tag(options):
    return "_".join("%s%s" % (_TAG.get(k, k[:2] + k.split("_")[-1][:1]), v) for k, v in sorted(options.items()))


def build_all(force=False, jobs=None):
    """Everything bench/tests need: emitter, runtime, and a module per stencils/*.iir fixture (plus the
    planner variants above).  Code generation is serial (milliseconds each); the nvcc compilations run
    `jobs` at a time (default: one per core)."""
    from concurrent.futures import ThreadPoolExecutor

    from . import codegen
    build_emitter(force)
    build_runtime(force)
    work = []  # (name, code, tag)
    sdir = os.path.join(ROOT, "stencils")
    for fn in sorted(os.listdir(sdir)):
        if not fn.endswith(".iir"):
            continue
        name = fn[:-4]
        with open(os.path.join(sdir, fn)) as f:
            iir = f.read()
        backend = "b200-ico" if '"gridType": "Unstructured"' in iir else "b200"
        try:
            code = codegen.compile_iir(iir, backend=backend)
        except codegen.CodeGenError as e:
            if "not yet supported" in str(e):
                continue
            raise
        work.append((name, code, ""))
        for opts in VARIANTS.get(name, []) + PRECISION_VARIANTS.get(name, []):
            work.append((name, codegen.compile_iir(iir, backend=backend, **opts), variant_tag(opts)))
    for name, rel in REFERENCE_IIR.items():
        path = os.path.join(REFERENCE_ROOT, rel)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            iir = f.read()
        work.append((name, codegen.compile_iir(iir, backend="b200"), "ref"))
        work.append((name, codegen.compile_iir(iir, backend="b200", use_tma=0, inline_temporaries=0),
                     "ref_plain"))
    jobs = jobs or os.cpu_count() or 1
    with ThreadPoolExecutor(max_workers=jobs) as pool:
        built = list(pool.map(lambda w: compile_module(w[0], w[1], tag=w[2], force=force), work))
    return built


def build_reference_tree(force=False, jobs=None):
    """Compiles a module for every optimized-IIR fixture of the reference tree the emitters accept
    (only where /root/reference exists; the modules travel to the GPU box).  Returns and writes the
    manifest dawn_b200/lib/stencils/reftree_manifest.json: file (relative to the reference root),
    backend, stencil name, module path, or the diagnostic the emitter gave."""
    import glob
    import json
    import re
    from concurrent.futures import ThreadPoolExecutor

    from . import codegen
    if not os.path.isdir(REFERENCE_ROOT):
        return []
    build_emitter(force)
    build_runtime(force)
    entries, work = [], []
    for path in sorted(glob.glob(os.path.join(REFERENCE_ROOT, "**", "*.iir"), recursive=True)):
        rel = os.path.relpath(path, REFERENCE_ROOT)
        with open(path) as f:
            iir = f.read()
        backend = "b200-ico" if re.search(r'"gridType"\s*:\s*"Unstructured"', iir) else "b200"
        mod = "reftree_" + re.sub(r"[^A-Za-z0-9]", "_", rel)[-60:]
        try:
            code = codegen.compile_iir(iir, backend=backend)
        except codegen.CodeGenError as e:
            entries.append({"file": rel, "backend": backend, "rejected": str(e)})
            continue
        m = re.search(r'setup_(\w+)\(void\*\* handle', code)
        entries.append({"file": rel, "backend": backend, "stencil": m.group(1), "module": mod})
        work.append((mod, code))
    def _compile(w):
        try:
            compile_module(w[0], w[1], force=force)
            return None
        except RuntimeError as e:
            return str(e)[-400:]
    with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 1) as pool:
        errors = list(pool.map(_compile, work))
    for (mod, _), err in zip(work, errors):
        if err:
            for e in entries:
                if e.get("module") == mod:
                    e["rejected"] = "nvcc: " + err
                    del e["module"]
    with open(os.path.join(LIB, "stencils", "reftree_manifest.json"), "w") as f:
        json.dump(entries, f, indent=1)
    return [e for e in entries if "module" in e]


def clean():
    shutil.rmtree(LIB, ignore_errors=True)
