"""One-off (not collected by pytest; run as `python tests/fuzz_more_cpu.py stage`): random unstructured programs through
the ico_stage / ico_hoist variants on the CPU emulation, against the oracle."""
import sys, json, re, os, tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tests/cpu_emu'); sys.path.insert(0,'/root/repo/tools')
import numpy as np
import make_fuzz_ico
import build as emu
from dawn_b200 import codegen
from dawn_b200.trimesh import TriMesh
from oracle import naive_interp
import test_kernel_emulation as T
fails=[]; staged=0; rejected=[]; n=0
for seed in range(100, 160):
    b = make_fuzz_ico.make(seed)
    iir = b.dumps()
    name = "fuzzico_%02d" % seed
    path = os.path.join(tempfile.gettempdir(), name + ".iir")
    open(path, "w").write(iir)
    for opts in (dict(), dict(ico_stage=1), dict(ico_hoist=1, ico_stage=2, levels_per_thread=3)):
        try:
            code = codegen.compile_iir(iir, backend="b200-ico", **opts)
        except codegen.CodeGenError as e:
            rejected.append((seed, str(e)[:100])); continue
        if "bulk_load_1d(" in code: staged += 1
        info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
        m, k = TriMesh(20, 14), 7
        f = T.random_fields(info, m, k, np.random.default_rng(seed))
        g = {n: 0.625 for n in info["globals"]} if info.get("globals") else None
        w = {n: v.copy() for n, v in f.items()}
        try:
            want = naive_interp.run_unstructured(path, m, k, w, globals_=g)
            got, _ = emu.run(name, code, info, m, k, f, globals_=g)
            T.compare(info, m, got, want); n += 1
        except AssertionError as e:
            fails.append((seed, opts, "MISMATCH", str(e)[:80]))
        except Exception as e:
            fails.append((seed, opts, "EXC", repr(e)[:120]))
print("compared:", n, "staged variants:", staged, "rejected:", len(rejected), rejected[:5], "fails:", len(fails))
for f in fails[:20]: print(f)
