"""One-off wider differential run on the CPU emulation (not collected by pytest): random unstructured programs
that overwrite gathered inputs (plain, chained, chained + L1 hand-over) and random Cartesian programs (default,
plain kernels, single precision) with seeds beyond the committed fixtures, against the oracle.
    python tests/fuzz_more_cpu.py ico|cart|stage     (stage: the same unstructured seeds through the staged-row and
hoisted-load variants on a mesh with odd element counts; round 2: 180 comparisons, 120 staged kernels, no mismatch)"""
import sys
KIND = sys.argv[1] if len(sys.argv) > 1 else "ico"
if KIND == "stage":
    exec(open(__file__.replace("fuzz_more_cpu.py", "fuzz_more_stage.py")).read())
elif KIND == "ico":
    import sys, json, re, os, tempfile
    sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tests/cpu_emu'); sys.path.insert(0,'/root/repo/tools')
    import numpy as np
    import make_fuzz_ico
    import build as emu
    from dawn_b200 import codegen
    from dawn_b200.trimesh import TriMesh
    from oracle import naive_interp
    import test_kernel_emulation as T
    fails=[]; chained=0; rejected=0
    for seed in range(100, 160):
        b = make_fuzz_ico.make(seed)
        iir = b.dumps()
        name = "fuzzico_%02d" % seed
        path = os.path.join(tempfile.gettempdir(), name + ".iir")
        open(path, "w").write(iir)
        for opts in (dict(), dict(ico_chain=1), dict(ico_chain=1, ico_chain_l1=1, co_schedule=0)):
            try:
                code = codegen.compile_iir(iir, backend="b200-ico", **opts)
            except codegen.CodeGenError as e:
                rejected += 1; continue
            if "chained with" in code: chained += 1
            info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
            m, k = TriMesh(11, 8), 6
            f = T.random_fields(info, m, k, np.random.default_rng(seed))
            g = {n: 0.625 for n in info["globals"]} if info.get("globals") else None
            d = json.loads(iir)["metadata"]
            w = {n: v.copy() for n, v in f.items()}
            try:
                want = naive_interp.run_unstructured(path, m, k, w, globals_=g)
                got, _ = emu.run(name, code, info, m, k, f, globals_=g)
                T.compare(info, m, got, want)
            except AssertionError as e:
                fails.append((seed, opts, "MISMATCH", str(e)[:80]))
            except Exception as e:
                fails.append((seed, opts, "EXC", repr(e)[:120]))
    print("chained variants:", chained, "rejected:", rejected, "fails:", len(fails))
    for f in fails[:20]: print(f)
else:
    import sys, json, re, os, tempfile
    sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tests/cpu_emu'); sys.path.insert(0,'/root/repo/tools')
    import numpy as np
    import make_fuzz
    import build as emu
    from dawn_b200 import codegen
    from oracle import naive_interp
    import test_kernel_emulation as T
    HALO=3
    fails=[]; rejected=0; n=0
    for seed in range(100, 140):
        b = make_fuzz.make(seed)
        iir = b.dumps(); name = "fuzz_%02d" % seed
        path = os.path.join(tempfile.gettempdir(), name + ".iir"); open(path, "w").write(iir)
        for opts in (dict(), dict(use_tma=0, column_ring=0), dict(single_precision=1)):
            try:
                code = codegen.compile_iir(iir, backend="b200", **opts)
            except codegen.CodeGenError as e:
                rejected += 1; continue
            info = json.loads(re.search(r'return "(\{.*\})";', code).group(1).encode().decode("unicode_escape"))
            I, J, K = (36, 11, 12)
            ni, nj = I + 2*HALO, J + 2*HALO
            rng = np.random.default_rng(seed)
            ft = np.float32 if opts.get("single_precision") else np.float64
            f = {}
            for fd in info["fields"]:
                d = fd["dims"]
                f[fd["name"]] = rng.uniform(0.5, 1.5, (K + 1 if "k" in d else 1, nj if "j" in d else 1, ni if "i" in d else 1)).astype(ft)
            g = {nm: 0.375 for nm in info["globals"]} if info.get("globals") else None
            try:
                out = {k_: v.copy() for k_, v in f.items()}
                naive_interp.run_cartesian(path, (ni, nj, K), out, halos=(HALO,)*4+(0,0), globals_=g)
                got, _ = emu.run_cartesian(name, code, info, (ni, nj, K), (HALO,)*4+(0,0), f, globals_=g)
                n += 1
                for fd in info["fields"]:
                    nm = fd["name"]
                    a, w = got[nm].reshape(out[nm].shape), out[nm]
                    if opts.get("single_precision") and info.get("globals"):
                        ok = np.allclose(a, w, rtol=1e-5, atol=0, equal_nan=True)
                    else:
                        ok = np.array_equal(a, w, equal_nan=True)
                    if not ok:
                        fails.append((seed, opts, nm)); break
            except Exception as e:
                fails.append((seed, opts, "EXC", repr(e)[:150]))
    print("ran", n, "rejected", rejected, "fails", len(fails))
    for f_ in fails[:20]: print(f_)
