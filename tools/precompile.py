"""Cross-compiles planner variants of one stencil here (no GPU needed) so that a GPU sweep
(tools/sweep.py) only has to load them.  Usage: python tools/precompile.py tridiagonal_solve 'ring_levels=4,tma_stages=4' ..."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dawn_b200 import build, codegen  # noqa: E402


def one(name, spec):
    opts = {k: int(v) for k, v in (kv.split("=") for kv in spec.split(",") if kv)}
    iir = open(os.path.join(ROOT, "stencils", name + ".iir")).read()
    backend = "b200-ico" if '"gridType": "Unstructured"' in iir else "b200"
    try:
        code = codegen.compile_iir(iir, backend=backend, **opts)
        so = build.compile_module(name, code, tag=build.variant_tag(opts))
        import re
        regs = ",".join(re.findall(r"Used (\d+) registers", open(so + ".ptxas.txt").read()))
        spill = ",".join(re.findall(r"(\d+) bytes spill stores", open(so + ".ptxas.txt").read()))
        return "%-60s regs=%s spill=%s" % (spec, regs, spill)
    except Exception as e:  # noqa: BLE001
        return "%-60s FAILED %s" % (spec, str(e)[:300])


if __name__ == "__main__":
    name = sys.argv[1]
    build.build_runtime()
    with ThreadPoolExecutor(8) as ex:
        for line in ex.map(lambda s: one(name, s), sys.argv[2:]):
            print(line, flush=True)
