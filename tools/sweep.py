"""GPU sweep of planner knobs for one workload: prints kernel ms per variant (kernel-only timing,
CUDA events, inputs resident).  Usage: python tools/sweep.py hori_diff 'rows_per_thread=1,2' 'max_blocks_per_sm=0,4,5' ...  (product of the axes)
       python tools/sweep.py tridiagonal list '' 'ring_levels=4,tma_stages=4' ...  (explicit variants, see
       tools/precompile.py to build them without a GPU first)"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import dawn_b200  # noqa: E402
import bench  # noqa: E402

HALO = 3
# further stencils of the reference's test list, with their algorithmic bytes per point
bench.STENCIL.update({"hd_smagorinsky": "hd_smagorinsky_stencil", "hori_diff_02": "hori_diff_stencil_02",
                      "lap": "lap"})
bench.SHAPE.update({"hd_smagorinsky": (1024, 1024, 80), "hori_diff_02": (1024, 1024, 80),
                    "lap": (1024, 1024, 80)})
bench.BYTES_PER_POINT.update({"hd_smagorinsky": 72, "hori_diff_02": 16, "lap": 16})


def main():
    workload = sys.argv[1]
    axes = []
    combos = None
    if len(sys.argv) > 2 and sys.argv[2] == "list":   # explicit variants: 'k=v,k=v' each ('' = default)
        combos = [[(kv.split("=")[0], int(kv.split("=")[1])) for kv in spec.split(",") if kv]
                  for spec in sys.argv[3:]]
    else:
        for a in sys.argv[2:]:
            k, vs = a.split("=")
            axes.append([(k, int(v)) for v in vs.split(",")])
    if os.environ.get("SWEEP_SHAPE"):
        bench.SHAPE[workload] = tuple(int(v) for v in os.environ["SWEEP_SHAPE"].lower().split("x"))
    I, J, K = bench.SHAPE[workload]
    name = bench.STENCIL[workload]
    ni, nj = I + 2 * HALO, J + 2 * HALO
    dom = dawn_b200.Domain(ni, nj, K)
    dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
    base = dawn_b200.load_stencil(name)
    stor = [dawn_b200.Storage(ni, nj, K + 1, f["dims"]) for f in base.fields]
    for s in stor:
        s.buf.uniform_(0.5, 1.5)
        if workload == "tridiagonal":
            s.buf.add_(4.0)
    pts = I * J * K
    for combo in (combos if combos is not None else itertools.product(*axes)):
        opts = {k: v for k, v in combo}
        try:
            mod = dawn_b200.load_stencil(name, **opts)
            st = mod(dom)
            sustained = os.environ.get("SWEEP_SUSTAINED", "0") == "1"
            for _ in range(400 if sustained else 5):   # sustained: let the power cap settle first
                st.run(*stor)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 300 if sustained else 30
            e0.record()
            for _ in range(n):
                st.run(*stor)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            gbs = bench.BYTES_PER_POINT[workload] * pts / (ms * 1e-3) / 1e9
            regs = ""
            pt = mod.path + ".ptxas.txt"
            if os.path.exists(pt):
                import re
                regs = ",".join(re.findall(r"Used (\d+) registers", open(pt).read()))
            print("%-70s %.4f ms  %.0f GB/s  %.1f%%  regs=%s" % (opts, ms, gbs, gbs / 65.415, regs),
                  flush=True)
        except Exception as e:  # noqa: BLE001
            print(opts, "FAILED", str(e)[:200], flush=True)


if __name__ == "__main__":
    main()
