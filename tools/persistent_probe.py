"""One-shot GPU check of the persistent column-ring kernel (codegen option persistent_ring=1) against the default
module of the same stencil: results bit for bit on the same inputs (the default module is the one the parity tests
pin on the oracle at these shapes), then kernel time of both, alternating.   python tools/persistent_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import dawn_b200  # noqa: E402
import bench  # noqa: E402

name = "tridiagonal_solve"
mods = [("default", dawn_b200.load_stencil(name)), ("persistent_ring=1", dawn_b200.load_stencil(name, persistent_ring=1))]
for shape in ((200, 77, 37), (512, 512, 80), (1024, 1024, 80)):
    I, J, K = shape
    ni, nj = I + 6, J + 6
    dom = dawn_b200.Domain(ni, nj, K)
    dom.set_halos(3, 3, 3, 3, 0, 0)
    base = [dawn_b200.Storage(ni, nj, K + 1, f["dims"]) for f in mods[0][1].fields]
    for s in base:
        s.buf.uniform_(0.5, 1.5)
    base[[f["name"] for f in mods[0][1].fields].index("b")].buf.add_(4.0)
    outs, sts = [], []
    for tag, mod in mods:
        stor = [dawn_b200.Storage(ni, nj, K + 1, f["dims"]) for f in mod.fields]
        for s, b in zip(stor, base):
            s.buf.copy_(b.buf)
        st = mod(dom)
        st.run(*stor)
        torch.cuda.synchronize()
        outs.append([s.buf.clone() for s in stor])
        sts.append((tag, st, stor))
    same = all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
    print("%dx%dx%d persistent == default bit for bit: %s   (last launch %s)" % (I, J, K, same, sts[1][1].last_launch()
          if hasattr(sts[1][1], "last_launch") else "?"), flush=True)
    if I < 512:
        continue
    for rep in range(2):
        for tag, st, stor in sts:
            for _ in range(5):
                st.run(*stor)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                st.run(*stor)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 30
            gbs = bench.BYTES_PER_POINT["tridiagonal"] * I * J * K / (ms * 1e-3) / 1e9
            print("  %-20s %dx%dx%d %.4f ms %.0f GB/s %.1f%%" % (tag, I, J, K, ms, gbs, gbs / 65.415), flush=True)
