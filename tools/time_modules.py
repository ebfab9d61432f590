"""Times explicitly named module files of one Cartesian workload (experiments with hand-edited generated code):
python tools/time_modules.py tridiagonal 512x512x80 dawn_b200/lib/stencils/a.so ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import dawn_b200  # noqa: E402
import bench  # noqa: E402
from dawn_b200.stencil import StencilModule  # noqa: E402

workload, shape = sys.argv[1], tuple(int(v) for v in sys.argv[2].split("x"))
I, J, K = shape
name = bench.STENCIL[workload]
ni, nj = I + 6, J + 6
dom = dawn_b200.Domain(ni, nj, K)
dom.set_halos(3, 3, 3, 3, 0, 0)
for path in sys.argv[3:]:
    mod = StencilModule(name, os.path.join(ROOT, path))
    stor = [dawn_b200.Storage(ni, nj, K + 1, f["dims"]) for f in mod.fields]
    for s in stor:
        s.buf.uniform_(0.5, 1.5)
        if workload == "tridiagonal":
            s.buf.add_(4.0)
    st = mod(dom)
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
    gbs = bench.BYTES_PER_POINT[workload] * I * J * K / (ms * 1e-3) / 1e9
    print("%-60s %s %.4f ms %.0f GB/s %.1f%%" % (os.path.basename(path), sys.argv[2], ms, gbs, gbs / 65.415), flush=True)
