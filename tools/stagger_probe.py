"""GPU probe of ring_stagger (start delay per phase for the first-wave blocks of a column-ring kernel):
kernel ms of tridiagonal_solve per delay, both bench shapes.  Module built with ring_stagger=-1 reads
DAWN_B200_STAGGER_NS / DAWN_B200_STAGGER_PHASES at every launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dawn_b200  # noqa: E402

HALO = 3


def main():
    delays = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,2000,4000,6000,8000,12000").split(",")]
    phases = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
    mod = dawn_b200.load_stencil("tridiagonal_solve", ring_stagger=-1)
    for (I, J, K) in ((512, 512, 80), (1024, 1024, 80)):
        ni, nj = I + 2 * HALO, J + 2 * HALO
        dom = dawn_b200.Domain(ni, nj, K)
        dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
        st = mod(dom)
        stor = [dawn_b200.Storage(ni, nj, K + 1, f["dims"]) for f in mod.fields]
        for s in stor:
            s.buf.uniform_(4.5, 5.5)
        for ph in phases:
            for d in delays:
                os.environ["DAWN_B200_STAGGER_NS"] = str(d)
                os.environ["DAWN_B200_STAGGER_PHASES"] = str(ph)
                for _ in range(10):
                    st.run(*stor)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 100
                e0.record()
                for _ in range(n):
                    st.run(*stor)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                gbs = 48 * I * J * K / (ms * 1e-3) / 1e9
                print("%dx%dx%d phases=%d delay=%5d ns  %.4f ms  %.0f GB/s  %.1f%%" % (
                    I, J, K, ph, d, ms, gbs, gbs / 65.415), flush=True)


if __name__ == "__main__":
    main()
