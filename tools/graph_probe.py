"""Launch-bound regime: a chain of small stencil runs (tutorial laplacian 128x128x80, config C1) issued
launch by launch from the host vs replayed as one CUDA graph (dawn_b200.TimeStep)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dawn_b200  # noqa: E402

HALO = 3


def main():
    I = J = 128
    K = 80
    chain = 20
    ni, nj = I + 2 * HALO, J + 2 * HALO
    dom = dawn_b200.Domain(ni, nj, K)
    dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
    cp = dawn_b200.load_stencil("copy_stencil")(dom)
    hd = dawn_b200.load_stencil("hori_diff_stencil_01")(dom)
    a = dawn_b200.Storage(ni, nj, K + 1, "ijk")
    b = dawn_b200.Storage(ni, nj, K + 1, "ijk")
    a.buf.uniform_(0.5, 1.5)
    step = dawn_b200.TimeStep()
    for n in range(chain // 2):
        step.run(hd, a, b)      # b = hd(a)
        step.run(cp, b, a)      # a = b
    step.capture()

    def timed(fn, reps=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_eager = timed(step.eager)
    t_graph = timed(step.replay)
    print("chain of %d stencil runs on %dx%dx%d: eager %.3f ms/step (%.1f us/launch), graph %.3f ms/step "
          "(%.1f us/launch), x%.2f" % (chain, I, J, K, t_eager, 1e3 * t_eager / chain, t_graph,
                                       1e3 * t_graph / chain, t_eager / t_graph))


if __name__ == "__main__":
    main()
