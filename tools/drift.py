"""Per-batch kernel time over consecutive batches (looks for power-cap drift)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, dawn_b200
HALO = 3; I = J = 1024; K = 80; ni, nj = I + 6, J + 6
mod = dawn_b200.load_stencil("hori_diff_type2_stencil")
dom = dawn_b200.Domain(ni, nj, K); st = mod(dom)
stor = [dawn_b200.Storage(ni, nj, K + 1, f["dims"]) for f in mod.fields]
for s in stor: s.buf.uniform_(0.5, 1.5)
for _ in range(5): st.run(*stor)
torch.cuda.synchronize()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
evs = [torch.cuda.Event(enable_timing=True) for _ in range(nb + 1)]
evs[0].record()
for b in range(nb):
    for _ in range(100): st.run(*stor)
    evs[b + 1].record()
torch.cuda.synchronize()
print(" ".join("%.4f" % (evs[b].elapsed_time(evs[b + 1]) / 100) for b in range(nb)))
