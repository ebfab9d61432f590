"""PCIe-bound end-to-end path: raw pinned H2D / D2H rates on this box and Stencil.run_from_host over
k_chunk choices (hori_diff_type2 1024x1024x80)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import dawn_b200  # noqa: E402

HALO = 3
I = J = 1024
K = 80
ni, nj = I + 2 * HALO, J + 2 * HALO
n = ni * nj * (K + 1)
host = torch.empty(n, dtype=torch.float64).pin_memory()
host.uniform_(0.5, 1.5)
dev = torch.empty(n, dtype=torch.float64, device="cuda")
for name, fn in (("H2D", lambda: dev.copy_(host, non_blocking=True)),
                 ("D2H", lambda: host.copy_(dev, non_blocking=True))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("%s %.2f GB in %.2f ms = %.1f GB/s" % (name, n * 8 / 1e9, dt * 1e3, n * 8 / dt / 1e9))
s2 = torch.cuda.Stream()
host2 = torch.empty(n, dtype=torch.float64).pin_memory()
dev2 = torch.empty(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        host2.copy_(dev2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("H2D + D2H concurrently: %.2f ms for %.2f GB each way = %.1f GB/s per direction" % (
    dt * 1e3, n * 8 / 1e9, n * 8 / dt / 1e9))

dom = dawn_b200.Domain(ni, nj, K)
dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
mod = dawn_b200.load_stencil("hori_diff_type2_stencil")
st = mod(dom)
arrays = []
for f in mod.fields:
    d = f["dims"]
    a = torch.empty((K + 1 if "k" in d else 1, nj if "j" in d else 1, ni if "i" in d else 1),
                    dtype=torch.float64).pin_memory()
    a.uniform_(0.5, 1.5)
    arrays.append(a)
for kc in (0, 20, 10, 8, 5, 4, 2):
    st._host_storages = None
    st.run_from_host(*arrays, upload_outputs=False, k_chunk=kc)
    t0 = time.perf_counter()
    for _ in range(4):
        h2d, d2h = st.run_from_host(*arrays, upload_outputs=False, k_chunk=kc)
    dt = (time.perf_counter() - t0) / 4
    print("k_chunk %2d: %.2f ms/step  %.3g updates/s  (h2d %.2f GB, d2h %.2f GB)" % (
        kc, dt * 1e3, I * J * K / dt, h2d / 1e9, d2h / 1e9))
