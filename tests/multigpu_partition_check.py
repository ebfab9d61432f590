"""Multi-GPU parity of a PARTITIONED unstructured mesh (dawn_b200/partition.py), launched under torchrun:
ICON nabla2 on an RCM-renumbered TriMesh cut into one part per GPU, `vec` halo copies refreshed from
their owners over NCCL point-to-point, generated kernels on the local meshes; owned results must equal
the oracle's on the undivided mesh bit for bit; then ICON nabla4 the same way with two halo rings (no exchange
between its passes).  Round 2: nabla2 green on 2 and 8 B200s, nabla4 on 2 (the host logic is covered by
tests/test_partition.py with gloo ranks); launched under torchrun, so not part of the single-GPU pytest run.
    torchrun --nproc-per-node N tests/multigpu_partition_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import bit_equal, iir_path  # noqa: E402
from oracle import naive_interp  # noqa: E402

import dawn_b200  # noqa: E402
from dawn_b200 import partition as pt, reorder as ro  # noqa: E402
from dawn_b200.trimesh import EDGE, TriMesh  # noqa: E402
from test_partition import LOC_OF, nabla2_fields  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    m, k = ro.reorder(TriMesh(40, 12 * world), "rcm"), 9
    f = nabla2_fields(m, k, 61)
    want = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k,
                                         {n: v.copy() for n, v in f.items()})
    p = pt.partition(m, world)[rank]
    mod = dawn_b200.load_stencil("ICON_laplacian_stencil")
    st = mod.on_mesh(p.mesh, k, splitters=p.splitters())
    tens = {}
    for fd in mod.fields:
        a = p.to_local(LOC_OF[fd["name"]], f[fd["name"]], axis=1)
        if fd["sparse"]:
            a = np.swapaxes(a, -1, -2)                       # oracle [k, n, nbh] -> device [k, nbh, n]
        tens[fd["name"]] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tens["vec"][:, p.n_owned[EDGE]:] = float("nan")          # stale halo copies
    p.exchange(EDGE, tens["vec"], dist)
    st.run(*[tens[fd["name"]] for fd in mod.fields])
    torch.cuda.synchronize()
    ok = True
    for n in ("rot_vec", "div_vec", "nabla2_vec"):
        loc = LOC_OF[n]
        no = p.n_owned[loc]
        v = p.mesh.valid(loc)[:no]
        got = tens[n].cpu().numpy()
        ok = ok and bit_equal(got[:, :no][:, v], want[n][:, p.global_ids[loc][:no]][:, v])
    # nabla4 = nabla2 of nabla2 on the same partition scheme with TWO halo rings: the whole stencil runs on the
    # local mesh, only `vec` is refreshed, no exchange between its passes (tests/test_partition.py on the CPU)
    from test_unstructured_oracle import _nabla_inputs, nabla4_oracle
    f4 = _nabla_inputs(m, k, 43)
    f4["nabla2_vec"] = np.full((k, m.num_edges), -1.0)
    f4["nabla4_vec"] = np.full((k, m.num_edges), -1.0)
    want4 = nabla4_oracle(m, k, f4)
    p2 = pt.partition(m, world, rings=2)[rank]
    mod4 = dawn_b200.load_stencil("ICON_nabla4_stencil")
    st4 = mod4.on_mesh(p2.mesh, k, splitters=p2.splitters())
    loc4 = {"geofac_rot": 2, "geofac_div": 0}
    t4 = {}
    for fd in mod4.fields:
        a = p2.to_local(loc4.get(fd["name"], EDGE), f4[fd["name"]], axis=1)
        if fd["sparse"]:
            a = np.swapaxes(a, -1, -2)
        t4[fd["name"]] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t4["vec"][:, p2.n_owned[EDGE]:] = float("nan")
    p2.exchange(EDGE, t4["vec"], dist)
    st4.run(*[t4[fd["name"]] for fd in mod4.fields])
    torch.cuda.synchronize()
    no = p2.n_owned[EDGE]
    v = p2.mesh.valid(EDGE)[:no]
    for n in ("nabla2_vec", "nabla4_vec"):
        ok = ok and bit_equal(t4[n].cpu().numpy()[:, :no][:, v], want4[n][:, p2.global_ids[EDGE][:no]][:, v])
    res = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_PARTITION_OK" if int(res) == 1 else "MULTIGPU_PARTITION_MISMATCH", "ranks", world)
    dist.destroy_process_group()
    sys.exit(0 if int(res) == 1 else 1)


if __name__ == "__main__":
    main()
