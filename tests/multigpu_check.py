"""Multi-GPU parity check, launched under torchrun (one rank per GPU): j-slab sharded
hori_diff_type2 with the NCCL halo exchange of the C-ABI runtime must reproduce the oracle's result
on the undivided domain, bit for bit.   torchrun --nproc-per-node N tests/multigpu_check.py"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import HALO, hori_diff_type2_inputs, run_oracle  # noqa: E402

import dawn_b200  # noqa: E402
from dawn_b200 import runtime as rt, slabs  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    I, J, K = 70, 16 * world + 5, 9
    f = hori_diff_type2_inputs(I, J, K)
    dom = (I + 2 * HALO, J + 2 * HALO, K)
    want = run_oracle("hori_diff_type2_stencil", dom, f)["out"]
    sl = slabs.decompose(J, world, rank, HALO)
    rows = slice(sl.j0, sl.j0 + sl.nj_local + 2 * HALO)
    loc = {k: (v[:, rows, :].copy() if v.shape[1] > 1 else v.copy()) for k, v in f.items()}
    if sl.has_lower:
        loc["in"][:, :HALO, :] = np.nan
    if sl.has_upper:
        loc["in"][:, -HALO:, :] = np.nan
    mod = dawn_b200.load_stencil("hori_diff_type2_stencil")
    ni, njl = dom[0], sl.nj_local + 2 * HALO
    d = dawn_b200.Domain(ni, njl, K)
    st = mod(d)
    stor = {fd["name"]: dawn_b200.Storage(ni, njl, K + 1, fd["dims"]).from_numpy(loc[fd["name"]])
            for fd in mod.fields}
    idbuf = (ctypes.c_char * 128)()
    if rank == 0:
        rt.check(rt.lib().b200rt_nccl_unique_id(idbuf))
    obj = [bytes(idbuf)]
    dist.broadcast_object_list(obj, src=0)
    idbuf = (ctypes.c_char * 128).from_buffer_copy(obj[0])
    comm = ctypes.c_void_p()
    rt.check(rt.lib().b200rt_comm_init_rank(ctypes.byref(comm), world, idbuf, rank))
    widths = slabs.exchange_widths(mod.fields, st.written_fields())
    stream = torch.cuda.current_stream().cuda_stream
    for name, (wm, wp) in widths.items():
        s = stor[name]
        plan = ctypes.c_void_p()
        rt.check(rt.lib().b200rt_halo_plan_create(ctypes.byref(plan), comm, rank, world, ni, njl, K,
                                                  s.stride_j, s.stride_k, HALO, wm, wp))
        rt.check(rt.lib().b200rt_halo_exchange(plan, s.buf.data_ptr() + 8 * s.lead, stream))
        rt.check(rt.lib().b200rt_halo_plan_destroy(plan))
    st.run(*[stor[fd["name"]] for fd in mod.fields])
    torch.cuda.synchronize()
    got = stor["out"].to_numpy()
    ok = np.array_equal(got[:K, HALO:-HALO, :], want[:K, HALO + sl.j0:HALO + sl.j0 + sl.nj_local, :],
                        equal_nan=True)
    res = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_OK" if int(res) == 1 else "MULTIGPU_MISMATCH", "ranks", world)
    rt.lib().b200rt_comm_destroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if int(res) == 1 else 1)


if __name__ == "__main__":
    main()
