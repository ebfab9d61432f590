"""Times the pieces of the multi-GPU step separately (torchrun, >=2 ranks)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import dawn_b200
from dawn_b200 import runtime as rt
HALO = 3
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
I = J = 1024; K = 80; ni, nj = I + 6, J + 6
mod = dawn_b200.load_stencil("hori_diff_type2_stencil")
dom = dawn_b200.Domain(ni, nj, K); st = mod(dom)
stor = [dawn_b200.Storage(ni, nj, K + 1, f["dims"]) for f in mod.fields]
for s in stor: s.buf.uniform_(0.5, 1.5)
idbuf = (ctypes.c_char * 128)()
if rank == 0: rt.check(rt.lib().b200rt_nccl_unique_id(idbuf))
obj = [bytes(idbuf)]; dist.broadcast_object_list(obj, src=0)
idbuf = (ctypes.c_char * 128).from_buffer_copy(obj[0])
comm = ctypes.c_void_p(); rt.check(rt.lib().b200rt_comm_init_rank(ctypes.byref(comm), world, idbuf, rank))
s_in = stor[1]; plan = ctypes.c_void_p()
rt.check(rt.lib().b200rt_halo_plan_create(ctypes.byref(plan), comm, rank, world, ni, nj, K, s_in.stride_j, s_in.stride_k, HALO, 2, 2))
def timeit(fn, n=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host
stream = torch.cuda.current_stream().cuda_stream
exch = lambda: rt.check(rt.lib().b200rt_halo_exchange(plan, s_in.buf.data_ptr() + 8 * s_in.lead, stream))
full = lambda: st.run(*stor)
inter = lambda: st.run(*stor, rows=(2, J - 2))
strips = lambda: (st.run(*stor, rows=(0, 2)), st.run(*stor, rows=(J - 2, J)))
for name, fn in (("full kernel", full), ("interior rows", inter), ("two strips", strips), ("halo exchange", exch),
                 ("exchange then full (sequential)", lambda: (exch(), full()))):
    ms, host = timeit(fn)
    if rank == 0: print("%-36s gpu %.4f ms   host %.4f ms/iter" % (name, ms, host), flush=True)
dist.destroy_process_group()
