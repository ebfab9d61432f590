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
    print("rank", rank, "communicator ready", flush=True)
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
    if os.environ.get("MULTIGPU_QUICK") == "1":
        # short form for a tight GPU budget: base parity (above), the peer-memory exchange, and the time
        # of one exchange of either kind at the bench shape
        ok_p2p = p2p_check(rank, world, comm, stream, st, mod, stor, loc, widths, want, sl, (ni, njl, K))
        res4 = torch.tensor([1 if ok_p2p else 0], device="cuda")
        dist.all_reduce(res4, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("MULTIGPU_P2P_OK" if int(res4) == 1 else "MULTIGPU_P2P_MISMATCH", "ranks", world, flush=True)
        if int(res4) == 1:
            exchange_times(rank, world, comm)
        rt.lib().b200rt_comm_destroy(comm)
        dist.destroy_process_group()
        sys.exit(0 if int(res) == 1 and int(res4) == 1 else 1)
    # ---- the same step (NCCL halo exchange + kernel) recorded once as a CUDA graph and replayed ------
    graph_ok = True
    import threading
    # a replayed NCCL send without its peer never returns: leave the process instead of hanging the box
    watchdog = threading.Timer(90.0, lambda: (print("rank", rank, "graph step timed out", flush=True),
                                              os._exit(3)))
    watchdog.daemon = True
    watchdog.start()
    try:
        plans = []
        step = dawn_b200.TimeStep()
        for name, (wm, wp) in widths.items():
            s = stor[name]
            plan = ctypes.c_void_p()
            rt.check(rt.lib().b200rt_halo_plan_create(ctypes.byref(plan), comm, rank, world, ni, njl, K,
                                                      s.stride_j, s.stride_k, HALO, wm, wp))
            plans.append(plan)
            step.exchange(plan, s.buf.data_ptr() + 8 * s.lead)
        step.run(st, *[stor[fd["name"]] for fd in mod.fields])
        print("rank", rank, "capturing the step", flush=True)
        captured = True
        try:
            step.capture()
        except Exception as e:  # noqa: BLE001
            print("rank", rank, "graph capture failed:", e, flush=True)
            captured = False
        # every rank must have a graph before anyone replays (a replayed ncclSend without its peer
        # would wait forever)
        flag = torch.tensor([1 if captured else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        print("rank", rank, "captured:", captured, flush=True)
        if int(flag) != 1:
            raise RuntimeError("graph capture failed on some rank")
        for rep in range(3):
            stor["in"].from_numpy(loc["in"])          # halo rows are NaN again
            stor["out"].from_numpy(loc["out"])        # -1 everywhere, as in the oracle's input
            torch.cuda.synchronize()
            if rep == 0:
                step.eager()                          # the same sequence, launch by launch
            else:
                step.replay()
            torch.cuda.synchronize()
            got = stor["out"].to_numpy()
            a = got[:K, HALO:-HALO, :]
            b = want[:K, HALO + sl.j0:HALO + sl.j0 + sl.nj_local, :]
            same = np.array_equal(a, b, equal_nan=True)
            if not same:
                bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
                print("rank", rank, "rep", rep, "mismatches", int(bad.sum()), "nan", int(np.isnan(a).sum()),
                      "untouched", int((a == -5.0).sum()), "rows", np.unique(np.nonzero(bad)[1])[:8],
                      flush=True)
            graph_ok = graph_ok and same
        step.close()                                  # graphs holding NCCL nodes go before the communicator
        for plan in plans:
            rt.check(rt.lib().b200rt_halo_plan_destroy(plan))
    except Exception as e:  # noqa: BLE001
        print("rank", rank, "graph step failed:", e, flush=True)
        graph_ok = False
    watchdog.cancel()
    res3 = torch.tensor([1 if graph_ok else 0], device="cuda")
    dist.all_reduce(res3, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_GRAPH_OK" if int(res3) == 1 else "MULTIGPU_GRAPH_MISMATCH", "ranks", world)
    ok_p2p = p2p_check(rank, world, comm, stream, st, mod, stor, loc, widths, want, sl, (ni, njl, K))
    res4 = torch.tensor([1 if ok_p2p else 0], device="cuda")
    dist.all_reduce(res4, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_P2P_OK" if int(res4) == 1 else "MULTIGPU_P2P_MISMATCH", "ranks", world)
    ok_icon = icon_check(rank, world, comm, stream)
    res2 = torch.tensor([1 if ok_icon else 0], device="cuda")
    dist.all_reduce(res2, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_ICON_OK" if int(res2) == 1 else "MULTIGPU_ICON_MISMATCH", "ranks", world)
    rt.lib().b200rt_comm_destroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if int(res) == 1 and int(res2) == 1 and int(res3) == 1 else 1)


def p2p_check(rank, world, comm, stream, st, mod, stor, loc, widths, want, sl, dims):
    """The peer-memory halo exchange (b200rt_halo_exchange_p2p): three exchanges in a row - the second and
    third need the credit returned by the neighbour's unpack - each followed by the kernel, bit-exact."""
    ni, njl, K = dims
    ok = True
    try:
        def gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        plans = []
        for name, (wm, wp) in widths.items():
            s = stor[name]
            plan = ctypes.c_void_p()
            rt.check(rt.lib().b200rt_halo_plan_create(ctypes.byref(plan), comm, rank, world, ni, njl, K,
                                                      s.stride_j, s.stride_k, HALO, wm, wp))
            rt.connect_halo_plan(plan, rank, world, gather)
            plans.append((plan, s))
        dist.barrier()
        for rep in range(3):
            stor["in"].from_numpy(loc["in"])          # halo rows are NaN again
            stor["out"].from_numpy(loc["out"])
            torch.cuda.synchronize()
            for plan, s in plans:
                rt.check(rt.lib().b200rt_halo_exchange_p2p(plan, s.buf.data_ptr() + 8 * s.lead, stream))
            st.run(*[stor[fd["name"]] for fd in mod.fields])
            torch.cuda.synchronize()
            a = stor["out"].to_numpy()[:K, HALO:-HALO, :]
            b = want[:K, HALO + sl.j0:HALO + sl.j0 + sl.nj_local, :]
            same = np.array_equal(a, b, equal_nan=True)
            if not same:
                print("rank", rank, "p2p rep", rep, "mismatches", int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum()),
                      flush=True)
            ok = ok and same
        for plan, _ in plans:
            rt.check(rt.lib().b200rt_halo_plan_p2p_status(plan))
        dist.barrier()                                # nobody unmaps a buffer a neighbour may still write
        for plan, _ in plans:
            rt.check(rt.lib().b200rt_halo_plan_destroy(plan))
    except Exception as e:  # noqa: BLE001
        print("rank", rank, "p2p exchange failed:", e, flush=True)
        ok = False
    return ok


def exchange_times(rank, world, comm, reps=200):
    """Device time of one halo exchange of `in` at the bench shape (1030 x 1030 x 80 storage, 2 rows each
    way), NCCL (pack, ncclSend/Recv, unpack) against peer memory (push, wait + unpack), max over ranks."""
    ni = nj = 1024 + 2 * HALO
    K = 80
    s = dawn_b200.Storage(ni, nj, K + 1, "ijk")
    s.buf.uniform_(0.0, 1.0)
    plan = ctypes.c_void_p()
    rt.check(rt.lib().b200rt_halo_plan_create(ctypes.byref(plan), comm, rank, world, ni, nj, K, s.stride_j,
                                              s.stride_k, HALO, 2, 2))

    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    rt.connect_halo_plan(plan, rank, world, gather)
    stream = torch.cuda.current_stream().cuda_stream
    ptr = s.buf.data_ptr() + 8 * s.lead
    for label, fn in (("nccl", rt.lib().b200rt_halo_exchange), ("p2p", rt.lib().b200rt_halo_exchange_p2p)):
        for _ in range(10):
            rt.check(fn(plan, ptr, stream))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            rt.check(fn(plan, ptr, stream))
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("EXCHANGE_US %s %.2f (back-to-back exchanges, %d ranks)" % (label, float(t), world), flush=True)
    rt.check(rt.lib().b200rt_halo_plan_p2p_status(plan))
    dist.barrier()
    rt.check(rt.lib().b200rt_halo_plan_destroy(plan))


def icon_check(rank, world, comm, stream):
    """ICON nabla2 on row slabs of a TriMesh: 1 halo quad row, `vec` exchanged over NCCL, owned edges
    must equal the oracle's result on the undivided mesh bit for bit."""
    from oracle import naive_interp
    from util import iir_path
    from dawn_b200.trimesh import CELL, EDGE, VERTEX, TriMesh
    nx, ny, K = 9, 5 * world + 2, 4
    gm = TriMesh(nx, ny)
    rng = np.random.default_rng(77)
    locs = {"vec": EDGE, "div_vec": CELL, "rot_vec": VERTEX, "nabla2_vec": EDGE,
            "primal_edge_length": EDGE, "dual_edge_length": EDGE, "tangent_orientation": EDGE,
            "geofac_rot": VERTEX, "geofac_div": CELL}
    G = {n: rng.uniform(0.5, 1.5, (K, gm.num(l))) for n, l in locs.items() if not n.startswith("geofac")}
    G["geofac_rot"] = rng.uniform(-1, 1, (K, gm.num_vertices, 6))
    G["geofac_div"] = rng.uniform(-1, 1, (K, gm.num_cells, 3))
    want = {n: v.copy() for n, v in G.items()}
    want["__tmp_nab_60"] = np.zeros((K, gm.num_edges))
    want["__tmp_nab_61"] = np.zeros((K, gm.num_edges))
    naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), gm, K, want)
    sl = slabs.mesh_slab(nx, ny, world, rank)
    lm = TriMesh(nx, sl.local_ny)
    mod = dawn_b200.load_stencil("ICON_laplacian_stencil")
    st = mod.on_mesh(lm, K)
    tens = {}
    for fd in mod.fields:
        l = locs[fd["name"]]
        per = sl.slots_per_row(l)
        a = G[fd["name"]][:, sl.row_offset * per: sl.row_offset * per + lm.num(l)].copy()
        if fd["name"] == "vec":  # halo rows must come from the neighbours
            if sl.halo_lo:
                a[:, :per] = np.nan
            if sl.halo_hi:
                a[:, (sl.local_ny - 1) * per: sl.local_ny * per] = np.nan
        if fd["sparse"]:
            a = np.ascontiguousarray(a.transpose(0, 2, 1))
        tens[fd["name"]] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    per = sl.slots_per_row(EDGE)
    plan = ctypes.c_void_p()
    rt.check(rt.lib().b200rt_halo_plan_create2(ctypes.byref(plan), comm, rank, world, per, sl.local_ny,
                                               K, per, lm.num_edges, sl.halo_lo, sl.halo_hi, 1, 1))
    rt.check(rt.lib().b200rt_halo_exchange(plan, tens["vec"].data_ptr(), stream))
    st.run(*[tens[fd["name"]] for fd in mod.fields])
    torch.cuda.synchronize()
    rt.check(rt.lib().b200rt_halo_plan_destroy(plan))
    got = tens["nabla2_vec"].cpu().numpy()
    o, g = sl.owned_slice(EDGE), sl.global_slice(EDGE)
    valid = gm.edge_valid[g]
    return bool(np.array_equal(got[:, o][:, valid], want["nabla2_vec"][:, g][:, valid]))


if __name__ == "__main__":
    main()
