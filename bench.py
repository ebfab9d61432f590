#!/usr/bin/env python
"""bench.py — headline benchmark: horizontal diffusion (hori_diff_type2_stencil, 4th-order lap +
flux limiter) on 1024x1024x80 doubles per GPU, grid-point updates/s and fraction of HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload hori_diff|tridiagonal]

N>1 is launched with torchrun (one rank per GPU): the global domain is 1024 x (1024*N) x 80 sharded
into j-slabs (weak scaling); every step exchanges the 2-row j halos of `in` with the slab neighbours
over NCCL and then runs the generated kernel.  Rank 0 prints ONE JSON line.

Timing: CUDA events on the launching stream, W untimed + exactly K timed steps between barriers,
max over ranks.  The three 0.68 GB fields of a step (2.0 GB) exceed the 126 MB L2, so no L2 flush is
needed between steps.  The oracle (oracle/) is used only for the `cpu_baseline` leg and for
`--impl reference`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

HALO = 3
BYTES_PER_POINT = {"hori_diff": 24, "tridiagonal": 48}  # SURVEY.md §8(d), DESIGN.md §4
STENCIL = {"hori_diff": "hori_diff_type2_stencil", "tridiagonal": "tridiagonal_solve"}
SHAPE = {"hori_diff": (1024, 1024, 80), "tridiagonal": (512, 512, 80)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region: nvidia-smi runs from before
    the warm-up with 10 ms period and its lines are collected by a reader thread as they arrive; samples are
    selected by their timestamps against the wall-clock bracket of the timed region.  A region of a few
    milliseconds can end before nvidia-smi has printed anything: the caller then keeps the same steps running
    (untimed) in small batches until a sample has arrived (`wants_more_load`), and that is said in `window`."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BEFORE, AFTER = 0.25, 0.05      # s around a region too short to hold a sample (the GPU is under the same load)

    def __init__(self, index=0, command="nvidia-smi"):
        self.proc = None
        self.index = index
        self.command = command
        self.t0 = self.t1 = self.t_same = None
        self.rows = []          # (timestamp, sm MHz, max sm MHz, [active reasons]); appended by the reader thread

    def start(self):
        import threading
        try:
            self.proc = subprocess.Popen(
                [self.command, "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "10"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except Exception:
            self.proc = None
            return
        self.reader = threading.Thread(target=self._pump, daemon=True)
        self.reader.start()

    def _pump(self):
        import datetime
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                self.rows.append((ts, float(parts[1]), float(parts[2]),
                                  [n for n, v in zip(self.NAMES, parts[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def _select(self):
        rows = list(self.rows)
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else 1e18
        sel = [r for r in rows if t0 <= r[0] <= t1]
        if sel:
            return sel, "timed region"
        sel = [r for r in rows if t0 - self.BEFORE <= r[0] <= t1 + self.AFTER]
        if sel:
            return sel, "timed region -%.2f s .. +%.2f s (same load)" % (self.BEFORE, self.AFTER)
        if self.t_same is not None:
            sel = [r for r in rows if t0 - self.BEFORE <= r[0] <= self.t_same]
            if sel:
                return sel, ("timed region -%.2f s .. +%.2f s (the same steps kept running, untimed, until "
                             "nvidia-smi had printed a sample)" % (self.BEFORE, self.t_same - t1))
        return [], "no sample"

    def wants_more_load(self, limit_s=1.5):
        """after mark_end(): True while no sample describes the timed region yet (and nvidia-smi is alive and the
        extension is younger than `limit_s`); the caller runs another small batch of the same steps and asks again"""
        if not self.proc or self.proc.poll() is not None or self.t1 is None:
            return False
        now = time.time()
        if now - self.t1 > limit_s:
            return False
        self.t_same = now
        return not self._select()[0]

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.reader.join(timeout=2)
        sel, window = self._select()
        sm = sorted(r[1] for r in sel)
        reasons = sorted({x for r in sel for x in r[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max((r[2] for r in sel), default=None), "samples": len(sel),
                "window": window, "reasons": reasons}


# --------------------------------------------------------------------------------------------------
# CPU legs (oracle): cpu_baseline and --impl reference
# --------------------------------------------------------------------------------------------------
def cpu_inputs(workload, I, J, K):
    from util import hori_diff_type2_inputs, tridiagonal_inputs
    return hori_diff_type2_inputs(I, J, K) if workload == "hori_diff" else tridiagonal_inputs(I, J, K)


def cpu_run(workload, f, K, threads):
    from oracle import capi
    if workload == "hori_diff":
        capi.hori_diff_type2(f["out"], f["in"], f["crlato"].reshape(-1), f["crlatu"].reshape(-1),
                             f["hdmask"], nk=K, threads=threads)
    else:
        capi.tridiagonal_solve(f["d"], f["a"], f["b"], f["c"], nk=K, threads=threads)


def cpu_baseline(workload):
    """The reference's naive algorithm (oracle/naive_stencils.c, loop shape of CXXNaiveCodeGen.cpp:
    380-495), single thread exactly as generated, full per-GPU domain, 1 warm-up + 3 timed runs."""
    I, J, K = SHAPE[workload]
    f = cpu_inputs(workload, I, J, K)
    ts = []
    for rep in range(4):
        g = {k: v.copy() for k, v in f.items()}
        t = time.perf_counter()
        cpu_run(workload, g, K, 1)
        ts.append(time.perf_counter() - t)
    ts = sorted(ts[1:])
    med = ts[len(ts) // 2]
    return {"value": I * J * K / med, "unit": "grid-point updates/s", "cores": 1, "kind": "port",
            "sample": "full %dx%dx%d domain, naive i-outer/j-inner loops, 1 warm-up + 3 timed runs, "
                      "median %.3f s; g++ -O3 -ffp-contract=off" % (I, J, K, med)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    I, J, K = SHAPE[workload]
    # every host thread this process may use; the hori_diff sample is split over k, so it takes at least
    # as many levels as there are threads (bounded sample: 16 levels on a 16-thread host, never > K)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    Ks = min(K, max(16, threads))
    if workload == "hori_diff":
        threads = min(threads, Ks)
    f = cpu_inputs(workload, I, J, Ks)
    for _ in range(args.warmup):
        cpu_run(workload, f, Ks, threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(workload, f, Ks, threads)
    dt = time.perf_counter() - t
    val = I * J * Ks * args.steps / dt
    sample = ("each step = levels [0,%d) of the %dx%dx%d domain (%d of %d levels), k split over %d "
              "pthreads" % (Ks, I, J, K, Ks, K, threads))
    if workload == "tridiagonal":
        sample = "each step = full %dx%dx%d solve, i columns split over %d pthreads" % (I, J, Ks, threads)
    line = {
        "impl": "reference", "metric": "grid-point updates/s", "value": val,
        "unit": "grid-point updates/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s %dx%dx%d double (reference naive C++ algorithm on host cores)" % (
            STENCIL[workload], I, J, K),
                   "sample": "a step of this arm is a bounded SAMPLE of the workload, the value is a rate: " + sample},
        "cpu_baseline": {"value": val, "unit": "grid-point updates/s", "cores": threads,
                         "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "grid-point updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU leg
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch

    import dawn_b200
    from dawn_b200 import runtime as rt

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    workload = args.workload
    I, J, K = SHAPE[workload]
    name = STENCIL[workload]
    ni, nj = I + 2 * HALO, J + 2 * HALO
    opts = {}
    for kv in args.opt:
        k, v = kv.split("=")
        opts[k] = int(v)
    mod = dawn_b200.load_stencil(name, **opts)
    dom = dawn_b200.Domain(ni, nj, K)
    dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
    st = mod(dom)

    # synthetic inputs: fillMath closed form of the reference drivers, generated on the device for
    # this rank's slab of the global domain (global j offset = rank * J)
    stor = {}
    es = 4 if mod.dtype == "float32" else 8      # ::dawn::float_type of the module (--opt single_precision=1)
    if es == 4 and world > 1:
        raise SystemExit("single precision is a 1-GPU bench line: the halo plans of the runtime move doubles")
    for fd in mod.fields:
        stor[fd["name"]] = dawn_b200.Storage(ni, nj, K + 1, fd["dims"], dtype=mod.dtype)
    from util import FILL
    keymap = {"in": "u", "out": None}

    def dev_fill(s, consts, gj0, gnj):
        a, b, c, d, e, f = consts
        nk_, nj_, ni_ = s.shape[2], s.shape[1], s.shape[0]
        ii = torch.arange(ni_, device="cuda", dtype=torch.float64) / (ni_ if "i" in s.dims else 1)
        jj = (torch.arange(nj_, device="cuda", dtype=torch.float64) + gj0) / gnj
        kk = torch.arange(nk_, device="cuda", dtype=torch.float64)
        x = ii[None, None, :] if "i" in s.dims else torch.zeros(1, 1, 1, device="cuda", dtype=torch.float64)
        y = jj[None, :, None] if "j" in s.dims else torch.zeros(1, 1, 1, device="cuda", dtype=torch.float64)
        v = kk[:, None, None] * 10e-3 + a * (b + torch.cos(np.pi * (x + c * y)) +
                                             torch.sin(d * np.pi * (x + e * y))) / f
        s.view().copy_(v.expand(nk_, nj_, ni_))

    gnj = J * world + 2 * HALO

    def fill_inputs():
        for fd in mod.fields:
            s = stor[fd["name"]]
            key = keymap.get(fd["name"], fd["name"])
            if key is None:
                s.fill(-1.0)
            else:
                dev_fill(s, FILL[key], rank * J, gnj)
                if workload == "tridiagonal" and fd["name"] == "b":
                    s.buf.add_(12.0)
        torch.cuda.synchronize()

    fill_inputs()
    fields = [stor[fd["name"]] for fd in mod.fields]

    # halo exchange plan for every input with a j extent (multi-GPU only)
    plans = []
    comm = None
    widths = {}
    if world > 1:
        import ctypes
        idbuf = (ctypes.c_char * 128)()
        if rank == 0:
            rt.check(rt.lib().b200rt_nccl_unique_id(idbuf))
        obj = [bytes(idbuf)]
        dist.broadcast_object_list(obj, src=0)
        idbuf = (ctypes.c_char * 128).from_buffer_copy(obj[0])
        comm = ctypes.c_void_p()
        rt.check(rt.lib().b200rt_comm_init_rank(ctypes.byref(comm), world, idbuf, rank))
        from dawn_b200 import slabs
        widths = slabs.exchange_widths(mod.fields, st.written_fields())
        for fd in mod.fields:
            if fd["name"] in widths:
                jm, jp = widths[fd["name"]]
                s = stor[fd["name"]]
                plan = ctypes.c_void_p()
                rt.check(rt.lib().b200rt_halo_plan_create(
                    ctypes.byref(plan), comm, rank, world, ni, nj, K, s.stride_j, s.stride_k, HALO,
                    jm, jp))
                plans.append((plan, s))
        if args.halo == "p2p":
            def gather(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out
            for plan, _ in plans:
                rt.connect_halo_plan(plan, rank, world, gather)
            dist.barrier()

    stream = torch.cuda.current_stream().cuda_stream
    launches_per_step = [0]
    # rows whose stencil footprint reaches into a slab neighbour's halo: computed after the exchange
    dep_lo = max([w[0] for w in widths.values()], default=0) if world > 1 else 0
    dep_hi = max([w[1] for w in widths.values()], default=0) if world > 1 else 0
    comm_stream = torch.cuda.Stream(priority=-1) if plans else None
    ev_ready, ev_halo = torch.cuda.Event(), torch.cuda.Event()

    def step():
        """One hot-path pass.  Multi-GPU: the halo exchange (pack kernel -> ncclSend/Recv -> unpack
        kernel) runs on its own stream while the rows that do not depend on halos are computed; the
        two boundary strips follow once the halos have landed."""
        if not plans:
            st.run(*fields)
            launches_per_step[0] = 0
            return
        main = torch.cuda.current_stream()
        ev_ready.record(main)                      # previous step done with `in`
        comm_stream.wait_event(ev_ready)
        n = 0
        for plan, s in plans:
            if args.halo == "p2p":  # pack + NVLink stores in one kernel, wait + unpack in a second one
                rt.check(rt.lib().b200rt_halo_exchange_p2p(plan, s.buf.data_ptr() + 8 * s.lead,
                                                           comm_stream.cuda_stream))
                n += 2
            else:
                rt.check(rt.lib().b200rt_halo_exchange(plan, s.buf.data_ptr() + 8 * s.lead,
                                                       comm_stream.cuda_stream))
                n += (2 if rank > 0 else 0) + (2 if rank + 1 < world else 0)
        # the strips are whole tiles of the kernel (module info "tile"): a 2-row strip would still occupy - and
        # load the inputs of - a full 12-row tile, work the interior launch then repeats
        tile_j = 0 if os.environ.get("B200_BENCH_THIN_STRIPS") == "1" else (mod.info.get("tile") or [0, 0])[1]
        lo = max(dep_lo, tile_j) if rank > 0 else 0
        hi = J - (max(dep_hi, tile_j) if rank + 1 < world else 0)
        if hi - lo < tile_j:
            lo, hi = (dep_lo if rank > 0 else 0), J - (dep_hi if rank + 1 < world else 0)
        # boundary strips follow the exchange on the (high-priority) communication stream ...
        if lo > 0:
            st.run(*fields, rows=(0, lo), stream=comm_stream.cuda_stream)
        if hi < J:
            st.run(*fields, rows=(hi, J), stream=comm_stream.cuda_stream)
        ev_halo.record(comm_stream)
        # ... while the rows that need no halo are computed on the main stream
        st.run(*fields, rows=(lo, hi))
        main.wait_event(ev_halo)
        launches_per_step[0] = n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps ------------------------------------------------------------------
    barrier()
    sampler.mark_begin()
    l0 = st.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        # the host barrier leaves the ranks' host threads tens of microseconds apart; a device-side all-reduce
        # in front of the first event lines the GPUs' streams up, so the timed region starts on all of them at once
        dist.all_reduce(torch.zeros(1, device="cuda"))
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = (st.launches() - l0) + launches_per_step[0] * args.steps
    if args.halo == "p2p":
        for plan, _ in plans:  # a wait that gave up means halos were stale: not a valid run
            rt.check(rt.lib().b200rt_halo_plan_p2p_status(plan))
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # a region of a few milliseconds can end before nvidia-smi has printed a line: the same steps keep running
    # (untimed, ~20 ms batches, the same count on every rank) until a sample describes this load
    batch = int(20.0 / max(ms_max / args.steps, 1e-3)) + 1
    while True:
        more = torch.tensor([1.0 if rank == 0 and sampler.wants_more_load() else 0.0], device="cuda")
        if dist is not None:
            dist.all_reduce(more, op=dist.ReduceOp.MAX)
        if float(more.item()) == 0.0:
            break
        for _ in range(batch):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity self-check: the launch shape that was just timed, against the oracle ----------------
    parity = parity_check(workload, mod, stor, fill_inputs, step, (I, J, K), rank, world,
                          max([w[0] for w in widths.values()], default=0),
                          max([w[1] for w in widths.values()], default=0))
    flag = torch.tensor([0 if parity["max_abs_diff"] == 0.0 and parity["differing_values"] == 0 else 1],
                        device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if float(flag.item()) != 0.0:
        # a fast kernel with different results is not a result: no value is printed
        sys.stderr.write("bench.py: rank %d parity check against the oracle: %s\n" % (rank, json.dumps(parity)))
        raise SystemExit(3)

    # ---- kernel-only time for the roofline: at N=1 a step IS one launch of the generated kernel, so
    # the timed region itself gives its average duration; at N>1 the same number of bare kernel
    # launches is timed right after the timed region (same thermal / power-cap state)
    if world == 1:
        kernel_ms = ms / args.steps
        kernels_per_run = (st.launches() - l0) // args.steps
    else:
        lk = st.launches()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(args.steps):
            st.run(*fields)
        ev[1].record()
        torch.cuda.synchronize()
        kernel_ms = ev[0].elapsed_time(ev[1]) / args.steps
        kernels_per_run = (st.launches() - lk) // args.steps

    # ---- end to end through the public API with HOST buffers --------------------------------------
    e2e_steps = min(args.steps, 5)
    from dawn_b200 import runtime as rt_
    with rt_.near_device(local) as near:  # pinned staging buffers on the GPU's own NUMA node
        host = [torch.empty(s.shape[2], s.shape[1], s.shape[0], dtype=s.dtype).pin_memory()
                for s in fields]
    for h, s in zip(host, fields):
        h.copy_(s.view())
    st2 = mod(dom)
    # inputs go up and the result comes down every step; `out` is write-only, so it is not an input
    # (uploaded once by the warm-up call); levels are independent, so the domain is streamed through
    # the device in chunks of 8 levels with H2D, kernels and D2H overlapping
    e2e_kw = dict(upload_outputs=False, k_chunk=8)
    h2d, d2h = st2.run_from_host(*host)  # warm-up (allocates the device-side storages)
    h2d, d2h = st2.run_from_host(*host, **e2e_kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h2d, d2h = st2.run_from_host(*host, **e2e_kw)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    pts = I * J * K
    # the same end-to-end step through the generated C++ CLASS (what a Dawn user's driver calls): storages with
    # pinned host mirrors, run() with its host <-> device pipeline, result read on the host
    e2e_cpp = None
    if rank == 0 and world == 1 and workload == "hori_diff" and not args.opt:
        e2e_cpp = cpp_class_e2e(I, J, K, e2e_steps)
    if rank == 0:
        peak, peak_src = measured_peak()
        bpp = BYTES_PER_POINT[workload] * es // 8
        achieved = bpp * pts / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)
        if os.path.exists(tp) and not args.shape and es == 8:  # the committed ncu capture is of the default shape
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": "grid-point updates/s", "value": pts * world * args.steps / (ms_max * 1e-3),
            "unit": "grid-point updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if es == 8 else "f32",
            "data": "synthetic (fillMath closed form of the reference drivers, generated on device)",
            "config": {
                "workload": "%s %dx%dx%d %s per GPU (halo 3, storages %dx%dx%d)%s" % (
                    name, I, J, K, "double" if es == 8 else "float", ni, nj, K + 1,
                    "" if world == 1 else ", global %dx%dx%d in %d j-slabs, %s halo exchange of "
                    "`in` (2 rows each way) every step, overlapped with the interior rows" % (
                        I, J * world, K, world,
                        "peer-memory (NVLink stores into the neighbour's landing buffer, device-side "
                        "flags)" if args.halo == "p2p" else "NCCL")),
                "l2": "inputs larger than L2 (%.1f GB per step vs 126 MB), no flush" % (bpp * pts / 1e9),
                "timing": "%d back-to-back steps = %.1f ms timed (%s); at N=1 roofline.achieved uses this "
                          "same region (a step is one kernel launch) - see clocks" % (
                              args.steps, ms_max,
                              "sustained figure" if ms_max >= 50.0 else
                              "a burst figure: shorter than the ~50 ms it takes the power cap to engage"),
                "codegen_options": opts,
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": None if traffic is None else
                "committed ncu capture profiles/traffic_%s.json (dram__bytes_read+write per launch), not "
                "measured in this run" % workload,
                "kernel": "%s (generated), %d launch(es)/step, %.4f ms/launch, %d algorithmic B/point"
                          % (name, kernels_per_run, kernel_ms / max(1, kernels_per_run), bpp),
                "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
            },
            "e2e": {"value": pts * world * e2e_steps / e2e_s, "unit": "grid-point updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "numa_local_host_buffers": bool(near.cpus),
                    "api": "dawn_b200.Stencil.run_from_host(pinned host buffers, upload_outputs=False, "
                           "k_chunk=8): inputs H2D + result D2H every step, pipelined over k chunks"},
            "e2e_python_mirror": None,
            "gpu_launches": launches,
            "clocks": clocks,
            "parity_checked": True, "max_abs_diff": parity["max_abs_diff"], "parity": parity,
            "last_launch": st.last_launch(),
        }
        if e2e_cpp is not None and "updates_per_s" in e2e_cpp and e2e_cpp.get("differing_values") == 0:
            # headline e2e = the generated C++ class; the Python mirror's figure is kept beside it
            line["e2e_python_mirror"] = line["e2e"]
            line["e2e"] = {"value": e2e_cpp["updates_per_s"], "unit": "grid-point updates/s",
                           "h2d_bytes_per_step": int(e2e_cpp["h2d_bytes_per_step"]),
                           "d2h_bytes_per_step": int(e2e_cpp["d2h_bytes_per_step"]), "steps": e2e_cpp["steps"],
                           "ms_per_step": e2e_cpp["seconds_per_step"] * 1e3,
                           "verified_against_oracle": True,
                           "api": "generated C++ class dawn_generated::b200::hori_diff_type2_stencil (the reference's "
                                  "class API): storages with pinned host mirrors, every step the host writes the four "
                                  "inputs' host views, run() streams them through the device in chunks of %d levels "
                                  "(set_host_pipeline) and leaves the result in the host mirror, the driver reads it "
                                  "there (tests/cpp/hori_diff_e2e.cu)" % e2e_cpp["chunk_levels"]}
        else:
            line.pop("e2e_python_mirror", None)
            if e2e_cpp is not None:
                line["e2e_cpp_class_error"] = e2e_cpp
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(workload)
            if not args.shape:
                # reported beside the CPU baseline, not a target either: what the reference's CUDA backend
                # output does on this GPU when it is merely recompiled
                line["reference_cuda_kernel"] = refcuda_baseline(workload)
                if workload == "hori_diff" and "ms_per_run" in line["reference_cuda_kernel"]:
                    # like for like: the b200 backend on the SAME stencil the reference text implements
                    line["reference_cuda_kernel"]["b200_same_stencil"] = same_stencil_as_reference_text()
        if world == 1 and workload == "hori_diff" and not args.shape and not args.opt and not args.no_sub_workloads:
            # the other two configurations the target names (BASELINE.json configs[2], [3]): measured in
            # the same run, each in its own process with its own roofline / clocks / cpu_baseline / e2e
            line["workloads"] = sub_workloads(args)
        print(json.dumps(line))
    for plan, _ in plans:
        rt.lib().b200rt_halo_plan_destroy(plan)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def parity_check(workload, mod, stor, fill_inputs, step, shape, rank, world, dep_lo, dep_hi):
    """One more step of the timed configuration on fresh inputs, compared bit for bit with the oracle
    (oracle/naive_stencils.c as the CHECKER) on a sample of the domain: hori_diff on whole k planes (levels
    are independent), the solver on bands of whole columns (columns are independent).  With slab neighbours
    the halo rows of the exchanged inputs are overwritten with NaN first, so a result that matches the
    oracle proves the exchange delivered them.  Returns a record for the JSON line."""
    import numpy as np
    import torch
    from oracle import capi
    I, J, K = shape
    fill_inputs()
    names = [fd["name"] for fd in mod.fields]
    single = mod.dtype == "float32"

    def interp(fields, nk):
        """single precision: the pinned interpreter on float32 storages (the C restatements are double)"""
        from oracle import naive_interp
        from util import iir_path
        some = next(iter(fields.values()))
        dims = (max(a.shape[2] for a in fields.values()), max(a.shape[1] for a in fields.values()), nk)
        return naive_interp.run_cartesian(iir_path(mod.name), dims, fields, halos=(HALO,) * 4 + (0, 0))
    if workload == "hori_diff":
        planes = sorted({0, 1, 19, 20, 21, K // 2, K - 2, K - 1} & set(range(K)))
        idx = torch.as_tensor(planes, device="cuda")
        take = lambda s: s.view().index_select(0, idx).cpu().numpy()  # noqa: E731
        before = {n: take(stor[n]) for n in ("in", "hdmask", "out")}
        vec = {n: stor[n].view().cpu().numpy().copy() for n in ("crlato", "crlatu")}
        if world > 1:   # poison what the exchange has to deliver
            v = stor["in"].view()
            if rank > 0:
                v[:, HALO - dep_lo:HALO, :] = float("nan")
            if rank + 1 < world:
                v[:, HALO + J:HALO + J + dep_hi, :] = float("nan")
        step()
        torch.cuda.synchronize()
        got = take(stor["out"])
        if single:
            want = interp({"out": before["out"].copy(), "in": before["in"], "crlato": vec["crlato"],
                           "crlatu": vec["crlatu"], "hdmask": before["hdmask"]}, len(planes))["out"]
        else:
            want = capi.hori_diff_type2(before["out"].copy(), before["in"], vec["crlato"].reshape(-1),
                                        vec["crlatu"].reshape(-1), before["hdmask"], halo=HALO, nk=len(planes))
        checks = [(got, want), (take(stor["in"]), before["in"])]
        what = "out on k planes %s (whole planes incl. halo ring)%s" % (
            planes, "; halo rows of `in` poisoned with NaN before the exchange" if world > 1 else "")
    else:
        bands = sorted({0, (J // 2 // 8) * 8, J - 8})
        rows = [r for b in bands for r in range(b, b + 8 + 2 * HALO)]  # band + its own halo rows
        idx = torch.as_tensor(rows, device="cuda")
        take = lambda s: s.view().index_select(1, idx).cpu().numpy()  # noqa: E731
        before = {n: take(stor[n]) for n in names}
        step()
        torch.cuda.synchronize()
        checks = []
        nb = 8 + 2 * HALO
        for q in range(len(bands)):
            sl = slice(q * nb, (q + 1) * nb)
            sub = {n: np.ascontiguousarray(before[n][:, sl, :]) for n in names}
            if single:
                r_ = interp(sub, K)
                wd, wc = r_["d"], r_["c"]
            else:
                wd, wc = capi.tridiagonal_solve(sub["d"], sub["a"], sub["b"], sub["c"], halo=HALO, nk=K)
            # only the band's own interior rows were computed from complete data on both sides
            for n, w in (("d", wd), ("c", wc)):
                g = take(stor[n])[:, sl, :]
                checks.append((g[:, HALO:-HALO, :], w[:, HALO:-HALO, :]))
        what = "d and c on 8-row bands of whole columns at interior rows %s" % bands
    ndiff, maxdiff, nvals = 0, 0.0, 0
    for g, w in checks:
        bad = g.view(np.int32 if single else np.int64) != w.view(np.int32 if single else np.int64)
        bad &= ~(np.isnan(g) & np.isnan(w))
        ndiff += int(bad.sum())
        nvals += g.size
        if bad.any():
            d = np.abs(g[bad] - w[bad])
            maxdiff = max(maxdiff, float(np.nanmax(d)) if np.isfinite(d).any() else float("inf"))
    return {"checked": what, "values_compared": nvals, "differing_values": ndiff, "max_abs_diff": maxdiff,
            "oracle": "oracle/naive_interp.py on float32 storages (single precision)" if single else
                      "oracle/naive_stencils.c (C restatement of the naive backend, bit-equal to the pinned "
                      "interpreter)"}


def run_icon(args):
    """ICON nabla2 (ICON_laplacian_stencil) on a synthetic triangle mesh: BASELINE.json configs[3]
    (~20 M edges x 65 levels).  Unit of work: edge-level updates.  Algorithmic bytes per level:
    40 B x E + 56 B x V + 32 B x C, neighbour tables (int32) once per run."""
    import numpy as np
    import torch

    import dawn_b200
    from dawn_b200.trimesh import TriMesh

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = ny = args.icon_n
    K = 65
    t0 = time.perf_counter()
    # weak scaling: every rank owns `ny` quad rows of a global nx x (ny*world) mesh plus one halo quad
    # row towards each neighbour; `vec` (the only field read through chains) is exchanged every step
    from dawn_b200 import runtime as rt, slabs
    sl = slabs.mesh_slab(nx, ny * world, world, rank)
    mesh = TriMesh(nx, sl.local_ny)
    if args.icon_reorder != "none":
        # locality-preserving renumbering (dawn_b200/reorder.py): the kernels are the same, only the
        # neighbour tables change; results map back bit for bit (tests/test_reorder.py)
        if world > 1:
            raise SystemExit("--icon-reorder is a 1-GPU option: the slab halo plan needs row-contiguous ids")
        from dawn_b200 import reorder as ro
        mesh = ro.reorder(mesh, args.icon_reorder)
    part = None
    if args.icon_partition != "none" and world > 1:
        # general partition (dawn_b200/partition.py) of the renumbered GLOBAL mesh instead of row slabs: cells
        # owned in contiguous id ranges, edges / vertices by the lowest-owner rule, one ring of halo cells,
        # splitter indices [Interior, Halo) = owned; `vec` halo copies refreshed point to point every step
        from dawn_b200 import partition as pt, reorder as ro
        gm = ro.reorder(TriMesh(nx, ny * world), args.icon_partition)
        part = pt.partition(gm, world)[rank]
        mesh = part.mesh
    nabla4 = args.workload == "icon_nabla4"
    diamond = args.workload == "icon_diamond"
    stencil_name = {"icon_nabla4": "ICON_nabla4_stencil", "icon_diamond": "ICON_laplacian_diamond_stencil"}.get(
        args.workload, "ICON_laplacian_stencil")
    passes = 2 if nabla4 else 1  # nabla4 = nabla2(nabla2(vec)): the nabla2 data flow twice
    if (nabla4 or diamond) and world > 1:
        # the slab plan below keeps one halo quad row and exchanges `vec` only; nabla4 would need a second
        # exchange (of nabla2_vec) between its passes, the diamond stencil one of u_vert / v_vert
        raise SystemExit("%s is a 1-GPU bench line; use icon_nabla2 for the multi-GPU runs" % args.workload)
    opts = {}
    for kv in args.opt:
        k_, v_ = kv.split("=")
        opts[k_] = int(v_)
    mod = dawn_b200.load_stencil(stencil_name, **opts)
    st = mod.on_mesh(mesh, K, splitters=None if part is None else part.splitters())
    t_mesh = time.perf_counter() - t0
    E, C, V = mesh.num_edges, mesh.num_cells, mesh.num_vertices
    n_of = {"cells": C, "edges": E, "vertices": V}
    gen = torch.Generator(device="cuda").manual_seed(7)
    tens = []
    for fd in mod.fields:
        n = n_of[fd["location"]]
        shape = (K, fd["sparse"], n) if fd["sparse"] else (K, n)
        t = torch.empty(shape, dtype=torch.float64, device="cuda")
        t.uniform_(0.5, 1.5, generator=gen)
        tens.append(t)
    torch.cuda.synchronize()
    plan = comm = None
    stream = torch.cuda.current_stream().cuda_stream
    if world > 1 and part is None:
        import ctypes
        idbuf = (ctypes.c_char * 128)()
        if rank == 0:
            rt.check(rt.lib().b200rt_nccl_unique_id(idbuf))
        obj = [bytes(idbuf)]
        dist.broadcast_object_list(obj, src=0)
        idbuf = (ctypes.c_char * 128).from_buffer_copy(obj[0])
        comm = ctypes.c_void_p()
        rt.check(rt.lib().b200rt_comm_init_rank(ctypes.byref(comm), world, idbuf, rank))
        per = sl.slots_per_row(1)
        plan = ctypes.c_void_p()
        rt.check(rt.lib().b200rt_halo_plan_create2(ctypes.byref(plan), comm, rank, world, per,
                                                   sl.local_ny, K, per, E, sl.halo_lo, sl.halo_hi, 1, 1))
    vec = None if diamond else tens[[fd["name"] for fd in mod.fields].index("vec")]

    def step():
        if plan is not None:
            rt.check(rt.lib().b200rt_halo_exchange(plan, vec.data_ptr(), stream))
        if part is not None:
            part.exchange(1, vec, dist)       # 1 = EDGE: gather at the owners, NCCL send / recv, scatter
        st.run(*tens)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler.mark_begin()
    l0 = st.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.all_reduce(torch.zeros(1, device="cuda"))   # device-side line-up of the ranks (see run_gpu)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = st.launches() - l0 + (0 if plan is None else args.steps * 2 * (sl.halo_lo + sl.halo_hi))

    # ---- parity self-check of the launch shape just timed: all K levels run on the GPU, a sample of
    # levels is compared with the oracle (the stencils have no vertical dependency, so a level can be
    # checked on its own); with slab neighbours the halo rows of `vec` are poisoned first
    parity = None
    if not nabla4:
        from util import unstructured_oracle_on_levels
        levels = [0, 8, K - 1]
        if plan is not None:
            per = sl.slots_per_row(1)
            if sl.halo_lo:
                vec[:, :per * sl.halo_lo] = float("nan")
            if sl.halo_hi:   # the slot rows of the upper halo quad row (one more row of slots follows them)
                vec[:, per * (sl.halo_lo + sl.nrows):per * (sl.halo_lo + sl.nrows + sl.halo_hi)] = float("nan")
        if part is not None:
            vec[:, part.n_owned[1]:] = float("nan")   # every halo copy of an edge
        step()
        torch.cuda.synchronize()
        written = set(mod.info.get("written", []))
        got = {fd["name"]: t.index_select(0, torch.as_tensor(levels, device="cuda")).cpu().numpy()
               for fd, t in zip(mod.fields, tens) if fd["name"] in written and not fd["sparse"]}
        want = unstructured_oracle_on_levels(stencil_name, mesh, mod.fields, tens, levels)
        loc_of = {"cells": 0, "edges": 1, "vertices": 2}   # dawn_b200.trimesh CELL, EDGE, VERTEX
        ndiff = nvals = 0
        for fd in mod.fields:
            if fd["name"] not in got:
                continue
            v = mesh.valid(loc_of[fd["location"]])
            if part is not None:   # owned elements: complete neighbourhoods, the values that count
                v = v & (np.arange(v.shape[0]) < part.n_owned[loc_of[fd["location"]]])
            g, w = got[fd["name"]][:, v], want[fd["name"]][:, v]
            ndiff += int((g.view(np.int64) != w.view(np.int64)).sum())
            nvals += g.size
        nan_in_vec = 0 if vec is None else int(torch.isnan(vec).sum().item())
        if part is not None:   # halo slots nobody owns (invalid edge slots) are never refreshed
            ev = torch.as_tensor(mesh.valid(1), device="cuda")
            nan_in_vec = int(torch.isnan(vec[:, ev]).sum().item())
        parity = {"checked": "%s on levels %s of %d, all valid elements%s" % (
                      sorted(got), levels, K,
                      "; halo rows of `vec` poisoned with NaN before the exchange" if plan is not None else
                      " this part owns; every halo copy of `vec` poisoned with NaN before the refresh"
                      if part is not None else ""),
                  "values_compared": nvals, "differing_values": ndiff + nan_in_vec,
                  "max_abs_diff": 0.0 if ndiff + nan_in_vec == 0 else float("nan"),
                  "oracle": "oracle/naive_interp.py (pinned against the reference's generated naive-ico text)"}
        flag = torch.tensor([float(ndiff + nan_in_vec != 0)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if float(flag.item()) != 0.0:
            sys.stderr.write("bench.py: rank %d parity check against the oracle: %s\n" % (rank, json.dumps(parity)))
            raise SystemExit(3)
    # work and bytes of the OWNED rows of one rank (halo rows are redundant work)
    own = TriMesh(nx, ny)
    E, C, V = own.num_edges, own.num_cells, own.num_vertices
    if part is not None:
        E, C, V = part.n_owned[1], part.n_owned[0], part.n_owned[2]
        cnt = torch.tensor([E, C, V], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt)                       # whole-job totals over the parts
        E, C, V = (int(x) // world for x in cnt.tolist())
    units = E * K * world
    bytes_run = passes * (K * (40 * E + 56 * V + 32 * C) + 4 * (6 * V + 3 * C + 2 * E + 2 * E))
    if diamond:
        # per edge and level: read 5 dense edge fields + 4 sparse ones (4 diamond slots each), write the
        # sparse vn_vert and 6 dense results = 31 doubles; per vertex: u_vert, v_vert once; E->C->V table once
        bytes_run = K * (248 * E + 16 * V) + 4 * 4 * E
    peak, peak_src = measured_peak()
    achieved = bytes_run / (ms * 1e-3) / 1e9
    line = {
        "metric": "edge-level updates/s", "value": units / (ms * 1e-3), "unit": "edge-level updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (uniform random fields generated on device; regular triangle mesh)",
        "config": {"workload": "%s on TriMesh %dx%d per GPU: %d edge slots, "
                               "%d cells, %d vertices x %d levels%s" % (
                                   "ICON_nabla4_stencil (nabla2 of nabla2; not in the reference, parity "
                                   "anchored on two passes of the nabla2 oracle)" if nabla4 else
                                   "ICON_laplacian_diamond_stencil (diamond nabla2 + Smagorinsky, "
                                   "dawn4py-tests/ICON_laplacian_diamond_stencil.py)" if diamond else
                                   "ICON_laplacian_stencil (nabla2)",
                                   nx, ny, E, C, V, K, "" if world == 1 else
                                   "; global mesh %dx%d, %s numbering, cut into %d parts of contiguous cell "
                                   "ranges (dawn_b200/partition.py: lowest-owner edges / vertices, one halo "
                                   "ring, splitter indices), `vec` halo copies refreshed point to point over "
                                   "NCCL every step; counts are per-part averages of OWNED elements" % (
                                       nx, ny * world, args.icon_partition, world) if part is not None else
                                   "; global mesh %dx%d in %d row slabs, 1 halo quad row, NCCL exchange "
                                   "of `vec` every step" % (nx, ny * world, world)),
                   "l2": "inputs larger than L2 (%.1f GB per step), no flush" % (bytes_run / 1e9),
                   "mesh_build_s": t_mesh, "numbering": args.icon_reorder, "codegen_options": opts},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None if (nabla4 or diamond) else icon_traffic(E),
                     "kernel": ("%d generated kernel launch/step (13 fused edge stages); %.0f algorithmic "
                                "bytes/step" % ((st.launches() - l0) // args.steps, bytes_run)) if diamond else
                               "%d generated kernel launches/step (vertex and cell stages co-scheduled in "
                               "one launch, 5 fused edge stages in the other%s); %.0f algorithmic "
                               "bytes/step" % ((st.launches() - l0) // args.steps,
                                               ", both twice" if nabla4 else "", bytes_run),
                     "peak_source": peak_src},
        "gpu_launches": launches, "clocks": clocks,
        "parity_checked": parity is not None, "max_abs_diff": None if parity is None else parity["max_abs_diff"],
        "parity": parity if parity is not None else
        "not checked in this run: nabla4 has a second pass whose inputs the first pass produces "
        "(tests/test_unstructured_gpu.py checks it at test sizes)",
    }
    if rank == 0 and world == 1 and not nabla4 and not diamond and not args.no_e2e:
        line["e2e"] = icon_e2e(mod, K, args.icon_e2e_n)
    if rank == 0 and world == 1 and not nabla4 and not diamond and not args.no_cpu_baseline:
        r = icon_reference(64, K, 3, 1)
        if r is not None:
            line["cpu_baseline"] = {"value": r[0], "unit": "edge-level updates/s", "cores": 1,
                                    "kind": "reference", "sample": r[2]}
    if rank == 0:
        print(json.dumps(line))
    if plan is not None:
        rt.lib().b200rt_halo_plan_destroy(plan)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def sub_workloads(args):
    """Runs `bench.py --workload tridiagonal` and `--workload icon_nabla2` as child processes (the main
    process keeps only ~3 GB of HBM; ICON at 20 M edges x 65 levels needs ~125 GB) and returns their JSON
    lines keyed by workload."""
    out = {}
    me = os.path.abspath(__file__)
    for key, extra, limit in (
            ("tridiagonal_solve_512x512x80", ["--workload", "tridiagonal", "--steps", str(max(args.steps, 50))], 300),
            ("icon_nabla2_20M_edges_x65", ["--workload", "icon_nabla2", "--steps", str(min(max(args.steps, 5), 20))], 600)):
        cmd = [sys.executable, me, "--gpus", "1", "--warmup", str(args.warmup)] + extra
        if args.no_cpu_baseline:
            cmd.append("--no-cpu-baseline")
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=limit)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode != 0 or not lines:
                out[key] = {"error": "exit code %d: %s" % (r.returncode, r.stderr.strip()[-300:])}
            else:
                out[key] = json.loads(lines[-1])
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": str(e)[:300]}
    return out


def icon_e2e(mod, K, nx):
    """The same metric end to end through the reference-facing entry point run_<N>_from_c_host (twin of the
    reference's, CudaIcoCodeGen.cpp:949-1028: element-major HOST arrays; device buffers allocated, inputs
    copied up and transposed to level-major on the GPU, kernels, results transposed back and copied down -
    every call).  Host buffers are pinned.  Measured on a smaller mesh than the device-resident line: the
    call is PCIe-bound, so its rate does not depend on the mesh size, and 20 M edges would need 110 GB of
    pinned host memory."""
    import ctypes
    import torch
    from dawn_b200.trimesh import TriMesh
    mesh = TriMesh(nx, nx)
    st = mod.on_mesh(mesh, K)
    n_of = {"cells": mesh.num_cells, "edges": mesh.num_edges, "vertices": mesh.num_vertices}
    host, h2d, d2h = [], 0, 0
    written = set(mod.info.get("written", []))
    for fd in mod.fields:
        n = n_of[fd["location"]] * max(1, fd["sparse"]) * K
        t = torch.empty(n, dtype=torch.float64).pin_memory()
        t.fill_(1.0)
        host.append(t)
        h2d += n * 8
        d2h += n * 8 if fd["name"] in written else 0
    fn = getattr(mod.lib, "run_%s_from_c_host" % mod.name)
    fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p]
    ptrs = (ctypes.c_void_p * len(host))(*[t.data_ptr() for t in host])
    stream = torch.cuda.current_stream().cuda_stream
    from dawn_b200 import runtime as rt
    rt.check(fn(st.handle, ptrs, len(host), stream))      # warm-up
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        rt.check(fn(st.handle, ptrs, len(host), stream))  # returns after the results are in the host buffers
    dt = (time.perf_counter() - t0) / reps
    return {"value": mesh.num_edges * K / dt, "unit": "edge-level updates/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": reps, "ms_per_step": dt * 1e3,
            "api": "run_%s_from_c_host (generated C ABI): pinned element-major host arrays of a TriMesh %dx%d "
                   "(%d edge slots) x %d levels; cudaMalloc + H2D + transpose + kernels + transpose + D2H "
                   "every call, as the reference's *_from_c_host does" % (mod.name, nx, nx, mesh.num_edges, K)}


def refcuda_child(workload):
    """Child process of the main bench line: times the reference's OWN pre-generated CUDA code, recompiled
    unmodified for sm_100a (oracle/ref_build/ref_cuda.cu -> oracle/_ref/libdawn_refcuda.so), on buffers of
    the bench shape.  hori_diff: dawn4py-tests/data/hori_diff_stencil_ref.cpp (4th-order diffusion with a
    coefficient field: in, out, coeff - the same 24 compulsory B/point as the headline stencil, whose
    CUDA text the reference does not ship); tridiagonal: data/tridiagonal_solve_stencil_ref.cpp (forward
    and backward sweep as two kernels).  Prints one JSON object."""
    import ctypes

    import torch

    import dawn_b200
    from oracle import capi
    lib = capi.refcuda()
    if lib is None:
        print(json.dumps({"unavailable": "oracle/_ref/libdawn_refcuda.so missing"}))
        return
    torch.cuda.set_device(0)
    I, J, K = SHAPE[workload]
    ni, nj = I + 2 * HALO, J + 2 * HALO
    n_fields = 3 if workload == "hori_diff" else 4
    stor = [dawn_b200.Storage(ni, nj, K + 1, "ijk") for _ in range(n_fields)]
    for n, st in enumerate(stor):
        st.buf.uniform_(0.5, 1.5)
        if workload == "tridiagonal" and n == 1:
            st.buf.add_(4.0)                                  # b: diagonally dominant
    torch.cuda.synchronize()
    ptrs = [ctypes.c_void_p(st.buf.data_ptr() + 8 * st.lead) for st in stor]
    fn = lib.refcuda_hori_diff if workload == "hori_diff" else lib.refcuda_tridiagonal_solve
    sec = ctypes.c_double(0.0)
    head = (ni, nj, K, HALO, stor[0].stride_j, stor[0].stride_k)
    rc = fn(*head, *ptrs, 3, ctypes.byref(sec))               # warm-up
    reps = 50
    rc = rc or fn(*head, *ptrs, reps, ctypes.byref(sec))
    if rc:
        print(json.dumps({"error": "reference CUDA kernel failed with code %d" % rc}))
        return
    ms = sec.value * 1e3 / reps
    peak, _ = measured_peak()
    gbs = BYTES_PER_POINT[workload] * I * J * K / (ms * 1e-3) / 1e9
    print(json.dumps({
        "ms_per_run": ms, "reps": reps, "achieved_gbs": gbs, "frac_of_peak": gbs / peak,
        "updates_per_s": I * J * K / (ms * 1e-3),
        "what": ("the reference's pre-generated CUDA code, unmodified, recompiled for sm_100a (-fmad=false): "
                 + ("dawn4py-tests/data/hori_diff_stencil_ref.cpp (in, out, coeff; 32x4 blocks + halo warps, "
                    "lap in a shared-memory IJ-cache)" if workload == "hori_diff" else
                    "dawn4py-tests/data/tridiagonal_solve_stencil_ref.cpp (two kernels, 32x1 blocks, "
                    "register K-caches)")
                 + "; %dx%dx%d, kernel time from its own timer_cuda events" % (I, J, K))}))


def cpp_class_e2e(I, J, K, steps):
    """Runs dawn_b200/lib/hori_diff_e2e (tests/cpp/hori_diff_e2e.cu, built by __graft_entry__.build()) and returns
    its JSON object, or a reason."""
    exe = os.path.join(ROOT, "dawn_b200", "lib", "hori_diff_e2e")
    if not os.path.exists(exe):
        return {"error": "dawn_b200/lib/hori_diff_e2e not built"}
    try:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        r = subprocess.run([exe, str(I), str(J), str(K), str(steps), "8", str(min(threads, K))],
                           capture_output=True, text=True, timeout=300)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not lines:
            return {"error": "exit code %d: %s" % (r.returncode, (r.stderr or r.stdout).strip()[-300:])}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:300]}


def same_stencil_as_reference_text():
    """The b200 module generated from stencils/hori_diff_stencil.iir - the dawn4py hori_diff (in, out, coeff)
    whose pre-generated CUDA text `reference_cuda_kernel` times - on the same shape, same storages layout,
    kernel-only timing with CUDA events."""
    import torch
    import dawn_b200
    I, J, K = SHAPE["hori_diff"]
    ni, nj = I + 2 * HALO, J + 2 * HALO
    mod = dawn_b200.load_stencil("hori_diff_stencil")
    dom = dawn_b200.Domain(ni, nj, K)
    dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
    st = mod(dom)
    stor = [dawn_b200.Storage(ni, nj, K + 1, fd["dims"]) for fd in mod.fields]
    for s_ in stor:
        s_.buf.uniform_(0.5, 1.5)
    for _ in range(5):
        st.run(*stor)
    torch.cuda.synchronize()
    reps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        st.run(*stor)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    peak, _ = measured_peak()
    gbs = 24 * I * J * K / (ms * 1e-3) / 1e9
    return {"ms_per_run": ms, "reps": reps, "achieved_gbs": gbs, "frac_of_peak": gbs / peak,
            "stencil": "hori_diff_stencil (dawn4py-tests/hori_diff_stencil.py: in, out, coeff), fields %s, 24 "
                       "algorithmic B/point, %dx%dx%d; bit-identical results: tests/test_reference_cuda_texts_gpu.py"
                       % ([fd["name"] for fd in mod.fields], I, J, K)}


def refcuda_baseline(workload):
    """Runs refcuda_child in its own process (a fault in code nobody has run on this GPU before must not
    reach the measured process) and returns its JSON object, or a reason."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--refcuda-child", workload],
                           capture_output=True, text=True, timeout=90)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": "child exited with %d: %s" % (r.returncode, r.stderr.strip()[-200:])}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def icon_reference(nx, K, reps, warm):
    """The reference's own generated naive-ico nabla2 (ICON_laplacian_stencil_ref.cpp, compiled unmodified
    into oracle/_ref) on the toylib mesh, single thread as the naive backend runs: (updates/s, ms per
    run, sample text) or None if oracle/_ref is not there."""
    from oracle import capi
    r = capi.icon_laplacian_reference_timed(nx, nx, K, warm + reps)
    if r is None:
        return None
    ne, t = r
    t = sorted(t[warm:])
    med = t[len(t) // 2]
    return ne * K / med, med * 1e3, (
        "ICON_laplacian_stencil_ref.cpp (reference's generated naive-ico code, unmodified) on a toylib "
        "%dx%d mesh: %d edges x %d levels per run, %d warm-up + %d timed runs, median %.2f s; "
        "g++ -O2 -ffp-contract=off" % (nx, nx, ne, K, warm, reps, med))


def run_icon_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, nx = 65, 64
    r = icon_reference(nx, K, max(1, args.steps), max(0, args.warmup))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference naive-ico build) missing"}))
        return
    val, ms, sample = r
    print(json.dumps({
        "impl": "reference", "metric": "edge-level updates/s", "value": val, "unit": "edge-level updates/s",
        "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ICON_laplacian_stencil (nabla2), reference naive-ico code on host cores"},
        "cpu_baseline": {"value": val, "unit": "edge-level updates/s", "cores": 1, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": val, "unit": "edge-level updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def icon_traffic(num_edges):
    """DRAM bytes per step (sum over the kernels of one step) from the committed ncu capture, only when
    it was taken on this mesh size."""
    tp = os.path.join(ROOT, "profiles", "traffic_icon_nabla2.json")
    if os.path.exists(tp):
        with open(tp) as f:
            d = json.load(f)
        if d.get("num_edges") == num_edges:
            return d.get("dram_bytes_per_step")
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hori_diff", choices=list(STENCIL) + ["icon_nabla2", "icon_nabla4", "icon_diamond"])
    ap.add_argument("--icon-n", type=int, default=None,
                    help="TriMesh nx=ny (default 2582 -> ~20 M edges; icon_diamond: 1500 -> 6.8 M edges, its "
                         "31 edge fields x 65 levels fill 110 GB)")
    ap.add_argument("--icon-reorder", default="none", choices=["none", "rcm", "morton"],
                    help="renumber the ICON mesh before the run (none = toylib's structured numbering)")
    ap.add_argument("--icon-partition", default="none", choices=["none", "rcm", "morton"],
                    help="multi-GPU ICON lines: general partition of the renumbered global mesh instead of row slabs")
    ap.add_argument("--opt", action="append", default=[], help="codegen option key=value (sweeps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="ICON lines: skip the host-buffer leg")
    ap.add_argument("--icon-e2e-n", type=int, default=1024,
                    help="TriMesh nx=ny of the ICON end-to-end leg (1024 -> 3.1 M edges, 16 GB of pinned host arrays)")
    ap.add_argument("--no-sub-workloads", action="store_true",
                    help="default line only: do not append the vertical-solve and ICON records")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "p2p"],
                    help="multi-GPU halo exchange: pack -> ncclSend/Recv -> unpack, or peer-memory stores")
    ap.add_argument("--refcuda-child", default=None, choices=list(STENCIL), help=argparse.SUPPRESS)
    ap.add_argument("--shape", default=None, help="IxJxK interior of the Cartesian workloads per GPU "
                    "(default: the BASELINE.json config, e.g. 1024x1024x80)")
    args = ap.parse_args()
    if args.shape and args.workload in SHAPE:
        SHAPE[args.workload] = tuple(int(v) for v in args.shape.lower().split("x"))
    if args.refcuda_child:
        refcuda_child(args.refcuda_child)
        return
    if args.icon_n is None:
        args.icon_n = 1500 if args.workload == "icon_diamond" else 2582
    if args.workload in ("icon_nabla2", "icon_nabla4", "icon_diamond"):
        if args.impl == "reference" and args.workload == "icon_nabla2":
            run_icon_reference(args)
        elif args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no compiled reference text for this stencil"}))
        else:
            run_icon(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
