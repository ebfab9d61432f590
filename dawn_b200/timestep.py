"""A model time step as ONE CUDA graph: a fixed sequence of generated-stencil runs and slab halo
exchanges, recorded once and replayed with a single launch.

The reference has no such layer (every generated run() synchronises the device before and after,
CudaCodeGen.cpp:435-450, and SURVEY.md §8(f) ranks "a CUDA-graph time-step wrapper chaining several
generated stencils + halo exchanges" as the next component).  What is recorded are the launches the
C ABI makes on the capturing stream: generated kernels, the pack/unpack kernels of
b200rt_halo_exchange and its ncclSend/ncclRecv group (NCCL supports stream capture).  Pointers and
domain sizes are baked into the graph, so the fields must stay allocated; their contents may change
between replays."""
import ctypes

from . import runtime as _rt


class TimeStep:
    def __init__(self):
        self._ops = []
        self._graph = None
        self.launches_per_step = 0

    # ---- recording the sequence ------------------------------------------------------------------
    def run(self, stencil, *fields, rows=None):
        """Append `stencil.run(*fields, rows=rows)` (Cartesian or unstructured instance)."""
        if rows is None:
            self._ops.append(lambda s: stencil.run(*fields, stream=s))
        else:
            self._ops.append(lambda s: stencil.run(*fields, stream=s, rows=rows))
        self._stencils = getattr(self, "_stencils", []) + [stencil]
        return self

    def exchange(self, plan, field_ptr):
        """Append a b200rt_halo_exchange of the field at device address `field_ptr` with `plan`
        (a b200_halo_plan* as ctypes.c_void_p)."""
        fn = _rt.lib().b200rt_halo_exchange
        self._ops.append(lambda s: _rt.check(fn(plan, ctypes.c_void_p(int(field_ptr)),
                                                ctypes.c_void_p(s))))
        return self

    def call(self, fn):
        """Append an arbitrary `fn(stream_ptr)` that only enqueues work on that stream."""
        self._ops.append(fn)
        return self

    # ---- execution -------------------------------------------------------------------------------
    def _enqueue(self, stream_ptr):
        for op in self._ops:
            op(stream_ptr)

    def eager(self):
        """Runs the sequence launch by launch on the current stream."""
        import torch
        self._enqueue(torch.cuda.current_stream().cuda_stream)

    def capture(self, warmup=1, capture_error_mode="thread_local"):
        """Records the sequence.  `warmup` eager passes come first: the first call of a generated
        module sets kernel attributes and allocates scratch temporaries, which must not happen while
        the stream is capturing.  capture_error_mode "thread_local": helper threads of the process
        (NCCL proxies, torch.distributed's watchdog) keep making CUDA calls during the capture."""
        import torch
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._enqueue(side.cuda_stream)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = sum(s.launches() for s in getattr(self, "_stencils", []))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=capture_error_mode):
            self._enqueue(torch.cuda.current_stream().cuda_stream)
        self.launches_per_step = sum(s.launches() for s in getattr(self, "_stencils", [])) - before
        self._graph = g
        return self

    def replay(self):
        """One time step: a single graph launch on the current stream."""
        if self._graph is None:
            raise RuntimeError("TimeStep.replay() before capture()")
        self._graph.replay()

    def close(self):
        """Drops the graph.  A graph that holds NCCL nodes must go before its communicator is
        destroyed (ncclCommDestroy waits for it otherwise)."""
        self._graph = None

    def step(self):
        """replay() if captured, else eager()."""
        if self._graph is None:
            self.eager()
        else:
            self._graph.replay()
