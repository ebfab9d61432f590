"""dawn_b200 — a B200-native (sm_100a) stencil execution backend behind Dawn's codegen boundary.

  codegen   : optimized IIR (proto3 JSON) -> CUDA translation unit     (C++ emitter behind a C ABI)
  runtime   : thin C-ABI runtime the generated host code calls
  stencil   : Python mirror of the generated-stencil interface (Domain / Storage / Stencil.run)
  timestep  : a sequence of stencil runs + halo exchanges recorded and replayed as one CUDA graph
  build     : in-tree builds of all native pieces for sm_100a
There is no CPU fallback anywhere in this package: without the CUDA modules it raises.
"""
from .codegen import CodeGenBackend, CodeGenError, compile_iir, parse_backend_string  # noqa: F401
from .stencil import Domain, Stencil, StencilModule, Storage, load_stencil  # noqa: F401
from .timestep import TimeStep  # noqa: F401
