/* dawn_b200_codegen.h — C ABI of the B200 code-generation backend (libdawn_b200_codegen.so).
 *
 * This is the string-level seam of the reference's backend selection:
 *   std::string dawn::codegen::run(const std::map<std::string, std::string>& serializedIIR,
 *                                  IIRSerializer::Format, Backend, const Options&)
 *                                                   dawn/src/dawn/CodeGen/Driver.h:38-40
 *   Backend dawn::codegen::parseBackendString(const std::string&)   CodeGen/Driver.cpp:47-64
 * i.e. serialized optimized IIR in, one translation unit of source text out.  A Dawn tree binds it
 * with the ~40-line adapter shown in INTEGRATION.md (Backend::B200 / Backend::B200Ico cases of the
 * switch in CodeGen/Driver.cpp:67-85).  Plain pointers and sizes only; no C++ types cross it.
 */
#ifndef DAWN_B200_CODEGEN_H
#define DAWN_B200_CODEGEN_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Values of dawn::codegen::Backend (CodeGen/Options.h:22) followed by the two new backends. */
enum dawn_b200_backend {
  DAWN_BACKEND_GRIDTOOLS = 0,
  DAWN_BACKEND_CXXNAIVE = 1,
  DAWN_BACKEND_CXXNAIVEICO = 2,
  DAWN_BACKEND_CUDAICO = 3,
  DAWN_BACKEND_CUDA = 4,
  DAWN_BACKEND_CXXOPT = 5,
  DAWN_BACKEND_B200 = 6,
  DAWN_BACKEND_B200ICO = 7
};

/* dawn::codegen::Options (CodeGen/Options.inc:36-52) + Blackwell planner knobs (0 = automatic).
 * The planner knobs take over what --block-size-i/j/k (Optimizer/Options.inc:44-46) and the cache
 * passes decide for the reference's CUDA backend. */
typedef struct dawn_b200_options {
  int max_halo_points;   /* --max-halo-points, default 3 */
  int use_parallel_ep;   /* accepted, unused */
  int max_blocks_per_sm; /* accepted, unused: occupancy comes from __launch_bounds__ */
  int nsms;              /* accepted, unused */
  int domain_size_i, domain_size_j, domain_size_k; /* accepted, unused (sizes are run-time) */
  int run_with_sync;     /* --run-with-sync, default 1 */
  int block_size_i;      /* threads along i per block */
  int block_size_j;      /* thread rows per block */
  int rows_per_thread;   /* j rows computed per thread (register micro-tile) */
  int block_size_k;      /* k levels per block for Parallel multistages */
  int inline_temporaries; /* default 1: recompute single-assignment temporaries in registers */
  int fuse_columns;      /* default 1: fuse horizontally pointwise multistages (solver sweeps) */
  int unroll_k;          /* k-loop unroll factor */
  int levels_per_thread; /* unstructured: k levels per thread */
  int block_size_ico;    /* unstructured: threads per block */
  int use_tma;           /* default 1: tile kernels stage read-only inputs through smem with TMA */
  int tma_stages;        /* depth of the TMA ring over k (0 = default 4) */
  int column_cache;      /* default 0: fused solver sweeps hand fields over through shared memory */
  int column_ring;       /* default 1: sequential column kernels stream inputs through a TMA ring */
  int ring_levels;       /* k levels per ring slot (0 = default 8) */
  int prefetch_k;        /* register prefetch depth of sequential column loops (0 = default 8) */
  int cols_per_thread;   /* i columns per thread of tile kernels without redundant halo stages (1|2) */
  int shared_temporaries; /* default 0; 1: TMA tile kernels keep widely reused temporaries in a shared-memory
                             IJ-cache (the reference's cache(IJ, local), CudaCodeGen ij_caches) */
  int co_schedule;        /* default 1: unstructured groups that are independent and gather a common field are
                             launched as one kernel whose blocks alternate between them (L2 reuse) */
  int ring_drain;         /* default 1: column-ring kernels drain an arrived slot into registers and refill
                             it at once (all slots in flight during the sweep); 2: TMA tile kernels with one
                             live stage also refill right after their hoisted loads */
  int ring_early;         /* default 1: the first TMA prologue of a column-ring kernel is issued at kernel
                             start, ahead of the short k ranges that precede the streamed range */
  int ring_stagger;       /* default 0; N > 0: the first-wave blocks of a column-ring kernel delay their start by
                             (block % resident blocks per SM) * N ns so that their forward (HBM-fed) and
                             backward (L2-fed) sweeps interleave; -1: N from $DAWN_B200_STAGGER_NS (probing) */
  int ico_full_blocks;    /* default 0: unstructured kernels run full blocks of levels through an unrolled
                             loop without per-level bound checks (more loads in flight per thread) */
  int pair_stores;        /* default 1: TMA tile kernels with two columns per thread store the two results of a row as
                             one 16-byte store when both cells are valid (fields the kernel never reads) */
  int ico_cache_hints;    /* unstructured kernels: fields that every stage of a launch reads at its own element only
                             (each value is needed once) are loaded evict-first (ld.global.cs) and results no later
                             launch reads are stored streaming (st.global.cs), so that the GATHERED fields keep
                             their lines in L2 for the neighbours that come back to them */
  int ico_chain;          /* default 0; 1: an unstructured launch whose results the NEXT launch gathers through a chain
                             (ICON nabla2: rot_vec / div_vec of the vertex + cell launch, read by the edge launch)
                             is fused with it into ONE launch: blocks follow a host-built schedule in which every
                             consumer block comes `ico_chain_lag` blocks after the last producer block it needs,
                             producers publish a per-block flag (st.release.gpu), consumers wait for the flags
                             of their neighbourhood (ld.acquire.gpu, bounded spin) and read the handed-over
                             fields from L2 (ld.global.cg) instead of HBM.  Measured on B200 (ICON nabla2, 20 M
                             edges x 65 levels, lag 2048): DRAM traffic 124.9 -> 115.2 GB per run (the re-read of
                             rot_vec / div_vec is gone, 1.10 x the algorithmic bytes), but the three roles mixed
                             in one launch reach only 61 % DRAM throughput against ~80 % of the two separate
                             launches: 23.2 ms against 20.7 ms.  Bit-exact, tested, off by default */
  int ico_chain_l1;       /* default 0: chained consumers read the handed-over fields with ld.global.cg; 1: ordinary
                             (L1-cached) loads behind a wait widened by one producer block on either side, so
                             that every cache line they touch is final (the ends of a level, shared with other
                             blocks of levels, still go through L2).  Measured slower: 28.3 against 24.6 ms */
  int ico_chain_lag;      /* producer blocks between a consumer block and the last producer it needs (0 = 2048:
                             below the ~1600 producer blocks of one residency generation - 148 SMs x 16 blocks -
                             consumers are dispatched before their producers finish and spin in their slots:
                             37.8 ms at 512) */
  int ring_producer;      /* default 0; 1: column-ring kernels get a dedicated producer warp that issues the TMA loads
                             (full / empty mbarriers per slot, no block-wide barrier in the sweeps) instead of
                             thread 0 between __syncthreads().  Measured on B200 (tridiagonal_solve): the lone-block
                             latency drops from 18.0 to 16.0 us, but full waves get slower (1024^2 x 80: 87.4 %
                             against 90.7 % of the HBM peak), so it stays off */
  int single_precision;   /* default 0: ::dawn::float_type is double; 1: the translation unit is generated for
                             float (the reference selects the type when the generated code is COMPILED,
                             driver-includes/defs.hpp:47-81; here the TMA box geometry and the shared-memory
                             pairs depend on the element size, so the emitter has to know - the emitted text
                             defines DAWN_PRECISION itself and static_asserts the size) */
  int ico_stage;          /* default 0; 1 | 2: unstructured kernels of Parallel multistages fetch the rows they read at
                             their OWN elements (dense fields and sparse weights of the kernel's location, never
                             written by the launch) with 1-D bulk copies (cp.async.bulk) into shared memory - one
                             copy of block_size elements per level and row, issued by the first lanes of the block
                             at kernel start, 1: one mbarrier per level, 2: one for the block - so that only the
                             gathers are left to the load/store path and the streams' bytes in flight do not cost
                             registers.  Taken per block at run time (full blocks, 16-byte aligned field bases);
                             other blocks run the plain code of the same kernel.  Measured on B200 (ICON nabla2,
                             20 M edges x 65 levels, bit-identical): 28.8 ms (8 levels per block), 23.7 ms (4 levels)
                             against 21.2 ms plain - the rows of a block cost 25-50 KB of shared memory, which
                             caps the resident warps (32 or 16 per SM instead of 64) and shrinks L1 for the
                             gathers; off */
  int ico_hoist;          /* default 0; 1: unstructured kernels load every read-only value of a level (own element and
                             top-level neighbours, no vertical offset) at the top of the level, so that the loads
                             of a level are in flight together - left at their uses they are split by the basic
                             blocks a double-precision division ends.  Measured on B200: 21.3 ms against 21.2 ms,
                             no gain at 64 resident warps (the 32-register budget spills what the hoisting holds);
                             off */
  int persistent_ring;    /* default 0; 1: column-ring kernels run as persistent blocks (one per resident slot of the device)
                             that loop over the column sets; while a block finishes the last streamed range of a set
                             it already requests the first ring slots of its NEXT set into the slots that range
                             frees, so the bulk-copy latency at the start of a set (~1/4 of a lone block's life in
                             the Thomas solver) is hidden behind the previous set's backward sweep.  Measured on B200
                             (tridiagonal_solve, 444 persistent blocks, bit-identical): 512^2 x 80 0.198 ms / 77.7 %
                             against 0.183 ms / 84.2 %, 1024^2 x 80 0.758 ms / 81 % against 0.689 ms / 89 % - the
                             hardware's own dispatch of independent blocks balances better than a static
                             round-robin over sets, and the lead of two short backward chunks does not cover the
                             latency it was meant to hide; off */
} dawn_b200_options;

/* Fills `opts` with the defaults above. */
void dawn_b200_default_options(dawn_b200_options* opts);

/* parseBackendString twin: returns the enum value, or -1 ("Backend not supported"). */
int dawn_b200_parse_backend(const char* backend_str);

/* Generates code for `n` serialized stencil instantiations (IIR as proto3 JSON text).
 *   names[i], iir_json[i], iir_len[i]: instantiation name and its serialized bytes.
 *   backend: DAWN_BACKEND_B200 or DAWN_BACKEND_B200ICO (others -> error: not provided here).
 * On success returns 0 and stores a malloc'ed, NUL-terminated translation unit in *out_code
 * (release with dawn_b200_free).  On failure returns non-zero; dawn_b200_codegen_last_error()
 * describes the problem (the reference throws std::runtime_error / SemanticError instead). */
int dawn_b200_codegen_run(const char* const* names, const char* const* iir_json,
                          const size_t* iir_len, int n, int backend, const dawn_b200_options* opts,
                          char** out_code, size_t* out_len);

/* Interface files for one unstructured instantiation (reference: --output-c-header and
 * --output-f90-interface of the cuda-ico backend, CudaIcoCodeGen.cpp:1869-1892): kind 0 = C header,
 * 1 = Fortran 2003 interface module.  Result is malloc'ed like out_code above. */
int dawn_b200_codegen_interface(const char* iir_json, size_t iir_len, int kind, char** out_text,
                                size_t* out_len);

/* Serialized-IIR formats (reference: IIRSerializer::Format { Json, Byte },
 * dawn/src/dawn/Serialization/IIRSerializer.h).  Every entry point above accepts either format and
 * tells them apart by the first byte; pass iir_len for Byte input (it may contain NULs). */
#define DAWN_B200_IIR_JSON 0
#define DAWN_B200_IIR_BYTE 1

/* Converts one serialized StencilInstantiation to `to_format`; result malloc'ed (dawn_b200_free). */
int dawn_b200_iir_convert(const char* data, size_t len, int to_format, char** out, size_t* out_len);

/* The wire schema the Byte reader implements, one "Message number name type" line per field and
 * one "enum Name Value number" line per enumerator (checked against the .proto files in tests). */
int dawn_b200_iir_schema(char** out_text, size_t* out_len);

void dawn_b200_free(void* p);
const char* dawn_b200_codegen_last_error(void);

/* ICOChainSize twin (dawn/src/dawn/CodeGen/IcoChainSizes.cpp:199-264): number of neighbours
 * reached by a chain of location types (0 cells, 1 edges, 2 vertices) on a regular triangular
 * mesh.  Returns -1 for an invalid chain. */
int dawn_b200_ico_chain_size(const int* chain, int chain_len, int include_center);

#ifdef __cplusplus
}
#endif
#endif
