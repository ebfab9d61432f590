// Public C++ interface of the B200 emitter.  Mirrors the reference's backend entry points
//   std::unique_ptr<TranslationUnit> cuda::run(context, Options)   (CodeGen/Cuda/CudaCodeGen.h:34-37)
//   std::string codegen::run(map<name, serializedIIR>, Format, Backend, Options) (CodeGen/Driver.h:38-40)
// with serialized IIR (JSON) as the input so it can live behind a C ABI (include/dawn_b200_codegen.h).
#pragma once
#include "iir.hpp"

#include <map>
#include <string>
#include <vector>

namespace b200 {
namespace codegen {

// Backend enum twin of dawn::codegen::Backend (CodeGen/Options.h:22) extended by the two new values.
enum class Backend { GridTools, CXXNaive, CXXNaiveIco, CUDAIco, CUDA, CXXOpt, B200, B200Ico };

// Accepts the strings of dawn::codegen::parseBackendString (CodeGen/Driver.cpp:47-64) plus
// "b200"/"B200" and "b200-ico"/"B200Ico"; unknown -> std::invalid_argument("Backend not supported").
Backend parseBackendString(const std::string& backendStr);

// Options twin of dawn::codegen::Options (CodeGen/Options.inc:36-52) + Blackwell planning knobs that
// replace PassSetBlockSize / PassSetCaches decisions (Optimizer/PassSetBlockSize.cpp:24-71).
struct Options {
  int MaxHaloSize = 3;
  bool UseParallelEP = false;
  int MaxBlocksPerSM = 0;
  int nsms = 0;
  int DomainSizeI = 0, DomainSizeJ = 0, DomainSizeK = 0;
  bool RunWithSync = true;
  // --- b200 planner knobs (0 = let the planner decide) ---
  int BlockSizeI = 0;     // tile width in i (threads along i)
  int BlockSizeJ = 0;     // thread rows per block
  int RowsPerThread = 0;  // consecutive j rows computed by one thread (register micro-tile)
  int ColsPerThread = 0;  // consecutive i columns per thread (1 or 2; plain-tile kernels only)
  int IcoFullBlocks = 0;   // unstructured: full level blocks run an unrolled loop without per-level bound checks
  int RingStagger = 0;     // ns of start delay per phase for first-wave blocks of column-ring kernels (-1: env)
  int RingEarly = 1;       // column-ring kernels issue their first TMA prologue at kernel start
  int RingDrain = 1;       // column-ring kernels copy a slot to registers on arrival and refill it at once
  int CoSchedule = 1;      // unstructured: independent groups gathering a common field share a launch
  int SharedTemporaries = 0; // 1: TMA tile kernels evaluate widely reused temporaries once per level
                             // into a shared-memory IJ-cache instead of recomputing them per thread
                             // (measured slower on B200: hori_diff 0.44 ms vs 0.35 ms recomputed)
  int BlockSizeK = 0;     // k levels per block for Parallel multistages
  int InlineTemporaries = 1; // 1: recompute cheap temporaries in registers, 0: stage through memory
  int FuseColumns = 1;    // 1: fuse consecutive horizontally-pointwise multistages into one kernel
  int UnrollK = 0;        // k-loop unroll factor (0 = planner default)
  int LevelsPerThread = 0; // unstructured: levels per thread
  int BlockSizeIco = 0;    // unstructured: threads per block
  int UseTMA = 1;          // 1: tile kernels stage their read-only inputs with TMA (when aligned)
  int TmaStages = 0;       // depth of the TMA shared-memory ring over k (0 = planner default 4)
  int ColumnCache = 0;     // 1: fused column kernels keep sweep-to-sweep fields in shared memory
  int ColumnRing = 1;      // 1: sequential column kernels stream their inputs through a TMA ring
  int RingLevels = 0;      // k levels per ring slot (0 = default 4)
  int PrefetchK = 0;       // register prefetch depth of sequential column loops (0 = default 8)
  int PairStores = 1;      // TMA tile kernels: 16-byte stores of the two columns of a micro-tile row
  int IcoCacheHints = 0;   // unstructured: evict-first loads of read-once fields, streaming stores of final results
  int IcoChain = 0;        // 1: unstructured: fuse a launch with the next one that gathers its results (L2 hand-over;
                           // measured: less DRAM traffic but slower on B200, see dawn_b200_codegen.h)
  int IcoChainL1 = 0;      // chained consumers: L1-cached hand-over reads behind a wait widened by one block each side
  int IcoChainLag = 0;     // schedule distance between a consumer block and its producers (0 = 2048 blocks)
  int RingProducer = 0;    // 1: column-ring kernels get a dedicated TMA producer warp instead of thread 0 +
                           // __syncthreads (measured slower in full waves on B200, see dawn_b200_codegen.h)
  int SinglePrecision = 0; // 1: generate for ::dawn::float_type = float (element size 4)
  int RingPersistent = 0;  // 1: column-ring kernels as persistent blocks that prefetch their next column set
  int IcoHoist = 0;        // unstructured: read-only loads of a level hoisted to its top (all in flight together)
  int IcoStage = 0;        // unstructured: own-element rows staged in shared memory by cp.async.bulk (1: one mbarrier
                           // per level, 2: one per block); see dawn_b200_codegen.h
};

// TranslationUnit twin (CodeGen/TranslationUnit.h:26-53)
struct TranslationUnit {
  std::string filename;
  std::vector<std::string> ppDefines;
  std::map<std::string, std::string> stencils; // stencil name -> generated code
  std::string globals;
  // Concatenation ppDefines + globals + stencils, like codegen::generate (CodeGen/Driver.cpp:105-159)
  std::string str() const;
};

// context: name -> serialized StencilInstantiation (JSON).
TranslationUnit run(const std::map<std::string, std::string>& context, Backend backend,
                    const Options& options);

// Per-backend entry points (twins of cuda::run / cudaico::run).
TranslationUnit runCartesian(const std::map<std::string, iir::Instantiation>& ctx, const Options&);
TranslationUnit runIco(const std::map<std::string, iir::Instantiation>& ctx, const Options&);
// interface files of the b200-ico C ABI (twins of --output-c-header / --output-f90-interface)
std::string icoCHeader(const iir::Instantiation& inst);
std::string icoF90Interface(const iir::Instantiation& inst);

} // namespace codegen
} // namespace b200
