// dawn-b200-codegen: command-line twin of `dawn-codegen` (dawn/src/dawn/dawn-codegen.cpp:71-143)
// for the two backends this project provides.  Reads optimized IIR (proto3 JSON or protobuf bytes,
// as written by `dawn-opt` / IIRSerializer in either format) from a file or stdin, writes the CUDA
// translation unit to -o or stdout.
//   dawn-opt x.sir | dawn-b200-codegen --backend=b200 -o x_b200.cu
#include "codegen.hpp"
#include "iir.hpp"
#include "proto_wire.hpp"

#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>

int main(int argc, char** argv) {
  using namespace b200::codegen;
  std::string backend = "b200", input, output, cHeader, f90Interface, convertTo;
  Options opts;
  auto intOpt = [&](const std::string& a, const char* name, int& dst) {
    std::string p = std::string("--") + name + "=";
    if(a.compare(0, p.size(), p) == 0) {
      dst = std::stoi(a.substr(p.size()));
      return true;
    }
    return false;
  };
  for(int n = 1; n < argc; ++n) {
    std::string a = argv[n];
    if(a == "-h" || a == "--help") {
      std::cout << "Usage: dawn-b200-codegen [-b|--backend b200|b200-ico] [-o out.cu] [in.iir]\n"
                   "  in.iir: serialized StencilInstantiation, proto3 JSON or protobuf bytes (auto-detected)\n"
                   "  --output-c-header=FILE --output-f90-interface=FILE  (b200-ico) interface files\n"
                   "  --convert=json|byte  re-serialize the IIR instead of generating code\n"
                   "  planner knobs: --block-size-i= --block-size-j= --rows-per-thread= "
                   "--block-size-k= --inline-temporaries= --fuse-columns= --unroll-k= "
                   "--levels-per-thread= --block-size-ico= --max-halo-points= --run-with-sync=\n";
      return 0;
    } else if((a == "-b" || a == "--backend") && n + 1 < argc)
      backend = argv[++n];
    else if(a.compare(0, 10, "--backend=") == 0)
      backend = a.substr(10);
    else if(a == "-o" && n + 1 < argc)
      output = argv[++n];
    // reference: --output-c-header / --output-f90-interface of the cuda-ico backend
    // (CudaIcoCodeGen.cpp:1869-1892, dawn-codegen.cpp)
    else if(a.compare(0, 18, "--output-c-header=") == 0)
      cHeader = a.substr(18);
    else if(a.compare(0, 23, "--output-f90-interface=") == 0)
      f90Interface = a.substr(23);
    else if(a.compare(0, 10, "--convert=") == 0) // json | byte: re-serialize the IIR instead of generating code
      convertTo = a.substr(10);
    else if(intOpt(a, "block-size-i", opts.BlockSizeI) || intOpt(a, "block-size-j", opts.BlockSizeJ) ||
            intOpt(a, "rows-per-thread", opts.RowsPerThread) ||
            intOpt(a, "block-size-k", opts.BlockSizeK) ||
            intOpt(a, "inline-temporaries", opts.InlineTemporaries) ||
            intOpt(a, "fuse-columns", opts.FuseColumns) || intOpt(a, "unroll-k", opts.UnrollK) ||
            intOpt(a, "levels-per-thread", opts.LevelsPerThread) ||
            intOpt(a, "block-size-ico", opts.BlockSizeIco) || intOpt(a, "use-tma", opts.UseTMA) ||
            intOpt(a, "tma-stages", opts.TmaStages) ||
            intOpt(a, "column-cache", opts.ColumnCache) || intOpt(a, "prefetch-k", opts.PrefetchK) ||
            intOpt(a, "column-ring", opts.ColumnRing) || intOpt(a, "ring-levels", opts.RingLevels) ||
            intOpt(a, "cols-per-thread", opts.ColsPerThread) ||
            intOpt(a, "shared-temporaries", opts.SharedTemporaries) ||
            intOpt(a, "co-schedule", opts.CoSchedule) || intOpt(a, "ring-drain", opts.RingDrain) ||
            intOpt(a, "ring-early", opts.RingEarly) || intOpt(a, "ring-stagger", opts.RingStagger) ||
            intOpt(a, "ico-full-blocks", opts.IcoFullBlocks) ||
            intOpt(a, "single-precision", opts.SinglePrecision) ||
            intOpt(a, "ring-producer", opts.RingProducer) ||
            intOpt(a, "pair-stores", opts.PairStores) || intOpt(a, "ico-cache-hints", opts.IcoCacheHints) || intOpt(a, "ico-stage", opts.IcoStage) || intOpt(a, "ico-hoist", opts.IcoHoist) || intOpt(a, "persistent-ring", opts.RingPersistent) || intOpt(a, "ico-chain", opts.IcoChain) || intOpt(a, "ico-chain-l1", opts.IcoChainL1) || intOpt(a, "ico-chain-lag", opts.IcoChainLag) ||
            intOpt(a, "max-blocks-sm", opts.MaxBlocksPerSM) ||
            intOpt(a, "max-halo-points", opts.MaxHaloSize)) {
    } else if(a.compare(0, 16, "--run-with-sync=") == 0)
      opts.RunWithSync = a.substr(16) != "0" && a.substr(16) != "false";
    else if(!a.empty() && a[0] != '-')
      input = a;
    else {
      std::cerr << "dawn-b200-codegen: unknown option " << a << "\n";
      return 2;
    }
  }
  try {
    std::string text;
    if(input.empty())
      text.assign(std::istreambuf_iterator<char>(std::cin), std::istreambuf_iterator<char>());
    else {
      std::ifstream f(input, std::ios::binary);
      if(!f)
        throw std::runtime_error("cannot open " + input);
      text.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    if(!convertTo.empty()) {
      b200::json::Value dom = b200::proto::looksLikeJson(text)
                                  ? b200::json::parse(text)
                                  : b200::proto::decode(text, "StencilInstantiation");
      if(b200::proto::looksLikeJson(text))
        b200::proto::normalizeJson(dom, "StencilInstantiation");
      std::string res;
      if(convertTo == "json")
        res = b200::json::write(dom);
      else if(convertTo == "byte")
        res = b200::proto::encode(dom, "StencilInstantiation");
      else
        throw std::runtime_error("--convert takes json or byte");
      if(output.empty())
        std::cout << res;
      else {
        std::ofstream o(output, std::ios::binary);
        o << res;
      }
      return 0;
    }
    if(!cHeader.empty() || !f90Interface.empty()) {
      auto inst = b200::iir::parseInstantiation(text);
      if(!inst.unstructured)
        throw std::runtime_error("interface files are generated for unstructured instantiations");
      if(!cHeader.empty()) {
        std::ofstream o(cHeader);
        o << icoCHeader(inst);
      }
      if(!f90Interface.empty()) {
        std::ofstream o(f90Interface);
        o << icoF90Interface(inst);
      }
    }
    auto tu = run({{"restoredIIR", text}}, parseBackendString(backend), opts);
    std::string code = tu.str();
    if(output.empty())
      std::cout << code;
    else {
      std::ofstream o(output);
      o << code;
    }
  } catch(std::exception& e) {
    std::cerr << "dawn-b200-codegen: error: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
