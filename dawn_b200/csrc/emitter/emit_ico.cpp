// placeholder until the unstructured emitter lands (replaced below in this round)
#include "codegen.hpp"
#include <stdexcept>
namespace b200 { namespace codegen {
int icoChainSize(const std::vector<iir::LocationType>&, bool) { return -1; }
TranslationUnit runIco(const std::map<std::string, iir::Instantiation>&, const Options&) {
  throw std::runtime_error("b200-ico backend not built");
}
}}
