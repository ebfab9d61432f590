// TEST INFRASTRUCTURE ONLY.  Runs the reference's pre-generated cxxnaive code
// dawn/test/unit-test/dawn/CodeGen/reference/nonoverlapping_stencil.cpp (included where it lies).
#include <cassert>
#include REF_FILE(nonoverlapping_stencil.cpp)
#include "ref_common.hpp"
extern "C" int ref_nonoverlapping_stencil(int ni, int nj, int nk, int nk_alloc, int halo,
                                          const double* in, double* out) {
  auto dom = refw::make_domain(ni, nj, nk, halo);
  refw::meta_data_t md(ni, nj, nk_alloc);
  auto s_in = refw::wrap_in(md, in, "in"), s_out = refw::wrap_in(md, out, "out");
  dawn_generated::cxxnaive::generated st(dom);
  st.run(s_in, s_out);
  refw::copy_out(s_out, out);
  return 0;
}
