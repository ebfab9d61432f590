// TEST INFRASTRUCTURE ONLY.  Runs the reference's pre-generated cxxnaive code
// dawn/test/unit-test/dawn/CodeGen/reference/update_dz_c.cpp (included where it lies).
#include <cassert>
#include REF_FILE(update_dz_c.cpp)
#include "ref_common.hpp"
// fields in API order: dp_ref, zs, area, ut, vt, gz, gz_x, gz_y, ws3 (update_dz_c.cpp:471-473)
extern "C" int ref_update_dz_c(int ni, int nj, int nk, int nk_alloc, int halo, double dt,
                               double** f) {
  auto dom = refw::make_domain(ni, nj, nk, halo);
  refw::meta_data_t md(ni, nj, nk_alloc);
  refw::storage_t s[9];
  static const char* names[9] = {"dp_ref", "zs", "area", "ut", "vt", "gz", "gz_x", "gz_y", "ws3"};
  for(int n = 0; n < 9; ++n)
    s[n] = refw::wrap_in(md, f[n], names[n]);
  dawn_generated::cxxnaive::update_dz_c st(dom);
  st.set_dt(dt);
  st.run(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
  for(int n = 0; n < 9; ++n)
    refw::copy_out(s[n], f[n]);
  return 0;
}
