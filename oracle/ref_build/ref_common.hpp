// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Helpers shared by the ref_*.cpp wrappers that
// expose the reference's own, unmodified, pre-generated naive C++ stencils and its verifier
// through a plain C ABI (dense i-fastest [k][j][i] double buffers).
#pragma once
#include <cstring>
#include <driver-includes/gridtools_includes.hpp>
#include <driver-includes/verify.hpp>

namespace refw {
using gridtools::dawn::domain;
using gridtools::dawn::meta_data_t;
using gridtools::dawn::storage_t;

inline domain make_domain(int ni, int nj, int nk, int halo) {
  domain dom(ni, nj, nk);
  dom.set_halos(halo, halo, halo, halo, 0, 0);
  return dom;
}
// buffers are (ni, nj, nk_alloc) dense, i fastest: same as the shim's storage layout.
inline storage_t wrap_in(const meta_data_t& md, const double* src, const char* name) {
  storage_t s(md, name);
  std::memcpy(s.raw(), src, sizeof(double) * md.padded_total_length());
  return s;
}
inline void copy_out(const storage_t& s, double* dst) {
  std::memcpy(dst, s.raw(), sizeof(double) * s.get_storage_info_ptr()->padded_total_length());
}
} // namespace refw
