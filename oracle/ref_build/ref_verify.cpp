// TEST INFRASTRUCTURE ONLY.  C ABI over the reference's verifier
// (dawn/src/driver-includes/verify.hpp:48-64 fillMath, :87-157 verify) compiled unmodified.
#define DAWN_GENERATED 1
#define GRIDTOOLS_DAWN_HALO_EXTENT 3
#include "ref_common.hpp"

extern "C" {
// mask bit0=i, bit1=j, bit2=k: which dimensions the storage has (ijk=7, j-only=2, ij=3 ...).
int ref_fillmath(double a, double b, double c, double d, double e, double f, int ni, int nj, int nk,
                 int mask, double* out) {
  using namespace gridtools::dawn;
  domain dom(ni, nj, nk);
  verifier verif(dom);
  if(mask == 7) {
    meta_data_ijk_t md(ni, nj, nk);
    storage_ijk_t s(md, "f");
    verif.fillMath(a, b, c, d, e, f, s);
    std::memcpy(out, s.raw(), sizeof(double) * md.padded_total_length());
  } else if(mask == 3) {
    meta_data_ij_t md(ni, nj, 1);
    storage_ij_t s(md, "f");
    verif.fillMath(a, b, c, d, e, f, s);
    std::memcpy(out, s.raw(), sizeof(double) * md.padded_total_length());
  } else if(mask == 2) {
    meta_data_j_t md(1, nj, 1);
    storage_j_t s(md, "f");
    verif.fillMath(a, b, c, d, e, f, s);
    std::memcpy(out, s.raw(), sizeof(double) * md.padded_total_length());
  } else if(mask == 1) {
    meta_data_i_t md(ni, 1, 1);
    storage_i_t s(md, "f");
    verif.fillMath(a, b, c, d, e, f, s);
    std::memcpy(out, s.raw(), sizeof(double) * md.padded_total_length());
  } else if(mask == 4) {
    meta_data_k_t md(1, 1, nk);
    storage_k_t s(md, "f");
    verif.fillMath(a, b, c, d, e, f, s);
    std::memcpy(out, s.raw(), sizeof(double) * md.padded_total_length());
  } else {
    return -1;
  }
  return 0;
}

// returns 1 if verifier::verify(a, b) passes on the interior of domain (ni,nj,nk) with halos.
int ref_verify(int ni, int nj, int nk, int nk_alloc, int halo, double precision, const double* a,
               const double* b) {
  using namespace gridtools::dawn;
  domain dom = refw::make_domain(ni, nj, nk, halo);
  verifier verif(dom, precision);
  meta_data_t md(ni, nj, nk_alloc);
  storage_t sa = refw::wrap_in(md, a, "a"), sb = refw::wrap_in(md, b, "b");
  return verif.verify(sa, sb, /*max_erros=*/0) ? 1 : 0; // 0: keep the reference from printing
}
}
