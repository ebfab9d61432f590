// gridtools::dawn::domain / halo with the reference's API (driver-includes/domain.hpp:27-86,
// halo.hpp:24-30): sizes INCLUDE the halos; default horizontal halo = GRIDTOOLS_DAWN_HALO_EXTENT.
#pragma once
#include "defs.hpp"
#include <array>

namespace gridtools {
namespace dawn {

struct halo {
  static constexpr int value = GRIDTOOLS_DAWN_HALO_EXTENT;
};

class domain {
  struct shape_t {
    unsigned int size[3];
    unsigned int lo[3], hi[3]; // halo widths at the low / high end of each dimension
  } s_;

public:
  domain(unsigned int i, unsigned int j, unsigned int k)
      : s_{{i, j, k}, {(unsigned)halo::value, (unsigned)halo::value, 0u},
           {(unsigned)halo::value, (unsigned)halo::value, 0u}} {}
  explicit domain(const std::array<unsigned int, 3>& d) : domain(d[0], d[1], d[2]) {}

  void set_halos(unsigned int iminus = 0, unsigned int iplus = 0, unsigned int jminus = 0,
                 unsigned int jplus = 0, unsigned int kminus = 0, unsigned int kplus = 0) {
    s_.lo[0] = iminus, s_.hi[0] = iplus;
    s_.lo[1] = jminus, s_.hi[1] = jplus;
    s_.lo[2] = kminus, s_.hi[2] = kplus;
  }

  unsigned int isize() const { return s_.size[0]; }
  unsigned int jsize() const { return s_.size[1]; }
  unsigned int ksize() const { return s_.size[2]; }
  unsigned int iminus() const { return s_.lo[0]; }
  unsigned int iplus() const { return s_.hi[0]; }
  unsigned int jminus() const { return s_.lo[1]; }
  unsigned int jplus() const { return s_.hi[1]; }
  unsigned int kminus() const { return s_.lo[2]; }
  unsigned int kplus() const { return s_.hi[2]; }

  std::array<unsigned int, 3> dims() const { return {s_.size[0], s_.size[1], s_.size[2]}; }
  std::array<unsigned int, 6> halos() const {
    return {s_.lo[0], s_.hi[0], s_.lo[1], s_.hi[1], s_.lo[2], s_.hi[2]};
  }
  // number of compute points per dimension (not in the reference; used by drivers and benches)
  unsigned int interior(int d) const { return s_.size[d] - s_.lo[d] - s_.hi[d]; }
};

} // namespace dawn
} // namespace gridtools
