// Access-extent constants exported by generated stencils (`static constexpr cartesian_extent
// <field>_extent`), API of the reference's driver-includes/extent.hpp:25-33.
#pragma once
#include <array>

namespace dawn {
namespace driver {

// {{iminus, iplus}, {jminus, jplus}, {kminus, kplus}}
using cartesian_extent = std::array<std::array<int, 2>, 3>;

struct unstructured_extent {
  bool horizontal;             // reads neighbours through a chain
  std::array<int, 2> vertical; // {kminus, kplus}
};

template <int D>
constexpr std::array<int, 2> get(const cartesian_extent& e) {
  return e[D];
}
template <int D>
constexpr auto get(const unstructured_extent& e) {
  if constexpr(D == 0)
    return e.horizontal;
  else
    return e.vertical;
}

} // namespace driver
} // namespace dawn
