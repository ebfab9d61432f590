// Run-time headers of generated b200-ico code (counterpart of the reference's
// driver-includes/{unstructured_interface,unstructured_domain,cuda_utils}.hpp for the cuda-ico
// backend): the mesh arrives as the plain b200_mesh_t of include/dawn_b200_rt.h.
#pragma once
#include <cassert>
#include <cstdint>
#include <exception>
#include <string>
#include <vector>

#include "defs.hpp"
#include "extent.hpp"
#include "math.hpp"
#include "b200_runtime.hpp"
#include "timer_b200.hpp"
#include "b200_chain.hpp"
#include "b200_tma.hpp"

namespace dawn_b200 {

// device pointer of the neighbour table for (chain, include_center) with exactly `nbh_size`
// neighbours per element, or nullptr
inline const int* find_table(const b200_mesh_t& mesh, const int* chain, int chain_len,
                             int include_center, int nbh_size) {
  for(int t = 0; t < mesh.num_tables; ++t) {
    const b200_nbh_table_t& tb = mesh.tables[t];
    if(tb.chain_len != chain_len || (tb.include_center != 0) != (include_center != 0) ||
       tb.nbh_size != nbh_size)
      continue;
    bool same = true;
    for(int n = 0; n < chain_len; ++n)
      same = same && tb.chain[n] == chain[n];
    if(same)
      return tb.table;
  }
  return nullptr;
}

} // namespace dawn_b200
