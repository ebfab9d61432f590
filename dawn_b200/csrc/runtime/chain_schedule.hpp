// Host-side schedule of a chained unstructured launch (codegen option ico_chain; include/dawn_b200_rt.h,
// b200rt_chain_plan_*).  Shared by the CUDA runtime (rt.cu) and the CPU emulation of the tests
// (tests/cpu_emu/emu_rt.cpp), so that what the emulation checks is the schedule the GPU runs.
//
// A chained launch holds up to two PRODUCER roles (e.g. ICON nabla2's vertex and cell stages) and one
// CONSUMER role (its edge stages) that gathers what the producers wrote.  Work is cut into blocks of
// `block` elements; the blocks of one level block run in the order of `sched` (role << 30 | block index),
// which is the order of blockIdx.x.  Every consumer block comes after all producer blocks its elements
// reach through the neighbour tables, plus `lag` more producer blocks, so that by the time it is
// dispatched its inputs are complete (it still checks the producers' flags) and still in L2.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

namespace dawn_b200 {

struct ChainSchedule {
  unsigned nb[2] = {0, 0}; // producer blocks per role
  unsigned nc = 0;         // consumer blocks
  std::vector<unsigned> sched; // nb[0] + nb[1] + nc entries
  std::vector<int> need;       // per consumer block: lo0, hi0, lo1, hi1 - producer block ranges, hi exclusive
};

struct ChainTable {
  const int* soa; // host copy of a neighbour table [nbh][n_consumer], -1 = missing
  int nbh;
  int role;       // producer role the table leads to
};

inline ChainSchedule buildChainSchedule(int n_consumer, int block, int n_roles, const int* n_producer,
                                        const std::vector<ChainTable>& tables, int lag) {
  ChainSchedule s;
  for(int r = 0; r < n_roles && r < 2; ++r)
    s.nb[r] = (unsigned)((n_producer[r] + block - 1) / block);
  s.nc = (unsigned)((n_consumer + block - 1) / block);
  s.need.assign((size_t)s.nc * 4, 0);
  for(unsigned c = 0; c < s.nc; ++c) {
    const int e0 = (int)c * block, e1 = std::min(n_consumer, e0 + block);
    int lo[2] = {1 << 30, 1 << 30}, hi[2] = {-1, -1};
    for(auto const& t : tables)
      for(int n = 0; n < t.nbh; ++n) {
        const int* row = t.soa + (size_t)n * n_consumer;
        for(int e = e0; e < e1; ++e) {
          const int v = row[e];
          if(v >= 0)
            lo[t.role] = std::min(lo[t.role], v), hi[t.role] = std::max(hi[t.role], v);
        }
      }
    for(int r = 0; r < 2; ++r)
      if(hi[r] >= 0)
        s.need[c * 4 + 2 * r] = lo[r] / block, s.need[c * 4 + 2 * r + 1] = hi[r] / block + 1;
  }
  // merged order: producers in the proportional interleave of a co-scheduled launch, each consumer block
  // as soon as `lag` producer blocks beyond everything it (and every earlier consumer block) needs are out
  const unsigned tot = s.nb[0] + s.nb[1];
  unsigned p[2] = {0, 0}, runmax[2] = {0, 0};
  s.sched.reserve((size_t)tot + s.nc);
  auto issueProducer = [&]() {
    int r;
    if(p[0] >= s.nb[0])
      r = 1;
    else if(p[1] >= s.nb[1])
      r = 0;
    else // the role that is behind in relative progress
      r = ((unsigned long long)p[0] * s.nb[1] <= (unsigned long long)p[1] * s.nb[0]) ? 0 : 1;
    s.sched.push_back(((unsigned)r << 30) | p[r]);
    ++p[r];
  };
  for(unsigned c = 0; c < s.nc; ++c) {
    unsigned target[2];
    for(int r = 0; r < 2; ++r) {
      // + 1: a consumer may also wait for the producer block after its range (whole cache lines final)
      if(s.need[c * 4 + 2 * r + 1] > s.need[c * 4 + 2 * r])
        runmax[r] = std::max(runmax[r], std::min(s.nb[r], (unsigned)s.need[c * 4 + 2 * r + 1] + 1u));
      const unsigned share = tot ? (unsigned)((unsigned long long)lag * s.nb[r] / tot) : 0u;
      target[r] = std::min(s.nb[r], runmax[r] + (runmax[r] ? share : 0u));
    }
    while(p[0] < target[0] || p[1] < target[1])
      issueProducer();
    s.sched.push_back((2u << 30) | c);
  }
  while(p[0] < s.nb[0] || p[1] < s.nb[1])
    issueProducer();
  return s;
}

} // namespace dawn_b200
