// Device-side helpers for the TMA shared-memory ring of generated tile kernels (sm_90+ PTX through
// libcu++'s cuda::ptx wrappers; SASS: UTMALDG / SYNCS).  One elected thread arms an mbarrier with the
// byte count of a pipeline stage and issues one cp.async.bulk.tensor per staged field; every thread
// of the block then waits on the barrier's phase parity before reading the stage.
#pragma once
#if defined(__CUDACC__)
#include <cuda.h>
#include <cuda/ptx>
#include <cstdint>

namespace dawn_b200 {
namespace tma {

__device__ __forceinline__ void barrier_init(unsigned long long* bar, unsigned count) {
  cuda::ptx::mbarrier_init(reinterpret_cast<uint64_t*>(bar), count);
}
// make the initialised barriers visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
  cuda::ptx::fence_mbarrier_init(cuda::ptx::sem_release, cuda::ptx::scope_cluster);
  cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
}
// orders this thread's earlier generic-proxy writes before later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async() { cuda::ptx::fence_proxy_async(); }
__device__ __forceinline__ void expect_bytes(unsigned long long* bar, unsigned bytes) {
  cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta,
                                       cuda::ptx::space_shared, reinterpret_cast<uint64_t*>(bar),
                                       bytes);
}
// plain arrival (release at CTA scope): consumers hand a drained ring slot back to the producer warp
__device__ __forceinline__ void arrive(unsigned long long* bar) {
  (void)cuda::ptx::mbarrier_arrive(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared,
                                   reinterpret_cast<uint64_t*>(bar));
}
__device__ __forceinline__ void load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                        int c2, unsigned long long* bar) {
  const int32_t coords[3] = {c0, c1, c2};
  cuda::ptx::cp_async_bulk_tensor(cuda::ptx::space_cluster, cuda::ptx::space_global, smem_dst, map,
                                  coords, reinterpret_cast<uint64_t*>(bar));
}
// 1-D bulk copy global -> shared (cp.async.bulk, SASS UBLKCP): `bytes` a multiple of 16, both addresses 16-byte
// aligned; completes its bytes on `bar`.  The unstructured kernels stage the rows they read at their own
// elements this way (ico_stage): the bytes in flight live in shared memory instead of registers.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes,
                                             unsigned long long* bar) {
  cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, smem_dst, gmem_src, bytes,
                           reinterpret_cast<uint64_t*>(bar));
}
__device__ __forceinline__ void wait(unsigned long long* bar, unsigned parity) {
  while(!cuda::ptx::mbarrier_try_wait_parity(reinterpret_cast<uint64_t*>(bar), parity)) {
  }
}

} // namespace tma
} // namespace dawn_b200
#endif
