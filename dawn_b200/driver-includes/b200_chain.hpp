// Device-side helpers of chained unstructured launches (codegen option ico_chain; host side:
// b200rt_chain_plan_* of include/dawn_b200_rt.h).  A producer block publishes one flag when all its
// threads have stored their results; a consumer block waits for the flags of the producer blocks its
// elements reach and then reads the handed-over fields with ld.global.cg (L2 only: L1 is not coherent
// and may hold a line from before the producer wrote it).
#pragma once
#if defined(__CUDACC__)
namespace dawn_b200 {
namespace chain {

// after __syncthreads(), by one thread: the block's stores are visible device-wide before the flag is
__device__ __forceinline__ void publish(unsigned* flag, unsigned epoch) {
  __threadfence();
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
// bounded spin (~2 s): blocks run in blockIdx order on every GPU so far, which makes the producers of a
// consumer block older than the block itself; if that ever failed, an error flag beats a hung device
__device__ __forceinline__ void wait(const unsigned* flag, unsigned epoch, unsigned* error) {
  unsigned v;
  for(unsigned spins = 0; spins < (1u << 24); ++spins) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if(v == epoch)
      return;
    __nanosleep(100);
  }
  *error = 1u;
}

// element `idx` of one level (`n` elements from `level`) of a handed-over field: an ordinary load that may stay in
// L1 - every line the consumer touches holds final values (its wait covers the neighbouring producer blocks) -
// except within a line's length of either end of the level, where the line is shared with a level that another
// block of levels produces (later, or earlier in a previous life of the line): those always come from L2
template <class T>
__device__ __forceinline__ T load(const T* level, int idx, int n) {
  return (idx >= 32 && idx < n - 32) ? level[idx] : __ldcg(level + idx);
}

} // namespace chain
} // namespace dawn_b200
#else
// CPU emulation of the tests (tests/cpu_emu): blocks run one after the other in blockIdx order, so a flag
// that is not there yet never will be
#include <atomic>
namespace dawn_b200 {
namespace chain {
inline void publish(unsigned* flag, unsigned epoch) {
  reinterpret_cast<std::atomic<unsigned>*>(flag)->store(epoch, std::memory_order_release);
}
inline void wait(const unsigned* flag, unsigned epoch, unsigned* error) {
  if(reinterpret_cast<const std::atomic<unsigned>*>(flag)->load(std::memory_order_acquire) != epoch)
    *error = 1u;
}
template <class T>
inline T load(const T* level, int idx, int) { return level[idx]; }
} // namespace chain
} // namespace dawn_b200
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline T __ldcs(const T* p) { return *p; }
template <class T>
inline void __stcs(T* p, T v) { *p = v; }
#endif
