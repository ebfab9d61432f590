// TEST INFRASTRUCTURE: host-memory stand-in for the entry points of include/dawn_b200_rt.h that generated
// b200-ico modules call ("device" memory is plain host memory in the emulation, tests/cpu_emu/emu_cuda.hpp).
#include "../../include/dawn_b200_rt.h"
#include "../../dawn_b200/csrc/runtime/chain_schedule.hpp" // the schedule the GPU runtime builds

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;
extern "C" {
const char* b200rt_last_error(void) { return g_err.c_str(); }
int b200rt_set_error(int code, const char* msg) {
  g_err = msg ? msg : "";
  return code;
}
// every "device" allocation sits between two guard bands; a kernel that writes past either end of its
// scratch is counted (emu_guard_violations) when the buffer is freed
static const size_t kGuard = 4096;
static long g_violations = 0;
long emu_guard_violations(void) { return g_violations; }
int b200rt_malloc(void** p, size_t bytes) {
  unsigned char* raw = static_cast<unsigned char*>(std::malloc(bytes + 2 * kGuard + sizeof(size_t)));
  if(!raw)
    return b200rt_set_error(B200_ERR_CUDA, "emu: out of memory");
  std::memcpy(raw, &bytes, sizeof(size_t));
  std::memset(raw + sizeof(size_t), 0xA5, kGuard);
  std::memset(raw + sizeof(size_t) + kGuard + bytes, 0xA5, kGuard);
  *p = raw + sizeof(size_t) + kGuard;
  return 0;
}
int b200rt_free(void* p) {
  if(!p)
    return 0;
  unsigned char* user = static_cast<unsigned char*>(p);
  unsigned char* raw = user - kGuard - sizeof(size_t);
  size_t bytes;
  std::memcpy(&bytes, raw, sizeof(size_t));
  for(size_t n = 0; n < kGuard; ++n)
    if(raw[sizeof(size_t) + n] != 0xA5 || user[bytes + n] != 0xA5) {
      ++g_violations;
      break;
    }
  std::free(raw);
  return 0;
}
int b200rt_malloc_host(void** p, size_t bytes) { return b200rt_malloc(p, bytes); }
int b200rt_free_host(void* p) { return b200rt_free(p); }
int b200rt_memcpy_d2d(void* d, const void* s, size_t n, void*) {
  std::memmove(d, s, n);
  return 0;
}
int b200rt_memset(void* p, int v, size_t bytes, void*) {
  std::memset(p, v, bytes);
  return 0;
}
int b200rt_memcpy_h2d(void* d, const void* s, size_t n, void*) {
  std::memcpy(d, s, n);
  return 0;
}
int b200rt_memcpy_d2h(void* d, const void* s, size_t n, void*) {
  std::memcpy(d, s, n);
  return 0;
}
int b200rt_stream_synchronize(void*) { return 0; }
int b200rt_check_last_launch(void) { return 0; }
int b200rt_event_create(void** e) {
  *e = std::malloc(1);
  return 0;
}
int b200rt_event_destroy(void* e) {
  std::free(e);
  return 0;
}
int b200rt_event_record(void*, void*) { return 0; }
int b200rt_stream_wait_event(void*, void*) { return 0; }
int b200rt_event_synchronize(void*) { return 0; }
int b200rt_event_elapsed_ms(void*, void*, float* ms) {
  *ms = 0.f;
  return 0;
}
// CUtensorMap stand-in: the geometry itself (struct emu_tensor_map of emu_cuda.hpp)
struct emu_tensor_map_rt {
  const void* base;
  int size[3];
  long long stride[3];
  int box[3];
  int elem_bytes;
};
static long g_tensor_maps = 0;
long emu_tensor_maps_encoded(void) { return g_tensor_maps; }
#ifdef EMU_BLOCK_THREADS
long emu_bulk_copies_issued(void) { return dawn_b200::tma::emu_bulk_copies.load(); } // ico_stage: staged blocks ran
#else
long emu_bulk_copies_issued(void) { return 0; }
#endif
int b200rt_tensor_map_3d_es(void* map, const void* base, int size_i, int size_j, int size_k, int stride_j,
                            int stride_k, int box_i, int box_j, int box_k, int elem_bytes) {
  // the alignment rules of the real encoder (rt.cu): what the emitted launchers promise
  const int per16 = 16 / elem_bytes;
  if((reinterpret_cast<uintptr_t>(base) & 15) || stride_j % per16 || stride_k % per16 || (box_i * elem_bytes) % 16)
    return b200rt_set_error(B200_ERR_LAYOUT, "emu: TMA alignment rules violated");
  ++g_tensor_maps;
  emu_tensor_map_rt m{base, {size_i, size_j, size_k}, {1, stride_j, stride_k}, {box_i, box_j, box_k}, elem_bytes};
  std::memset(map, 0, 128);
  std::memcpy(map, &m, sizeof(m));
  return 0;
}
int b200rt_tensor_map_3d(void* map, const void* base, int size_i, int size_j, int size_k, int stride_j,
                         int stride_k, int box_i, int box_j, int box_k) {
  return b200rt_tensor_map_3d_es(map, base, size_i, size_j, size_k, stride_j, stride_k, box_i, box_j, box_k, 8);
}
// the *_from_host / verify paths are GPU-tested (tests/test_unstructured_gpu.py); not part of the emulation
int b200rt_reshape(const double*, double*, int, int, int, void*) {
  return b200rt_set_error(B200_ERR_RUNTIME, "emu: b200rt_reshape is not emulated");
}
int b200rt_reshape_back(const double*, double*, int, int, int, void*) {
  return b200rt_set_error(B200_ERR_RUNTIME, "emu: b200rt_reshape_back is not emulated");
}
int b200rt_verify_field(const double*, const double*, size_t, double, double, int, b200_verification_metrics*,
                        void*) {
  return b200rt_set_error(B200_ERR_RUNTIME, "emu: b200rt_verify_field is not emulated");
}
int b200rt_reshape_f32(const float*, float*, int, int, int, void*) {
  return b200rt_set_error(B200_ERR_RUNTIME, "emu: b200rt_reshape_f32 is not emulated");
}
int b200rt_reshape_back_f32(const float*, float*, int, int, int, void*) {
  return b200rt_set_error(B200_ERR_RUNTIME, "emu: b200rt_reshape_back_f32 is not emulated");
}
int b200rt_verify_field_f32(const float*, const float*, size_t, double, double, int, b200_verification_metrics*,
                            void*) {
  return b200rt_set_error(B200_ERR_RUNTIME, "emu: b200rt_verify_field_f32 is not emulated");
}
int b200rt_metrics_record_add(b200_metrics_record*, const char*, const char*, const b200_verification_metrics*) {
  return 0;
}
// chained unstructured launches: the same schedule as the GPU runtime, arrays in host memory
struct b200_chain_plan {
  b200_chain_view v{};
  unsigned epoch = 0;
  std::vector<unsigned> sched, flags;
  std::vector<int> need;
  unsigned error = 0;
};
int b200rt_chain_plan_create(b200_chain_plan** plan, int n_consumer, int block, int n_roles, const int* n_producer,
                             const int* const* tables, const int* table_role, const int* table_nbh, int n_tables,
                             int level_blocks, int lag_blocks, void*) {
  std::vector<dawn_b200::ChainTable> tabs;
  for(int t = 0; t < n_tables; ++t)
    tabs.push_back({tables[t], table_nbh[t], table_role[t]});
  dawn_b200::ChainSchedule sc =
      dawn_b200::buildChainSchedule(n_consumer, block, n_roles, n_producer, tabs, lag_blocks > 0 ? lag_blocks : 2048);
  auto* p = new b200_chain_plan();
  p->sched = sc.sched, p->need = sc.need;
  p->flags.assign((size_t)level_blocks * (sc.nb[0] + sc.nb[1]) + 1, 0u);
  p->v.sched = p->sched.data(), p->v.need = p->need.data(), p->v.flags = p->flags.data(), p->v.error = &p->error;
  p->v.sched_len = (unsigned)sc.sched.size(), p->v.nb0 = sc.nb[0], p->v.nb1 = sc.nb[1], p->v.nc = sc.nc;
  p->v.level_blocks = (unsigned)level_blocks;
  *plan = p;
  return 0;
}
int b200rt_chain_plan_view(const b200_chain_plan* plan, b200_chain_view* view) {
  *view = plan->v;
  return 0;
}
unsigned b200rt_chain_plan_next_epoch(b200_chain_plan* plan) { return ++plan->epoch; }
int b200rt_chain_plan_status(b200_chain_plan* plan, void*) {
  if(plan && plan->error)
    return b200rt_set_error(B200_ERR_RUNTIME, "emu: a consumer block ran before one of its producers");
  return 0;
}
int b200rt_chain_plan_destroy(b200_chain_plan* plan) {
  delete plan;
  return 0;
}
}
