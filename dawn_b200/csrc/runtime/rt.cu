// libdawn_b200rt.so — implementation of include/dawn_b200_rt.h (sm_100a).
#include "../../../include/dawn_b200_rt.h"
#include "chain_schedule.hpp"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <limits>
#include <cstring>
#include <map>
#include <string>
#include <vector>

struct b200_graph {
  cudaGraphExec_t exec;
};

// stencil -> per-iteration {field -> metrics} (std::map: keys sorted like nlohmann::json objects)
struct b200_metrics_record {
  std::map<std::string, std::vector<std::map<std::string, b200_verification_metrics>>> stencils;
  std::string text;
};

namespace {
thread_local std::string g_err;

int cudaFail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return B200_ERR_CUDA;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if(e_ != cudaSuccess)                                                                          \
      return cudaFail(e_, #call);                                                                  \
  } while(0)

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
} // namespace

extern "C" {

const char* b200rt_last_error(void) { return g_err.c_str(); }
int b200rt_set_error(int code, const char* msg) {
  g_err = msg ? msg : "";
  return code;
}

int b200rt_device_count(int* n) {
  CK(cudaGetDeviceCount(n));
  return 0;
}
int b200rt_set_device(int device) {
  CK(cudaSetDevice(device));
  return 0;
}
int b200rt_device_synchronize(void) {
  CK(cudaDeviceSynchronize());
  return 0;
}
int b200rt_malloc(void** p, size_t bytes) {
  CK(cudaMalloc(p, bytes ? bytes : 1));
  return 0;
}
int b200rt_free(void* p) {
  if(p)
    CK(cudaFree(p));
  return 0;
}
int b200rt_malloc_host(void** p, size_t bytes) {
  CK(cudaMallocHost(p, bytes ? bytes : 1));
  return 0;
}
int b200rt_free_host(void* p) {
  if(p)
    CK(cudaFreeHost(p));
  return 0;
}
int b200rt_memset(void* p, int v, size_t bytes, void* s) {
  CK(cudaMemsetAsync(p, v, bytes, S(s)));
  return 0;
}
int b200rt_memcpy_h2d(void* d, const void* h, size_t bytes, void* s) {
  CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, S(s)));
  return 0;
}
int b200rt_memcpy_d2h(void* h, const void* d, size_t bytes, void* s) {
  CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, S(s)));
  return 0;
}
int b200rt_memcpy_d2d(void* d, const void* s0, size_t bytes, void* s) {
  CK(cudaMemcpyAsync(d, s0, bytes, cudaMemcpyDeviceToDevice, S(s)));
  return 0;
}
int b200rt_memcpy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                    size_t height, void* s) {
  CK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, S(s)));
  return 0;
}
int b200rt_stream_create(void** s) {
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = st;
  return 0;
}
int b200rt_stream_destroy(void* s) {
  CK(cudaStreamDestroy(S(s)));
  return 0;
}
int b200rt_stream_synchronize(void* s) {
  CK(cudaStreamSynchronize(S(s)));
  return 0;
}
int b200rt_stream_wait_event(void* s, void* e) {
  CK(cudaStreamWaitEvent(S(s), static_cast<cudaEvent_t>(e), 0));
  return 0;
}
int b200rt_event_create(void** e) {
  cudaEvent_t ev;
  CK(cudaEventCreate(&ev));
  *e = ev;
  return 0;
}
int b200rt_event_destroy(void* e) {
  CK(cudaEventDestroy(static_cast<cudaEvent_t>(e)));
  return 0;
}
int b200rt_event_record(void* e, void* s) {
  CK(cudaEventRecord(static_cast<cudaEvent_t>(e), S(s)));
  return 0;
}
int b200rt_event_synchronize(void* e) {
  CK(cudaEventSynchronize(static_cast<cudaEvent_t>(e)));
  return 0;
}
int b200rt_event_elapsed_ms(void* a, void* b, float* ms) {
  CK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b)));
  return 0;
}
int b200rt_check_last_launch(void) {
  CK(cudaGetLastError());
  return 0;
}

int b200rt_tensor_map_3d(void* map, const void* base, int size_i, int size_j, int size_k,
                         int stride_j, int stride_k, int box_i, int box_j, int box_k) {
  return b200rt_tensor_map_3d_es(map, base, size_i, size_j, size_k, stride_j, stride_k, box_i, box_j,
                                 box_k, 8);
}

int b200rt_tensor_map_3d_es(void* map, const void* base, int size_i, int size_j, int size_k,
                            int stride_j, int stride_k, int box_i, int box_j, int box_k,
                            int elem_bytes) {
  if(elem_bytes != 8 && elem_bytes != 4)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_tensor_map_3d_es: elem_bytes must be 4 or 8");
  const int per16 = 16 / elem_bytes; // elements per 16-byte unit
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if(!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if(qres != cudaDriverEntryPointSuccess || !fn)
      return b200rt_set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = reinterpret_cast<encode_fn>(fn);
  }
  if((reinterpret_cast<uintptr_t>(base) & 15) || (stride_j % per16) || (stride_k % per16) || box_i > 256 ||
     box_j > 256 || box_k > 256 || box_k < 1 || ((box_i * elem_bytes) & 15))
    return b200rt_set_error(B200_ERR_LAYOUT, "b200rt_tensor_map_3d: TMA alignment rules violated");
  cuuint64_t dims[3] = {(cuuint64_t)size_i, (cuuint64_t)size_j, (cuuint64_t)size_k};
  cuuint64_t strides[2] = {(cuuint64_t)stride_j * elem_bytes, (cuuint64_t)stride_k * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_i, (cuuint32_t)box_j, (cuuint32_t)box_k};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(static_cast<CUtensorMap*>(map),
                      elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                      const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if(r != CUDA_SUCCESS) {
    char buf[96];
    std::snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return b200rt_set_error(B200_ERR_CUDA, buf);
  }
  return 0;
}
} // extern "C"

// ------------------------------------------------------------------------------------------------
// Fused verification: one pass, block-level reduction, one atomic per block and metric.
// ------------------------------------------------------------------------------------------------
namespace {

struct VerifyAcc {
  double max_rel, min_rel, max_abs, min_abs;
  unsigned long long n_failed;
};

__device__ inline double atomicMaxD(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if(__longlong_as_double(assumed) >= v)
      break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while(assumed != old);
  return __longlong_as_double(old);
}
__device__ inline double atomicMinD(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if(__longlong_as_double(assumed) <= v)
      break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while(assumed != old);
  return __longlong_as_double(old);
}

template <class T>
__global__ void __launch_bounds__(256)
verify_kernel(const T* __restrict__ dsl, const T* __restrict__ ref, size_t n,
              double rel_tol, double abs_tol, VerifyAcc* acc) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double mxr = 0, mnr = inf, mxa = 0, mna = inf;
  unsigned long long bad = 0;
  for(size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n;
      idx += (size_t)gridDim.x * blockDim.x) {
    const T a = __ldg(&dsl[idx]), b = __ldg(&ref[idx]);
    // cuda_verify.hpp:25-45 (expected = reference value b, actual = a): both errors are 0 for equal
    // values, else |b - a| and |(b - a) / b| - so a zero reference with a nonzero value gives inf
    // (FieldType arithmetic where the reference's templates use it: the absolute error and isclose's
    // difference are formed in T, the relative error in double)
    const double ae = (b != a) ? (double)fabs(b - a) : 0.0;
    const double re = (b != a) ? fabs(((double)b - a) / b) : 0.0;
    // isclose_kernel (cuda_verify.hpp:58-69): false whenever a NaN is involved
    const bool close = fabs(a - b) <= (abs_tol + rel_tol * fabs(b));
    // thrust's maximum / minimum leave the result open when NaN errors are reduced; here they are
    // left out of the four extrema (fmax / fmin drop them) and counted as failures
    mxr = fmax(mxr, re), mnr = fmin(mnr, re);
    mxa = fmax(mxa, ae), mna = fmin(mna, ae);
    bad += close ? 0ull : 1ull;
  }
  for(int o = 16; o > 0; o >>= 1) {
    mxr = fmax(mxr, __shfl_xor_sync(0xffffffffu, mxr, o));
    mnr = fmin(mnr, __shfl_xor_sync(0xffffffffu, mnr, o));
    mxa = fmax(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
    mna = fmin(mna, __shfl_xor_sync(0xffffffffu, mna, o));
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  __shared__ double s[4][8];
  __shared__ unsigned long long sb[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if(l == 0)
    s[0][w] = mxr, s[1][w] = mnr, s[2][w] = mxa, s[3][w] = mna, sb[w] = bad;
  __syncthreads();
  if(threadIdx.x == 0) {
    for(int q = 1; q < 8; ++q) {
      mxr = fmax(mxr, s[0][q]), mnr = fmin(mnr, s[1][q]);
      mxa = fmax(mxa, s[2][q]), mna = fmin(mna, s[3][q]);
      bad += sb[q];
    }
    atomicMaxD(&acc->max_rel, mxr);
    atomicMinD(&acc->min_rel, mnr);
    atomicMaxD(&acc->max_abs, mxa);
    atomicMinD(&acc->min_abs, mna);
    if(bad)
      atomicAdd(&acc->n_failed, bad);
  }
}

// [elem][sparse][k] (k fastest)  ->  [k][sparse][elem] (elem fastest); 32x32 smem transpose per
// sparse slot so that both the read and the write are coalesced.
template <class T>
__global__ void __launch_bounds__(256)
reshape_kernel(const T* __restrict__ in, T* __restrict__ out, int k_size, int n_elem,
               int sparse, bool back) {
  __shared__ T tile[32][33];
  const int s = blockIdx.z;
  const int e0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
  if(!back) {
    for(int r = ty; r < 32; r += 8) {
      int e = e0 + r, k = k0 + tx;
      if(e < n_elem && k < k_size)
        tile[r][tx] = in[((size_t)e * sparse + s) * k_size + k];
    }
    __syncthreads();
    for(int r = ty; r < 32; r += 8) {
      int k = k0 + r, e = e0 + tx;
      if(e < n_elem && k < k_size)
        out[((size_t)k * sparse + s) * n_elem + e] = tile[tx][r];
    }
  } else {
    for(int r = ty; r < 32; r += 8) {
      int k = k0 + r, e = e0 + tx;
      if(e < n_elem && k < k_size)
        tile[r][tx] = in[((size_t)k * sparse + s) * n_elem + e];
    }
    __syncthreads();
    for(int r = ty; r < 32; r += 8) {
      int e = e0 + r, k = k0 + tx;
      if(e < n_elem && k < k_size)
        out[((size_t)e * sparse + s) * k_size + k] = tile[tx][r];
    }
  }
}

__global__ void __launch_bounds__(256)
nbh_transpose_kernel(const int* __restrict__ rm, int* __restrict__ soa, int n_elem, int nbh) {
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if(idx >= (size_t)n_elem * nbh)
    return;
  int e = (int)(idx % n_elem), n = (int)(idx / n_elem);
  soa[idx] = __ldg(&rm[(size_t)e * nbh + n]);
}

// pack / unpack of j-rows for the slab halo exchange: buf[k][row][i]
__global__ void __launch_bounds__(256)
halo_pack_kernel(const double* __restrict__ field, double* __restrict__ buf, int ni, int nk, int sj,
                 int sk, int j_first, int rows, bool unpack, double* __restrict__ field_w) {
  const size_t total = (size_t)ni * rows * nk;
  for(size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
      idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % ni);
    const int r = (int)((idx / ni) % rows);
    const int k = (int)(idx / ((size_t)ni * rows));
    const size_t f = (size_t)i + (size_t)(j_first + r) * sj + (size_t)k * sk;
    if(unpack)
      field_w[f] = buf[idx];
    else
      buf[idx] = __ldg(&field[f]);
  }
}

// ---- peer-memory halo exchange: flags at the head of every plan's IPC-shared arena ----------------
// (one 32-byte sector each; epochs count exchanges from 1)
enum : int {
  HF_ARRIVED_LO = 0,    // written by rank-1: its push of this epoch has landed in my recv_lo
  HF_ARRIVED_HI = 4,    // written by rank+1: ... in my recv_hi
  HF_CREDIT_LO = 8,     // written by rank-1: it has consumed its recv_hi up to epoch-1, push `epoch` may go
  HF_CREDIT_HI = 12,    // written by rank+1: same for its recv_lo
  HF_PUSH_EPOCH = 16,   // local: epoch of the next push
  HF_UNPACK_EPOCH = 20, // local: epoch of the next unpack
  HF_PUSH_DONE = 24,    // local: blocks of the running push kernel that have finished
  HF_UNPACK_DONE = 28,
  HF_ERROR = 32,        // local: a wait ran into its time limit (the neighbour never showed up)
  HF_WORDS = 64
};
constexpr long long kSpinLimitCycles = 20ll << 30; // ~11 s at 2 GHz: never hang the device on a lost peer

__device__ __forceinline__ void spin_until(volatile unsigned long long* flag, unsigned long long want,
                                           unsigned long long* err) {
  const long long t0 = clock64();
  while(*flag < want) {
    __nanosleep(40);
    if(clock64() - t0 > kSpinLimitCycles) {
      atomicExch(err, 1ull);
      return;
    }
  }
}

// Pack + transfer in one kernel: the boundary rows of my slab are stored straight into the landing
// buffers of the slab neighbours (peer memory over NVLink), then the last block raises `arrived` there.
__global__ void __launch_bounds__(256)
halo_push_kernel(const double* __restrict__ field, unsigned long long* my, double* peer_lo_buf,
                 unsigned long long* peer_lo_flags, double* peer_hi_buf,
                 unsigned long long* peer_hi_flags, int ni, int nk, int sj, int sk, int j_lo_first,
                 int rows_lo, int j_hi_first, int rows_hi) {
  volatile unsigned long long* vmy = my;
  const unsigned long long epoch = vmy[HF_PUSH_EPOCH];
  if(threadIdx.x == 0) {
    // the neighbour must be done reading what the previous push left in its landing buffer
    if(peer_lo_buf)
      spin_until(&vmy[HF_CREDIT_LO], epoch, &my[HF_ERROR]);
    if(peer_hi_buf)
      spin_until(&vmy[HF_CREDIT_HI], epoch, &my[HF_ERROR]);
  }
  __syncthreads();
  const size_t total_lo = peer_lo_buf ? (size_t)ni * rows_lo * nk : 0;
  const size_t total_hi = peer_hi_buf ? (size_t)ni * rows_hi * nk : 0;
  for(size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total_lo + total_hi;
      idx += (size_t)gridDim.x * blockDim.x) {
    const bool lo = idx < total_lo;
    const size_t e = lo ? idx : idx - total_lo;
    const int rows = lo ? rows_lo : rows_hi;
    const int i = (int)(e % ni);
    const int r = (int)((e / ni) % rows);
    const int k = (int)(e / ((size_t)ni * rows));
    const size_t f = (size_t)i + (size_t)((lo ? j_lo_first : j_hi_first) + r) * sj + (size_t)k * sk;
    (lo ? peer_lo_buf : peer_hi_buf)[e] = __ldg(&field[f]);
  }
  __threadfence_system(); // my peer stores are visible system-wide before this block counts as done
  __syncthreads();
  if(threadIdx.x == 0) {
    const unsigned prev = atomicAdd(reinterpret_cast<unsigned*>(&my[HF_PUSH_DONE]), 1u);
    if(prev == gridDim.x - 1) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(&my[HF_PUSH_DONE]) = 0u;
      vmy[HF_PUSH_EPOCH] = epoch + 1;
      if(peer_lo_flags) // I am the HIGH neighbour of rank-1
        *static_cast<volatile unsigned long long*>(&peer_lo_flags[HF_ARRIVED_HI]) = epoch;
      if(peer_hi_flags)
        *static_cast<volatile unsigned long long*>(&peer_hi_flags[HF_ARRIVED_LO]) = epoch;
      __threadfence_system();
    }
  }
}

// Waits for the neighbours' rows of this epoch, moves them from the landing buffers into the halo rows
// of the field, then tells the neighbours that the buffers are free again.
__global__ void __launch_bounds__(256)
halo_unpack_p2p_kernel(double* __restrict__ field, unsigned long long* my, const double* recv_lo,
                       const double* recv_hi, unsigned long long* peer_lo_flags,
                       unsigned long long* peer_hi_flags, int ni, int nk, int sj, int sk,
                       int j_lo_first, int rows_lo, int j_hi_first, int rows_hi) {
  volatile unsigned long long* vmy = my;
  const unsigned long long epoch = vmy[HF_UNPACK_EPOCH];
  if(threadIdx.x == 0) {
    if(recv_lo)
      spin_until(&vmy[HF_ARRIVED_LO], epoch, &my[HF_ERROR]);
    if(recv_hi)
      spin_until(&vmy[HF_ARRIVED_HI], epoch, &my[HF_ERROR]);
    __threadfence_system();
  }
  __syncthreads();
  const size_t total_lo = recv_lo ? (size_t)ni * rows_lo * nk : 0;
  const size_t total_hi = recv_hi ? (size_t)ni * rows_hi * nk : 0;
  for(size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total_lo + total_hi;
      idx += (size_t)gridDim.x * blockDim.x) {
    const bool lo = idx < total_lo;
    const size_t e = lo ? idx : idx - total_lo;
    const int rows = lo ? rows_lo : rows_hi;
    const int i = (int)(e % ni);
    const int r = (int)((e / ni) % rows);
    const int k = (int)(e / ((size_t)ni * rows));
    const size_t f = (size_t)i + (size_t)((lo ? j_lo_first : j_hi_first) + r) * sj + (size_t)k * sk;
    field[f] = __ldcg(&(lo ? recv_lo : recv_hi)[e]); // written by a peer: read at L2, never from L1
  }
  __threadfence_system();
  __syncthreads();
  if(threadIdx.x == 0) {
    const unsigned prev = atomicAdd(reinterpret_cast<unsigned*>(&my[HF_UNPACK_DONE]), 1u);
    if(prev == gridDim.x - 1) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(&my[HF_UNPACK_DONE]) = 0u;
      vmy[HF_UNPACK_EPOCH] = epoch + 1;
      if(peer_lo_flags) // rank-1 pushes into my recv_lo and waits on ITS credit from the high side
        *static_cast<volatile unsigned long long*>(&peer_lo_flags[HF_CREDIT_HI]) = epoch + 1;
      if(peer_hi_flags)
        *static_cast<volatile unsigned long long*>(&peer_hi_flags[HF_CREDIT_LO]) = epoch + 1;
      __threadfence_system();
    }
  }
}
} // namespace

template <class T>
static int verify_impl(const T* dsl, const T* ref, size_t n, double rel_tol, double abs_tol, int iteration,
                       b200_verification_metrics* out, void* stream) {
  if(!out)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_verify_field: out is null");
  const double inf = std::numeric_limits<double>::infinity(); // thrust::reduce's initial values
  VerifyAcc init{0.0, inf, 0.0, inf, 0ull};
  VerifyAcc* d = nullptr;
  CK(cudaMalloc(&d, sizeof(VerifyAcc)));
  CK(cudaMemcpyAsync(d, &init, sizeof(init), cudaMemcpyHostToDevice, S(stream)));
  if(n) {
    int blocks = (int)((n + 255) / 256);
    if(blocks > 148 * 8)
      blocks = 148 * 8;
    verify_kernel<<<blocks, 256, 0, S(stream)>>>(dsl, ref, n, rel_tol, abs_tol, d);
  }
  VerifyAcc h;
  cudaError_t e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, S(stream));
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(S(stream));
  cudaFree(d);
  if(e != cudaSuccess)
    return cudaFail(e, "b200rt_verify_field");
  out->iteration = iteration;
  out->max_rel_err = h.max_rel, out->max_abs_err = h.max_abs;
  out->min_rel_err = h.min_rel, out->min_abs_err = h.min_abs; // inf over an empty field, as thrust's
  out->n_failed = (long long)h.n_failed;
  out->is_valid = h.n_failed == 0;
  return 0;
}
template <class T>
static int reshape_impl(const T* in, T* out, int k_size, int n_elem, int sparse,
                        bool back, void* stream) {
  if(k_size <= 0 || n_elem <= 0 || sparse <= 0)
    return 0;
  dim3 grid((n_elem + 31) / 32, (k_size + 31) / 32, sparse);
  reshape_kernel<<<grid, 256, 0, S(stream)>>>(in, out, k_size, n_elem, sparse, back);
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

int b200rt_verify_field(const double* dsl, const double* ref, size_t n, double rel_tol,
                        double abs_tol, int iteration, b200_verification_metrics* out, void* stream) {
  return verify_impl(dsl, ref, n, rel_tol, abs_tol, iteration, out, stream);
}
int b200rt_verify_field_f32(const float* dsl, const float* ref, size_t n, double rel_tol,
                            double abs_tol, int iteration, b200_verification_metrics* out, void* stream) {
  return verify_impl(dsl, ref, n, rel_tol, abs_tol, iteration, out, stream);
}

int b200rt_graph_begin(void* stream) {
  // thread-local capture mode: helper threads of the process (NCCL proxies, a framework's watchdog) keep
  // making CUDA calls while this thread records
  CK(cudaStreamBeginCapture(S(stream), cudaStreamCaptureModeThreadLocal));
  return 0;
}
int b200rt_graph_end(void* stream, b200_graph** graph) {
  if(!graph)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_graph_end: null argument");
  cudaGraph_t g = nullptr;
  CK(cudaStreamEndCapture(S(stream), &g));
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
  cudaGraphDestroy(g);
  if(e != cudaSuccess)
    return cudaFail(e, "cudaGraphInstantiate");
  *graph = new b200_graph{exec};
  return 0;
}
int b200rt_graph_launch(b200_graph* graph, void* stream) {
  if(!graph)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_graph_launch: null graph");
  CK(cudaGraphLaunch(graph->exec, S(stream)));
  return 0;
}
int b200rt_graph_destroy(b200_graph* graph) {
  if(graph) {
    cudaGraphExecDestroy(graph->exec);
    delete graph;
  }
  return 0;
}

int b200rt_metrics_record_create(b200_metrics_record** rec) {
  if(!rec)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_metrics_record_create: null argument");
  *rec = new b200_metrics_record();
  return 0;
}
int b200rt_metrics_record_add(b200_metrics_record* rec, const char* stencil, const char* field,
                              const b200_verification_metrics* m) {
  if(!rec || !stencil || !field || !m || m->iteration < 0)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_metrics_record_add: bad argument");
  auto& iterations = rec->stencils[stencil];
  // to_json.cpp:19-24: a new entry is appended when the iteration does not exist yet, otherwise the
  // field is inserted into the existing entry unless it is already there
  if((int)iterations.size() < m->iteration + 1)
    iterations.emplace_back();
  auto& entry = (int)iterations.size() > m->iteration ? iterations[m->iteration] : iterations.back();
  entry.emplace(field, *m);
  return 0;
}
const char* b200rt_metrics_record_json(b200_metrics_record* rec) {
  if(!rec)
    return "null";
  auto num = [](double v) {
    char buf[40];
    if(v != v || v - v != 0)
      return std::string("null"); // nlohmann dumps non-finite numbers as null
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s(buf);
    if(s.find_first_of(".eEn") == std::string::npos)
      s += ".0";
    return s;
  };
  auto quote = [](const std::string& s) {
    std::string o = "\"";
    for(char c : s) {
      if(c == '"' || c == '\\')
        o += '\\';
      o += c;
    }
    return o + "\"";
  };
  std::string& o = rec->text;
  if(rec->stencils.empty()) {
    o = "null";
    return o.c_str();
  }
  o = "{";
  bool firstS = true;
  for(auto const& st : rec->stencils) {
    o += (firstS ? "" : ",") + quote(st.first) + ":[";
    firstS = false;
    for(size_t it = 0; it < st.second.size(); ++it) {
      o += it ? ",{" : "{";
      bool firstF = true;
      for(auto const& f : st.second[it]) {
        auto const& m = f.second;
        o += (firstF ? "" : ",") + quote(f.first) + ":{\"field_is_valid\":" +
             (m.is_valid ? "true" : "false") + ",\"max_absolute_error\":" + num(m.max_abs_err) +
             ",\"max_relative_error\":" + num(m.max_rel_err) + ",\"min_absolute_error\":" +
             num(m.min_abs_err) + ",\"min_relative_error\":" + num(m.min_rel_err) + "}";
        firstF = false;
      }
      o += "}";
    }
    o += "]";
  }
  o += "}";
  return o.c_str();
}
int b200rt_metrics_record_write(b200_metrics_record* rec, const char* path) {
  if(!rec || !path)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_metrics_record_write: null argument");
  FILE* f = std::fopen(path, "w");
  if(!f)
    return b200rt_set_error(B200_ERR_RUNTIME, "b200rt_metrics_record_write: cannot open file");
  std::fputs(b200rt_metrics_record_json(rec), f);
  std::fclose(f);
  return 0;
}
int b200rt_metrics_record_destroy(b200_metrics_record* rec) {
  delete rec;
  return 0;
}

int b200rt_reshape(const double* in, double* out, int k, int n, int sp, void* s) {
  return reshape_impl(in, out, k, n, sp, false, s);
}
int b200rt_reshape_back(const double* in, double* out, int k, int n, int sp, void* s) {
  return reshape_impl(in, out, k, n, sp, true, s);
}
int b200rt_reshape_f32(const float* in, float* out, int k, int n, int sp, void* s) {
  return reshape_impl(in, out, k, n, sp, false, s);
}
int b200rt_reshape_back_f32(const float* in, float* out, int k, int n, int sp, void* s) {
  return reshape_impl(in, out, k, n, sp, true, s);
}

int b200rt_upload_nbh_table(const int* row_major, int n_elem, int nbh, int* d_soa, void* stream) {
  size_t cnt = (size_t)n_elem * nbh;
  if(!cnt)
    return 0;
  int* tmp = nullptr;
  CK(cudaMalloc(&tmp, cnt * sizeof(int)));
  cudaError_t e = cudaMemcpyAsync(tmp, row_major, cnt * sizeof(int), cudaMemcpyHostToDevice, S(stream));
  if(e == cudaSuccess) {
    nbh_transpose_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, S(stream)>>>(tmp, d_soa, n_elem, nbh);
    e = cudaGetLastError();
  }
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(S(stream));
  cudaFree(tmp);
  if(e != cudaSuccess)
    return cudaFail(e, "b200rt_upload_nbh_table");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// NCCL (resolved at run time so that the library loads on boxes without libnccl)
// ------------------------------------------------------------------------------------------------
namespace {
struct Id128 {
  char internal[128];
};
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;

int loadNccl() {
  if(g_nccl.lib)
    return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for(const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if(g_nccl.lib)
      break;
  }
  if(!g_nccl.lib)
    return b200rt_set_error(B200_ERR_NCCL, "libnccl.so.2 not found");
#define SYM(field, name)                                                                           \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(g_nccl.lib, name);                              \
  if(!g_nccl.field)                                                                                \
    return b200rt_set_error(B200_ERR_NCCL, "missing NCCL symbol " name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}
int ncclFail(int r, const char* what) {
  g_err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return B200_ERR_NCCL;
}
#define NK(call)                                                                                   \
  do {                                                                                             \
    int r_ = (call);                                                                               \
    if(r_ != 0)                                                                                    \
      return ncclFail(r_, #call);                                                                  \
  } while(0)
constexpr int kNcclFloat64 = 8; // ncclDouble
} // namespace

int b200rt_nccl_unique_id(void* id128) {
  if(int e = loadNccl())
    return e;
  NK(g_nccl.GetUniqueId(id128));
  return 0;
}
int b200rt_comm_init_rank(void** comm, int nranks, const void* id128, int rank) {
  if(int e = loadNccl())
    return e;
  Id128 id;
  std::memcpy(&id, id128, sizeof(id));
  NK(g_nccl.CommInitRank(comm, nranks, id, rank));
  return 0;
}
int b200rt_comm_destroy(void* comm) {
  if(int e = loadNccl())
    return e;
  NK(g_nccl.CommDestroy(comm));
  return 0;
}

// ---- chained unstructured launches ------------------------------------------------------------------
struct b200_chain_plan {
  b200_chain_view v{};
  unsigned epoch = 0;
};

int b200rt_chain_plan_create(b200_chain_plan** plan, int n_consumer, int block, int n_roles,
                             const int* n_producer, const int* const* tables, const int* table_role,
                             const int* table_nbh, int n_tables, int level_blocks, int lag_blocks,
                             void* stream) {
  if(!plan || n_consumer < 0 || block <= 0 || n_roles < 1 || n_roles > 2 || !n_producer || level_blocks < 1)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_chain_plan_create: bad argument");
  // host copies of the neighbour tables (index work, once per stencil instance)
  std::vector<std::vector<int>> host(n_tables);
  std::vector<dawn_b200::ChainTable> tabs;
  for(int t = 0; t < n_tables; ++t) {
    if(!tables[t] || table_role[t] < 0 || table_role[t] >= n_roles || table_nbh[t] <= 0)
      return b200rt_set_error(B200_ERR_ARG, "b200rt_chain_plan_create: bad table");
    host[t].resize((size_t)table_nbh[t] * n_consumer);
    CK(cudaMemcpyAsync(host[t].data(), tables[t], host[t].size() * sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    tabs.push_back({host[t].data(), table_nbh[t], table_role[t]});
  }
  CK(cudaStreamSynchronize(S(stream)));
  dawn_b200::ChainSchedule sc =
      dawn_b200::buildChainSchedule(n_consumer, block, n_roles, n_producer, tabs, lag_blocks > 0 ? lag_blocks : 2048);
  auto* p = new b200_chain_plan();
  p->v.sched_len = (unsigned)sc.sched.size();
  p->v.nb0 = sc.nb[0], p->v.nb1 = sc.nb[1], p->v.nc = sc.nc, p->v.level_blocks = (unsigned)level_blocks;
  unsigned *d_sched = nullptr, *d_flags = nullptr, *d_err = nullptr;
  int* d_need = nullptr;
  const size_t nflags = (size_t)level_blocks * (sc.nb[0] + sc.nb[1]);
  cudaError_t e = cudaMalloc(&d_sched, std::max<size_t>(1, sc.sched.size()) * sizeof(unsigned));
  if(e == cudaSuccess) e = cudaMalloc(&d_need, std::max<size_t>(1, sc.need.size()) * sizeof(int));
  if(e == cudaSuccess) e = cudaMalloc(&d_flags, std::max<size_t>(1, nflags) * sizeof(unsigned));
  if(e == cudaSuccess) e = cudaMalloc(&d_err, sizeof(unsigned));
  if(e == cudaSuccess) e = cudaMemcpyAsync(d_sched, sc.sched.data(), sc.sched.size() * sizeof(unsigned), cudaMemcpyHostToDevice, S(stream));
  if(e == cudaSuccess) e = cudaMemcpyAsync(d_need, sc.need.data(), sc.need.size() * sizeof(int), cudaMemcpyHostToDevice, S(stream));
  if(e == cudaSuccess) e = cudaMemsetAsync(d_flags, 0, std::max<size_t>(1, nflags) * sizeof(unsigned), S(stream));
  if(e == cudaSuccess) e = cudaMemsetAsync(d_err, 0, sizeof(unsigned), S(stream));
  if(e == cudaSuccess) e = cudaStreamSynchronize(S(stream)); // the host vectors go out of scope
  if(e != cudaSuccess) {
    cudaFree(d_sched), cudaFree(d_need), cudaFree(d_flags), cudaFree(d_err);
    delete p;
    return cudaFail(e, "b200rt_chain_plan_create");
  }
  p->v.sched = d_sched, p->v.need = d_need, p->v.flags = d_flags, p->v.error = d_err;
  *plan = p;
  return 0;
}
int b200rt_chain_plan_view(const b200_chain_plan* plan, b200_chain_view* view) {
  if(!plan || !view)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_chain_plan_view: null argument");
  *view = plan->v;
  return 0;
}
unsigned b200rt_chain_plan_next_epoch(b200_chain_plan* plan) {
  if(++plan->epoch == 0) // wrapped: 0 is the "never published" value of the flags
    ++plan->epoch;
  return plan->epoch;
}
int b200rt_chain_plan_status(b200_chain_plan* plan, void* stream) {
  if(!plan)
    return 0;
  unsigned err = 0;
  CK(cudaMemcpyAsync(&err, plan->v.error, sizeof(err), cudaMemcpyDeviceToHost, S(stream)));
  CK(cudaStreamSynchronize(S(stream)));
  if(err)
    return b200rt_set_error(B200_ERR_RUNTIME, "chained launch: a consumer block gave up waiting for its producers "
                                              "(blocks were not dispatched in index order?); results are invalid");
  return 0;
}
int b200rt_chain_plan_destroy(b200_chain_plan* plan) {
  if(!plan)
    return 0;
  cudaFree(const_cast<unsigned*>(plan->v.sched)), cudaFree(const_cast<int*>(plan->v.need));
  cudaFree(plan->v.flags), cudaFree(plan->v.error);
  delete plan;
  return 0;
}
} // extern "C"

struct b200_halo_plan {
  void* comm;
  int rank, nranks, ni, nj, nk, sj, sk, halo, halo_hi, wminus, wplus;
  double *send_lo = nullptr, *send_hi = nullptr, *recv_lo = nullptr, *recv_hi = nullptr;
  // flags + both landing buffers live in ONE allocation that the slab neighbours map through CUDA IPC
  unsigned long long* arena = nullptr;
  size_t cap = 0; // bytes of one landing buffer (a multiple of 256)
  void *peer_lo = nullptr, *peer_hi = nullptr; // mapped arenas of rank-1 / rank+1
  bool connected = false;
};
namespace {
constexpr size_t kArenaHead = HF_WORDS * sizeof(unsigned long long);
inline double* arenaRecvLo(void* arena) {
  return reinterpret_cast<double*>(static_cast<char*>(arena) + kArenaHead);
}
inline double* arenaRecvHi(void* arena, size_t cap) {
  return reinterpret_cast<double*>(static_cast<char*>(arena) + kArenaHead + cap);
}
} // namespace

extern "C" {
int b200rt_halo_plan_create(b200_halo_plan** plan, void* comm, int rank, int nranks, int ni, int nj,
                            int nk, int stride_j, int stride_k, int halo_j, int width_minus,
                            int width_plus) {
  return b200rt_halo_plan_create2(plan, comm, rank, nranks, ni, nj, nk, stride_j, stride_k, halo_j,
                                  halo_j, width_minus, width_plus);
}

int b200rt_halo_plan_create2(b200_halo_plan** plan, void* comm, int rank, int nranks, int ni, int nj,
                             int nk, int stride_j, int stride_k, int halo_lo, int halo_hi,
                             int width_minus, int width_plus) {
  const bool has_lo = rank > 0, has_hi = rank + 1 < nranks;
  if(!plan || width_minus < 0 || width_plus < 0 || (has_lo && width_minus > halo_lo) ||
     (has_hi && width_plus > halo_hi))
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_plan_create: bad widths");
  auto* p = new b200_halo_plan{comm, rank,     nranks,   ni,     nj,         nk,
                               stride_j, stride_k, halo_lo, halo_hi, width_minus, width_plus};
  // my low halo (width_minus rows) is filled from rank-1's top interior rows; my high halo
  // (width_plus rows) from rank+1's bottom interior rows.
  size_t lo = (size_t)ni * nk * (size_t)std::max(width_minus, width_plus) * sizeof(double);
  if(lo == 0)
    lo = sizeof(double);
  CK(cudaMalloc(&p->send_lo, lo));
  CK(cudaMalloc(&p->send_hi, lo));
  p->cap = (lo + 255) / 256 * 256;
  CK(cudaMalloc(&p->arena, kArenaHead + 2 * p->cap));
  p->recv_lo = arenaRecvLo(p->arena);
  p->recv_hi = arenaRecvHi(p->arena, p->cap);
  unsigned long long head[HF_WORDS] = {};
  head[HF_CREDIT_LO] = head[HF_CREDIT_HI] = 1; // both landing buffers are free for epoch 1
  head[HF_PUSH_EPOCH] = head[HF_UNPACK_EPOCH] = 1;
  CK(cudaMemcpy(p->arena, head, sizeof(head), cudaMemcpyHostToDevice));
  *plan = p;
  return 0;
}

int b200rt_halo_plan_ipc_handle(b200_halo_plan* p, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  if(!p || !handle64)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_plan_ipc_handle: null argument");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, p->arena));
  std::memcpy(handle64, &h, sizeof(h));
  return 0;
}

int b200rt_halo_plan_connect(b200_halo_plan* p, const void* handle_lo, const void* handle_hi) {
  if(!p)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_plan_connect: null plan");
  const bool has_lo = p->rank > 0, has_hi = p->rank + 1 < p->nranks;
  if((has_lo && !handle_lo) || (has_hi && !handle_hi))
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_plan_connect: a neighbour's handle is missing");
  cudaIpcMemHandle_t h;
  if(has_lo) {
    std::memcpy(&h, handle_lo, sizeof(h));
    CK(cudaIpcOpenMemHandle(&p->peer_lo, h, cudaIpcMemLazyEnablePeerAccess));
  }
  if(has_hi) {
    std::memcpy(&h, handle_hi, sizeof(h));
    CK(cudaIpcOpenMemHandle(&p->peer_hi, h, cudaIpcMemLazyEnablePeerAccess));
  }
  p->connected = true;
  return 0;
}

int b200rt_halo_exchange_p2p(b200_halo_plan* p, double* field, void* stream) {
  if(!p || !field)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_exchange_p2p: null argument");
  if(p->nranks > 1 && !p->connected)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_exchange_p2p: plan not connected to its neighbours");
  const bool has_lo = p->rank > 0, has_hi = p->rank + 1 < p->nranks;
  if(!has_lo && !has_hi)
    return 0;
  const int j_int0 = p->halo, j_int1 = p->nj - p->halo_hi; // interior rows [j_int0, j_int1)
  auto* flo = static_cast<unsigned long long*>(p->peer_lo);
  auto* fhi = static_cast<unsigned long long*>(p->peer_hi);
  // at most one block per SM: a block that waits for a neighbour must not crowd out the stencil kernel
  // running beside it, and 148 x 256 threads saturate an NVLink port several times over
  auto blocksFor = [&](size_t total) {
    return (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, 148));
  };
  { // my first `wplus` interior rows -> high landing buffer of rank-1; my last `wminus` -> low one of rank+1
    const bool to_lo = has_lo && p->wplus, to_hi = has_hi && p->wminus;
    size_t total = (size_t)p->ni * p->nk * ((to_lo ? p->wplus : 0) + (to_hi ? p->wminus : 0));
    halo_push_kernel<<<blocksFor(total), 256, 0, S(stream)>>>(
        field, p->arena, to_lo ? arenaRecvHi(p->peer_lo, p->cap) : nullptr, to_lo ? flo : nullptr,
        to_hi ? arenaRecvLo(p->peer_hi) : nullptr, to_hi ? fhi : nullptr, p->ni, p->nk, p->sj, p->sk, j_int0,
        p->wplus, j_int1 - p->wminus, p->wminus);
  }
  { // rank-1's rows -> my low halo rows; rank+1's rows -> my high halo rows
    const bool from_lo = has_lo && p->wminus, from_hi = has_hi && p->wplus;
    size_t total = (size_t)p->ni * p->nk * ((from_lo ? p->wminus : 0) + (from_hi ? p->wplus : 0));
    halo_unpack_p2p_kernel<<<blocksFor(total), 256, 0, S(stream)>>>(
        field, p->arena, from_lo ? p->recv_lo : nullptr, from_hi ? p->recv_hi : nullptr,
        from_lo ? flo : nullptr, from_hi ? fhi : nullptr, p->ni, p->nk, p->sj, p->sk,
        j_int0 - p->wminus, p->wminus, j_int1, p->wplus);
  }
  CK(cudaGetLastError());
  return 0;
}

int b200rt_halo_plan_p2p_status(b200_halo_plan* p) {
  if(!p)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_plan_p2p_status: null plan");
  unsigned long long err = 0;
  CK(cudaMemcpy(&err, p->arena + HF_ERROR, sizeof(err), cudaMemcpyDeviceToHost));
  if(err)
    return b200rt_set_error(B200_ERR_RUNTIME, "peer-memory halo exchange: a wait for a slab neighbour ran "
                                              "into its time limit");
  return 0;
}

int b200rt_halo_exchange(b200_halo_plan* p, double* field, void* stream) {
  if(!p)
    return b200rt_set_error(B200_ERR_ARG, "b200rt_halo_exchange: null plan");
  if(p->nranks > 1)
    if(int e = loadNccl())
      return e;
  const bool has_lo = p->rank > 0, has_hi = p->rank + 1 < p->nranks;
  auto launch = [&](const double* f, double* buf, int j_first, int rows, bool unpack, double* fw) {
    size_t total = (size_t)p->ni * rows * p->nk;
    if(!total)
      return;
    int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
    halo_pack_kernel<<<blocks, 256, 0, S(stream)>>>(f, buf, p->ni, p->nk, p->sj, p->sk, j_first,
                                                    rows, unpack, fw);
  };
  // rows I send down (to rank-1): my first `wplus` interior rows fill its high halo;
  // rows I send up (to rank+1): my last `wminus` interior rows fill its low halo.
  const int j_int0 = p->halo, j_int1 = p->nj - p->halo_hi; // interior rows [j_int0, j_int1)
  if(has_lo && p->wplus)
    launch(field, p->send_lo, j_int0, p->wplus, false, nullptr);
  if(has_hi && p->wminus)
    launch(field, p->send_hi, j_int1 - p->wminus, p->wminus, false, nullptr);
  if(has_lo || has_hi) {
    NK(g_nccl.GroupStart());
    size_t n_lo_send = (size_t)p->ni * p->nk * p->wplus, n_hi_send = (size_t)p->ni * p->nk * p->wminus;
    size_t n_lo_recv = (size_t)p->ni * p->nk * p->wminus, n_hi_recv = (size_t)p->ni * p->nk * p->wplus;
    if(has_lo && n_lo_send)
      NK(g_nccl.Send(p->send_lo, n_lo_send, kNcclFloat64, p->rank - 1, p->comm, S(stream)));
    if(has_lo && n_lo_recv)
      NK(g_nccl.Recv(p->recv_lo, n_lo_recv, kNcclFloat64, p->rank - 1, p->comm, S(stream)));
    if(has_hi && n_hi_send)
      NK(g_nccl.Send(p->send_hi, n_hi_send, kNcclFloat64, p->rank + 1, p->comm, S(stream)));
    if(has_hi && n_hi_recv)
      NK(g_nccl.Recv(p->recv_hi, n_hi_recv, kNcclFloat64, p->rank + 1, p->comm, S(stream)));
    NK(g_nccl.GroupEnd());
  }
  if(has_lo && p->wminus)
    launch(nullptr, p->recv_lo, j_int0 - p->wminus, p->wminus, true, field);
  if(has_hi && p->wplus)
    launch(nullptr, p->recv_hi, j_int1, p->wplus, true, field);
  CK(cudaGetLastError());
  return 0;
}

int b200rt_halo_plan_destroy(b200_halo_plan* p) {
  if(!p)
    return 0;
  if(p->peer_lo)
    cudaIpcCloseMemHandle(p->peer_lo);
  if(p->peer_hi)
    cudaIpcCloseMemHandle(p->peer_hi);
  cudaFree(p->send_lo), cudaFree(p->send_hi), cudaFree(p->arena);
  delete p;
  return 0;
}
}
