// TEST / BENCH INFRASTRUCTURE ONLY (oracle/_ref build).  Stand-in for the headers the reference's
// generated CUDA text includes (`#include <driver-includes/gridtools_includes.hpp>`, CudaCodeGen.cpp:731)
// so that its pre-generated .cpp/.cu files compile UNMODIFIED for sm_100a without GridTools: a domain,
// device storages that wrap caller-owned device pointers, the views / traits / timer names the
// generated wrapper class mentions.  API actually used by the texts: domain::{i,j,k}size/minus/plus,
// storage_ijk_t::{strides, get_storage_info_ptr()->index, sync}, gridtools::make_device_view +
// data_view::data, gridtools::halo, storage_traits_t::{storage_info_t, data_store_t} (temporaries the
// kernels only see when they are not IJ-cached: 5-D block-private arrays), meta_data_t / storage_t for
// fields the wrapper class allocates itself, timer_cuda::{start, pause, reset, total_time},
// ::dawn::driver::cartesian_extent.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <memory>
#include <string>

namespace dawn {
using float_type = double;
namespace driver {
struct cartesian_extent {
  int iminus, iplus, jminus, jplus, kminus, kplus;
};
} // namespace driver
} // namespace dawn

namespace gridtools {
template <int... Is>
struct halo {
  static constexpr int at(int d) {
    constexpr int v[] = {Is..., 0};
    return v[d];
  }
};

namespace dawn {
// driver-includes/domain.hpp:42-58
class domain {
  std::array<unsigned, 3> dims_;
  std::array<unsigned, 6> halos_{0, 0, 0, 0, 0, 0};

public:
  domain(unsigned i, unsigned j, unsigned k) : dims_{i, j, k} {}
  void set_halos(unsigned im, unsigned ip, unsigned jm, unsigned jp, unsigned km, unsigned kp) {
    halos_ = {im, ip, jm, jp, km, kp};
  }
  unsigned isize() const { return dims_[0]; }
  unsigned jsize() const { return dims_[1]; }
  unsigned ksize() const { return dims_[2]; }
  unsigned iminus() const { return halos_[0]; }
  unsigned iplus() const { return halos_[1]; }
  unsigned jminus() const { return halos_[2]; }
  unsigned jplus() const { return halos_[3]; }
  unsigned kminus() const { return halos_[4]; }
  unsigned kplus() const { return halos_[5]; }
};

struct device_storage_info {
  std::array<int, 3> str{1, 0, 0};
  int index(int i, int j, int k) const { return i * str[0] + j * str[1] + k * str[2]; }
};

struct device_buffer {
  ::dawn::float_type* ptr = nullptr;
  ~device_buffer() { cudaFree(ptr); }
};

// dense ijk layout, i fastest (what the wrapper class allocates its own fields with, e.g. versioned ones)
struct meta_data_t {
  int ni, nj, nk;
  meta_data_t(int i, int j, int k) : ni(i), nj(j), nk(k) {}
};

// a storage that wraps a device buffer the caller owns (element (0,0,0), i-stride 1), or owns a dense one
class device_storage {
  ::dawn::float_type* ptr_ = nullptr;
  std::shared_ptr<device_storage_info> info_;
  std::shared_ptr<device_buffer> own_;

public:
  device_storage() = default;
  device_storage(::dawn::float_type* dev, int stride_j, int stride_k)
      : ptr_(dev), info_(std::make_shared<device_storage_info>()) {
    info_->str = {1, stride_j, stride_k};
  }
  device_storage(const meta_data_t& md, const std::string& = "")
      : info_(std::make_shared<device_storage_info>()), own_(std::make_shared<device_buffer>()) {
    info_->str = {1, md.ni, md.ni * md.nj};
    const size_t bytes = sizeof(::dawn::float_type) * md.ni * md.nj * md.nk;
    cudaMalloc(&own_->ptr, bytes);
    cudaMemset(own_->ptr, 0, bytes);
    ptr_ = own_->ptr;
  }
  const std::array<int, 3>& strides() const { return info_->str; }
  const device_storage_info* get_storage_info_ptr() const { return info_.get(); }
  void sync() const {} // nothing lives on the host
  ::dawn::float_type* device_ptr() const { return ptr_; }
  void view_strides(int* s) const { s[0] = 1, s[1] = info_->str[1], s[2] = info_->str[2], s[3] = s[4] = 0; }
};
using storage_ijk_t = device_storage;
using storage_t = device_storage;

// storages of temporaries: [block i + extent][block j + extent][blocks in i][blocks in j][k]
// (CudaCodeGen.cpp:672-694), dimension 0 fastest; allocated when a device view is first asked for -
// IJ-cached temporaries are declared by the wrapper class but never reach a kernel
struct storage_traits_t {
  template <int Id, int N, class Halo>
  struct storage_info_t {
    static_assert(N == 5, "temporaries of the CUDA backend are 5-dimensional");
    int len[5], str[5];
    storage_info_t(int a, int b, int c, int d, int e) : len{a, b, c, d, e} {
      str[0] = 1;
      for(int n = 1; n < 5; ++n)
        str[n] = str[n - 1] * len[n - 1];
    }
    template <int D>
    int begin() const {
      return Halo::at(D);
    }
    template <int D>
    int stride() const {
      return str[D];
    }
    size_t size() const { return (size_t)str[4] * len[4]; }
  };
  template <class T, class Info>
  struct data_store_t {
    Info info;
    std::shared_ptr<device_buffer> buf = std::make_shared<device_buffer>();
    data_store_t(const Info& i) : info(i) {}
    const Info* get_storage_info_ptr() const { return &info; }
    T* device_ptr() const {
      if(!buf->ptr) {
        cudaMalloc(&buf->ptr, sizeof(T) * info.size());
        cudaMemset(buf->ptr, 0, sizeof(T) * info.size());
      }
      return buf->ptr;
    }
    void view_strides(int* s) const {
      for(int n = 0; n < 5; ++n)
        s[n] = info.str[n];
    }
  };
};

// driver-includes/timer_cuda.hpp:39-84 (start/pause bracket the launches with events; pause waits)
class timer_cuda {
  std::string name_;
  cudaEvent_t a_, b_;
  double total_ = 0;

public:
  explicit timer_cuda(std::string name) : name_(std::move(name)) {
    cudaEventCreate(&a_);
    cudaEventCreate(&b_);
  }
  ~timer_cuda() {
    cudaEventDestroy(a_);
    cudaEventDestroy(b_);
  }
  void start() { cudaEventRecord(a_, 0); }
  void pause() {
    cudaEventRecord(b_, 0);
    cudaEventSynchronize(b_);
    float ms = 0;
    cudaEventElapsedTime(&ms, a_, b_);
    total_ += ms * 1e-3;
  }
  void reset() { total_ = 0; }
  double total_time() const { return total_; }
};
} // namespace dawn

// passed to kernels by value (temporaries): pointer + up to five strides
template <class Store>
class data_view {
  ::dawn::float_type* ptr_ = nullptr;
  int str_[5] = {1, 0, 0, 0, 0};

public:
  data_view() = default;
  data_view(::dawn::float_type* p, const int* s) : ptr_(p) {
    for(int n = 0; n < 5; ++n)
      str_[n] = s[n];
  }
  __host__ __device__ ::dawn::float_type* data() const { return ptr_; }
  __host__ __device__ ::dawn::float_type& operator()(int i, int j, int a, int b, int k) const {
    return ptr_[i * str_[0] + j * str_[1] + a * str_[2] + b * str_[3] + k * str_[4]];
  }
};
template <class Store>
data_view<Store> make_device_view(const Store& s) {
  int str[5];
  s.view_strides(str);
  return data_view<Store>(s.device_ptr(), str);
}
} // namespace gridtools
