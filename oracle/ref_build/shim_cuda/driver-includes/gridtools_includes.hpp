// TEST / BENCH INFRASTRUCTURE ONLY (oracle/_ref build).  Stand-in for the headers the reference's
// generated CUDA text includes (`#include <driver-includes/gridtools_includes.hpp>`, CudaCodeGen.cpp:731)
// so that its pre-generated .cpp/.cu files compile UNMODIFIED for sm_100a without GridTools: a domain,
// device storages that wrap caller-owned device pointers, the views / traits / timer names the
// generated wrapper class mentions.  API actually used by the texts: domain::{i,j,k}size/minus/plus,
// storage_ijk_t::{strides, get_storage_info_ptr()->index, sync}, gridtools::make_device_view +
// data_view::data, gridtools::halo, storage_traits_t::{storage_info_t, data_store_t} (temporaries the
// kernels never touch when they are IJ-cached), timer_cuda::{start, pause, reset, total_time},
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
struct halo {};

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

// a storage that wraps a device buffer the caller owns (element (0,0,0), i-stride 1)
class device_storage {
  ::dawn::float_type* ptr_ = nullptr;
  std::shared_ptr<device_storage_info> info_;

public:
  device_storage() = default;
  device_storage(::dawn::float_type* dev, int stride_j, int stride_k)
      : ptr_(dev), info_(std::make_shared<device_storage_info>()) {
    info_->str = {1, stride_j, stride_k};
  }
  const std::array<int, 3>& strides() const { return info_->str; }
  const device_storage_info* get_storage_info_ptr() const { return info_.get(); }
  void sync() const {} // nothing lives on the host
  ::dawn::float_type* device_ptr() const { return ptr_; }
};
using storage_ijk_t = device_storage;
using storage_t = device_storage;

// storages of temporaries: the IJ-cached ones are declared and constructed by the wrapper class but never
// passed to a kernel
struct storage_traits_t {
  template <int Id, int N, class Halo>
  struct storage_info_t {
    template <class... A>
    storage_info_t(A...) {}
  };
  template <class T, class Info>
  struct data_store_t {
    data_store_t(const Info&) {}
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

template <class Store>
class data_view {
  ::dawn::float_type* ptr_ = nullptr;

public:
  data_view() = default;
  explicit data_view(::dawn::float_type* p) : ptr_(p) {}
  ::dawn::float_type* data() const { return ptr_; }
};
template <class Store>
data_view<Store> make_device_view(const Store& s) {
  return data_view<Store>(s.device_ptr());
}
} // namespace gridtools
