// B200-native storages with the API surface generated code and reference-style drivers use from
// GridTools 1.0.x (cf. driver-includes/storage_runtime.hpp:39-97):
//   meta_data_t md(i, j, k);  storage_t f(md, "name");  auto v = make_host_view(f); v(i,j,k) = ..;
//   f.sync();  f.strides();  f.get_storage_info_ptr()->index(i,j,k) / total_length<D>();
// A storage owns a pinned host mirror and an HBM buffer (allocated lazily through the C-ABI
// runtime).  Layout: i fastest, rows padded to a multiple of 128 bytes and shifted so that the first
// interior point (i = halo) of every row starts a 128-byte line (16 doubles / 32 floats) -> warps reading a row issue full,
// aligned 128 B transactions.  Masked dimensions have stride 0.
#pragma once
#include "b200_runtime.hpp"
#include "defs.hpp"
#include "domain.hpp"

#include <array>
#include <cstring>
#include <memory>
#include <string>

namespace gridtools {

using uint_t = unsigned int;
using int_t = int;

enum class access_mode { read_write = 0, read_only = 1 };

template <int... Is>
struct halo {};
template <bool... Bs>
struct selector {};

namespace b200_detail {

template <bool MI, bool MJ, bool MK>
class storage_info {
  std::array<uint_t, 3> len_{1, 1, 1};
  std::array<int, 3> str_{0, 0, 0};
  int lead_ = 0;       // elements before (0,0,0): aligns i = halo to 128 B
  size_t elems_ = 1;   // allocation size in elements (lead included)

public:
  static constexpr bool has_i = MI, has_j = MJ, has_k = MK;
  storage_info() = default;
  storage_info(uint_t i, uint_t j, uint_t k) {
    len_ = {MI ? i : 1u, MJ ? j : 1u, MK ? k : 1u};
    const int h = ::gridtools::dawn::halo::value;
    size_t s = 1;
    const int line = 128 / (int)sizeof(::dawn::float_type); // elements per 128-byte line
    if(MI) {
      str_[0] = 1;
      lead_ = (line - h % line) % line;
      s = ((size_t)len_[0] + line - 1) / line * line;
    }
    if(MJ) {
      str_[1] = (int)s;
      s *= len_[1];
    }
    if(MK) {
      str_[2] = (int)s;
      s *= len_[2];
    }
    elems_ = s + lead_ + line;
  }
  template <int D>
  uint_t total_length() const {
    return len_[D];
  }
  template <int D>
  int stride() const {
    return str_[D];
  }
  const std::array<int, 3>& strides() const { return str_; }
  int index(int i, int j, int k) const { return i * str_[0] + j * str_[1] + k * str_[2]; }
  int lead() const { return lead_; }
  size_t padded_total_length() const { return elems_; }
};

template <class T>
struct buffers {
  T* host = nullptr;
  T* dev = nullptr;
  size_t elems = 0;
  bool host_dirty = false; // host has newer data than the device
  bool dev_dirty = false;  // device has newer data than the host
  void* stream = nullptr;  // stream of the last device-side producer (kernel or upload) of `dev`
  ~buffers() {
    b200rt_free_host(host);
    b200rt_free(dev);
  }
};

} // namespace b200_detail

template <class T, class Info>
class data_store {
  std::shared_ptr<Info> info_;
  std::shared_ptr<b200_detail::buffers<T>> buf_;
  std::string name_;

  void ensure_device() const {
    if(!buf_->dev) {
      void* p = nullptr;
      ::dawn_b200::check(b200rt_malloc(&p, buf_->elems * sizeof(T)));
      buf_->dev = static_cast<T*>(p);
      buf_->host_dirty = true;
    }
  }

public:
  using data_t = T;
  using storage_info_t = Info;
  data_store() = default;
  explicit data_store(const Info& info, const std::string& name = "")
      : info_(std::make_shared<Info>(info)), buf_(std::make_shared<b200_detail::buffers<T>>()),
        name_(name) {
    buf_->elems = info.padded_total_length();
    void* p = nullptr;
    ::dawn_b200::check(b200rt_malloc_host(&p, buf_->elems * sizeof(T)));
    buf_->host = static_cast<T*>(p);
    std::memset(buf_->host, 0, buf_->elems * sizeof(T));
    buf_->host_dirty = true;
  }
  data_store(const Info& info, T init, const std::string& name = "") : data_store(info, name) {
    for(size_t n = 0; n < buf_->elems; ++n)
      buf_->host[n] = init;
  }
  bool valid() const { return static_cast<bool>(buf_); }
  const std::string& name() const { return name_; }
  const Info* get_storage_info_ptr() const { return info_.get(); }
  const std::array<int, 3>& strides() const { return info_->strides(); }

  // make both copies coherent (GridTools' data_store::sync).  Copies are ordered on the stream that
  // last produced the device copy - generated stencils launch on their own (possibly non-blocking)
  // stream, which the legacy null stream does not wait for - and the host waits for that stream.
  void sync() const {
    if(!buf_)
      return;
    if(buf_->dev_dirty && buf_->dev) {
      ::dawn_b200::check(b200rt_memcpy_d2h(buf_->host, buf_->dev, buf_->elems * sizeof(T), buf_->stream));
      ::dawn_b200::check(b200rt_stream_synchronize(buf_->stream));
      buf_->dev_dirty = false;
    } else if(buf_->host_dirty && buf_->dev) {
      ::dawn_b200::check(b200rt_memcpy_h2d(buf_->dev, buf_->host, buf_->elems * sizeof(T), buf_->stream));
      ::dawn_b200::check(b200rt_stream_synchronize(buf_->stream));
      buf_->host_dirty = false;
    }
  }
  // host pointer to element (0,0,0); read_write access marks the host copy as the newer one
  T* host_ptr(bool will_write) const {
    if(buf_->dev_dirty)
      sync();
    if(will_write)
      buf_->host_dirty = true;
    return buf_->host + info_->lead();
  }
  // descriptor of the up-to-date device copy for work on `stream` (uploads the host copy on that stream
  // if it is newer; if another stream produced the device copy, the host waits for it first)
  b200_field_t device_descriptor(void* stream = nullptr) const {
    ensure_device();
    if(buf_->dev_dirty && buf_->stream != stream)
      ::dawn_b200::check(b200rt_stream_synchronize(buf_->stream));
    if(buf_->host_dirty) {
      ::dawn_b200::check(b200rt_memcpy_h2d(buf_->dev, buf_->host, buf_->elems * sizeof(T), stream));
      ::dawn_b200::check(b200rt_stream_synchronize(stream));
      buf_->host_dirty = false;
      buf_->stream = stream;
    }
    return b200_field_t{buf_->dev + info_->lead(), info_->template stride<1>(),
                        info_->template stride<2>()};
  }
  T* device_ptr() const { return static_cast<T*>(device_descriptor().data); }

  // ---- pieces of the host <-> device pipeline of generated run() methods (set_host_pipeline): levels
  // [k0, k1) of a field with a k dimension are one contiguous range (k is the slowest dimension); fields
  // without one travel whole.  All copies are asynchronous on `stream`; the caller orders and waits.
  bool host_is_newer() const { return buf_ && (buf_->host_dirty || !buf_->dev); }
  bool device_is_newer() const { return buf_ && buf_->dev && buf_->dev_dirty; }
  void* producer_stream() const { return buf_ ? buf_->stream : nullptr; }
  b200_field_t device_descriptor_no_upload() const {
    ensure_device();
    return b200_field_t{buf_->dev + info_->lead(), info_->template stride<1>(), info_->template stride<2>()};
  }
  void level_range(int k0, int k1, size_t& first, size_t& count) const {
    if(Info::has_k) {
      const size_t sk = (size_t)info_->template stride<2>();
      first = (size_t)info_->lead() + (size_t)k0 * sk, count = (size_t)(k1 - k0) * sk;
      if(k0 == 0)
        first = 0, count += (size_t)info_->lead(); // the lead pad travels with the first level
    } else
      first = 0, count = buf_->elems;
    if(first + count > buf_->elems)
      count = buf_->elems - first;
  }
  void upload_levels(int k0, int k1, void* stream) const {
    ensure_device();
    size_t first, count;
    level_range(k0, k1, first, count);
    ::dawn_b200::check(b200rt_memcpy_h2d(buf_->dev + first, buf_->host + first, count * sizeof(T), stream));
  }
  void download_levels(int k0, int k1, void* stream) const {
    size_t first, count;
    level_range(k0, k1, first, count);
    ::dawn_b200::check(b200rt_memcpy_d2h(buf_->host + first, buf_->dev + first, count * sizeof(T), stream));
  }
  // both copies hold the same data (after the pipeline has moved every level); `stream` = last device producer
  void mark_coherent(void* stream) const {
    buf_->host_dirty = buf_->dev_dirty = false;
    buf_->stream = stream;
  }
  // a kernel on `stream` writes the device copy
  void mark_device_dirty(void* stream = nullptr) const {
    buf_->dev_dirty = true;
    buf_->stream = stream;
  }
};

template <class Store, access_mode Mode = access_mode::read_write>
class data_view {
  typename Store::data_t* ptr_ = nullptr;
  typename Store::storage_info_t info_;

public:
  using data_t = typename Store::data_t;
  data_view() = default;
  explicit data_view(const Store& s)
      : ptr_(s.host_ptr(Mode == access_mode::read_write)), info_(*s.get_storage_info_ptr()) {}
  data_t& operator()(int i, int j, int k) const { return ptr_[info_.index(i, j, k)]; }
  data_t* data() const { return ptr_; }
};

template <access_mode Mode = access_mode::read_write, class Store>
data_view<Store, Mode> make_host_view(const Store& s) {
  return data_view<Store, Mode>(s);
}

namespace dawn {

using meta_data_ijk_t = ::gridtools::b200_detail::storage_info<true, true, true>;
using meta_data_ij_t = ::gridtools::b200_detail::storage_info<true, true, false>;
using meta_data_i_t = ::gridtools::b200_detail::storage_info<true, false, false>;
using meta_data_j_t = ::gridtools::b200_detail::storage_info<false, true, false>;
using meta_data_k_t = ::gridtools::b200_detail::storage_info<false, false, true>;
using meta_data_scalar_t = ::gridtools::b200_detail::storage_info<false, false, false>;
using meta_data_t = meta_data_ijk_t;

using storage_ijk_t = ::gridtools::data_store<::dawn::float_type, meta_data_ijk_t>;
using storage_ij_t = ::gridtools::data_store<::dawn::float_type, meta_data_ij_t>;
using storage_i_t = ::gridtools::data_store<::dawn::float_type, meta_data_i_t>;
using storage_j_t = ::gridtools::data_store<::dawn::float_type, meta_data_j_t>;
using storage_k_t = ::gridtools::data_store<::dawn::float_type, meta_data_k_t>;
using storage_scalar_t = ::gridtools::data_store<::dawn::float_type, meta_data_scalar_t>;
using storage_t = storage_ijk_t;

} // namespace dawn
} // namespace gridtools
