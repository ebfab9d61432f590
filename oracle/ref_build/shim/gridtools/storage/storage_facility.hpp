// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not part of the shipped product.
//
// Minimal stand-in for the three GridTools 1.0.x headers that the reference's
// `dawn/src/driver-includes/storage_runtime.hpp:26-28` includes.  GridTools is a
// network-fetched dependency of the reference (README.md:36, "GridTools == 1.0.4")
// and is absent from this image, so the reference's own pre-generated naive
// stencils (dawn/test/unit-test/dawn/CodeGen/reference/*.cpp) and its
// `verify.hpp` are compiled UNMODIFIED against this shim.
//
// Only the API surface those files actually touch is provided:
//   halo<...>, selector<...>, backend::{mc,x86,cuda}, access_mode,
//   storage_traits<B>::{storage_info_t, special_storage_info_t, data_store_t},
//   storage_info(i,j,k) / total_length<D>() / index(i,j,k) / stride<D>() / padded_total_length(),
//   data_store(info[,name]) / sync() / name() / get_storage_info_ptr() / strides(),
//   data_view::operator()(i,j,k) / data(),  make_host_view[<mode>](store).
//
// Layout: i fastest, then j, then k (masked dimensions get stride 0), no padding.
// Parity never depends on the layout because every comparison goes through
// view(i,j,k) (driver-includes/verify.hpp:135-150).
#pragma once

#include <array>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

namespace gridtools {

using uint_t = unsigned int;
using int_t = int;

template <int... Is>
struct halo {};

template <bool... Bs>
struct selector {
  static constexpr std::array<bool, sizeof...(Bs)> value{Bs...};
};

namespace backend {
struct mc {};
struct x86 {};
struct cuda {};
struct naive {};
} // namespace backend

enum class access_mode { read_write = 0, read_only = 1 };

namespace shim_detail {

template <bool MI, bool MJ, bool MK>
class storage_info_impl {
  std::array<uint_t, 3> dims_{1, 1, 1};
  std::array<int, 3> strides_{0, 0, 0};
  std::size_t size_ = 1;

public:
  storage_info_impl() = default;
  storage_info_impl(uint_t i, uint_t j, uint_t k) {
    dims_ = {MI ? i : 1u, MJ ? j : 1u, MK ? k : 1u};
    int s = 1;
    strides_[0] = MI ? s : 0;
    s *= dims_[0];
    strides_[1] = MJ ? s : 0;
    s *= dims_[1];
    strides_[2] = MK ? s : 0;
    s *= dims_[2];
    size_ = static_cast<std::size_t>(s);
  }
  template <int D>
  uint_t total_length() const {
    return dims_[D];
  }
  template <int D>
  int stride() const {
    return strides_[D];
  }
  const std::array<int, 3>& strides() const { return strides_; }
  int index(int i, int j, int k) const {
    return i * strides_[0] + j * strides_[1] + k * strides_[2];
  }
  std::size_t padded_total_length() const { return size_; }
};

} // namespace shim_detail

template <class T, class Info>
class data_store {
  std::shared_ptr<Info> info_;
  std::shared_ptr<std::vector<T>> data_;
  std::string name_;

public:
  using data_t = T;
  using storage_info_t = Info;
  data_store() = default;
  explicit data_store(const Info& info, const std::string& name = "")
      : info_(std::make_shared<Info>(info)),
        data_(std::make_shared<std::vector<T>>(info.padded_total_length(), T())), name_(name) {}
  data_store(const Info& info, T init, const std::string& name = "")
      : info_(std::make_shared<Info>(info)),
        data_(std::make_shared<std::vector<T>>(info.padded_total_length(), init)), name_(name) {}
  void sync() const {}
  const std::string& name() const { return name_; }
  const Info* get_storage_info_ptr() const { return info_.get(); }
  const std::array<int, 3>& strides() const { return info_->strides(); }
  T* raw() const { return data_->data(); }
  bool valid() const { return static_cast<bool>(data_); }
};

template <class Store, access_mode Mode = access_mode::read_write>
class data_view {
  typename Store::data_t* ptr_ = nullptr;
  typename Store::storage_info_t info_;

public:
  using data_t = typename Store::data_t;
  data_view() = default;
  explicit data_view(const Store& s) : ptr_(s.raw()), info_(*s.get_storage_info_ptr()) {}
  // allow read_only -> read_write style conversions used by generated code
  template <access_mode M2>
  data_view(const data_view<Store, M2>& o) : ptr_(o.data()), info_(o.info()) {}
  data_t& operator()(int i, int j, int k) const { return ptr_[info_.index(i, j, k)]; }
  data_t* data() const { return ptr_; }
  const typename Store::storage_info_t& info() const { return info_; }
};

template <access_mode Mode = access_mode::read_write, class Store>
data_view<Store, Mode> make_host_view(const Store& s) {
  return data_view<Store, Mode>(s);
}

template <class Backend>
struct storage_traits {
  template <int Id, int Dims, class Halo = halo<0, 0, 0>>
  using storage_info_t = shim_detail::storage_info_impl<true, true, true>;

  template <int Id, class Selector, class Halo = halo<0, 0, 0>>
  using special_storage_info_t =
      shim_detail::storage_info_impl<Selector::value[0], Selector::value[1], Selector::value[2]>;

  template <class T, class Info>
  using data_store_t = data_store<T, Info>;
};

} // namespace gridtools
