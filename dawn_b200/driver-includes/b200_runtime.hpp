// C++ conveniences over the C-ABI runtime (include/dawn_b200_rt.h) used by generated host code.
#pragma once
#include "../../include/dawn_b200_rt.h"
#include "defs.hpp"

#include <stdexcept>
#include <string>

namespace dawn_b200 {

inline int set_error(int code, const char* msg) { return b200rt_set_error(code, msg); }
inline int check_launch() { return b200rt_check_last_launch(); }
[[noreturn]] inline void throw_last_error() {
  throw std::runtime_error(std::string("dawn_b200: ") + b200rt_last_error());
}
inline void check(int err) {
  if(err)
    throw_last_error();
}

// Device scratch for temporaries that cannot be kept in registers; (re)allocated on first use with
// the strides of the ijk fields it is used with.
struct scratch_field {
  ::dawn::float_type* data = nullptr;
  int stride_j = 0, stride_k = 0, levels = 0;
  scratch_field() = default;
  scratch_field(const scratch_field&) = delete;
  ~scratch_field() { b200rt_free(data); }
  int ensure(int sj, int sk, int nlevels) {
    if(data && sj == stride_j && sk == stride_k && nlevels <= levels)
      return 0;
    b200rt_free(data);
    data = nullptr;
    void* p = nullptr;
    size_t bytes = (size_t)sk * (size_t)(nlevels + 1) * sizeof(::dawn::float_type);
    if(int e = b200rt_malloc(&p, bytes))
      return e;
    if(int e = b200rt_memset(p, 0, bytes, nullptr))
      return e;
    data = static_cast<::dawn::float_type*>(p);
    stride_j = sj, stride_k = sk, levels = nlevels;
    return 0;
  }
};

} // namespace dawn_b200
