// Math functions callable from generated kernels.  The IIR names them
// `gridtools::dawn::math::<fn>` (reference: driver-includes/math.hpp, 20 functions); the emitter
// maps them onto ::dawn_b200::math::<fn>, usable on host and device.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define DAWN_B200_FN __host__ __device__ __forceinline__
#else
#define DAWN_B200_FN inline
#endif

namespace dawn_b200 {
namespace math {
template <class T> DAWN_B200_FN T fabs(T x) { return ::fabs(x); }
template <class T> DAWN_B200_FN T abs(T x) { return x < T(0) ? -x : x; }
template <class T> DAWN_B200_FN T floor(T x) { return ::floor(x); }
template <class T> DAWN_B200_FN T ceil(T x) { return ::ceil(x); }
template <class T> DAWN_B200_FN T trunc(T x) { return ::trunc(x); }
template <class T> DAWN_B200_FN T sqrt(T x) { return ::sqrt(x); }
template <class T> DAWN_B200_FN T exp(T x) { return ::exp(x); }
template <class T> DAWN_B200_FN T log(T x) { return ::log(x); }
template <class T> DAWN_B200_FN T sin(T x) { return ::sin(x); }
template <class T> DAWN_B200_FN T cos(T x) { return ::cos(x); }
template <class T> DAWN_B200_FN T tan(T x) { return ::tan(x); }
template <class T> DAWN_B200_FN T asin(T x) { return ::asin(x); }
template <class T> DAWN_B200_FN T acos(T x) { return ::acos(x); }
template <class T> DAWN_B200_FN T atan(T x) { return ::atan(x); }
template <class T, class U> DAWN_B200_FN auto pow(T x, U y) { return ::pow(x, y); }
template <class T, class U> DAWN_B200_FN auto fmod(T x, U y) { return ::fmod(x, y); }
template <class T, class U> DAWN_B200_FN auto min(T x, U y) -> decltype(x + y) { return x < y ? x : y; }
template <class T, class U> DAWN_B200_FN auto max(T x, U y) -> decltype(x + y) { return x > y ? x : y; }
template <class T> DAWN_B200_FN bool isnan(T x) { return x != x; }
template <class T> DAWN_B200_FN bool isinf(T x) { return !(x != x) && ((x - x) != (x - x)); }
} // namespace math
} // namespace dawn_b200

// host-side spelling used by reference-style drivers
namespace gridtools {
namespace dawn {
namespace math = ::dawn_b200::math;
}
} // namespace gridtools
