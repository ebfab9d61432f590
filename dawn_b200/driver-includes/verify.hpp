// Host-side verifier for drivers: fill / fillMath / verify with the reference's semantics
// (driver-includes/verify.hpp:28-64 fillMath closed form, :87-157 + :207-214 comparison:
// absolute below 1e-3, relative otherwise; default precision 1e-10 for double).
#pragma once
#include "gridtools_includes.hpp"

#include <cmath>
#include <cstdio>

namespace gridtools {
namespace dawn {

class verifier {
  domain dom_;
  double precision_;

  template <class Storage, class Fn>
  static void each_point(Storage& s, Fn&& fn) {
    auto const& info = *s.get_storage_info_ptr();
    auto view = make_host_view(s);
    const uint_t d0 = info.template total_length<0>(), d1 = info.template total_length<1>(),
                 d2 = info.template total_length<2>();
    for(uint_t k = 0; k < d2; ++k)
      for(uint_t j = 0; j < d1; ++j)
        for(uint_t i = 0; i < d0; ++i)
          view(i, j, k) = fn(d0, d1, i, j, k);
  }

public:
  // default precision as the reference's (verify.hpp:30-32): 1e-10 for double, 1e-4 for float
  explicit verifier(const domain& dom, double precision = sizeof(::dawn::float_type) == 8 ? 1e-10 : 1e-4)
      : dom_(dom), precision_(precision) {}

  template <class V, class... Storages>
  void fill(V value, Storages&... storages) const {
    (each_point(storages, [&](uint_t, uint_t, uint_t, uint_t, uint_t) { return value; }), ...);
  }

  template <class... Storages>
  void fillMath(double a, double b, double c, double d, double e, double f,
                Storages&... storages) const {
    const double pi = std::atan(1.) * 4.;
    auto fn = [&](uint_t d0, uint_t d1, uint_t i, uint_t j, uint_t k) {
      const double x = i / (double)d0, y = j / (double)d1;
      return k * 10e-3 + a * (b + std::cos(pi * (x + c * y)) + std::sin(d * pi * (x + e * y))) / f;
    };
    (each_point(storages, fn), ...);
  }

  template <class S1, class S2>
  bool verify(S1& s1, S2& s2, int max_errors = 10) const {
    s1.sync();
    s2.sync();
    auto v1 = make_host_view<access_mode::read_only>(s1);
    auto v2 = make_host_view<access_mode::read_only>(s2);
    bool ok = true;
    for(int k = dom_.kminus(); k < (int)(dom_.ksize() - dom_.kplus()); ++k)
      for(int j = dom_.jminus(); j < (int)(dom_.jsize() - dom_.jplus()); ++j)
        for(int i = dom_.iminus(); i < (int)(dom_.isize() - dom_.iplus()); ++i) {
          const double x = v1(i, j, k), y = v2(i, j, k);
          bool close;
          if(std::fabs(x) < 1e-3 && std::fabs(y) < 1e-3)
            close = std::fabs(x - y) < precision_;
          else
            close = std::fabs((x - y) / x) < precision_;
          if(!close) {
            ok = false;
            if(--max_errors >= 0)
              std::fprintf(stderr, "( %d, %d, %d ) : %s = %.17g ; %s = %.17g\n", i, j, k,
                           s1.name().c_str(), x, s2.name().c_str(), y);
          }
        }
    return ok;
  }
};

} // namespace dawn
} // namespace gridtools
