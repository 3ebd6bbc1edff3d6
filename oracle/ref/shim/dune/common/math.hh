// Stand-in for <dune/common/math.hh>: only Dune::power and Dune::binomial are used on the path.
#ifndef SHIM_DUNE_COMMON_MATH_HH
#define SHIM_DUNE_COMMON_MATH_HH
#include <cmath>
namespace Dune {
  template<class Base, class Exponent>
  constexpr Base power(Base m, Exponent p) {
    Base r = 1;
    for (Exponent i = 0; i < p; ++i) r *= m;
    return r;
  }
  template<class T>
  constexpr T binomial(const T& n, const T& k) noexcept {
    if (k < 0 || k > n) return 0;
    T kk = (2*k > n) ? n-k : k;
    T bin = 1;
    for (T i = n-kk; i < n; ++i) bin *= i+1;
    T den = 1;
    for (T i = 1; i <= kk; ++i) den *= i;
    return bin/den;
  }
}
#endif
