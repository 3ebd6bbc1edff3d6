// Stand-in for <dune/common/binaryfunctions.hh>.
#ifndef SHIM_DUNE_COMMON_BINARYFUNCTIONS_HH
#define SHIM_DUNE_COMMON_BINARYFUNCTIONS_HH
#include <algorithm>
namespace Dune {
  template<class T> struct Min { T operator()(const T& a, const T& b) const { return std::min(a,b); } };
  template<class T> struct Max { T operator()(const T& a, const T& b) const { return std::max(a,b); } };
}
#endif
