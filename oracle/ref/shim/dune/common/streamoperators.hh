// Stand-in for <dune/common/streamoperators.hh>: operator<< for std::array.
#ifndef SHIM_DUNE_COMMON_STREAMOPERATORS_HH
#define SHIM_DUNE_COMMON_STREAMOPERATORS_HH
#include <array>
#include <ostream>
namespace Dune {
  template<class T, std::size_t N>
  inline std::ostream& operator<<(std::ostream& s, const std::array<T,N>& a) {
    s << "[";
    for (std::size_t i = 0; i < N; ++i) s << (i ? "," : "") << a[i];
    return s << "]";
  }
}
#endif
