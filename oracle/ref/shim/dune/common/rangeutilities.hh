// Stand-in for <dune/common/rangeutilities.hh>: IntegralRange only.
#ifndef SHIM_DUNE_COMMON_RANGEUTILITIES_HH
#define SHIM_DUNE_COMMON_RANGEUTILITIES_HH
#include <cstddef>
namespace Dune {
  template<class T>
  class IntegralRange {
  public:
    struct iterator {
      T v;
      T operator*() const { return v; }
      iterator& operator++() { ++v; return *this; }
      bool operator!=(const iterator& o) const { return v!=o.v; }
      bool operator==(const iterator& o) const { return v==o.v; }
    };
    IntegralRange(T from, T to) : from_(from), to_(to) {}
    iterator begin() const { return {from_}; }
    iterator end() const { return {to_}; }
    std::size_t size() const { return to_-from_; }
    bool empty() const { return from_==to_; }
  private:
    T from_, to_;
  };
}
#endif
