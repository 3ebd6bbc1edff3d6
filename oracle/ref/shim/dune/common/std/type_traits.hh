// Stand-in for <dune/common/std/type_traits.hh>: the detection idiom used by
// dune/grid/geometrygrid/coordfunction.hh (Std::is_detected). Test infrastructure only.
#ifndef SHIM_DUNE_COMMON_STD_TYPE_TRAITS_HH
#define SHIM_DUNE_COMMON_STD_TYPE_TRAITS_HH
#include <type_traits>
namespace Dune { namespace Std {
  namespace Impl {
    template<class AlwaysVoid, template<class...> class Op, class... Args> struct detector : std::false_type {};
    template<template<class...> class Op, class... Args> struct detector<std::void_t<Op<Args...>>, Op, Args...> : std::true_type {};
  }
  template<template<class...> class Op, class... Args> using is_detected = Impl::detector<void, Op, Args...>;
} }
#endif
