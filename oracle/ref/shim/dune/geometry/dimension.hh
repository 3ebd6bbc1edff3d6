// Stand-in for <dune/geometry/dimension.hh>.
#ifndef SHIM_DUNE_GEOMETRY_DIMENSION_HH
#define SHIM_DUNE_GEOMETRY_DIMENSION_HH
#include <type_traits>
namespace Dune {
  template<int dim> struct Dim : public std::integral_constant<int,dim> {};
  template<int codim> struct Codim : public std::integral_constant<int,codim> {};
}
#endif
