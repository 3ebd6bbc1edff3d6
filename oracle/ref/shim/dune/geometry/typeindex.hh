// Stand-in for <dune/geometry/typeindex.hh>: a dense index over (dim, cube) suffices because the
// only use on the path is as an array slot inside MultipleCodimMultipleGeomTypeMapper.
#ifndef SHIM_DUNE_GEOMETRY_TYPEINDEX_HH
#define SHIM_DUNE_GEOMETRY_TYPEINDEX_HH
#include "type.hh"
namespace Dune {
  struct GlobalGeometryTypeIndex {
    static constexpr std::size_t size(std::size_t maxdim) { return 2*(maxdim+1); }
    static constexpr std::size_t index(const GeometryType& gt) { return 2*gt.dim() + (gt.isNone() ? 1 : 0); }
  };
}
#endif
