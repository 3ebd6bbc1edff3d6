// Stand-in for <dune/geometry/referenceelements.hh>: cube reference elements, only the queries
// used by the mapper (type of a sub-entity) and by the oracle driver.
#ifndef SHIM_DUNE_GEOMETRY_REFERENCEELEMENTS_HH
#define SHIM_DUNE_GEOMETRY_REFERENCEELEMENTS_HH
#include "type.hh"
namespace Dune {
  template<class ct, int dim>
  struct CubeReferenceElement {
    GeometryType type(int /*i*/, int codim) const { return GeometryTypes::cube(dim-codim); }
    GeometryType type() const { return GeometryTypes::cube(dim); }
    int size(int codim) const {            // C(dim,codim) * 2^codim
      int b = 1; for (int k = 0; k < codim; ++k) b = b*(dim-k)/(k+1);
      return b << codim;
    }
  };
  template<class ct, int dim>
  struct ReferenceElements {
    static CubeReferenceElement<ct,dim> general(const GeometryType&) { return {}; }
    static CubeReferenceElement<ct,dim> cube() { return {}; }
  };
}
#endif
