// Stand-in for <dune/geometry/referenceelements.hh>: cube reference elements, only the queries used by the mapper
// (type / number of sub-entities), by dune/grid/geometrygrid/{cornerstorage,intersection}.hh (face normals, sub-entity
// corners, barycentres) and by the oracle driver. dune-geometry is not vendored: RESTATED (numbering of the reference
// cube: face k has direction k/2 and side k%2, corner j has bit i set where it sits at 1 along axis i; sub-entity
// corners in ascending corner number). Test infrastructure only.
#ifndef SHIM_DUNE_GEOMETRY_REFERENCEELEMENTS_HH
#define SHIM_DUNE_GEOMETRY_REFERENCEELEMENTS_HH
#include <dune/common/fvector.hh>
#include "type.hh"
namespace Dune {
  template<class ct, int dim>
  struct CubeReferenceElement {
    typedef FieldVector<ct,dim> Coordinate;
    struct SubGeometry { ct volume() const { return ct(1); } };
    GeometryType type(int /*i*/, int codim) const { return GeometryTypes::cube(dim-codim); }
    GeometryType type() const { return GeometryTypes::cube(dim); }
    int size(int codim) const {            // C(dim,codim) * 2^codim
      int b = 1; for (int k = 0; k < codim; ++k) b = b*(dim-k)/(k+1);
      return b << codim;
    }
    ct volume() const { return ct(1); }
    // faces (codim 1) and the element itself (codim 0) only: what GeoGrid::Intersection asks for
    int size(int /*i*/, int c, int cc) const {
      if (c == 0) return size(cc);
      if (c == 1 && cc == dim) return 1 << (dim-1);
      return cc == c ? 1 : -1;
    }
    int subEntity(int i, int c, int ii, int cc) const {
      if (c == 0) return ii;
      if (c == 1 && cc == dim) {           // ii-th corner of face i: bits of ii spread over the axes != dir, side at dir
        const int dir = i/2, side = i%2;
        int v = 0, b = 0;
        for (int ax = 0; ax < dim; ++ax) { if (ax == dir) { v |= side << ax; continue; } v |= ((ii >> b) & 1) << ax; ++b; }
        return v;
      }
      return cc == c ? i : -1;
    }
    Coordinate position(int i, int c) const {       // barycentre of sub-entity (i, c): c = 0, 1 or dim
      Coordinate x(ct(0.5));
      if (c == 1) x[i/2] = ct(i%2);
      if (c == dim) for (int ax = 0; ax < dim; ++ax) x[ax] = ct((i >> ax) & 1);
      return x;
    }
    Coordinate integrationOuterNormal(int face) const { Coordinate n(ct(0)); n[face/2] = (face%2) ? ct(1) : ct(-1); return n; }
    template<int codim> SubGeometry geometry(int /*i*/) const { return SubGeometry(); }
  };
  // dim = 0 (the reference point: barycentre of a 1-D grid's intersections)
  template<class ct>
  struct CubeReferenceElement<ct,0> {
    typedef FieldVector<ct,0> Coordinate;
    GeometryType type() const { return GeometryTypes::cube(0); }
    int size(int) const { return 1; }
    ct volume() const { return ct(1); }
    Coordinate position(int, int) const { return Coordinate(); }
  };
  template<class ct, int dim>
  struct ReferenceElements {
    static CubeReferenceElement<ct,dim> general(const GeometryType&) { return {}; }
    static CubeReferenceElement<ct,dim> cube() { return {}; }
  };
  template<class ct, int dim>
  inline CubeReferenceElement<ct,dim> referenceElement(const GeometryType&) { return {}; }
}
#endif
