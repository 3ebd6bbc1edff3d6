// Stand-in for dune-geometry's <dune/geometry/axisalignedcubegeometry.hh>.
// dune-geometry is NOT vendored in /root/reference and not installed here, so this restates the
// published semantics of Dune::AxisAlignedCubeGeometry (dune-geometry >= 2.9; see SURVEY.md
// Appendix C.1). "geometry parity unpinned": there is no upstream binary to check this against.
#ifndef SHIM_DUNE_GEOMETRY_AXISALIGNEDCUBEGEOMETRY_HH
#define SHIM_DUNE_GEOMETRY_AXISALIGNEDCUBEGEOMETRY_HH
#include <bitset>
#include <cassert>
#include <dune/common/fvector.hh>
#include <dune/common/fmatrix.hh>
#include "type.hh"
namespace Dune {
  template<class CoordType, unsigned int dim, unsigned int coorddim>
  class AxisAlignedCubeGeometry {
  public:
    constexpr static int mydimension = dim;
    constexpr static int coorddimension = coorddim;
    typedef CoordType ctype;
    typedef FieldVector<ctype,dim> LocalCoordinate;
    typedef FieldVector<ctype,coorddim> GlobalCoordinate;
    typedef ctype Volume;
    typedef FieldMatrix<ctype,dim,coorddim> JacobianTransposed;
    typedef FieldMatrix<ctype,coorddim,dim> JacobianInverseTransposed;

    AxisAlignedCubeGeometry(const GlobalCoordinate lower, const GlobalCoordinate upper)
      : lower_(lower), upper_(upper), axes_((1ull<<coorddim)-1) {}
    AxisAlignedCubeGeometry(const GlobalCoordinate lower, const GlobalCoordinate upper, const std::bitset<coorddim>& axes)
      : lower_(lower), upper_(upper), axes_(axes)
    {
      for (std::size_t i=0; i<coorddim; i++) if (!axes_[i]) upper_[i] = lower_[i];
    }
    AxisAlignedCubeGeometry(const GlobalCoordinate lower) : lower_(lower), upper_(lower), axes_(0) {}

    GeometryType type() const { return GeometryTypes::cube(dim); }
    bool affine() const { return true; }
    int corners() const { return 1<<dim; }

    GlobalCoordinate global(const LocalCoordinate& local) const {
      GlobalCoordinate r;
      std::size_t lc = 0;
      for (std::size_t i=0; i<coorddim; i++)
        r[i] = axes_[i] ? lower_[i] + local[lc++]*(upper_[i]-lower_[i]) : lower_[i];
      return r;
    }
    LocalCoordinate local(const GlobalCoordinate& global) const {
      LocalCoordinate r;
      std::size_t lc = 0;
      for (std::size_t i=0; i<coorddim; i++)
        if (axes_[i]) r[lc++] = (global[i]-lower_[i])/(upper_[i]-lower_[i]);
      return r;
    }
    JacobianTransposed jacobianTransposed(const LocalCoordinate&) const {
      JacobianTransposed r;
      std::size_t lc = 0;
      for (std::size_t i=0; i<coorddim; i++) if (axes_[i]) r[lc++][i] = upper_[i]-lower_[i];
      return r;
    }
    JacobianInverseTransposed jacobianInverseTransposed(const LocalCoordinate&) const {
      JacobianInverseTransposed r;
      std::size_t lc = 0;
      for (std::size_t i=0; i<coorddim; i++) if (axes_[i]) r[i][lc++] = CoordType(1.0)/(upper_[i]-lower_[i]);
      return r;
    }
    Volume integrationElement(const LocalCoordinate&) const { return volume(); }
    GlobalCoordinate center() const {
      GlobalCoordinate r;
      if (dim == 0) return lower_;
      for (std::size_t i=0; i<coorddim; i++) r[i] = CoordType(0.5)*(lower_[i]+upper_[i]);
      return r;
    }
    GlobalCoordinate corner(int k) const {
      GlobalCoordinate r;
      unsigned int mask = 1;
      for (std::size_t i=0; i<coorddim; i++) {
        if (!axes_[i]) r[i] = lower_[i];
        else { r[i] = (k & mask) ? upper_[i] : lower_[i]; mask <<= 1; }
      }
      return r;
    }
    Volume volume() const {
      ctype vol = 1;
      for (std::size_t i=0; i<coorddim; i++) if (axes_[i]) vol *= upper_[i]-lower_[i];
      return vol;
    }
    const GlobalCoordinate& lower() const { return lower_; }
    const GlobalCoordinate& upper() const { return upper_; }
  private:
    GlobalCoordinate lower_, upper_;
    std::bitset<coorddim> axes_;
  };
}
#endif
