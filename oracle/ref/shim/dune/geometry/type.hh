// Stand-in for <dune/geometry/type.hh>: cube geometry types only (YaspGrid has no others).
#ifndef SHIM_DUNE_GEOMETRY_TYPE_HH
#define SHIM_DUNE_GEOMETRY_TYPE_HH
#include <ostream>
namespace Dune {
  class GeometryType {
  public:
    constexpr GeometryType() : dim_(0), none_(true) {}
    constexpr GeometryType(unsigned int topologyId, unsigned int dim, bool none = false) : dim_(dim), none_(none) { (void)topologyId; }
    constexpr unsigned int dim() const { return dim_; }
    constexpr bool isCube() const { return !none_; }
    constexpr bool isNone() const { return none_; }
    constexpr bool isVertex() const { return !none_ && dim_==0; }
    constexpr bool isLine() const { return !none_ && dim_==1; }
    // the queries of dune/grid/io/file/vtk/common.hh (cube types only exist here)
    constexpr bool isQuadrilateral() const { return !none_ && dim_==2; }
    constexpr bool isHexahedron() const { return !none_ && dim_==3; }
    constexpr bool isTriangle() const { return false; }
    constexpr bool isTetrahedron() const { return false; }
    constexpr bool isPyramid() const { return false; }
    constexpr bool isPrism() const { return false; }
    constexpr unsigned int id() const { return none_ ? 0u : ((1u << dim_) - 1u); }
    constexpr bool operator==(const GeometryType& o) const { return dim_==o.dim_ && none_==o.none_; }
    constexpr bool operator!=(const GeometryType& o) const { return !(*this==o); }
    constexpr bool operator<(const GeometryType& o) const { return dim_<o.dim_; }
  private:
    unsigned int dim_;
    bool none_;
  };
  inline std::ostream& operator<<(std::ostream& s, const GeometryType& t) { return s << "(cube, " << t.dim() << ")"; }
  namespace GeometryTypes {
    inline constexpr GeometryType cube(unsigned int dim) { return GeometryType((1u<<dim)-1u, dim); }
    inline constexpr GeometryType none(unsigned int dim) { return GeometryType(0u, dim, true); }
    inline constexpr GeometryType vertex = GeometryType(0u, 0u);
  }
}
#endif
