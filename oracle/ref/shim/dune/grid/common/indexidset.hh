// Minimal stand-in for the IndexSet/IdSet facades of <dune/grid/common/indexidset.hh>
// (reference :113-248, :466-481): pure CRTP forwarding, no arithmetic.
#ifndef SHIM_DUNE_GRID_COMMON_INDEXIDSET_HH
#define SHIM_DUNE_GRID_COMMON_INDEXIDSET_HH
#include <cstddef>
#include <dune/geometry/type.hh>
namespace Dune {
  template<class GridImp, class IndexSetImp, class IndexTypeImp, class TypesImp>
  class IndexSet {
  public:
    typedef IndexTypeImp IndexType;
    typedef TypesImp Types;
    static const int dimension = std::remove_const<GridImp>::type::dimension;
    template<int cc, class E> IndexType index(const E& e) const { return asImp().template index<cc>(e); }
    template<class E> IndexType index(const E& e) const { return asImp().template index<E::codimension>(e); }
    template<int cc, class E> IndexType subIndex(const E& e, int i, unsigned int codim) const { return asImp().template subIndex<cc>(e,i,codim); }
    template<class E> IndexType subIndex(const E& e, int i, unsigned int codim) const { return asImp().template subIndex<E::codimension>(e,i,codim); }
    Types types(int codim) const { return asImp().types(codim); }
    auto size(GeometryType type) const { return asImp().size(type); }
    auto size(int codim) const { return asImp().size(codim); }
    template<class E> bool contains(const E& e) const { return asImp().contains(e); }
  protected:
    IndexSet() = default;
    IndexSetImp& asImp() { return static_cast<IndexSetImp&>(*this); }
    const IndexSetImp& asImp() const { return static_cast<const IndexSetImp&>(*this); }
  };
  template<class GridImp, class IdSetImp, class IdTypeImp>
  class IdSet {
  public:
    typedef IdTypeImp IdType;
    template<class E> IdType id(const E& e) const { return asImp().template id<E::codimension>(e); }
    template<class E0> IdType subId(const E0& e, int i, unsigned int codim) const { return asImp().subId(e,i,codim); }
  protected:
    IdSet() = default;
    const IdSetImp& asImp() const { return static_cast<const IdSetImp&>(*this); }
  };
}
#endif
