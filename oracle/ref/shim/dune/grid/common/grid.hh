// Minimal stand-in for the interface ("facade") layer of <dune/grid/common/grid.hh> and the
// headers it pulls in (entity.hh, geometry.hh, intersection.hh, intersectioniterator.hh,
// entityiterator.hh, entityseed.hh, gridview.hh, defaultgridview.hh). The reference facades only
// forward to the implementation classes; this file reproduces that forwarding
// so that dune/grid/yaspgrid.hh and dune/grid/yaspgrid/*.hh compile UNMODIFIED from
// /root/reference for oracle/_ref. Test infrastructure only; no arithmetic lives here.
#ifndef SHIM_DUNE_GRID_COMMON_GRID_HH
#define SHIM_DUNE_GRID_COMMON_GRID_HH
#include <iterator>
#include <cstddef>
#include <array>
#include <cstddef>
#include <type_traits>
#include <vector>
#include <dune/common/fvector.hh>
#include <dune/geometry/type.hh>
#include <dune/geometry/referenceelements.hh>
#include <dune/grid/common/gridenums.hh>
#include <dune/grid/common/exceptions.hh>
#include <dune/grid/common/datahandleif.hh>
#include <dune/grid/common/indexidset.hh>

namespace Dune {

  // ---- Geometry facade (reference common/geometry.hh:194-371) ---------------------------------
  template<int mydim, int cdim, class GridImp, template<int,int,class> class GeometryImp>
  class Geometry {
  public:
    typedef GeometryImp<mydim,cdim,GridImp> Implementation;
    typedef typename std::remove_const<GridImp>::type::ctype ctype;
    constexpr static int mydimension = mydim;
    constexpr static int coorddimension = cdim;
    typedef FieldVector<ctype,mydim> LocalCoordinate;
    typedef FieldVector<ctype,cdim> GlobalCoordinate;
    explicit Geometry(const Implementation& impl) : impl_(impl) {}
    Implementation& impl() { return impl_; }
    const Implementation& impl() const { return impl_; }
    GeometryType type() const { return impl_.type(); }
    bool affine() const { return impl_.affine(); }
    int corners() const { return impl_.corners(); }
    GlobalCoordinate corner(int i) const { return impl_.corner(i); }
    GlobalCoordinate global(const LocalCoordinate& x) const { return impl_.global(x); }
    LocalCoordinate local(const GlobalCoordinate& y) const { return impl_.local(y); }
    ctype integrationElement(const LocalCoordinate& x) const { return impl_.integrationElement(x); }
    ctype volume() const { return impl_.volume(); }
    GlobalCoordinate center() const { return impl_.center(); }
    auto jacobianTransposed(const LocalCoordinate& x) const { return impl_.jacobianTransposed(x); }
    auto jacobianInverseTransposed(const LocalCoordinate& x) const { return impl_.jacobianInverseTransposed(x); }
  private:
    Implementation impl_;
  };

  // ---- EntitySeed facade ----------------------------------------------------------------------
  template<class GridImp, class EntitySeedImp>
  class EntitySeed {
  public:
    typedef EntitySeedImp Implementation;
    static const int codimension = EntitySeedImp::codimension;
    EntitySeed() {}
    EntitySeed(const Implementation& impl) : impl_(impl) {}
    bool isValid() const { return impl_.isValid(); }
    Implementation& impl() { return impl_; }
    const Implementation& impl() const { return impl_; }
  private:
    Implementation impl_;
  };

  // ---- Entity facade (reference common/entity.hh) ---------------------------------------------
  template<int cd, int dim, class GridImp, template<int,int,class> class EntityImp>
  class Entity {
  public:
    typedef EntityImp<cd,dim,GridImp> Implementation;
    constexpr static int codimension = cd;
    constexpr static int dimension = dim;
    constexpr static int mydimension = dim-cd;
    typedef typename std::remove_const<GridImp>::type::template Codim<cd>::Geometry Geometry;
    typedef typename std::remove_const<GridImp>::type::template Codim<cd>::EntitySeed EntitySeed;
    Entity() {}
    Entity(const Implementation& impl) : impl_(impl) {}
    Entity(Implementation&& impl) : impl_(std::move(impl)) {}
    Implementation& impl() { return impl_; }
    const Implementation& impl() const { return impl_; }
    int level() const { return impl_.level(); }
    PartitionType partitionType() const { return impl_.partitionType(); }
    Geometry geometry() const { return impl_.geometry(); }
    GeometryType type() const { return impl_.type(); }
    unsigned int subEntities(unsigned int codim) const { return impl_.subEntities(codim); }
    EntitySeed seed() const { return impl_.seed(); }
    bool operator==(const Entity& o) const { return impl_.equals(o.impl_); }
    bool operator!=(const Entity& o) const { return !impl_.equals(o.impl_); }
    // codim-0 only (instantiated lazily)
    template<int cc> auto subEntity(int i) const { return impl_.template subEntity<cc>(i); }
    auto father() const { return impl_.father(); }
    bool hasFather() const { return impl_.hasFather(); }
    bool isLeaf() const { return impl_.isLeaf(); }
    auto geometryInFather() const { return impl_.geometryInFather(); }
    auto ileafbegin() const { return impl_.ileafbegin(); }
    auto ileafend() const { return impl_.ileafend(); }
    auto ilevelbegin() const { return impl_.ilevelbegin(); }
    auto ilevelend() const { return impl_.ilevelend(); }
    auto hbegin(int maxlevel) const { return impl_.hbegin(maxlevel); }
    auto hend(int maxlevel) const { return impl_.hend(maxlevel); }
    // reference common/entity.hh:656-665 (default via intersections) is not needed by the oracle
  private:
    Implementation impl_;
  };

  template<int cd, int dim, class GridImp, template<int,int,class> class EntityImp>
  class EntityDefaultImplementation {
  public:
    constexpr static int codimension = cd;
    constexpr static int dimension = dim;
    constexpr static int mydimension = dim-cd;
  };

  // ---- EntityIterator facade (reference common/entityiterator.hh) -----------------------------
  template<int codim, class GridImp, class IteratorImp>
  class EntityIterator {
  public:
    typedef IteratorImp Implementation;
    typedef typename IteratorImp::Entity Entity;
    // std::iterator_traits members (reference common/entityiterator.hh:26-47); std::distance in utility/globalindexset.hh:340
    typedef std::forward_iterator_tag iterator_category;
    typedef Entity value_type;
    typedef std::ptrdiff_t difference_type;
    typedef const Entity* pointer;
    typedef const Entity& reference;
    EntityIterator() {}
    EntityIterator(const IteratorImp& imp) : imp_(imp) {}
    EntityIterator& operator++() { imp_.increment(); return *this; }
    const Entity& operator*() const { return imp_.dereference(); }
    const Entity* operator->() const { return &imp_.dereference(); }
    bool operator==(const EntityIterator& o) const { return imp_.equals(o.imp_); }
    bool operator!=(const EntityIterator& o) const { return !imp_.equals(o.imp_); }
    Implementation& impl() { return imp_; }
    const Implementation& impl() const { return imp_; }
  private:
    IteratorImp imp_;
  };

  // ---- Intersection facades (reference common/intersection.hh:216-411) ------------------------
  template<class GridImp, class IntersectionImp>
  class Intersection {
  public:
    typedef IntersectionImp Implementation;
    Intersection() {}
    Intersection(const Implementation& impl) : impl_(impl) {}
    Implementation& impl() { return impl_; }
    const Implementation& impl() const { return impl_; }
    bool boundary() const { return impl_.boundary(); }
    bool neighbor() const { return impl_.neighbor(); }
    bool conforming() const { return impl_.conforming(); }
    auto inside() const { return impl_.inside(); }
    auto outside() const { return impl_.outside(); }
    int boundarySegmentIndex() const { return impl_.boundarySegmentIndex(); }
    auto geometry() const { return impl_.geometry(); }
    auto geometryInInside() const { return impl_.geometryInInside(); }
    auto geometryInOutside() const { return impl_.geometryInOutside(); }
    GeometryType type() const { return impl_.type(); }
    int indexInInside() const { return impl_.indexInInside(); }
    int indexInOutside() const { return impl_.indexInOutside(); }
    auto centerUnitOuterNormal() const { return impl_.centerUnitOuterNormal(); }
    template<class X> auto outerNormal(const X& x) const { return impl_.outerNormal(x); }
    template<class X> auto unitOuterNormal(const X& x) const { return impl_.unitOuterNormal(x); }
    template<class X> auto integrationOuterNormal(const X& x) const { return impl_.integrationOuterNormal(x); }
    bool operator==(const Intersection& o) const { return impl_.equals(o.impl_); }
    bool operator!=(const Intersection& o) const { return !impl_.equals(o.impl_); }
  private:
    Implementation impl_;
  };

  template<class GridImp, class IntersectionIteratorImp, class IntersectionImp>
  class IntersectionIterator {
  public:
    typedef IntersectionIteratorImp Implementation;
    typedef Dune::Intersection<GridImp,IntersectionImp> Intersection;
    IntersectionIterator() {}
    IntersectionIterator(const Implementation& imp) : imp_(imp) {}
    IntersectionIterator& operator++() { imp_.increment(); return *this; }
    const Intersection& operator*() const { return imp_.dereference(); }
    const Intersection* operator->() const { return &imp_.dereference(); }
    bool operator==(const IntersectionIterator& o) const { return imp_.equals(o.imp_); }
    bool operator!=(const IntersectionIterator& o) const { return !imp_.equals(o.imp_); }
  private:
    Implementation imp_;
  };

  // ---- Grid views (reference common/gridview.hh, common/defaultgridview.hh) -------------------
  template<class GridImp> struct DefaultLevelGridViewTraits { typedef GridImp Grid; static const bool leaf = false; };
  template<class GridImp> struct DefaultLeafGridViewTraits  { typedef GridImp Grid; static const bool leaf = true;  };

  template<class ViewTraits>
  class GridView {
  public:
    typedef typename std::remove_const<typename ViewTraits::Grid>::type Grid;
    typedef GridView Traits;          // GridView::Traits::Codim<cd>::Iterator (used by utility/globalindexset.hh:102)
    typedef typename Grid::Traits GTraits;
    typedef typename std::conditional<ViewTraits::leaf, typename GTraits::LeafIndexSet, typename GTraits::LevelIndexSet>::type IndexSet;
    typedef typename GTraits::LeafIntersectionIterator IntersectionIterator;
    typedef typename GTraits::LeafIntersection Intersection;
    typedef typename GTraits::Communication Communication;
    typedef typename Grid::ctype ctype;
    constexpr static int dimension = Grid::dimension;
    constexpr static int dimensionworld = Grid::dimensionworld;
    template<int cd> struct Codim {
      typedef typename GTraits::template Codim<cd>::Entity Entity;
      typedef typename GTraits::template Codim<cd>::Geometry Geometry;
      typedef typename GTraits::template Codim<cd>::template Partition<All_Partition>::LevelIterator Iterator;
      template<PartitionIteratorType pit> struct Partition {
        typedef typename GTraits::template Codim<cd>::template Partition<pit>::LevelIterator Iterator;
      };
    };
    GridView(const Grid& g, int level) : grid_(&g), level_(level) {}
    const Grid& grid() const { return *grid_; }
    int level() const { return level_; }
    const IndexSet& indexSet() const {
      if constexpr (ViewTraits::leaf) return grid_->leafIndexSet(); else return grid_->levelIndexSet(level_);
    }
    int size(int codim) const { return grid_->size(level_, codim); }
    int size(const GeometryType& t) const { return grid_->size(level_, t); }
    template<int cd> auto begin() const { return grid_->template lbegin<cd,All_Partition>(level_); }
    template<int cd> auto end() const { return grid_->template lend<cd,All_Partition>(level_); }
    template<int cd, PartitionIteratorType pit> auto begin() const { return grid_->template lbegin<cd,pit>(level_); }
    template<int cd, PartitionIteratorType pit> auto end() const { return grid_->template lend<cd,pit>(level_); }
    auto ibegin(const typename Codim<0>::Entity& e) const { return e.impl().ilevelbegin(); }
    auto iend(const typename Codim<0>::Entity& e) const { return e.impl().ilevelend(); }
    const Communication& comm() const { return grid_->comm(); }
    int overlapSize(int codim) const { return grid_->overlapSize(level_, codim); }
    int ghostSize(int codim) const { return grid_->ghostSize(level_, codim); }
    template<class DataHandleImp, class DataType>
    void communicate(CommDataHandleIF<DataHandleImp,DataType>& data, InterfaceType iftype, CommunicationDirection dir) const {
      grid_->communicate(data, iftype, dir, level_);
    }
  private:
    const Grid* grid_;
    int level_;
  };

  // ---- GridTraits (reference common/grid.hh:1009-1108) ----------------------------------------
  template<int dim, int dimw, class GridImp,
      template<int,int,class> class GeometryImp,
      template<int,int,class> class EntityImp,
      template<int,PartitionIteratorType,class> class LevelIteratorImp,
      template<class> class LeafIntersectionImp,
      template<class> class LevelIntersectionImp,
      template<class> class LeafIntersectionIteratorImp,
      template<class> class LevelIntersectionIteratorImp,
      template<class> class HierarchicIteratorImp,
      template<int,PartitionIteratorType,class> class LeafIteratorImp,
      class LevelIndexSetImp, class LeafIndexSetImp,
      class GlobalIdSetImp, class GIDType, class LocalIdSetImp, class LIDType, class CCType,
      template<class> class LevelGridViewTraits,
      template<class> class LeafGridViewTraits,
      template<int,class> class EntitySeedImp,
      template<int,int,class> class LocalGeometryImp = GeometryImp,
      class LevelIndexType = unsigned int,
      class LevelGeometryTypes = std::vector<GeometryType>,
      class LeafIndexType = LevelIndexType,
      class LeafGeometryTypes = LevelGeometryTypes>
  struct GridTraits {
    typedef GridImp Grid;
    typedef Dune::Intersection<const GridImp, LeafIntersectionImp<const GridImp>> LeafIntersection;
    typedef Dune::Intersection<const GridImp, LevelIntersectionImp<const GridImp>> LevelIntersection;
    typedef Dune::IntersectionIterator<const GridImp, LeafIntersectionIteratorImp<const GridImp>, LeafIntersectionImp<const GridImp>> LeafIntersectionIterator;
    typedef Dune::IntersectionIterator<const GridImp, LevelIntersectionIteratorImp<const GridImp>, LevelIntersectionImp<const GridImp>> LevelIntersectionIterator;
    typedef Dune::EntityIterator<0, const GridImp, HierarchicIteratorImp<const GridImp>> HierarchicIterator;
    template<int cd>
    struct Codim {
      typedef GeometryImp<dim-cd, dimw, const GridImp> GeometryImpl;
      typedef LocalGeometryImp<dim-cd, dim, const GridImp> LocalGeometryImpl;
      typedef Dune::Geometry<dim-cd, dimw, const GridImp, GeometryImp> Geometry;
      typedef Dune::Geometry<dim-cd, dim, const GridImp, LocalGeometryImp> LocalGeometry;
      typedef Dune::Entity<cd, dim, const GridImp, EntityImp> Entity;
      typedef Dune::EntitySeed<const GridImp, EntitySeedImp<cd, const GridImp>> EntitySeed;
      template<PartitionIteratorType pitype>
      struct Partition {
        typedef Dune::EntityIterator<cd, const GridImp, LevelIteratorImp<cd,pitype,const GridImp>> LevelIterator;
        typedef Dune::EntityIterator<cd, const GridImp, LeafIteratorImp<cd,pitype,const GridImp>> LeafIterator;
      };
      typedef typename Partition<All_Partition>::LeafIterator LeafIterator;
      typedef typename Partition<All_Partition>::LevelIterator LevelIterator;
    };
    typedef Dune::GridView<LeafGridViewTraits<const GridImp>> LeafGridView;
    typedef Dune::GridView<LevelGridViewTraits<const GridImp>> LevelGridView;
    typedef IndexSet<const GridImp,LevelIndexSetImp,LevelIndexType,LevelGeometryTypes> LevelIndexSet;
    typedef IndexSet<const GridImp,LeafIndexSetImp,LeafIndexType,LeafGeometryTypes> LeafIndexSet;
    typedef IdSet<const GridImp,GlobalIdSetImp,GIDType> GlobalIdSet;
    typedef IdSet<const GridImp,LocalIdSetImp,LIDType> LocalIdSet;
    typedef CCType Communication;
  };

  // ---- Grid base (reference common/grid.hh Grid / GridDefaultImplementation) ------------------
  template<int dim, int dimworld, class ct, class GridFamily>
  class GridDefaultImplementation {
  public:
    typedef typename GridFamily::Traits Traits;
    constexpr static int dimension = dim;
    constexpr static int dimensionworld = dimworld;
    typedef ct ctype;
    typedef typename Traits::Communication Communication;
    typedef typename Traits::LeafGridView LeafGridView;
    typedef typename Traits::LevelGridView LevelGridView;
    typedef typename Traits::LeafIntersectionIterator LeafIntersectionIterator;
    typedef typename Traits::LevelIntersectionIterator LevelIntersectionIterator;
    typedef typename Traits::LeafIntersection LeafIntersection;
    typedef typename Traits::LevelIntersection LevelIntersection;
    typedef typename Traits::HierarchicIterator HierarchicIterator;
    typedef typename Traits::LevelIndexSet LevelIndexSet;
    typedef typename Traits::LeafIndexSet LeafIndexSet;
    typedef typename Traits::GlobalIdSet GlobalIdSet;
    typedef typename Traits::LocalIdSet LocalIdSet;
    template<int cd> struct Codim : public Traits::template Codim<cd> {};
    // reference common/grid.hh:878-891
    LevelGridView levelGridView(int level) const { return LevelGridView(asImp(), level); }
    LeafGridView leafGridView() const { return LeafGridView(asImp(), asImp().maxLevel()); }
  protected:
    const typename Traits::Grid& asImp() const { return static_cast<const typename Traits::Grid&>(*this); }
  };
}
#endif
