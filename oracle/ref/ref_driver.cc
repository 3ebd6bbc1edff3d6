// ref_driver.cc — oracle/_ref/liboracle_ref.so: the REFERENCE's own YaspGrid code driven through
// oracle/oracle_api.h. TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Built by oracle/Makefile from the sources where they lie under /root/reference:
//   dune/grid/yaspgrid.hh, dune/grid/yaspgrid/{coordinates,torus,partitioning,ygrid,yaspgridgeometry,
//   yaspgridentity,yaspgridintersection,yaspgridintersectioniterator,yaspgridleveliterator,
//   yaspgridindexsets,yaspgrididset,yaspgridentityseed,yaspgridhierarchiciterator}.hh,
//   dune/grid/yaspgrid/backuprestore.hh,
//   dune/grid/common/{gridenums,exceptions,datahandleif,mapper,mcmgmapper,capabilities,backuprestore}.hh
// are included unmodified; dune-common, dune-geometry and MPI (absent from this image) are replaced
// by the stand-ins in oracle/ref/shim (ranks = threads of this process, see shim/mpi.h).
// Every loop below is written the way a dune-grid user writes it: per-entity iterators,
// per-intersection facade calls, a CommDataHandleIF data handle.
#define HAVE_MPI 1
#include <dune/grid/yaspgrid.hh>
#include <dune/grid/common/mcmgmapper.hh>
#include <dune/grid/utility/globalindexset.hh>
// GeometryGrid: the reference's own corner and intersection rules (compiled unmodified; they pull in
// geometrygrid/{declaration,coordfunction,coordfunctioncaller,hostcorners}.hh)
// the VTK file syntax: VTUWriter + the four DataArrayWriter families, unmodified
#include <dune/grid/io/file/vtk/vtuwriter.hh>
#include <dune/grid/geometrygrid/cornerstorage.hh>
#include <dune/grid/geometrygrid/intersection.hh>

#include <chrono>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <thread>

#include "../oracle_api.h"
#include "../common/fv_problem.hh"
#include "../common/multilinear.hh"


// ---- GeometryGrid<YaspGrid<dim>, AnalyticalCoordFunction> as far as the path needs it ----------------------------------
// The reference's GeoGrid::CoordVector (corners = F(hostGeometry.corner(i)) through HostCorners + CoordFunctionCaller,
// geometrygrid/cornerstorage.hh:54-60), GeoGrid::IntersectionCoordVector (face corners = insideGeo.global(hostLocalGeo.corner(i)),
// :148-154) and GeoGrid::Intersection (normals, geometrygrid/intersection.hh:103-172) are instantiated UNMODIFIED on a
// small traits class; only the geometry implementation they are given is ours: the multilinear map restated in
// oracle/common/multilinear.hh (dune-geometry's MultiLinearGeometry is not vendored - unpinned arithmetic, pinned rules).
namespace orcgeo {
  template<int dim>
  struct CoordFunction : public Dune::AnalyticalCoordFunction<double, dim, dim, CoordFunction<dim>> {
    int def = 0; const double* params = nullptr;
    CoordFunction(int d, const double* p) : def(d), params(p) {}
    void evaluate(const Dune::FieldVector<double,dim>& x, Dune::FieldVector<double,dim>& y) const { orc::deformation(def, params, dim, x.data(), y.data()); }
  };

  template<int mydim, int cdim> struct JacInvT {          // MultiLinearGeometry::JacobianInverseTransposed: cdim x mydim + detInv()
    double m[cdim][mydim]; double detInv_;
    template<class X, class Y> void mv(const X& x, Y& y) const { for (int i = 0; i < cdim; ++i) { y[i] = 0; for (int j = 0; j < mydim; ++j) y[i] += m[i][j]*x[j]; } }
    double detInv() const { return detInv_; }
    double det() const { return 1.0/detInv_; }
    const double* operator[](int i) const { return m[i]; }
  };

  template<int mydim, int cdim, class Grid>
  class MLGeometry {
  public:
    typedef Dune::FieldVector<double,mydim> LocalCoordinate;
    typedef Dune::FieldVector<double,cdim> GlobalCoordinate;
    typedef JacInvT<mydim,cdim> JacobianInverseTransposed;
    static const int ncorners = 1 << mydim;
    MLGeometry() : grid_(nullptr), valid_(false) {}
    explicit MLGeometry(const Grid& g) : grid_(&g), valid_(false) {}
    template<class CoordVector>
    MLGeometry(const Grid& g, const Dune::GeometryType&, const CoordVector& coords) : grid_(&g), valid_(true) {
      std::array<GlobalCoordinate, ncorners> c;
      coords.calculate(c);                               // the reference's corner rule fills the corners
      for (int k = 0; k < ncorners; ++k) for (int i = 0; i < cdim; ++i) c_[k][i] = c[k][i];
    }
    explicit operator bool() const { return valid_; }
    const Grid& grid() const { return *grid_; }
    Dune::GeometryType type() const { return Dune::GeometryTypes::cube(mydim); }
    int corners() const { return ncorners; }
    GlobalCoordinate corner(int k) const { GlobalCoordinate y; for (int i = 0; i < cdim; ++i) y[i] = c_[k][i]; return y; }
    GlobalCoordinate global(const LocalCoordinate& x) const {
      GlobalCoordinate y; const double (*cit)[cdim] = c_;
      orc::ml_global<cdim>(mydim, cit, x.data(), 1.0, false, y.data());
      return y;
    }
    void jacobianTransposed(const LocalCoordinate& x, double (*jt)[cdim]) const {
      const double (*cit)[cdim] = c_;
      orc::ml_jt<cdim>(mydim, cit, x.data(), 1.0, false, jt);
    }
    JacobianInverseTransposed jacobianInverseTransposed(const LocalCoordinate& x) const {
      static_assert(mydim == cdim, "inverse only for elements");
      double jt[mydim][cdim], inv[cdim][cdim];
      jacobianTransposed(x, jt);
      JacobianInverseTransposed r;
      r.detInv_ = orc::right_inv<cdim>(jt, inv);          // MatrixHelper::rightInvA returns sqrt(det(A A^T))
      for (int i = 0; i < cdim; ++i) for (int j = 0; j < mydim; ++j) r.m[i][j] = inv[i][j];
      return r;
    }
    double integrationElement(const LocalCoordinate& x) const {
      double jt[mydim > 0 ? mydim : 1][cdim];
      jacobianTransposed(x, jt);
      return orc::sqrt_det_aat_rect<mydim,cdim>(jt);
    }
    GlobalCoordinate center() const { return global(Dune::referenceElement<double,mydim>(type()).position(0, 0)); }
    double volume() const { return integrationElement(Dune::referenceElement<double,mydim>(type()).position(0, 0))*1.0; }
    const double (*raw() const)[cdim] { return c_; }
  private:
    const Grid* grid_; bool valid_;
    double c_[ncorners][cdim];
  };

  template<int dim, class HostGridType>
  struct Grid {
    typedef Grid<dim, HostGridType> This;
    struct Traits {
      typedef double ctype;
      static const int dimension = dim, dimensionworld = dim;
      typedef HostGridType HostGrid;
      typedef orcgeo::CoordFunction<dim> CoordFunction;
      template<int cd> struct Codim {
        typedef MLGeometry<dim-cd, dim, const This> GeometryImpl;
        typedef GeometryImpl Geometry;
        typedef typename HostGrid::template Codim<1>::LocalGeometry LocalGeometry;
        typedef int Entity; typedef int EntityImpl;          // inside() / outside() are never instantiated
      };
    };
  };
}

namespace {

thread_local std::string g_err;

struct WorldBase {
  int P = 1;
  int gdim = 0;
  virtual ~WorldBase() {}
  virtual int maxLevel() = 0;
  virtual void torusDims(int*) = 0;
  virtual void domainSize(double*) = 0;
  virtual int64_t size(int rank, int level, int codim) = 0;
  virtual int nbs(int rank) = 0;
  virtual int overlapSize(int level) = 0;
  virtual int boxes(int rank, int level, int codim, int* out) = 0;
  virtual int list(int rank, int level, int codim, int which, int* out, int cap) = 0;
  virtual int coordinates(int rank, int level, int axis, double* out, int* first) = 0;
  virtual int64_t entities(int rank, int level, int codim, int pit, uint32_t*, uint8_t*, int32_t*, double*, double*, double*, double*) = 0;
  virtual int64_t subindices(int rank, int level, int pit, int cc, uint32_t*) = 0;
  virtual int64_t mapperSize(int rank, int level, const uint32_t* block) = 0;
  virtual int64_t mapperIndex(int rank, int level, const uint32_t* block, int codim, int pit, uint32_t*) = 0;
  virtual int64_t mapperSubIndex(int rank, int level, const uint32_t* block, int pit, int cc, uint32_t*) = 0;
  virtual int64_t intersections(int rank, int level, int pit, uint8_t*, uint8_t*, int32_t*, int32_t*, double*, double*, double*) = 0;
  virtual int64_t geometryEval(int rank, int level, int pit, int nqp, const double* qp, double*, double*, double*) = 0;
  virtual int64_t geometryEvalCodim(int rank, int level, int codim, int pit, int nqp, const double* qp, double*, double*, double*, double*) = 0;
  virtual int64_t geogridEval(int rank, int level, int pit, int def, const double* params, int nqp, const double* qp,
                              double*, double*, double*, double*, double*) = 0;
  virtual int64_t geogridIntersections(int rank, int level, int pit, int def, const double* params, int nqp, const double* qp,
                                       double*, double*, double*, double*, double*, double*, double*) = 0;
  virtual int64_t communicate(int level, const uint32_t* block, int iftype, int dir, int op, int narrays, double** arrays, const int* items) = 0;
  virtual int64_t communicateVariable(int level, const uint32_t* block, int iftype, int dir, const int64_t* const* offsets, const double* const* data,
                                      int64_t* nCalls, int64_t* nItems, uint32_t** rIndex, int64_t** rN, double** rData) = 0;
  virtual std::string backup(int rank) = 0;
  virtual int64_t globalIndex(int level, int codim, int32_t** out) = 0;
  virtual int fvInit(int rank, int level, int problem, const double* params, double* c) = 0;
  virtual double fvRun(int level, int problem, const double* params, double** c, double** upd, int steps, double* dt_last) = 0;
};

// run f(rank) on P threads (ranks of the in-process MPI); rethrow the first exception text
template<class F>
void on_all_ranks(int P, F&& f) {
  if (P == 1) { f(0); return; }
  std::vector<std::thread> th;
  std::vector<std::string> errs(P);
  for (int p = 0; p < P; ++p)
    th.emplace_back([&, p] {
      try { f(p); }
      catch (const std::exception& e) { errs[p] = e.what(); if (errs[p].empty()) errs[p] = "exception"; }
      catch (...) { errs[p] = "unknown exception"; }
    });
  for (auto& t : th) t.join();
  for (int p = 0; p < P; ++p)
    if (!errs[p].empty()) { Dune::Exception e; e.message(errs[p]); throw e; }
}

// compile-time dispatch over (codim, PartitionIteratorType)
template<int dim, int cd, class F>
void dispatch_pit(int pit, F&& f) {
  using namespace Dune;
  switch (pit) {
    case Interior_Partition:       f(std::integral_constant<int,cd>(), std::integral_constant<PartitionIteratorType,Interior_Partition>()); break;
    case InteriorBorder_Partition: f(std::integral_constant<int,cd>(), std::integral_constant<PartitionIteratorType,InteriorBorder_Partition>()); break;
    case Overlap_Partition:        f(std::integral_constant<int,cd>(), std::integral_constant<PartitionIteratorType,Overlap_Partition>()); break;
    case OverlapFront_Partition:   f(std::integral_constant<int,cd>(), std::integral_constant<PartitionIteratorType,OverlapFront_Partition>()); break;
    case All_Partition:            f(std::integral_constant<int,cd>(), std::integral_constant<PartitionIteratorType,All_Partition>()); break;
    case Ghost_Partition:          f(std::integral_constant<int,cd>(), std::integral_constant<PartitionIteratorType,Ghost_Partition>()); break;
    default: DUNE_THROW(Dune::GridError, "bad partition iterator type " << pit);
  }
}
template<int dim, class F>
void dispatch(int codim, int pit, F&& f) {
  if (codim == 0) dispatch_pit<dim,0>(pit, f);
  else if (codim == 1) dispatch_pit<dim,1>(pit, f);
  else if (codim == 2) { if constexpr (dim >= 2) dispatch_pit<dim,2>(pit, f); }
  else if (codim == 3) { if constexpr (dim >= 3) dispatch_pit<dim,3>(pit, f); }
  else DUNE_THROW(Dune::GridError, "bad codim " << codim);
}

// the data handle: flat arrays indexed by an MCMG mapper, modelled on the reference's own
// NumPyCommDataHandle (dune/python/grid/numpycommdatahandle.hh:41-142): per entity, for each array,
// for each mapper index of the entity, for each item -> one write / one read+combine.
template<class Mapper>
class ArrayDataHandle : public Dune::CommDataHandleIF<ArrayDataHandle<Mapper>, double> {
public:
  ArrayDataHandle(const Mapper& m, const uint32_t* block, int narrays, double* const* arrays, const int* items, int op)
    : mapper_(m), block_(block), n_(narrays), arrays_(arrays), items_(items), op_(op), itemSize_(0), written(0)
  { for (int a = 0; a < n_; ++a) itemSize_ += items_[a]; }
  bool contains(int /*dim*/, int codim) const { return block_[codim] > 0; }
  bool fixedSize(int, int) const { return true; }
  template<class Entity> std::size_t size(const Entity& e) const { return mapper_.size(e.type())*itemSize_; }
  template<class Buf, class Entity> void gather(Buf& buf, const Entity& e) const {
    for (int a = 0; a < n_; ++a)
      for (const auto index : mapper_.indices(e))
        for (int i = 0; i < items_[a]; ++i) { buf.write(arrays_[a][std::size_t(index)*items_[a]+i]); ++written; }
  }
  template<class Buf, class Entity> void scatter(Buf& buf, const Entity& e, std::size_t) {
    for (int a = 0; a < n_; ++a)
      for (const auto index : mapper_.indices(e))
        for (int i = 0; i < items_[a]; ++i) {
          double& local = arrays_[a][std::size_t(index)*items_[a]+i];
          double remote; buf.read(remote);
          if (op_ == 2) { if (remote >= 0) local = std::min(remote, local); }       // MinimumExchange, utility/globalindexset.hh:153-160
          else if (op_ == 3) { if (remote >= 0) local = remote; }                     // IndexExchange, :262-289
          else local = (op_ == 1) ? remote+local : remote;
        }
  }
  mutable int64_t written;
private:
  const Mapper& mapper_; const uint32_t* block_; int n_; double* const* arrays_; const int* items_; int op_; std::size_t itemSize_;
};

// a data handle of variable size (fixedSize() == false): the CSR arrays say how many items an entity carries; scatter records
// the calls it receives (entity's mapper index, n, the n items) in call order
template<class Mapper>
class VarDataHandle : public Dune::CommDataHandleIF<VarDataHandle<Mapper>, double> {
public:
  VarDataHandle(const Mapper& m, const uint32_t* block, const int64_t* off, const double* data, uint32_t* rIndex, int64_t* rN, double* rData)
    : written(0), calls(0), items(0), mapper_(m), block_(block), off_(off), data_(data), rIndex_(rIndex), rN_(rN), rData_(rData) {}
  bool contains(int /*dim*/, int codim) const { return block_[codim] > 0; }
  bool fixedSize(int, int) const { return false; }
  template<class Entity> std::size_t size(const Entity& e) const { const auto i = mapper_.index(e); return std::size_t(off_[i+1] - off_[i]); }
  template<class Buf, class Entity> void gather(Buf& buf, const Entity& e) const {
    const auto i = mapper_.index(e);
    for (int64_t k = off_[i]; k < off_[i+1]; ++k) { buf.write(data_[k]); ++written; }
  }
  template<class Buf, class Entity> void scatter(Buf& buf, const Entity& e, std::size_t n) {
    if (rIndex_) rIndex_[calls] = uint32_t(mapper_.index(e));
    if (rN_) rN_[calls] = int64_t(n);
    for (std::size_t k = 0; k < n; ++k) { double x; buf.read(x); if (rData_) rData_[items] = x; ++items; }
    ++calls;
  }
  mutable int64_t written;
  int64_t calls, items;
private:
  const Mapper& mapper_; const uint32_t* block_; const int64_t* off_; const double* data_; uint32_t* rIndex_; int64_t* rN_; double* rData_;
};

template<int dim, class CC>
struct World : WorldBase {
  typedef Dune::YaspGrid<dim,CC> Grid;
  typedef typename Grid::LevelGridView GV;
  typedef Dune::MultipleCodimMultipleGeomTypeMapper<GV> Mapper;
  std::unique_ptr<vmpi_world> vw;
  std::vector<vmpi_rank> rk;
  std::vector<std::unique_ptr<Grid>> grids;
  Dune::Yasp::DefaultPartitioning<dim> defPart;
  Dune::Yasp::PowerDPartitioning<dim> powPart;
  std::unique_ptr<Dune::Yasp::FixedSizePartitioning<dim>> fixPart;

  World(const orc_spec& s, int nranks) {
    P = nranks; gdim = dim;
    vw.reset(new vmpi_world(P)); rk.resize(P); grids.resize(P);
    for (int p = 0; p < P; ++p) rk[p] = vmpi_rank{vw.get(), p};
    const Dune::Yasp::Partitioning<dim>* part = &defPart;
    if (s.partitioner == 1) part = &powPart;
    if (s.partitioner == 2) {
      std::array<int,dim> fd; for (int i = 0; i < dim; ++i) fd[i] = s.fixed_dims[i];
      fixPart.reset(new Dune::Yasp::FixedSizePartitioning<dim>(fd)); part = fixPart.get();
    }
    std::array<int,dim> N; for (int i = 0; i < dim; ++i) N[i] = s.N[i];
    std::bitset<dim> per(s.periodic);
    on_all_ranks(P, [&](int p) {
      typename Grid::Communication comm(&rk[p]);
      if constexpr (std::is_same_v<CC, Dune::EquidistantCoordinates<double,dim>>) {
        Dune::FieldVector<double,dim> L; for (int i = 0; i < dim; ++i) L[i] = s.upper[i];
        grids[p].reset(new Grid(L, N, per, s.overlap, comm, part));
      } else if constexpr (std::is_same_v<CC, Dune::EquidistantOffsetCoordinates<double,dim>>) {
        Dune::FieldVector<double,dim> lo, up; for (int i = 0; i < dim; ++i) { lo[i] = s.lower[i]; up[i] = s.upper[i]; }
        grids[p].reset(new Grid(lo, up, N, per, s.overlap, comm, part));
      } else {
        std::array<std::vector<double>,dim> c;
        for (int i = 0; i < dim; ++i) c[i].assign(s.tensor[i], s.tensor[i]+s.N[i]+1);
        grids[p].reset(new Grid(c, per, s.overlap, comm, part));
      }
      grids[p]->refineOptions(s.keep_overlap != 0);
      if (s.refine > 0) grids[p]->globalRefine(s.refine);
    });
  }

  // every rank restored by the reference's BackupRestoreFacility from its own text (yaspgrid/backuprestore.hh:146-218, 268-346)
  World(int nranks, const char* const* texts) {
    P = nranks; gdim = dim;
    vw.reset(new vmpi_world(P)); rk.resize(P); grids.resize(P);
    for (int p = 0; p < P; ++p) rk[p] = vmpi_rank{vw.get(), p};
    on_all_ranks(P, [&](int p) {
      typename Grid::Communication comm(&rk[p]);
      std::istringstream in(texts[p]);
      grids[p].reset(Dune::BackupRestoreFacility<Grid>::restore(in, comm));
    });
  }
  std::string backup(int rank) override {
    std::ostringstream s;
    Dune::BackupRestoreFacility<Grid>::backup(g(rank), s);
    return s.str();
  }

  Grid& g(int rank) { return *grids[rank]; }
  int maxLevel() override { return g(0).maxLevel(); }
  void torusDims(int* d) override { for (int i = 0; i < dim; ++i) d[i] = g(0).torus().dims(i); }
  void domainSize(double* L) override { for (int i = 0; i < dim; ++i) L[i] = g(0).domainSize()[i]; }
  int64_t size(int rank, int level, int codim) override { return g(rank).size(level, codim); }
  int nbs(int rank) override { return int(g(rank).numBoundarySegments()); }
  int overlapSize(int level) override { return g(0).overlapSize(level, 0); }

  int boxes(int rank, int level, int codim, int* out) override {
    auto gl = g(rank).begin(level);
    int k = 0;
    for (auto it = gl->overlapfront[codim].dataBegin(); it != gl->overlapfront[codim].dataEnd(); ++it, ++k) {
      int* o = out + 26*k;
      std::array<int,dim> zero; zero.fill(0); (void)zero;
      o[0] = int(it->shift().to_ulong());
      std::array<int,dim> org = it->origin();
      o[1] = gl->overlapfront[codim].superindex(org, k);          // index offset of the component
      const Dune::YGridComponent<CC>* b[4] = { &*it, gl->overlap[codim].dataBegin()+k,
                                               gl->interiorborder[codim].dataBegin()+k, gl->interior[codim].dataBegin()+k };
      for (int q = 0; q < 4; ++q)
        for (int i = 0; i < 3; ++i) {
          o[2+6*q+i]   = i < dim ? b[q]->origin(i) : 0;
          o[2+6*q+3+i] = i < dim ? b[q]->size(i) : 1;
        }
    }
    return k;
  }

  int list(int rank, int level, int codim, int which, int* out, int cap) override {
    auto gl = g(rank).begin(level);
    const Dune::YGridList<CC>* l = nullptr;
    switch (which) {
      case 0: l = &gl->send_overlapfront_overlapfront[codim]; break;
      case 1: l = &gl->recv_overlapfront_overlapfront[codim]; break;
      case 2: l = &gl->send_overlap_overlapfront[codim]; break;
      case 3: l = &gl->recv_overlapfront_overlap[codim]; break;
      case 4: l = &gl->send_interiorborder_interiorborder[codim]; break;
      case 5: l = &gl->recv_interiorborder_interiorborder[codim]; break;
      case 6: l = &gl->send_interiorborder_overlapfront[codim]; break;
      case 7: l = &gl->recv_overlapfront_interiorborder[codim]; break;
      default: return -1;
    }
    int n = 0;
    for (auto is = l->begin(); is != l->end(); ++is, ++n) {
      if (n >= cap) continue;
      int* o = out + 10*n;
      o[0] = is->rank; o[1] = is->distance; o[2] = int(is->grid.shift().to_ulong());
      typename Dune::YGrid<CC>::Iterator it(is->yg);
      o[3] = it.superindex() - is->grid.superindex(is->grid.origin());   // the artificial index offset
      for (int i = 0; i < 3; ++i) { o[4+i] = i < dim ? is->grid.origin(i) : 0; o[7+i] = i < dim ? is->grid.size(i) : 1; }
    }
    return n;
  }

  int coordinates(int rank, int level, int axis, double* out, int* first) override {
    auto gl = g(rank).begin(level);
    const auto& b = *gl->overlapfront[dim].dataBegin();     // vertex box = all stored vertex indices
    *first = b.origin(axis);
    for (int i = 0; i < b.size(axis); ++i) out[i] = gl->coords.coordinate(axis, b.origin(axis)+i);
    return b.size(axis);
  }

  int64_t entities(int rank, int level, int codim, int pit, uint32_t* index, uint8_t* ptype, int32_t* coord,
                   double* lower, double* upper, double* center, double* volume) override {
    int64_t n = 0;
    Grid& gr = g(rank);
    const auto& is = gr.levelIndexSet(level);
    dispatch<dim>(codim, pit, [&](auto cdc, auto pitc) {
      constexpr int cd = decltype(cdc)::value;
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<cd,pt>(level);
      for (auto it = gr.template lbegin<cd,pt>(level); it != end; ++it, ++n) {
        const auto& e = *it;
        if (index) index[n] = is.index(e);
        if (ptype) ptype[n] = uint8_t(e.partitionType());
        if (coord) for (int i = 0; i < dim; ++i) coord[n*dim+i] = e.impl().transformingsubiterator().coord(i);
        if (lower || upper || center || volume) {
          auto geo = e.geometry();
          if (lower) { auto c = geo.corner(0); for (int i = 0; i < dim; ++i) lower[n*dim+i] = c[i]; }
          if (upper) { auto c = geo.corner(geo.corners()-1); for (int i = 0; i < dim; ++i) upper[n*dim+i] = c[i]; }
          if (center) { auto c = geo.center(); for (int i = 0; i < dim; ++i) center[n*dim+i] = c[i]; }
          if (volume) volume[n] = geo.volume();
        }
      }
    });
    return n;
  }

  int64_t subindices(int rank, int level, int pit, int cc, uint32_t* out) override {
    int64_t n = 0;
    Grid& gr = g(rank);
    const auto& is = gr.levelIndexSet(level);
    dispatch<dim>(0, pit, [&](auto, auto pitc) {
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<0,pt>(level);
      for (auto it = gr.template lbegin<0,pt>(level); it != end; ++it, ++n) {
        const auto& e = *it;
        const int ns = e.subEntities(cc);
        for (int i = 0; i < ns; ++i) out[n*ns+i] = is.subIndex(e, i, cc);
      }
    });
    return n;
  }

  Mapper makeMapper(int rank, int level, const uint32_t* block) {
    std::array<uint32_t,4> b{{0,0,0,0}};
    for (int i = 0; i <= dim; ++i) b[i] = block[i];
    return Mapper(g(rank).levelGridView(level), [b](Dune::GeometryType gt, int dimgrid) { return std::size_t(b[dimgrid - int(gt.dim())]); });
  }
  int64_t mapperSize(int rank, int level, const uint32_t* block) override { return makeMapper(rank, level, block).size(); }
  int64_t mapperIndex(int rank, int level, const uint32_t* block, int codim, int pit, uint32_t* out) override {
    Mapper m = makeMapper(rank, level, block);
    Grid& gr = g(rank);
    int64_t n = 0;
    dispatch<dim>(codim, pit, [&](auto cdc, auto pitc) {
      constexpr int cd = decltype(cdc)::value;
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<cd,pt>(level);
      for (auto it = gr.template lbegin<cd,pt>(level); it != end; ++it, ++n) out[n] = m.index(*it);
    });
    return n;
  }
  int64_t mapperSubIndex(int rank, int level, const uint32_t* block, int pit, int cc, uint32_t* out) override {
    Mapper m = makeMapper(rank, level, block);
    Grid& gr = g(rank);
    int64_t n = 0;
    dispatch<dim>(0, pit, [&](auto, auto pitc) {
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<0,pt>(level);
      for (auto it = gr.template lbegin<0,pt>(level); it != end; ++it, ++n) {
        const auto& e = *it;
        const int ns = e.subEntities(cc);
        for (int i = 0; i < ns; ++i) out[n*ns+i] = m.subIndex(e, i, cc);
      }
    });
    return n;
  }

  int64_t intersections(int rank, int level, int pit, uint8_t* neighbor, uint8_t* boundary, int32_t* outside, int32_t* bsi,
                        double* normal, double* fcenter, double* fvolume) override {
    Grid& gr = g(rank);
    GV gv = gr.levelGridView(level);
    const auto& is = gr.levelIndexSet(level);
    int64_t n = 0;
    dispatch<dim>(0, pit, [&](auto, auto pitc) {
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<0,pt>(level);
      for (auto it = gr.template lbegin<0,pt>(level); it != end; ++it, ++n) {
        const auto& e = *it;
        auto iend = gv.iend(e);
        for (auto iit = gv.ibegin(e); iit != iend; ++iit) {
          const auto& in = *iit;
          const int64_t k = n*2*dim + in.indexInInside();
          if (neighbor) neighbor[k] = in.neighbor();
          if (boundary) boundary[k] = in.boundary();
          if (outside) outside[k] = in.neighbor() ? int32_t(is.index(in.outside())) : -1;
          if (bsi) bsi[k] = in.boundary() ? in.boundarySegmentIndex() : -1;
          if (normal) { auto nn = in.centerUnitOuterNormal(); for (int i = 0; i < dim; ++i) normal[k*dim+i] = nn[i]; }
          if (fcenter || fvolume) {
            auto geo = in.geometry();
            if (fcenter) { auto c = geo.center(); for (int i = 0; i < dim; ++i) fcenter[k*dim+i] = c[i]; }
            if (fvolume) fvolume[k] = geo.volume();
          }
        }
      }
    });
    return n;
  }

  int64_t geometryEval(int rank, int level, int pit, int nqp, const double* qp, double* global, double* jit, double* detJ) override {
    Grid& gr = g(rank);
    int64_t n = 0;
    dispatch<dim>(0, pit, [&](auto, auto pitc) {
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<0,pt>(level);
      for (auto it = gr.template lbegin<0,pt>(level); it != end; ++it, ++n) {
        auto geo = it->geometry();
        for (int q = 0; q < nqp; ++q) {
          Dune::FieldVector<double,dim> x; for (int i = 0; i < dim; ++i) x[i] = qp[q*dim+i];
          const int64_t k = n*nqp + q;
          if (global) { auto y = geo.global(x); for (int i = 0; i < dim; ++i) global[k*dim+i] = y[i]; }
          if (jit) { auto J = geo.jacobianInverseTransposed(x); for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) jit[(k*dim+i)*dim+j] = J[i][j]; }
          if (detJ) detJ[k] = geo.integrationElement(x);
        }
      }
    });
    return n;
  }

  int64_t geometryEvalCodim(int rank, int level, int codim, int pit, int nqp, const double* qp, double* global, double* jt, double* jit, double* detJ) override {
    Grid& gr = g(rank);
    int64_t n = 0;
    dispatch<dim>(codim, pit, [&](auto cdc, auto pitc) {
      constexpr int cd = decltype(cdc)::value;
      constexpr int mydim = dim-cd;
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<cd,pt>(level);
      for (auto it = gr.template lbegin<cd,pt>(level); it != end; ++it, ++n) {
        auto geo = it->geometry();
        for (int q = 0; q < nqp; ++q) {
          Dune::FieldVector<double,mydim> x; for (int i = 0; i < mydim; ++i) x[i] = qp[q*mydim+i];
          const int64_t k = n*nqp + q;
          if (global) { auto y = geo.global(x); for (int i = 0; i < dim; ++i) global[k*dim+i] = y[i]; }
          if (jt) { auto J = geo.jacobianTransposed(x); for (int i = 0; i < mydim; ++i) for (int j = 0; j < dim; ++j) jt[(k*mydim+i)*dim+j] = J[i][j]; }
          if (jit) { auto J = geo.jacobianInverseTransposed(x); for (int i = 0; i < dim; ++i) for (int j = 0; j < mydim; ++j) jit[(k*dim+i)*mydim+j] = J[i][j]; }
          if (detJ) detJ[k] = geo.integrationElement(x);
        }
      }
    });
    return n;
  }

  // GeometryGrid<YaspGrid, AnalyticalCoordFunction>: corners = F(hostGeometry.corner(i))
  // (reference geometrygrid/cornerstorage.hh:54-60, hostcorners.hh:40-43, coordfunctioncaller.hh:40-43), geometry =
  // MultiLinearGeometry over those corners (geometrygrid/geometry.hh:66,197-204) — the multilinear part is the
  // restatement in oracle/common/multilinear.hh (dune-geometry is not vendored).
  int64_t geogridEval(int rank, int level, int pit, int def, const double* params, int nqp, const double* qp,
                      double* corners, double* global, double* jt, double* jit, double* detJ) override {
    Grid& gr = g(rank);
    int64_t n = 0;
    constexpr int nc = 1 << dim;
    const orcgeo::CoordFunction<dim> cf(def, params);
    dispatch<dim>(0, pit, [&](auto, auto pitc) {
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<0,pt>(level);
      for (auto it = gr.template lbegin<0,pt>(level); it != end; ++it, ++n) {
        // corners through the reference's own GeoGrid::CoordVector (HostCorners + CoordFunctionCaller)
        typedef orcgeo::Grid<dim, Grid> GG;
        typedef typename std::decay_t<decltype(*it)> HostElement;
        std::array<Dune::FieldVector<double,dim>, nc> cc;
        Dune::GeoGrid::CoordVector<dim, const GG, false>(*it, cf).calculate(cc);
        double c[nc][dim];
        for (int k = 0; k < nc; ++k) {
          for (int i = 0; i < dim; ++i) c[k][i] = cc[k][i];
          if (corners) for (int i = 0; i < dim; ++i) corners[(n*nc+k)*dim+i] = c[k][i];
        }
        for (int q = 0; q < nqp; ++q) {
          const int64_t k = n*nqp + q;
          const double* x = qp + q*dim;
          double J[dim][dim], Ji[dim][dim], y[dim];
          if (global) { orc::multilinear_global<dim>(c, x, y); for (int i = 0; i < dim; ++i) global[k*dim+i] = y[i]; }
          orc::multilinear_jt<dim>(c, x, J);
          if (jt) for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) jt[(k*dim+i)*dim+j] = J[i][j];
          if (jit) { orc::right_inv<dim>(J, Ji); for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) jit[(k*dim+i)*dim+j] = Ji[i][j]; }
          if (detJ) detJ[k] = orc::sqrt_det_aat<dim>(J);
        }
      }
    });
    return n;
  }


  // GeoGrid::Intersection (geometrygrid/intersection.hh, compiled unmodified) over the Yasp host intersections
  int64_t geogridIntersections(int rank, int level, int pit, int def, const double* params, int nqp, const double* qp,
                               double* corners, double* center, double* volume, double* ion, double* on, double* uon, double* cuon) override {
    typedef orcgeo::Grid<dim, Grid> GG;
    typedef typename GG::Traits::template Codim<0>::GeometryImpl ElementGeo;
    Grid& gr = g(rank);
    GV gv = gr.levelGridView(level);
    const GG gg;
    const orcgeo::CoordFunction<dim> cf(def, params);
    constexpr int nfc = 1 << (dim-1);
    int64_t n = 0;
    dispatch<dim>(0, pit, [&](auto, auto pitc) {
      constexpr Dune::PartitionIteratorType pt = decltype(pitc)::value;
      auto end = gr.template lend<0,pt>(level);
      for (auto it = gr.template lbegin<0,pt>(level); it != end; ++it, ++n) {
        const auto& e = *it;
        Dune::GeoGrid::CoordVector<dim, const GG, false> coords(e, cf);
        const ElementGeo insideGeo(gg, e.type(), coords);                                  // geometrygrid/geometry.hh:138-145
        auto iend = gv.iend(e);
        for (auto iit = gv.ibegin(e); iit != iend; ++iit) {
          typedef typename std::decay_t<decltype(*iit)>::Implementation HostIntersectionImpl;
          Dune::GeoGrid::Intersection<const GG, HostIntersectionImpl> is(iit->impl(), insideGeo);
          const int64_t k = n*2*dim + is.indexInInside();
          auto geo = is.geometry();
          if (corners) for (int c = 0; c < nfc; ++c) { auto x = geo.corner(c); for (int i = 0; i < dim; ++i) corners[(k*nfc+c)*dim+i] = x[i]; }
          if (center) { auto x = geo.center(); for (int i = 0; i < dim; ++i) center[k*dim+i] = x[i]; }
          if (volume) volume[k] = geo.volume();
          for (int q = 0; q < nqp; ++q) {
            Dune::FieldVector<double,dim-1> x; for (int i = 0; i < dim-1; ++i) x[i] = qp[q*(dim-1)+i];
            if (ion) { auto v = is.integrationOuterNormal(x); for (int i = 0; i < dim; ++i) ion[((k*nqp)+q)*dim+i] = v[i]; }
            if (on) { auto v = is.outerNormal(x); for (int i = 0; i < dim; ++i) on[((k*nqp)+q)*dim+i] = v[i]; }
            if (uon) { auto v = is.unitOuterNormal(x); for (int i = 0; i < dim; ++i) uon[((k*nqp)+q)*dim+i] = v[i]; }
          }
          if (cuon) { auto v = is.centerUnitOuterNormal(); for (int i = 0; i < dim; ++i) cuon[k*dim+i] = v[i]; }
        }
      }
    });
    return n;
  }

  int64_t communicate(int level, const uint32_t* block, int iftype, int dir, int op, int narrays, double** arrays, const int* items) override {
    std::vector<int64_t> sent(P, 0);
    on_all_ranks(P, [&](int p) {
      Mapper m = makeMapper(p, level, block);
      ArrayDataHandle<Mapper> dh(m, block, narrays, arrays + std::size_t(p)*narrays, items, op);
      g(p).levelGridView(level).communicate(dh, Dune::InterfaceType(iftype), Dune::CommunicationDirection(dir));
      sent[p] = dh.written;
    });
    int64_t s = 0; for (auto v : sent) s += v;
    return s;
  }

  int64_t communicateVariable(int level, const uint32_t* block, int iftype, int dir, const int64_t* const* offsets, const double* const* data,
                              int64_t* nCalls, int64_t* nItems, uint32_t** rIndex, int64_t** rN, double** rData) override {
    for (int c = 0; c <= dim; ++c) if (block[c] > 1) DUNE_THROW(Dune::GridError, "variable-size handle: block must be 0 or 1");
    std::vector<int64_t> sent(P, 0);
    on_all_ranks(P, [&](int p) {
      Mapper m = makeMapper(p, level, block);
      VarDataHandle<Mapper> dh(m, block, offsets[p], data[p], rIndex ? rIndex[p] : nullptr, rN ? rN[p] : nullptr, rData ? rData[p] : nullptr);
      g(p).levelGridView(level).communicate(dh, Dune::InterfaceType(iftype), Dune::CommunicationDirection(dir));
      sent[p] = dh.written; nCalls[p] = dh.calls; nItems[p] = dh.items;
    });
    int64_t s = 0; for (auto v : sent) s += v;
    return s;
  }

  // the reference's own GlobalIndexSet, constructed collectively on all rank threads
  int64_t globalIndex(int level, int codim, int32_t** out) override {
    std::vector<int64_t> sizes(P, 0);
    on_all_ranks(P, [&](int p) {
      Grid& gr = g(p);
      GV gv = gr.levelGridView(level);
      Dune::GlobalIndexSet<GV> gis(gv, codim);
      const auto& is = gr.levelIndexSet(level);
      for (auto it = gv.template begin<0>(); it != gv.template end<0>(); ++it) {
        const auto& e = *it;
        if (codim == 0) out[p][is.index(e)] = gis.index(e);
        else for (unsigned int i = 0; i < e.subEntities(codim); ++i) out[p][is.subIndex(e, i, codim)] = gis.subIndex(e, i, codim);
      }
      sizes[p] = gis.size(codim);
    });
    return sizes[0];
  }

  int fvInit(int rank, int level, int problem, const double* params, double* c) override {
    Grid& gr = g(rank);
    GV gv = gr.levelGridView(level);
    Mapper m(gv, Dune::mcmgElementLayout());
    for (auto it = gv.template begin<0>(); it != gv.template end<0>(); ++it) {
      auto x = it->geometry().center();
      c[m.index(*it)] = orc::fv_initial(problem, params, dim, x.data());
    }
    return 0;
  }

  // One explicit upwind finite-volume step per rank, in the style of dune-grid-howto's parallel evolve
  // (not vendored in the reference; definition in SURVEY.md §8d): element loop over the interior
  // partition, intersection loop, mapper lookups, then communicate(InteriorBorder_All, Forward).
  double fvRun(int level, int problem, const double* params, double** c, double** upd, int steps, double* dt_last) override {
    std::vector<double> dts(P, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    on_all_ranks(P, [&](int p) {
      Grid& gr = g(p);
      GV gv = gr.levelGridView(level);
      Mapper mapper(gv, Dune::mcmgElementLayout());
      const uint32_t block[4] = {1,0,0,0};
      const int items[1] = {1};
      std::vector<double> update(mapper.size(), 0.0);
      double* cc = c[p];
      for (int s = 0; s < steps; ++s) {
        double dt = std::numeric_limits<double>::infinity();
        auto endit = gv.template end<0,Dune::Interior_Partition>();
        for (auto it = gv.template begin<0,Dune::Interior_Partition>(); it != endit; ++it) {
          const auto& e = *it;
          const double volume = e.geometry().volume();
          const auto i = mapper.index(e);
          double sumfactor = 0.0, u = 0.0;
          auto isend = gv.iend(e);
          for (auto is = gv.ibegin(e); is != isend; ++is) {
            const auto& in = *is;
            auto igeo = in.geometry();
            auto fc = igeo.center();
            auto nrm = in.centerUnitOuterNormal();
            Dune::FieldVector<double,dim> vel;
            orc::fv_velocity(problem, params, dim, fc.data(), vel.data());
            const double factor = (vel*nrm)*igeo.volume()/volume;
            if (factor >= 0) { sumfactor += factor; u -= cc[i]*factor; }
            else if (in.neighbor()) { const auto j = mapper.index(in.outside()); u -= cc[j]*factor; }
            else if (in.boundary()) { u -= orc::fv_boundary(problem, params, dim, fc.data())*factor; }
          }
          update[i] = u;
          dt = std::min(dt, 1.0/sumfactor);
        }
        dt = gr.comm().min(dt);
        dt *= 0.99;
        for (auto it = gv.template begin<0,Dune::Interior_Partition>(); it != endit; ++it) {
          const auto i = mapper.index(*it);
          cc[i] += dt*update[i];
        }
        ArrayDataHandle<Mapper> dh(mapper, block, 1, &cc, items, 0);
        gv.communicate(dh, Dune::InteriorBorder_All_Interface, Dune::ForwardCommunication);
        dts[p] = dt;
      }
      if (upd && upd[p]) std::memcpy(upd[p], update.data(), update.size()*sizeof(double));
    });
    auto t1 = std::chrono::steady_clock::now();
    if (dt_last) *dt_last = dts[0];
    return std::chrono::duration<double>(t1-t0).count();
  }
};

template<int dim>
WorldBase* make_world(const orc_spec& s, int nranks) {
  switch (s.coord_mode) {
    case 0: return new World<dim, Dune::EquidistantCoordinates<double,dim>>(s, nranks);
    case 1: return new World<dim, Dune::EquidistantOffsetCoordinates<double,dim>>(s, nranks);
    case 2: return new World<dim, Dune::TensorProductCoordinates<double,dim>>(s, nranks);
  }
  DUNE_THROW(Dune::GridError, "bad coord_mode " << s.coord_mode);
}

template<int dim>
WorldBase* restore_world(int coord_mode, int nranks, const char* const* texts) {
  switch (coord_mode) {
    case 0: return new World<dim, Dune::EquidistantCoordinates<double,dim>>(nranks, texts);
    case 1: return new World<dim, Dune::EquidistantOffsetCoordinates<double,dim>>(nranks, texts);
    case 2: return new World<dim, Dune::TensorProductCoordinates<double,dim>>(nranks, texts);
  }
  DUNE_THROW(Dune::GridError, "bad coord_mode " << coord_mode);
}

template<class F, class R>
R guarded(R onerr, F&& f) {
  try { g_err.clear(); return f(); }
  catch (const std::exception& e) { g_err = e.what(); if (g_err.empty()) g_err = "exception"; }
  catch (...) { g_err = "unknown exception"; }
  return onerr;
}
inline WorldBase* W(void* w) { return static_cast<WorldBase*>(w); }

template<int d>
int partition_d(const int* N, int P, int overlap, int partitioner, const int* fixed, int* out) {
  std::array<int,d> size, dims, fd;
  for (int i = 0; i < d; ++i) { size[i] = N[i]; fd[i] = fixed ? fixed[i] : 1; }
  if (partitioner == 0) Dune::Yasp::DefaultPartitioning<d>().partition(size, P, dims, overlap);
  else if (partitioner == 1) Dune::Yasp::PowerDPartitioning<d>().partition(size, P, dims, overlap);
  else Dune::Yasp::FixedSizePartitioning<d>(fd).partition(size, P, dims, overlap);
  for (int i = 0; i < d; ++i) out[i] = dims[i];
  return 0;
}

} // namespace

extern "C" {

const char* orc_kind(void) { return "reference"; }
const char* orc_last_error(void) { return g_err.c_str(); }

int orc_partition(int dim, const int* N, int P, int overlap, int partitioner, const int* fixed, int* dims_out) {
  return guarded(-1, [&] {
    switch (dim) {
      case 1: return partition_d<1>(N, P, overlap, partitioner, fixed, dims_out);
      case 2: return partition_d<2>(N, P, overlap, partitioner, fixed, dims_out);
      case 3: return partition_d<3>(N, P, overlap, partitioner, fixed, dims_out);
      case 4: return partition_d<4>(N, P, overlap, partitioner, fixed, dims_out);
    }
    return -1;
  });
}
int orc_entity_shift(int dim, int i, int cc) {
  switch (dim) { case 1: return int(Dune::Yasp::entityShift<1>(i,cc).to_ulong()); case 2: return int(Dune::Yasp::entityShift<2>(i,cc).to_ulong());
                 case 3: return int(Dune::Yasp::entityShift<3>(i,cc).to_ulong()); case 4: return int(Dune::Yasp::entityShift<4>(i,cc).to_ulong()); }
  return -1;
}
int orc_entity_move(int dim, int i, int cc) {
  switch (dim) { case 1: return int(Dune::Yasp::entityMove<1>(i,cc).to_ulong()); case 2: return int(Dune::Yasp::entityMove<2>(i,cc).to_ulong());
                 case 3: return int(Dune::Yasp::entityMove<3>(i,cc).to_ulong()); case 4: return int(Dune::Yasp::entityMove<4>(i,cc).to_ulong()); }
  return -1;
}
int orc_sub_entities(int dim, int d, int c) {
  switch (dim) { case 1: return Dune::Yasp::subEnt<1>(d,c); case 2: return Dune::Yasp::subEnt<2>(d,c);
                 case 3: return Dune::Yasp::subEnt<3>(d,c); case 4: return Dune::Yasp::subEnt<4>(d,c); }
  return -1;
}

void* orc_create(const orc_spec* spec, int nranks) {
  return guarded((void*)nullptr, [&]() -> void* {
    if (spec->dim == 2) return make_world<2>(*spec, nranks);
    if (spec->dim == 3) return make_world<3>(*spec, nranks);
    DUNE_THROW(Dune::GridError, "oracle supports dim 2 and 3");
  });
}
void orc_destroy(void* w) { delete W(w); }
int orc_max_level(void* w) { return W(w)->maxLevel(); }
void orc_torus_dims(void* w, int* d) { W(w)->torusDims(d); }
void orc_domain_size(void* w, double* L) { W(w)->domainSize(L); }
int64_t orc_size(void* w, int rank, int level, int codim) { return guarded(int64_t(-1), [&]{ return W(w)->size(rank, level, codim); }); }
int orc_num_boundary_segments(void* w, int rank) { return W(w)->nbs(rank); }
int orc_overlap_size(void* w, int level) { return guarded(-1, [&]{ return W(w)->overlapSize(level); }); }
int orc_boxes(void* w, int rank, int level, int codim, int* out) { return guarded(-1, [&]{ return W(w)->boxes(rank, level, codim, out); }); }
int orc_list(void* w, int rank, int level, int codim, int which, int* out, int cap) { return guarded(-1, [&]{ return W(w)->list(rank, level, codim, which, out, cap); }); }
int orc_coordinates(void* w, int rank, int level, int axis, double* out, int* first) { return guarded(-1, [&]{ return W(w)->coordinates(rank, level, axis, out, first); }); }
int64_t orc_entities(void* w, int rank, int level, int codim, int pit, uint32_t* index, uint8_t* ptype, int32_t* coord,
                     double* lower, double* upper, double* center, double* volume) {
  return guarded(int64_t(-1), [&]{ return W(w)->entities(rank, level, codim, pit, index, ptype, coord, lower, upper, center, volume); });
}
int64_t orc_subindices(void* w, int rank, int level, int pit, int cc, uint32_t* out) { return guarded(int64_t(-1), [&]{ return W(w)->subindices(rank, level, pit, cc, out); }); }
int64_t orc_mapper_size(void* w, int rank, int level, const uint32_t* block) { return guarded(int64_t(-1), [&]{ return W(w)->mapperSize(rank, level, block); }); }
int64_t orc_mapper_index(void* w, int rank, int level, const uint32_t* block, int codim, int pit, uint32_t* out) {
  return guarded(int64_t(-1), [&]{ return W(w)->mapperIndex(rank, level, block, codim, pit, out); });
}
int64_t orc_mapper_subindex(void* w, int rank, int level, const uint32_t* block, int pit, int cc, uint32_t* out) {
  return guarded(int64_t(-1), [&]{ return W(w)->mapperSubIndex(rank, level, block, pit, cc, out); });
}
int64_t orc_intersections(void* w, int rank, int level, int pit, uint8_t* neighbor, uint8_t* boundary, int32_t* outside, int32_t* bsi,
                          double* normal, double* fc, double* fv) {
  return guarded(int64_t(-1), [&]{ return W(w)->intersections(rank, level, pit, neighbor, boundary, outside, bsi, normal, fc, fv); });
}
int64_t orc_geometry_eval(void* w, int rank, int level, int pit, int nqp, const double* qp, double* global, double* jit, double* detJ) {
  return guarded(int64_t(-1), [&]{ return W(w)->geometryEval(rank, level, pit, nqp, qp, global, jit, detJ); });
}
int64_t orc_geometry_eval_codim(void* w, int rank, int level, int codim, int pit, int nqp, const double* qp, double* global, double* jt, double* jit, double* detJ) {
  return guarded(int64_t(-1), [&]{ return W(w)->geometryEvalCodim(rank, level, codim, pit, nqp, qp, global, jt, jit, detJ); });
}
int64_t orc_geogrid_eval(void* w, int rank, int level, int pit, int def, const double* params, int nqp, const double* qp,
                         double* corners, double* global, double* jt, double* jit, double* detJ) {
  return guarded(int64_t(-1), [&]{ return W(w)->geogridEval(rank, level, pit, def, params, nqp, qp, corners, global, jt, jit, detJ); });
}
int64_t orc_geogrid_intersections(void* w, int rank, int level, int pit, int def, const double* params, int nqp, const double* qp,
                                  double* corners, double* center, double* volume, double* ion, double* on, double* uon, double* cuon) {
  return guarded(int64_t(-1), [&]{ return W(w)->geogridIntersections(rank, level, pit, def, params, nqp, qp, corners, center, volume, ion, on, uon, cuon); });
}
int64_t orc_communicate(void* w, int level, const uint32_t* block, int iftype, int dir, int op, int narrays, double** arrays, const int* items) {
  return guarded(int64_t(-1), [&]{ return W(w)->communicate(level, block, iftype, dir, op, narrays, arrays, items); });
}
int64_t orc_communicate_variable(void* w, int level, const uint32_t* block, int iftype, int dir, const int64_t* const* offsets,
                                 const double* const* data, int64_t* n_calls, int64_t* n_items, uint32_t** recv_index, int64_t** recv_n, double** recv_data) {
  return guarded(int64_t(-1), [&]{ return W(w)->communicateVariable(level, block, iftype, dir, offsets, data, n_calls, n_items, recv_index, recv_n, recv_data); });
}
int64_t orc_global_index(void* w, int level, int codim, int32_t** out) {
  return guarded(int64_t(-1), [&] { return W(w)->globalIndex(level, codim, out); });
}
int orc_backup(void* w, int rank, char* out, int cap) {
  return guarded(-1, [&] {
    const std::string t = W(w)->backup(rank);
    if (out && cap > int(t.size())) std::memcpy(out, t.c_str(), t.size()+1);
    return int(t.size());
  });
}
void* orc_restore(int dim, int coord_mode, int nranks, const char* const* texts) {
  return guarded((void*)nullptr, [&]() -> void* {
    if (dim == 2) return restore_world<2>(coord_mode, nranks, texts);
    if (dim == 3) return restore_world<3>(coord_mode, nranks, texts);
    DUNE_THROW(Dune::GridError, "oracle supports dim 2 and 3");
  });
}
int64_t orc_vtu_piece(int outputtype, unsigned ncells, unsigned npoints, int narrays, const int* section, const char* const* names,
                      const int* ncomps, const int* nitems, const int* precision, const double* const* values,
                      const char* point_scalars, const char* point_vectors, const char* cell_scalars, const char* cell_vectors,
                      char* out, int64_t cap) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    std::ostringstream s;
    {
      Dune::VTK::VTUWriter writer(s, Dune::VTK::OutputType(outputtype), Dune::VTK::unstructuredGrid);
      static const Dune::VTK::Precision precs[4] = { Dune::VTK::Precision::float32, Dune::VTK::Precision::float64,
                                                     Dune::VTK::Precision::int32, Dune::VTK::Precision::uint8 };
      auto arrays = [&](int sec) {
        for (int a = 0; a < narrays; ++a) if (section[a] == sec) {
          std::unique_ptr<Dune::VTK::DataArrayWriter> p(writer.makeArrayWriter(names[a], ncomps[a], nitems[a], precs[precision[a]]));
          if (!p->writeIsNoop())
            for (std::int64_t k = 0; k < std::int64_t(ncomps[a])*nitems[a]; ++k) p->write(values[a][k]);
        }
      };
      auto all = [&] {                                               // VTKWriter::writeAllData, vtkwriter.hh:1205-1217
        bool any = false;
        for (int a = 0; a < narrays; ++a) any = any || section[a] == 0;
        if (any) { writer.beginPointData(point_scalars, point_vectors); arrays(0); writer.endPointData(); }       // :1342-1353
        any = false;
        for (int a = 0; a < narrays; ++a) any = any || section[a] == 1;
        if (any) { writer.beginCellData(cell_scalars, cell_vectors); arrays(1); writer.endCellData(); }           // :1300-1338
        writer.beginPoints(); arrays(2); writer.endPoints();
        writer.beginCells(); arrays(3); writer.endCells();
      };
      writer.beginMain(ncells, npoints);
      all();
      writer.endMain();
      if (writer.beginAppended()) all();
      writer.endAppended();
    }
    const std::string t = s.str();
    if (out && cap >= std::int64_t(t.size())) std::memcpy(out, t.data(), t.size());
    return std::int64_t(t.size());
  });
}
int orc_fv_init(void* w, int rank, int level, int problem, const double* params, double* c) {
  return guarded(-1, [&]{ return W(w)->fvInit(rank, level, problem, params, c); });
}
double orc_fv_step(void* w, int level, int problem, const double* params, double** c, double** upd) {
  return guarded(-1.0, [&]{ double dt = 0; W(w)->fvRun(level, problem, params, c, upd, 1, &dt); return dt; });
}
double orc_fv_run(void* w, int level, int problem, const double* params, double** c, int steps, double* dt_last) {
  return guarded(-1.0, [&]{ return W(w)->fvRun(level, problem, params, c, nullptr, steps, dt_last); });
}

} // extern "C"
