// bulkgrid.hh — header-only C++17 "bulk GridView" layer over the C ABI of libdgb.so (include/dgb.h).
//
// It keeps the operator surface of the reference for the YaspGrid leaf-sweep path, with per-entity calls turned into
// per-view bulk calls that return device arrays in the reference's iteration order:
//
//   Dune::YaspGrid<dim,Coordinates>(L, s, periodic, overlap, comm)   dune/grid/yaspgrid.hh:904-909,974-980,1043-1047
//     -> DuneB200::BulkYaspGrid<dim,Coordinates>(L, s, periodic, overlap, rank, nranks, comm)   (Coordinates = one of the tag
//        types EquidistantCoordinates / EquidistantOffsetCoordinates / TensorProductCoordinates<ct,dim>, coordinates.hh:27,129,243)
//   grid.leafGridView() / levelGridView(l), globalRefine, refineOptions   dune/grid/common/grid.hh:878-891, yaspgrid.hh:1216,1270
//   gv.size(codim), gv.overlapSize(codim), gv.ghostSize(codim)            dune/grid/common/gridview.hh:186-198,343-355
//   for (e : elements(gv, ps)) { is.index(e); e.partitionType(); e.geometry().center()/volume() }
//     -> gv.elements<codim>(pit)  (index, partitionType, lower, upper, center, volume arrays)     rangegenerators.hh:881
//   for (is : intersections(gv,e)) { is.neighbor(); is.boundary(); is.outside(); is.boundarySegmentIndex();
//                                    is.centerUnitOuterNormal(); is.geometry().center()/volume() }
//     -> gv.intersections(pit)                                                                     rangegenerators.hh:845
//   geo.global(x) / jacobianInverseTransposed(x) / integrationElement(x) at quadrature points
//     -> gv.geometry(qp, pit)                                                                      common/geometry.hh:194-371
//   MultipleCodimMultipleGeomTypeMapper<GV>(gv, layout).index(e)/subIndex(e,i,cc)/size()           common/mcmgmapper.hh:129-343
//     -> BulkMapper(gv, layout).index(codim,pit) / subIndex(cc,pit) / size()
//   for (e : elements(gv)) sum += f(e);  for (is : intersections(gv,e)) sum += f(is)   with a compiled-in device functor
//     -> gv.forEachElement(functorId, args) / gv.forEachIntersection(functorId, args)             (csrc/dgb_functors.cuh)
//   is.unitOuterNormal / integrationOuterNormal / geometryInInside / geometryInOutside             yaspgridintersection.hh:163-228
//     -> gv.intersectionNormals(pit), gv.geometryInInside(k), gv.geometryInOutside(k)
//   gv.communicate(dataHandle, InteriorBorder_All_Interface, ForwardCommunication)                 common/gridview.hh:334-341
//     -> gv.communicate(BulkDataHandle{mapper, arrays...}, iftype, dir)   (the shape of NumPyCommDataHandle,
//        dune/python/grid/numpycommdatahandle.hh:41-142)
//
// Exceptions mirror the reference's: GridError (yaspgrid.hh:788,1053,1219), RangeError (:1745,1817), Exception.
// Only the CUDA runtime API header is needed (no nvcc): g++ -std=c++17 -I include -I $CUDA/include x.cc -L... -ldgb -lcudart
#ifndef DUNE_GRID_B200_BULKGRID_HH
#define DUNE_GRID_B200_BULKGRID_HH
#include <array>
#include <bitset>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <iterator>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <cuda_runtime_api.h>

#include "../dgb.h"

namespace DuneB200 {

// dune/grid/common/gridenums.hh — identical names and values
enum PartitionType { InteriorEntity = 0, BorderEntity = 1, OverlapEntity = 2, FrontEntity = 3, GhostEntity = 4 };
enum InterfaceType { InteriorBorder_InteriorBorder_Interface = 0, InteriorBorder_All_Interface = 1,
                     Overlap_OverlapFront_Interface = 2, Overlap_All_Interface = 3, All_All_Interface = 4 };
enum PartitionIteratorType { Interior_Partition = 0, InteriorBorder_Partition = 1, Overlap_Partition = 2,
                             OverlapFront_Partition = 3, All_Partition = 4, Ghost_Partition = 5 };
enum CommunicationDirection { ForwardCommunication = 0, BackwardCommunication = 1 };
enum class CommOp { set = 0, add = 1 };                      // dune/python/grid/mapper.hh:117-137

struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct GridError : Exception { using Exception::Exception; };
struct RangeError : Exception { using Exception::Exception; };
struct CudaError : Exception { using Exception::Exception; };

inline void check(int rc) {
  if (rc == DGB_OK) return;
  const std::string msg = dgb_last_error();
  switch (rc) {
    case DGB_EGRID: throw GridError(msg);
    case DGB_ERANGE: throw RangeError(msg);
    case DGB_ECUDA: case DGB_ENCCL: throw CudaError(msg);
    default: throw Exception(msg);
  }
}

// caller-owned device memory (RAII); the library never allocates user data
template<class T>
class DeviceArray {
  T* p_ = nullptr; std::size_t n_ = 0;
public:
  DeviceArray() = default;
  explicit DeviceArray(std::size_t n) : n_(n) {
    if (n && cudaMalloc(reinterpret_cast<void**>(&p_), n*sizeof(T)) != cudaSuccess) throw CudaError("cudaMalloc failed");
  }
  DeviceArray(DeviceArray&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
  DeviceArray& operator=(DeviceArray&& o) noexcept { std::swap(p_, o.p_); std::swap(n_, o.n_); return *this; }
  DeviceArray(const DeviceArray&) = delete;
  DeviceArray& operator=(const DeviceArray&) = delete;
  ~DeviceArray() { if (p_) cudaFree(p_); }
  T* data() { return p_; }
  const T* data() const { return p_; }
  std::size_t size() const { return n_; }
  std::vector<T> toHost(cudaStream_t s = nullptr) const {
    std::vector<T> h(n_);
    if (n_) { cudaMemcpyAsync(h.data(), p_, n_*sizeof(T), cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
    return h;
  }
  void fromHost(const std::vector<T>& h, cudaStream_t s = nullptr) {
    if (h.size() != n_) throw Exception("DeviceArray::fromHost: size mismatch");
    if (n_) { cudaMemcpyAsync(p_, h.data(), n_*sizeof(T), cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); }
  }
};

// the NCCL communicator of the torus exchange (replaces the MPI communicator handed to YaspGrid)
class Communicator {
  dgb_comm* c_ = nullptr;
public:
  Communicator() = default;
  Communicator(const void* ncclUniqueId128, int rank, int nranks) { check(dgb_comm_init(ncclUniqueId128, rank, nranks, &c_)); }
  Communicator(Communicator&& o) noexcept : c_(o.c_) { o.c_ = nullptr; }
  Communicator(const Communicator&) = delete;
  ~Communicator() { if (c_) dgb_comm_destroy(c_); }
  static std::array<char,128> uniqueId() { std::array<char,128> id{}; check(dgb_comm_unique_id(id.data())); return id; }
  dgb_comm* handle() const { return c_; }
};

// coordinate containers of YaspGrid (dune/grid/yaspgrid/coordinates.hh:27,129,243) as tag types: they select the constructor
template<class ct, int dim> struct EquidistantCoordinates { typedef ct ctype; static constexpr int mode = DGB_COORD_EQUIDISTANT; };
template<class ct, int dim> struct EquidistantOffsetCoordinates { typedef ct ctype; static constexpr int mode = DGB_COORD_EQUIDISTANT_OFFSET; };
template<class ct, int dim> struct TensorProductCoordinates { typedef ct ctype; static constexpr int mode = DGB_COORD_TENSORPRODUCT; };

template<int dim> class BulkYaspGridBase;
template<int dim, class Coordinates> class BulkYaspGrid;
template<int dim> class BulkGridView;
template<int dim> class BulkMapper;
template<int dim> struct BulkVariableDataHandle;

// result of elements<codim>(): arrays in iteration order of the partition; vectors are component-major [d*n + e]
struct ElementRange {
  std::int64_t n = 0;
  DeviceArray<std::uint32_t> index;            // indexSet().index(e)
  DeviceArray<std::uint8_t> partitionType;     // e.partitionType()
  DeviceArray<double> lower, upper, center;    // geometry().corner(0), corner(last), center()
  DeviceArray<double> volume;                  // geometry().volume()
};
// result of intersections(): per cell and face k in reference order (dir = k/2, side = k%2): [k*n + cell]
struct IntersectionRange {
  std::int64_t n = 0;
  DeviceArray<std::uint8_t> neighbor, boundary;
  DeviceArray<std::int32_t> outside;               // indexSet().index(is.outside()) or -1
  DeviceArray<std::int32_t> boundarySegmentIndex;  // or -1
  DeviceArray<double> center;                      // [(k*dim+d)*n + cell]
  DeviceArray<double> volume;
};
struct GeometryRange {
  std::int64_t n = 0; int nqp = 0;
  DeviceArray<double> global;                      // [(q*dim+d)*n + e]
  DeviceArray<double> jacobianInverseTransposed;   // [((q*dim+r)*dim+c)*n + e]
  DeviceArray<double> integrationElement;          // [q*n + e]
};

// MCMGLayout for Yasp: block size per codimension (mcmgmapper.hh:64-110; one geometry type per codim)
template<int dim> std::array<std::uint32_t,4> mcmgElementLayout() { std::array<std::uint32_t,4> l{}; l[0] = 1; return l; }
template<int dim> std::array<std::uint32_t,4> mcmgVertexLayout() { std::array<std::uint32_t,4> l{}; l[dim] = 1; return l; }

template<int dim>
class BulkGridView {
  const BulkYaspGridBase<dim>* grid_; dgb_view v_;
  friend class BulkYaspGridBase<dim>; friend class BulkMapper<dim>;
  BulkGridView(const BulkYaspGridBase<dim>* g, const dgb_view& v) : grid_(g), v_(v) {}
public:
  enum { dimension = dim };
  const dgb_view& descriptor() const { return v_; }
  const BulkYaspGridBase<dim>& grid() const { return *grid_; }
  int level() const { return v_.level; }
  std::int64_t size(int codim) const { std::int64_t n; check(dgb_view_size(&v_, codim, &n)); return n; }
  std::int64_t size(int codim, PartitionIteratorType pit) const { std::int64_t n; check(dgb_view_partition_size(&v_, codim, pit, &n)); return n; }
  int overlapSize(int codim) const { int n; check(dgb_view_overlap_size(&v_, codim, &n)); return n; }
  int ghostSize(int codim) const { int n; check(dgb_view_ghost_size(&v_, codim, &n)); return n; }

  template<int codim = 0>
  ElementRange elements(PartitionIteratorType pit = All_Partition, bool geometry = true, cudaStream_t s = nullptr) const {
    ElementRange r; r.n = size(codim, pit);
    r.index = DeviceArray<std::uint32_t>(r.n); r.partitionType = DeviceArray<std::uint8_t>(r.n);
    if (geometry) { r.lower = DeviceArray<double>(dim*r.n); r.upper = DeviceArray<double>(dim*r.n);
                    r.center = DeviceArray<double>(dim*r.n); r.volume = DeviceArray<double>(r.n); }
    check(dgb_elements_materialize(&v_, codim, pit, r.index.data(), r.partitionType.data(), nullptr, r.lower.data(), r.upper.data(),
                                   r.center.data(), r.volume.data(), s));
    return r;
  }
  IntersectionRange intersections(PartitionIteratorType pit = All_Partition, bool geometry = true, cudaStream_t s = nullptr) const {
    IntersectionRange r; r.n = size(0, pit);
    const std::size_t m = 2*dim*r.n;
    r.neighbor = DeviceArray<std::uint8_t>(m); r.boundary = DeviceArray<std::uint8_t>(m);
    r.outside = DeviceArray<std::int32_t>(m); r.boundarySegmentIndex = DeviceArray<std::int32_t>(m);
    if (geometry) { r.center = DeviceArray<double>(dim*m); r.volume = DeviceArray<double>(m); }
    check(dgb_intersections_materialize(&v_, pit, r.neighbor.data(), r.boundary.data(), r.outside.data(), r.boundarySegmentIndex.data(),
                                        r.center.data(), r.volume.data(), s));
    return r;
  }
  // YaspIntersection::centerUnitOuterNormal (yaspgridintersection.hh:177-182): +-e_dir for face k, the same for all cells
  std::array<double,dim> centerUnitOuterNormal(int k) const { std::array<double,dim> n{}; n[k/2] = (k%2) ? 1.0 : -1.0; return n; }
  GeometryRange geometry(const std::vector<std::array<double,dim>>& qp, PartitionIteratorType pit = All_Partition, cudaStream_t s = nullptr) const {
    GeometryRange r; r.n = size(0, pit); r.nqp = int(qp.size());
    r.global = DeviceArray<double>(std::size_t(r.nqp)*dim*r.n);
    r.jacobianInverseTransposed = DeviceArray<double>(std::size_t(r.nqp)*dim*dim*r.n);
    r.integrationElement = DeviceArray<double>(std::size_t(r.nqp)*r.n);
    check(dgb_geometry_eval(&v_, pit, r.nqp, qp.empty() ? nullptr : qp[0].data(), r.global.data(), r.jacobianInverseTransposed.data(),
                            r.integrationElement.data(), s));
    return r;
  }
  // unitOuterNormal (== outerNormal == centerUnitOuterNormal) and integrationOuterNormal of every face of every cell: [(k*dim+d)*n + cell]
  std::pair<DeviceArray<double>, DeviceArray<double>> intersectionNormals(PartitionIteratorType pit = All_Partition, cudaStream_t s = nullptr) const {
    const std::size_t m = std::size_t(2*dim*dim)*size(0, pit);
    DeviceArray<double> u(m), w(m);
    check(dgb_intersection_normals(&v_, pit, u.data(), w.data(), s));
    return {std::move(u), std::move(w)};
  }
  // YaspIntersection::geometryInInside / geometryInOutside of face k: (lower, upper) of the unit face in local coordinates
  std::pair<std::array<double,dim>, std::array<double,dim>> geometryInInside(int k) const { return localFace(k, 0); }
  std::pair<std::array<double,dim>, std::array<double,dim>> geometryInOutside(int k) const { return localFace(k, 1); }
  // sum over the element / intersection loop of a compiled-in device functor (csrc/dgb_functors.cuh), one fused kernel;
  // `values` (optional, caller-owned: n resp. 2*dim*n doubles) receives the loop body's value per element / face
  double forEachElement(int functorId, const std::vector<double>& args = {}, PartitionIteratorType pit = All_Partition,
                        double* d_values = nullptr, cudaStream_t s = nullptr) const { return forEach(false, functorId, args, pit, d_values, s); }
  double forEachIntersection(int functorId, const std::vector<double>& args = {}, PartitionIteratorType pit = All_Partition,
                             double* d_values = nullptr, cudaStream_t s = nullptr) const { return forEach(true, functorId, args, pit, d_values, s); }
  // communicate: the data handle is "arrays indexed by a mapper" with a combine operation
  template<class DataHandle>
  void communicate(DataHandle& dh, InterfaceType iftype, CommunicationDirection dir, cudaStream_t s = nullptr) const;
  void communicate(BulkVariableDataHandle<dim>& dh, InterfaceType iftype, CommunicationDirection dir, cudaStream_t s = nullptr) const;
private:
  std::pair<std::array<double,dim>, std::array<double,dim>> localFace(int k, int outside) const {
    double lo[3], hi[3]; check(dgb_intersection_local_geometry(dim, k, outside, lo, hi));
    std::array<double,dim> a{}, b{}; for (int i = 0; i < dim; ++i) { a[i] = lo[i]; b[i] = hi[i]; }
    return {a, b};
  }
  double forEach(bool faces, int id, const std::vector<double>& args, PartitionIteratorType pit, double* d_values, cudaStream_t s) const {
    DeviceArray<double> sum(1);
    check((faces ? dgb_for_each_intersection : dgb_for_each_element)(&v_, pit, id, args.empty() ? nullptr : args.data(), int(args.size()), d_values, sum.data(), s));
    return sum.toHost(s)[0];
  }
};

template<int dim>
class BulkMapper {
  const BulkGridView<dim>* gv_; dgb_mapper m_;
public:
  BulkMapper(const BulkGridView<dim>& gv, const std::array<std::uint32_t,4>& layout) : gv_(&gv) { check(dgb_mapper_init(&gv.v_, layout.data(), &m_)); }
  const dgb_mapper& descriptor() const { return m_; }
  std::int64_t size() const { std::int64_t n; check(dgb_mapper_size(&m_, &n)); return n; }
  DeviceArray<std::uint32_t> index(int codim, PartitionIteratorType pit = All_Partition, cudaStream_t s = nullptr) const {
    DeviceArray<std::uint32_t> out(gv_->size(codim, pit));
    check(dgb_mapper_index(&gv_->v_, &m_, codim, pit, out.data(), s));
    return out;
  }
  // [i*n + cell]: subIndex(e, i, cc) for all cells of the partition
  DeviceArray<std::uint32_t> subIndex(int cc, PartitionIteratorType pit = All_Partition, cudaStream_t s = nullptr) const {
    std::size_t nsub = 1; for (int k = 0; k < cc; ++k) nsub = nsub*(dim-k)/(k+1);
    nsub <<= cc;
    DeviceArray<std::uint32_t> out(nsub*gv_->size(0, pit));
    check(dgb_mapper_subindex(&gv_->v_, &m_, cc, pit, out.data(), s));
    return out;
  }
};

// NumPyCommDataHandle-shaped data handle: arrays[a] holds itemSize[a] doubles per mapper entry
template<int dim>
struct BulkDataHandle {
  const BulkMapper<dim>& mapper;
  std::vector<double*> arrays;
  std::vector<std::int32_t> itemSizes;
  CommOp op = CommOp::set;
};

// A data handle of VARIABLE size (CommDataHandleIF::fixedSize() == false, common/datahandleif.hh:97-125): size(e) and gather(e) are the
// CSR pair (offsets by mapper index, items); after communicate() the scatter(buff, e, n) calls the reference would make are in
// recvIndex / recvOffsets / recvData, in the reference's order (dgb.h, dgb_commvar_*)
template<int dim>
struct BulkVariableDataHandle {
  const BulkMapper<dim>& mapper;
  const std::int64_t* sendOffsets;       // device, mapper.size()+1 entries
  const double* sendData;                // device
  DeviceArray<std::uint32_t> recvIndex;
  DeviceArray<std::int64_t> recvOffsets;
  DeviceArray<double> recvData;
};

// everything that does not depend on the coordinate container; BulkYaspGrid<dim,Coordinates> below adds the constructor of its container
template<int dim>
class BulkYaspGridBase {
protected:
  dgb_grid* g_ = nullptr; dgb_comm* comm_ = nullptr;
  std::vector<std::vector<double>> keep_;
  void create(dgb_grid_spec& s, int rank, int nranks) { check(dgb_grid_create(&s, rank, nranks, &g_)); }
  static dgb_grid_spec base(const std::array<int,dim>& s, std::bitset<dim> periodic, int overlap) {
    dgb_grid_spec sp{}; sp.dim = dim; sp.overlap = overlap; sp.periodic = unsigned(periodic.to_ulong()); sp.partitioner = DGB_PART_DEFAULT;
    for (int i = 0; i < 3; ++i) { sp.N[i] = i < dim ? s[i] : 1; sp.upper[i] = 1.0; sp.fixed_dims[i] = 1; }
    return sp;
  }
public:
  enum { dimension = dim, dimensionworld = dim };
  static_assert(dim == 2 || dim == 3, "the bulk layer covers YaspGrid<2> and YaspGrid<3> (dgb_grid_create rejects other dimensions)");
  BulkYaspGridBase() = default;
  // EquidistantCoordinates (yaspgrid.hh:904-909)
  BulkYaspGridBase(const std::array<double,dim>& L, const std::array<int,dim>& s, std::bitset<dim> periodic = {}, int overlap = 1,
               int rank = 0, int nranks = 1, const Communicator* comm = nullptr) : comm_(comm ? comm->handle() : nullptr) {
    dgb_grid_spec sp = base(s, periodic, overlap); sp.coord_mode = DGB_COORD_EQUIDISTANT;
    for (int i = 0; i < dim; ++i) sp.upper[i] = L[i];
    create(sp, rank, nranks);
  }
  // EquidistantOffsetCoordinates (yaspgrid.hh:974-980)
  BulkYaspGridBase(const std::array<double,dim>& lowerleft, const std::array<double,dim>& upperright, const std::array<int,dim>& s,
               std::bitset<dim> periodic = {}, int overlap = 1, int rank = 0, int nranks = 1, const Communicator* comm = nullptr)
      : comm_(comm ? comm->handle() : nullptr) {
    dgb_grid_spec sp = base(s, periodic, overlap); sp.coord_mode = DGB_COORD_EQUIDISTANT_OFFSET;
    for (int i = 0; i < dim; ++i) { sp.lower[i] = lowerleft[i]; sp.upper[i] = upperright[i]; }
    create(sp, rank, nranks);
  }
  // TensorProductCoordinates (yaspgrid.hh:1043-1047)
  BulkYaspGridBase(const std::array<std::vector<double>,dim>& coords, std::bitset<dim> periodic = {}, int overlap = 1,
               int rank = 0, int nranks = 1, const Communicator* comm = nullptr) : comm_(comm ? comm->handle() : nullptr) {
    std::array<int,dim> s{};
    for (int i = 0; i < dim; ++i) s[i] = int(coords[i].size())-1;
    dgb_grid_spec sp = base(s, periodic, overlap); sp.coord_mode = DGB_COORD_TENSORPRODUCT;
    keep_.assign(coords.begin(), coords.end());
    for (int i = 0; i < dim; ++i) sp.tensor[i] = keep_[i].data();
    create(sp, rank, nranks);
  }
  BulkYaspGridBase(const BulkYaspGridBase&) = delete;
  ~BulkYaspGridBase() { if (g_) dgb_grid_destroy(g_); }
  // adopts a grid built by the C ABI (BackupRestoreFacility::restore)
  struct Adopt {};
  BulkYaspGridBase(Adopt, dgb_grid* g, const Communicator* comm) : g_(g), comm_(comm ? comm->handle() : nullptr) {}

  dgb_grid* handle() const { return g_; }
  dgb_comm* comm() const { return comm_; }
  void refineOptions(bool keepPhysicalOverlap) { check(dgb_grid_refine_options(g_, keepPhysicalOverlap)); }
  void globalRefine(int refCount) { check(dgb_grid_global_refine(g_, refCount)); }
  int maxLevel() const { int m; check(dgb_grid_max_level(g_, &m)); return m; }
  std::int64_t numBoundarySegments() const { std::int64_t n; check(dgb_grid_num_boundary_segments(g_, &n)); return n; }
  std::array<double,dim> domainSize() const { double L[3]; check(dgb_grid_domain_size(g_, L)); std::array<double,dim> r{}; for (int i = 0; i < dim; ++i) r[i] = L[i]; return r; }
  std::array<int,dim> torusDims() const { std::int32_t d[3]; check(dgb_grid_torus_dims(g_, d)); std::array<int,dim> r{}; for (int i = 0; i < dim; ++i) r[i] = d[i]; return r; }
  BulkGridView<dim> leafGridView() const { dgb_view v; check(dgb_view_leaf(g_, &v)); return BulkGridView<dim>(this, v); }
  BulkGridView<dim> levelGridView(int level) const { dgb_view v; check(dgb_view_level(g_, level, &v)); return BulkGridView<dim>(this, v); }
};

// Dune::YaspGrid<dim, Coordinates> (yaspgrid.hh:164): the coordinate container picks the constructor, as in the reference
template<int dim, class Coordinates = EquidistantCoordinates<double,dim>>
class BulkYaspGrid : public BulkYaspGridBase<dim> {
  typedef BulkYaspGridBase<dim> Base;
public:
  typedef Coordinates CoordinateContainer;
  typedef typename Coordinates::ctype ctype;
  static_assert(std::is_same<ctype, double>::value, "the device kernels compute in double precision");
  // yaspgrid.hh:904-909 - only for Coordinates = EquidistantCoordinates
  BulkYaspGrid(const std::array<double,dim>& L, const std::array<int,dim>& s, std::bitset<dim> periodic = {}, int overlap = 1,
               int rank = 0, int nranks = 1, const Communicator* comm = nullptr) : Base(L, s, periodic, overlap, rank, nranks, comm) {
    static_assert(Coordinates::mode == DGB_COORD_EQUIDISTANT, "this constructor belongs to YaspGrid<dim, EquidistantCoordinates>");
  }
  // yaspgrid.hh:974-980 - only for Coordinates = EquidistantOffsetCoordinates
  BulkYaspGrid(const std::array<double,dim>& lowerleft, const std::array<double,dim>& upperright, const std::array<int,dim>& s,
               std::bitset<dim> periodic = {}, int overlap = 1, int rank = 0, int nranks = 1, const Communicator* comm = nullptr)
      : Base(lowerleft, upperright, s, periodic, overlap, rank, nranks, comm) {
    static_assert(Coordinates::mode == DGB_COORD_EQUIDISTANT_OFFSET, "this constructor belongs to YaspGrid<dim, EquidistantOffsetCoordinates>");
  }
  // yaspgrid.hh:1043-1047 - only for Coordinates = TensorProductCoordinates
  explicit BulkYaspGrid(const std::array<std::vector<double>,dim>& coords, std::bitset<dim> periodic = {}, int overlap = 1,
                        int rank = 0, int nranks = 1, const Communicator* comm = nullptr) : Base(coords, periodic, overlap, rank, nranks, comm) {
    static_assert(Coordinates::mode == DGB_COORD_TENSORPRODUCT, "this constructor belongs to YaspGrid<dim, TensorProductCoordinates>");
  }
  BulkYaspGrid(typename Base::Adopt a, dgb_grid* g, const Communicator* comm) : Base(a, g, comm) {}
};

// Dune::BackupRestoreFacility<YaspGrid<dim,Coordinates>> (dune/grid/yaspgrid/backuprestore.hh): same static members and the
// same file format / file naming; the coordinate container of the grid type tells restore how to read the text.
template<class Grid> struct BackupRestoreFacility;
template<int dim, class Coordinates>
struct BackupRestoreFacility<BulkYaspGrid<dim,Coordinates>> {
  typedef BulkYaspGrid<dim,Coordinates> Grid;
  static Grid* restore(std::istream& stream, const Communicator* comm, int rank = 0, int nranks = 1) { return restore(stream, Coordinates::mode, rank, nranks, comm); }
  static Grid* restore(const std::string& filename, const Communicator* comm, int rank = 0, int nranks = 1) { return restore(filename, Coordinates::mode, rank, nranks, comm); }
  static std::string text(const Grid& grid) {
    std::int64_t need = 0; check(dgb_grid_backup(grid.handle(), nullptr, 0, &need));
    std::string t(std::size_t(need), '\0'); check(dgb_grid_backup(grid.handle(), &t[0], need, &need)); t.resize(std::size_t(need)-1);
    return t;
  }
  static void backup(const Grid& grid, std::ostream& stream) { stream << text(grid); }
  // :91-104 one file written by rank 0; tensor product :208-220 one file per rank, named filename + rank
  static void backup(const Grid& grid, const std::string& filename) {
    const dgb_view v = grid.leafGridView().descriptor();
    const bool tensor = v.coord_mode == DGB_COORD_TENSORPRODUCT;
    if (!tensor && v.rank != 0) return;
    std::ofstream file(tensor ? filename + std::to_string(v.rank) : filename);
    if (file) backup(grid, file);
    else std::cerr << "ERROR: BackupRestoreFacility::backup: couldn't open file `" << filename << "'" << std::endl;
  }
  static Grid* restore(std::istream& stream, int coordMode, int rank = 0, int nranks = 1, const Communicator* comm = nullptr) {
    const std::string t((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
    dgb_grid* g = nullptr; check(dgb_grid_restore(t.c_str(), coordMode, rank, nranks, &g));
    return new Grid(typename Grid::Adopt(), g, comm);
  }
  static Grid* restore(const std::string& filename, int coordMode, int rank = 0, int nranks = 1, const Communicator* comm = nullptr) {
    std::ifstream file(coordMode == DGB_COORD_TENSORPRODUCT ? filename + std::to_string(rank) : filename);
    if (file) return restore(file, coordMode, rank, nranks, comm);
    std::cerr << "ERROR: BackupRestoreFacility::restore: couldn't open file `" << filename << "'" << std::endl;
    return nullptr;
  }
};

// Dune::GlobalIndexSet<GridView>(gridview, codim) (dune/grid/utility/globalindexset.hh:318-497). The per-entity calls
// index(e) / subIndex(e,i,codim) become one device array: index()[e] for the entity with local index e of the codim.
template<int dim>
class GlobalIndexSet {
  DeviceArray<std::int32_t> idx_; std::int64_t n_ = 0; int codim_;
public:
  typedef int Index;
  GlobalIndexSet(const BulkGridView<dim>& gv, int codim, cudaStream_t s = nullptr) : idx_(std::size_t(gv.size(codim))), codim_(codim) {
    check(dgb_global_index(gv.grid().handle(), gv.level(), codim, idx_.data(), &n_, gv.grid().comm(), s));
  }
  const DeviceArray<std::int32_t>& index() const { return idx_; }
  unsigned int size(unsigned int codim) const { return int(codim) == codim_ ? unsigned(n_) : 0u; }   // :487-490
};

// Dune::VTKWriter<GridView>(gridView) with addCellData / addVertexData / write(name) (dune/grid/io/file/vtk/vtkwriter.hh:638,
// 698, 764, 985): the data are device arrays indexed by the element / vertex index, ncomp values per entity (P0 / P1 containers)
template<int dim>
class VTKWriter {
  const BulkGridView<dim>& gv_;
  struct Field { std::string name; const double* data; int ncomp; bool vector; };
  std::vector<Field> cell_, vertex_;
  static std::vector<dgb_vtk_field> pack(const std::vector<Field>& f) {
    std::vector<dgb_vtk_field> o;
    for (const Field& x : f) o.push_back(dgb_vtk_field{x.name.c_str(), x.data, x.ncomp, x.vector ? 1 : 0, 0});
    return o;
  }
public:
  explicit VTKWriter(const BulkGridView<dim>& gv) : gv_(gv) {}
  void addCellData(const double* d_values, const std::string& name, int ncomps = 1) { cell_.push_back({name, d_values, ncomps, ncomps > 1}); }
  void addVertexData(const double* d_values, const std::string& name, int ncomps = 1) { vertex_.push_back({name, d_values, ncomps, ncomps > 1}); }
  void clear() { cell_.clear(); vertex_.clear(); }
  // write(name, VTK::OutputType) (vtkwriter.hh:803-807); the enumerators carry the values of VTK::OutputType (common.hh:43-51)
  enum OutputType { ascii = DGB_VTK_ASCII, base64 = DGB_VTK_BASE64, appendedraw = DGB_VTK_APPENDEDRAW, appendedbase64 = DGB_VTK_APPENDEDBASE64 };
  std::string write(const std::string& name, OutputType type = ascii, const std::string& path = "", cudaStream_t s = nullptr) const {
    auto c = pack(cell_), v = pack(vertex_);
    char out[4096];
    check(dgb_vtk_write_ex(gv_.grid().handle(), gv_.level(), name.c_str(), path.c_str(), c.data(), int(c.size()), v.data(), int(v.size()), 0,
                           int(type), out, 4096, s));
    return out;
  }
  std::string write(const std::string& name, const std::string& path, cudaStream_t s = nullptr) const { return write(name, ascii, path, s); }
};

template<int dim>
template<class DataHandle>
void BulkGridView<dim>::communicate(DataHandle& dh, InterfaceType iftype, CommunicationDirection dir, cudaStream_t s) const {
  check(dgb_communicate(grid_->handle(), v_.level, &dh.mapper.descriptor(), iftype, dir, int(dh.op), dh.arrays.data(), dh.itemSizes.data(),
                        int(dh.arrays.size()), grid_->comm(), s));
}

template<int dim>
void BulkGridView<dim>::communicate(BulkVariableDataHandle<dim>& dh, InterfaceType iftype, CommunicationDirection dir, cudaStream_t s) const {
  dgb_commvar* h = nullptr;
  check(dgb_commvar_create(grid_->handle(), v_.level, &dh.mapper.descriptor(), iftype, dir, dh.sendOffsets, dh.sendData, s, &h));
  struct Guard { dgb_commvar* h; ~Guard() { cudaDeviceSynchronize(); dgb_commvar_destroy(h); } } guard{h};
  check(dgb_commvar_exchange(h, grid_->comm(), s));
  std::int64_t ne = 0, ni = 0;
  check(dgb_commvar_finish(h, s, &ne, &ni));
  dh.recvIndex = DeviceArray<std::uint32_t>(std::size_t(ne));
  dh.recvOffsets = DeviceArray<std::int64_t>(std::size_t(ne)+1);
  dh.recvData = DeviceArray<double>(std::size_t(ni));
  check(dgb_commvar_result(h, dh.recvIndex.data(), dh.recvOffsets.data(), dh.recvData.data(), s));
  cudaStreamSynchronize(s);
}

// the explicit upwind finite-volume transport functor (include/dgb.h)
template<int dim>
class FiniteVolume {
  dgb_fv* fv_ = nullptr;
public:
  enum Mode { exact = DGB_FV_EXACT, separable = DGB_FV_SEPARABLE };
  FiniteVolume(BulkYaspGridBase<dim>& grid, int level, int problem, const std::vector<double>& params, Mode mode = separable) {
    check(dgb_fv_create(grid.handle(), level, problem, params.data(), int(params.size()), mode, grid.comm(), &fv_));
  }
  // collective (like MPI_Win_create): the arrays the time loop alternates between; peers then store halo cells straight into them
  int registerFields(const std::vector<double*>& d_fields, cudaStream_t s = nullptr) {
    check(dgb_fv_register_fields(fv_, d_fields.data(), int(d_fields.size()), s));
    int n = 0; check(dgb_fv_registered_fields(fv_, &n)); return n;
  }
  void fence(cudaStream_t s = nullptr) { check(dgb_fv_fence(fv_, s)); }        // completes the halo of the last registered output array
  FiniteVolume(const FiniteVolume&) = delete;
  ~FiniteVolume() { if (fv_) dgb_fv_destroy(fv_); }
  void initial(double* d_c, cudaStream_t s = nullptr) { check(dgb_fv_init_field(fv_, d_c, s)); }
  void step(const double* d_c_in, double* d_c_out, cudaStream_t s = nullptr) { check(dgb_fv_step(fv_, d_c_in, d_c_out, s)); }
  double stepHost(const double* h_c_in, double* h_c_out, cudaStream_t s = nullptr) { double dt; check(dgb_fv_step_host(fv_, h_c_in, h_c_out, &dt, s)); return dt; }
  double lastDt(cudaStream_t s = nullptr) { double dt; check(dgb_fv_last_dt(fv_, &dt, s)); return dt; }
};

} // namespace DuneB200
#endif
