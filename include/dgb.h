/* dgb.h — C ABI of libdgb.so, the B200-native bulk implementation of dune-grid's YaspGrid leaf-sweep
 * hot path. Plain C types only; every data buffer is caller-owned DEVICE memory unless a
 * parameter says "host". All functions return DGB_OK (0) or a negative error code; the message is
 * available from dgb_last_error() (thread-local). Work is enqueued on the caller's CUDA stream
 * (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream) and functions return
 * before the device has finished, except where a host result is returned.
 *
 * There is NO CPU fallback: any entry point that needs the device fails with DGB_ECUDA when no CUDA
 * device is usable. Descriptor-only queries (create/refine/view/mapper_init/sizes/lists) are pure host
 * arithmetic and work without a GPU.
 *
 * The reference has no FFI for this path (it is a C++ template facade); each entry point below cites
 * the reference interface it replaces (paths relative to the dune-grid source tree).
 */
#ifndef DGB_H
#define DGB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DGB_MAX_DIM 3
#define DGB_MAX_COMP 8              /* 2^dim shift components (yaspgrid.hh:199-206) */

enum { DGB_OK = 0,
       DGB_EGRID = -1,              /* Dune::GridError   (yaspgrid.hh:788,1053,1219; partitioning.hh:66,137) */
       DGB_ERANGE = -2,             /* Dune::RangeError  (yaspgrid.hh:1745,1817) */
       DGB_EARG = -3,               /* Dune::Exception   (torus.hh:87; partitioning.hh:159) / bad argument */
       DGB_ECUDA = -4, DGB_ENCCL = -5, DGB_ENOTIMPL = -6 };

/* coordinate containers: dune/grid/yaspgrid/coordinates.hh:27,129,243 */
enum { DGB_COORD_EQUIDISTANT = 0, DGB_COORD_EQUIDISTANT_OFFSET = 1, DGB_COORD_TENSORPRODUCT = 2 };
/* partitioners: dune/grid/yaspgrid/partitioning.hh:45,122,145 */
enum { DGB_PART_DEFAULT = 0, DGB_PART_POWER_D = 1, DGB_PART_FIXED = 2 };
/* dune/grid/common/gridenums.hh — identical numeric values */
enum { DGB_InteriorEntity = 0, DGB_BorderEntity = 1, DGB_OverlapEntity = 2, DGB_FrontEntity = 3, DGB_GhostEntity = 4 };
enum { DGB_InteriorBorder_InteriorBorder_Interface = 0, DGB_InteriorBorder_All_Interface = 1,
       DGB_Overlap_OverlapFront_Interface = 2, DGB_Overlap_All_Interface = 3, DGB_All_All_Interface = 4 };
enum { DGB_Interior_Partition = 0, DGB_InteriorBorder_Partition = 1, DGB_Overlap_Partition = 2,
       DGB_OverlapFront_Partition = 3, DGB_All_Partition = 4, DGB_Ghost_Partition = 5 };
enum { DGB_ForwardCommunication = 0, DGB_BackwardCommunication = 1 };
enum { DGB_OP_SET = 0, DGB_OP_ADD = 1,         /* dune/python/grid/mapper.hh:117-137 commOp */
       DGB_OP_MIN_NONNEG = 2,                  /* if (remote >= 0) local = min(remote, local): MinimumExchange, utility/globalindexset.hh:153-160 */
       DGB_OP_SET_NONNEG = 3 };                /* if (remote >= 0) local = remote: IndexExchange, utility/globalindexset.hh:262-289 */
/* index of the four nested partition boxes of a component (yaspgrid.hh:199-206) */
enum { DGB_BOX_OVERLAPFRONT = 0, DGB_BOX_OVERLAP = 1, DGB_BOX_INTERIORBORDER = 2, DGB_BOX_INTERIOR = 3 };

const char* dgb_last_error(void);
const char* dgb_version(void);
int dgb_device_count(int* n);                       /* DGB_ECUDA if the CUDA runtime cannot see a device */

/* ---- grid life cycle: YaspGrid constructors yaspgrid.hh:904-909 (equidistant), :974-980 (offset), :1043-1047
 *      (tensor product); refineOptions :1270; globalRefine :1216 -------------------------------------------- */
typedef struct dgb_grid_spec {
  int32_t dim;                      /* 2 or 3 */
  int32_t coord_mode;               /* DGB_COORD_* */
  double  lower[DGB_MAX_DIM];       /* offset mode: lower left corner */
  double  upper[DGB_MAX_DIM];       /* equidistant: L; offset: upper right corner */
  int32_t N[DGB_MAX_DIM];           /* coarse cells per direction */
  const double* tensor[DGB_MAX_DIM];/* HOST pointers, N[i]+1 monotone coordinates (tensor mode) */
  uint32_t periodic;                /* bit i: direction i periodic */
  int32_t overlap;
  int32_t partitioner;              /* DGB_PART_* */
  int32_t fixed_dims[DGB_MAX_DIM];  /* DGB_PART_FIXED */
} dgb_grid_spec;

typedef struct dgb_grid dgb_grid;
int dgb_grid_create(const dgb_grid_spec* spec, int rank, int nranks, dgb_grid** out);
int dgb_grid_destroy(dgb_grid* g);
int dgb_grid_refine_options(dgb_grid* g, int keepPhysicalOverlap);
int dgb_grid_global_refine(dgb_grid* g, int refCount);
int dgb_grid_max_level(const dgb_grid* g, int* maxLevel);
int dgb_grid_torus_dims(const dgb_grid* g, int32_t* dims);             /* torus.hh:112 */
int dgb_grid_num_boundary_segments(const dgb_grid* g, int64_t* n);     /* yaspgrid.hh:1456 */
int dgb_grid_domain_size(const dgb_grid* g, double* L);                /* yaspgrid.hh:1462 */
/* BackupRestoreFacility<YaspGrid<dim,Coordinates>> (yaspgrid/backuprestore.hh, format version 2).
 * dgb_grid_backup: the text backup(grid, std::ostream&) writes (:105-128 equidistant / offset, :222-246 tensor product -
 * there every rank writes its own file holding its LOCAL coordinates); *needed = length incl. the terminating NUL, the
 * text is copied when cap >= *needed. Numbers are formatted like the reference's operator<< (6 significant digits).
 * dgb_grid_restore: restore(std::istream&, comm) (:146-218, :268-346): FixedSizePartitioning with the stored torus,
 * then one refineOptions(keep_l) + globalRefine(1) per stored level. `text` is this rank's file. */
int dgb_grid_backup(const dgb_grid* g, char* text, int64_t cap, int64_t* needed);
int dgb_grid_restore(const char* text, int coord_mode, int rank, int nranks, dgb_grid** out);
/* Dune::GlobalIndexSet<GridView>(gridview, codim) (dune/grid/utility/globalindexset.hh:318-440): globally unique consecutive
 * indices of the entities of one codimension. d_out[e] = index(entity with local index e) for all size(codim) entities of
 * the level (int32, like the reference's Index); *global_size = size(codim). Same algorithm, as device kernels: owner =
 * smallest rank that holds the entity as interior or border (communicate All_All with DGB_OP_MIN_NONNEG), owned entities
 * numbered in the order the reference meets them (cells in index order, sub-entities in reference-element order) by a
 * device-wide prefix sum, rank offsets from an all-gather of the counts, non-owners served by communicate All_All with
 * DGB_OP_SET_NONNEG. Collective over the ranks of `comm` (NULL on one rank); synchronises the stream. */
int dgb_global_index(dgb_grid* g, int level, int codim, int32_t* d_out, int64_t* global_size, struct dgb_comm* comm, void* stream);
/* VTKWriter<GridView>(gridView, VTK::conforming) with addCellData / addVertexData and write(name, VTK::ascii)
 * (dune/grid/io/file/vtk/vtkwriter.hh:638-642, 698, 764, 985-1030): cells of the InteriorBorder partition, vertices numbered at
 * their first visit, sections PointData / CellData / Points / Cells, vectors padded to 3 components. Fields are DEVICE arrays
 * indexed by the element index (cell fields) or the vertex index (vertex fields) of the level, ncomp doubles per entity, like the
 * containers of P0VTKFunction / P1VTKFunction. One rank: <path>/<name>.vtu; several ranks: every rank writes
 * s####-p####-<name>.vtu and rank 0 the header s####-<name>.pvtu (:858-909). `written` receives the file to open (piece or
 * header). Parity of this component is unpinned (the upstream writer cannot be built without dune-common). */
typedef struct dgb_vtk_field {
  const char* name;
  const double* d_data;
  int32_t ncomp;                    /* FieldInfo::size() */
  int32_t is_vector;                /* FieldInfo::Type::vector (written with 3 components) or scalar */
  int32_t precision;                /* 0 VTK::Precision::float32 (the reference's default), 1 float64 */
} dgb_vtk_field;
int dgb_vtk_write(const dgb_grid* g, int level, const char* name, const char* path, const dgb_vtk_field* cell_fields,
                  int ncell_fields, const dgb_vtk_field* vertex_fields, int nvertex_fields, int coord_precision,
                  char* written, int cap, void* stream);
/* the same with the output type of VTKWriter::write(name, type): VTK::OutputType, dune/grid/io/file/vtk/common.hh:43-51 (same values) -
 * inline base64, appended raw binary, appended base64; array syntax of dataarraywriter.hh:196-420. The syntax of all four types is
 * pinned byte for byte by the reference's own VTUWriter / DataArrayWriter classes (compiled into the test oracle). */
enum { DGB_VTK_ASCII = 0, DGB_VTK_BASE64 = 1, DGB_VTK_APPENDEDRAW = 2, DGB_VTK_APPENDEDBASE64 = 3 };
int dgb_vtk_write_ex(const dgb_grid* g, int level, const char* name, const char* path, const dgb_vtk_field* cell_fields,
                     int ncell_fields, const dgb_vtk_field* vertex_fields, int nvertex_fields, int coord_precision,
                     int outputtype, char* written, int cap, void* stream);
/* standalone processor-grid choice, partitioning.hh:56-117 */
int dgb_partition(int dim, const int32_t* N, int P, int overlap, int partitioner, const int32_t* fixed, int32_t* dims_out);

/* ---- views: Grid::leafGridView/levelGridView (common/grid.hh:878-891). A view is a POD descriptor that is
 *      copied by value into kernel parameters; leaf == level(maxLevel) for Yasp (yaspgrid.hh:1356-1381). ---- */
typedef struct dgb_box { int32_t origin[DGB_MAX_DIM]; int32_t size[DGB_MAX_DIM]; } dgb_box;
typedef struct dgb_component {
  uint32_t shift;                   /* bit i set <=> entities extend along axis i (ygrid.hh:155-165) */
  int32_t  codim;
  int32_t  index_offset[4];         /* per partition box kind: YGrid::_indexOffset of this component (ygrid.hh:771-791) */
  dgb_box  box[4];                  /* DGB_BOX_*; index arithmetic always uses box[0] as the enclosing array */
} dgb_component;
typedef struct dgb_view {
  int32_t dim, level, rank, nranks;
  int32_t N[DGB_MAX_DIM];           /* cells of the whole level (yaspgrid.hh:264) */
  uint32_t periodic;
  int32_t overlap;                  /* overlap width on this level */
  int32_t ncomp;                    /* 2^dim */
  int32_t comp_begin[DGB_MAX_DIM+2];/* components of codim c are comp[comp_begin[c] .. comp_begin[c+1]) in ascending shift */
  int32_t shiftmap[DGB_MAX_COMP];   /* shift bits -> component number inside its codim (ygrid.hh:567) */
  dgb_component comp[DGB_MAX_COMP];
  /* coordinates (coordinates.hh): vertex coordinate along axis d at GLOBAL index i */
  int32_t coord_mode;
  double  h[DGB_MAX_DIM];           /* equidistant modes */
  double  origin[DGB_MAX_DIM];      /* offset mode */
  const double* tensor[DGB_MAX_DIM];/* DEVICE pointers (tensor mode), entry 0 = global index tensor_offset[d] */
  const double* tensor_host[DGB_MAX_DIM];
  int32_t tensor_offset[DGB_MAX_DIM];
  int32_t tensor_size[DGB_MAX_DIM]; /* number of stored coordinates */
  double  L[DGB_MAX_DIM];           /* domain size, periodic wrap shift (yaspgridentity.hh:497-510) */
  /* boundary segments are numbered on the level-0 local box (yaspgridintersection.hh:105-162) */
  int32_t bs_origin[DGB_MAX_DIM], bs_size[DGB_MAX_DIM], bs_sides[DGB_MAX_DIM];
  int32_t N0[DGB_MAX_DIM];
} dgb_view;

int dgb_view_leaf(const dgb_grid* g, dgb_view* v);
int dgb_view_level(const dgb_grid* g, int level, dgb_view* v);        /* DGB_ERANGE like yaspgrid.hh:1745 */
int dgb_view_size(const dgb_view* v, int codim, int64_t* n);          /* GridView::size(codim), yaspgrid.hh:1425 */
int dgb_view_partition_size(const dgb_view* v, int codim, int pitype, int64_t* n);  /* length of begin<cd,pit>..end */
int dgb_view_overlap_size(const dgb_view* v, int codim, int* n);      /* gridview.hh overlapSize, yaspgrid.hh:1399 */
int dgb_view_ghost_size(const dgb_view* v, int codim, int* n);        /* always 0, yaspgrid.hh:1413 */
/* send/recv box lists of communicate (yaspgrid.hh:208-226,561-671). which: 0 send_of_of 1 recv_of_of 2 send_o_of
 * 3 recv_of_o 4 send_ib_ib 5 recv_ib_ib 6 send_ib_of 7 recv_of_ib. Entry = 10 ints: rank, distance, shift,
 * indexOffset, origin[3], size[3]. Host arithmetic only — no messages are exchanged to build them. */
int dgb_view_comm_list(const dgb_grid* g, int level, int codim, int which, int32_t* out, int cap, int* count);

/* ---- mapper: MultipleCodimMultipleGeomTypeMapper (common/mcmgmapper.hh:212-249,373-407). block[codim] replaces
 *      MCMGLayout (Yasp has one geometry type per codim, yaspgridindexsets.hh:90-93). ------------------------- */
typedef struct dgb_mapper { int32_t dim; uint32_t block[4]; uint32_t offset[4]; uint32_t size; } dgb_mapper;
#define DGB_INVALID_OFFSET 0xffffffffu
int dgb_mapper_init(const dgb_view* v, const uint32_t* block, dgb_mapper* m);
int dgb_mapper_size(const dgb_mapper* m, int64_t* n);
/* mapper.index(e) for every entity of begin<codim,pitype>()..end() in iteration order -> d_out[n] */
int dgb_mapper_index(const dgb_view* v, const dgb_mapper* m, int codim, int pitype, uint32_t* d_out, void* stream);
/* mapper.subIndex(e,i,cc) for every cell of the partition -> d_out[i*n + cell] (sub-entity major) */
int dgb_mapper_subindex(const dgb_view* v, const dgb_mapper* m, int cc, int pitype, uint32_t* d_out, void* stream);

/* Field checkpoint for restarts (not in the reference, which leaves user data to the user; the grid part of a restart is
 * dgb_grid_backup/restore above): file `filename`.`rank` = one header line, this rank's grid backup text, then the raw
 * doubles of a DEVICE array indexed by `mapper` (item_size doubles per index). Restore checks that header and grid text
 * match the grid and mapper it is given (DGB_EARG otherwise) and copies the data back to the device. Both synchronise. */
int dgb_field_backup(const dgb_grid* g, int level, const dgb_mapper* m, const double* d_field, int item_size,
                     const char* filename, void* stream);
int dgb_field_restore(const dgb_grid* g, int level, const dgb_mapper* m, double* d_field, int item_size,
                      const char* filename, void* stream);

/* ---- element sweep (rangegenerators.hh:881 elements(gv); yaspgridentity.hh:270-305,480-514; AxisAlignedCubeGeometry).
 *      All outputs are optional (NULL = skip), in iteration order of the partition; vector quantities are stored
 *      component-major: out[d*n + e]. index = indexSet.index(e). ------------------------------------------------ */
int dgb_elements_materialize(const dgb_view* v, int codim, int pitype,
                             uint32_t* d_index, uint8_t* d_ptype, int32_t* d_coord,
                             double* d_lower, double* d_upper, double* d_center, double* d_volume, void* stream);
/* ---- intersection sweep (rangegenerators.hh:845 intersections(gv,e); yaspgridintersection.hh:40-288). Per cell of the
 *      partition and face k in reference order (dir=k/2, side=k%2): out[k*n + cell]; face centre out[(k*dim+d)*n + cell].
 *      outside = indexSet.index(outside()) if neighbor() else -1; bsi = boundarySegmentIndex() if boundary() else -1. */
int dgb_intersections_materialize(const dgb_view* v, int pitype,
                                  uint8_t* d_neighbor, uint8_t* d_boundary, int32_t* d_outside, int32_t* d_bsi,
                                  double* d_center, double* d_volume, void* stream);
/* YaspIntersection::unitOuterNormal / outerNormal / centerUnitOuterNormal (= +-e_dir) and integrationOuterNormal (= face volume
 * times it) for every cell of the partition and face k: out[(k*dim+d)*n + cell] (yaspgridintersection.hh:163-190); NULL = skip */
int dgb_intersection_normals(const dgb_view* v, int pitype, double* d_unitOuterNormal, double* d_integrationOuterNormal, void* stream);
/* geometryInInside() (in_outside = 0) / geometryInOutside() (1) of face k: lower and upper corner of the unit face in the
 * element's local coordinates - the same for every cell of a YaspGrid (yaspgridintersection.hh:195-228). Host arithmetic. */
int dgb_intersection_local_geometry(int dim, int k, int in_outside, double* lower, double* upper);
/* ---- Geometry::global / jacobianInverseTransposed / integrationElement at nqp local points (host array nqp*dim) for all
 *      cells of the partition: global[(q*dim+d)*n+e], jit[((q*dim+r)*dim+c)*n+e], detJ[q*n+e] (common/geometry.hh:194-371) */
int dgb_geometry_eval(const dgb_view* v, int pitype, int nqp, const double* qp_host,
                      double* d_global, double* d_jit, double* d_detJ, void* stream);
/* the same for the entities of ANY codimension (faces, edges, vertices at quadrature points): AxisAlignedCubeGeometry<mydim,dim>,
 * mydim = dim-codim, local points qp_host (nqp*mydim): global[(q*dim+d)*n+e], jacobianTransposed jt[((q*mydim+r)*dim+c)*n+e],
 * jacobianInverseTransposed jit[((q*dim+r)*mydim+c)*n+e], integrationElement detJ[q*n+e]; entities in the iteration order of the
 * partition (components of the codim in ascending shift). yaspgridentity.hh:270-274, 849-851; NULL = skip. */
int dgb_geometry_eval_codim(const dgb_view* v, int codim, int pitype, int nqp, const double* qp_host,
                            double* d_global, double* d_jt, double* d_jit, double* d_detJ, void* stream);
/* ---- GeometryGrid over the Yasp host view with a compiled-in analytic coordinate function
 *      (geometrygrid/cornerstorage.hh:54-60, coordfunctioncaller.hh:40-43, geometry.hh:197-204). deformation: 0 identity,
 *      1 the sine field of SURVEY.md §8d config 5 (amplitude params_host[0]). corners[(k*dim+d)*n+e], jt like jit. */
int dgb_geogrid_eval(const dgb_view* v, int pitype, int deformation, const double* params_host, int nparams,
                     int nqp, const double* qp_host, double* d_corners, double* d_global, double* d_jt, double* d_jit,
                     double* d_detJ, void* stream);
/* GeoGrid::Intersection of GeometryGrid over the Yasp host view (geometrygrid/intersection.hh:103-172, cornerstorage.hh:122-165)
 * for every cell of the partition and its 2*dim faces in reference order (k: dir = k/2, side = k%2), NULL = skip:
 *   corners[((k*2^(dim-1)+i)*dim+d)*n+e]  geometry().corner(i) = insideGeo.global(hostLocalGeometry.corner(i))
 *   center[(k*dim+d)*n+e], volume[k*n+e]  geometry().center() / .volume() of the multilinear face geometry
 *   at the nqp (<= 16) face-local points qp_face_host (nqp x (dim-1)), out[((k*nqp+q)*dim+d)*n+e]:
 *     integrationOuterNormal = J^-T(x) n_ref |det J(x)|, outerNormal = J^-T(x) n_ref, unitOuterNormal, x = geometryInInside().global(local)
 *   centerUnitOuterNormal[(k*dim+d)*n+e]  = unitOuterNormal(face centre) */
int dgb_geogrid_intersections(const dgb_view* v, int pitype, int deformation, const double* params_host, int nparams,
                              int nqp, const double* qp_face_host, double* d_corners, double* d_center, double* d_volume,
                              double* d_integrationOuterNormal, double* d_outerNormal, double* d_unitOuterNormal,
                              double* d_centerUnitOuterNormal, void* stream);
/* reduction variant: sum over cells and points of weight_q * detJ -> *d_sum (one double, device) */
int dgb_geogrid_volume(const dgb_view* v, int pitype, int deformation, const double* params_host, int nparams,
                       int nqp, const double* qp_host, const double* weights_host, double* d_sum, void* stream);

/* ---- functor sweeps: `for (e : elements(gv, pitype)) sum += body(e)` and `for (e : ...) for (is : intersections(gv, e)) sum +=
 *      body(is)` (rangegenerators.hh:845,881) as ONE fused device kernel each (sweep + functor + reduction). The loop bodies are
 *      compiled-in device functors selected by id - structs in csrc/dgb_functors.cuh with element(ctx) / face(ctx) methods that see
 *      what the reference's entity / geometry / intersection facades offer; adding one = adding a struct there and listing it in
 *      the registry macro. No host callbacks, no CPU fallback. Shipped: the three loops of doc/recipes/recipe-integration.cc.
 *      d_per_element[e] / d_per_face[k*n + e] (iteration order of the partition; NULL = not stored) receive the bodies' values,
 *      *d_sum (device double; NULL = no reduction) their sum in a fixed, reproducible order. args_host: <= 16 doubles. ---------- */
enum { DGB_FUNCTOR_MIDPOINT_EXP_NORM = 0,      /* recipe-integration.cc:90-98   u(center)*volume, u(x) = exp(|x|) */
       DGB_FUNCTOR_GAUSS5_EXP_NORM = 1,        /* recipe-integration.cc:100-111 order-5 Gauss rule of the same u */
       DGB_FUNCTOR_BOUNDARY_FLUX_X = 2,        /* recipe-integration.cc:113-124 faces without neighbour: center . normal * volume */
       DGB_FUNCTOR_MEASURE = 3 };              /* args[0] * volume of the element / of the face */
int dgb_functor_count(int* n);
int dgb_functor_info(int functor_id, const char** name, int* has_element_body, int* has_intersection_body);
int dgb_for_each_element(const dgb_view* v, int pitype, int functor_id, const double* args_host, int nargs,
                         double* d_per_element, double* d_sum, void* stream);
int dgb_for_each_intersection(const dgb_view* v, int pitype, int functor_id, const double* args_host, int nargs,
                              double* d_per_face, double* d_sum, void* stream);

/* ---- communication: GridView::communicate(DataHandle, InterfaceType, CommunicationDirection) (gridview.hh:334,
 *      yaspgrid.hh:1470-1730) for flat arrays indexed by a mapper, the shape of the reference's own
 *      NumPyCommDataHandle (dune/python/grid/numpycommdatahandle.hh:41-142). Collective over all ranks. ---------- */
typedef struct dgb_comm dgb_comm;
int dgb_comm_unique_id(void* id128);                                   /* 128-byte ncclUniqueId, rank 0 */
int dgb_comm_init(const void* id128, int rank, int nranks, dgb_comm** out);
int dgb_comm_destroy(dgb_comm* c);
int dgb_communicate(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, int op,
                    double* const* d_arrays, const int32_t* item_sizes, int narrays, dgb_comm* comm, void* stream);
/* The three phases of dgb_communicate as separate calls, for transports other than NCCL (e.g. the reference's MPI
 * Torus) and for tests that emulate several ranks on one device. Buffers are owned by the library.
 * plan_info: out[0]=#send segments, out[1]=#recv segments, then per segment (peer, offset, count) in doubles, send
 * segments first; cap = capacity of out in int64. pack fills the send buffer (returned in *d_sendbuf); unpack scatters
 * the recv buffer (*d_recvbuf must have been filled by the transport). */
int dgb_comm_plan_info(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir,
                       const int32_t* item_sizes, int narrays, int64_t* out, int cap);
int dgb_comm_pack(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, double* const* d_arrays,
                  const int32_t* item_sizes, int narrays, double** d_sendbuf, double** d_recvbuf, void* stream);
int dgb_comm_unpack(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, int op, double* const* d_arrays,
                    const int32_t* item_sizes, int narrays, void* stream);
/* Data handles of VARIABLE size: CommDataHandleIF::fixedSize(dim, codim) == false (common/datahandleif.hh:97-125;
 * YaspGrid::communicateCodim, yaspgrid.hh:1562-1634: "variable size case: sender side determines the size"; the reference's own
 * checkCommunication uses it, test/checkcommunicate.hh:96-102). Bulk form of the handle:
 *   size(e)            = d_send_offsets[i+1] - d_send_offsets[i],  i = mapper.index(e)   (CSR, mapper.size()+1 entries)
 *   gather(buff, e)    = the items d_send_data[d_send_offsets[i] ...]
 *   scatter(buff, e, n): the calls the reference would make are RETURNED, in its order (codim dim..0, receive boxes in list
 *                        order, entities x fastest): d_index[k] = mapper.index(e) of call k, its n = d_offsets[k+1]-d_offsets[k]
 *                        items at d_data[d_offsets[k] ...]. What the handle does with them (set, add, append) is the caller's.
 * The mapper layout must have block 0 or 1 per codim (contains(dim, codim) = block != 0). As in the reference there are two
 * message rounds: per-entity sizes, then the items. Collective. Life cycle:
 *   dgb_commvar_create (local: sizes, offsets and packed items of every send box) -> dgb_commvar_exchange (both rounds over
 *   NCCL; device copies for messages to the own rank) -> dgb_commvar_finish (counts) -> dgb_commvar_result -> dgb_commvar_destroy.
 * Other transports (the reference's MPI torus; tests that emulate several ranks on one device) replace dgb_commvar_exchange:
 *   move the words of d_send_sizes segment by segment into the partners' d_recv_sizes (dgb_commvar_segments: out[0] = number of
 *   segments, then per segment peer, first entity slot, entity slots, first item, items; receive side: the item columns are
 *   valid after dgb_commvar_sizes_received), call dgb_commvar_sizes_received, move the items the same way, then finish. */
typedef struct dgb_commvar dgb_commvar;
int dgb_commvar_create(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, const int64_t* d_send_offsets,
                       const double* d_send_data, void* stream, dgb_commvar** out);
int dgb_commvar_exchange(dgb_commvar* h, dgb_comm* comm, void* stream);
int dgb_commvar_segments(dgb_commvar* h, int recv_side, int64_t* out, int cap);
int dgb_commvar_buffers(dgb_commvar* h, int64_t** d_send_sizes, double** d_send_items, int64_t** d_recv_sizes, double** d_recv_items);
int dgb_commvar_sizes_received(dgb_commvar* h, void* stream);
int dgb_commvar_finish(dgb_commvar* h, void* stream, int64_t* n_entities, int64_t* n_items);
int dgb_commvar_result(dgb_commvar* h, uint32_t* d_index, int64_t* d_offsets, double* d_data, void* stream);
int dgb_commvar_destroy(dgb_commvar* h);
int dgb_allreduce_min(double* d_value, dgb_comm* comm, void* stream);  /* comm().min(x), yaspgrid.hh:1316 pattern */
int dgb_allreduce_max(double* d_value, dgb_comm* comm, void* stream);
/* bytes this rank sent in the most recent dgb_communicate / fv step exchange (for NVLink GB/s reporting) */
int dgb_comm_last_bytes(const dgb_grid* g, int64_t* bytes);

/* ---- the finite-volume transport functor: element loop + intersection loop + mapper +
 *      geometry + CFL min-reduction + communicate(InteriorBorder_All_Interface, ForwardCommunication). -------------- */
typedef struct dgb_fv dgb_fv;
enum { DGB_FV_EXACT = 0,           /* per-cell arithmetic in the reference order: bit-identical to the CPU oracle */
       DGB_FV_SEPARABLE = 1 };     /* tensor-product factorised face factors: <= 1e-13 relative */
/* `problem` selects a compiled-in transport problem (velocity, inflow value, initial value: the FV registry of
 * csrc/dgb_functors.cuh): 0 constant velocity, 1 rotation in the x-y plane, 2 shear flow with a z-dependent velocity. The
 * production kernel k_fv_ring evaluates the x/y face factors once per COLUMN of cells where the problem allows it: a velocity that
 * does not depend on z and a constant inflow value (`column_invariant` in the registry; problems 0 and 1). Any other problem runs
 * on the same kernel in its general mode, which evaluates the functor at every face of every plane (same results as the z-march
 * kernel k_fv_march, which remains the kernel of DGB_FV_EXACT). All problems must be time independent: the CFL bound
 * reduced while step n runs is the one step n+2 uses. */
int dgb_fv_create(dgb_grid* g, int level, int problem, const double* params_host, int nparams, int mode,
                  dgb_comm* comm, dgb_fv** out);
/* dgb_fv_destroy is COLLECTIVE on a multi-rank grid whose steps used the fused exchange (like MPI_Win_free): every rank closes
 * its mappings of the peers' exchange blocks, all ranks meet in a barrier on `comm` (which must still be alive), then the blocks
 * are freed. dgb_fv_abandon is the non-collective form (e.g. from a garbage collector): it frees everything local and
 * deliberately leaks the exported exchange block, because freeing memory another process may still have mapped is undefined. */
int dgb_fv_destroy(dgb_fv* fv);
int dgb_fv_abandon(dgb_fv* fv);
int dgb_fv_init_field(dgb_fv* fv, double* d_c, void* stream);          /* c0 at the cell centres, all stored cells */
/* one step: c_out = c_in + dt*update(c_in) on interior cells, halo of c_out refreshed, dt = 0.99*min(1/sum_out) */
int dgb_fv_step(dgb_fv* fv, const double* d_c_in, double* d_c_out, void* stream);
/* Registration of the arrays a time loop alternates between (collective over the ranks of the grid; the analogue of
 * MPI_Win_create / ncclCommRegister): every rank passes its `nfields` (<= 4) device arrays, indexed by the element mapper,
 * in the same order. Afterwards a dgb_fv_step whose d_c_out is one of them lets the peers' step kernels store the halo
 * cells STRAIGHT INTO that array over NVLink (CUDA IPC mapping): no receive buffer, no unpack launch. Contract: all ranks
 * pass the array of the same registration slot to the same step; d_c_in != d_c_out; while an array is registered, its
 * overlap cells may be overwritten by the peers as soon as this rank has entered the step that writes it. If the arrays
 * cannot be mapped on every rank (not cudaMalloc memory, no peer access) nothing is registered on any rank and the steps
 * keep the receive-buffer form - same results. The arrays must stay allocated until dgb_fv_destroy. */
int dgb_fv_register_fields(dgb_fv* fv, double* const* d_fields, int nfields, void* stream);
int dgb_fv_registered_fields(const dgb_fv* fv, int* n);                /* 0 = the registration was declined */
/* Completion of the halo of a REGISTERED output array. With overlap 1 and the production kernel, a step that writes a registered
 * array delivers the x faces (one cell per row of the receiver: every 8-byte store would be a read-modify-write of a DRAM
 * sector) into the receiver's contiguous exchange buffer instead of into the array; the next step reads its x neighbours from
 * that buffer, so in a time loop the overlap COLUMN of the array is never written. dgb_fv_fence(fv, stream) unpacks it for
 * the most recent output array - call it before anything but dgb_fv_step reads that array's overlap cells (the analogue of
 * MPI_Win_fence). dgb_fv_last_dt, dgb_fv_step_host, dgb_fv_update, dgb_fv_step_local and a dgb_fv_step whose input is another
 * array do it themselves. Arrays that were not registered are always complete when dgb_fv_step returns (stream order).
 * DGB_FV_HYBRID=0 in the environment keeps the x faces going into the array as well (DIRECT). Not collective. */
int dgb_fv_fence(dgb_fv* fv, void* stream);
/* the kernel of dgb_fv_step alone: no halo exchange, no reduction across ranks. *d_pending_max then points at the
 * device double holding this rank's max outflow sum, which the caller must max-reduce over ranks before the next
 * step (dgb_fv_step does both itself). For transports other than NCCL and single-device rank emulation. */
int dgb_fv_step_local(dgb_fv* fv, const double* d_c_in, double* d_c_out, double** d_pending_max, void* stream);
int dgb_fv_bootstrap_local(dgb_fv* fv, double** d_pending_max, void* stream);
int dgb_fv_last_dt(dgb_fv* fv, double* dt_host, void* stream);         /* synchronises the stream */
/* optional: also materialise the update array (tests) */
int dgb_fv_update(dgb_fv* fv, const double* d_c_in, double* d_update, void* stream);
/* the same step on HOST buffers: H2D copy of c, step, D2H copy of the result (incl. refreshed overlap cells). 3-D grids
 * with pinned buffers are pipelined in z-slabs (upload / kernel / download of different slabs run concurrently, the halo
 * exchange rides on the slab kernels and is unpacked straight into h_c_out); pageable buffers take the serial form.
 * Collective on a multi-rank grid, synchronises the stream. */
int dgb_fv_step_host(dgb_fv* fv, const double* h_c_in, double* h_c_out, double* dt_host, void* stream);
/* how the last dgb_fv_step_host moved the field: */
#define DGB_FV_HOST_NONE 0
#define DGB_FV_HOST_SERIAL 1        /* H2D of the whole field, step, D2H of the whole field */
#define DGB_FV_HOST_PIPELINED 2     /* z-slabs: upload, step kernel and download of different slabs run concurrently on three streams */
int dgb_fv_host_path(const dgb_fv* fv, int* path);
int dgb_fv_launch_count(const dgb_fv* fv, int64_t* kernels);           /* kernels launched by this object so far */
/* which kernel the last step launched for the (bulk of the) interior cells: info8 = { kind, tuning variant, tile width, tile
 * height, bytes per staging copy (16 = paired cp.async, 8 = per cell, 5280 = one TMA box per plane tile, 0 = not a ring kernel),
 * z planes per CTA, 1 if the launch packed the send boxes into the peers' memory (fused exchange), staging form of the ring
 * kernel (0 8-byte cp.async, 1 aligned 16-byte pairs, 2 phase-aware 16-byte pairs, 3 tensor map / TMA) } */
enum { DGB_FV_KERNEL_NONE = 0, DGB_FV_KERNEL_STEP2D = 1, DGB_FV_KERNEL_MARCH = 2, DGB_FV_KERNEL_RING = 3, DGB_FV_KERNEL_RING_BULK = 4 };
int dgb_fv_kernel_info(const dgb_fv* fv, int32_t* info8);
/* host arithmetic only (no device needed): the z planes one CTA of the 3-D ring kernel (128 x 4 tile) marches for a region of
 * nx x ny x nz interior cells - packs: the launch also packs the send boxes; tensor_map: the tensor-map staging form. Chunks of
 * equal length leave a ragged last wave on the 148 SMs, so the rules take m long chunks and a short last one. */
int dgb_fv_chunk_rule(int nx, int ny, int nz, int packs, int tensor_map, int* zchunk);
/* how the last dgb_fv_step refreshed the overlap cells: */
#define DGB_FV_EXCHANGE_NONE 0      /* no step yet, or nothing to exchange */
#define DGB_FV_EXCHANGE_SERIAL 1    /* step kernel, then communicate() (2-D, or no room for a shell/bulk split) */
#define DGB_FV_EXCHANGE_OVERLAPPED 2 /* shell kernel + NCCL send/recv on a second stream, concurrent with the bulk kernel */
#define DGB_FV_EXCHANGE_FUSED 3     /* one step kernel stores into the receivers' buffers over peer-mapped memory and signals */
#define DGB_FV_EXCHANGE_DIRECT 4    /* as FUSED, and the stores go straight into the receivers' registered arrays: no unpack */
#define DGB_FV_EXCHANGE_HYBRID 5    /* as DIRECT, except the x-face boxes: they stay in the receiver's contiguous exchange buffer, where
                                       the receiver's next step kernel reads its x neighbours; see dgb_fv_fence */
/* Limits of the fused form (beyond them dgb_fv_step keeps the NCCL forms, same results): at most 64 ranks (control block),
 * at most 32 send boxes per rank (26 in 3-D with overlap; more only with several codims, which the FV exchange does not
 * have). dgb_communicate takes at most 8 arrays per call. A peer that does not deliver its step within 20 s makes the
 * waiting kernel give up WITHOUT unpacking; the time-out is sticky: every later dgb_fv_step / dgb_fv_step_host / dgb_fv_last_dt
 * of the object returns DGB_ENCCL and launches nothing. */
int dgb_fv_exchange_mode(const dgb_fv* fv, int* mode);

#ifdef __cplusplus
}
#endif
#endif /* DGB_H */
