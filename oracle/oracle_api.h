/* oracle_api.h — C interface shared by the two CPU oracles of this repo.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load these
 * libraries, and only as the checker or the timed CPU baseline.
 *
 * Two libraries export exactly these symbols:
 *   oracle/_ref/liboracle_ref.so   the reference's own dune/grid/yaspgrid.hh, dune/grid/yaspgrid/ headers and
 *                                  dune/grid/common/mcmgmapper.hh compiled UNMODIFIED from /root/reference against
 *                                  the stand-in headers in oracle/ref/shim (dune-common / dune-geometry / MPI are not
 *                                  available here); ranks are threads (oracle/ref/shim/mpi.h).
 *   oracle/liboracle_port.so       a self-contained C++ restatement of the same algorithms (oracle/port/), which does
 *                                  not need /root/reference and is what travels in git.
 * Pinning: everything integer (partitioning, boxes, lists, indices, partition types, intersections' topology, communicate,
 * GlobalIndexSet, BackupRestore text) is pinned by the reference itself - the _ref library IS its code, the port is checked
 * against it and against the golden vectors generated from it (tests/golden/). PARITY UNPINNED for the floating-point
 * geometry (AxisAlignedCubeGeometry, MultiLinearGeometry, FieldMatrix inverse): that arithmetic lives in dune-geometry /
 * dune-common (Depends: dune-geometry >= 2.12, no commit pinned), which are not vendored in the reference tree and not in this
 * image; both libraries use the stand-ins of oracle/ref/shim and oracle/common, restated from the published algorithms.
 * Every function simulates ALL ranks of a run inside one process ("world").
 *
 * Conventions: entity arrays are in the reference's iteration order for the requested
 * PartitionIteratorType (x fastest, components of a codim in ascending shift value); any output
 * pointer may be NULL (= skip). Enum values are those of dune/grid/common/gridenums.hh.
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_spec {
  int dim;                 /* 2 or 3 */
  int coord_mode;          /* 0 EquidistantCoordinates (L = upper), 1 EquidistantOffsetCoordinates, 2 TensorProductCoordinates */
  double lower[3];
  double upper[3];
  int N[3];                /* coarse cells per direction (yaspgrid.hh:904-909) */
  const double* tensor[3]; /* coord_mode 2: N[i]+1 monotone coordinates per axis (yaspgrid.hh:1043) */
  unsigned periodic;       /* bit i = direction i periodic */
  int overlap;
  int partitioner;         /* 0 DefaultPartitioning, 1 PowerDPartitioning, 2 FixedSizePartitioning (partitioning.hh) */
  int fixed_dims[3];
  int refine;              /* number of globalRefine levels added after construction (yaspgrid.hh:1216) */
  int keep_overlap;        /* refineOptions(keepPhysicalOverlap) (yaspgrid.hh:1270) */
} orc_spec;

const char* orc_kind(void);            /* "reference" or "port" */
const char* orc_last_error(void);

/* a1: processor grid only (partitioning.hh:56-117); returns 0 or -1 (GridError) */
int orc_partition(int dim, const int* N, int P, int overlap, int partitioner, const int* fixed, int* dims_out);
/* KATs of test-yaspgrid-entityshifttable.cc: entityShift / entityMove bitsets and subEnt counts */
int orc_entity_shift(int dim, int i, int cc);
int orc_entity_move(int dim, int i, int cc);
int orc_sub_entities(int dim, int d, int c);

void* orc_create(const orc_spec* spec, int nranks);   /* NULL + orc_last_error() on GridError/Exception */
void  orc_destroy(void* w);
int   orc_max_level(void* w);
void  orc_torus_dims(void* w, int* dims);
void  orc_domain_size(void* w, double* L);
int64_t orc_size(void* w, int rank, int level, int codim);          /* yaspgrid.hh:1425 */
int   orc_num_boundary_segments(void* w, int rank);                 /* yaspgrid.hh:1456 */
int   orc_overlap_size(void* w, int level);

/* a2: nested boxes. out: per component 26 ints = shift, indexOffset(overlapfront), then 4 boxes
 * (overlapfront, overlap, interiorborder, interior) × (origin[3], size[3]). Returns #components. */
int orc_boxes(void* w, int rank, int level, int codim, int* out);
/* a2/a10: send/recv lists. which: 0 send_of_of 1 recv_of_of 2 send_o_of 3 recv_of_o 4 send_ib_ib 5 recv_ib_ib
 * 6 send_ib_of 7 recv_of_ib. Entry = 10 ints: rank, distance, shift, indexOffset, origin[3], size[3]. Returns #entries. */
int orc_list(void* w, int rank, int level, int codim, int which, int* out, int cap);
/* a3: coordinate(axis, i) for i = first .. first+count-1 where [first,first+count) is the vertex range of the stored box */
int orc_coordinates(void* w, int rank, int level, int axis, double* out, int* first);

/* a4,a5,a8: entity sweep. Returns the entity count. coord: n×dim ints; lower/upper/center: n×dim; */
int64_t orc_entities(void* w, int rank, int level, int codim, int pitype,
                     uint32_t* index, uint8_t* ptype, int32_t* coord,
                     double* lower, double* upper, double* center, double* volume);
/* a5: indexSet.subIndex(e,i,cc) for all cells of the partition, n × subEntities(cc) */
int64_t orc_subindices(void* w, int rank, int level, int pitype, int cc, uint32_t* out);
/* a6: MultipleCodimMultipleGeomTypeMapper with layout(gt) = block[dim - gt.dim()] */
int64_t orc_mapper_size(void* w, int rank, int level, const uint32_t* block);
int64_t orc_mapper_index(void* w, int rank, int level, const uint32_t* block, int codim, int pitype, uint32_t* out);
int64_t orc_mapper_subindex(void* w, int rank, int level, const uint32_t* block, int pitype, int cc, uint32_t* out);
/* a7: intersections of all cells of the partition; per cell 2*dim entries in reference order.
 * outside_index = indexSet.index(outside()) if neighbor() else -1; bsi = boundarySegmentIndex() if boundary() else -1;
 * normal, face_center: n×2dim×dim; face_volume: n×2dim. */
int64_t orc_intersections(void* w, int rank, int level, int pitype,
                          uint8_t* neighbor, uint8_t* boundary, int32_t* outside_index, int32_t* bsi,
                          double* normal, double* face_center, double* face_volume);
/* a8: geometry at quadrature points for codim-0 cells: global n×nqp×dim, jit n×nqp×dim×dim (row major), detJ n×nqp */
int64_t orc_geometry_eval(void* w, int rank, int level, int pitype, int nqp, const double* qp,
                          double* global, double* jit, double* detJ);
/* a8 for any codimension: e.geometry() of the entities of `codim` (AxisAlignedCubeGeometry<dim-codim, dim>; only cells are wrapped
 * periodically, yaspgridentity.hh:270-274) at nqp local points qp (nqp x mydim, mydim = dim-codim): global n x nqp x dim,
 * jt (jacobianTransposed) n x nqp x mydim x dim, jit n x nqp x dim x mydim, detJ n x nqp. Returns the entity count. */
int64_t orc_geometry_eval_codim(void* w, int rank, int level, int codim, int pitype, int nqp, const double* qp,
                                double* global, double* jt, double* jit, double* detJ);
/* a9: GeometryGrid over the Yasp host level: corners n×2^dim×dim, then as above with jt too */
int64_t orc_geogrid_eval(void* w, int rank, int level, int pitype, int deformation, const double* params,
                         int nqp, const double* qp, double* corners, double* global, double* jt, double* jit, double* detJ);
/* a9, second half: GeoGrid::Intersection (geometrygrid/intersection.hh:103-172, cornerstorage.hh:122-165) for all cells of the
 * partition and their 2*dim faces in reference order. corners: n x 2dim x 2^(dim-1) x dim = insideGeo.global(hostLocalGeo.corner(i));
 * center, volume: n x 2dim (x dim) of the face geometry (MultiLinearGeometry<dim-1,dim> over those corners); at the nqp face-local
 * points qp (nqp x (dim-1)): integrationOuterNormal = J^-T(x) n_ref |det J(x)|, outerNormal = J^-T(x) n_ref, unitOuterNormal:
 * n x 2dim x nqp x dim; centerUnitOuterNormal: n x 2dim x dim. The reference headers for these rules are compiled into the
 * reference library; the multilinear arithmetic under them is the restatement of oracle/common/multilinear.hh (unpinned). */
int64_t orc_geogrid_intersections(void* w, int rank, int level, int pitype, int deformation, const double* params,
                                  int nqp, const double* qp_face, double* corners, double* center, double* volume,
                                  double* integrationOuterNormal, double* outerNormal, double* unitOuterNormal,
                                  double* centerUnitOuterNormal);
/* a10: communicate on all ranks at once. arrays[rank*narrays + a] is indexed by the mapper (block) with item_sizes[a]
 * doubles per mapper entry. op 0 = set (local = remote), 1 = add (local = local + remote). Returns doubles sent in total. */
int64_t orc_communicate(void* w, int level, const uint32_t* block, int iftype, int dir, int op,
                        int narrays, double** arrays, const int* item_sizes);
/* a10 with a data handle of VARIABLE size (CommDataHandleIF::fixedSize == false; yaspgrid.hh:1562-1634), on all ranks at once.
 * The handle: contains(dim, codim) = block[codim] != 0 (block 0 or 1), size(e) = offsets[rank][i+1] - offsets[rank][i] and
 * gather(e) = data[rank][offsets[rank][i] ...] with i = mapper.index(e). Every scatter(buff, e, n) call is RECORDED in call
 * order: recv_index[rank][k] = mapper.index(e), recv_n[rank][k] = n, the n items appended to recv_data[rank]. n_calls[rank] /
 * n_items[rank] always receive the counts; the three output tables may be NULL (counting pass). Returns the items sent in
 * total, or -1. */
int64_t orc_communicate_variable(void* w, int level, const uint32_t* block, int iftype, int dir,
                                 const int64_t* const* offsets, const double* const* data, int64_t* n_calls, int64_t* n_items,
                                 uint32_t** recv_index, int64_t** recv_n, double** recv_data);
/* a11: FV transport. problem 0: constant velocity params[0..2], inflow value params[3], initial
 * shell centre params[4..6], radii params[7..8]; problem 1: rotation about (0.5,0.5) with angular speed params[0],
 * inflow params[3], same initial condition. c[rank] is indexed by the element index (All partition). */
int orc_fv_init(void* w, int rank, int level, int problem, const double* params, double* c);
/* one step on all ranks (one thread per rank): returns dt (global min · 0.99). If upd != NULL the per-rank update
 * arrays are also returned (interior cells; 0 elsewhere). */
double orc_fv_step(void* w, int level, int problem, const double* params, double** c, double** upd);
/* the same step repeated `steps` times with rank threads kept alive; returns seconds of wall time for the loop */
double orc_fv_run(void* w, int level, int problem, const double* params, double** c, int steps, double* dt_last);

/* (f2) Dune::GlobalIndexSet<LevelGridView>(gridview, codim) (dune/grid/utility/globalindexset.hh:318-440) built on all ranks
 * at once. out[rank] receives size(level, codim) ints: the global index of the entity with local index e, read back
 * through subIndex(cell, i, codim) / index(cell). Returns size(codim), the number of global entities, or -1.
 * orc_communicate additionally understands op 2 (if remote >= 0: local = min(remote, local), MinimumExchange :153-160)
 * and op 3 (if remote >= 0: local = remote, IndexExchange :262-289). */
int64_t orc_global_index(void* w, int level, int codim, int32_t** out);

/* (f3) BackupRestoreFacility<YaspGrid<dim,CC>> (yaspgrid/backuprestore.hh). orc_backup: the text backup(grid, std::ostream&)
 * writes for `rank` (:105-128, :222-246); returns its length without the NUL (copied when it fits into cap) or -1.
 * orc_restore: a world whose rank r is built by restore(std::istream&, comm) (:146-218, :268-346) from texts[r]. */
int   orc_backup(void* w, int rank, char* out, int cap);
void* orc_restore(int dim, int coord_mode, int nranks, const char* const* texts);

/* (f1) one .vtu piece written by the reference's own VTK::VTUWriter / DataArrayWriterFactory / Base64Stream / RawStream
 * (dune/grid/io/file/vtk/{vtuwriter,dataarraywriter,streams,b64enc,common}.hh, compiled unmodified into the REFERENCE library only; -1 from
 * the port) in the call sequence of VTKWriter::writeDataFile / writeAllData (vtkwriter.hh:1177-1217): PointData, CellData, Points, Cells,
 * then - appended output types - the same arrays again into the <AppendedData> section. outputtype: 0 ascii, 1 base64, 2 appendedraw,
 * 3 appendedbase64 (VTK::OutputType). Arrays: `narrays` entries in file order; section[a]: 0 PointData, 1 CellData, 2 Points, 3 Cells;
 * precision[a]: 0 Float32, 1 Float64, 2 Int32, 3 UInt8; values[a] holds ncomps[a]*nitems[a] doubles (converted like the writer's
 * write<T>()). scalars / vectors: the attribute strings of the two data sections ("" = none). Returns the byte count of the file;
 * the bytes are copied when cap is large enough. */
int64_t orc_vtu_piece(int outputtype, unsigned ncells, unsigned npoints, int narrays, const int* section, const char* const* names,
                      const int* ncomps, const int* nitems, const int* precision, const double* const* values,
                      const char* point_scalars, const char* point_vectors, const char* cell_scalars, const char* cell_vectors,
                      char* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
