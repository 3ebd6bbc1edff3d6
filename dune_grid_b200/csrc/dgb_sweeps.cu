// dgb_sweeps.cu — bulk element / intersection / mapper / geometry sweeps.
// One launch per shift component of the requested codim; a thread owns one entity (i0 fastest, so a warp
// writes 32 consecutive entries of every output array: fully coalesced stores, no loads except the optional
// tensor-product coordinate tables). All kernels are store-bandwidth bound: roofline = bytes written / HBM BW.
#include "dgb_internal.hh"
#include "dgb_device.cuh"
#include "dgb_geogrid.cuh"

namespace dgb {

static constexpr int BX = 128;      // threads per block along i0

struct SweepGeom {                   // where this launch walks and where it writes
  int comp;                          // component number in view.comp
  int kind;                          // DGB_BOX_* of the partition
  long long base;                    // output position of the component's first entity
  long long n;                       // total entities of the sweep (array stride of component-major outputs)
};

// a thread walks SW_ROWS consecutive rows (i1) of its column: 8x fewer, 8x longer-lived blocks than one entity per
// thread, which matters for the 4-byte-per-entity outputs (mapper indices), where block scheduling was the cost
// (ROWS = 8); kernels that write hundreds of bytes per entity keep one entity per thread (ROWS = 1)
template<int SW_ROWS>
__device__ __forceinline__ bool sweep_entity(const dgb_view& v, const SweepGeom& s, int row, int* coord, long long& pos) {
  const dgb_box& b = v.comp[s.comp].box[s.kind];
  const int i0 = blockIdx.x*blockDim.x + threadIdx.x;
  const int i1 = blockIdx.y*SW_ROWS + row, i2 = blockIdx.z;
  if (i0 >= b.size[0] || i1 >= b.size[1]) return false;
  coord[0] = b.origin[0] + i0;
  coord[1] = b.origin[1] + i1;
  coord[2] = (v.dim > 2) ? b.origin[2] + i2 : 0;
  pos = s.base + ((long long)i2*b.size[1] + i1)*b.size[0] + i0;
  return true;
}

static dim3 sweep_grid(const dgb_view& v, int comp, int kind, int SW_ROWS) {
  const dgb_box& b = v.comp[comp].box[kind];
  return dim3((b.size[0]+BX-1)/BX, (b.size[1]+SW_ROWS-1)/SW_ROWS, v.dim > 2 ? b.size[2] : 1);
}
static bool sweep_empty(const dgb_view& v, int comp, int kind) {
  const dgb_box& b = v.comp[comp].box[kind];
  for (int i = 0; i < v.dim; ++i) if (b.size[i] <= 0) return true;
  return false;
}

// ---- K1: entities ------------------------------------------------------------------------------------
struct EntityOut { uint32_t* index; uint8_t* ptype; int32_t* coord; double *lower, *upper, *center, *volume; uint32_t block, offset; };

__global__ void __launch_bounds__(BX) k_entities(const __grid_constant__ dgb_view v, const SweepGeom s, const EntityOut o) {
  for (int row = 0; row < 8; ++row) {
  int coord[3]; long long pos;
  if (!sweep_entity<8>(v, s, row, coord, pos)) continue;
  const dgb_component& c = v.comp[s.comp];
  if (o.index) o.index[pos] = entity_index(v, c, s.kind, coord)*o.block + o.offset;
  if (o.ptype) o.ptype[pos] = partition_type(v, c, coord);
  if (o.coord) for (int i = 0; i < v.dim; ++i) o.coord[i*s.n + pos] = coord[i];
  if (o.lower || o.upper || o.center || o.volume) {
    // yaspgridentity.hh:270-274 (no wrap for 0 < codim), :491-514 (codim 0 wraps), AxisAlignedCubeGeometry
    double vol = 1.0;
    for (int i = 0; i < v.dim; ++i) {
      const bool ext = c.shift >> i & 1u;
      double ll = vertex_coord(v, i, coord[i]);
      double ur = ext ? vertex_coord(v, i, coord[i]+1) : ll;
      if (c.codim == 0) { const double sh = periodic_shift(v, i, coord[i]); if (sh != 0.0) { ll += sh; ur += sh; } }
      if (o.lower) o.lower[i*s.n + pos] = ll;
      if (o.upper) o.upper[i*s.n + pos] = ur;
      if (o.center) o.center[i*s.n + pos] = 0.5*(ll+ur);
      if (ext) vol *= ur-ll;
    }
    if (o.volume) o.volume[pos] = vol;
  }
  }
}

// ---- K2: sub-entity indices of cells -------------------------------------------------------------------
__global__ void __launch_bounds__(BX) k_subindex(const __grid_constant__ dgb_view v, const SweepGeom s, const SubTable t,
                                                 int cc, uint32_t block, uint32_t offset, uint32_t* out) {
  for (int row = 0; row < 8; ++row) {
  int coord[3]; long long pos;
  if (!sweep_entity<8>(v, s, row, coord, pos)) continue;
  for (int i = 0; i < t.n; ++i) {
    // yaspgridentity.hh:764-776: coord += move, component = shiftmapping[shift], index in the overlapfront numbering
    int sc[3];
    for (int j = 0; j < 3; ++j) sc[j] = coord[j] + (t.move[i] >> j & 1);
    const dgb_component& c = v.comp[v.comp_begin[cc] + v.shiftmap[t.shift[i]]];
    out[(long long)i*s.n + pos] = entity_index(v, c, DGB_BOX_OVERLAPFRONT, sc)*block + offset;
  }
  }
}

// ---- K3: intersections -------------------------------------------------------------------------------
struct FaceOut { uint8_t *neighbor, *boundary; int32_t *outside, *bsi; double *center, *volume; };

__global__ void __launch_bounds__(BX) k_intersections(const __grid_constant__ dgb_view v, const SweepGeom s, const FaceOut o) {
  for (int row = 0; row < 1; ++row) {
  int coord[3]; long long pos;
  if (!sweep_entity<1>(v, s, row, coord, pos)) continue;
  const dgb_component& c = v.comp[0];
  const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
  const int inside = int(entity_index(v, c, s.kind, coord));
  const bool geo = o.center || o.volume;
  double ll[3], ur[3];
  if (geo)
    for (int i = 0; i < v.dim; ++i) {
      const double sh = periodic_shift(v, i, coord[i]);
      ll[i] = vertex_coord(v, i, coord[i]); ur[i] = vertex_coord(v, i, coord[i]+1);
      if (sh != 0.0) { ll[i] += sh; ur[i] += sh; }
    }
  int stride = 1;
  for (int dir = 0; dir < v.dim; ++dir) {
    for (int side = 0; side < 2; ++side) {
      const int k = 2*dir + side;
      const bool nb = face_neighbor(v, coord, dir, side);
      const bool bd = face_boundary(v, coord, dir, side);
      if (o.neighbor) o.neighbor[k*s.n + pos] = nb;
      if (o.boundary) o.boundary[k*s.n + pos] = bd;
      // outside index = inside index -+ superincrement(dir) (yaspgridintersection.hh:40-57, ygrid.hh:362-366)
      if (o.outside) o.outside[k*s.n + pos] = nb ? inside + (side ? stride : -stride) : -1;
      if (o.bsi) o.bsi[k*s.n + pos] = bd ? boundary_segment_index(v, coord, dir, side) : -1;
      if (geo) {
        // yaspgridintersection.hh:232-269: the face is flat in `dir` at vertex coord+side, wrapped like the cell
        double vol = 1.0;
        for (int i = 0; i < v.dim; ++i) {
          double lo = ll[i], hi = ur[i];
          if (i == dir) { lo = side ? ur[i] : ll[i]; hi = lo; }
          else vol *= hi-lo;
          if (o.center) o.center[((long long)k*v.dim + i)*s.n + pos] = 0.5*(lo+hi);
        }
        if (o.volume) o.volume[k*s.n + pos] = vol;
      }
    }
    stride *= of.size[dir];
  }
  }
}

// ---- K4: axis-aligned geometry at quadrature points ---------------------------------------------------
struct QuadPoints { int n; double x[27*3]; };
struct GeoOut { double *global, *jit, *detJ; };

__global__ void __launch_bounds__(BX) k_geometry(const __grid_constant__ dgb_view v, const SweepGeom s, const QuadPoints q, const GeoOut o) {
  for (int row = 0; row < 1; ++row) {
  int coord[3]; long long pos;
  if (!sweep_entity<1>(v, s, row, coord, pos)) continue;
  const int dim = v.dim;
  double ll[3], w[3], vol = 1.0;
  for (int i = 0; i < dim; ++i) {
    const double sh = periodic_shift(v, i, coord[i]);
    double lo = vertex_coord(v, i, coord[i]), hi = vertex_coord(v, i, coord[i]+1);
    if (sh != 0.0) { lo += sh; hi += sh; }
    ll[i] = lo; w[i] = hi-lo; vol *= w[i];
  }
  for (int p = 0; p < q.n; ++p) {
    for (int i = 0; i < dim; ++i) {
      if (o.global) o.global[((long long)p*dim + i)*s.n + pos] = ll[i] + q.x[p*dim+i]*w[i];
      if (o.jit) for (int j = 0; j < dim; ++j) o.jit[(((long long)p*dim + i)*dim + j)*s.n + pos] = (i == j) ? 1.0/w[i] : 0.0;
    }
    if (o.detJ) o.detJ[(long long)p*s.n + pos] = vol;
  }
  }
}

// ---- K8: GeometryGrid (multilinear geometry over deformed corners) ------------------------------------
struct GeoGridOut { double *corners, *global, *jt, *jit, *detJ; };

template<int DIM>
__global__ void __launch_bounds__(BX) k_geogrid(const __grid_constant__ dgb_view v, const SweepGeom s, const QuadPoints q,
                                                const Deformation def, const GeoGridOut o) {
  for (int row = 0; row < 1; ++row) {
  int coord[3]; long long pos;
  if (!sweep_entity<1>(v, s, row, coord, pos)) continue;
  constexpr int NC = 1 << DIM;
  double ll[DIM], ur[DIM];
  for (int i = 0; i < DIM; ++i) {
    const double sh = periodic_shift(v, i, coord[i]);
    ll[i] = vertex_coord(v, i, coord[i]); ur[i] = vertex_coord(v, i, coord[i]+1);
    if (sh != 0.0) { ll[i] += sh; ur[i] += sh; }
  }
  double c[NC][DIM];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    double x[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = (k >> i & 1) ? ur[i] : ll[i];      // AxisAlignedCubeGeometry::corner(k)
    deform<DIM>(def, x, c[k]);
    if (o.corners)
#pragma unroll
      for (int i = 0; i < DIM; ++i) o.corners[((long long)k*DIM + i)*s.n + pos] = c[k][i];
  }
  for (int p = 0; p < q.n; ++p) {
    double x[DIM], y[DIM], J[DIM][DIM], Ji[DIM][DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = q.x[p*DIM+i];
    if (o.global) {
      ml_global<DIM>(c, x, y);
#pragma unroll
      for (int i = 0; i < DIM; ++i) o.global[((long long)p*DIM + i)*s.n + pos] = y[i];
    }
    ml_jt<DIM>(c, x, J);
    if (o.jt)
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) o.jt[(((long long)p*DIM + i)*DIM + j)*s.n + pos] = J[i][j];
    const double det = ml_det<DIM>(J);
    if (o.jit) {
      ml_inverse_transposed<DIM>(J, det, Ji);
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) o.jit[(((long long)p*DIM + i)*DIM + j)*s.n + pos] = Ji[i][j];
    }
    if (o.detJ) o.detJ[(long long)p*s.n + pos] = fabs(det);
  }
  }
}

// reduction variant: sum_q w_q |det J(x_q)| over all cells; block partial via warp shuffles, one atomicAdd per block
template<int DIM>
__global__ void __launch_bounds__(BX) k_geogrid_volume(const __grid_constant__ dgb_view v, const SweepGeom s, const QuadPoints q,
                                                       const QuadPoints wts, const Deformation def, double* sum) {
  double acc = 0.0;
  for (int row = 0; row < 8; ++row) {
  int coord[3]; long long pos;
  if (sweep_entity<8>(v, s, row, coord, pos)) {
    constexpr int NC = 1 << DIM;
    double ll[DIM], ur[DIM], c[NC][DIM];
    for (int i = 0; i < DIM; ++i) {
      const double sh = periodic_shift(v, i, coord[i]);
      ll[i] = vertex_coord(v, i, coord[i]); ur[i] = vertex_coord(v, i, coord[i]+1);
      if (sh != 0.0) { ll[i] += sh; ur[i] += sh; }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      double x[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) x[i] = (k >> i & 1) ? ur[i] : ll[i];
      deform<DIM>(def, x, c[k]);
    }
    for (int p = 0; p < q.n; ++p) {
      double x[DIM], J[DIM][DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) x[i] = q.x[p*DIM+i];
      ml_jt<DIM>(c, x, J);
      acc += wts.x[p]*fabs(ml_det<DIM>(J));
    }
  }
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  __shared__ double part[BX/32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < BX/32; ++w) t += part[w];
    atomicAdd(sum, t);
  }
}

// ---- host drivers ------------------------------------------------------------------------------------
template<class Launch>
static int for_each_component(const dgb_view& v, int codim, int pitype, int rows, Launch&& launch) {
  if (codim < 0 || codim > v.dim) return fail(DGB_EARG, "bad codim");
  const int kind = box_kind_of(pitype);
  if (kind == -2) return fail(DGB_EGRID, "YaspLevelIterator with this codim or partition type not implemented");
  if (kind == -1) return DGB_OK;                     // Ghost_Partition: empty range (yaspgrid.hh:1827)
  if (int rc = require_device()) return rc;
  SweepGeom s; s.kind = kind; s.base = 0; s.n = partition_size(v, codim, pitype);
  for (int j = v.comp_begin[codim]; j < v.comp_begin[codim+1]; ++j) {
    s.comp = j;
    if (!sweep_empty(v, j, kind)) launch(s, sweep_grid(v, j, kind, rows));
    long long t = 1; for (int i = 0; i < v.dim; ++i) t *= v.comp[j].box[kind].size[i];
    s.base += t;
  }
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

static int check_tensor(const dgb_view& v) {
  if (v.coord_mode == DGB_COORD_TENSORPRODUCT)
    for (int i = 0; i < v.dim; ++i) if (!v.tensor[i]) return fail(DGB_ECUDA, "tensor coordinates are not on the device (view created without a GPU?)");
  return DGB_OK;
}

static int fill_qp(const dgb_view& v, int nqp, const double* qp_host, QuadPoints& q) {
  if (nqp < 0 || nqp > 27) return fail(DGB_EARG, "at most 27 quadrature points per call");
  q.n = nqp;
  for (int i = 0; i < nqp*v.dim; ++i) q.x[i] = qp_host[i];
  return DGB_OK;
}

} // namespace dgb

using namespace dgb;

extern "C" {

int dgb_mapper_index(const dgb_view* v, const dgb_mapper* m, int codim, int pitype, uint32_t* d_out, void* stream) {
  if (codim < 0 || codim > v->dim) return fail(DGB_EARG, "bad codim");
  if (m->offset[codim] == DGB_INVALID_OFFSET) return fail(DGB_EARG, "codim not contained in the mapper layout");
  EntityOut o{}; o.index = d_out; o.block = m->block[codim]; o.offset = m->offset[codim];
  return for_each_component(*v, codim, pitype, 8, [&](const SweepGeom& s, dim3 grid) {
    k_entities<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o);
  });
}

int dgb_mapper_subindex(const dgb_view* v, const dgb_mapper* m, int cc, int pitype, uint32_t* d_out, void* stream) {
  if (cc < 0 || cc > v->dim) return fail(DGB_EARG, "bad codim");
  if (m->offset[cc] == DGB_INVALID_OFFSET) return fail(DGB_EARG, "codim not contained in the mapper layout");
  SubTable t; t.n = sub_entities(v->dim, cc);
  for (int i = 0; i < t.n; ++i) { t.shift[i] = (unsigned char)sub_shift(v->dim, i, cc); t.move[i] = (unsigned char)sub_move(v->dim, i, cc); }
  const uint32_t block = m->block[cc], offset = m->offset[cc];
  return for_each_component(*v, 0, pitype, 8, [&](const SweepGeom& s, dim3 grid) {
    k_subindex<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, t, cc, block, offset, d_out);
  });
}

int dgb_elements_materialize(const dgb_view* v, int codim, int pitype, uint32_t* d_index, uint8_t* d_ptype, int32_t* d_coord,
                             double* d_lower, double* d_upper, double* d_center, double* d_volume, void* stream) {
  if ((d_lower || d_upper || d_center || d_volume)) if (int rc = check_tensor(*v)) return rc;
  EntityOut o{d_index, d_ptype, d_coord, d_lower, d_upper, d_center, d_volume, 1u, 0u};
  return for_each_component(*v, codim, pitype, 8, [&](const SweepGeom& s, dim3 grid) {
    k_entities<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o);
  });
}

int dgb_intersections_materialize(const dgb_view* v, int pitype, uint8_t* d_neighbor, uint8_t* d_boundary, int32_t* d_outside,
                                  int32_t* d_bsi, double* d_center, double* d_volume, void* stream) {
  if (d_center || d_volume) if (int rc = check_tensor(*v)) return rc;
  FaceOut o{d_neighbor, d_boundary, d_outside, d_bsi, d_center, d_volume};
  return for_each_component(*v, 0, pitype, 1, [&](const SweepGeom& s, dim3 grid) {
    k_intersections<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o);
  });
}

int dgb_geometry_eval(const dgb_view* v, int pitype, int nqp, const double* qp_host, double* d_global, double* d_jit,
                      double* d_detJ, void* stream) {
  if (int rc = check_tensor(*v)) return rc;
  QuadPoints q; if (int rc = fill_qp(*v, nqp, qp_host, q)) return rc;
  GeoOut o{d_global, d_jit, d_detJ};
  return for_each_component(*v, 0, pitype, 1, [&](const SweepGeom& s, dim3 grid) {
    k_geometry<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, o);
  });
}

int dgb_geogrid_eval(const dgb_view* v, int pitype, int deformation, const double* params_host, int nparams, int nqp,
                     const double* qp_host, double* d_corners, double* d_global, double* d_jt, double* d_jit, double* d_detJ, void* stream) {
  if (int rc = check_tensor(*v)) return rc;
  QuadPoints q; if (int rc = fill_qp(*v, nqp, qp_host, q)) return rc;
  Deformation def; def.id = deformation;
  for (int i = 0; i < 4; ++i) def.p[i] = (params_host && i < nparams) ? params_host[i] : 0.0;
  if (deformation < 0 || deformation > 1) return fail(DGB_EARG, "unknown deformation id");
  GeoGridOut o{d_corners, d_global, d_jt, d_jit, d_detJ};
  return for_each_component(*v, 0, pitype, 1, [&](const SweepGeom& s, dim3 grid) {
    if (v->dim == 3) k_geogrid<3><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, def, o);
    else k_geogrid<2><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, def, o);
  });
}

int dgb_geogrid_volume(const dgb_view* v, int pitype, int deformation, const double* params_host, int nparams, int nqp,
                       const double* qp_host, const double* weights_host, double* d_sum, void* stream) {
  if (int rc = check_tensor(*v)) return rc;
  QuadPoints q; if (int rc = fill_qp(*v, nqp, qp_host, q)) return rc;
  QuadPoints w; w.n = nqp; for (int i = 0; i < nqp; ++i) w.x[i] = weights_host[i];
  Deformation def; def.id = deformation;
  for (int i = 0; i < 4; ++i) def.p[i] = (params_host && i < nparams) ? params_host[i] : 0.0;
  if (deformation < 0 || deformation > 1) return fail(DGB_EARG, "unknown deformation id");
  if (int rc = require_device()) return rc;
  DGB_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), (cudaStream_t)stream));
  return for_each_component(*v, 0, pitype, 8, [&](const SweepGeom& s, dim3 grid) {
    if (v->dim == 3) k_geogrid_volume<3><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, w, def, d_sum);
    else k_geogrid_volume<2><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, w, def, d_sum);
  });
}

} // extern "C"
