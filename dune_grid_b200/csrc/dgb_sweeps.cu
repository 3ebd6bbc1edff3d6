// dgb_sweeps.cu — bulk element / intersection / mapper / geometry sweeps.
// One launch per shift component of the requested codim; a thread owns one entity (i0 fastest, so a warp
// writes 32 consecutive entries of every output array: fully coalesced stores, no loads except the optional
// tensor-product coordinate tables). All kernels are store-bandwidth bound: roofline = bytes written / HBM BW.
#include <cstdlib>
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

// A thread owns one column (i0, i2) and walks ENT_ROWS rows (i1); axes 0 and 2 (coordinates, table loads, periodic
// shifts) are evaluated once per column, position and consecutive index advance by constant strides.
static constexpr int ENT_ROWS = 8;

__global__ void __launch_bounds__(BX) k_entities(const __grid_constant__ dgb_view v, const SweepGeom s, const EntityOut o) {
  const dgb_component& c = v.comp[s.comp];
  const dgb_box& b = c.box[s.kind];
  const int dim = v.dim;
  const int i0 = blockIdx.x*BX + threadIdx.x, i1 = blockIdx.y*ENT_ROWS, i2 = blockIdx.z;
  if (i0 >= b.size[0]) return;
  int coord[3] = { b.origin[0]+i0, b.origin[1]+i1, dim > 2 ? b.origin[2]+i2 : 0 };
  long long pos = s.base + ((long long)i2*b.size[1] + i1)*b.size[0] + i0;
  uint32_t idx = entity_index(v, c, s.kind, coord)*o.block + o.offset;
  const uint32_t didx = uint32_t(c.box[DGB_BOX_OVERLAPFRONT].size[0])*o.block;
  const bool geo = o.lower || o.upper || o.center || o.volume;
  const int rows = min(ENT_ROWS, b.size[1]-i1);
  double ll[3] = {0, 0, 0}, ur[3] = {0, 0, 0};
  // yaspgridentity.hh:270-274 (no wrap for 0 < codim), :491-514 (codim 0 wraps), AxisAlignedCubeGeometry.
  // ALL coordinate loads of the thread (tensor-product tables) are issued up front, before the first store: ncu showed the
  // graded 384^3 case stalled on the per-row table loads (long scoreboard), not on DRAM (profiles/r02_sweeps_ncu.json)
  double y1[ENT_ROWS+1];
  if (geo) {
#pragma unroll
    for (int r = 0; r <= ENT_ROWS; ++r) y1[r] = (dim > 1 && r <= rows) ? vertex_coord(v, 1, coord[1]+r) : 0.0;
  }
  auto axis = [&](int i) {
    const bool ext = c.shift >> i & 1u;
    ll[i] = vertex_coord(v, i, coord[i]);
    ur[i] = ext ? vertex_coord(v, i, coord[i]+1) : ll[i];
    if (c.codim == 0) { const double sh = periodic_shift(v, i, coord[i]); if (sh != 0.0) { ll[i] += sh; ur[i] += sh; } }
  };
  if (geo) { axis(0); if (dim > 2) axis(2); }
  const bool ext1 = c.shift >> 1 & 1u;
#pragma unroll
  for (int r = 0; r < ENT_ROWS; ++r) {
    if (r < rows) {
      if (o.index) o.index[pos] = idx;
      if (o.ptype) o.ptype[pos] = partition_type(v, c, coord);
      if (o.coord) { long long p = pos; for (int i = 0; i < dim; ++i) { o.coord[p] = coord[i]; p += s.n; } }
      if (geo) {
        if (dim > 1) {
          ll[1] = y1[r]; ur[1] = ext1 ? y1[r+1] : y1[r];
          if (c.codim == 0) { const double sh = periodic_shift(v, 1, coord[1]); if (sh != 0.0) { ll[1] += sh; ur[1] += sh; } }
        }
        double vol = 1.0;
        long long p = pos;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i < dim) {
            if (o.lower) o.lower[p] = ll[i];
            if (o.upper) o.upper[p] = ur[i];
            if (o.center) o.center[p] = 0.5*(ll[i]+ur[i]);
            if (c.shift >> i & 1u) vol *= ur[i]-ll[i];
            p += s.n;
          }
        }
        if (o.volume) o.volume[pos] = vol;
      }
      pos += b.size[0]; idx += didx; ++coord[1];
    }
  }
}

// ---- K1b / K2b: mapper indices, incremental form ------------------------------------------------------------------
// The 4-byte-per-entity outputs are integer-issue-bound when every entity recomputes its position and index from its
// coordinates (64-bit multiplies) - 16-byte stores did not help (measured). Here a thread owns one column (i0) and walks
// IDX_ROWS consecutive rows (i1): position and consecutive index are computed once, then advance by constant strides
// (row length of the swept box / of the enclosing overlapfront box), so an entity costs two adds and one coalesced store.
static constexpr int IDX_ROWS = 32;

template<int IDX_ROWS>
__global__ void __launch_bounds__(BX) k_index_rows(const __grid_constant__ dgb_view v, const SweepGeom s, uint32_t block, uint32_t offset,
                                                   uint32_t* __restrict__ out) {
  const dgb_component& c = v.comp[s.comp];
  const dgb_box& b = c.box[s.kind];
  const int i0 = blockIdx.x*BX + threadIdx.x, i1 = blockIdx.y*IDX_ROWS;
  if (i0 >= b.size[0]) return;
  const int coord[3] = { b.origin[0]+i0, b.origin[1]+i1, v.dim > 2 ? b.origin[2]+int(blockIdx.z) : 0 };
  uint32_t* p = out + (s.base + ((long long)blockIdx.z*b.size[1] + i1)*b.size[0] + i0);
  uint32_t val = entity_index(v, c, s.kind, coord)*block + offset;
  const uint32_t dval = uint32_t(c.box[DGB_BOX_OVERLAPFRONT].size[0])*block;     // next row of the overlapfront numbering
  const int rows = min(IDX_ROWS, b.size[1]-i1), dp = b.size[0];
  for (int r = 0; r < rows; ++r) { *p = val; p += dp; val += dval; }
}

__global__ void __launch_bounds__(BX) k_subindex_rows(const __grid_constant__ dgb_view v, const SweepGeom s, const SubTable t, int cc,
                                                      uint32_t block, uint32_t offset, uint32_t* __restrict__ out) {
  const dgb_box& b = v.comp[s.comp].box[s.kind];
  const int i0 = blockIdx.x*BX + threadIdx.x, i1 = blockIdx.y*IDX_ROWS;
  if (i0 >= b.size[0]) return;
  const int coord[3] = { b.origin[0]+i0, b.origin[1]+i1, v.dim > 2 ? b.origin[2]+int(blockIdx.z) : 0 };
  const long long pos = s.base + ((long long)blockIdx.z*b.size[1] + i1)*b.size[0] + i0;
  const int rows = min(IDX_ROWS, b.size[1]-i1), dp = b.size[0];
  for (int i = 0; i < t.n; ++i) {
    // yaspgridentity.hh:764-776: coord += move, component = shiftmapping[shift], index in the overlapfront numbering
    const dgb_component& c = v.comp[v.comp_begin[cc] + v.shiftmap[t.shift[i]]];
    const int sc[3] = { coord[0] + (t.move[i] & 1), coord[1] + (t.move[i] >> 1 & 1), coord[2] + (t.move[i] >> 2 & 1) };
    uint32_t val = entity_index(v, c, DGB_BOX_OVERLAPFRONT, sc)*block + offset;
    const uint32_t dval = uint32_t(c.box[DGB_BOX_OVERLAPFRONT].size[0])*block;
    uint32_t* p = out + ((long long)i*s.n + pos);
    for (int r = 0; r < rows; ++r) { *p = val; p += dp; val += dval; }
  }
}

// ---- K3: intersections -------------------------------------------------------------------------------
struct FaceOut { uint8_t *neighbor, *boundary; int32_t *outside, *bsi; double *center, *volume; };

// A thread owns one column (i0, i2) and walks FACE_ROWS rows (i1): everything that belongs to axes 0 and 2 (vertex
// coordinates incl. the tensor-product table loads, periodic shifts, neighbour / boundary flags) is computed once per
// column; per row only axis 1 is refreshed and the output positions advance by the row length.
static constexpr int FACE_ROWS_DEFAULT = 2;     // round 1, per-row loads (256^3, 48 output streams): 1: 0.76 ms, 2: 0.74, 4: 1.06, 8: 1.29, 16: 0.89

template<int FACE_ROWS>
__global__ void __launch_bounds__(BX) k_intersections(const __grid_constant__ dgb_view v, const SweepGeom s, const FaceOut o) {
  const dgb_component& c = v.comp[0];
  const dgb_box& b = c.box[s.kind];
  const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
  const int dim = v.dim;
  const int i0 = blockIdx.x*BX + threadIdx.x, i1 = blockIdx.y*FACE_ROWS, i2 = blockIdx.z;
  if (i0 >= b.size[0]) return;
  int coord[3] = { b.origin[0]+i0, b.origin[1]+i1, dim > 2 ? b.origin[2]+i2 : 0 };
  long long pos = s.base + ((long long)i2*b.size[1] + i1)*b.size[0] + i0;
  int inside = int(entity_index(v, c, s.kind, coord));
  const bool geo = o.center || o.volume;
  const int rows = min(FACE_ROWS, b.size[1]-i1);
  double ll[3] = {0, 0, 0}, ur[3] = {0, 0, 0};
  bool nb[6], bd[6];
  // every coordinate load of the thread (tensor-product tables: x, z and the rows+1 y vertices it walks) is issued before the
  // first store; ncu had the graded 384^3 sweep waiting on the per-row loads (long scoreboard 10.6 cycles per issue, DRAM 55 %)
  double y1[FACE_ROWS+1];
  if (geo) {
#pragma unroll
    for (int r = 0; r <= FACE_ROWS; ++r) y1[r] = (dim > 1 && r <= rows) ? vertex_coord(v, 1, coord[1]+r) : 0.0;
  }
  auto axis = [&](int i) {
    if (geo) {
      const double sh = periodic_shift(v, i, coord[i]);
      ll[i] = vertex_coord(v, i, coord[i]); ur[i] = vertex_coord(v, i, coord[i]+1);
      if (sh != 0.0) { ll[i] += sh; ur[i] += sh; }
    }
    nb[2*i] = face_neighbor(v, coord, i, 0); nb[2*i+1] = face_neighbor(v, coord, i, 1);
    bd[2*i] = face_boundary(v, coord, i, 0); bd[2*i+1] = face_boundary(v, coord, i, 1);
  };
  axis(0);
  if (dim > 2) axis(2);
  int stride[3] = { 1, of.size[0], of.size[0]*of.size[1] };
#pragma unroll
  for (int r = 0; r < FACE_ROWS; ++r) {
    if (r < rows) {
      if (dim > 1) {
        if (geo) {
          const double sh = periodic_shift(v, 1, coord[1]);
          ll[1] = y1[r]; ur[1] = y1[r+1];
          if (sh != 0.0) { ll[1] += sh; ur[1] += sh; }
        }
        nb[2] = face_neighbor(v, coord, 1, 0); nb[3] = face_neighbor(v, coord, 1, 1);
        bd[2] = face_boundary(v, coord, 1, 0); bd[3] = face_boundary(v, coord, 1, 1);
      }
      long long pf = pos;                       // position in the one-value-per-face arrays (face-major)
      long long pc = pos;                       // position in the face centre array ((face, axis)-major)
#pragma unroll
      for (int dir = 0; dir < 3; ++dir) {
        if (dir < dim) {
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            const int k = 2*dir + side;
            if (o.neighbor) o.neighbor[pf] = nb[k];
            if (o.boundary) o.boundary[pf] = bd[k];
            // outside index = inside index -+ superincrement(dir) (yaspgridintersection.hh:40-57, ygrid.hh:362-366)
            if (o.outside) o.outside[pf] = nb[k] ? inside + (side ? stride[dir] : -stride[dir]) : -1;
            if (o.bsi) o.bsi[pf] = bd[k] ? boundary_segment_index(v, coord, dir, side) : -1;
            if (geo) {
              // yaspgridintersection.hh:232-269: the face is flat in `dir` at vertex coord+side, wrapped like the cell
              double vol = 1.0;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                if (i < dim) {
                  double lo = ll[i], hi = ur[i];
                  if (i == dir) { lo = side ? ur[i] : ll[i]; hi = lo; }
                  else vol *= hi-lo;
                  if (o.center) o.center[pc] = 0.5*(lo+hi);
                  pc += s.n;
                }
              }
              if (o.volume) o.volume[pf] = vol;
            }
            pf += s.n;
          }
        }
      }
      pos += b.size[0]; inside += of.size[0]; ++coord[1];
    }
  }
}



// ---- linear walk: the form for boxes whose rows are NOT a multiple of 32 entities ---------------------------------------
// Measured (tools/intersections_probe.py, profiles/r02_intersections_probe.jsonl): the column form above reaches 90-93 % of
// the copy peak on 384^3 / 512^3 boxes but 64 % on the 386-wide box of an x-periodic grid (BASELINE config 4) - tensor-product
// coordinates cost nothing, a 382-cell periodic grid (384 stored) runs at 102 %. With 386-entity rows every row of every
// output stream starts at another offset inside its 32-byte sector / 128-byte line, so every warp store straddles one more
// line and leaves two partial sectors, and the fourth CTA of a row has 2 live lanes. Here the box is walked in its LINEAR
// order instead (which is the order of every output stream): a CTA owns LIN_ROWS*128 consecutive entities, thread t the
// entities chunk + j*128 + t, so every warp store is a whole, aligned run of 32 entities in all 48 streams whatever the row
// length. Coordinates follow by carry propagation (i0 += 128, wrap into i1, i2); per-axis values are refreshed only when
// their index changes.
static constexpr int LIN_ROWS = 8;

struct LinWalk {
  int i0, i1, i2;            // position inside the box
  int s0, s1;
  __device__ __forceinline__ void init(long long e, const dgb_box& b, int dim) {
    s0 = b.size[0]; s1 = dim > 1 ? b.size[1] : 1;
    const long long q = e / s0;
    i0 = int(e - q*s0);
    i2 = dim > 2 ? int(q / s1) : 0;
    i1 = int(q - (long long)i2*s1);
  }
  // advance by BX entities; returns a mask: bit 1 = i1 changed, bit 2 = i2 changed
  __device__ __forceinline__ int step() {
    int changed = 0;
    i0 += BX;
    while (i0 >= s0) { i0 -= s0; ++i1; changed |= 1; if (i1 == s1) { i1 = 0; ++i2; changed |= 2; } }
    return changed;
  }
};

__global__ void __launch_bounds__(BX) k_intersections_linear(const __grid_constant__ dgb_view v, const SweepGeom s, const FaceOut o, long long count) {
  const dgb_component& c = v.comp[0];
  const dgb_box& b = c.box[s.kind];
  const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
  const int dim = v.dim;
  long long e = (long long)blockIdx.x*(BX*LIN_ROWS) + threadIdx.x;
  if (e >= count) return;
  LinWalk w; w.init(e, b, dim);
  const bool geo = o.center || o.volume;
  const int stride[3] = { 1, of.size[0], of.size[0]*of.size[1] };
  double ll[3] = {0, 0, 0}, ur[3] = {0, 0, 0};
  bool nb[6], bd[6];
  int coord[3] = { b.origin[0]+w.i0, b.origin[1]+w.i1, dim > 2 ? b.origin[2]+w.i2 : 0 };
  auto axis = [&](int i) {
    if (geo) {
      const double sh = periodic_shift(v, i, coord[i]);
      ll[i] = vertex_coord(v, i, coord[i]); ur[i] = vertex_coord(v, i, coord[i]+1);
      if (sh != 0.0) { ll[i] += sh; ur[i] += sh; }
    }
    nb[2*i] = face_neighbor(v, coord, i, 0); nb[2*i+1] = face_neighbor(v, coord, i, 1);
    bd[2*i] = face_boundary(v, coord, i, 0); bd[2*i+1] = face_boundary(v, coord, i, 1);
  };
  axis(0); if (dim > 1) axis(1); if (dim > 2) axis(2);
  for (int j = 0; j < LIN_ROWS; ++j) {
    // the x coordinates of the NEXT entity are requested before this entity's 48 stores are issued: with tensor-product tables the
    // loop waited on these two loads every iteration (ncu: long scoreboard 13 cycles per issue, DRAM 72 %)
    const bool more = j+1 < LIN_ROWS && e + BX < count;
    int ch = 0;
    double nx0 = 0.0, nx1 = 0.0;
    if (more) {
      ch = w.step();
      if (geo) { nx0 = vertex_coord(v, 0, b.origin[0]+w.i0); nx1 = vertex_coord(v, 0, b.origin[0]+w.i0+1); }
    }
    const long long pos = s.base + e;
    const int inside = int(entity_index(v, c, s.kind, coord));
    long long pf = pos, pc = pos;
#pragma unroll
    for (int dir = 0; dir < 3; ++dir) {
      if (dir < dim) {
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int k = 2*dir + side;
          if (o.neighbor) o.neighbor[pf] = nb[k];
          if (o.boundary) o.boundary[pf] = bd[k];
          if (o.outside) o.outside[pf] = nb[k] ? inside + (side ? stride[dir] : -stride[dir]) : -1;
          if (o.bsi) o.bsi[pf] = bd[k] ? boundary_segment_index(v, coord, dir, side) : -1;
          if (geo) {
            double vol = 1.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              if (i < dim) {
                double lo = ll[i], hi = ur[i];
                if (i == dir) { lo = side ? ur[i] : ll[i]; hi = lo; }
                else vol *= hi-lo;
                if (o.center) o.center[pc] = 0.5*(lo+hi);
                pc += s.n;
              }
            }
            if (o.volume) o.volume[pf] = vol;
          }
          pf += s.n;
        }
      }
    }
    if (!more) break;
    e += BX;
    coord[0] = b.origin[0]+w.i0;
    if (geo) {
      const double sh = periodic_shift(v, 0, coord[0]);
      ll[0] = nx0; ur[0] = nx1;
      if (sh != 0.0) { ll[0] += sh; ur[0] += sh; }
    }
    nb[0] = face_neighbor(v, coord, 0, 0); nb[1] = face_neighbor(v, coord, 0, 1);
    bd[0] = face_boundary(v, coord, 0, 0); bd[1] = face_boundary(v, coord, 0, 1);
    if (ch & 1) { coord[1] = b.origin[1]+w.i1; axis(1); }
    if (ch & 2) { coord[2] = b.origin[2]+w.i2; axis(2); }
  }
}

// entities of one component in linear order (boxes whose rows are not a multiple of 32 entities): the form of k_intersections_linear
__global__ void __launch_bounds__(BX) k_entities_linear(const __grid_constant__ dgb_view v, const SweepGeom s, const EntityOut o, long long count) {
  const dgb_component& c = v.comp[s.comp];
  const dgb_box& b = c.box[s.kind];
  const int dim = v.dim;
  long long e = (long long)blockIdx.x*(BX*LIN_ROWS) + threadIdx.x;
  if (e >= count) return;
  LinWalk w; w.init(e, b, dim);
  const bool geo = o.lower || o.upper || o.center || o.volume;
  double ll[3] = {0, 0, 0}, ur[3] = {0, 0, 0};
  int coord[3] = { b.origin[0]+w.i0, b.origin[1]+w.i1, dim > 2 ? b.origin[2]+w.i2 : 0 };
  // yaspgridentity.hh:270-274 (no wrap for 0 < codim), :491-514 (codim 0 wraps), AxisAlignedCubeGeometry
  auto axis = [&](int i) {
    if (!geo) return;
    const bool ext = c.shift >> i & 1u;
    ll[i] = vertex_coord(v, i, coord[i]);
    ur[i] = ext ? vertex_coord(v, i, coord[i]+1) : ll[i];
    if (c.codim == 0) { const double sh = periodic_shift(v, i, coord[i]); if (sh != 0.0) { ll[i] += sh; ur[i] += sh; } }
  };
  axis(0); if (dim > 1) axis(1); if (dim > 2) axis(2);
  const bool ext0 = c.shift & 1u;
  for (int j = 0; j < LIN_ROWS; ++j) {
    // the x coordinates of the next entity are requested before this entity's stores (see k_intersections_linear)
    const bool more = j+1 < LIN_ROWS && e + BX < count;
    int ch = 0;
    double nx0 = 0.0, nx1 = 0.0;
    if (more) {
      ch = w.step();
      if (geo) { nx0 = vertex_coord(v, 0, b.origin[0]+w.i0); if (ext0) nx1 = vertex_coord(v, 0, b.origin[0]+w.i0+1); }
    }
    const long long pos = s.base + e;
    if (o.index) o.index[pos] = entity_index(v, c, s.kind, coord)*o.block + o.offset;
    if (o.ptype) o.ptype[pos] = partition_type(v, c, coord);
    if (o.coord) { long long p = pos; for (int i = 0; i < dim; ++i) { o.coord[p] = coord[i]; p += s.n; } }
    if (geo) {
      double vol = 1.0;
      long long p = pos;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < dim) {
          if (o.lower) o.lower[p] = ll[i];
          if (o.upper) o.upper[p] = ur[i];
          if (o.center) o.center[p] = 0.5*(ll[i]+ur[i]);
          if (c.shift >> i & 1u) vol *= ur[i]-ll[i];
          p += s.n;
        }
      }
      if (o.volume) o.volume[pos] = vol;
    }
    if (!more) break;
    e += BX;
    coord[0] = b.origin[0]+w.i0;
    if (geo) {
      ll[0] = nx0; ur[0] = ext0 ? nx1 : nx0;
      if (c.codim == 0) { const double sh = periodic_shift(v, 0, coord[0]); if (sh != 0.0) { ll[0] += sh; ur[0] += sh; } }
    }
    if (ch & 1) { coord[1] = b.origin[1]+w.i1; axis(1); }
    if (ch & 2) { coord[2] = b.origin[2]+w.i2; axis(2); }
  }
}

// mapper indices of one component in linear order (see above): two adds per entity, every warp store aligned
__global__ void __launch_bounds__(BX) k_index_linear(const __grid_constant__ dgb_view v, const SweepGeom s, uint32_t block, uint32_t offset,
                                                     uint32_t* __restrict__ out, long long count) {
  const dgb_component& c = v.comp[s.comp];
  const dgb_box& b = c.box[s.kind];
  const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
  constexpr int ROWS = 32;
  long long e = (long long)blockIdx.x*(BX*ROWS) + threadIdx.x;
  if (e >= count) return;
  LinWalk w; w.init(e, b, v.dim);
  const int coord[3] = { b.origin[0]+w.i0, b.origin[1]+w.i1, v.dim > 2 ? b.origin[2]+w.i2 : 0 };
  uint32_t val = entity_index(v, c, s.kind, coord)*block + offset;
  // moving by BX entities inside a row adds BX to the index; wrapping into the next row adds the gap between the rows of the
  // enclosing overlapfront numbering, wrapping into the next plane the gap between its planes
  const uint32_t rowGap = uint32_t(of.size[0]-b.size[0])*block, planeGap = uint32_t(of.size[0])*uint32_t(of.size[1]-b.size[1])*block;
  uint32_t* p = out + s.base + e;
  for (int j = 0; j < ROWS; ++j) {
    *p = val;
    e += BX; p += BX;
    if (e >= count) break;
    const int i1old = w.i1, i2old = w.i2;
    w.step();
    val += uint32_t(BX)*block;
    int rows = (w.i2 - i2old)*w.s1 + (w.i1 - i1old);        // rows advanced (may be > 1 for narrow boxes)
    val += uint32_t(rows)*rowGap + uint32_t(w.i2 - i2old)*planeGap;
  }
}

// ---- K1c / K2c: mapper indices and sub-indices, 16-byte stores, ONE launch per call -------------------------------------------
// Every stream of indices (a shift component of mapper.index, one sub-entity number of mapper.subIndex) is an affine function of
// the position inside the swept box: value = val0 + i0*block + i1*rowStride + i2*planeStride (numbering of the enclosing
// overlapfront array, ygrid.hh:472-479). A thread owns the 16-byte vectors g, g+BX, ... of the OUTPUT array (absolute positions,
// so any stream offset is aligned); the four values of a vector are consecutive unless the vector straddles a row end (carry
// walk) or the ends of the stream (scalar stores). The position advances by 4*BX entities with a carry; no division in the loop.
struct IdxSeg { long long base, count; int s0, s1; uint32_t val0, rowStride, planeStride; unsigned firstBlock; };
static constexpr int IDX_MAX_SEGS = 12;
struct IdxSegs { int n; uint32_t block; IdxSeg s[IDX_MAX_SEGS]; };
static constexpr int VEC_ITER = 8;

__global__ void __launch_bounds__(BX) k_index_vec(const __grid_constant__ IdxSegs P, uint32_t* __restrict__ out) {
  int k = 0;
  while (k+1 < P.n && blockIdx.x >= P.s[k+1].firstBlock) ++k;
  const IdxSeg& S = P.s[k];
  const uint32_t block = P.block;
  const long long g = (S.base >> 2) + (long long)(blockIdx.x - S.firstBlock)*(BX*VEC_ITER) + threadIdx.x;
  const long long e0 = 4*g - S.base;               // entity of the vector's first lane; -3..-1 only for the first vector of a stream
  if (e0 >= S.count) return;
  const int s0 = S.s0, s1 = S.s1;
  int i0, i1, i2;
  {
    const long long ee = e0 < 0 ? 0 : e0, q = ee / s0;
    i0 = int(ee - q*s0); i2 = int(q / s1); i1 = int(q - (long long)i2*s1);
  }
  uint32_t* p = out + (S.base + e0);
  // 32-bit bookkeeping inside the loop (ncu: the 64-bit form was issue bound at 72 instructions per vector): `left` = entities from
  // lane 0 of the current vector to the end of the stream, clamped (a thread advances by VEC_ITER*4*BX entities at most)
  const long long leftLL = S.count - e0;
  int left = leftLL > (1ll << 30) ? (1 << 30) : int(leftLL);
  int skip = e0 < 0 ? int(-e0) : 0;                // lanes of the stream's first vector that lie before the stream
#pragma unroll 2
  for (int j = 0; j < VEC_ITER; ++j) {
    const uint32_t val = S.val0 + uint32_t(i0)*block + uint32_t(i1)*S.rowStride + uint32_t(i2)*S.planeStride;
    if (skip == 0 && left >= 4 && i0 + 3 < s0) {
      *reinterpret_cast<uint4*>(p) = make_uint4(val, val + block, val + 2*block, val + 3*block);
    } else {
      int a0 = i0, a1 = i1, a2 = i2;               // (a0,a1,a2) is entity 0 until the lane reaches it
      for (int l = skip; l < 4 && l < left; ++l) {
        p[l] = S.val0 + uint32_t(a0)*block + uint32_t(a1)*S.rowStride + uint32_t(a2)*S.planeStride;
        if (++a0 == s0) { a0 = 0; if (++a1 == s1) { a1 = 0; ++a2; } }
      }
    }
    left -= 4*BX;
    if (left <= 0) break;
    p += 4*BX;
    i0 += 4*BX - skip;                             // the walk state stood at entity 0, not at e0, in the stream's first vector
    skip = 0;
    if (i0 >= s0) {
      if (s0 >= BX) { do { i0 -= s0; ++i1; } while (i0 >= s0); }
      else { const int r = i0 / s0; i0 -= r*s0; i1 += r; }
      if (i1 >= s1) { const int r = i1 / s1; i1 -= r*s1; i2 += r; }
    }
  }
}

// host side: one segment per stream; returns false if the call has more streams than a launch takes
static bool idx_add_segment(IdxSegs& P, unsigned& blocks, const dgb_view& v, int comp, int kind, const dgb_box& walk, const int* move,
                            long long base, uint32_t block, uint32_t offset) {
  if (P.n >= IDX_MAX_SEGS) return false;
  const dgb_component& c = v.comp[comp];
  const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
  long long count = 1; for (int i = 0; i < v.dim; ++i) count *= walk.size[i];
  if (count <= 0) return true;
  IdxSeg& S = P.s[P.n];
  S.base = base; S.count = count; S.s0 = walk.size[0]; S.s1 = v.dim > 1 ? walk.size[1] : 1;
  long long idx = c.index_offset[kind], stride = 1;
  long long strides[3] = {0, 0, 0};
  for (int i = 0; i < v.dim; ++i) { idx += (long long)(walk.origin[i] + move[i] - of.origin[i])*stride; strides[i] = stride; stride *= of.size[i]; }
  S.val0 = uint32_t(idx)*block + offset;
  S.rowStride = uint32_t(strides[1])*block; S.planeStride = uint32_t(strides[2])*block;
  S.firstBlock = blocks;
  const long long vectors = ((base + count + 3) >> 2) - (base >> 2);
  blocks += unsigned((vectors + BX*VEC_ITER - 1)/(BX*VEC_ITER));
  ++P.n;
  return true;
}

// ---- K3b: intersection normals -----------------------------------------------------------------------------------------
// YaspIntersection::{outerNormal,unitOuterNormal,centerUnitOuterNormal} = +-e_dir (yaspgridintersection.hh:163-182),
// integrationOuterNormal = geometry().volume() * centerUnitOuterNormal() (:184-190): out[(k*dim+d)*n + cell]
__global__ void __launch_bounds__(BX) k_intersection_normals(const __grid_constant__ dgb_view v, const SweepGeom s, double* __restrict__ unit,
                                                             double* __restrict__ integration) {
  int coord[3]; long long pos;
  if (!sweep_entity<1>(v, s, 0, coord, pos)) return;
  const int dim = v.dim;
  double w[3] = {1.0, 1.0, 1.0};
  if (integration)
    for (int i = 0; i < dim; ++i) {
      const double sh = periodic_shift(v, i, coord[i]);
      double lo = vertex_coord(v, i, coord[i]), hi = vertex_coord(v, i, coord[i]+1);
      if (sh != 0.0) { lo += sh; hi += sh; }
      w[i] = hi-lo;
    }
  for (int k = 0; k < 2*dim; ++k) {
    const int dir = k >> 1;
    const double sgn = (k & 1) ? 1.0 : -1.0;
    double vol = 1.0;
    for (int i = 0; i < dim; ++i) if (i != dir) vol *= w[i];
    for (int d = 0; d < dim; ++d) {
      const double n = d == dir ? sgn : 0.0;
      const long long at = ((long long)k*dim + d)*s.n + pos;
      if (unit) unit[at] = n;
      if (integration) integration[at] = vol*n;
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


// ---- K4b: axis-aligned geometry of the entities of ANY codimension at local points ---------------------------------------
// AxisAlignedCubeGeometry<mydim, dim> (yaspgrid/yaspgridgeometry.hh:30-77; entity geometries yaspgridentity.hh:270-274, 491-514,
// 849-851): the axes in the component's shift span the entity, the others are flat at the lower corner; only cells are wrapped
// periodically. global[(q*dim+d)*n+e], jt[((q*mydim+r)*dim+c)*n+e], jit[((q*dim+r)*mydim+c)*n+e], detJ[q*n+e].
struct GeoOutCodim { double *global, *jt, *jit, *detJ; };

__global__ void __launch_bounds__(BX) k_geometry_codim(const __grid_constant__ dgb_view v, const SweepGeom s, const QuadPoints q, const GeoOutCodim o) {
  int coord[3]; long long pos;
  if (!sweep_entity<1>(v, s, 0, coord, pos)) return;
  const dgb_component& c = v.comp[s.comp];
  const int dim = v.dim, mydim = dim - c.codim;
  double ll[3], w[3], vol = 1.0;
  int local[3];                                        // local coordinate number of a spanned axis, -1 for a flat one
  int lc = 0;
  for (int i = 0; i < dim; ++i) {
    const bool ext = c.shift >> i & 1u;
    double lo = vertex_coord(v, i, coord[i]), hi = ext ? vertex_coord(v, i, coord[i]+1) : lo;
    if (c.codim == 0) { const double sh = periodic_shift(v, i, coord[i]); if (sh != 0.0) { lo += sh; hi += sh; } }
    ll[i] = lo; w[i] = hi-lo;
    local[i] = ext ? lc++ : -1;
    if (ext) vol *= w[i];
  }
  for (int p = 0; p < q.n; ++p) {
    for (int i = 0; i < dim; ++i) {
      if (o.global) o.global[((long long)p*dim + i)*s.n + pos] = local[i] >= 0 ? ll[i] + q.x[p*mydim + local[i]]*w[i] : ll[i];
      for (int r = 0; r < mydim; ++r) {
        if (o.jt) o.jt[(((long long)p*mydim + r)*dim + i)*s.n + pos] = (local[i] == r) ? w[i] : 0.0;
        if (o.jit) o.jit[(((long long)p*dim + i)*mydim + r)*s.n + pos] = (local[i] == r) ? 1.0/w[i] : 0.0;
      }
    }
    if (o.detJ) o.detJ[(long long)p*s.n + pos] = vol;
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


// ---- K8b: GeometryGrid intersections ---------------------------------------------------------------------------------
// GeoGrid::Intersection over the Yasp host intersection (geometrygrid/intersection.hh:103-172): per cell the deformed corners
// are evaluated once (8 coordinate-function calls in 3-D); per face k (dir = k/2, side = k%2) the face corners are
// insideGeo.global(hostLocalGeometry.corner(i)) (cornerstorage.hh:148-154; the host's local face geometry is the unit face of
// yaspgridintersection.hh:195-209), the face geometry is the multilinear map over them (centre, volume), and the normals at a
// face-local point come from the ELEMENT's Jacobian at x = geometryInInside().global(local): outerNormal = J^-T n_ref,
// integrationOuterNormal = J^-T n_ref |det J|, unitOuterNormal = outerNormal / |outerNormal|.
struct FacePoints { int n; int center; double x[16*2]; };     // center: index of a point that IS the face centre (its unit normal is centerUnitOuterNormal), -1 if none
struct GeoFaceOut { double *corners, *center, *volume, *ion, *on, *uon, *cuon; };

template<int DIM>
__device__ __forceinline__ void geoface_normals(const double (*c)[DIM], int dir, int side, const double* xl, double* vI, double* vO, double* vU) {
  double x[DIM], J[DIM][DIM], Ji[DIM][DIM], nrm[DIM];
  int b = 0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) x[i] = (i == dir) ? double(side) : 0.0 + xl[b++]*(1.0 - 0.0);     // AxisAlignedCubeGeometry::global of the unit face
  ml_jt<DIM>(c, x, J);
  const double det = ml_det<DIM>(J);
  ml_inverse_transposed<DIM>(J, det, Ji);
  const double sgn = side ? 1.0 : -1.0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) nrm[i] = Ji[i][dir]*sgn;                                        // jit.mv(refNormal), refNormal = +-e_dir
  if (vO) {
#pragma unroll
    for (int i = 0; i < DIM; ++i) vO[i] = nrm[i];
  }
  if (vI) {
    const double a = fabs(det);                                                               // jit.detInv()
#pragma unroll
    for (int i = 0; i < DIM; ++i) vI[i] = nrm[i]*a;
  }
  if (vU) {
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) s2 += nrm[i]*nrm[i];
    const double f = 1.0/sqrt(s2);
#pragma unroll
    for (int i = 0; i < DIM; ++i) vU[i] = nrm[i]*f;
  }
}

template<int DIM>
__global__ void __launch_bounds__(BX) k_geogrid_faces(const __grid_constant__ dgb_view v, const SweepGeom s, const FacePoints q,
                                                      const Deformation def, const GeoFaceOut o) {
  int coord[3]; long long pos;
  if (!sweep_entity<1>(v, s, 0, coord, pos)) return;
  constexpr int NC = 1 << DIM, NFC = 1 << (DIM-1);
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
    for (int i = 0; i < DIM; ++i) x[i] = (k >> i & 1) ? ur[i] : ll[i];
    deform<DIM>(def, x, c[k]);
  }
  double half[DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i) half[i] = 0.5;
#pragma unroll
  for (int f = 0; f < 2*DIM; ++f) {
    const int dir = f/2, side = f % 2;
    if (o.corners || o.center || o.volume) {
      double fc[NFC][DIM];
#pragma unroll
      for (int i = 0; i < NFC; ++i) {
        double xl[DIM];
        int b = 0;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) xl[ax] = (ax == dir) ? double(side) : double((i >> b++) & 1);
        ml_global<DIM>(c, xl, fc[i]);
        if (o.corners)
#pragma unroll
          for (int d = 0; d < DIM; ++d) o.corners[((long long)(f*NFC + i)*DIM + d)*s.n + pos] = fc[i][d];
      }
      if (o.center) {
        double y[DIM];
        MLGlobal<DIM, DIM-1, 0, false>::run(fc, half, 1.0, y);
#pragma unroll
        for (int d = 0; d < DIM; ++d) o.center[((long long)f*DIM + d)*s.n + pos] = y[d];
      }
      if (o.volume) {
        double jt[DIM > 1 ? DIM-1 : 1][DIM];
        MLJt<DIM, DIM-1, 0, false>::run(fc, half, 1.0, jt);
        double vol;
        if (DIM == 3) {                                   // FieldMatrixHelper::sqrtDetAAT<2,3>
          const double v0 = jt[0][0]*jt[1][1] - jt[0][1]*jt[1][0];
          const double v1 = jt[0][0]*jt[1][2 % DIM] - jt[1][0]*jt[0][2 % DIM];
          const double v2 = jt[0][1]*jt[1][2 % DIM] - jt[0][2 % DIM]*jt[1][1];
          vol = sqrt(v0*v0 + v1*v1 + v2*v2);
        } else {                                          // <1,2>: Cholesky factor of the 1x1 matrix A A^T
          double a = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) a += jt[0][d]*jt[0][d];
          vol = sqrt(a);
        }
        o.volume[(long long)f*s.n + pos] = vol;
      }
    }
    for (int p = 0; p < q.n; ++p) {
      double vI[DIM], vO[DIM], vU[DIM];
      const bool isCenter = o.cuon && p == q.center;             // same expression, same point: no second evaluation for centerUnitOuterNormal
      geoface_normals<DIM>(c, dir, side, q.x + p*(DIM-1), o.ion ? vI : nullptr, o.on ? vO : nullptr, (o.uon || isCenter) ? vU : nullptr);
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const long long at = ((long long)(f*q.n + p)*DIM + d)*s.n + pos;
        if (o.ion) o.ion[at] = vI[d];
        if (o.on) o.on[at] = vO[d];
        if (o.uon) o.uon[at] = vU[d];
        if (isCenter) o.cuon[((long long)f*DIM + d)*s.n + pos] = vU[d];
      }
    }
    if (o.cuon && q.center < 0) {
      double vU[DIM];
      geoface_normals<DIM>(c, dir, side, half, nullptr, nullptr, vU);
#pragma unroll
      for (int d = 0; d < DIM; ++d) o.cuon[((long long)f*DIM + d)*s.n + pos] = vU[d];
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
  static const int rows = [] { const char* e = getenv("DGB_IDX_ROWS"); const int r = e ? atoi(e) : 32; return r == 64 || r == 128 ? r : 32; }();      // measured (403 M faces): 32: 0.325 ms, 64: 0.365, 128: 0.396
  static const int linear = [] { const char* e = getenv("DGB_SWEEP_LINEAR"); return e ? atoi(e) : -1; }();
  static const bool vec = [] { const char* e = getenv("DGB_IDX_VEC"); return !e || atoi(e) != 0; }();
  if (vec && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {       // one launch, 16-byte stores (k_index_vec)
    IdxSegs P; P.n = 0; P.block = o.block;
    unsigned blocks = 0;
    const int zero[3] = {0, 0, 0};
    const int rc = for_each_component(*v, codim, pitype, rows, [&](const SweepGeom& s, dim3) {
      idx_add_segment(P, blocks, *v, s.comp, s.kind, v->comp[s.comp].box[s.kind], zero, s.base, o.block, o.offset);
    });
    if (rc != DGB_OK) return rc;
    if (blocks) k_index_vec<<<blocks, BX, 0, (cudaStream_t)stream>>>(P, d_out);
    DGB_CUDA(cudaGetLastError());
    return DGB_OK;
  }
  return for_each_component(*v, codim, pitype, rows, [&](const SweepGeom& s, dim3 grid) {
    const dgb_box& b = v->comp[s.comp].box[s.kind];
    if (linear == 1 || (linear == -1 && b.size[0] % 32 != 0)) {
      long long count = 1; for (int i = 0; i < v->dim; ++i) count *= b.size[i];
      const unsigned nb = unsigned((count + BX*32 - 1)/(BX*32));
      k_index_linear<<<nb, BX, 0, (cudaStream_t)stream>>>(*v, s, o.block, o.offset, d_out, count);
      return;
    }
    if (rows == 32) k_index_rows<32><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o.block, o.offset, d_out);
    else if (rows == 128) k_index_rows<128><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o.block, o.offset, d_out);
    else k_index_rows<64><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o.block, o.offset, d_out);
  });
}

int dgb_mapper_subindex(const dgb_view* v, const dgb_mapper* m, int cc, int pitype, uint32_t* d_out, void* stream) {
  if (cc < 0 || cc > v->dim) return fail(DGB_EARG, "bad codim");
  if (m->offset[cc] == DGB_INVALID_OFFSET) return fail(DGB_EARG, "codim not contained in the mapper layout");
  SubTable t; t.n = sub_entities(v->dim, cc);
  for (int i = 0; i < t.n; ++i) { t.shift[i] = (unsigned char)sub_shift(v->dim, i, cc); t.move[i] = (unsigned char)sub_move(v->dim, i, cc); }
  const uint32_t block = m->block[cc], offset = m->offset[cc];
  static const bool vec = [] { const char* e = getenv("DGB_IDX_VEC"); return !e || atoi(e) != 0; }();
  if (vec && t.n <= IDX_MAX_SEGS && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
    // codim 0 has one component: one stream per sub-entity number, all in one launch of k_index_vec
    IdxSegs P; P.n = 0; P.block = block;
    unsigned blocks = 0;
    const int rc = for_each_component(*v, 0, pitype, IDX_ROWS, [&](const SweepGeom& s, dim3) {
      for (int i = 0; i < t.n; ++i) {
        // yaspgridentity.hh:764-776: coord += move, component = shiftmapping[shift], index in the overlapfront numbering
        const int move[3] = { t.move[i] & 1, t.move[i] >> 1 & 1, t.move[i] >> 2 & 1 };
        idx_add_segment(P, blocks, *v, v->comp_begin[cc] + v->shiftmap[t.shift[i]], DGB_BOX_OVERLAPFRONT, v->comp[s.comp].box[s.kind], move,
                        (long long)i*s.n + s.base, block, offset);
      }
    });
    if (rc != DGB_OK) return rc;
    if (blocks) k_index_vec<<<blocks, BX, 0, (cudaStream_t)stream>>>(P, d_out);
    DGB_CUDA(cudaGetLastError());
    return DGB_OK;
  }
  return for_each_component(*v, 0, pitype, IDX_ROWS, [&](const SweepGeom& s, dim3 grid) {
    k_subindex_rows<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, t, cc, block, offset, d_out);
  });
}

int dgb_elements_materialize(const dgb_view* v, int codim, int pitype, uint32_t* d_index, uint8_t* d_ptype, int32_t* d_coord,
                             double* d_lower, double* d_upper, double* d_center, double* d_volume, void* stream) {
  if ((d_lower || d_upper || d_center || d_volume)) if (int rc = check_tensor(*v)) return rc;
  EntityOut o{d_index, d_ptype, d_coord, d_lower, d_upper, d_center, d_volume, 1u, 0u};
  static const int linear = [] { const char* e = getenv("DGB_SWEEP_LINEAR"); return e ? atoi(e) : -1; }();      // -1 automatic, 0 never, 1 always
  return for_each_component(*v, codim, pitype, ENT_ROWS, [&](const SweepGeom& s, dim3 grid) {
    const dgb_box& b = v->comp[s.comp].box[s.kind];
    // rows that do not keep the warps aligned: linear walk - where it was measured to win (tools/entities_probe.py, 384^3, fraction of
    // the copy peak, column walk -> linear walk: cells 0.82 -> 0.98 (tensor, x periodic); faces 0.93 -> 1.06 / 0.81 -> 0.97; edges
    // 0.84 -> 1.02 equidistant but 0.78 -> 0.55 with tensor-product tables; vertices 0.75 -> 0.67 / 0.76 -> 0.58)
    const int cd = v->comp[s.comp].codim;
    const bool pays = cd <= 1 || (cd == 2 && v->coord_mode != DGB_COORD_TENSORPRODUCT);
    if (linear == 1 || (linear == -1 && b.size[0] % 32 != 0 && pays)) {
      long long count = 1; for (int i = 0; i < v->dim; ++i) count *= b.size[i];
      const unsigned nb = unsigned((count + BX*LIN_ROWS - 1)/(BX*LIN_ROWS));
      k_entities_linear<<<nb, BX, 0, (cudaStream_t)stream>>>(*v, s, o, count);
      return;
    }
    k_entities<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o);
  });
}

int dgb_intersections_materialize(const dgb_view* v, int pitype, uint8_t* d_neighbor, uint8_t* d_boundary, int32_t* d_outside,
                                  int32_t* d_bsi, double* d_center, double* d_volume, void* stream) {
  if (d_center || d_volume) if (int rc = check_tensor(*v)) return rc;
  FaceOut o{d_neighbor, d_boundary, d_outside, d_bsi, d_center, d_volume};
  static const int rows = [] { const char* e = getenv("DGB_FACE_ROWS"); const int r = e ? atoi(e) : FACE_ROWS_DEFAULT; return r; }();
  static const int linear = [] { const char* e = getenv("DGB_SWEEP_LINEAR"); return e ? atoi(e) : -1; }();      // -1 automatic, 0 never, 1 always
  return for_each_component(*v, 0, pitype, rows, [&](const SweepGeom& s, dim3 grid) {
    const dgb_box& b = v->comp[s.comp].box[s.kind];
    if (linear == 1 || (linear == -1 && b.size[0] % 32 != 0)) {       // rows that do not keep the warps aligned: linear walk
      long long count = 1; for (int i = 0; i < v->dim; ++i) count *= b.size[i];
      const unsigned nb = unsigned((count + BX*LIN_ROWS - 1)/(BX*LIN_ROWS));
      k_intersections_linear<<<nb, BX, 0, (cudaStream_t)stream>>>(*v, s, o, count);
      return;
    }
    switch (rows) {                                   // rows walked per thread (tuning knob DGB_FACE_ROWS)
      case 1: k_intersections<1><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o); break;
      case 2: k_intersections<2><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o); break;
      case 8: k_intersections<8><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o); break;
      case 16: k_intersections<16><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o); break;
      default: k_intersections<4><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, o); break;
    }
  });
}

int dgb_intersection_normals(const dgb_view* v, int pitype, double* d_unitOuterNormal, double* d_integrationOuterNormal, void* stream) {
  if (d_integrationOuterNormal) if (int rc = check_tensor(*v)) return rc;
  return for_each_component(*v, 0, pitype, 1, [&](const SweepGeom& s, dim3 grid) {
    k_intersection_normals<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, d_unitOuterNormal, d_integrationOuterNormal);
  });
}

// geometryInInside() / geometryInOutside() of face k: the unit face of the reference cube, identical for every cell
// (yaspgridintersection.hh:195-228): lower / upper corner in the local coordinates of the inside / outside element
int dgb_intersection_local_geometry(int dim, int k, int in_outside, double* lower, double* upper) {
  if (dim < 1 || dim > 3 || k < 0 || k >= 2*dim) return fail(DGB_EARG, "bad dimension or face number");
  const int dir = k/2, face = k % 2;
  for (int i = 0; i < dim; ++i) { lower[i] = 0.0; upper[i] = 1.0; }
  const double at = in_outside ? (face == 1 ? 0.0 : 1.0) : (face == 0 ? 0.0 : 1.0);
  lower[dir] = upper[dir] = at;
  return DGB_OK;
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

int dgb_geometry_eval_codim(const dgb_view* v, int codim, int pitype, int nqp, const double* qp_host, double* d_global, double* d_jt,
                            double* d_jit, double* d_detJ, void* stream) {
  if (codim < 0 || codim > v->dim) return fail(DGB_EARG, "bad codim");
  if (int rc = check_tensor(*v)) return rc;
  if (nqp < 0 || nqp > 27) return fail(DGB_EARG, "at most 27 quadrature points per call");
  QuadPoints q; q.n = nqp;
  for (int i = 0; i < nqp*(v->dim-codim); ++i) q.x[i] = qp_host[i];
  GeoOutCodim o{d_global, d_jt, d_jit, d_detJ};
  return for_each_component(*v, codim, pitype, 1, [&](const SweepGeom& s, dim3 grid) {
    k_geometry_codim<<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, o);
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

int dgb_geogrid_intersections(const dgb_view* v, int pitype, int deformation, const double* params_host, int nparams, int nqp,
                              const double* qp_face_host, double* d_corners, double* d_center, double* d_volume,
                              double* d_integrationOuterNormal, double* d_outerNormal, double* d_unitOuterNormal,
                              double* d_centerUnitOuterNormal, void* stream) {
  if (int rc = check_tensor(*v)) return rc;
  if (nqp < 0 || nqp > 16) return fail(DGB_EARG, "at most 16 face points per call");
  if (deformation < 0 || deformation > 1) return fail(DGB_EARG, "unknown deformation id");
  FacePoints q; q.n = nqp;
  for (int i = 0; i < nqp*(v->dim-1); ++i) q.x[i] = qp_face_host[i];
  q.center = -1;
  for (int p = 0; p < nqp && q.center < 0; ++p) {
    bool c = true;
    for (int i = 0; i < v->dim-1; ++i) c = c && qp_face_host[p*(v->dim-1)+i] == 0.5;
    if (c) q.center = p;
  }
  Deformation def; def.id = deformation;
  for (int i = 0; i < 4; ++i) def.p[i] = (params_host && i < nparams) ? params_host[i] : 0.0;
  GeoFaceOut o{d_corners, d_center, d_volume, d_integrationOuterNormal, d_outerNormal, d_unitOuterNormal, d_centerUnitOuterNormal};
  return for_each_component(*v, 0, pitype, 1, [&](const SweepGeom& s, dim3 grid) {
    if (v->dim == 3) k_geogrid_faces<3><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, def, o);
    else k_geogrid_faces<2><<<grid, BX, 0, (cudaStream_t)stream>>>(*v, s, q, def, o);
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
