// dgb_device.cuh — device-side closed forms shared by all kernels: vertex coordinates, entity boxes,
// consecutive indices, partition types, intersections, boundary-segment ids. Nothing is ever stored per
// entity; everything is recomputed from (component, i0, i1, i2) and the by-value dgb_view descriptor.
// The whole library is compiled with -fmad=false so fp64 expressions round exactly like the reference's
// scalar C++ (g++ -ffp-contract=off); FMA is used only where written explicitly as fma().
#ifndef DGB_DEVICE_CUH
#define DGB_DEVICE_CUH
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/dgb.h"

namespace dgb {

// vertex coordinate along axis d at global index i (coordinates.hh:65-68,169-172,278-281)
__device__ __forceinline__ double vertex_coord(const dgb_view& v, int d, int i) {
  if (v.coord_mode == DGB_COORD_EQUIDISTANT) return double(i)*v.h[d];
  if (v.coord_mode == DGB_COORD_EQUIDISTANT_OFFSET) return v.origin[d] + double(i)*v.h[d];
  return __ldg(v.tensor[d] + (i - v.tensor_offset[d]));
}

// periodic wrap of element / face coordinates (yaspgridentity.hh:497-510, yaspgridintersection.hh:252-264):
// cells of the periodic overlap live outside [0,N) and are shifted back by the domain size.
__device__ __forceinline__ double periodic_shift(const dgb_view& v, int d, int cellcoord) {
  if (!(v.periodic >> d & 1u)) return 0.0;
  if (cellcoord < 0) return v.L[d];
  if (cellcoord + 1 > v.N[d]) return -v.L[d];
  return 0.0;
}

// consecutive index of `coord` in component `c` (ygrid.hh:472-479 + 764-767); `kind` selects whose YGrid
// offset table is used (the partition the iterator walks, see dgb_host.hh build_boxes)
__device__ __forceinline__ uint32_t entity_index(const dgb_view& v, const dgb_component& c, int kind, const int* coord) {
  const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
  int idx = c.index_offset[kind], stride = 1;
  for (int i = 0; i < v.dim; ++i) { idx += (coord[i] - of.origin[i])*stride; stride *= of.size[i]; }
  return uint32_t(idx);
}

__device__ __forceinline__ bool box_contains(const dgb_box& b, int dim, const int* coord) {
  for (int i = 0; i < dim; ++i)
    if (coord[i] < b.origin[i] || coord[i] >= b.origin[i] + b.size[i]) return false;
  return true;
}

// yaspgridentity.hh:294-305 (generic), :480-488 (codim 0: Interior or Overlap only), :863-874 (vertices)
__device__ __forceinline__ uint8_t partition_type(const dgb_view& v, const dgb_component& c, const int* coord) {
  if (box_contains(c.box[DGB_BOX_INTERIOR], v.dim, coord)) return DGB_InteriorEntity;
  if (c.codim == 0) return DGB_OverlapEntity;
  if (box_contains(c.box[DGB_BOX_INTERIORBORDER], v.dim, coord)) return DGB_BorderEntity;
  if (box_contains(c.box[DGB_BOX_OVERLAP], v.dim, coord)) return DGB_OverlapEntity;
  if (box_contains(c.box[DGB_BOX_OVERLAPFRONT], v.dim, coord)) return DGB_FrontEntity;
  return DGB_GhostEntity;
}

// sub-entity tables (yaspgridentity.hh:154-233): bit j of shift = sub-entity extends along axis j, bit j of move =
// the sub-entity lives on the neighbouring cell in +j. Closed form of the reference-cube numbering:
// sub-entities of codim cc are grouped by their set of "flat" axes F (|F| = cc); groups are ordered by the
// reference element's recursive prism construction, positions inside a group by the binary number of the
// selected sides. The tables are tiny, so they are generated on the host with the reference recursion and
// passed by value (see SubTable in dgb_sweeps.cu).
struct SubTable { int n; unsigned char shift[12]; unsigned char move[12]; };

// face k of a cell (yaspgridintersection.hh:40-57,62-82): dir = k/2, side = k%2
__device__ __forceinline__ bool face_boundary(const dgb_view& v, const int* coord, int dir, int side) {
  if (v.periodic >> dir & 1u) return false;
  const int q = coord[dir] + side;
  return q == 0 || q == v.N[dir];
}
__device__ __forceinline__ bool face_neighbor(const dgb_view& v, const int* coord, int dir, int side) {
  const dgb_box& ov = v.comp[0].box[DGB_BOX_OVERLAP];
  const int q = coord[dir] + side;
  return q > ov.origin[dir] && q <= ov.origin[dir] + ov.size[dir] - 1;
}
// yaspgridintersection.hh:105-162 (numbering on the level-0 local box, highest axis fastest)
__device__ __forceinline__ int boundary_segment_index(const dgb_view& v, const int* coord, int dir, int side) {
  int pos[3], fsize[3];
  int vol = 1;
  for (int i = 0; i < v.dim; ++i) { pos[i] = coord[i] / (1 << v.level) - v.bs_origin[i]; vol *= v.bs_size[i]; }
  for (int k = 0; k < v.dim; ++k) fsize[k] = vol / v.bs_size[k];
  int index = 0, localoffset = 1;
  for (int k = v.dim-1; k >= 0; --k) {
    if (k == dir) continue;
    index += pos[k]*localoffset;
    localoffset *= v.bs_size[k];
  }
  for (int k = 0; k < dir; ++k) index += v.bs_sides[k]*fsize[k];
  index += side*(v.bs_sides[dir] > 1)*fsize[dir];
  return index;
}

// non-negative doubles compare like their bit patterns: max-reduction through an unsigned 64-bit atomic
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double val) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(val));
}

} // namespace dgb
#endif
