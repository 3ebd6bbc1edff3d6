// dgb_gis.cu — Dune::GlobalIndexSet<GridView> (dune/grid/utility/globalindexset.hh) as device kernels.
//
// The reference (ctor :318-440) builds std::maps keyed by global ids while it walks cells and their sub-entities:
//   (1) UniqueEntityPartition (:176-203): assignment[e] = rank if the entity is interior or border here, else -1; then
//       communicate(MinimumExchange, All_All_Interface, Forward): if (x >= 0) assignment = min(x, assignment)
//   (2) nLocal = #entities owned here; allgather of the counts; myoffset = sum of the counts of lower ranks (:338-362)
//   (3) cells in iteration order, sub-entities i = 0..subEntities(codim)-1: the FIRST time an entity is met it gets
//       myoffset + counter++ if owned here, else -1 (:384-428); codim 0: interior cells in iteration order (:384-402)
//   (4) communicate(IndexExchange, All_All_Interface, Forward): if (x >= 0) index = x (:262-289)
// Here: (1) one kernel + dgb_communicate(op MIN_NONNEG); (3) "met for the first time at (cell, i)" is decided
// arithmetically - (cell, i) is the first visit of its entity iff in every flat direction of the entity the cell is the
// lower of the two cells touching it, or there is no lower cell in the stored box - and the counter is an exclusive
// device-wide prefix sum (cub::DeviceScan) over the flags in (cell, i) order; (2) the total of the same scan, gathered
// over the ranks with NCCL; (4) dgb_communicate(op SET_NONNEG). Indices travel as doubles (exact below 2^53) because the
// exchange plans move doubles; the result is int32 like the reference's `typedef int Index`.
#include <cub/device/device_scan.cuh>
#include "dgb_internal.hh"
#include "dgb_device.cuh"
#include "dgb_comm.hh"

namespace dgb {

static constexpr int GT = 256;

// entity with consecutive index e of codim `codim`: component, coordinates
__device__ __forceinline__ const dgb_component& gis_entity(const dgb_view& v, int codim, long long e, int* coord) {
  int j = v.comp_begin[codim];
  while (j+1 < v.comp_begin[codim+1] && e >= v.comp[j+1].index_offset[DGB_BOX_OVERLAPFRONT]) ++j;
  const dgb_component& c = v.comp[j];
  const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
  unsigned l = unsigned(e - c.index_offset[DGB_BOX_OVERLAPFRONT]);
  for (int i = 0; i < 3; ++i) {
    if (i < v.dim) { const unsigned s = unsigned(of.size[i]); coord[i] = of.origin[i] + int(l % s); l /= s; }
    else coord[i] = 0;
  }
  return c;
}

// (1) before the exchange: own rank where the entity is interior or border (codim 0: interior), -1 elsewhere; gidx = -1
__global__ void __launch_bounds__(GT) k_gis_assign(const __grid_constant__ dgb_view v, int codim, long long n, double* own, double* gidx) {
  const long long e = (long long)blockIdx.x*GT + threadIdx.x;
  if (e >= n) return;
  int coord[3];
  const dgb_component& c = gis_entity(v, codim, e, coord);
  const uint8_t pt = partition_type(v, c, coord);
  own[e] = (pt == DGB_InteriorEntity || (codim != 0 && pt == DGB_BorderEntity)) ? double(v.rank) : -1.0;
  gidx[e] = -1.0;
}

// (3a) flags[cell*nsub + i] = 1 iff sub-entity i of the cell is owned here and met for the first time at this (cell, i)
__global__ void __launch_bounds__(GT) k_gis_flags(const __grid_constant__ dgb_view v, int codim, const SubTable t, long long ncells,
                                                  const double* __restrict__ own, int* __restrict__ flags) {
  const long long cell = (long long)blockIdx.x*GT + threadIdx.x;
  if (cell >= ncells) return;
  int coord[3];
  const dgb_component& cc = gis_entity(v, 0, cell, coord);
  const dgb_box& cells = cc.box[DGB_BOX_OVERLAPFRONT];
  for (int i = 0; i < t.n; ++i) {
    const dgb_component& c = v.comp[v.comp_begin[codim] + v.shiftmap[t.shift[i]]];
    const int sc[3] = { coord[0] + (t.move[i] & 1), coord[1] + (t.move[i] >> 1 & 1), coord[2] + (t.move[i] >> 2 & 1) };
    bool first = true;
    for (int j = 0; j < v.dim; ++j)
      if (!(t.shift[i] >> j & 1) && !(t.move[i] >> j & 1) && coord[j] > cells.origin[j]) first = false;   // the cell below met it
    const uint32_t idx = entity_index(v, c, DGB_BOX_OVERLAPFRONT, sc);
    flags[cell*t.n + i] = (first && own[idx] == double(v.rank)) ? 1 : 0;
  }
}

// (3b) owned entities: index = myoffset + number of owned entities met earlier
__global__ void __launch_bounds__(GT) k_gis_number(const __grid_constant__ dgb_view v, int codim, const SubTable t, long long ncells,
                                                   const int* __restrict__ flags, const int* __restrict__ scan, long long myoffset,
                                                   double* __restrict__ gidx) {
  const long long cell = (long long)blockIdx.x*GT + threadIdx.x;
  if (cell >= ncells) return;
  int coord[3];
  gis_entity(v, 0, cell, coord);
  for (int i = 0; i < t.n; ++i) {
    if (!flags[cell*t.n + i]) continue;
    const dgb_component& c = v.comp[v.comp_begin[codim] + v.shiftmap[t.shift[i]]];
    const int sc[3] = { coord[0] + (t.move[i] & 1), coord[1] + (t.move[i] >> 1 & 1), coord[2] + (t.move[i] >> 2 & 1) };
    gidx[entity_index(v, c, DGB_BOX_OVERLAPFRONT, sc)] = double(myoffset + scan[cell*t.n + i]);
  }
}

__global__ void __launch_bounds__(GT) k_gis_total(const int* flags, const int* scan, long long m, long long* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total = m > 0 ? (long long)scan[m-1] + flags[m-1] : 0;
}
__global__ void __launch_bounds__(GT) k_gis_final(const double* __restrict__ gidx, long long n, int32_t* __restrict__ out) {
  const long long e = (long long)blockIdx.x*GT + threadIdx.x;
  if (e < n) out[e] = int32_t(gidx[e]);
}

struct DevBuf {                       // scratch owned by one call
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) { DGB_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); return DGB_OK; }
  template<class T> T* as() { return static_cast<T*>(p); }
};

} // namespace dgb

using namespace dgb;

extern "C" int dgb_global_index(dgb_grid* g, int level, int codim, int32_t* d_out, int64_t* global_size, dgb_comm* comm, void* stream_) {
  return guarded([&]() -> int {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!g || !d_out || !global_size) throw ArgError("null argument");
    if (int rc = require_device()) return rc;
    dgb_view v; if (int rc = make_view(g, level, &v)) return rc;
    if (codim < 0 || codim > v.dim) throw ArgError("bad codim");
    if (v.nranks > 1 && !comm) throw ArgError("a dgb_comm is required on a multi-rank grid");
    const long long n = partition_size(v, codim, DGB_All_Partition), ncells = partition_size(v, 0, DGB_All_Partition);
    SubTable t; t.n = sub_entities(v.dim, codim);
    for (int i = 0; i < t.n; ++i) { t.shift[i] = (unsigned char)sub_shift(v.dim, i, codim); t.move[i] = (unsigned char)sub_move(v.dim, i, codim); }
    const long long m = ncells*t.n;
    if (m >= (1ll << 31)) throw ArgError("global index: more than 2^31 (cell, sub-entity) pairs on one rank");
    uint32_t block[4] = {0, 0, 0, 0}; block[codim] = 1;
    dgb_mapper mapper; dgb_mapper_init(&v, block, &mapper);
    DevBuf own, gidx, flags, scan, tmp, counts;
    if (int rc = own.alloc(n*sizeof(double))) return rc;
    if (int rc = gidx.alloc(n*sizeof(double))) return rc;
    if (int rc = flags.alloc(m*sizeof(int))) return rc;
    if (int rc = scan.alloc(m*sizeof(int))) return rc;
    if (int rc = counts.alloc((v.nranks+1)*sizeof(long long))) return rc;
    const int32_t items[1] = {1};
    if (n > 0) k_gis_assign<<<unsigned((n+GT-1)/GT), GT, 0, stream>>>(v, codim, n, own.as<double>(), gidx.as<double>());   // a rank may hold no entity at all
    DGB_CUDA(cudaGetLastError());
    if (codim != 0) {                                                  // UniqueEntityPartition, :176-203
      double* arrays[1] = { own.as<double>() };
      if (int rc = dgb_communicate(g, level, &mapper, DGB_All_All_Interface, DGB_ForwardCommunication, DGB_OP_MIN_NONNEG, arrays, items, 1, comm, stream)) return rc;
    }
    if (ncells > 0) k_gis_flags<<<unsigned((ncells+GT-1)/GT), GT, 0, stream>>>(v, codim, t, ncells, own.as<double>(), flags.as<int>());
    DGB_CUDA(cudaGetLastError());
    size_t tmpBytes = 0;
    DGB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, flags.as<int>(), scan.as<int>(), int(m), stream));
    if (int rc = tmp.alloc(tmpBytes)) return rc;
    if (m > 0) DGB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, flags.as<int>(), scan.as<int>(), int(m), stream));
    // counts of all ranks -> my offset and the global size (:338-362)
    long long* dCounts = counts.as<long long>();
    k_gis_total<<<1, GT, 0, stream>>>(flags.as<int>(), scan.as<int>(), m, dCounts + v.nranks);
    if (int rc = nccl_allgather_bytes(dCounts + v.nranks, dCounts, sizeof(long long), comm, stream)) return rc;
    std::vector<long long> hc(v.nranks);
    DGB_CUDA(cudaMemcpyAsync(hc.data(), dCounts, v.nranks*sizeof(long long), cudaMemcpyDeviceToHost, stream));
    DGB_CUDA(cudaStreamSynchronize(stream));
    long long myoffset = 0, total = 0;
    for (int r = 0; r < v.nranks; ++r) { if (r < v.rank) myoffset += hc[r]; total += hc[r]; }
    if (total >= (1ll << 31)) throw ArgError("global index: more than 2^31 entities (the reference's Index is int)");
    *global_size = total;
    if (ncells > 0) k_gis_number<<<unsigned((ncells+GT-1)/GT), GT, 0, stream>>>(v, codim, t, ncells, flags.as<int>(), scan.as<int>(), myoffset, gidx.as<double>());
    DGB_CUDA(cudaGetLastError());
    {                                                                  // IndexExchange, :430-431
      double* arrays[1] = { gidx.as<double>() };
      if (int rc = dgb_communicate(g, level, &mapper, DGB_All_All_Interface, DGB_ForwardCommunication, DGB_OP_SET_NONNEG, arrays, items, 1, comm, stream)) return rc;
    }
    if (n > 0) k_gis_final<<<unsigned((n+GT-1)/GT), GT, 0, stream>>>(gidx.as<double>(), n, d_out);
    DGB_CUDA(cudaGetLastError());
    DGB_CUDA(cudaStreamSynchronize(stream));
    return DGB_OK;
  });
}
