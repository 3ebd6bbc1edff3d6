// dgb_internal.hh — definitions shared by the translation units of libdgb.so (not installed).
#ifndef DGB_INTERNAL_HH
#define DGB_INTERNAL_HH
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "dgb_host.hh"

namespace dgb {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define DGB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
  return dgb::fail(DGB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)

// translate C++ exceptions of the host layer to C error codes at the ABI boundary
template<class F> int guarded(F&& f) {
  try { return f(); }
  catch (const GridError& e) { return fail(DGB_EGRID, e.what()); }
  catch (const RangeError& e) { return fail(DGB_ERANGE, e.what()); }
  catch (const ArgError& e) { return fail(DGB_EARG, e.what()); }
  catch (const std::exception& e) { return fail(DGB_EARG, e.what()); }
}

// one cached exchange plan (see dgb_comm.cu)
struct CommPlan;

} // namespace dgb

struct dgb_comm;

struct dgb_grid {
  dgb::RankGrid me;
  std::vector<dgb::RankGrid> all;            // integer descriptors of every rank (index = rank)
  bool keepOverlap = true;
  int64_t nBoundarySegments = 0;
  // device copies of tensor coordinate arrays per level (owned)
  struct DevCoords { double* d[3] = {nullptr,nullptr,nullptr}; };
  std::vector<DevCoords> devCoords;
  std::map<std::string, std::shared_ptr<dgb::CommPlan>> plans;
  int64_t lastBytesSent = 0;
  std::mutex mtx;
};

namespace dgb {
// fills a view (device tensor pointers are uploaded lazily on first use with a device present)
int make_view(const dgb_grid* g, int level, dgb_view* v);
// kind index (DGB_BOX_*) of the box a partition iterator walks, -1 for Ghost (yaspgrid.hh:1819-1830)
inline int box_kind_of(int pitype) {
  switch (pitype) {
    case DGB_Interior_Partition: return DGB_BOX_INTERIOR;
    case DGB_InteriorBorder_Partition: return DGB_BOX_INTERIORBORDER;
    case DGB_Overlap_Partition: return DGB_BOX_OVERLAP;
    case DGB_OverlapFront_Partition: case DGB_All_Partition: return DGB_BOX_OVERLAPFRONT;
    case DGB_Ghost_Partition: return -1;
  }
  return -2;
}
int64_t partition_size(const dgb_view& v, int codim, int pitype);
int require_device();
}
#endif
