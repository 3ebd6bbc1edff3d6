// dgb_comm.hh — exchange plan shared between dgb_comm.cu and dgb_fv.cu (internal).
#ifndef DGB_COMM_HH
#define DGB_COMM_HH
#include <memory>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/dgb.h"

struct dgb_grid;
struct dgb_comm;

namespace dgb {
// one box of an exchange plan: the entities of a sub-box of one shift component
struct PlanBox {
  int origin[3], size[3];
  int of_origin[3], of_size[3];     // enclosing overlapfront box of the component
  long long buf_offset;             // first double of this box in the send/recv buffer
  long long count;                  // entities
  unsigned index_base;              // mapper: index = (component offset + local superindex)*block + offset
  unsigned block, map_offset;
};
struct CommPlan {
  struct Segment { int peer; long long offset, count; };     // doubles
  std::vector<Segment> sendSeg, recvSeg;
  long long sendTotal = 0, recvTotal = 0, maxSendBox = 0, maxRecvBox = 0;
  int nSendBoxes = 0, nRecvBoxes = 0, items = 0;
  std::vector<PlanBox> hostSend, hostRecv;               // hostRecv is ordered by unpack wave
  std::vector<int> recvWaveEnd;                          // end index (exclusive) of each wave in hostRecv
  PlanBox* dSendBoxes = nullptr; PlanBox* dRecvBoxes = nullptr;
  double* dSendBuf = nullptr; double* dRecvBuf = nullptr;
  bool uploaded = false;
  long long launches = 0;
  cudaEvent_t trace[3] = {nullptr, nullptr, nullptr};     // optional timing events: after pack, after transfer, after unpack (DGB_FV_TRACE)
  // direct mode (plan_enable_direct): the packing kernel stores straight into the RECEIVER's exchange buffer through
  // peer-mapped memory (NVLink), double buffered by step parity; one small all-reduce orders "all senders have written"
  // before the unpack. dDst[parity][box] = where send box `box` starts in its receiver's buffer.
  bool direct = false;
  double* dRecvBuf2[2] = {nullptr, nullptr};              // receive buffers of the direct mode, by step parity
  double** dDst[2] = {nullptr, nullptr};                  // device arrays of nSendBoxes pointers
  std::vector<void*> ipcOpened;
  double* dBarrier = nullptr;
  int upload();
  ~CommPlan();
};
int get_plan(dgb_grid* g, int level, const dgb_mapper& m, int iftype, int dir, const int32_t* item_sizes, int narrays,
             std::shared_ptr<CommPlan>& out);
// packed = true: the send buffer has already been filled (in-kernel pack of the FV shell kernel)
int run_plan(dgb_grid* g, CommPlan& plan, int myrank, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays,
             dgb_comm* comm, cudaStream_t stream, bool packed = false);
int nccl_allreduce(double* d, int op /*0 min, 1 max*/, dgb_comm* comm, cudaStream_t stream);
// collective over all ranks: exchange IPC handles of the plans' receive buffers and the segment offsets, map the peers'
// buffers. Returns DGB_OK with plan.direct = false when peer mapping is not available (the NCCL path stays in use).
int plan_enable_direct(dgb_grid* g, CommPlan& plan, int myrank, int nranks, dgb_comm* comm, cudaStream_t stream);
// the part of a direct exchange after the packing kernel: barrier over the ranks, unpack from buffer `parity`
int run_plan_direct(dgb_grid* g, CommPlan& plan, int parity, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays,
                    dgb_comm* comm, cudaStream_t stream);
}
#endif
