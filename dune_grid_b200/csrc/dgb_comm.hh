// dgb_comm.hh — exchange plan shared between dgb_comm.cu and dgb_fv.cu (internal).
#ifndef DGB_COMM_HH
#define DGB_COMM_HH
#include <memory>
#include <mutex>
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
  // the same entities in the partner's own frame, and the partner's enclosing overlapfront box of the component: where a
  // sender finds the receiver's copy of a cell (registered fields: stores go straight into the partner's array)
  int peer_origin[3], peer_of_origin[3], peer_of_size[3];
  int peer;
  int face;                         // 2*axis + (partner on the high side) if the partner is a FACE neighbour (one non-zero torus offset), else -1
  int list_pos;                     // position in the reference's processing order (codim dim..0, list order): hostRecv is re-ordered by wave
};
struct FusedState;
struct CommPlan {
  struct Segment { int peer; long long offset, count; };     // doubles
  // peer-memory form of run_plan (see dgb_comm.cu, "communicate over peer-mapped memory"): set up collectively at the first use
  std::shared_ptr<FusedState> peer;
  bool peerTried = false;
  // does ANY rank exchange with another rank under this plan (evaluated for every rank from its integer descriptors, so all
  // ranks agree)? The peer-memory form is collective: it is taken by all ranks or by none
  bool anyRemote = false;
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
  // A cached plan owns ONE send / receive buffer pair: users on different streams are serialised through this event (every
  // use waits for the previous one and records its own end); host-side bookkeeping is guarded by `mtx`.
  cudaEvent_t busy = nullptr;
  std::mutex mtx;
  int enter(cudaStream_t stream);                         // stream waits for the previous user of the buffers
  int leave(cudaStream_t stream);                         // marks the end of this use
  int upload();
  ~CommPlan();
};
int get_plan(dgb_grid* g, int level, const dgb_mapper& m, int iftype, int dir, const int32_t* item_sizes, int narrays,
             std::shared_ptr<CommPlan>& out);
// packed = true: the send buffer has already been filled (in-kernel pack of the FV shell kernel)
int run_plan(dgb_grid* g, CommPlan& plan, int myrank, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays,
             dgb_comm* comm, cudaStream_t stream, bool packed = false);
int nccl_allreduce(double* d, int op /*0 min, 1 max*/, dgb_comm* comm, cudaStream_t stream);
// rank-ordered concatenation of `bytes` bytes per rank (d_out holds nranks*bytes); one rank: a copy
int nccl_allgather_bytes(const void* d_in, void* d_out, size_t bytes, dgb_comm* comm, cudaStream_t stream);
int comm_rank_count(const dgb_comm* comm, int* rank, int* nranks);
// what the first unpack launch of a fused step waits for before it touches the receive buffer (dgb_fv.cu): the flags of
// all ranks reaching `target`; CTA 0 then also stores the maximum of the published CFL values into `slot`
struct HaloWait { struct FusedCtl* ctl; int nranks; unsigned target; int parity; double* slot; unsigned* hostTimeout; };
// recv boxes scattered from `buf` (laid out like the plan's receive buffer) instead of plan.dRecvBuf
int plan_unpack_from(CommPlan& plan, const double* buf, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays, cudaStream_t stream,
                     const HaloWait* wait = nullptr);

// scatter (set) a sub-list of the plan's receive boxes - `host` / `dev` hold the same boxes, disjoint - from `buf` into one array
int plan_unpack_boxes(CommPlan& plan, const std::vector<PlanBox>& host, PlanBox* dev, const double* buf, double* d_array, cudaStream_t stream);

// ---- fused exchange: the FV step kernel itself packs into the RECEIVERS' buffers through peer-mapped memory (NVLink) and
// signals completion with a flag; no NCCL call and no second stream in the steady state (dgb_fv.cu) -------------------
// One block per rank, shared with every other rank by CUDA IPC: control words + two receive buffers (by step parity).
struct FusedCtl {
  unsigned flags[64];        // flags[s] = number of steps whose data (and CFL value) rank s has delivered here
  unsigned counter;          // CTAs of the running step kernel that have finished (local)
  unsigned timeout;          // set by the wait kernel if a peer did not deliver in time
  unsigned pad[62];
  double cfl[2][64];         // cfl[parity][s] = local outflow maximum of rank s
};
// what the step kernel needs (device memory, one per parity)
struct FusedPack {
  const PlanBox* boxes;      // send boxes of the plan
  double* const* dst;        // dst[b] = first double of box b in its receiver's buffer of this parity (or in the receiver's field)
  const int* dstRow;         // registered fields: row / plane stride (doubles) of the receiver's stored box per send box;
  const int* dstPlane;       //   NULL: box-linear layout of the exchange buffer
  unsigned* counter;
  int nBoxes, nranks;
  unsigned* peerFlag[64];    // &ctl(r).flags[myrank]
  double* peerCfl[64];       // &ctl(r).cfl[parity][myrank]
};
struct FusedState {
  bool on = false;
  void* block = nullptr;
  FusedCtl* ctl = nullptr;
  double* recv[2] = {nullptr, nullptr};
  FusedPack* dPack = nullptr;            // [2]
  double** dDst[2] = {nullptr, nullptr};
  std::vector<void*> opened;
  FusedPack hostPack[2];                 // host copies of dPack[0..1]
  std::vector<double*> hostDst[2];       // host copies of dDst[0..1]
  // registered fields (dgb_fv_register_fields): peers store halo cells straight into these arrays, no receive buffer, no unpack
  // `hybrid`: like dPack, but the x-face boxes (one cell per row of the receiver: 8-byte stores into 32-byte sectors, a DRAM
  // read-modify-write each) keep going to the receiver's contiguous exchange buffer; the receiver's NEXT step kernel reads
  // its x neighbours from that buffer (dgb_fv.cu, FvArgs::xh) and its overlap column is refreshed only on demand
  struct Direct { double* field = nullptr; FusedPack* dPack = nullptr; double** dDst = nullptr; int* dRow = nullptr; int* dPlane = nullptr;
                  FusedPack* dHybrid = nullptr; double** dDstH[2] = {nullptr, nullptr}; int* dRowH = nullptr; int* dPlaneH = nullptr;
                  // host copies (the step kernels take the send boxes in their parameter space)
                  std::vector<double*> hDst, hDstH[2]; std::vector<int> hRow, hPlane, hRowH, hPlaneH; };
  // this rank's x-face receive boxes (low side, high side): where the step kernel finds the x halo in the exchange buffer
  struct XFace { bool on = false; long long buf_offset = 0; int y0 = 0, z0 = 0, ny = 0; };
  XFace xface[2];
  bool hybridOk = false;                 // every rank can use the hybrid form (same verdict everywhere)
  PlanBox* dXRecv = nullptr; int nXRecv = 0; std::vector<PlanBox> hostXRecv;      // the x-face receive boxes, for the on-demand unpack
  std::vector<Direct> direct;
  int direct_slot(const double* p) const { for (std::size_t k = 0; k < direct.size(); ++k) if (direct[k].field == p) return int(k); return -1; }
  void release_direct();
  unsigned epoch = 0;                    // fused steps done
  // sticky time-out word in mapped pinned host memory: the waiting kernels store 1 there when a peer did not deliver within
  // 20 s; the host reads it (no synchronisation needed) at the start of every later step and refuses to launch
  unsigned* hTimeout = nullptr;          // host address
  unsigned* dTimeout = nullptr;          // device alias of the same word
  bool multi = false;                    // the block is exported to other processes (teardown must be collective)
  bool timed_out() const { return hTimeout && *const_cast<volatile unsigned*>(hTimeout) != 0u; }
  // closes the peers' mappings; `free_block` = false leaves the exported block allocated (non-collective teardown: freeing a
  // block other processes may still have mapped is undefined behaviour, so it is deliberately leaked until process exit)
  void release(bool free_block = true);
  ~FusedState() { release(!multi); }
};
// collective over all ranks: allocate and publish this rank's block, map the peers'. Leaves fs.on = false (and DGB_OK) when
// peer mapping is not available on every rank - the caller then keeps the NCCL path.
// `local_ok` = false vetoes the fused form for ALL ranks (e.g. a rank without interior cells launches no step kernel and
// could never signal)
// `max_boxes`: most send boxes the user of the state can handle (the FV step kernels keep them in their parameter space)
int fused_setup(CommPlan& plan, int myrank, int nranks, dgb_comm* comm, cudaStream_t stream, FusedState& fs, bool local_ok = true, int max_boxes = 32);
// collective teardown of a fused state shared between processes: every rank closes its imported mappings, all ranks meet in
// a barrier (all-reduce on the communicator), only then the exported blocks are freed
int fused_teardown(FusedState& fs, dgb_comm* comm, cudaStream_t stream);
// collective: every rank publishes IPC handles of its `nfields` arrays (all indexed like the plan's mapper, one double per
// cell), maps the arrays of the ranks it sends to and prepares, per array, the destinations of its send boxes INSIDE the
// receivers' arrays. Leaves fs.direct empty (DGB_OK) if any rank cannot (all ranks agree).
int fused_register(CommPlan& plan, int myrank, int nranks, dgb_comm* comm, cudaStream_t stream, FusedState& fs,
                   double* const* fields, int nfields);
// a plan of its own for one user (not the per-grid cache): its buffers are then private to that user's stream
std::shared_ptr<CommPlan> build_plan(dgb_grid* g, int level, const dgb_mapper& m, int iftype, int dir, const int32_t* item_sizes, int narrays);
// ncclCommGetAsyncError of the communicator (DGB_OK without a communicator): DGB_ENCCL once a peer has failed
int nccl_check_async(dgb_comm* comm);
}
#endif
