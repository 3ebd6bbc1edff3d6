// dgb_comm.cu — GridView::communicate for mapper-indexed device arrays.
//
// Reference: YaspGrid::communicateCodim (dune/grid/yaspgrid.hh:1490-1730) allocates a buffer per neighbour box
// on every call, packs entity by entity through the data-handle facade, and exchanges with MPI_Isend/Irecv
// (torus.hh:387-440). Here an exchange PLAN is built once per (level, layout, interface, direction, item sizes):
// all send boxes of all codims are packed by ONE kernel into one contiguous device buffer (one segment per
// peer rank, message layout identical to the reference's: codim dim..0, boxes in list order, entities x-fastest,
// per entity array-major / block-major / item-minor), segments travel as one ncclSend/ncclRecv pair per peer
// inside a single group over NVLink, and ONE kernel scatters (set or add) every received segment. Messages
// to the own rank (periodic wrap on a 1-wide torus direction, torus.hh:395-403) are copied device-to-device.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <nccl.h>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <sstream>
#include "dgb_internal.hh"
#include "dgb_device.cuh"
#include "dgb_comm.hh"

namespace dgb {

// ---- NCCL through dlopen: libdgb.so must load on a CPU-only box and must share the process's NCCL ------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool load() {
    if (lib) return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { error = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define DGB_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) { error = std::string("libnccl lacks ") + name; lib = nullptr; return false; }
    DGB_SYM(GetUniqueId, "ncclGetUniqueId") DGB_SYM(CommInitRank, "ncclCommInitRank") DGB_SYM(CommDestroy, "ncclCommDestroy")
    DGB_SYM(Send, "ncclSend") DGB_SYM(Recv, "ncclRecv") DGB_SYM(AllReduce, "ncclAllReduce") DGB_SYM(AllGather, "ncclAllGather")
    DGB_SYM(GroupStart, "ncclGroupStart") DGB_SYM(GroupEnd, "ncclGroupEnd") DGB_SYM(GetErrorString, "ncclGetErrorString")
    DGB_SYM(CommGetAsyncError, "ncclCommGetAsyncError")
#undef DGB_SYM
    return true;
  }
};
static NcclApi g_nccl;

#define DGB_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) \
  return dgb::fail(DGB_ENCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); } while (0)

} // namespace dgb

struct dgb_comm { ncclComm_t comm; int rank, nranks; };

namespace dgb {

// ---- device side ------------------------------------------------------------------------------------------
struct ArrayTable { int n; int per_entity; double* ptr[8]; int item[8]; };

static constexpr int PACK_THREADS = 256;

// One launch serves a group of boxes: the 1-D grid is the concatenation of the boxes' CTA ranges (start[b]..start[b+1]),
// so a one-cell corner box costs one CTA and a face its ceil(count/256) (capped, grid-stride beyond). Thread t of a box
// handles buffer double t: entity e = t / per_entity (x fastest inside the box), then (array, block entry, item) in
// reference order. Index arithmetic is 32-bit whenever the box's doubles fit (they do for every BASELINE config).
static constexpr int HALO_MAX_GROUP = 128;
struct HaloGroup { int nboxes; int start[HALO_MAX_GROUP+1]; };

template<int MODE, typename IT>
__device__ __forceinline__ void halo_box(const PlanBox& b, const ArrayTable& arrays, double* __restrict__ buffer, IT first, IT step, IT total) {
  const IT per = IT(arrays.per_entity), s0 = IT(b.size[0]), s1 = IT(b.size[1]);
  for (IT t = first; t < total; t += step) {
    const IT e = per == 1 ? t : t / per;
    int r = per == 1 ? 0 : int(t - e*per);
    const IT q = e / s0; const int i0 = int(e - q*s0);
    const IT i2 = q / s1; const int i1 = int(q - i2*s1);
    const long long local = ((long long)(b.origin[2]+int(i2)-b.of_origin[2])*b.of_size[1] + (b.origin[1]+i1-b.of_origin[1]))*b.of_size[0]
                            + (b.origin[0]+i0-b.of_origin[0]);
    const unsigned idx = (b.index_base + unsigned(local))*b.block + b.map_offset;
    // (array a, entry r inside the array's block*item run)
    int a = 0;
    while (r >= int(b.block)*arrays.item[a]) { r -= int(b.block)*arrays.item[a]; ++a; }
    double* p = arrays.ptr[a] + (size_t)idx*arrays.item[a] + r;
    double* w = buffer + b.buf_offset + t;
    if (MODE == 0) *w = *p;
    else if (MODE == 1) *p = *w;
    else if (MODE == 2) *p = *w + *p;
    else if (MODE == 3) { const double x = *w; if (x >= 0.0) *p = fmin(x, *p); }      // MinimumExchange (globalindexset.hh:153-160)
    else { const double x = *w; if (x >= 0.0) *p = x; }                               // IndexExchange (globalindexset.hh:262-289)
  }
}

// the receiving side of the fused FV exchange, folded into the first unpack launch: thread s waits until rank s has
// delivered epoch `target` (a 20 s timeout marks a peer that never arrives instead of hanging the device), CTA 0 reduces
// the published CFL maxima
__device__ __forceinline__ bool halo_wait(const HaloWait& hw) {
  __shared__ int s_timed_out;
  const int s = threadIdx.x;
  if (s == 0) s_timed_out = 0;
  __syncthreads();
  if (s < hw.nranks) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned f;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&hw.ctl->flags[s]) : "memory");
      if (int(f - hw.target) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (hw.ctl->timeout || t1 - t0 > 20000000000ull) {              // sticky: also seen by the host before the next step
        hw.ctl->timeout = 1u; s_timed_out = 1;
        if (hw.hostTimeout) { *reinterpret_cast<volatile unsigned*>(hw.hostTimeout) = 1u; __threadfence_system(); }
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (s_timed_out) return false;                                      // never scatter a buffer a peer has not completed
  if (blockIdx.x == 0 && s == 0 && hw.slot) {
    double m = 0.0;
    for (int r = 0; r < hw.nranks; ++r) m = fmax(m, *reinterpret_cast<volatile double*>(&hw.ctl->cfl[hw.parity][r]));
    *hw.slot = m;
  }
  return true;
}

template<int MODE>   // 0 pack (gather), unpack: 1 set, 2 add, 3 min of non-negative, 4 set if non-negative
__global__ void __launch_bounds__(PACK_THREADS) k_halo(const PlanBox* __restrict__ boxes, const ArrayTable arrays, double* __restrict__ buffer,
                                                       const __grid_constant__ HaloGroup grp, const HaloWait hw) {
  if (hw.ctl && !halo_wait(hw)) return;
  int bi = 0;
  while (bi+1 < grp.nboxes && int(blockIdx.x) >= grp.start[bi+1]) ++bi;          // CTA-uniform
  const PlanBox& b = boxes[bi];
  const long long total = b.count*arrays.per_entity;
  const unsigned cta = blockIdx.x - unsigned(grp.start[bi]), nctas = unsigned(grp.start[bi+1]-grp.start[bi]);
  if (total < (1ll << 31) - (long long)nctas*PACK_THREADS)
    halo_box<MODE, unsigned>(b, arrays, buffer, cta*PACK_THREADS + threadIdx.x, nctas*PACK_THREADS, unsigned(total));
  else
    halo_box<MODE, long long>(b, arrays, buffer, (long long)cta*PACK_THREADS + threadIdx.x, (long long)nctas*PACK_THREADS, total);
}

// ---- communicate over peer-mapped memory: the pack kernel stores every send box straight into its place in the RECEIVER's buffer
// (CUDA IPC mapping of the partner's exchange block, the stores travel over NVLink), the last CTA of the exchange publishes the
// epoch in every rank's control block, and the first unpack launch waits for the flags of all ranks (halo_wait). Two launches,
// no NCCL call, no staging copy; the protocol (two receive buffers by exchange parity, flags = exchanges delivered) is the one
// of the fused FV step (dgb_fv.cu).
struct PeerSignal { const FusedPack* pack; unsigned total, epoch; };
__global__ void __launch_bounds__(PACK_THREADS) k_halo_peer(const PlanBox* __restrict__ boxes, double* const* __restrict__ dst, const ArrayTable arrays,
                                                            const __grid_constant__ HaloGroup grp, const PeerSignal sig) {
  int bi = 0;
  while (bi+1 < grp.nboxes && int(blockIdx.x) >= grp.start[bi+1]) ++bi;          // CTA-uniform
  const PlanBox& b = boxes[bi];
  double* buffer = dst[bi] - b.buf_offset;                                       // halo_box writes buffer + buf_offset + t
  const long long total = b.count*arrays.per_entity;
  const unsigned cta = blockIdx.x - unsigned(grp.start[bi]), nctas = unsigned(grp.start[bi+1]-grp.start[bi]);
  if (total < (1ll << 31) - (long long)nctas*PACK_THREADS)
    halo_box<0, unsigned>(b, arrays, buffer, cta*PACK_THREADS + threadIdx.x, nctas*PACK_THREADS, unsigned(total));
  else
    halo_box<0, long long>(b, arrays, buffer, (long long)cta*PACK_THREADS + threadIdx.x, (long long)nctas*PACK_THREADS, total);
  __syncthreads();                                   // the CTA's peer stores are ordered before thread 0's fence
  if (threadIdx.x == 0) {
    const FusedPack& P = *sig.pack;
    __threadfence_system();
    if (atomicAdd(P.counter, 1u) == sig.total-1) {   // the last CTA of the exchange (the counter runs over all its launches)
      atomicExch(P.counter, 0u);
      __threadfence_system();
      for (int r = 0; r < P.nranks; ++r) *reinterpret_cast<volatile unsigned*>(P.peerFlag[r]) = sig.epoch;
    }
  }
}

// a rank that sends nothing still publishes its epoch; a rank that receives nothing still waits for the others (every rank
// leaves exchange n only after all ranks have delivered it: the buffers of exchange n+2 may then be written)
__global__ void k_halo_signal(const PeerSignal sig) {
  if (threadIdx.x == 0) {
    const FusedPack& P = *sig.pack;
    __threadfence_system();
    for (int r = 0; r < P.nranks; ++r) *reinterpret_cast<volatile unsigned*>(P.peerFlag[r]) = sig.epoch;
  }
}
__global__ void __launch_bounds__(64) k_halo_wait_only(const HaloWait hw) { halo_wait(hw); }

// ---- data handles of variable size (CommDataHandleIF::fixedSize == false, yaspgrid.hh:1562-1634) ----------------------------------
// The sender decides how many items an entity carries. Bulk form: the send side is a CSR pair (offsets by mapper index, items);
// every message is preceded by a message of per-entity sizes (the reference's first exchange round); the receiver gets the
// sequence of scatter(buff, e, n) calls the reference would make - in its order - as (mapper index of e, offsets, items).
struct VarBuf {
  void* p = nullptr;
  ~VarBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) { if (p) { cudaFree(p); p = nullptr; } DGB_CUDA(cudaMalloc(&p, bytes ? bytes : 8)); return DGB_OK; }
  template<class T> T* as() const { return static_cast<T*>(p); }
};

// one thread per entity slot of a box: the mapper index of the entity (block 1), its size on the send side, its place in the
// reference's scatter order on the receive side
__global__ void __launch_bounds__(PACK_THREADS) k_var_slots(const PlanBox b, const long long* __restrict__ csr, long long* __restrict__ slotIndex,
                                                            long long* __restrict__ slotSize, long long* __restrict__ slotOut, long long outBase) {
  const long long s0 = b.size[0], s1 = b.size[1];
  for (long long t = (long long)blockIdx.x*PACK_THREADS + threadIdx.x; t < b.count; t += (long long)gridDim.x*PACK_THREADS) {
    const long long q = t / s0; const int i0 = int(t - q*s0);
    const long long i2 = q / s1; const int i1 = int(q - i2*s1);
    const long long local = ((long long)(b.origin[2]+int(i2)-b.of_origin[2])*b.of_size[1] + (b.origin[1]+i1-b.of_origin[1]))*b.of_size[0]
                            + (b.origin[0]+i0-b.of_origin[0]);
    const unsigned idx = (b.index_base + unsigned(local))*b.block + b.map_offset;
    slotIndex[b.buf_offset + t] = idx;
    if (slotSize) slotSize[b.buf_offset + t] = csr[idx+1] - csr[idx];
    if (slotOut) slotOut[b.buf_offset + t] = outBase + t;
  }
}

// one warp per slot: n[slot] doubles from src[srcTable[srcSel ? srcSel[slot] : slot] ...] to dst[dstTable[dstSel ? dstSel[slot] : slot] ...]
__global__ void __launch_bounds__(PACK_THREADS) k_var_copy(long long nslots, const long long* __restrict__ n, const double* __restrict__ src,
                                                           const long long* __restrict__ srcTable, const long long* __restrict__ srcSel,
                                                           double* __restrict__ dst, const long long* __restrict__ dstTable, const long long* __restrict__ dstSel) {
  const int lane = threadIdx.x & 31;
  for (long long w = ((long long)blockIdx.x*PACK_THREADS + threadIdx.x) >> 5; w < nslots; w += ((long long)gridDim.x*PACK_THREADS) >> 5) {
    const long long cnt = n[w];
    const double* sp = src + srcTable[srcSel ? srcSel[w] : w];
    double* dp = dst + dstTable[dstSel ? dstSel[w] : w];
    for (long long j = lane; j < cnt; j += 32) dp[j] = sp[j];
  }
}

__global__ void __launch_bounds__(PACK_THREADS) k_var_permute(long long nslots, const long long* __restrict__ pos, const long long* __restrict__ idx,
                                                              const long long* __restrict__ size, long long* __restrict__ outIndex, long long* __restrict__ outSize) {
  for (long long t = (long long)blockIdx.x*PACK_THREADS + threadIdx.x; t < nslots; t += (long long)gridDim.x*PACK_THREADS) {
    outIndex[pos[t]] = idx[t]; outSize[pos[t]] = size[t];
  }
}

__global__ void __launch_bounds__(PACK_THREADS) k_var_index32(long long n, const long long* __restrict__ in, uint32_t* __restrict__ out) {
  for (long long t = (long long)blockIdx.x*PACK_THREADS + threadIdx.x; t < n; t += (long long)gridDim.x*PACK_THREADS) out[t] = uint32_t(in[t]);
}

// ---- plan construction (host) -------------------------------------------------------------------------
static void add_boxes(const dgb_grid* g, int level, int codim, int which, const dgb_mapper& m, int per_entity_codim,
                      std::vector<PlanBox>& boxes, std::vector<int>& peer) {
  std::vector<ListEntry> l;
  comm_list(g->me, g->all, level, codim, which, l);
  const LevelBoxes& lb = g->me.levels[level].b;
  for (const ListEntry& e : l) {
    PlanBox b;
    const dgb_component& c = lb.comp[lb.comp_begin[codim] + lb.shiftmap[e.shift]];
    const dgb_box& of = c.box[DGB_BOX_OVERLAPFRONT];
    for (int i = 0; i < 3; ++i) { b.origin[i] = e.box.origin[i]; b.size[i] = e.box.size[i]; b.of_origin[i] = of.origin[i]; b.of_size[i] = of.size[i]; }
    b.count = 1; for (int i = 0; i < lb.dim; ++i) b.count *= e.box.size[i];
    for (int i = lb.dim; i < 3; ++i) { b.size[i] = 1; b.of_size[i] = 1; b.origin[i] = 0; b.of_origin[i] = 0; }
    b.index_base = unsigned(e.index_offset);
    b.block = m.block[codim]; b.map_offset = m.offset[codim];
    {
      const LevelBoxes& pl = g->all[e.rank].levels[level].b;
      const dgb_box& pof = pl.comp[pl.comp_begin[codim] + pl.shiftmap[e.shift]].box[DGB_BOX_OVERLAPFRONT];
      for (int i = 0; i < 3; ++i) {
        b.peer_origin[i] = i < lb.dim ? e.box.origin[i] - e.move[i] : 0;
        b.peer_of_origin[i] = i < lb.dim ? pof.origin[i] : 0; b.peer_of_size[i] = i < lb.dim ? pof.size[i] : 1;
      }
      b.peer = e.rank;
      int nz = 0, ax = -1;
      for (int i = 0; i < lb.dim; ++i) if (e.delta[i] != 0) { ++nz; ax = i; }
      b.face = (nz == 1) ? 2*ax + (e.delta[ax] > 0 ? 1 : 0) : -1;
    }
    b.buf_offset = 0;
    b.list_pos = int(boxes.size());
    (void)per_entity_codim;
    boxes.push_back(b);
    peer.push_back(e.rank);
  }
}

std::shared_ptr<CommPlan> build_plan(dgb_grid* g, int level, const dgb_mapper& m, int iftype, int dir,
                                     const int32_t* item_sizes, int narrays) {
  auto plan = std::make_shared<CommPlan>();
  const int dim = g->me.spec.dim;
  int sendWhich, recvWhich;
  switch (iftype) {                                   // yaspgrid.hh:1506-1525
    case DGB_InteriorBorder_InteriorBorder_Interface: sendWhich = 4; recvWhich = 5; break;
    case DGB_InteriorBorder_All_Interface: sendWhich = 6; recvWhich = 7; break;
    case DGB_Overlap_OverlapFront_Interface: case DGB_Overlap_All_Interface: sendWhich = 2; recvWhich = 3; break;
    case DGB_All_All_Interface: sendWhich = 0; recvWhich = 1; break;
    default: throw ArgError("unknown interface type");
  }
  if (dir == DGB_BackwardCommunication) std::swap(sendWhich, recvWhich);       // :1528-1529
  else if (dir != DGB_ForwardCommunication) throw ArgError("unknown communication direction");
  int items = 0; for (int a = 0; a < narrays; ++a) items += item_sizes[a];
  std::vector<PlanBox> sb, rb; std::vector<int> sp, rp;
  for (int codim = dim; codim >= 0; --codim) {        // YaspCommunicateMeta: codim dim..0 (:122-143)
    if (m.block[codim] == 0) continue;                // DataHandle::contains
    add_boxes(g, level, codim, sendWhich, m, items, sb, sp);
    add_boxes(g, level, codim, recvWhich, m, items, rb, rp);
  }
  // group by peer (stable): one contiguous segment per peer, box order inside a segment = list order
  auto layout = [&](std::vector<PlanBox>& boxes, const std::vector<int>& peer, std::vector<CommPlan::Segment>& segs, long long& total) {
    std::vector<int> peers(peer);
    std::sort(peers.begin(), peers.end()); peers.erase(std::unique(peers.begin(), peers.end()), peers.end());
    long long off = 0;
    for (int p : peers) {
      CommPlan::Segment s; s.peer = p; s.offset = off;
      for (std::size_t k = 0; k < boxes.size(); ++k) if (peer[k] == p) {
        boxes[k].buf_offset = off;
        off += boxes[k].count*(long long)boxes[k].block*items;
      }
      s.count = off - s.offset;
      segs.push_back(s);
    }
    total = off;
  };
  layout(sb, sp, plan->sendSeg, plan->sendTotal);
  layout(rb, rp, plan->recvSeg, plan->recvTotal);
  plan->nSendBoxes = int(sb.size()); plan->nRecvBoxes = int(rb.size());
  plan->maxSendBox = 0; plan->maxRecvBox = 0;
  for (auto& b : sb) plan->maxSendBox = std::max(plan->maxSendBox, b.count*(long long)b.block*items);
  for (auto& b : rb) plan->maxRecvBox = std::max(plan->maxRecvBox, b.count*(long long)b.block*items);
  plan->items = items;
  // Unpack hazards: the reference scatters recv boxes one after another (yaspgrid.hh:1694-1729), so when two boxes
  // of one component overlap (border/front entities shared by several neighbours) the later one wins (set) or the
  // sums accumulate in list order (add). Boxes are therefore scheduled in "waves": a box runs one wave after the
  // last earlier box it overlaps; boxes of one wave are disjoint and share a launch.
  std::vector<int> wave(rb.size(), 0);
  auto overlaps = [&](const PlanBox& a, const PlanBox& b) {
    if (a.index_base != b.index_base || a.map_offset != b.map_offset) return false;       // different component / codim
    for (int i = 0; i < 3; ++i)
      if (a.origin[i]+a.size[i] <= b.origin[i] || b.origin[i]+b.size[i] <= a.origin[i]) return false;
    return true;
  };
  int nwaves = rb.empty() ? 0 : 1;
  for (std::size_t k = 0; k < rb.size(); ++k)
    for (std::size_t j = 0; j < k; ++j)
      if (overlaps(rb[j], rb[k])) { wave[k] = std::max(wave[k], wave[j]+1); nwaves = std::max(nwaves, wave[k]+1); }
  std::vector<PlanBox> ordered;
  plan->recvWaveEnd.clear();
  for (int w = 0; w < nwaves; ++w) {
    for (std::size_t k = 0; k < rb.size(); ++k) if (wave[k] == w) ordered.push_back(rb[k]);
    plan->recvWaveEnd.push_back(int(ordered.size()));
  }
  plan->hostSend = sb; plan->hostRecv = ordered;
  for (std::size_t r = 0; r < g->all.size() && !plan->anyRemote; ++r) {
    std::vector<ListEntry> l;
    for (int codim = dim; codim >= 0 && !plan->anyRemote; --codim) {
      if (m.block[codim] == 0) continue;
      for (int which : { sendWhich, recvWhich }) {
        comm_list(g->all[r], g->all, level, codim, which, l);
        for (const ListEntry& e : l) if (e.rank != int(r)) plan->anyRemote = true;
      }
    }
  }
  return plan;
}

int CommPlan::upload() {
  if (uploaded) return DGB_OK;
  if (int rc = require_device()) return rc;
  if (nSendBoxes) { DGB_CUDA(cudaMalloc(&dSendBoxes, nSendBoxes*sizeof(PlanBox)));
                    DGB_CUDA(cudaMemcpy(dSendBoxes, hostSend.data(), nSendBoxes*sizeof(PlanBox), cudaMemcpyHostToDevice)); }
  if (nRecvBoxes) { DGB_CUDA(cudaMalloc(&dRecvBoxes, nRecvBoxes*sizeof(PlanBox)));
                    DGB_CUDA(cudaMemcpy(dRecvBoxes, hostRecv.data(), nRecvBoxes*sizeof(PlanBox), cudaMemcpyHostToDevice)); }
  if (sendTotal) DGB_CUDA(cudaMalloc(&dSendBuf, sendTotal*sizeof(double)));
  if (recvTotal) DGB_CUDA(cudaMalloc(&dRecvBuf, recvTotal*sizeof(double)));
  uploaded = true;
  return DGB_OK;
}
CommPlan::~CommPlan() {
  if (busy) cudaEventDestroy(busy);
  if (dSendBoxes) cudaFree(dSendBoxes);
  if (dRecvBoxes) cudaFree(dRecvBoxes);
  if (dSendBuf) cudaFree(dSendBuf);
  if (dRecvBuf) cudaFree(dRecvBuf);
}

// CTAs of one box: enough to cover it, capped (grid-stride beyond; DGB_HALO_MAXBLOCKS overrides the cap)
static int halo_ctas(long long doubles) {
  static const long long cap = [] { const char* e = getenv("DGB_HALO_MAXBLOCKS"); long long v = e ? atoll(e) : 2048; return v > 0 ? v : 2048; }();
  long long bx = (doubles + PACK_THREADS - 1)/PACK_THREADS;
  return int(std::max(1ll, std::min(bx, cap)));
}

static int make_table(const CommPlan& plan, double* const* d_arrays, const int32_t* item_sizes, int narrays, ArrayTable& at) {
  if (narrays > 8) return fail(DGB_EARG, "at most 8 arrays per communicate call");
  at.n = narrays; at.per_entity = 0;
  for (int a = 0; a < narrays; ++a) { at.ptr[a] = d_arrays[a]; at.item[a] = item_sizes[a]; }
  for (int a = narrays; a < 8; ++a) { at.ptr[a] = nullptr; at.item[a] = 1 << 30; }
  (void)plan;
  return DGB_OK;
}

// boxes of one launch share the mapper block size (per-entity run length) and, for unpack, the wave
static void launch_boxes(CommPlan& plan, const ArrayTable& at, int mode, const std::vector<PlanBox>& host, PlanBox* dev, double* buf,
                         const std::vector<int>* waveEnd, cudaStream_t stream, const HaloWait* wait = nullptr) {
  std::size_t k = 0, wv = 0;
  HaloWait hw{nullptr, 0, 0u, 0, nullptr, nullptr};
  if (wait) hw = *wait;                                   // only the first launch waits
  while (k < host.size()) {
    std::size_t limit = host.size();
    if (waveEnd) { while (wv < waveEnd->size() && std::size_t((*waveEnd)[wv]) <= k) ++wv; limit = std::size_t((*waveEnd)[wv]); }
    HaloGroup grp; grp.nboxes = 0; grp.start[0] = 0;
    std::size_t e = k;
    while (e < limit && host[e].block == host[k].block && grp.nboxes < HALO_MAX_GROUP) {
      grp.start[grp.nboxes+1] = grp.start[grp.nboxes] + halo_ctas(host[e].count*(long long)host[e].block*plan.items);
      ++grp.nboxes; ++e;
    }
    ArrayTable t = at; t.per_entity = int(host[k].block)*plan.items;
    const unsigned grid = unsigned(grp.start[grp.nboxes]);
    if (mode == 0) k_halo<0><<<grid, PACK_THREADS, 0, stream>>>(dev+k, t, buf, grp, hw);
    else if (mode == 1) k_halo<1><<<grid, PACK_THREADS, 0, stream>>>(dev+k, t, buf, grp, hw);
    else if (mode == 2) k_halo<2><<<grid, PACK_THREADS, 0, stream>>>(dev+k, t, buf, grp, hw);
    else if (mode == 3) k_halo<3><<<grid, PACK_THREADS, 0, stream>>>(dev+k, t, buf, grp, hw);
    else k_halo<4><<<grid, PACK_THREADS, 0, stream>>>(dev+k, t, buf, grp, hw);
    hw.ctl = nullptr;
    ++plan.launches;
    k = e;
  }
}

int plan_pack(CommPlan& plan, double* const* d_arrays, const int32_t* item_sizes, int narrays, cudaStream_t stream) {
  if (int rc = plan.upload()) return rc;
  ArrayTable at; if (int rc = make_table(plan, d_arrays, item_sizes, narrays, at)) return rc;
  if (plan.nSendBoxes) launch_boxes(plan, at, 0, plan.hostSend, plan.dSendBoxes, plan.dSendBuf, nullptr, stream);
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

int plan_unpack(CommPlan& plan, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays, cudaStream_t stream) {
  if (int rc = plan.upload()) return rc;
  ArrayTable at; if (int rc = make_table(plan, d_arrays, item_sizes, narrays, at)) return rc;
  if (plan.nRecvBoxes) launch_boxes(plan, at, op + 1, plan.hostRecv, plan.dRecvBoxes, plan.dRecvBuf, &plan.recvWaveEnd, stream);
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

static int run_plan_peer(dgb_grid* g, CommPlan& plan, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays, cudaStream_t stream) {
  FusedState& fs = *plan.peer;
  if (fs.timed_out()) return fail(DGB_ENCCL, "communicate: a rank did not deliver an earlier exchange within 20 s (ranks out of step, or a peer has died)");
  if (int rc = plan.upload()) return rc;
  ArrayTable at; if (int rc = make_table(plan, d_arrays, item_sizes, narrays, at)) return rc;
  const unsigned epoch = ++fs.epoch;
  const int parity = int(epoch & 1u);
  // pack: the launches of launch_boxes with the peer kernel; the counter of "CTAs done" runs over all of them
  struct Launch { std::size_t first; HaloGroup grp; int per; };
  std::vector<Launch> launches;
  unsigned total = 0;
  for (std::size_t k = 0; k < plan.hostSend.size();) {
    Launch L; L.first = k; L.grp.nboxes = 0; L.grp.start[0] = 0; L.per = int(plan.hostSend[k].block)*plan.items;
    std::size_t e = k;
    while (e < plan.hostSend.size() && plan.hostSend[e].block == plan.hostSend[k].block && L.grp.nboxes < HALO_MAX_GROUP) {
      L.grp.start[L.grp.nboxes+1] = L.grp.start[L.grp.nboxes] + halo_ctas(plan.hostSend[e].count*(long long)plan.hostSend[e].block*plan.items);
      ++L.grp.nboxes; ++e;
    }
    total += unsigned(L.grp.start[L.grp.nboxes]);
    launches.push_back(L);
    k = e;
  }
  int64_t bytes = 0;
  for (auto& sg : plan.sendSeg) if (sg.peer != g->me.rank) bytes += sg.count*8;
  PeerSignal sig{fs.dPack + parity, total, epoch};
  if (launches.empty()) {
    // a rank without anything to send still has to publish its epoch: one CTA, no box
    HaloGroup grp; grp.nboxes = 0; grp.start[0] = 0;
    (void)grp;
    sig.total = 1;
    k_halo_signal<<<1, 32, 0, stream>>>(sig);
    ++plan.launches;
  }
  for (auto& L : launches) {
    ArrayTable t = at; t.per_entity = L.per;
    k_halo_peer<<<unsigned(L.grp.start[L.grp.nboxes]), PACK_THREADS, 0, stream>>>(plan.dSendBoxes + L.first, fs.dDst[parity] + L.first, t, L.grp, sig);
    ++plan.launches;
  }
  DGB_CUDA(cudaGetLastError());
  if (plan.trace[0]) cudaEventRecord(plan.trace[0], stream);
  g->lastBytesSent = bytes;
  if (plan.trace[1]) cudaEventRecord(plan.trace[1], stream);
  HaloWait hw{fs.ctl, fs.hostPack[0].nranks, epoch, parity, nullptr, fs.dTimeout};
  int rc = DGB_OK;
  if (plan.nRecvBoxes) rc = plan_unpack_from(plan, fs.recv[parity], op, d_arrays, item_sizes, narrays, stream, &hw);
  else { k_halo_wait_only<<<1, 64, 0, stream>>>(hw); ++plan.launches; }       // stay in step with the ranks that do receive
  if (plan.trace[2]) cudaEventRecord(plan.trace[2], stream);
  DGB_CUDA(cudaGetLastError());
  return rc;
}

// the exchange itself; `arrays` are device pointers
int run_plan(dgb_grid* g, CommPlan& plan, int myrank, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays,
             dgb_comm* comm, cudaStream_t stream, bool packed) {
  bool remote = false;
  for (auto& s : plan.sendSeg) if (s.peer != myrank) remote = true;
  for (auto& s : plan.recvSeg) if (s.peer != myrank) remote = true;
  if (remote && !comm) return fail(DGB_EARG, "this exchange has remote partners: a dgb_comm is required");
  if (remote) if (int rc = nccl_check_async(comm)) return rc;
  if (int rc = plan.enter(stream)) return rc;
  if (!packed) {
    // peer-memory form (default; DGB_COMM_PEER=0 keeps the NCCL form): set up collectively at the first use of the plan - every
    // rank runs the same sequence of communicate calls (the contract of the call), so every rank is here with the same plan
    static const bool wantPeer = [] { const char* e = getenv("DGB_COMM_PEER"); return !e || atoi(e) != 0; }();
    // (a single-rank periodic exchange keeps the copy form: measured 0.049 vs 0.057 ms for two 1 M-cell faces, nothing travels)
    if (remote && !plan.anyRemote) return fail(DGB_EGRID, "communicate: the ranks disagree on the partners of this exchange");
    if (plan.anyRemote && !comm) return fail(DGB_EARG, "this exchange has remote partners on some rank: a dgb_comm is required on every rank");
    if (wantPeer && plan.anyRemote && comm && !plan.peerTried) {
      plan.peerTried = true;
      int nranks = 1; if (comm) nranks = comm->nranks;
      auto fs = std::make_shared<FusedState>();
      if (int rc = fused_setup(plan, myrank, nranks, comm, stream, *fs, true, 1 << 20)) return rc;
      if (fs->on) plan.peer = fs;
    }
    if (plan.peer && plan.peer->on) {
      const int rc = run_plan_peer(g, plan, op, d_arrays, item_sizes, narrays, stream);
      if (int rc2 = plan.leave(stream)) return rc2;
      return rc;
    }
  }
  if (packed) { if (int rc = plan.upload()) return rc; }
  else if (int rc = plan_pack(plan, d_arrays, item_sizes, narrays, stream)) return rc;
  if (plan.trace[0]) cudaEventRecord(plan.trace[0], stream);
  int64_t bytes = 0;
  if (remote) {
    DGB_NCCL(g_nccl.GroupStart());
    for (auto& s : plan.sendSeg) if (s.peer != myrank) { DGB_NCCL(g_nccl.Send(plan.dSendBuf + s.offset, s.count, ncclDouble, s.peer, comm->comm, stream)); bytes += s.count*8; }
    for (auto& s : plan.recvSeg) if (s.peer != myrank) DGB_NCCL(g_nccl.Recv(plan.dRecvBuf + s.offset, s.count, ncclDouble, s.peer, comm->comm, stream));
    DGB_NCCL(g_nccl.GroupEnd());
  }
  for (auto& s : plan.sendSeg) if (s.peer == myrank) {
    for (auto& r : plan.recvSeg) if (r.peer == myrank) {
      if (r.count != s.count) return fail(DGB_EGRID, "local sends/receives do not match in exchange");
      DGB_CUDA(cudaMemcpyAsync(plan.dRecvBuf + r.offset, plan.dSendBuf + s.offset, s.count*sizeof(double), cudaMemcpyDeviceToDevice, stream));
    }
  }
  g->lastBytesSent = bytes;
  if (plan.trace[1]) cudaEventRecord(plan.trace[1], stream);
  const int rcu = plan_unpack(plan, op, d_arrays, item_sizes, narrays, stream);
  if (plan.trace[2]) cudaEventRecord(plan.trace[2], stream);
  if (int rc = plan.leave(stream)) return rc;
  return rcu;
}

int plan_unpack_from(CommPlan& plan, const double* buf, int op, double* const* d_arrays, const int32_t* item_sizes, int narrays, cudaStream_t stream,
                     const HaloWait* wait) {
  if (int rc = plan.upload()) return rc;
  ArrayTable at; if (int rc = make_table(plan, d_arrays, item_sizes, narrays, at)) return rc;
  if (plan.nRecvBoxes) launch_boxes(plan, at, op + 1, plan.hostRecv, plan.dRecvBoxes, const_cast<double*>(buf), &plan.recvWaveEnd, stream, wait);
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

int plan_unpack_boxes(CommPlan& plan, const std::vector<PlanBox>& host, PlanBox* dev, const double* buf, double* d_array, cudaStream_t stream) {
  if (host.empty()) return DGB_OK;
  const int32_t items[1] = {1};
  double* arrays[1] = { d_array };
  ArrayTable at; if (int rc = make_table(plan, arrays, items, 1, at)) return rc;
  launch_boxes(plan, at, DGB_OP_SET + 1, host, dev, const_cast<double*>(buf), nullptr, stream);
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

// ---- fused exchange over peer-mapped memory: set-up ------------------------------------------------------------------
// what every rank publishes: the IPC handle of its block and, per sender rank, the offset (in doubles) of that sender's
// segment in the receive buffers (-1: no message from that rank)
struct FusedInfo { cudaIpcMemHandle_t handle; long long segOffset[64]; long long segCount[64]; long long recvTotal; int ok; int pad; };

void FusedState::release_direct() {
  for (auto& d : direct) {
    if (d.dPack) cudaFree(d.dPack); if (d.dDst) cudaFree(d.dDst); if (d.dRow) cudaFree(d.dRow); if (d.dPlane) cudaFree(d.dPlane);
    if (d.dHybrid) cudaFree(d.dHybrid); if (d.dRowH) cudaFree(d.dRowH); if (d.dPlaneH) cudaFree(d.dPlaneH);
    for (int k = 0; k < 2; ++k) if (d.dDstH[k]) cudaFree(d.dDstH[k]);
  }
  direct.clear();
  if (dXRecv) cudaFree(dXRecv);
  dXRecv = nullptr; nXRecv = 0; hostXRecv.clear(); hybridOk = false; xface[0] = XFace(); xface[1] = XFace();
}

void FusedState::release(bool free_block) {
  release_direct();
  for (void* p : opened) cudaIpcCloseMemHandle(p);
  opened.clear();
  for (int k = 0; k < 2; ++k) { if (dDst[k]) cudaFree(dDst[k]); dDst[k] = nullptr; }
  if (dPack) cudaFree(dPack);
  if (block && free_block) cudaFree(block);
  if (hTimeout) cudaFreeHost(hTimeout);
  dPack = nullptr; block = nullptr; ctl = nullptr; recv[0] = recv[1] = nullptr; hTimeout = dTimeout = nullptr; on = false; multi = false;
}

int fused_teardown(FusedState& fs, dgb_comm* comm, cudaStream_t stream) {
  const bool collective = fs.multi && comm && comm->nranks > 1;
  if (!fs.block && !collective) { fs.release(); return DGB_OK; }
  cudaStreamSynchronize(stream);
  for (void* p : fs.opened) cudaIpcCloseMemHandle(p);              // 1. nobody keeps a mapping of a peer's block ...
  fs.opened.clear();
  int rc = DGB_OK;
  if (collective) {                                                // 2. ... and every rank knows that ...
    double* d = nullptr;
    if (cudaMalloc(&d, sizeof(double)) == cudaSuccess) {
      cudaMemsetAsync(d, 0, sizeof(double), stream);
      if (g_nccl.AllReduce(d, d, 1, ncclDouble, ncclMax, comm->comm, stream) != ncclSuccess) rc = DGB_ENCCL;
      cudaStreamSynchronize(stream);
      cudaFree(d);
    } else rc = DGB_ECUDA;
  }
  fs.release(rc == DGB_OK);                                         // 3. ... before any exported block is freed
  return rc;
}

int fused_setup(CommPlan& plan, int myrank, int nranks, dgb_comm* comm, cudaStream_t stream, FusedState& fs, bool local_ok, int max_boxes) {
  fs.on = false;
  if (nranks > 64) return DGB_OK;                              // FusedCtl capacity (same verdict on every rank)
  const bool multi = nranks > 1;
  if (multi && !comm) return DGB_OK;
  // Every rank takes part in BOTH collectives below (all-gather of the handles, all-reduce of the verdict) whatever fails
  // locally: a local failure only turns this rank's vote into "no", it never leaves the peers blocked in a collective.
  bool ok = local_ok && plan.nSendBoxes <= max_boxes;          // FUSED_MAX_BOXES of the step kernels
  std::string why;
  auto soft = [&](cudaError_t e, const char* what) { if (e != cudaSuccess) { cudaGetLastError(); if (ok) why = std::string(what) + ": " + cudaGetErrorString(e); ok = false; } return e == cudaSuccess; };
  if (plan.upload() != DGB_OK) ok = false;
  auto round256 = [](size_t b) { return (b + 255)/256*256; };
  const size_t ctlBytes = round256(sizeof(FusedCtl));
  auto buf_bytes = [&](long long doubles) { return round256(size_t(doubles)*sizeof(double)); };
  const size_t blockBytes = ctlBytes + 2*buf_bytes(plan.recvTotal);
  if (soft(cudaMalloc(&fs.block, blockBytes), "cudaMalloc(fused block)")) soft(cudaMemset(fs.block, 0, blockBytes), "cudaMemset(fused block)");
  else fs.block = nullptr;
  if (soft(cudaHostAlloc(reinterpret_cast<void**>(&fs.hTimeout), sizeof(unsigned), cudaHostAllocMapped), "cudaHostAlloc(timeout word)")) {
    *fs.hTimeout = 0u;
    soft(cudaHostGetDevicePointer(reinterpret_cast<void**>(&fs.dTimeout), fs.hTimeout, 0), "cudaHostGetDevicePointer");
  } else fs.hTimeout = nullptr;
  fs.multi = multi;
  fs.ctl = static_cast<FusedCtl*>(fs.block);
  if (fs.block)
    for (int k = 0; k < 2; ++k) fs.recv[k] = reinterpret_cast<double*>(static_cast<char*>(fs.block) + ctlBytes + k*buf_bytes(plan.recvTotal));
  std::vector<FusedInfo> all(nranks);
  FusedInfo mine; std::memset(&mine, 0, sizeof(mine));
  mine.recvTotal = plan.recvTotal;
  for (int r = 0; r < 64; ++r) { mine.segOffset[r] = -1; mine.segCount[r] = 0; }
  for (auto& s : plan.recvSeg) { mine.segOffset[s.peer] = s.offset; mine.segCount[s.peer] = s.count; }
  if (multi && fs.block) soft(cudaIpcGetMemHandle(&mine.handle, fs.block), "cudaIpcGetMemHandle");
  mine.ok = ok ? 1 : 0;
  int hard = DGB_OK;                                           // failures of the collectives themselves: nothing left to agree on
  if (multi) {
    FusedInfo* dAll = nullptr;
    if (cudaMalloc(&dAll, sizeof(FusedInfo)*nranks) != cudaSuccess) { cudaGetLastError(); hard = fail(DGB_ECUDA, "fused_setup: cudaMalloc of the handle table"); }
    else {
      cudaMemcpyAsync(dAll + myrank, &mine, sizeof(FusedInfo), cudaMemcpyHostToDevice, stream);
      if (g_nccl.AllGather(dAll + myrank, dAll, sizeof(FusedInfo), ncclChar, comm->comm, stream) != ncclSuccess) hard = fail(DGB_ENCCL, "fused_setup: ncclAllGather");
      cudaMemcpyAsync(all.data(), dAll, sizeof(FusedInfo)*nranks, cudaMemcpyDeviceToHost, stream);
      if (cudaStreamSynchronize(stream) != cudaSuccess && hard == DGB_OK) hard = fail(DGB_ECUDA, "fused_setup: exchange of the IPC handles");
      cudaFree(dAll);
    }
    if (hard != DGB_OK) { fs.release(false); return hard; }
  } else all[0] = mine;
  for (int r = 0; r < nranks; ++r) ok = ok && all[r].ok;
  // map every rank's block: flags and CFL values go to all ranks, field data to the ranks this rank sends to
  std::vector<char*> base(nranks, nullptr);
  for (int r = 0; r < nranks && ok; ++r) {
    if (r == myrank) { base[r] = static_cast<char*>(fs.block); continue; }
    void* p = nullptr;
    if (!soft(cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) break;
    fs.opened.push_back(p);
    base[r] = static_cast<char*>(p);
  }
  if (ok)
    for (auto& s : plan.sendSeg)
      if (all[s.peer].segOffset[myrank] < 0 || all[s.peer].segCount[myrank] != s.count) ok = false;
  // destination of every send box: receiver's buffer + offset of my segment there + offset of the box inside my segment
  FusedPack pk[2];
  for (int k = 0; k < 2 && ok; ++k) {
    std::vector<double*> dst(plan.nSendBoxes);
    for (int b = 0; b < plan.nSendBoxes && ok; ++b) {
      const PlanBox& pb = plan.hostSend[b];
      const CommPlan::Segment* seg = nullptr;
      for (auto& s : plan.sendSeg) if (pb.buf_offset >= s.offset && pb.buf_offset < s.offset + s.count) seg = &s;
      if (!seg) { ok = false; why = "send box outside every segment"; break; }
      char* peerBuf = base[seg->peer] + ctlBytes + k*buf_bytes(all[seg->peer].recvTotal);
      dst[b] = reinterpret_cast<double*>(peerBuf) + all[seg->peer].segOffset[myrank] + (pb.buf_offset - seg->offset);
    }
    if (ok && plan.nSendBoxes) {
      if (soft(cudaMalloc(&fs.dDst[k], plan.nSendBoxes*sizeof(double*)), "cudaMalloc(box destinations)"))
        soft(cudaMemcpy(fs.dDst[k], dst.data(), plan.nSendBoxes*sizeof(double*), cudaMemcpyHostToDevice), "cudaMemcpy(box destinations)");
    }
    std::memset(&pk[k], 0, sizeof(FusedPack));
    pk[k].boxes = plan.dSendBoxes; pk[k].dst = fs.dDst[k]; pk[k].counter = fs.ctl ? &fs.ctl->counter : nullptr;
    pk[k].nBoxes = plan.nSendBoxes; pk[k].nranks = nranks;
    for (int r = 0; r < nranks && ok; ++r) {
      FusedCtl* c = reinterpret_cast<FusedCtl*>(base[r]);
      pk[k].peerFlag[r] = &c->flags[myrank];
      pk[k].peerCfl[r] = &c->cfl[k][myrank];
    }
  }
  if (ok && soft(cudaMalloc(&fs.dPack, 2*sizeof(FusedPack)), "cudaMalloc(FusedPack)"))
    soft(cudaMemcpy(fs.dPack, pk, 2*sizeof(FusedPack), cudaMemcpyHostToDevice), "cudaMemcpy(FusedPack)");
  fs.hostPack[0] = pk[0]; fs.hostPack[1] = pk[1];
  for (int k = 0; k < 2; ++k) {                       // host copy of the buffered destinations (hybrid form of registered fields)
    fs.hostDst[k].assign(plan.nSendBoxes, nullptr);
    if (ok && plan.nSendBoxes) cudaMemcpy(fs.hostDst[k].data(), fs.dDst[k], plan.nSendBoxes*sizeof(double*), cudaMemcpyDeviceToHost);
  }
  if (multi) {                                           // every rank must reach the same verdict: min over ranks of 1/0
    double v = ok ? 1.0 : 0.0;
    double* dScratch = nullptr;
    if (cudaMalloc(&dScratch, sizeof(double)) != cudaSuccess) { cudaGetLastError(); fs.release(false); return fail(DGB_ECUDA, "fused_setup: cudaMalloc of the verdict word"); }
    cudaMemcpyAsync(dScratch, &v, sizeof(double), cudaMemcpyHostToDevice, stream);
    const bool sent = g_nccl.AllReduce(dScratch, dScratch, 1, ncclDouble, ncclMin, comm->comm, stream) == ncclSuccess;
    cudaMemcpyAsync(&v, dScratch, sizeof(double), cudaMemcpyDeviceToHost, stream);
    const bool synced = cudaStreamSynchronize(stream) == cudaSuccess;
    cudaFree(dScratch);
    if (!sent || !synced) { fs.release(false); return fail(DGB_ENCCL, "fused_setup: verdict all-reduce failed"); }
    ok = v > 0.5;
  }
  if (!ok) {
    // all ranks are here with the same verdict; peers may still hold mappings of this rank's block: collective teardown
    if (int rc = fused_teardown(fs, comm, stream)) return rc;
    return DGB_OK;
  }
  fs.epoch = 0;
  fs.on = true;
  return DGB_OK;
}

// base address of the allocation that holds `p` (cuMemGetAddressRange through the driver library, which - like NCCL - is
// loaded at run time so that libdgb.so still loads on a box without a GPU driver)
static bool allocation_base(const void* p, char** base) {
  typedef int (*fn_t)(unsigned long long*, size_t*, unsigned long long);
  static fn_t fn = [] { void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL); return h ? reinterpret_cast<fn_t>(dlsym(h, "cuMemGetAddressRange_v2")) : nullptr; }();
  if (!fn) return false;
  unsigned long long b = 0; size_t n = 0;
  if (fn(&b, &n, reinterpret_cast<unsigned long long>(p)) != 0) return false;
  *base = reinterpret_cast<char*>(b);
  return true;
}

static constexpr int MAX_REG_FIELDS = 4;
struct RegInfo { cudaIpcMemHandle_t handle[MAX_REG_FIELDS]; unsigned long long offset[MAX_REG_FIELDS]; int ok; int pad; };

int fused_register(CommPlan& plan, int myrank, int nranks, dgb_comm* comm, cudaStream_t stream, FusedState& fs,
                   double* const* fields, int nfields) {
  fs.release_direct();
  if (!fs.on) return DGB_OK;
  if (nfields < 1 || nfields > MAX_REG_FIELDS) return fail(DGB_EARG, "dgb_fv_register_fields: 1 to 4 fields");
  const bool multi = nranks > 1;
  // local vote: the plan must be a plain codim-0, one-double-per-cell exchange without overlapping receive boxes
  bool ok = plan.items == 1 && plan.recvWaveEnd.size() <= 1;
  for (auto& b : plan.hostSend) if (b.block != 1 || b.map_offset != 0 || b.index_base != 0) ok = false;
  for (auto& b : plan.hostRecv) if (b.block != 1 || b.map_offset != 0 || b.index_base != 0) ok = false;
  RegInfo mine; std::memset(&mine, 0, sizeof(mine));
  if (multi) {
    for (int f = 0; f < nfields && ok; ++f) {
      char* base = nullptr;
      if (!allocation_base(fields[f], &base)) { ok = false; break; }
      if (cudaIpcGetMemHandle(&mine.handle[f], base) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      mine.offset[f] = (unsigned long long)(reinterpret_cast<const char*>(fields[f]) - base);
    }
  }
  mine.ok = ok ? 1 : 0;
  std::vector<RegInfo> all(nranks);
  if (multi) {
    RegInfo* dAll = nullptr;
    if (cudaMalloc(&dAll, sizeof(RegInfo)*nranks) != cudaSuccess) { cudaGetLastError(); return fail(DGB_ECUDA, "fused_register: cudaMalloc"); }
    cudaMemcpyAsync(dAll + myrank, &mine, sizeof(RegInfo), cudaMemcpyHostToDevice, stream);
    const bool sent = g_nccl.AllGather(dAll + myrank, dAll, sizeof(RegInfo), ncclChar, comm->comm, stream) == ncclSuccess;
    cudaMemcpyAsync(all.data(), dAll, sizeof(RegInfo)*nranks, cudaMemcpyDeviceToHost, stream);
    const bool synced = cudaStreamSynchronize(stream) == cudaSuccess;
    cudaFree(dAll);
    if (!sent || !synced) return fail(DGB_ENCCL, "fused_register: exchange of the field handles failed");
  } else all[0] = mine;
  for (int r = 0; r < nranks; ++r) ok = ok && all[r].ok;
  // map the arrays of the ranks this rank sends to (one mapping per distinct allocation)
  std::vector<std::vector<char*>> fieldOf(nranks, std::vector<char*>(nfields, nullptr));
  std::vector<std::pair<cudaIpcMemHandle_t, char*>> mapped;
  for (int f = 0; f < nfields; ++f) fieldOf[myrank][f] = reinterpret_cast<char*>(fields[f]);
  if (ok)
    for (auto& sg : plan.sendSeg) {
      if (sg.peer == myrank) continue;
      for (int f = 0; f < nfields && ok; ++f) {
        char* base = nullptr;
        for (auto& m : mapped) if (!std::memcmp(&m.first, &all[sg.peer].handle[f], sizeof(cudaIpcMemHandle_t))) base = m.second;
        if (!base) {
          void* p = nullptr;
          if (cudaIpcOpenMemHandle(&p, all[sg.peer].handle[f], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
          fs.opened.push_back(p);
          base = static_cast<char*>(p);
          mapped.emplace_back(all[sg.peer].handle[f], base);
        }
        fieldOf[sg.peer][f] = base + all[sg.peer].offset[f];
      }
    }
  std::vector<FusedState::Direct> dir(ok ? nfields : 0);
  for (int f = 0; f < nfields && ok; ++f) {
    const int nb = plan.nSendBoxes;
    std::vector<double*> dst(nb); std::vector<int> row(nb), plane(nb);
    for (int b = 0; b < nb; ++b) {
      const PlanBox& pb = plan.hostSend[b];
      const long long idx = ((long long)(pb.peer_origin[2]-pb.peer_of_origin[2])*pb.peer_of_size[1] + (pb.peer_origin[1]-pb.peer_of_origin[1]))*pb.peer_of_size[0]
                            + (pb.peer_origin[0]-pb.peer_of_origin[0]);
      dst[b] = reinterpret_cast<double*>(fieldOf[pb.peer][f]) + idx;
      row[b] = pb.peer_of_size[0]; plane[b] = pb.peer_of_size[0]*pb.peer_of_size[1];
    }
    FusedState::Direct& d = dir[f];
    d.field = fields[f];
    d.hDst = dst; d.hRow = row; d.hPlane = plane;
    auto up = [&](void** dev, const void* host, size_t bytes) {
      if (bytes == 0) return true;
      if (cudaMalloc(dev, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
      return cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    ok = ok && up(reinterpret_cast<void**>(&d.dDst), dst.data(), nb*sizeof(double*)) && up(reinterpret_cast<void**>(&d.dRow), row.data(), nb*sizeof(int))
            && up(reinterpret_cast<void**>(&d.dPlane), plane.data(), nb*sizeof(int));
    FusedPack pk[2] = { fs.hostPack[0], fs.hostPack[1] };
    for (int k = 0; k < 2; ++k) { pk[k].dst = d.dDst; pk[k].dstRow = d.dRow; pk[k].dstPlane = d.dPlane; }
    ok = ok && up(reinterpret_cast<void**>(&d.dPack), pk, 2*sizeof(FusedPack));
    // hybrid: x-face boxes keep their place in the receiver's exchange buffer (box-linear), everything else goes to the field
    std::vector<int> rowH(row), planeH(plane);
    for (int b = 0; b < nb; ++b) if (plan.hostSend[b].face == 0 || plan.hostSend[b].face == 1) { rowH[b] = plan.hostSend[b].size[0]; planeH[b] = plan.hostSend[b].size[0]*plan.hostSend[b].size[1]; }
    ok = ok && up(reinterpret_cast<void**>(&d.dRowH), rowH.data(), nb*sizeof(int)) && up(reinterpret_cast<void**>(&d.dPlaneH), planeH.data(), nb*sizeof(int));
    d.hRowH = rowH; d.hPlaneH = planeH;
    FusedPack ph[2] = { fs.hostPack[0], fs.hostPack[1] };
    for (int k = 0; k < 2; ++k) {
      std::vector<double*> dstH(dst);
      for (int b = 0; b < nb; ++b) if (plan.hostSend[b].face == 0 || plan.hostSend[b].face == 1) dstH[b] = fs.hostDst[k][b];
      ok = ok && up(reinterpret_cast<void**>(&d.dDstH[k]), dstH.data(), nb*sizeof(double*));
      d.hDstH[k] = dstH;
      ph[k].dst = d.dDstH[k]; ph[k].dstRow = d.dRowH; ph[k].dstPlane = d.dPlaneH;
    }
    ok = ok && up(reinterpret_cast<void**>(&d.dHybrid), ph, 2*sizeof(FusedPack));
  }
  // hybrid form: this rank's x-face receive boxes must be whole faces, one cell thick (overlap 1), at most one per side
  bool hyb = ok;
  std::vector<PlanBox> xrecv;
  FusedState::XFace xf[2];
  for (auto& b : plan.hostRecv) if (b.face == 0 || b.face == 1) {
    if (b.size[0] != 1 || xf[b.face].on) hyb = false;
    xf[b.face].on = true; xf[b.face].buf_offset = b.buf_offset; xf[b.face].y0 = b.origin[1]; xf[b.face].z0 = b.origin[2]; xf[b.face].ny = b.size[1];
    xrecv.push_back(b);
  }
  for (auto& b : plan.hostSend) if ((b.face == 0 || b.face == 1) && b.size[0] != 1) hyb = false;
  bool hyb_all = hyb && ok;
  if (multi) {                                           // same verdicts everywhere: (registered, hybrid) = min over ranks
    double v[2] = { ok ? 1.0 : 0.0, hyb_all ? 1.0 : 0.0 };
    double* dScratch = nullptr;
    if (cudaMalloc(&dScratch, 2*sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(DGB_ECUDA, "fused_register: cudaMalloc"); }
    cudaMemcpyAsync(dScratch, v, 2*sizeof(double), cudaMemcpyHostToDevice, stream);
    const bool sent = g_nccl.AllReduce(dScratch, dScratch, 2, ncclDouble, ncclMin, comm->comm, stream) == ncclSuccess;
    cudaMemcpyAsync(v, dScratch, 2*sizeof(double), cudaMemcpyDeviceToHost, stream);
    const bool synced = cudaStreamSynchronize(stream) == cudaSuccess;
    cudaFree(dScratch);
    if (!sent || !synced) return fail(DGB_ENCCL, "fused_register: verdict all-reduce failed");
    ok = v[0] > 0.5; hyb_all = v[1] > 0.5;
  }
  if (ok) {
    fs.direct = dir;
    fs.hybridOk = hyb_all;
    if (hyb_all) {
      fs.xface[0] = xf[0]; fs.xface[1] = xf[1];
      fs.hostXRecv = xrecv; fs.nXRecv = int(xrecv.size());
      if (fs.nXRecv) {
        if (cudaMalloc(&fs.dXRecv, fs.nXRecv*sizeof(PlanBox)) != cudaSuccess || cudaMemcpy(fs.dXRecv, xrecv.data(), fs.nXRecv*sizeof(PlanBox), cudaMemcpyHostToDevice) != cudaSuccess)
          return fail(DGB_ECUDA, "fused_register: x-face receive boxes");
      }
    }
  } else {
    fs.direct = dir; fs.release_direct();               // frees the partly built tables
  }
  return DGB_OK;
}

// ---- variable-size exchange (host) ---------------------------------------------------------------------------------------------------
} // namespace dgb

struct dgb_commvar {
  dgb_grid* g = nullptr;
  std::shared_ptr<dgb::CommPlan> plan;
  int myrank = 0, stage = 0;                 // 0 created (send side packed), 1 sizes received, 2 finished
  const long long* csr = nullptr; const double* data = nullptr;
  dgb::VarBuf sendIdx, sendSizes, sendOff, sendItems, recvSizes, recvOff, recvItems, recvIdx, outPos, outIndex, outSizes, outOffsets, outData, tmp;
  std::vector<long long> sendSegItem, recvSegItem;      // per segment: first item, items (pairs)
  long long sendItemsTotal = 0, recvItemsTotal = 0;
};

namespace dgb {

static unsigned var_ctas(long long n, int per = 1) { return unsigned(std::max(1ll, std::min((n*per + PACK_THREADS - 1)/PACK_THREADS, 4096ll))); }

// exclusive prefix sum of n sizes into off[0..n] (off[n] = total)
static int var_scan(dgb_commvar& h, const long long* sizes, long long* off, long long n, cudaStream_t stream) {
  DGB_CUDA(cudaMemsetAsync(off, 0, sizeof(long long), stream));
  if (n <= 0) return DGB_OK;
  if (n >= (1ll << 31)) return fail(DGB_EARG, "variable-size communicate: more than 2^31 entities in one exchange");
  size_t bytes = 0;
  DGB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, sizes, off + 1, int(n), stream));
  if (int rc = h.tmp.alloc(bytes)) return rc;
  DGB_CUDA(cub::DeviceScan::InclusiveSum(h.tmp.p, bytes, sizes, off + 1, int(n), stream));
  return DGB_OK;
}

// item ranges of the plan's segments from a device offsets array (synchronises the stream)
static int var_segment_items(const std::vector<CommPlan::Segment>& segs, const long long* dOff, std::vector<long long>& out, long long total_slots,
                             long long& total, cudaStream_t stream) {
  std::vector<long long> edge(2*segs.size() + 1, 0);
  for (std::size_t k = 0; k < segs.size(); ++k) {
    DGB_CUDA(cudaMemcpyAsync(&edge[2*k], dOff + segs[k].offset, sizeof(long long), cudaMemcpyDeviceToHost, stream));
    DGB_CUDA(cudaMemcpyAsync(&edge[2*k+1], dOff + segs[k].offset + segs[k].count, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  }
  DGB_CUDA(cudaMemcpyAsync(&edge[2*segs.size()], dOff + total_slots, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  DGB_CUDA(cudaStreamSynchronize(stream));
  out.resize(2*segs.size());
  for (std::size_t k = 0; k < segs.size(); ++k) { out[2*k] = edge[2*k]; out[2*k+1] = edge[2*k+1] - edge[2*k]; }
  total = edge[2*segs.size()];
  return DGB_OK;
}

static int var_create(dgb_grid* g, int level, const dgb_mapper& m, int iftype, int dir, const long long* csr, const double* data,
                      cudaStream_t stream, dgb_commvar** out) {
  if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
  for (int c = 0; c < 4; ++c) if (m.block[c] > 1) throw ArgError("variable-size communicate: the mapper layout must have block size 1 on the communicated codims");
  if (!csr) throw ArgError("variable-size communicate: null offsets");
  if (int rc = require_device()) return rc;
  std::unique_ptr<dgb_commvar> h(new dgb_commvar);
  h->g = g; h->myrank = g->me.rank; h->csr = csr; h->data = data;
  const int32_t one[1] = {1};
  if (int rc = get_plan(g, level, m, iftype, dir, one, 1, h->plan)) return rc;
  CommPlan& P = *h->plan;
  const long long ns = P.sendTotal, nr = P.recvTotal;
  if (int rc = h->sendIdx.alloc(ns*8)) return rc;
  if (int rc = h->sendSizes.alloc(ns*8)) return rc;
  if (int rc = h->sendOff.alloc((ns+1)*8)) return rc;
  if (int rc = h->recvSizes.alloc(nr*8)) return rc;
  for (const PlanBox& b : P.hostSend)
    if (b.count > 0) k_var_slots<<<var_ctas(b.count), PACK_THREADS, 0, stream>>>(b, csr, h->sendIdx.as<long long>(), h->sendSizes.as<long long>(), nullptr, 0);
  DGB_CUDA(cudaGetLastError());
  if (int rc = var_scan(*h, h->sendSizes.as<long long>(), h->sendOff.as<long long>(), ns, stream)) return rc;
  if (int rc = var_segment_items(P.sendSeg, h->sendOff.as<long long>(), h->sendSegItem, ns, h->sendItemsTotal, stream)) return rc;
  if (int rc = h->sendItems.alloc(h->sendItemsTotal*8)) return rc;
  if (ns > 0 && h->sendItemsTotal > 0) {
    if (!data) throw ArgError("variable-size communicate: null data");
    k_var_copy<<<var_ctas(ns, 32), PACK_THREADS, 0, stream>>>(ns, h->sendSizes.as<long long>(), data, csr, h->sendIdx.as<long long>(),
                                                             h->sendItems.as<double>(), h->sendOff.as<long long>(), nullptr);
  }
  DGB_CUDA(cudaGetLastError());
  *out = h.release();
  return DGB_OK;
}

static int var_sizes_received(dgb_commvar& h, cudaStream_t stream) {
  if (h.stage != 0) return fail(DGB_EARG, "variable-size communicate: sizes were already received");
  CommPlan& P = *h.plan;
  const long long nr = P.recvTotal;
  if (int rc = h.recvOff.alloc((nr+1)*8)) return rc;
  if (int rc = var_scan(h, h.recvSizes.as<long long>(), h.recvOff.as<long long>(), nr, stream)) return rc;
  if (int rc = var_segment_items(P.recvSeg, h.recvOff.as<long long>(), h.recvSegItem, nr, h.recvItemsTotal, stream)) return rc;
  if (h.recvItemsTotal < 0) return fail(DGB_EARG, "variable-size communicate: negative size received");
  if (int rc = h.recvItems.alloc(h.recvItemsTotal*8)) return rc;
  h.stage = 1;
  return DGB_OK;
}

// 8-byte words of the send buffer to the partners / from the partners into the receive buffer; `item` = false: the size round
static int var_transport(dgb_commvar& h, bool item, dgb_comm* comm, cudaStream_t stream) {
  CommPlan& P = *h.plan;
  bool remote = false;
  for (auto& s : P.sendSeg) if (s.peer != h.myrank) remote = true;
  for (auto& s : P.recvSeg) if (s.peer != h.myrank) remote = true;
  if (remote && !comm) return fail(DGB_EARG, "this exchange has remote partners: a dgb_comm is required");
  if (remote) if (int rc = nccl_check_async(comm)) return rc;
  const char* sb = item ? h.sendItems.as<char>() : h.sendSizes.as<char>();
  char* rb = item ? h.recvItems.as<char>() : h.recvSizes.as<char>();
  auto soff = [&](std::size_t k) { return item ? h.sendSegItem[2*k] : P.sendSeg[k].offset; };
  auto scnt = [&](std::size_t k) { return item ? h.sendSegItem[2*k+1] : P.sendSeg[k].count; };
  auto roff = [&](std::size_t k) { return item ? h.recvSegItem[2*k] : P.recvSeg[k].offset; };
  auto rcnt = [&](std::size_t k) { return item ? h.recvSegItem[2*k+1] : P.recvSeg[k].count; };
  int64_t bytes = 0;
  if (remote) {
    DGB_NCCL(g_nccl.GroupStart());
    for (std::size_t k = 0; k < P.sendSeg.size(); ++k) if (P.sendSeg[k].peer != h.myrank) {
      DGB_NCCL(g_nccl.Send(sb + 8*soff(k), size_t(scnt(k)), ncclInt64, P.sendSeg[k].peer, comm->comm, stream)); bytes += 8*scnt(k); }
    for (std::size_t k = 0; k < P.recvSeg.size(); ++k) if (P.recvSeg[k].peer != h.myrank)
      DGB_NCCL(g_nccl.Recv(rb + 8*roff(k), size_t(rcnt(k)), ncclInt64, P.recvSeg[k].peer, comm->comm, stream));
    DGB_NCCL(g_nccl.GroupEnd());
  }
  for (std::size_t k = 0; k < P.sendSeg.size(); ++k) if (P.sendSeg[k].peer == h.myrank)
    for (std::size_t j = 0; j < P.recvSeg.size(); ++j) if (P.recvSeg[j].peer == h.myrank) {
      if (rcnt(j) != scnt(k)) return fail(DGB_EGRID, "local sends/receives do not match in exchange");
      DGB_CUDA(cudaMemcpyAsync(rb + 8*roff(j), sb + 8*soff(k), size_t(8*scnt(k)), cudaMemcpyDeviceToDevice, stream));
    }
  if (item) h.g->lastBytesSent = bytes;
  return DGB_OK;
}

static int var_finish(dgb_commvar& h, cudaStream_t stream) {
  if (h.stage != 1) return fail(DGB_EARG, "variable-size communicate: finish before the sizes were received (or twice)");
  CommPlan& P = *h.plan;
  const long long nr = P.recvTotal;
  if (int rc = h.recvIdx.alloc(nr*8)) return rc;
  if (int rc = h.outPos.alloc(nr*8)) return rc;
  if (int rc = h.outIndex.alloc(nr*8)) return rc;
  if (int rc = h.outSizes.alloc(nr*8)) return rc;
  if (int rc = h.outOffsets.alloc((nr+1)*8)) return rc;
  if (int rc = h.outData.alloc(h.recvItemsTotal*8)) return rc;
  // the reference scatters box after box in list order, codim dim..0 (yaspgrid.hh:1694-1729): position of a box's first entity there
  std::vector<long long> base(P.hostRecv.size(), 0);
  {
    std::vector<std::size_t> order(P.hostRecv.size());
    for (std::size_t k = 0; k < order.size(); ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return P.hostRecv[a].list_pos < P.hostRecv[b].list_pos; });
    long long at = 0;
    for (std::size_t k : order) { base[k] = at; at += P.hostRecv[k].count; }
  }
  for (std::size_t k = 0; k < P.hostRecv.size(); ++k) {
    const PlanBox& b = P.hostRecv[k];
    if (b.count > 0) k_var_slots<<<var_ctas(b.count), PACK_THREADS, 0, stream>>>(b, nullptr, h.recvIdx.as<long long>(), nullptr, h.outPos.as<long long>(), base[k]);
  }
  if (nr > 0) k_var_permute<<<var_ctas(nr), PACK_THREADS, 0, stream>>>(nr, h.outPos.as<long long>(), h.recvIdx.as<long long>(), h.recvSizes.as<long long>(),
                                                                      h.outIndex.as<long long>(), h.outSizes.as<long long>());
  DGB_CUDA(cudaGetLastError());
  if (int rc = var_scan(h, h.outSizes.as<long long>(), h.outOffsets.as<long long>(), nr, stream)) return rc;
  if (nr > 0 && h.recvItemsTotal > 0)
    k_var_copy<<<var_ctas(nr, 32), PACK_THREADS, 0, stream>>>(nr, h.recvSizes.as<long long>(), h.recvItems.as<double>(), h.recvOff.as<long long>(), nullptr,
                                                             h.outData.as<double>(), h.outOffsets.as<long long>(), h.outPos.as<long long>());
  DGB_CUDA(cudaGetLastError());
  h.stage = 2;
  return DGB_OK;
}

int CommPlan::enter(cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(mtx);
  if (!busy) DGB_CUDA(cudaEventCreateWithFlags(&busy, cudaEventDisableTiming));
  else DGB_CUDA(cudaStreamWaitEvent(stream, busy, 0));
  return DGB_OK;
}
int CommPlan::leave(cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(mtx);
  if (busy) DGB_CUDA(cudaEventRecord(busy, stream));
  return DGB_OK;
}

int nccl_check_async(dgb_comm* comm) {
  if (!comm || !g_nccl.CommGetAsyncError) return DGB_OK;
  ncclResult_t st = ncclSuccess;
  if (g_nccl.CommGetAsyncError(comm->comm, &st) != ncclSuccess) return fail(DGB_ENCCL, "ncclCommGetAsyncError failed");
  if (st != ncclSuccess && st != ncclInProgress) return fail(DGB_ENCCL, std::string("NCCL communicator reports an asynchronous error: ") + g_nccl.GetErrorString(st));
  return DGB_OK;
}

int get_plan(dgb_grid* g, int level, const dgb_mapper& m, int iftype, int dir, const int32_t* item_sizes, int narrays,
             std::shared_ptr<CommPlan>& out) {
  std::ostringstream key;
  key << level << '|' << iftype << '|' << dir;
  for (int c = 0; c < 4; ++c) key << '|' << m.block[c] << ':' << m.offset[c];
  for (int a = 0; a < narrays; ++a) key << '|' << item_sizes[a];
  std::lock_guard<std::mutex> lk(g->mtx);
  auto it = g->plans.find(key.str());
  if (it == g->plans.end()) it = g->plans.emplace(key.str(), build_plan(g, level, m, iftype, dir, item_sizes, narrays)).first;
  out = it->second;
  return DGB_OK;
}

int comm_rank_count(const dgb_comm* comm, int* rank, int* nranks) { *rank = comm ? comm->rank : 0; *nranks = comm ? comm->nranks : 1; return DGB_OK; }
int nccl_allgather_bytes(const void* d_in, void* d_out, size_t bytes, dgb_comm* comm, cudaStream_t stream) {
  if (!comm || comm->nranks == 1) { DGB_CUDA(cudaMemcpyAsync(d_out, d_in, bytes, cudaMemcpyDeviceToDevice, stream)); return DGB_OK; }
  DGB_NCCL(g_nccl.AllGather(d_in, d_out, bytes, ncclChar, comm->comm, stream));
  return DGB_OK;
}

int nccl_allreduce(double* d, int op, dgb_comm* comm, cudaStream_t stream) {
  if (!comm || comm->nranks == 1) return DGB_OK;
  DGB_NCCL(g_nccl.AllReduce(d, d, 1, ncclDouble, op == 0 ? ncclMin : ncclMax, comm->comm, stream));
  return DGB_OK;
}

} // namespace dgb

using namespace dgb;

extern "C" {

int dgb_comm_unique_id(void* id128) {
  if (!g_nccl.load()) return fail(DGB_ENCCL, g_nccl.error);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  DGB_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
  return DGB_OK;
}

int dgb_comm_init(const void* id128, int rank, int nranks, dgb_comm** out) {
  if (int rc = require_device()) return rc;
  if (!g_nccl.load()) return fail(DGB_ENCCL, g_nccl.error);
  ncclUniqueId id; memcpy(&id, id128, sizeof(id));
  ncclComm_t c;
  DGB_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
  *out = new dgb_comm{c, rank, nranks};
  return DGB_OK;
}

int dgb_comm_destroy(dgb_comm* c) {
  if (!c) return DGB_OK;
  g_nccl.CommDestroy(c->comm);
  delete c;
  return DGB_OK;
}

int dgb_communicate(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, int op, double* const* d_arrays,
                    const int32_t* item_sizes, int narrays, dgb_comm* comm, void* stream) {
  return guarded([&]() -> int {
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    if (narrays < 1) throw ArgError("no arrays");
    if (op < DGB_OP_SET || op > DGB_OP_SET_NONNEG) throw ArgError("unknown combine op");
    std::shared_ptr<CommPlan> plan;
    if (int rc = get_plan(g, level, *m, iftype, dir, item_sizes, narrays, plan)) return rc;
    return run_plan(g, *plan, g->me.rank, op, d_arrays, item_sizes, narrays, comm, (cudaStream_t)stream);
  });
}

int dgb_comm_plan_info(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, const int32_t* item_sizes, int narrays,
                       int64_t* out, int cap) {
  return guarded([&]() -> int {
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    std::shared_ptr<CommPlan> plan;
    if (int rc = get_plan(g, level, *m, iftype, dir, item_sizes, narrays, plan)) return rc;
    const int need = 2 + 3*int(plan->sendSeg.size() + plan->recvSeg.size());
    if (cap < need) throw ArgError("plan_info: output too small");
    out[0] = int64_t(plan->sendSeg.size()); out[1] = int64_t(plan->recvSeg.size());
    int64_t* o = out + 2;
    for (auto& s : plan->sendSeg) { *o++ = s.peer; *o++ = s.offset; *o++ = s.count; }
    for (auto& s : plan->recvSeg) { *o++ = s.peer; *o++ = s.offset; *o++ = s.count; }
    return DGB_OK;
  });
}

int dgb_comm_pack(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, double* const* d_arrays,
                  const int32_t* item_sizes, int narrays, double** d_sendbuf, double** d_recvbuf, void* stream) {
  return guarded([&]() -> int {
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    std::shared_ptr<CommPlan> plan;
    if (int rc = get_plan(g, level, *m, iftype, dir, item_sizes, narrays, plan)) return rc;
    if (int rc = plan->enter((cudaStream_t)stream)) return rc;
    if (int rc = plan_pack(*plan, d_arrays, item_sizes, narrays, (cudaStream_t)stream)) return rc;
    if (int rc = plan->leave((cudaStream_t)stream)) return rc;
    if (d_sendbuf) *d_sendbuf = plan->dSendBuf;
    if (d_recvbuf) *d_recvbuf = plan->dRecvBuf;
    return DGB_OK;
  });
}

int dgb_comm_unpack(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, int op, double* const* d_arrays,
                    const int32_t* item_sizes, int narrays, void* stream) {
  return guarded([&]() -> int {
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    if (op < DGB_OP_SET || op > DGB_OP_SET_NONNEG) throw ArgError("unknown combine op");
    std::shared_ptr<CommPlan> plan;
    if (int rc = get_plan(g, level, *m, iftype, dir, item_sizes, narrays, plan)) return rc;
    if (int rc = plan->enter((cudaStream_t)stream)) return rc;
    if (int rc = plan_unpack(*plan, op, d_arrays, item_sizes, narrays, (cudaStream_t)stream)) return rc;
    return plan->leave((cudaStream_t)stream);
  });
}

/* ---- data handles of variable size ---- */
int dgb_commvar_create(dgb_grid* g, int level, const dgb_mapper* m, int iftype, int dir, const int64_t* d_send_offsets,
                       const double* d_send_data, void* stream, dgb_commvar** out) {
  return guarded([&]() -> int {
    if (!g || !m || !out) throw ArgError("null argument");
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t");
    return var_create(g, level, *m, iftype, dir, reinterpret_cast<const long long*>(d_send_offsets), d_send_data, (cudaStream_t)stream, out);
  });
}
int dgb_commvar_exchange(dgb_commvar* h, dgb_comm* comm, void* stream) {
  return guarded([&]() -> int {
    if (!h) throw ArgError("null argument");
    if (int rc = var_transport(*h, false, comm, (cudaStream_t)stream)) return rc;     // the sizes first (yaspgrid.hh:1566-1612)
    if (int rc = var_sizes_received(*h, (cudaStream_t)stream)) return rc;
    return var_transport(*h, true, comm, (cudaStream_t)stream);                        // then the data (:1637-1681)
  });
}
int dgb_commvar_segments(dgb_commvar* h, int recv_side, int64_t* out, int cap) {
  return guarded([&]() -> int {
    if (!h || !out) throw ArgError("null argument");
    const auto& segs = recv_side ? h->plan->recvSeg : h->plan->sendSeg;
    const auto& items = recv_side ? h->recvSegItem : h->sendSegItem;
    if (cap < 1 + 5*int(segs.size())) throw ArgError("segments: output too small");
    out[0] = int64_t(segs.size());
    for (std::size_t k = 0; k < segs.size(); ++k) {
      int64_t* o = out + 1 + 5*k;
      o[0] = segs[k].peer; o[1] = segs[k].offset; o[2] = segs[k].count;
      o[3] = items.size() > 2*k ? items[2*k] : -1; o[4] = items.size() > 2*k ? items[2*k+1] : -1;
    }
    return DGB_OK;
  });
}
int dgb_commvar_buffers(dgb_commvar* h, int64_t** d_send_sizes, double** d_send_items, int64_t** d_recv_sizes, double** d_recv_items) {
  if (!h) return fail(DGB_EARG, "null argument");
  if (d_send_sizes) *d_send_sizes = h->sendSizes.as<int64_t>();
  if (d_send_items) *d_send_items = h->sendItems.as<double>();
  if (d_recv_sizes) *d_recv_sizes = h->recvSizes.as<int64_t>();
  if (d_recv_items) *d_recv_items = h->stage >= 1 ? h->recvItems.as<double>() : nullptr;
  return DGB_OK;
}
int dgb_commvar_sizes_received(dgb_commvar* h, void* stream) {
  return guarded([&]() -> int { if (!h) throw ArgError("null argument"); return var_sizes_received(*h, (cudaStream_t)stream); });
}
int dgb_commvar_finish(dgb_commvar* h, void* stream, int64_t* n_entities, int64_t* n_items) {
  return guarded([&]() -> int {
    if (!h) throw ArgError("null argument");
    if (int rc = var_finish(*h, (cudaStream_t)stream)) return rc;
    if (n_entities) *n_entities = h->plan->recvTotal;
    if (n_items) *n_items = h->recvItemsTotal;
    return DGB_OK;
  });
}
int dgb_commvar_result(dgb_commvar* h, uint32_t* d_index, int64_t* d_offsets, double* d_data, void* stream_) {
  return guarded([&]() -> int {
    if (!h) throw ArgError("null argument");
    if (h->stage != 2) throw ArgError("variable-size communicate: result before finish");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long nr = h->plan->recvTotal;
    if (d_index && nr > 0) k_var_index32<<<var_ctas(nr), PACK_THREADS, 0, stream>>>(nr, h->outIndex.as<long long>(), d_index);
    DGB_CUDA(cudaGetLastError());
    if (d_offsets) DGB_CUDA(cudaMemcpyAsync(d_offsets, h->outOffsets.p, size_t(nr+1)*8, cudaMemcpyDeviceToDevice, stream));
    if (d_data && h->recvItemsTotal > 0) DGB_CUDA(cudaMemcpyAsync(d_data, h->outData.p, size_t(h->recvItemsTotal)*8, cudaMemcpyDeviceToDevice, stream));
    return DGB_OK;
  });
}
int dgb_commvar_destroy(dgb_commvar* h) { delete h; return DGB_OK; }

int dgb_allreduce_min(double* d_value, dgb_comm* comm, void* stream) { return nccl_allreduce(d_value, 0, comm, (cudaStream_t)stream); }
int dgb_allreduce_max(double* d_value, dgb_comm* comm, void* stream) { return nccl_allreduce(d_value, 1, comm, (cudaStream_t)stream); }

} // extern "C"
