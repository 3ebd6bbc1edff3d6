// dgb_fv.cu — the explicit first-order upwind finite-volume transport step as ONE
// fused device kernel per time step: element loop + intersection loop + mapper index arithmetic + axis-aligned
// geometry + CFL reduction, followed by communicate(InteriorBorder_All_Interface, ForwardCommunication).
//
// CPU structure being replaced (dune-grid-howto style, restated in oracle/ref/ref_driver.cc fvRun): pass 1 walks
// elements(gv, interior) and intersections(gv,e), writes update[i] and the local dt; comm().min(dt); pass 2 does
// c += dt*update; then gv.communicate(...). Bytes per cell there: read c + neighbours, write update, read update,
// read+write c = 40 B. Here: read c (8 B) + write c_new (8 B) = 16 B per cell per step; neighbour values come from
// L1/L2. The CFL bound depends only on geometry and velocity (not on c), so the kernel of step n also reduces
// max_i(sum of outflow factors) for step n+1 (velocities are time independent); one geometry-only bootstrap
// reduction precedes the first step. min_i fl(1/s_i) == fl(1/max_i s_i) because correctly rounded division is monotone.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <cuda.h>
#include "dgb_internal.hh"
#include "dgb_device.cuh"
#include "dgb_functors.cuh"
#include "dgb_comm.hh"

namespace dgb {

struct FvProblem { int id; double p[12]; };

// velocity, inflow value and initial value come from the compiled-in problem registry (dgb_functors.cuh; the CPU statement of the
// same functions is oracle/common/fv_problem.hh). Runtime dispatch by id for the general kernels, by template for the fast ones.
template<int ID> struct FvProblemById;
#define DGB_BY_ID(T_) template<> struct FvProblemById<T_::id> { using type = T_; };
DGB_FOR_EACH_FV_PROBLEM(DGB_BY_ID)
#undef DGB_BY_ID

template<int DIM>
__device__ __forceinline__ void fv_velocity(const FvProblem& q, const double* x, double* u) {
  double xx[3] = { x[0], x[1], DIM > 2 ? x[2] : 0.0 }, uu[3] = { 0.0, 0.0, 0.0 };
#define DGB_VEL(T_) if (q.id == T_::id) T_::velocity(q.p, xx, uu);
  DGB_FOR_EACH_FV_PROBLEM(DGB_VEL)
#undef DGB_VEL
#pragma unroll
  for (int i = 0; i < DIM; ++i) u[i] = uu[i];
}
__device__ __forceinline__ double fv_boundary(const FvProblem& q, const double* x) {
  double b = 0.0;
#define DGB_BND(T_) if (q.id == T_::id) b = T_::boundary(q.p, x);
  DGB_FOR_EACH_FV_PROBLEM(DGB_BND)
#undef DGB_BND
  return b;
}
template<int DIM>
__device__ __forceinline__ double fv_initial(const FvProblem& q, const double* x) {
  double c = 0.0;
#define DGB_INI(T_) if (q.id == T_::id) c = T_::initial(q.p, DIM, x);
  DGB_FOR_EACH_FV_PROBLEM(DGB_INI)
#undef DGB_INI
  return c;
}
static bool fv_problem_column_invariant(int id) {
#define DGB_CI(T_) if (id == T_::id) return T_::column_invariant;
  DGB_FOR_EACH_FV_PROBLEM(DGB_CI)
#undef DGB_CI
  return false;
}

static constexpr int FBX = 128, FBY = 4;

struct FvRegion { int origin[3], size[3]; };         // cells to update (a sub-box of the interior box)

// per-axis tables for DGB_FV_SEPARABLE: reciprocal cell widths, cell centres and vertex coordinates
struct FvTables { const double* rw[3]; const double* xc[3]; const double* xv[3]; int first[3]; };

// the send boxes of the fused exchange travel in the kernel PARAMETER space (constant cache): every CTA intersects them with its
// tile at its start, and a global-memory read there costs each of the ~2000 CTAs of a step a DRAM/L2 round trip before it streams
struct FvSendBoxes {
  int n;
  struct Box { int origin[3], size[3]; int row, plane; double* dst; } box[32];      // row / plane: strides (doubles) of the destination
};

struct FvArgs {
  FvRegion region;
  FvProblem problem;
  FvTables tab;
  const double* c;       // field at t_n, indexed by the element index (overlapfront array)
  double* cnew;          // field at t_{n+1} (NULL: reduction only)
  double* update;        // optional: materialise the update array
  const double* slotIn;  // max_i sum_out of this step (produced by the previous launch)
  double* slotOut;       // max_i sum_out for the next step
  double* slotZero;      // slot to clear for the step after next
  double* dtOut;         // dt actually used
  const double* consts;  // device constants {0, boundary value} (upwind sources of faces without a stored neighbour)
  const FusedPack* pack; // fused exchange (3-D step kernels): pack into the receivers' buffers and signal; NULL = off
  unsigned packEpoch;    // value published in the receivers' flags when this launch has delivered everything
  int packSignal;        // 1: the last CTA of this launch publishes the epoch (the last slab launch of a step); 0: pack and fence only
  // hybrid exchange of registered fields (phase-aware ring kernel only): the x neighbours beyond the low / high end of the region
  // are NOT read from the field's overlap column (stale) but from the exchange buffer the peer's previous step filled:
  // xh[side][(z - xhZ0)*xhNy + (y - xhY0)]; NULL = read the field
  const double* xh[2];
  int xhY0[2], xhZ0[2], xhNy[2];
  // deferred completion of the PREVIOUS step's exchange (registered fields: no wait kernel between two steps). The CTAs that touch
  // the rank's boundary - they read halo cells and store into peer memory - first wait until every rank has delivered epoch
  // waitEpoch; the first CTA also reduces the published CFL maxima into *waitSlot (what k_fv_wait does). NULL: nothing pending.
  FusedCtl* waitCtl; unsigned waitEpoch; int waitParity, waitRanks; double* waitSlot; unsigned* waitTimeout;
  unsigned debugSkipFaces;   // developer knob DGB_FV_SKIP_FACES (timing experiments only, wrong results): bit f = do not pack face boxes f
  FvSendBoxes send;          // valid when pack != NULL
};

// ---- fused exchange, device side (dgb_comm.hh: FusedCtl / FusedPack) -------------------------------------------------
// Called by every CTA of a 3-D step kernel after its last store to c_new: the cells of the CTA's tile x chunk that lie in
// a send box of communicate(InteriorBorder_All, Forward) are copied to that box's place in the RECEIVER's buffer (peer-
// mapped memory: the stores travel over NVLink while the other CTAs are still computing). All threads of the CTA share
// the copy, whatever the shape of the intersection (an x-face is one cell per row: 1 x BY x ZC cells here).
// fv_fused_prepare (kernel start, one thread per send box): the intersection of the box with the CTA's tile x chunk.
// fv_fused_pack (after the CTA's last store): all threads share the copy of every non-empty intersection.
struct FusedItem { double* dst0; long long src0; int ex, ey, n, dsy, dsz; };
static constexpr int FUSED_MAX_BOXES = 32;
__device__ __forceinline__ void fv_fused_prepare(const FvArgs& a, const dgb_box& of, FusedItem* items, int x0, int nx, int y0, int ny,
                                                 int z0, int nz, int tid) {
  if (tid < a.send.n) {
    const FvSendBoxes::Box& pb = a.send.box[tid];
    const int lx = max(x0, pb.origin[0]), hx = min(x0+nx, pb.origin[0]+pb.size[0]);
    const int ly = max(y0, pb.origin[1]), hy = min(y0+ny, pb.origin[1]+pb.size[1]);
    const int lz = max(z0, pb.origin[2]), hz = min(z0+nz, pb.origin[2]+pb.size[2]);
    FusedItem it;
    it.n = 0;
    if (lx < hx && ly < hy && lz < hz) {
      it.ex = hx-lx; it.ey = hy-ly; it.n = it.ex*it.ey*(hz-lz);
      // destination layout: the box's place in the receiver's exchange buffer (box-linear), or - registered fields - the
      // receiver's FIELD itself (row and plane strides of the receiver's stored box)
      it.dsy = pb.row; it.dsz = pb.plane;
      it.dst0 = pb.dst + ((long long)(lz-pb.origin[2])*it.dsz + (long long)(ly-pb.origin[1])*it.dsy) + (lx-pb.origin[0]);
      it.src0 = ((long long)(lz-of.origin[2])*of.size[1] + (ly-of.origin[1]))*of.size[0] + (lx-of.origin[0]);
      items[tid] = it;
    } else items[tid].n = 0;
  }
}
template<int NT>
__device__ __forceinline__ void fv_fused_pack(const FvArgs& a, const dgb_box& of, const FusedItem* items, int tid) {
  const int nBoxes = a.send.n;
  const int sy = of.size[0]; const long long sz = (long long)of.size[0]*of.size[1];
  __syncthreads();                                   // the CTA's stores to c_new are visible to all its threads
  for (int b = 0; b < nBoxes; ++b) {
    const int n = items[b].n;
    if (n == 0) continue;
    const FusedItem it = items[b];
    // U independent loads in flight per thread before the first store: a y-face tile is BX x 1 x ZC cells, i.e. 64 cells
    // per thread, and one load-store round trip per cell would keep the CTA (and the kernel's tail) alive for ~60 us
    constexpr int U = 8;
    for (int t0 = tid; t0 < n; t0 += NT*U) {
      double val[U]; long long d[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + u*NT;
        if (t < n) {
          const int i0 = t % it.ex, q = t / it.ex, i1 = q % it.ey, i2 = q / it.ey;
          val[u] = __ldcg(a.cnew + it.src0 + i2*sz + i1*sy + i0);
          d[u] = (long long)i2*it.dsz + i1*it.dsy + i0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) if (t0 + u*NT < n) it.dst0[d[u]] = val[u];
    }
  }
}
// Tile row of a CTA of the 3-D step kernels. With the fused exchange the LAST tile row is scheduled first: a y-face is
// BX x 1 x ZC cells of one CTA, the heaviest pack, and must not sit in the kernel's tail.
__device__ __forceinline__ int fv_tile_row(const FvArgs& a) {
  if (!a.pack || (a.debugSkipFaces & 0x200u)) return int(blockIdx.y);
  return blockIdx.y == 0 ? int(gridDim.y)-1 : int(blockIdx.y)-1;
}
// Called by ONE thread per CTA after a __syncthreads that follows fv_fused_pack and the CTA's CFL atomic: the last CTA of
// the grid publishes the rank's outflow maximum and then the epoch in every rank's control block.
__device__ __forceinline__ void fv_fused_signal(const FvArgs& a, const FusedItem* items) {
  const FusedPack& P = *a.pack;
  // the CTA's peer stores (and its CFL atomic) must be performed system-wide before it is counted; a CTA that stored nothing into
  // peer memory needs no system-scope fence (it would only wait for its own streaming stores to drain: measured ~14 us per step)
  bool stored = false;
  for (int b = 0; b < a.send.n; ++b) stored = stored || items[b].n > 0;
  if (stored) __threadfence_system(); else if (!(a.debugSkipFaces & 0x100u)) __threadfence();
  if (!a.packSignal) return;                         // an earlier slab launch of the step: the last launch signals for all
  const unsigned total = gridDim.x*gridDim.y*gridDim.z;
  // (tried in round 2: counting with a reduction that returns nothing + one publishing CTA that waits for the count, so that no CTA
  // idles for the round trip of the returning atomic - no measurable difference on the periodic proxy, 0.4216 vs 0.4248 ms within
  // the box-to-box spread; the simpler form stays)
  if (atomicAdd(P.counter, 1u) == total-1) {
    atomicExch(P.counter, 0u);
    __threadfence();
    const double m = __longlong_as_double((long long)atomicAdd(reinterpret_cast<unsigned long long*>(a.slotOut), 0ull));
    for (int r = 0; r < P.nranks; ++r) *reinterpret_cast<volatile double*>(P.peerCfl[r]) = m;
    __threadfence_system();
    for (int r = 0; r < P.nranks; ++r) *reinterpret_cast<volatile unsigned*>(P.peerFlag[r]) = a.packEpoch;
  }
}
// see FvArgs::waitCtl. fv_peer_wait_issue (at the very top of the kernel) only ISSUES the flag load, so that its latency overlaps the
// table loads of the column set-up; fv_peer_wait_finish - called by all threads of the CTA before the CTA's first read of c, with
// the CTA-uniform `boundary` - spins if the peer has not delivered yet and publishes the result to the CTA.
__device__ __forceinline__ unsigned fv_peer_wait_issue(const FvArgs& a, bool boundary, int tid) {
  unsigned f = a.waitEpoch;
  const bool first = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (a.waitCtl && (boundary || first) && tid < a.waitRanks)
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&a.waitCtl->flags[tid]) : "memory");
  return f;
}
__device__ __forceinline__ void fv_peer_wait_finish(const FvArgs& a, bool boundary, int tid, unsigned f) {
  if (!a.waitCtl) return;
  const bool first = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (!boundary && !first) return;
  if (tid < a.waitRanks && int(f - a.waitEpoch) < 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&a.waitCtl->flags[tid]) : "memory");
      if (int(f - a.waitEpoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (a.waitCtl->timeout || t1 - t0 > 20000000000ull) {              // 20 s: a peer is gone; sticky, reported by the next API call
        a.waitCtl->timeout = 1u;
        if (a.waitTimeout) { *reinterpret_cast<volatile unsigned*>(a.waitTimeout) = 1u; __threadfence_system(); }
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (first && tid == 0) {
    double m = 0.0;
    for (int r = 0; r < a.waitRanks; ++r) m = fmax(m, *reinterpret_cast<volatile double*>(&a.waitCtl->cfl[a.waitParity][r]));
    *a.waitSlot = m;
  }
}
// does the CTA's tile x chunk lie within `overlap` cells of a side of the rank's interior box that has an overlap layer (a neighbour)?
__device__ __forceinline__ bool fv_touches_rank_boundary(const dgb_view& v, int x0, int nx, int y0, int ny, int z0, int nz) {
  const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
  const dgb_box& in = v.comp[0].box[DGB_BOX_INTERIOR];
  const int lo[3] = { x0, y0, z0 }, n[3] = { nx, ny, nz };
  bool t = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (of.origin[i] < in.origin[i] && lo[i] < in.origin[i] + v.overlap) t = true;
    if (of.origin[i]+of.size[i] > in.origin[i]+in.size[i] && lo[i]+n[i] > in.origin[i]+in.size[i] - v.overlap) t = true;
  }
  return t;
}

// the receiving side: thread s waits until rank s has delivered epoch `target`, then the reduced CFL value is stored
__global__ void k_fv_wait(FusedCtl* ctl, int nranks, unsigned target, int parity, double* slot, unsigned* hostTimeout) {
  __shared__ int s_timed_out;
  const int s = threadIdx.x;
  if (s == 0) s_timed_out = 0;
  __syncthreads();
  if (s < nranks) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned f;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&ctl->flags[s]) : "memory");
      if (int(f - target) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (ctl->timeout || t1 - t0 > 20000000000ull) {                    // 20 s: a peer is gone; do not hang the device
        ctl->timeout = 1u; s_timed_out = 1;                               // sticky: dgb_fv_step refuses every later step
        if (hostTimeout) { *reinterpret_cast<volatile unsigned*>(hostTimeout) = 1u; __threadfence_system(); }
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (s == 0 && !s_timed_out) {
    double m = 0.0;
    for (int r = 0; r < nranks; ++r) m = fmax(m, *reinterpret_cast<volatile double*>(&ctl->cfl[parity][r]));
    *slot = m;
  }
}

// EXACT: the arithmetic of the per-entity CPU loop, operation by operation (bit-identical results).
// SEPARABLE: face factor (u.n)|F|/|V| = (u.n)/w_dir taken from per-axis reciprocal tables (<= 2 ulp difference).
// PERM = 1 (thin-x slabs of the exchange shell): the fast thread index runs along y, blockIdx.z along x.
template<int DIM, int MODE, bool REDUCE_ONLY, int PERM = 0>
__global__ void __launch_bounds__(FBX*FBY) k_fv_step(const __grid_constant__ dgb_view v, const __grid_constant__ FvArgs a) {
  const int t0 = blockIdx.x*FBX + threadIdx.x, t1 = blockIdx.y*FBY + threadIdx.y, t2 = blockIdx.z;
  const int i0 = PERM ? t2 : t0, i1 = PERM ? t0 : t1, i2 = PERM ? t1 : t2;
  const bool active = i0 < a.region.size[0] && i1 < a.region.size[1] && (DIM < 3 || i2 < a.region.size[2]);
  double sum = 0.0;
  if (active) {
    const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
    int coord[3];
    coord[0] = a.region.origin[0] + i0; coord[1] = a.region.origin[1] + i1; coord[2] = DIM > 2 ? a.region.origin[2] + i2 : 0;
    long long stride[3]; stride[0] = 1; stride[1] = of.size[0]; stride[2] = (long long)of.size[0]*of.size[1];
    long long idx = 0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) idx += (coord[i]-of.origin[i])*stride[i];
    double ll[DIM], ur[DIM], w[DIM], xc[DIM], rw[DIM];
    double vol = 1.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      if (MODE == DGB_FV_SEPARABLE) {
        const int t = coord[i]-a.tab.first[i];
        ll[i] = __ldg(a.tab.xv[i]+t); ur[i] = __ldg(a.tab.xv[i]+t+1);
        xc[i] = __ldg(a.tab.xc[i]+t); rw[i] = __ldg(a.tab.rw[i]+t);
      } else {
        ll[i] = vertex_coord(v, i, coord[i]); ur[i] = vertex_coord(v, i, coord[i]+1);
        w[i] = ur[i]-ll[i]; vol *= w[i];
        xc[i] = 0.5*(ll[i]+ur[i]);
      }
    }
    const double ci = REDUCE_ONLY ? 0.0 : __ldg(a.c + idx);
    double upd = 0.0;
#pragma unroll
    for (int dir = 0; dir < DIM; ++dir) {
      double area = 1.0;
      if (MODE == DGB_FV_EXACT) {
#pragma unroll
        for (int i = 0; i < DIM; ++i) if (i != dir) area *= w[i];
      }
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        double fc[DIM], u[DIM];
#pragma unroll
        for (int i = 0; i < DIM; ++i) fc[i] = (i == dir) ? (side ? ur[i] : ll[i]) : xc[i];
        fv_velocity<DIM>(a.problem, fc, u);
        const double un = side ? u[dir] : -u[dir];
        const double factor = (MODE == DGB_FV_EXACT) ? (un*area)/vol : un*rw[dir];
        if (factor >= 0.0) {
          sum += factor;
          if (!REDUCE_ONLY) upd -= ci*factor;
        } else if (!REDUCE_ONLY) {
          if (face_neighbor(v, coord, dir, side)) {
            const double cj = __ldg(a.c + idx + (side ? stride[dir] : -stride[dir]));
            upd -= cj*factor;
          } else if (face_boundary(v, coord, dir, side)) {
            upd -= fv_boundary(a.problem, fc)*factor;
          }
        }
      }
    }
    if (!REDUCE_ONLY) {
      const double dt = 0.99*(1.0/(*a.slotIn));
      if (a.update) a.update[idx] = upd;
      if (a.cnew) a.cnew[idx] = ci + dt*upd;
      if (a.dtOut && i0 == 0 && i1 == 0 && i2 == 0) *a.dtOut = dt;
    }
  }
  // block max of the outflow sums -> one atomic per block (warp shuffles, then shared memory across the warps)
  for (int off = 16; off > 0; off >>= 1) sum = fmax(sum, __shfl_down_sync(0xffffffffu, sum, off));
  __shared__ double part[FBX*FBY/32];
  const int tid = threadIdx.y*FBX + threadIdx.x;
  if ((tid & 31) == 0) part[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int k = 0; k < FBX*FBY/32; ++k) m = fmax(m, part[k]);
    if (m > 0.0) atomic_max_nonneg(a.slotOut, m);
    if (a.slotZero && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) *a.slotZero = 0.0;
  }
}

// ---- exchange shell: all slabs in one launch, with the pack of communicate() fused in -------------------------
// The shell slabs (<= 6 thin boxes of interior cells that some neighbour receives) are updated by one launch:
// blockIdx.y selects the slab, threads run along the slab's longest contiguous axis (y for the thin-x slabs, whose
// cells are one per row anyway). Each new value is stored to c_new AND to every send box of the exchange plan that
// contains the cell ("in-kernel pack": the separate gather kernel, its strided re-read of c_new and one stream
// dependency disappear). Same per-cell arithmetic as k_fv_step.
struct FvShell { FvRegion slab[6]; int perm[6]; int n; const PlanBox* boxes; int nBoxes; double* sendBuf; };

template<int MODE>
__global__ void __launch_bounds__(256) k_fv_shell(const __grid_constant__ dgb_view v, const __grid_constant__ FvArgs a, const __grid_constant__ FvShell sh) {
  const FvRegion& r = sh.slab[blockIdx.y];
  const int perm = sh.perm[blockIdx.y];
  const long long cells = (long long)r.size[0]*r.size[1]*r.size[2];
  const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
  const long long stride[3] = { 1, of.size[0], (long long)of.size[0]*of.size[1] };
  const double dt = 0.99*(1.0/(*a.slotIn));
  double sum_max = 0.0;
  for (long long t = (long long)blockIdx.x*256 + threadIdx.x; t < cells; t += (long long)gridDim.x*256) {
    int i[3];
    if (perm) { i[1] = int(t % r.size[1]); const long long q = t / r.size[1]; i[2] = int(q % r.size[2]); i[0] = int(q / r.size[2]); }
    else { i[0] = int(t % r.size[0]); const long long q = t / r.size[0]; i[1] = int(q % r.size[1]); i[2] = int(q / r.size[1]); }
    int coord[3]; long long idx = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { coord[k] = r.origin[k] + i[k]; idx += (coord[k]-of.origin[k])*stride[k]; }
    double ll[3], ur[3], w[3], xc[3], rw[3], vol = 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (MODE == DGB_FV_SEPARABLE) {
        const int tt = coord[k]-a.tab.first[k];
        ll[k] = __ldg(a.tab.xv[k]+tt); ur[k] = __ldg(a.tab.xv[k]+tt+1); xc[k] = __ldg(a.tab.xc[k]+tt); rw[k] = __ldg(a.tab.rw[k]+tt);
      } else {
        ll[k] = vertex_coord(v, k, coord[k]); ur[k] = vertex_coord(v, k, coord[k]+1);
        w[k] = ur[k]-ll[k]; vol *= w[k]; xc[k] = 0.5*(ll[k]+ur[k]);
      }
    }
    const double ci = __ldg(a.c + idx);
    double upd = 0.0, sum = 0.0;
#pragma unroll
    for (int dir = 0; dir < 3; ++dir) {
      double area = 1.0;
      if (MODE == DGB_FV_EXACT) {
#pragma unroll
        for (int k = 0; k < 3; ++k) if (k != dir) area *= w[k];
      }
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        double fc[3], u[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) fc[k] = (k == dir) ? (side ? ur[k] : ll[k]) : xc[k];
        fv_velocity<3>(a.problem, fc, u);
        const double un = side ? u[dir] : -u[dir];
        const double factor = (MODE == DGB_FV_EXACT) ? (un*area)/vol : un*rw[dir];
        if (factor >= 0.0) { sum += factor; upd -= ci*factor; }
        else if (face_neighbor(v, coord, dir, side)) upd -= __ldg(a.c + idx + (side ? stride[dir] : -stride[dir]))*factor;
        else if (face_boundary(v, coord, dir, side)) upd -= fv_boundary(a.problem, fc)*factor;
      }
    }
    sum_max = fmax(sum_max, sum);
    const double cn = ci + dt*upd;
    a.cnew[idx] = cn;
    for (int b = 0; b < sh.nBoxes; ++b) {                            // in-kernel pack (codim 0, one double per cell)
      const PlanBox& pb = sh.boxes[b];
      const int j0 = coord[0]-pb.origin[0], j1 = coord[1]-pb.origin[1], j2 = coord[2]-pb.origin[2];
      if (j0 >= 0 && j0 < pb.size[0] && j1 >= 0 && j1 < pb.size[1] && j2 >= 0 && j2 < pb.size[2])
        sh.sendBuf[pb.buf_offset + ((long long)j2*pb.size[1] + j1)*pb.size[0] + j0] = cn;
    }
  }
  if (a.dtOut && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.dtOut = dt;
  for (int off = 16; off > 0; off >>= 1) sum_max = fmax(sum_max, __shfl_down_sync(0xffffffffu, sum_max, off));
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum_max;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int k = 0; k < 8; ++k) m = fmax(m, part[k]);
    if (m > 0.0) atomic_max_nonneg(a.slotOut, m);
  }
}

// ---- 3-D production kernel: z-marching with a register prefetch queue ---------------------------------------
// A CTA owns a (BX x BY) column tile of the region and marches over ZC consecutive z-planes. Everything that
// does not depend on z (x/y coordinates, reciprocal widths, x/y neighbour and boundary flags, velocity terms that
// are z-invariant, the time step) is computed once per thread; the z-neighbours of a cell are carried in registers
// (c_below, c_here, c_above). The value of the plane above is not loaded when it is needed but PF planes ahead,
// into a register queue: every thread keeps PF+1 independent 8-byte loads in flight, which is what covers the
// HBM latency (Little's law: 6.5 TB/s x ~0.8 us = 5 MB in flight = 35 KB per SM = 1024 threads x (PF+1) x 8 B
// for PF = 4). The x/y inflow neighbours are the streaming loads of adjacent threads / rows issued PF iterations
// earlier, so they hit L1 (tile interior) or L2 (tile edge). HBM traffic stays at 8 B read + 8 B write per cell;
// the CFL max is reduced per thread over its column, then by warp shuffles, then once per CTA.
template<int PROBLEM>
__device__ __forceinline__ void fv_velocity3(const FvProblem& q, const double* x, double* u) { FvProblemById<PROBLEM>::type::velocity(q.p, x, u); }

template<int MODE, int PROBLEM, bool REDUCE_ONLY, int PF, int BX, int BY>
__global__ void __launch_bounds__(BX*BY) k_fv_march(const __grid_constant__ dgb_view v, const __grid_constant__ FvArgs a, const int zc) {
  const int i0 = blockIdx.x*BX + threadIdx.x;
  const int by = fv_tile_row(a);
  const int i1 = by*BY + threadIdx.y;
  const int z0 = blockIdx.z*zc;
  const int zn = min(zc, a.region.size[2]-z0);
  double sum_max = 0.0;
  if (!REDUCE_ONLY) {
    const bool bnd = fv_touches_rank_boundary(v, a.region.origin[0]+blockIdx.x*BX, min(BX, a.region.size[0]-(int)blockIdx.x*BX), a.region.origin[1]+by*BY,
                                              min(BY, a.region.size[1]-by*BY), a.region.origin[2]+z0, zn);
    const unsigned f = fv_peer_wait_issue(a, bnd, threadIdx.y*BX + threadIdx.x);
    fv_peer_wait_finish(a, bnd, threadIdx.y*BX + threadIdx.x, f);
  }
  if (i0 < a.region.size[0] && i1 < a.region.size[1]) {
    const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
    const dgb_box& ov = v.comp[0].box[DGB_BOX_OVERLAP];
    const int cx = a.region.origin[0]+i0, cy = a.region.origin[1]+i1, cz0 = a.region.origin[2]+z0;
    const long long sy = of.size[0], sz = (long long)of.size[0]*of.size[1];
    const long long idx0 = (long long)(cz0-of.origin[2])*sz + (long long)(cy-of.origin[1])*sy + (cx-of.origin[0]);
    // z-invariant geometry of the column
    double llx, urx, lly, ury, xcx, xcy, rwx = 0.0, rwy = 0.0, wx = 0.0, wy = 0.0;
    if (MODE == DGB_FV_SEPARABLE) {
      const int tx = cx-a.tab.first[0], ty = cy-a.tab.first[1];
      llx = __ldg(a.tab.xv[0]+tx); urx = __ldg(a.tab.xv[0]+tx+1); xcx = __ldg(a.tab.xc[0]+tx); rwx = __ldg(a.tab.rw[0]+tx);
      lly = __ldg(a.tab.xv[1]+ty); ury = __ldg(a.tab.xv[1]+ty+1); xcy = __ldg(a.tab.xc[1]+ty); rwy = __ldg(a.tab.rw[1]+ty);
    } else {
      llx = vertex_coord(v, 0, cx); urx = vertex_coord(v, 0, cx+1); wx = urx-llx; xcx = 0.5*(llx+urx);
      lly = vertex_coord(v, 1, cy); ury = vertex_coord(v, 1, cy+1); wy = ury-lly; xcy = 0.5*(lly+ury);
    }
    // neighbour / boundary flags of the four x/y faces (yaspgridintersection.hh:62-82)
    const bool nbx0 = cx > ov.origin[0] && cx <= ov.origin[0]+ov.size[0]-1;
    const bool nbx1 = cx+1 > ov.origin[0] && cx+1 <= ov.origin[0]+ov.size[0]-1;
    const bool nby0 = cy > ov.origin[1] && cy <= ov.origin[1]+ov.size[1]-1;
    const bool nby1 = cy+1 > ov.origin[1] && cy+1 <= ov.origin[1]+ov.size[1]-1;
    const bool perx = v.periodic & 1u, pery = v.periodic >> 1 & 1u, perz = v.periodic >> 2 & 1u;
    const bool bdx0 = !perx && cx == 0, bdx1 = !perx && cx+1 == v.N[0];
    const bool bdy0 = !pery && cy == 0, bdy1 = !pery && cy+1 == v.N[1];
    const double dt = REDUCE_ONLY ? 0.0 : 0.99*(1.0/(*a.slotIn));
    const double* p = a.c + idx0;
    // last plane offset (relative to cz0) that holds a stored cell this thread may read: the plane above the
    // chunk if it exists (yaspgridintersection.hh:75-82, neighbor() of the +z face)
    const int kload = min(zn, ov.origin[2]+ov.size[2]-1-cz0);
    double c_below = 0.0, c_here = 0.0, c_above = 0.0;
    double q[PF > 0 ? PF : 1];
    if (!REDUCE_ONLY) {
      const bool nb_first = cz0 > ov.origin[2] && cz0 <= ov.origin[2]+ov.size[2]-1;
      if (nb_first) c_below = __ldg(p - sz);
      c_here = __ldg(p);
#pragma unroll
      for (int j = 0; j < PF; ++j) q[j] = (1+j <= kload) ? __ldg(p + (long long)(1+j)*sz) : 0.0;
    }
    double llz = (MODE == DGB_FV_SEPARABLE) ? __ldg(a.tab.xv[2]+(cz0-a.tab.first[2])) : vertex_coord(v, 2, cz0);
    for (int k = 0; k < zn; ++k) {
      const int cz = cz0+k;
      double urz, xcz, rwz = 0.0, wz = 0.0;
      if (MODE == DGB_FV_SEPARABLE) {
        const int tz = cz-a.tab.first[2];
        urz = __ldg(a.tab.xv[2]+tz+1); xcz = __ldg(a.tab.xc[2]+tz); rwz = __ldg(a.tab.rw[2]+tz);
      } else {
        urz = vertex_coord(v, 2, cz+1); wz = urz-llz; xcz = 0.5*(llz+urz);
      }
      const bool nbz0 = cz > ov.origin[2] && cz <= ov.origin[2]+ov.size[2]-1;
      const bool nbz1 = cz+1 > ov.origin[2] && cz+1 <= ov.origin[2]+ov.size[2]-1;
      const bool bdz0 = !perz && cz == 0, bdz1 = !perz && cz+1 == v.N[2];
      if (!REDUCE_ONLY) {
        if (PF > 0) {
          c_above = q[0];
#pragma unroll
          for (int j = 0; j+1 < PF; ++j) q[j] = q[j+1];
          q[PF > 0 ? PF-1 : 0] = (k+1+PF <= kload) ? __ldg(p + (long long)(1+PF)*sz) : 0.0;
        } else if (nbz1) c_above = __ldg(p + sz);
      }
      double vol = 1.0, ax = 1.0, ay = 1.0, az = 1.0;
      if (MODE == DGB_FV_EXACT) { vol = (wx*wy)*wz; ax = wy*wz; ay = wx*wz; az = wx*wy; }
      double sum = 0.0, upd = 0.0;
      // one face: un = u.n, factor = un*|F|/|V|; outflow uses the cell value, inflow the neighbour or boundary value
#define DGB_FACE(DIRX, SIDE, FCX, FCY, FCZ, AREA, RW, NB, BD, CJ)                                             \
      {                                                                                                        \
        const double fc[3] = { FCX, FCY, FCZ };                                                                \
        double u[3]; fv_velocity3<PROBLEM>(a.problem, fc, u);                                                  \
        const double un = (SIDE) ? u[DIRX] : -u[DIRX];                                                         \
        const double factor = (MODE == DGB_FV_EXACT) ? (un*(AREA))/vol : un*(RW);                              \
        if (factor >= 0.0) { sum += factor; if (!REDUCE_ONLY) upd -= c_here*factor; }                          \
        else if (!REDUCE_ONLY) {                                                                               \
          if (NB) upd -= (CJ)*factor;                                                                          \
          else if (BD) upd -= fv_boundary(a.problem, fc)*factor;                                               \
        }                                                                                                      \
      }
      DGB_FACE(0, 0, llx, xcy, xcz, ax, rwx, nbx0, bdx0, __ldg(p - 1))
      DGB_FACE(0, 1, urx, xcy, xcz, ax, rwx, nbx1, bdx1, __ldg(p + 1))
      DGB_FACE(1, 0, xcx, lly, xcz, ay, rwy, nby0, bdy0, __ldg(p - sy))
      DGB_FACE(1, 1, xcx, ury, xcz, ay, rwy, nby1, bdy1, __ldg(p + sy))
      DGB_FACE(2, 0, xcx, xcy, llz, az, rwz, nbz0, bdz0, c_below)
      DGB_FACE(2, 1, xcx, xcy, urz, az, rwz, nbz1, bdz1, c_above)
#undef DGB_FACE
      sum_max = fmax(sum_max, sum);
      if (!REDUCE_ONLY) {
        if (a.update) a.update[idx0 + (long long)k*sz] = upd;
        if (a.cnew) __stcs(a.cnew + idx0 + (long long)k*sz, c_here + dt*upd);
        c_below = c_here; c_here = c_above;
      }
      p += sz; llz = urz;
    }
    if (!REDUCE_ONLY && a.dtOut && i0 == 0 && i1 == 0 && z0 == 0) *a.dtOut = dt;
  }
  const int tid = threadIdx.y*BX + threadIdx.x;
  __shared__ FusedItem s_march_items[FUSED_MAX_BOXES];
  if (!REDUCE_ONLY && a.pack) {
    const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
    fv_fused_prepare(a, of, s_march_items, a.region.origin[0]+blockIdx.x*BX, min(BX, a.region.size[0]-(int)blockIdx.x*BX),
                     a.region.origin[1]+by*BY, min(BY, a.region.size[1]-by*BY), a.region.origin[2]+z0, zn, tid);
    fv_fused_pack<BX*BY>(a, of, s_march_items, tid);
  }
  for (int off = 16; off > 0; off >>= 1) sum_max = fmax(sum_max, __shfl_down_sync(0xffffffffu, sum_max, off));
  __shared__ double part[BX*BY/32];
  if ((tid & 31) == 0) part[tid >> 5] = sum_max;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int k = 0; k < BX*BY/32; ++k) m = fmax(m, part[k]);
    if (m > 0.0) atomic_max_nonneg(a.slotOut, m);
    if (a.slotZero && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) *a.slotZero = 0.0;
    if (!REDUCE_ONLY && a.pack) fv_fused_signal(a, s_march_items);
  }
}

// ---- 3-D separable-mode kernel, shared-memory plane ring ------------------------------------------------------
// Same arithmetic as k_fv_march_sep, different data path: the CTA's (BX x BY) tile of every z-plane, with a
// one-cell halo in x and y, is staged in a ring of D+2 shared-memory slots by cp.async (LDGSTS, 8-byte
// granularity: the rows of a stored box with odd width, e.g. 513 on a partitioned grid, are only 8-byte aligned,
// which rules out TMA boxes and 16-byte copies). D planes are in flight per CTA independent of the register
// budget; the x/y upwind neighbours and the plane above are read from shared memory, so a cell costs one global
// load (plus the halo share), three shared loads and one store. One __syncthreads per plane publishes the slot.
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// per-thread operands of the ring loop
struct FvRingLane {
  double f0, f1, f2, f3, tbx, tby, un5, dt;   // as FvSepLane
  double c_below, zhigh;
  const double* g;                            // own column in c, plane 0 of the chunk
  double* w;                                  // own column in cnew
  double* own;                                // own cell in slot 0 of the ring
  const double* hg;                           // halo duty of this thread: its halo cell in c, plane 0 of the chunk ...
  int hoff;                                   // ... and its shared-memory element offset relative to `own`
  int ox, oy;                                 // shared-memory element offsets of the upwind x / y neighbour (0: none)
  int poff;                                   // A16: shared-memory offset (relative to `own`) of the cell PAIR this thread stages
  bool active, hasHalo, loader;               // loader (A16): this thread stages a pair; g then points at the pair
};
// slow path extras
struct FvRingSlow { double alt0, alt1, alt2, alt3, un4; unsigned flags; unsigned xlE, xlO, xrE, xrO; };    // x*: shared addresses of the x neighbours (pair staging)
// reciprocal z widths of the whole stored box, passed in the kernel parameter space: indexed uniformly per CTA,
// they are read through the constant cache (LDC) and cost no load/store-unit wavefront
#define DGB_RWZ_MAX 1040
struct FvZTable { double rw[DGB_RWZ_MAX]; };

// MODE 0..7: direction triple DX*4+DY*2+DZ of the fast path; 8: per-lane flags.
// Threads outside the region (partial tiles) stage nothing and store nothing but run the same arithmetic (on the
// clamped column of the last valid thread of their row / the last valid row): no divergence around the barriers.
// A16 (rows 16-byte aligned: even row length, even region offset, aligned field): the own cells are staged as PAIRS by half
// of the threads with 16-byte cp.async - half the LDGSTS instructions and shared-memory store wavefronts of the 8-byte form;
// the row pitch grows to BX+4 so that pair destinations are 16-byte aligned (left halo in column 1, cells from column 2).
template<int D, int BX, int BY, int MODE, bool ZPARAM, bool A16>
__device__ __forceinline__ void fv_ring_loop(FvRingLane L, const FvRingSlow S, const long long sz, const double* rwz,
                                             const FvZTable& zt, const int zbase, const int zn, const int kload) {
  constexpr int R = D+2, PX = A16 ? BX+4 : BX+2, PY = BY+2, SLOT = PX*PY;
  constexpr bool DX = MODE & 4, DY = MODE & 2, DZ = MODE & 1;
  const long long sz8 = sz*8;
  const char* gl = reinterpret_cast<const char*>(L.g) + (D+1)*sz8;      // own cell of the next plane to stage
  const char* hl = reinterpret_cast<const char*>(L.hg) + (D+1)*sz8;     // halo cell of the next plane to stage
  char* w = reinterpret_cast<char*>(L.w);
  double c_below = L.c_below, c_here = 0.0;
  const bool self4 = S.un4 >= 0.0, self5 = L.un5 >= 0.0;
  // stage plane k+1+D into the slot that plane k-1 occupied (its readers passed the barrier of iteration k)
#define DGB_RING_CELL(GUARD, K)                                                                                \
  {                                                                                                            \
    cp_async_wait<D-1>();                                                                                      \
    __syncthreads();                                                                                           \
    if (!(GUARD) || (K)+1+D <= kload) {                                                                        \
      if (A16) { if (L.loader) cp_async16(const_cast<double*>(prev) + L.poff, gl); }                           \
      else if (L.active) cp_async8(const_cast<double*>(prev), gl);                                             \
      if (L.hasHalo) cp_async8(const_cast<double*>(prev) + L.hoff, hl);                                        \
    }                                                                                                          \
    cp_async_commit();                                                                                         \
    gl += sz8; hl += sz8;                                                                                      \
    const double c_above = (!(GUARD) || (K)+1 <= kload) ? *nxt : L.zhigh;                                      \
    const double rw = ZPARAM ? zt.rw[zbase + (K)] : __ldg(rwz + (K));                                          \
    double upd;                                                                                                \
    if (MODE < 8) {                                                                                            \
      const double f5 = L.un5*rw;                                                                              \
      const double nx = cur[L.ox], ny = cur[L.oy];                                                             \
      upd = 0.0 - (DX ? nx : c_here)*L.f0;                                                                     \
      if (DX) upd -= L.tbx;                                                                                    \
      upd -= (DX ? c_here : nx)*L.f1;                                                                          \
      if (!DX) upd -= L.tbx;                                                                                   \
      upd -= (DY ? ny : c_here)*L.f2;                                                                          \
      if (DY) upd -= L.tby;                                                                                    \
      upd -= (DY ? c_here : ny)*L.f3;                                                                          \
      if (!DY) upd -= L.tby;                                                                                   \
      upd -= (DZ ? c_below : c_here)*(-f5);                                                                    \
      upd -= (DZ ? c_here : c_above)*f5;                                                                       \
    } else {                                                                                                   \
      double val;                                                                                              \
      val = S.alt0; if (S.flags & 16u) val = cur[-1];   if (S.flags & 1u) val = c_here; upd = 0.0 - val*L.f0;  \
      val = S.alt1; if (S.flags & 32u) val = cur[1];    if (S.flags & 2u) val = c_here; upd -= val*L.f1;       \
      val = S.alt2; if (S.flags & 64u) val = cur[-PX];  if (S.flags & 4u) val = c_here; upd -= val*L.f2;       \
      val = S.alt3; if (S.flags & 128u) val = cur[PX];  if (S.flags & 8u) val = c_here; upd -= val*L.f3;       \
      upd -= (self4 ? c_here : c_below)*(S.un4*rw);                                                            \
      upd -= (self5 ? c_here : c_above)*(L.un5*rw);                                                            \
    }                                                                                                          \
    if (L.active) __stcs(reinterpret_cast<double*>(w), c_here + L.dt*upd);                                     \
    c_below = c_here; c_here = c_above;                                                                        \
    w += sz8;                                                                                                  \
  }
  // plane 0 of the own column (A16: staged by another thread, so its arrival has to be published first)
  cp_async_wait<D>();
  if (A16) __syncthreads();
  c_here = *L.own;
  // the loops are deliberately not unrolled: the ring position rotates through three pointers (previous,
  // current, next slot), which keeps the body small enough for the register allocator not to interleave cells
  const double* prev = L.own + (R-1)*SLOT;               // slot of plane k-1 (free), k, k+1
  const double* cur = L.own;
  const double* nxt = L.own + SLOT;
  const double* const last = L.own + (R-1)*SLOT;
  const int kmain = max(0, min(zn, kload-D));            // cells whose staging and reads all hit stored planes
  int k = 0;
#pragma unroll 1
  for (; k < kmain; ++k) {
    DGB_RING_CELL(false, k)
    prev = cur; cur = nxt; nxt = (nxt == last) ? L.own : nxt + SLOT;
  }
#pragma unroll 1
  for (; k < zn; ++k) {
    DGB_RING_CELL(true, k)
    prev = cur; cur = nxt; nxt = (nxt == last) ? L.own : nxt + SLOT;
  }
#undef DGB_RING_CELL
}

// ---- paired, phase-aware staging (STG = 2) --------------------------------------------------------------------------
// The rows of a rank-local box of a partitioned grid are only 8-byte aligned: its row length is odd (513 for the 2x2x2 split of
// 1024^3) or the interior starts at an odd offset, and with an odd plane stride the alignment of a row alternates from
// plane to plane as well. 8-byte LDGSTS staging costs twice the instructions and shared-memory store wavefronts of 16-byte
// staging, which is why those boxes ran 11 % slower than the aligned single-GPU box. Here EVERY staged row
// is copied as 16-byte aligned PAIRS whatever its alignment: a cell is placed in shared memory at column
//     2 + (x - x0) + phi(row, plane),     phi = parity of the element address of the tile's first cell of that row,
// so that even element addresses land on even columns. The aligned pairs of a row then sit at fixed columns (2p, 2p+1) and
// cover the tile's cells, both x neighbours and at most three surplus cells; thread t < (BY+2)*(BX+4)/2 stages pair
// t % ((BX+4)/2) of row t / ((BX+4)/2) (tile rows and the y neighbour rows alike), so no separate halo duty exists.
// A thread's own cell moves by the phase of its row: its column alternates from plane to plane when the plane stride is
// odd, and the y neighbour's row has the other phase when the row stride is odd. The ring has an even number of slots and
// the loop is unrolled by two, so "even" and "odd" planes of the chunk always use the same slots and every phase-dependent
// offset is a per-thread constant (one for even, one for odd planes): the hot loop has no phase arithmetic at all. Slot
// positions are CTA-uniform byte offsets (uPrev/uCur/uNxt), per-thread positions are byte offsets inside a slot.
// .ca, not .cg: measured with ncu (profiles/r02_fv_staging_notes.md) - 16-byte copies that bypass L1 are merged into 32-byte
// sector requests per aligned LANE PAIR only, so rows whose pairs start at an odd multiple of 16 bytes fetched every
// sector twice (L2 -> SM traffic 2.16 GB instead of 1.39 GB per 513^3 launch); through L1 the second request hits
__device__ __forceinline__ void cp_async16_b(unsigned smem, const void* gmem) {
#ifdef DGB_PAIR_CG
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(smem), "l"(gmem) : "memory");
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" :: "r"(smem), "l"(gmem) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8_b(unsigned smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned smem) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(smem) : "memory");
  return v;
}
// a pair that may stick out of the array by one element at either end (first / last element of the field)
__device__ __forceinline__ void stage_pair_checked(unsigned dst, const char* src, const char* lo, const char* hi) {
  if (src >= lo && src + 16 <= hi) cp_async16_b(dst, src);
  else {
    if (src >= lo && src + 8 <= hi) cp_async8_b(dst, src);
    if (src + 8 >= lo && src + 16 <= hi) cp_async8_b(dst + 8, src + 8);
  }
}
struct FvPairLane {
  double f0, f1, f2, f3, tbx, tby, un5, dt;   // as FvRingLane
  double c_below, zhigh;
  char* w;                                    // own column in cnew
  const char* glE; const char* glO;           // loader: source of this thread's pair in the next even / odd plane to stage
  unsigned aE, aO;                            // shared address of the own cell in slot 0 for an even / odd plane of the chunk
  unsigned xE, xO, yE, yO;                    // ... of the upwind x / y neighbour (= own cell if none); slow path: yE/yO = phase step
  unsigned ld;                                // loader: shared address of the pair in slot 0 (side duty: of the side cell)
  unsigned step2;                             // loader: bytes from a plane to the next plane of the same parity in the source
  bool active, loader, side;                  // side duty (hybrid exchange): one 8-byte x halo cell per plane from the exchange buffer
};

// MODE 9 (problems whose velocity depends on z: column_invariant == false): nothing is hoisted out of the z loop - every plane
// evaluates the velocity functor at its six face centres and takes the upwind decisions there, in the order and with the
// expressions of k_fv_march (same bits). Per-thread operands re-use the lane: f0..f3 = llx, urx, lly, ury; tbx, tby = xcx, xcy;
// un5 = rwx; S.alt0 = rwy; S.flags: bit m = face m has a stored neighbour, bit 4+m = face m lies on the physical boundary.
struct FvGeneral {
  const FvProblem* q;
  const double* xvz; const double* xcz;     // z vertex / centre tables, entry 0 = first plane of the chunk
  int cz0, ovz0, ovz1, nz; bool perz;
  double sum;                                // max over the planes of the outflow sum (CFL)
};

template<int D, int BX, int BY, int MODE, bool ZPARAM, int PROBLEM = 0>
__device__ __forceinline__ void fv_pair_loop(FvPairLane L, const FvRingSlow S, const long long sz, const double* rwz, const FvZTable& zt,
                                             const int zbase, const int zn, const int kload, const int kmain,
                                             const char* lo, const char* hi, FvGeneral* G = nullptr) {
  constexpr int R = D+2, PX = BX+4, PY = BY+2;
  constexpr unsigned SLOTB = (PX*PY + 2*PY)*8u, LASTB = (R-1)*SLOTB, ROWB = PX*8u;       // + 2*PY side cells (x halo from the exchange buffer)
  static_assert(R % 2 == 0, "even and odd planes must keep their slots");
  constexpr bool DX = MODE & 4, DY = MODE & 2, DZ = MODE & 1;
  const long long sz8 = sz*8;
  double c_below = L.c_below, c_here = 0.0;
  const bool self4 = S.un4 >= 0.0, self5 = L.un5 >= 0.0;
  unsigned uPrev = LASTB, uCur = 0u, uNxt = SLOTB;       // CTA-uniform: slot of plane k-1 (free), k, k+1
  /* one face of the general mode: the expressions and the order of DGB_FACE in k_fv_march (separable) */
#define DGB_GFACE(DIRX, SIDE, FCX, FCY, FCZ, RW, NB, BD, CJ)                                                   \
      {                                                                                                        \
        const double fc[3] = { FCX, FCY, FCZ };                                                                \
        double u[3]; fv_velocity3<PROBLEM>(*G->q, fc, u);                                                      \
        const double un = (SIDE) ? u[DIRX] : -u[DIRX];                                                         \
        const double factor = un*(RW);                                                                         \
        if (factor >= 0.0) { sum += factor; upd -= c_here*factor; }                                            \
        else if (NB) upd -= (CJ)*factor;                                                                       \
        else if (BD) upd -= fv_boundary(*G->q, fc)*factor;                                                     \
      }
#define DGB_PAIR_CELL(GUARD, K, AOWN, AX, AY, ANXT, GL, XL, XR)                                               \
  {                                                                                                            \
    cp_async_wait<D-1>();                                                                                      \
    __syncthreads();                                                                                           \
    if (!(GUARD)) { if (L.loader) cp_async16_b(uPrev + L.ld, GL); else if (L.side) cp_async8_b(uPrev + L.ld, GL); }   \
    else if ((K)+1+D <= kload && L.loader) stage_pair_checked(uPrev + L.ld, GL, lo, hi);                       \
    else if ((K)+1+D < zn && L.side) cp_async8_b(uPrev + L.ld, GL);                                            \
    cp_async_commit();                                                                                         \
    GL += L.step2;                                                                                             \
    const double c_above = (!(GUARD) || (K)+1 <= kload) ? lds_f64(uNxt + (ANXT)) : L.zhigh;                    \
    const double rw = ZPARAM ? zt.rw[zbase + (K)] : __ldg(rwz + (K));                                          \
    double upd;                                                                                                \
    if (MODE < 8) {                                                                                            \
      const double f5 = L.un5*rw;                                                                              \
      const double nx = lds_f64(uCur + (AX)), ny = lds_f64(uCur + (AY));                                       \
      upd = 0.0 - (DX ? nx : c_here)*L.f0;                                                                     \
      if (DX) upd -= L.tbx;                                                                                    \
      upd -= (DX ? c_here : nx)*L.f1;                                                                          \
      if (!DX) upd -= L.tbx;                                                                                   \
      upd -= (DY ? ny : c_here)*L.f2;                                                                          \
      if (DY) upd -= L.tby;                                                                                    \
      upd -= (DY ? c_here : ny)*L.f3;                                                                          \
      if (!DY) upd -= L.tby;                                                                                   \
      upd -= (DZ ? c_below : c_here)*(-f5);                                                                    \
      upd -= (DZ ? c_here : c_above)*f5;                                                                       \
    } else if (MODE == 9) {                                                                                    \
      const unsigned own = uCur + (AOWN);                                                                      \
      const int cz = G->cz0 + (K);                                                                             \
      const double llz = __ldg(G->xvz + (K)), urz = __ldg(G->xvz + (K) + 1), xcz = __ldg(G->xcz + (K));         \
      const bool nbz0 = cz > G->ovz0 && cz <= G->ovz1, nbz1 = cz+1 > G->ovz0 && cz+1 <= G->ovz1;                \
      const bool bdz0 = !G->perz && cz == 0, bdz1 = !G->perz && cz+1 == G->nz;                                  \
      double sum = 0.0;                                                                                        \
      upd = 0.0;                                                                                               \
      DGB_GFACE(0, 0, L.f0, L.tby, xcz, L.un5, S.flags & 1u, S.flags & 16u, lds_f64(uCur + (XL)))             \
      DGB_GFACE(0, 1, L.f1, L.tby, xcz, L.un5, S.flags & 2u, S.flags & 32u, lds_f64(uCur + (XR)))             \
      DGB_GFACE(1, 0, L.tbx, L.f2, xcz, S.alt0, S.flags & 4u, S.flags & 64u, lds_f64(own - ROWB + (AY)))      \
      DGB_GFACE(1, 1, L.tbx, L.f3, xcz, S.alt0, S.flags & 8u, S.flags & 128u, lds_f64(own + ROWB + (AY)))     \
      DGB_GFACE(2, 0, L.tbx, L.tby, llz, rw, nbz0, bdz0, c_below)                                              \
      DGB_GFACE(2, 1, L.tbx, L.tby, urz, rw, nbz1, bdz1, c_above)                                              \
      G->sum = fmax(G->sum, sum);                                                                              \
    } else {                                                                                                   \
      /* slow path: AY is the phase step (in bytes) towards the rows above and below */                       \
      const unsigned own = uCur + (AOWN);                                                                      \
      double val;                                                                                              \
      val = S.alt0; if (S.flags & 16u) val = lds_f64(uCur + (XL));          if (S.flags & 1u) val = c_here; upd = 0.0 - val*L.f0;  \
      val = S.alt1; if (S.flags & 32u) val = lds_f64(uCur + (XR));          if (S.flags & 2u) val = c_here; upd -= val*L.f1;       \
      val = S.alt2; if (S.flags & 64u) val = lds_f64(own - ROWB + (AY));    if (S.flags & 4u) val = c_here; upd -= val*L.f2;       \
      val = S.alt3; if (S.flags & 128u) val = lds_f64(own + ROWB + (AY));   if (S.flags & 8u) val = c_here; upd -= val*L.f3;       \
      upd -= (self4 ? c_here : c_below)*(S.un4*rw);                                                            \
      upd -= (self5 ? c_here : c_above)*(L.un5*rw);                                                            \
    }                                                                                                          \
    if (L.active) __stcs(reinterpret_cast<double*>(L.w), c_here + L.dt*upd);                                   \
    c_below = c_here; c_here = c_above;                                                                        \
    L.w += sz8;                                                                                                \
    uPrev = uCur; uCur = uNxt; uNxt = (uNxt == LASTB) ? 0u : uNxt + SLOTB;                                     \
  }
  // plane 0 of the own column was staged by another thread: publish its arrival first
  cp_async_wait<D>();
  __syncthreads();
  c_here = lds_f64(L.aE);
  int k = 0;
  // an even plane of the chunk stages plane k+1+D, which is odd (D is even), and vice versa
#pragma unroll 1
  for (; k+1 < kmain; k += 2) {
    DGB_PAIR_CELL(false, k, L.aE, L.xE, L.yE, L.aO, L.glO, S.xlE, S.xrE)
    DGB_PAIR_CELL(false, k+1, L.aO, L.xO, L.yO, L.aE, L.glE, S.xlO, S.xrO)
  }
#pragma unroll 1
  for (; k < zn; k += 2) {
    DGB_PAIR_CELL(true, k, L.aE, L.xE, L.yE, L.aO, L.glO, S.xlE, S.xrE)
    if (k+1 < zn) DGB_PAIR_CELL(true, k+1, L.aO, L.xO, L.yO, L.aE, L.glE, S.xlO, S.xrO)
  }
#undef DGB_PAIR_CELL
#undef DGB_GFACE
}

// ---- tensor-map staging (STG = 3) ------------------------------------------------------------------------------------
// Boxes whose row and plane strides are multiples of 16 bytes (even row length, 16-byte aligned field: the single-GPU 512^3
// box) can be described to the TMA engine as a 3-D tensor of doubles. ONE thread of the CTA then stages a whole plane tile
// (+ its x neighbours and the y neighbour row of the inflow side: a box of (BX+4) x (BY+1) cells; (BX+4) x (BY+2) where
// the CTA's columns do not share a flow direction) with ONE cp.async.bulk.tensor instruction per plane. A box may start at
// any EVEN column (its first byte must be 16-byte aligned - an odd start raises an illegal-instruction error, tools/probes/
// tma_probe.cu), any row and plane, also outside the array: cells beyond the array are zero filled (they are never read: the neighbour flags of
// the column decide). The other 511 threads issue no staging instruction at all. Completion is a transaction count on one
// mbarrier per ring slot; the barrier of iteration k that frees the slot of plane k-1 stays a __syncthreads.
struct FvTmaMaps { alignas(64) unsigned char inflow[128]; alignas(64) unsigned char full[128]; };   // two CUtensorMap objects
struct FvNoTma {};
template<int STG> struct FvTmaParam { typedef FvNoTma type; };
template<> struct FvTmaParam<3> { typedef FvTmaMaps type; };

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra WAIT_%=;\n}\n"
               :: "r"(bar), "r"(parity) : "memory");
}
// arm the slot's barrier with the box size and start the copy of the box at (x, y, z) of the tensor
__device__ __forceinline__ void tma_stage_plane(unsigned dst, const void* map, int x, int y, int z, unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  unsigned p;
  asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
  return p != 0u;
}
struct FvTmaLane {
  double f0, f1, f2, f3, tbx, tby, un5, dt;   // as FvRingLane
  double c_below, zhigh;
  char* w;                                    // own column in cnew
  unsigned aOwn, aX, aY;                      // shared address of the own cell / the upwind x / y neighbour in slot 0 (= own cell if none)
  bool active, issuer;                        // issuer: the warp that stages (one elected lane)
};

template<int D, int BX, int BY, int MODE, bool ZPARAM>
__device__ __forceinline__ void fv_tma_loop(FvTmaLane L, const FvRingSlow S, const long long sz, const double* rwz, const FvZTable& zt,
                                            const int zbase, const int zn, const int kload, const void* map, const unsigned bar0,
                                            const unsigned ringBase, const int bx0, const int by0, const int bz0, const unsigned bytes) {
  constexpr int R = D+2;
  static_assert((R & (R-1)) == 0, "ring slots: a power of two");
  constexpr int LOGR = R == 8 ? 3 : R == 4 ? 2 : R == 16 ? 4 : 1;
  constexpr unsigned ROWB = (BX+4)*8u;
  constexpr bool DX = MODE & 4, DY = MODE & 2, DZ = MODE & 1;
  const long long sz8 = sz*8;
  constexpr unsigned SLOTB = unsigned(((BX+4)*(BY+2)*8 + 127)/128*128), LASTB = (R-1)*SLOTB;
  double c_below = L.c_below, c_here = 0.0;
  const bool self4 = S.un4 >= 0.0, self5 = L.un5 >= 0.0;
#define DGB_TMA_CELL(GUARD, K, UPREV, UCUR, UNXT, BARW, PARW, BARI)                                           \
  {                                                                                                            \
    if (!(GUARD) || (K)+1 <= kload) mbar_wait(bar0 + (BARW), (PARW));                                          \
    __syncthreads();                                                                                           \
    if (L.issuer && (!(GUARD) || (K)+1+D <= kload)) {                                                          \
      if (elect_one()) tma_stage_plane(ringBase + (UPREV), map, bx0, by0, bz0 + (K)+1+D, bar0 + (BARI), bytes); \
    }                                                                                                          \
    const double c_above = (!(GUARD) || (K)+1 <= kload) ? lds_f64((UNXT) + L.aOwn) : L.zhigh;                  \
    const double rw = ZPARAM ? zt.rw[zbase + (K)] : __ldg(rwz + (K));                                          \
    double upd;                                                                                                \
    if (MODE < 8) {                                                                                            \
      const double f5 = L.un5*rw;                                                                              \
      const double nx = lds_f64((UCUR) + L.aX), ny = lds_f64((UCUR) + L.aY);                                   \
      upd = 0.0 - (DX ? nx : c_here)*L.f0;                                                                     \
      if (DX) upd -= L.tbx;                                                                                    \
      upd -= (DX ? c_here : nx)*L.f1;                                                                          \
      if (!DX) upd -= L.tbx;                                                                                   \
      upd -= (DY ? ny : c_here)*L.f2;                                                                          \
      if (DY) upd -= L.tby;                                                                                    \
      upd -= (DY ? c_here : ny)*L.f3;                                                                          \
      if (!DY) upd -= L.tby;                                                                                   \
      upd -= (DZ ? c_below : c_here)*(-f5);                                                                    \
      upd -= (DZ ? c_here : c_above)*f5;                                                                       \
    } else {                                                                                                   \
      const unsigned own = (UCUR) + L.aOwn;                                                                    \
      double val;                                                                                              \
      val = S.alt0; if (S.flags & 16u) val = lds_f64(own - 8u);     if (S.flags & 1u) val = c_here; upd = 0.0 - val*L.f0;  \
      val = S.alt1; if (S.flags & 32u) val = lds_f64(own + 8u);     if (S.flags & 2u) val = c_here; upd -= val*L.f1;       \
      val = S.alt2; if (S.flags & 64u) val = lds_f64(own - ROWB);   if (S.flags & 4u) val = c_here; upd -= val*L.f2;       \
      val = S.alt3; if (S.flags & 128u) val = lds_f64(own + ROWB);  if (S.flags & 8u) val = c_here; upd -= val*L.f3;       \
      upd -= (self4 ? c_here : c_below)*(S.un4*rw);                                                            \
      upd -= (self5 ? c_here : c_above)*(L.un5*rw);                                                            \
    }                                                                                                          \
    if (L.active) __stcs(reinterpret_cast<double*>(L.w), c_here + L.dt*upd);                                   \
    c_below = c_here; c_here = c_above;                                                                        \
    L.w += sz8;                                                                                                \
  }
  /* slot s of the ring: compile-time offsets while the plane number is a multiple of R at s = 0 */            
#define DGB_TMA_FIXED(K, S_)                                                                                   \
  DGB_TMA_CELL(false, (K)+(S_), unsigned(((S_)+R-1) % R)*SLOTB, unsigned(S_)*SLOTB, unsigned(((S_)+1) % R)*SLOTB,  \
               unsigned(((S_)+1) % R)*8u, ((S_) == R-1 ? par ^ 1u : par), unsigned(((S_)+R-1) % R)*8u)
  mbar_wait(bar0, 0u);                                   // plane 0 (always stored: the chunk starts on an interior plane)
  c_here = lds_f64(L.aOwn);
  const int kmain = max(0, min(zn, kload-D));            // cells whose staging and reads all hit stored planes
  int k = 0;
  unsigned par = 0u;                                     // parity of the use of the slots by the planes k .. k+R-1
  if constexpr (R == 8) {
#pragma unroll 1
    for (; k+R <= kmain; k += R) {
      DGB_TMA_FIXED(k, 0) DGB_TMA_FIXED(k, 1) DGB_TMA_FIXED(k, 2) DGB_TMA_FIXED(k, 3)
      DGB_TMA_FIXED(k, 4) DGB_TMA_FIXED(k, 5) DGB_TMA_FIXED(k, 6) DGB_TMA_FIXED(k, 7)
      par ^= 1u;
    }
  } else if constexpr (R == 16) {
#pragma unroll 1
    for (; k+R <= kmain; k += R) {
      DGB_TMA_FIXED(k, 0) DGB_TMA_FIXED(k, 1) DGB_TMA_FIXED(k, 2) DGB_TMA_FIXED(k, 3)
      DGB_TMA_FIXED(k, 4) DGB_TMA_FIXED(k, 5) DGB_TMA_FIXED(k, 6) DGB_TMA_FIXED(k, 7)
      DGB_TMA_FIXED(k, 8) DGB_TMA_FIXED(k, 9) DGB_TMA_FIXED(k, 10) DGB_TMA_FIXED(k, 11)
      DGB_TMA_FIXED(k, 12) DGB_TMA_FIXED(k, 13) DGB_TMA_FIXED(k, 14) DGB_TMA_FIXED(k, 15)
      par ^= 1u;
    }
  }
  unsigned uPrev = unsigned((k+R-1) & (R-1))*SLOTB, uCur = unsigned(k & (R-1))*SLOTB, uNxt = unsigned((k+1) & (R-1))*SLOTB;
#define DGB_TMA_ROTATE uPrev = uCur; uCur = uNxt; uNxt = (uNxt == LASTB) ? 0u : uNxt + SLOTB;
#pragma unroll 1
  for (; k < kmain; ++k) {
    DGB_TMA_CELL(false, k, uPrev, uCur, uNxt, unsigned((k+1) & (R-1))*8u, unsigned((k+1) >> LOGR) & 1u, unsigned((k+1+D) & (R-1))*8u)
    DGB_TMA_ROTATE
  }
#pragma unroll 1
  for (; k < zn; ++k) {
    DGB_TMA_CELL(true, k, uPrev, uCur, uNxt, unsigned((k+1) & (R-1))*8u, unsigned((k+1) >> LOGR) & 1u, unsigned((k+1+D) & (R-1))*8u)
    DGB_TMA_ROTATE
  }
#undef DGB_TMA_ROTATE
#undef DGB_TMA_FIXED
#undef DGB_TMA_CELL
}

template<int PROBLEM, int D, int BX, int BY, int MINB, bool ZPARAM, int STG = 0>      // STG: 0 8-byte staging, 1 aligned pairs, 2 phase-aware pairs, 3 tensor map (TMA)
__global__ void __launch_bounds__(BX*BY, MINB) k_fv_ring(const __grid_constant__ dgb_view v, const __grid_constant__ FvArgs a,
                                                         const __grid_constant__ FvZTable zt, const int zc,
                                                         const __grid_constant__ typename FvTmaParam<STG>::type tmaps) {
  constexpr bool A16 = STG == 1;
  constexpr int PX = STG ? BX+4 : BX+2, PY = BY+2, SLOT = PX*PY, XO = STG ? 2 : 1;      // XO: column of the tile's first cell
  static_assert(STG == 3 || 2*BX + 2*BY <= BX*BY, "one halo cell per thread");
  static_assert(STG == 3 || PY*(PX/2) <= BX*BY, "one pair per thread");
  extern __shared__ __align__(16) double ring[];       // [R][PY][PX]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.x*BX + tx;
  const int by = fv_tile_row(a);
  const int i1 = by*BY + ty;
  const int z0 = blockIdx.z*zc;
  const int zn = min(zc, a.region.size[2]-z0);
  const int cz0 = a.region.origin[2]+z0;
  const bool active = i0 < a.region.size[0] && i1 < a.region.size[1];
  const int zbase = cz0-a.tab.first[2];
  const double* rwz = a.tab.rw[2] + zbase;
  // deferred completion of the previous step's exchange: boundary CTAs wait for the peers before their first read of c; the flag
  // load is issued here, its latency overlaps the table loads below
  const bool rankBoundary = fv_touches_rank_boundary(v, a.region.origin[0]+blockIdx.x*BX, min(BX, a.region.size[0]-(int)blockIdx.x*BX), a.region.origin[1]+by*BY,
                                                     min(BY, a.region.size[1]-by*BY), cz0, zn);
  const unsigned peerFlag = fv_peer_wait_issue(a, rankBoundary, ty*BX + tx);
  double rwmax = 0.0;
  for (int k = tx & 31; k < zn; k += 32) rwmax = fmax(rwmax, __ldg(rwz + k));
  for (int o = 16; o > 0; o >>= 1) rwmax = fmax(rwmax, __shfl_xor_sync(0xffffffffu, rwmax, o));
  double sum_max = 0.0;
  double f0 = 0, f1 = 0, f2 = 0, f3 = 0, un4 = 0, un5 = 0, alt0 = 0, alt1 = 0, alt2 = 0, alt3 = 0;
  unsigned flags = 0;                 // bit m: face m flows out; 4+m: upwind value is a stored neighbour; 8+m: boundary value
  unsigned gflags = 0;                // general mode: bit m: face m has a stored neighbour; 4+m: face m lies on the physical boundary
  constexpr bool GENERAL = !FvProblemById<PROBLEM>::type::column_invariant;
  double g_llx = 0, g_urx = 0, g_lly = 0, g_ury = 0, g_xcx = 0, g_xcy = 0, g_rwx = 0, g_rwy = 0;
  const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
  const dgb_box& ov = v.comp[0].box[DGB_BOX_OVERLAP];
  const int sy = of.size[0];
  const long long sz = (long long)sy*of.size[1];
  const bool perz = v.periodic >> 2 & 1u;
  const int ovz0 = ov.origin[2], ovz1 = ovz0 + ov.size[2]-1;
  const double bval = fv_boundary(a.problem, nullptr);
  // threads outside the region shadow the last valid column of their row / the last valid row (stores suppressed)
  const int cx = a.region.origin[0]+min(i0, a.region.size[0]-1), cy = a.region.origin[1]+min(i1, a.region.size[1]-1);
  {
    const int ttx = cx-a.tab.first[0], tty = cy-a.tab.first[1];
    const double llx = __ldg(a.tab.xv[0]+ttx), urx = __ldg(a.tab.xv[0]+ttx+1), xcx = __ldg(a.tab.xc[0]+ttx), rwx = __ldg(a.tab.rw[0]+ttx);
    const double lly = __ldg(a.tab.xv[1]+tty), ury = __ldg(a.tab.xv[1]+tty+1), xcy = __ldg(a.tab.xc[1]+tty), rwy = __ldg(a.tab.rw[1]+tty);
    const double xcz = __ldg(a.tab.xc[2]+(cz0-a.tab.first[2]));
    const bool perx = v.periodic & 1u, pery = v.periodic >> 1 & 1u;
    if (GENERAL) { g_llx = llx; g_urx = urx; g_lly = lly; g_ury = ury; g_xcx = xcx; g_xcy = xcy; g_rwx = rwx; g_rwy = rwy; }
    double sxy = 0.0;
#define DGB_COLFACE(K, F, ALT, DIRX, SIDE, FCX, FCY, RW, NB, BD)                                               \
    {                                                                                                          \
      const double fc[3] = { FCX, FCY, xcz };                                                                  \
      double u[3]; fv_velocity3<PROBLEM>(a.problem, fc, u);                                                    \
      const double un = (SIDE) ? u[DIRX] : -u[DIRX];                                                           \
      F = un*(RW);                                                                                             \
      if (NB) gflags |= 1u << K;                                                                               \
      if (BD) gflags |= 16u << K;                                                                              \
      if (F >= 0.0) { sxy += F; flags |= 1u << K; }                                                            \
      else if (NB) flags |= 16u << K;                                                                          \
      else if (BD) { ALT = bval; flags |= 256u << K; }                                                         \
    }
    DGB_COLFACE(0, f0, alt0, 0, 0, llx, xcy, rwx, cx > ov.origin[0] && cx <= ov.origin[0]+ov.size[0]-1, !perx && cx == 0)
    DGB_COLFACE(1, f1, alt1, 0, 1, urx, xcy, rwx, cx+1 > ov.origin[0] && cx+1 <= ov.origin[0]+ov.size[0]-1, !perx && cx+1 == v.N[0])
    DGB_COLFACE(2, f2, alt2, 1, 0, xcx, lly, rwy, cy > ov.origin[1] && cy <= ov.origin[1]+ov.size[1]-1, !pery && cy == 0)
    DGB_COLFACE(3, f3, alt3, 1, 1, xcx, ury, rwy, cy+1 > ov.origin[1] && cy+1 <= ov.origin[1]+ov.size[1]-1, !pery && cy+1 == v.N[1])
#undef DGB_COLFACE
    {
      const double fc[3] = { xcx, xcy, xcz };
      double u[3]; fv_velocity3<PROBLEM>(a.problem, fc, u);
      un4 = -u[2]; un5 = u[2];
    }
    double sum = sxy;
    if (un4 >= 0.0) sum += un4*rwmax;
    if (un5 >= 0.0) sum += un5*rwmax;
    sum_max = sum;
  }
  int dx = 2, dy = 2, dz = 2;
  {
    const bool o0 = flags & 1u, o1 = flags & 2u, o2 = flags & 4u, o3 = flags & 8u;
    if ((!o0 && o1) || (o0 && o1 && f0 == 0.0 && f1 == 0.0)) dx = 1; else if (o0 && !o1) dx = 0;
    if ((!o2 && o3) || (o2 && o3 && f2 == 0.0 && f3 == 0.0)) dy = 1; else if (o2 && !o3) dy = 0;
    dz = (un5 >= 0.0) ? 1 : 0;
  }
  fv_peer_wait_finish(a, rankBoundary, ty*BX + tx, peerFlag);
  // fused exchange: which send boxes this CTA's tile x chunk meets (read at the very end; the barriers below publish it)
  __shared__ FusedItem s_items[FUSED_MAX_BOXES];
  if (a.pack && !(a.debugSkipFaces & 0x400u))
    fv_fused_prepare(a, of, s_items, a.region.origin[0]+blockIdx.x*BX, min(BX, a.region.size[0]-(int)blockIdx.x*BX),
                     a.region.origin[1]+by*BY, min(BY, a.region.size[1]-by*BY), cz0, zn, ty*BX + tx);
  else if (a.pack && ty*BX + tx < FUSED_MAX_BOXES) s_items[ty*BX + tx].n = 0;
  // one code path per CTA (the loop contains barriers): fast only if every thread agrees on the directions
  __shared__ int s_pat[2];
  if (tx == 0 && ty == 0) { s_pat[0] = dx*4 + dy*2 + dz; s_pat[1] = 1; }
  __syncthreads();
  const int cand = s_pat[0];
  if (dx == 2 || dy == 2 || dx*4 + dy*2 + dz != cand) s_pat[1] = 0;
  __syncthreads();
  const int pat = (s_pat[1] && !GENERAL) ? cand : 8;    // general problems: all halo sides are staged, decisions per plane

  const double zlow = (!perz && ovz0 == 0) ? bval : 0.0;
  const long long idx0 = (long long)(cz0-of.origin[2])*sz + (long long)(cy-of.origin[1])*sy + (cx-of.origin[0]);
  const int kload = min(zn, ovz1-cz0);                   // last plane offset holding stored cells
  if constexpr (STG == 2) {
    FvPairLane L; FvRingSlow S;
    L.active = active;
    L.f0 = f0; L.f1 = f1; L.f2 = f2; L.f3 = f3; L.un5 = un5; L.tbx = 0.0; L.tby = 0.0;
    S.alt0 = alt0; S.alt1 = alt1; S.alt2 = alt2; S.alt3 = alt3; S.un4 = un4; S.flags = flags;
    L.zhigh = (!perz && ovz1+1 == v.N[2]) ? bval : 0.0;
    L.dt = 0.99*(1.0/(*a.slotIn));
    L.w = reinterpret_cast<char*>(a.cnew + idx0);
    const int t = ty*BX + tx;
    const int tx0 = a.region.origin[0]+blockIdx.x*BX, ty0 = a.region.origin[1]+by*BY;
    const int nxv = min(BX, a.region.size[0]-(int)blockIdx.x*BX), nyv = min(BY, a.region.size[1]-by*BY);
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(ring);
    constexpr unsigned SLOTB = (SLOT + 2*PY)*8u;           // a slot = PY rows of PX cells + 2*PY side cells (x halo of the low / high side)
    // element offset of (tx0, ty0-1, cz0) = first tile cell of staged row 0; `par` folds in the alignment of the field itself
    const long long e00 = (long long)(cz0-of.origin[2])*sz + (long long)(ty0-1-of.origin[1])*sy + (tx0-of.origin[0]);
    const unsigned par = unsigned(reinterpret_cast<uintptr_t>(a.c) >> 3) + unsigned(e00);
    const unsigned psy = unsigned(sy) & 1u, psz = unsigned(sz) & 1u;
    auto phase = [&](int rr, int kk) -> unsigned { return (par + unsigned(rr)*psy + unsigned(kk)*psz) & 1u; };
    // own cell (threads outside the region shadow the last valid column / row, like cx and cy)
    const int rrO = min(ty, nyv-1)+1, ctx = min(tx, nxv-1);
    const unsigned ph0 = phase(rrO, 0), ph1 = phase(rrO, 1);
    L.aE = ring0 + unsigned(rrO*PX + 2 + ctx + int(ph0))*8u;
    L.aO = ring0 + unsigned(rrO*PX + 2 + ctx + int(ph1))*8u;
    // hybrid exchange: does this tile touch the low / high end of the region, beyond which the x neighbour comes from the exchange
    // buffer (a side cell of the slot) instead of the field's overlap column?
    const bool sideLo = a.xh[0] && blockIdx.x == 0, sideHi = a.xh[1] && int(blockIdx.x) == int(gridDim.x)-1;
    const unsigned sideCellLo = ring0 + unsigned(SLOT + rrO)*8u, sideCellHi = ring0 + unsigned(SLOT + PY + rrO)*8u;
    const bool edgeLo = sideLo && ctx == 0, edgeHi = sideHi && ctx == nxv-1;
    // phase step towards the neighbouring rows (they have the other phase when the row stride is odd), in bytes
    const int dE = psy ? (ph0 ? -8 : 8) : 0, dO = psy ? (ph1 ? -8 : 8) : 0;
    L.xE = L.aE; L.xO = L.aO; L.yE = L.aE; L.yO = L.aO;
    S.xlE = edgeLo ? sideCellLo : L.aE - 8u; S.xlO = edgeLo ? sideCellLo : L.aO - 8u;
    S.xrE = edgeHi ? sideCellHi : L.aE + 8u; S.xrO = edgeHi ? sideCellHi : L.aO + 8u;
    if (pat < 8) {
      const int mx = dx ? 0 : 1, my = dy ? 2 : 3;
      int ox = dx ? -8 : 8, oy = dy ? -PX*8 : PX*8;
      if (!(flags >> (4+mx) & 1u) && !(flags >> mx & 1u)) { ox = 0; L.tbx = (flags >> (8+mx) & 1u) ? bval*(dx ? f0 : f1) : 0.0; if (dx) L.f0 = 0.0; else L.f1 = 0.0; }
      if (!(flags >> (4+my) & 1u) && !(flags >> my & 1u)) { oy = 0; L.tby = (flags >> (8+my) & 1u) ? bval*(dy ? f2 : f3) : 0.0; if (dy) L.f2 = 0.0; else L.f3 = 0.0; }
      if (flags >> mx & 1u) ox = 0;
      if (flags >> my & 1u) oy = 0;
      L.xE = unsigned(int(L.aE) + ox); L.xO = unsigned(int(L.aO) + ox);
      if (ox < 0 && edgeLo) L.xE = L.xO = sideCellLo;
      if (ox > 0 && edgeHi) L.xE = L.xO = sideCellHi;
      if (oy) { L.yE = unsigned(int(L.aE) + oy + dE); L.yO = unsigned(int(L.aO) + oy + dO); }
    } else { L.yE = unsigned(dE); L.yO = unsigned(dO); }
    // loader duty: pair p of staged row rr (row 0 / nyv+1: the y neighbour rows, if stored and - fast path - on the inflow side)
    constexpr int NP = PX/2, NLD = PY*NP;
    static_assert(NLD + 2*PY <= BX*BY, "pair loaders and side loaders are different threads");
    const int rr = t / NP, p = t - rr*NP;
    bool need = rr >= 1 && rr <= nyv;
    if (rr == 0 || rr == nyv+1) {
      const int gy = ty0-1+rr;
      need = gy >= ov.origin[1] && gy < ov.origin[1]+ov.size[1];
      if (need && pat < 8) need = (rr == 0) == bool(pat & 2);
    }
    L.loader = need && rr < PY;
    L.ld = ring0 + unsigned(rr*PX + 2*p)*8u;
    const char* lo = reinterpret_cast<const char*>(a.c);
    const char* hi = lo + sz*of.size[2]*8;
    const long long sz8 = sz*8;
    L.step2 = unsigned(2*sz8);
    // source of the pair in plane 0 / plane 1 of the chunk; plane j: same parity class, (j or j-1) planes further
    const char* srcE = lo + (e00 + (long long)rr*sy - 2 - (long long)phase(rr, 0) + 2*p)*8;
    const char* srcO = lo + (e00 + (long long)rr*sy + sz - 2 - (long long)phase(rr, 1) + 2*p)*8;
    long long srcStep = sz8;
    // side duty (hybrid exchange): thread NLD + side*PY + row copies the x halo cell of that tile row from the exchange buffer
    L.side = false;
    if (t >= NLD && t < NLD + 2*PY) {
      const int sd = (t - NLD)/PY, srow = (t - NLD) - sd*PY;
      bool on = (sd ? sideHi : sideLo) && srow >= 1 && srow <= nyv;
      if (on && pat < 8) on = (sd == 0) == bool(pat & 4);            // fast path: only the inflow side is read
      if (on) {
        L.side = true;
        L.ld = ring0 + unsigned(SLOT + sd*PY + srow)*8u;
        srcStep = (long long)a.xhNy[sd]*8;
        srcE = reinterpret_cast<const char*>(a.xh[sd] + ((long long)(cz0-a.xhZ0[sd])*a.xhNy[sd] + (ty0+srow-1-a.xhY0[sd])));
        srcO = srcE + srcStep;
        L.step2 = unsigned(2*srcStep);
      }
    }
#pragma unroll
    for (int j = 0; j <= D; ++j) {
      const char* src = (j & 1) ? srcO + (long long)(j-1)*srcStep : srcE + (long long)j*srcStep;
      if (j <= kload && L.loader) stage_pair_checked(L.ld + j*SLOTB, src, lo, hi);
      else if (j < zn && L.side) cp_async8_b(L.ld + j*SLOTB, src);
      cp_async_commit();
    }
    L.glO = srcO + (long long)D*srcStep;                 // plane D+1 (odd), staged by the first (even) iteration
    L.glE = srcE + (long long)(D+2)*srcStep;             // plane D+2
    L.c_below = (cz0 > ovz0) ? __ldg(a.c + idx0 - sz) : zlow;
    // cells whose staging and reads all hit stored planes; the last plane of the ARRAY is staged with bounds checks (its last
    // pair may stick out of the field by one element), and so is the plane above the chunk where side cells are in use (the
    // exchange buffer holds interior planes only)
    int kmain = max(0, min(zn, kload-D));
    if (cz0 + kload >= of.origin[2]+of.size[2]-1) kmain = max(0, min(kmain, kload-D-1));
    if (sideLo || sideHi) kmain = max(0, min(kmain, zn-D-1));
    if constexpr (GENERAL) {
      L.f0 = g_llx; L.f1 = g_urx; L.f2 = g_lly; L.f3 = g_ury; L.tbx = g_xcx; L.tby = g_xcy; L.un5 = g_rwx; S.alt0 = g_rwy; S.flags = gflags;
      FvGeneral G{ &a.problem, a.tab.xv[2] + zbase, a.tab.xc[2] + zbase, cz0, ovz0, ovz1, v.N[2], perz, 0.0 };
      fv_pair_loop<D, BX, BY, 9, ZPARAM, PROBLEM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi, &G);
      sum_max = G.sum;
    } else
    switch (pat) {
      case 0: fv_pair_loop<D, BX, BY, 0, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      case 1: fv_pair_loop<D, BX, BY, 1, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      case 2: fv_pair_loop<D, BX, BY, 2, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      case 3: fv_pair_loop<D, BX, BY, 3, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      case 4: fv_pair_loop<D, BX, BY, 4, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      case 5: fv_pair_loop<D, BX, BY, 5, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      case 6: fv_pair_loop<D, BX, BY, 6, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      case 7: fv_pair_loop<D, BX, BY, 7, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
      default: fv_pair_loop<D, BX, BY, 8, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, kmain, lo, hi); break;
    }
    cp_async_wait<0>();
    if (active && a.dtOut && i0 == 0 && i1 == 0 && z0 == 0) *a.dtOut = L.dt;
  } else if constexpr (STG == 3) {
    constexpr int R = D+2, PXT = BX+4;                    // the box starts at an even column (16-byte aligned start: a TMA rule for 8-byte elements)
    constexpr unsigned SLOTB = ((PXT*(BY+2)*8u + 127u)/128u)*128u;      // TMA destinations: 128-byte aligned
    FvTmaLane L; FvRingSlow S;
    L.active = active;
    L.f0 = f0; L.f1 = f1; L.f2 = f2; L.f3 = f3; L.un5 = un5; L.tbx = 0.0; L.tby = 0.0;
    S.alt0 = alt0; S.alt1 = alt1; S.alt2 = alt2; S.alt3 = alt3; S.un4 = un4; S.flags = flags;
    S.xlE = S.xlO = S.xrE = S.xrO = 0u;
    L.zhigh = (!perz && ovz1+1 == v.N[2]) ? bval : 0.0;
    L.dt = 0.99*(1.0/(*a.slotIn));
    L.w = reinterpret_cast<char*>(a.cnew + idx0);
    const int t = ty*BX + tx;
    const int tx0 = a.region.origin[0]+blockIdx.x*BX, ty0 = a.region.origin[1]+by*BY;
    const int nxv = min(BX, a.region.size[0]-(int)blockIdx.x*BX), nyv = min(BY, a.region.size[1]-by*BY);
    const unsigned ring0 = ((unsigned)__cvta_generic_to_shared(ring) + 127u) & ~127u;
    const unsigned bar0 = ring0 + R*SLOTB;
    // fast path: the box holds the tile rows and the y neighbour row of the inflow side; otherwise both neighbour rows
    const bool full = pat >= 8, lowRow = full || (pat & 2);
    const int rowO = min(ty, nyv-1) + (lowRow ? 1 : 0), ctx = min(tx, nxv-1);
    const int bx0 = (tx0-of.origin[0]-1) & ~1;           // even column at or below the tile's left neighbour (-2 at the array's edge)
    L.aOwn = ring0 + unsigned(rowO*PXT + (tx0-of.origin[0]-bx0) + ctx)*8u;
    L.aX = L.aOwn; L.aY = L.aOwn;
    if (pat < 8) {
      const int mx = dx ? 0 : 1, my = dy ? 2 : 3;
      int ox = dx ? -8 : 8, oy = dy ? -PXT*8 : PXT*8;
      if (!(flags >> (4+mx) & 1u) && !(flags >> mx & 1u)) { ox = 0; L.tbx = (flags >> (8+mx) & 1u) ? bval*(dx ? f0 : f1) : 0.0; if (dx) L.f0 = 0.0; else L.f1 = 0.0; }
      if (!(flags >> (4+my) & 1u) && !(flags >> my & 1u)) { oy = 0; L.tby = (flags >> (8+my) & 1u) ? bval*(dy ? f2 : f3) : 0.0; if (dy) L.f2 = 0.0; else L.f3 = 0.0; }
      if (flags >> mx & 1u) ox = 0;
      if (flags >> my & 1u) oy = 0;
      L.aX = unsigned(int(L.aOwn) + ox); L.aY = unsigned(int(L.aOwn) + oy);
    }
    L.issuer = t < 32;                                     // warp 0; one elected lane issues
    const void* map = full ? static_cast<const void*>(tmaps.full) : static_cast<const void*>(tmaps.inflow);
    const unsigned bytes = PXT*(full ? BY+2 : BY+1)*8u;
    const int by0 = ty0-of.origin[1]-(lowRow ? 1 : 0), bz0 = cz0-of.origin[2];
    if (t == 0) {
#pragma unroll
      for (int s = 0; s < R; ++s) mbar_init(bar0 + 8u*s, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      // the barriers (shared memory) for the async proxy; and - registered fields - the halo cells peers stored into c through the
      // generic proxy, acquired by fv_peer_wait_finish above, before the TMA engine (async proxy) reads them
      asm volatile("fence.proxy.async;\n" ::: "memory");
#pragma unroll
      for (int j = 0; j <= D; ++j)
        if (j <= kload) tma_stage_plane(ring0 + j*SLOTB, map, bx0, by0, bz0+j, bar0 + 8u*j, bytes);
    }
    __syncthreads();                                       // the barriers exist before anybody waits on them
    L.c_below = (cz0 > ovz0) ? __ldg(a.c + idx0 - sz) : zlow;
    switch (pat) {
      case 0: fv_tma_loop<D, BX, BY, 0, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      case 1: fv_tma_loop<D, BX, BY, 1, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      case 2: fv_tma_loop<D, BX, BY, 2, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      case 3: fv_tma_loop<D, BX, BY, 3, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      case 4: fv_tma_loop<D, BX, BY, 4, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      case 5: fv_tma_loop<D, BX, BY, 5, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      case 6: fv_tma_loop<D, BX, BY, 6, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      case 7: fv_tma_loop<D, BX, BY, 7, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
      default: fv_tma_loop<D, BX, BY, 8, ZPARAM>(L, S, sz, rwz, zt, zbase, zn, kload, map, bar0, ring0, bx0, by0, bz0, bytes); break;
    }
    if (active && a.dtOut && i0 == 0 && i1 == 0 && z0 == 0) *a.dtOut = L.dt;
  } else {
  FvRingLane L; FvRingSlow S;
  L.active = active;
  L.f0 = f0; L.f1 = f1; L.f2 = f2; L.f3 = f3; L.un5 = un5; L.tbx = 0.0; L.tby = 0.0; L.ox = 0; L.oy = 0;
  S.alt0 = alt0; S.alt1 = alt1; S.alt2 = alt2; S.alt3 = alt3; S.un4 = un4; S.flags = flags;
  L.zhigh = (!perz && ovz1+1 == v.N[2]) ? bval : 0.0;
  L.dt = 0.99*(1.0/(*a.slotIn));
  L.g = a.c + idx0; L.w = a.cnew + idx0;
  L.own = ring + (ty+1)*PX + (tx+XO);
  L.poff = 0; L.loader = false;
  if (A16) {                                           // thread t < BX*BY/2 stages the pair (2j, 2j+1) of row r
    const int t = ty*BX + tx, r = t/(BX/2), j = t - r*(BX/2);
    const int nxv = min(BX, a.region.size[0]-(int)blockIdx.x*BX), nyv = min(BY, a.region.size[1]-by*BY);
    L.loader = r < nyv && 2*j < nxv;
    const int gx = a.region.origin[0]+blockIdx.x*BX+2*j, gy = a.region.origin[1]+by*BY+min(r, nyv-1);
    if (L.loader) L.g = a.c + ((long long)(cz0-of.origin[2])*sz + (long long)(gy-of.origin[1])*sy + (gx-of.origin[0]));
    L.poff = ((r+1)*PX + XO + 2*j) - ((ty+1)*PX + tx + XO);
  }
  // halo duty: thread t stages halo cell t of the tile (left column, right column, bottom row, top row), if that
  // cell is a stored neighbour and - on the fast path, where all columns share the flow direction - on an inflow side
  {
    const int t = ty*BX + tx;
    const int nxv = min(BX, a.region.size[0]-blockIdx.x*BX), nyv = min(BY, a.region.size[1]-by*BY);   // valid columns / rows
    int hx = 0, hy = 0, side = -1;
    if (t < BY) { side = 0; hx = -1; hy = t; }
    else if (t < 2*BY) { side = 1; hx = nxv; hy = t-BY; }
    else if (t < 2*BY+BX) { side = 2; hx = t-2*BY; hy = -1; }
    else if (t < 2*BY+2*BX) { side = 3; hx = t-2*BY-BX; hy = nyv; }
    const int gx = a.region.origin[0]+blockIdx.x*BX+hx, gy = a.region.origin[1]+by*BY+hy;
    bool need = side >= 0 && (side < 2 ? hy < nyv : hx < nxv)
                && gx >= ov.origin[0] && gx < ov.origin[0]+ov.size[0] && gy >= ov.origin[1] && gy < ov.origin[1]+ov.size[1];
    if (need && pat < 8) {
      const bool px = pat & 4, py = pat & 2;
      need = (side == 0 && px) || (side == 1 && !px) || (side == 2 && py) || (side == 3 && !py);
    }
    L.hasHalo = need;
    L.hg = need ? a.c + ((long long)(cz0-of.origin[2])*sz + (long long)(gy-of.origin[1])*sy + (gx-of.origin[0])) : a.c;
    L.hoff = ((hy+1)*PX + (hx+XO)) - ((ty+1)*PX + (tx+XO));
  }
  // prologue: planes 0..D of the chunk
#pragma unroll
  for (int j = 0; j <= D; ++j) {
    if (j <= kload) {
      if (A16) { if (L.loader) cp_async16(L.own + j*SLOT + L.poff, L.g + (long long)j*sz); }
      else if (active) cp_async8(L.own + j*SLOT, L.g + (long long)j*sz);
      if (L.hasHalo) cp_async8(L.own + j*SLOT + L.hoff, L.hg + (long long)j*sz);
    }
    cp_async_commit();
  }
  L.c_below = (cz0 > ovz0) ? __ldg(a.c + idx0 - sz) : zlow;
  if (pat < 8) {
    const int mx = dx ? 0 : 1, my = dy ? 2 : 3;
    L.ox = dx ? -1 : 1; L.oy = dy ? -PX : PX;
    if (!(flags >> (4+mx) & 1u) && !(flags >> mx & 1u)) { L.ox = 0; L.tbx = (flags >> (8+mx) & 1u) ? bval*(dx ? f0 : f1) : 0.0; if (dx) L.f0 = 0.0; else L.f1 = 0.0; }
    if (!(flags >> (4+my) & 1u) && !(flags >> my & 1u)) { L.oy = 0; L.tby = (flags >> (8+my) & 1u) ? bval*(dy ? f2 : f3) : 0.0; if (dy) L.f2 = 0.0; else L.f3 = 0.0; }
    if (flags >> mx & 1u) L.ox = 0;
    if (flags >> my & 1u) L.oy = 0;
  }
  switch (pat) {
    case 0: fv_ring_loop<D, BX, BY, 0, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    case 1: fv_ring_loop<D, BX, BY, 1, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    case 2: fv_ring_loop<D, BX, BY, 2, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    case 3: fv_ring_loop<D, BX, BY, 3, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    case 4: fv_ring_loop<D, BX, BY, 4, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    case 5: fv_ring_loop<D, BX, BY, 5, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    case 6: fv_ring_loop<D, BX, BY, 6, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    case 7: fv_ring_loop<D, BX, BY, 7, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
    default: fv_ring_loop<D, BX, BY, 8, ZPARAM, A16>(L, S, sz, rwz, zt, zbase, zn, kload); break;
  }
  cp_async_wait<0>();
  if (active && a.dtOut && i0 == 0 && i1 == 0 && z0 == 0) *a.dtOut = L.dt;
  }
  const int tid = ty*BX + tx;
  if (a.pack && !(a.debugSkipFaces & 0x400u)) fv_fused_pack<BX*BY>(a, of, s_items, tid);
  for (int o = 16; o > 0; o >>= 1) sum_max = fmax(sum_max, __shfl_down_sync(0xffffffffu, sum_max, o));
  __shared__ double part[BX*BY/32];
  if ((tid & 31) == 0) part[tid >> 5] = sum_max;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int k = 0; k < BX*BY/32; ++k) m = fmax(m, part[k]);
    if (m > 0.0) atomic_max_nonneg(a.slotOut, m);
    if (a.slotZero && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) *a.slotZero = 0.0;
    if (a.pack) fv_fused_signal(a, s_items);
  }
}

// initial condition at the cell centres of all stored cells (geometry().center(), periodic wrap included)
template<int DIM>
__global__ void __launch_bounds__(FBX*FBY) k_fv_init(const __grid_constant__ dgb_view v, const FvProblem q, double* c) {
  const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
  const int i0 = blockIdx.x*FBX + threadIdx.x, i1 = blockIdx.y*FBY + threadIdx.y, i2 = blockIdx.z;
  if (i0 >= of.size[0] || i1 >= of.size[1]) return;
  int coord[3] = { of.origin[0]+i0, of.origin[1]+i1, DIM > 2 ? of.origin[2]+i2 : 0 };
  double x[DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const double sh = periodic_shift(v, i, coord[i]);
    double lo = vertex_coord(v, i, coord[i]), hi = vertex_coord(v, i, coord[i]+1);
    if (sh != 0.0) { lo += sh; hi += sh; }
    x[i] = 0.5*(lo+hi);
  }
  c[((long long)i2*of.size[1] + i1)*of.size[0] + i0] = fv_initial<DIM>(q, x);
}

// per-axis tables of the separable mode
__global__ void k_fv_tables(const __grid_constant__ dgb_view v, int axis, int first, int ncells, double* rw, double* xc, double* xv) {
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t > ncells) return;
  const double lo = vertex_coord(v, axis, first+t);
  xv[t] = lo;
  if (t < ncells) {
    const double hi = vertex_coord(v, axis, first+t+1);
    rw[t] = 1.0/(hi-lo);
    xc[t] = 0.5*(lo+hi);
  }
}

} // namespace dgb

using namespace dgb;

struct dgb_fv {
  dgb_grid* grid = nullptr;
  dgb_comm* comm = nullptr;
  int level = 0, mode = 0;
  int zchunk = 0;                // planes marched by one CTA of the 3-D kernel; 0 = auto (DGB_FV_ZCHUNK overrides)
  int variant = -1;              // tuning variant of the 3-D kernel; -1 = automatic (DGB_FV_VARIANT overrides)
  int staging = 3;               // staging of the ring kernel: 3 automatic (default), 4 tensor map (TMA) where possible, 2 phase-aware pairs, 1 aligned pairs, 0 8 bytes per cell (DGB_FV_STAGING=auto|tma|pair|a16|8)
  int tmaTile[2] = {0, 0};       // tile shape the cached tensor maps were built for
  std::vector<std::pair<const double*, dgb::FvTmaMaps>> tmaCache;   // tensor maps of the arrays the time loop alternates between
  std::unique_ptr<dgb::FvZTable> ztab;   // reciprocal z widths for the kernel parameter space (ring kernel)
  bool zparam = false;
  dgb_view view;
  dgb_mapper mapper;
  FvProblem problem;
  FvTables tab;
  double* dTables = nullptr;
  double* dSlots = nullptr;        // 4 rotating max-sum slots + dt + the constants {0, boundary value}: see slot_args
  long long step = 0;
  bool bootstrapped = false;
  bool fusedTried = false;       // fused_setup has been attempted
  int exchangeMode = DGB_FV_EXCHANGE_NONE;
  dgb::FusedState fused;         // fused exchange (pack + peer stores + signal inside the step kernel)
  bool seedPending = false;      // dgb_fv_bootstrap_local: the reduced seed still has to be copied to the second slot
  int64_t launches = 0;
  int kinfo[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // last step kernel: see dgb_fv_kernel_info
  int hostPath = 0;              // last dgb_fv_step_host: DGB_FV_HOST_*
  // hybrid exchange of registered fields: the x-overlap column of `xhArray` is stale, its values are in the exchange buffer of
  // parity `xhParity` (read there by the next step, or unpacked on demand: fv_flush_xhalo)
  bool xhPending = false; double* xhArray = nullptr; int xhParity = 0;
  bool wantPair = false;         // the next launch_step must use the phase-aware pair staging (the only reader of FvArgs::xh)
  // registered fields: the wait for the peers' flags of the last step has not been enqueued yet (it rides on the next step kernel, or
  // on dgb_fv_fence): epoch to wait for, parity of its CFL values, slot that receives their maximum
  bool waitPending = false; unsigned waitEpoch = 0; int waitParity = 0, waitSlot = 0;
  std::shared_ptr<CommPlan> plan;
  double* dStage[2] = {nullptr, nullptr};     // device staging for the host-buffer entry point
  double* dHostVote = nullptr;                // one device word: do all ranks take the pipelined host path?
  size_t stageBytes = 0;
  cudaStream_t sIn = nullptr, sOut = nullptr;  // copy streams of the pipelined host-buffer step
  // exchange shell / bulk split of the interior box (3-D, ranks with overlap cells): the shell slabs are the interior
  // cells some neighbour receives; they are updated first so that the exchange runs concurrently with the bulk
  bool split = false;
  dgb::FvRegion bulk; dgb::FvRegion shell[6]; int shellPerm[6]; int nShell = 0;
  cudaStream_t sComm = nullptr;
  cudaEvent_t evShell = nullptr, evBulk = nullptr, evDone = nullptr, evRed[2] = {nullptr, nullptr};
  cudaEvent_t tr[4] = {nullptr, nullptr, nullptr, nullptr};   // DGB_FV_TRACE: step start, shell end, bulk end, step end (timed)
  int traceLeft = 0;
  std::vector<cudaEvent_t> evIn, evK;
  cudaEvent_t evStart = nullptr, evEnd = nullptr;
};

namespace dgb {

// planes per CTA: long columns amortise the ring prologue; short ones keep >= ~2 waves of 32x8 tiles on 148 SMs
// With the fused exchange every CTA ends with its share of the pack (a load-store round trip + a system-scope fence) and the step
// ends at a kernel boundary (the flag wait), so shorter CTAs leave a shorter tail: 128 planes measured 2 % faster than 256 on the
// periodic 513^3 proxy (0.434 vs 0.444 ms), without exchange 256 and 128 are equal (0.381 / 0.382 ms).
static int auto_zchunk(const int* size, int bx, int by, bool packs) {
  const long long tiles = (long long)((size[0]+bx-1)/bx)*((size[1]+by-1)/by);
  const int resident = 148*1024/(bx*by);                       // CTAs the ring kernel keeps on the chip
  // measured with the exchange (tools/fused_proxy.py): 513^3 rank box (512 tiles): 128 planes 0.423 ms, 256 planes 0.433; 510x1024x1024
  // (1024 tiles): 128 planes 1.634 ms, 256 planes 1.567 - long columns as soon as they still give >= 2048 CTAs (~7 waves)
  int zc = (bx >= 128 && (!packs || tiles*((size[2]+255)/256) >= 2048)) ? 256 : 128;
  // round 2, session 3: chunks of equal length leave a ragged last wave; m chunks of ~224 planes and a SHORT last chunk (dispatched
  // last) fill it. Periodic 513^3 proxy with the hybrid exchange: 4 x 128 planes 0.423 ms, 2 x 256 0.429, 224 + 224 + 63 0.411,
  // 240 + 240 + 31 0.413, 208 + 208 + 95 0.417 (profiles/r02_fv_tma_probe.md)
  if (packs && bx >= 128 && zc == 128 && size[2] >= 256) {
    const int m = std::max(1, (size[2] + 112)/224);
    zc = std::max(32, int(size[2]/(m + 0.28))/8*8);
  }
  while (zc > 16 && tiles*((size[2]+zc-1)/zc) < 2*resident) zc /= 2;
  return zc;
}

// planes per CTA of the tensor-map form (14 planes in flight). Measured on 512^3 (512 tiles, 296 resident CTAs): chunks of equal
// length leave a ragged last wave (2 x 256: 0.357 ms, 4 x 128: 0.348, 3 x 171: 0.352); a short last chunk - dispatched last - fills it
// (3 x 160 + 32: 0.342, 2 x 224 + 64: 0.344, 3 x 144 + 80: 0.345). m full chunks of ~170 planes and a remainder of a fifth of a chunk.
static int tma_zchunk(const int* size, int bx, int by) {
  const long long tiles = (long long)((size[0]+bx-1)/bx)*((size[1]+by-1)/by);
  const int resident = 148*1024/(bx*by);
  const int nz = size[2];
  const int m = std::max(1, (nz + 85)/170);
  int zc = (int(std::ceil(nz/(m + 0.2))) + 7)/8*8;
  zc = std::max(16, std::min(zc, nz));
  while (zc > 16 && tiles*((nz+zc-1)/zc) < 2*resident) zc /= 2;
  return zc;
}

// the ring kernel of a registered problem; problems whose velocity depends on z exist in the phase-aware pair form only
template<int D, int BX, int BY, int MINB, bool ZP, int STG>
static auto ring_kernel(int pb) -> decltype(&k_fv_ring<0, D, BX, BY, MINB, ZP, STG>) {
  if constexpr (STG == 2) { if (pb == 2) return &k_fv_ring<2, D, BX, BY, MINB, ZP, 2>; }
  return pb == 1 ? &k_fv_ring<1, D, BX, BY, MINB, ZP, STG> : &k_fv_ring<0, D, BX, BY, MINB, ZP, STG>;
}

// ---- tensor maps of the TMA staging form (cuTensorMapEncodeTiled through the runtime's driver entry point: no link to libcuda) ----
template<int STG> struct FvTmaArg { static FvNoTma get(const FvTmaMaps&) { return FvNoTma{}; } };
template<> struct FvTmaArg<3> { static const FvTmaMaps& get(const FvTmaMaps& m) { return m; } };
static constexpr size_t fv_tma_slot_bytes(int bx, int by) { return (size_t(bx+4)*(by+2)*8 + 127)/128*128; }
typedef CUresult (*tmap_encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_t tmap_encoder() {
  static tmap_encode_t fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<tmap_encode_t>(p);
  }();
  return fn;
}
// the stored box of `base` as a tensor of doubles (x fastest); boxes of (bx+4) x (by+1) and (bx+4) x (by+2) cells of one plane
static int fv_tma_maps(dgb_fv* fv, const double* base, const dgb_box& of, int bx, int by, FvTmaMaps& out) {
  static_assert(sizeof(CUtensorMap) == 128, "FvTmaMaps holds two CUtensorMap objects");
  if (fv->tmaTile[0] != bx || fv->tmaTile[1] != by) { fv->tmaCache.clear(); fv->tmaTile[0] = bx; fv->tmaTile[1] = by; }
  for (auto& e : fv->tmaCache) if (e.first == base) { out = e.second; return DGB_OK; }
  const cuuint64_t dims[3] = { cuuint64_t(of.size[0]), cuuint64_t(of.size[1]), cuuint64_t(of.size[2]) };
  const cuuint64_t strides[2] = { dims[0]*8, dims[0]*dims[1]*8 };
  const cuuint32_t estr[3] = { 1, 1, 1 };
  for (int m = 0; m < 2; ++m) {
    const cuuint32_t box[3] = { cuuint32_t(bx+4), cuuint32_t(m ? by+2 : by+1), 1 };
    CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(m ? out.full : out.inflow);
    const CUresult r = tmap_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DGB_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
  }
  if (fv->tmaCache.size() >= 8) fv->tmaCache.erase(fv->tmaCache.begin());
  fv->tmaCache.emplace_back(base, out);
  return DGB_OK;
}

template<bool REDUCE_ONLY>
static int launch_step(dgb_fv* fv, const FvArgs& a, cudaStream_t stream) {
  const dgb_view& v = fv->view;
  dim3 block(FBX, FBY, 1);
  dim3 grid((a.region.size[0]+FBX-1)/FBX, (a.region.size[1]+FBY-1)/FBY, v.dim > 2 ? a.region.size[2] : 1);
  if (a.region.size[0] <= 0 || a.region.size[1] <= 0 || (v.dim > 2 && a.region.size[2] <= 0)) return DGB_OK;
  if (v.dim == 3) {
    // tile choice (measured, tools/fv_sweep.py): 128x4 columns for wide regions, 32x8 otherwise
    int variant = fv->variant >= 0 ? fv->variant : (a.region.size[0] >= 256 ? 3 : 0);
    if (variant != 99 && variant != 0 && variant != 2 && variant != 3 && variant != 7 && variant != 9) variant = a.region.size[0] >= 256 ? 3 : 0;
    const bool wide = variant == 3 || variant == 7;
    // staging of the ring kernel: 2 = 16-byte pairs with per-row phases (default: works for every alignment, row length and
    // region offset), 1 = 16-byte pairs of rows that are all 16-byte aligned (round-1 form, 128x4 tile only), 0 = 8 bytes per cell
    const dgb_box& ofb = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
    // measured (profiles/r02_fv_staging.jsonl): aligned pairs win where every row is 16-byte aligned (512^3 single rank: 0.350 ms vs
    // 0.360 phase-aware pairs vs 0.360 8-byte), phase-aware pairs win on the odd rank boxes of a partitioned grid (513^3: 0.382 vs 0.391)
    const bool a16ok = variant == 3 && fv->zparam && (reinterpret_cast<uintptr_t>(a.c) & 15u) == 0 && ofb.size[0] % 2 == 0
                       && (a.region.origin[0]-ofb.origin[0]) % 2 == 0 && a.region.size[0] % 2 == 0;
    // tensor-map staging (kernel form 3): strides that are multiples of 16 bytes, no side cells (hybrid exchange)
    const bool tmaok = variant == 3 && fv->zparam && (reinterpret_cast<uintptr_t>(a.c) & 15u) == 0 && ofb.size[0] % 2 == 0
                       && !fv->wantPair && !a.xh[0] && !a.xh[1] && tmap_encoder() != nullptr;
    // measured on the 512^3 box (profiles/r02_fv_tma_probe.md): tensor-map staging with 14 planes in flight and chunks of 160 planes
    // 0.342 ms, aligned pairs 0.352 ms; periodic 512^3-stored box with the in-kernel pack 0.4626 against 0.4757 ms. The automatic
    // choice takes it wherever it is eligible (not with side cells - hybrid exchange)
    int stg = fv->staging;
    if (stg == 4) stg = tmaok ? 33 : 3;                 // DGB_FV_STAGING=tma where possible
    // (not for thin z-slabs - the pipelined host path launches 16-plane slabs, where a ring of 16 slots never turns over: its
    // end-to-end rate fell from 5.70 to 5.5 Gcell/s with the tensor-map form)
    if (stg == 3) stg = (tmaok && a.region.size[2] >= 64) ? 33 : a16ok ? 1 : 2;      // automatic
    if (stg == 33) stg = 3;
    if (fv->wantPair) stg = 2;
    if (stg == 1 && !a16ok) stg = 0;
    if (stg == 0 && !(variant == 0 || variant == 3)) stg = 2;
    static const int tmaShape = [] { const char* e = getenv("DGB_FV_TMA_SHAPE"); return e ? atoi(e) : 1; }();
    const int zc = fv->zchunk > 0 ? fv->zchunk
                   : (stg == 3 && !a.pack && tmaShape == 1) ? tma_zchunk(a.region.size, 128, 4)
                   : auto_zchunk(a.region.size, wide ? 128 : 32, wide ? 4 : 8, a.pack != nullptr);
    const bool exact = fv->mode == DGB_FV_EXACT;
    const int pb = fv->problem.id;
    // the ring kernel evaluates the x/y face factors once per COLUMN: only for problems that declare a z-independent velocity and a
    // constant inflow value (column_invariant); every other registered problem runs on the general z-march kernel
    // (round 2: problems that do not declare it run on the same kernel in its general mode - velocity functor and upwind decisions
    // per plane, phase-aware pair staging - instead of the z-march kernel: 1.10 -> see DESIGN.md 4.1)
    static const bool generalRing = [] { const char* e = getenv("DGB_FV_GENERAL_RING"); return !e || atoi(e) != 0; }();
    const bool invariant = fv_problem_column_invariant(pb);
    const bool ring = !exact && !(REDUCE_ONLY || a.update || !a.cnew) && variant != 99 && (invariant || generalRing);
    if (ring && !invariant) stg = 2;
    if (!REDUCE_ONLY) {
      int bx = 128, by = 4;                                                   // the z-march kernels
      if (ring) switch (variant) { case 2: bx = 64; by = 4; break; case 3: case 7: bx = 128; by = 4; break; case 9: bx = 64; by = 8; break; default: bx = 32; by = 8; }
      const int info[8] = { ring ? DGB_FV_KERNEL_RING : DGB_FV_KERNEL_MARCH, ring ? variant : 99, bx, by,
                            ring ? (stg == 3 ? (bx+4)*(by+1)*8 : stg ? 16 : 8) : 0, zc, a.pack ? 1 : 0, ring ? stg : 0 };
      for (int k = 0; k < 8; ++k) fv->kinfo[k] = info[k];
    }
#define DGB_MARCH(MODE_, PROB_, PF_, BX_, BY_)                                                                       \
    { dim3 mblock(BX_, BY_, 1);                                                                                      \
      dim3 mgrid((a.region.size[0]+BX_-1)/BX_, (a.region.size[1]+BY_-1)/BY_, (a.region.size[2]+zc-1)/zc);          \
      k_fv_march<MODE_, PROB_, REDUCE_ONLY, PF_, BX_, BY_><<<mgrid, mblock, 0, stream>>>(v, a, zc); }
#define DGB_MARCH_ANY(MODE_)                                                                                         \
    { if (pb == 0) DGB_MARCH(MODE_, 0, 4, 128, 4) else if (pb == 1) DGB_MARCH(MODE_, 1, 4, 128, 4) else DGB_MARCH(MODE_, 2, 4, 128, 4) }
    static_assert(DGB_FV_PROBLEMS == 3, "add the new problem to DGB_MARCH_ANY");
    if (exact) DGB_MARCH_ANY(DGB_FV_EXACT)
    else if (!ring) DGB_MARCH_ANY(DGB_FV_SEPARABLE)
    else {
      FvTmaMaps tmaps{};
      // shapes of the tensor-map form (DGB_FV_TMA_SHAPE): 1 = 128x4 tile, 14 planes in flight (default); 0 = 6 planes in flight
      // (a 256x2 tile would need a 260-cell box row: cuTensorMapEncodeTiled refuses box dimensions above 256)
      if (stg == 3) { if (int rc = fv_tma_maps(fv, a.c, ofb, 128, 4, tmaps)) return rc; }
#define DGB_RING_LAUNCH(D_, BX_, BY_, MINB_, ZP_, STG_)                                                              \
    { auto kf = ring_kernel<D_, BX_, BY_, MINB_, ZP_, STG_>(pb);                                                     \
      DGB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));                    \
      kf<<<mgrid, mblock, smem, stream>>>(v, a, *fv->ztab, zc, FvTmaArg<STG_>::get(tmaps)); }
#define DGB_RING(D_, BX_, BY_, MINB_, WITH8_, WITHA16_)                                                             \
    { dim3 mblock(BX_, BY_, 1);                                                                                      \
      dim3 mgrid((a.region.size[0]+BX_-1)/BX_, (a.region.size[1]+BY_-1)/BY_, (a.region.size[2]+zc-1)/zc);          \
      const size_t smem = stg == 3 ? size_t(D_+2)*(fv_tma_slot_bytes(BX_, BY_) + 8) + 128                          \
                          : size_t(D_+2)*((BX_+(stg ? 4 : 2))*(BY_+2) + (stg == 2 ? 2*(BY_+2) : 0))*sizeof(double);  \
      if (WITHA16_ && stg == 3) DGB_RING_LAUNCH(D_, BX_, BY_, MINB_, true, (WITHA16_ ? 3 : 2))                       \
      else if (WITHA16_ && stg == 1) DGB_RING_LAUNCH(D_, BX_, BY_, MINB_, true, (WITHA16_ ? 1 : 2))                  \
      else if (WITH8_ && stg == 0) { if (fv->zparam) DGB_RING_LAUNCH(D_, BX_, BY_, MINB_, true, (WITH8_ ? 0 : 2))    \
                                     else DGB_RING_LAUNCH(D_, BX_, BY_, MINB_, false, (WITH8_ ? 0 : 2)) }            \
      else if (fv->zparam) DGB_RING_LAUNCH(D_, BX_, BY_, MINB_, true, 2)                                             \
      else DGB_RING_LAUNCH(D_, BX_, BY_, MINB_, false, 2) }
#define DGB_RING_TMA(D_, BX_, BY_)                                                                                  \
    { dim3 mblock(BX_, BY_, 1);                                                                                      \
      dim3 mgrid((a.region.size[0]+BX_-1)/BX_, (a.region.size[1]+BY_-1)/BY_, (a.region.size[2]+zc-1)/zc);          \
      const size_t smem = size_t(D_+2)*(fv_tma_slot_bytes(BX_, BY_) + 8) + 128;                                     \
      DGB_RING_LAUNCH(D_, BX_, BY_, 2, true, 3) }
      if (stg == 3 && tmaShape == 1) DGB_RING_TMA(14, 128, 4)
      else
      switch (variant) {                      // tuning variants: ring depth, tile, occupancy target (tools/fv_sweep.py)
        default:
        case 0: DGB_RING(8, 32, 8, 4, true, false) break;                  // production choice, narrow regions
        case 2: DGB_RING(6, 64, 4, 4, false, false) break;
        case 3: DGB_RING(6, 128, 4, 2, true, true) break;                  // production choice, wide regions
        case 7: DGB_RING(8, 128, 4, 2, false, false) break;
        case 9: DGB_RING(8, 64, 8, 2, false, false) break;
        // (tried: the phase-aware pair form with 14 planes in flight, 128x4 - 0.454 ms kernel only on the 513^3 box against 0.384 with 6)
      }
#undef DGB_RING_TMA
#undef DGB_RING
#undef DGB_RING_LAUNCH
    }
#undef DGB_MARCH_ANY
#undef DGB_MARCH
  } else {
    if (!REDUCE_ONLY) { const int info[8] = { DGB_FV_KERNEL_STEP2D, -1, FBX, FBY, 0, 1, 0, 0 }; for (int k = 0; k < 8; ++k) fv->kinfo[k] = info[k]; }
    if (fv->mode == DGB_FV_EXACT) k_fv_step<2, DGB_FV_EXACT, REDUCE_ONLY><<<grid, block, 0, stream>>>(v, a);
    else k_fv_step<2, DGB_FV_SEPARABLE, REDUCE_ONLY><<<grid, block, 0, stream>>>(v, a);
  }
  ++fv->launches;
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

// all exchange-shell slabs in one launch; pack = true also fills the send buffer of the FV exchange plan
static int launch_shell(dgb_fv* fv, const FvArgs& a, bool pack, cudaStream_t stream) {
  if (!fv->nShell) return DGB_OK;
  FvShell sh{};
  long long most = 0;
  for (int k = 0; k < fv->nShell; ++k) {
    sh.slab[k] = fv->shell[k]; sh.perm[k] = fv->shellPerm[k];
    most = std::max(most, (long long)fv->shell[k].size[0]*fv->shell[k].size[1]*fv->shell[k].size[2]);
  }
  sh.n = fv->nShell;
  if (pack) {
    if (int rc = fv->plan->upload()) return rc;
    sh.boxes = fv->plan->dSendBoxes; sh.nBoxes = fv->plan->nSendBoxes; sh.sendBuf = fv->plan->dSendBuf;
  }
  dim3 grid((unsigned)std::min<long long>((most+255)/256, 4096), (unsigned)fv->nShell, 1);
  if (fv->mode == DGB_FV_EXACT) k_fv_shell<DGB_FV_EXACT><<<grid, 256, 0, stream>>>(fv->view, a, sh);
  else k_fv_shell<DGB_FV_SEPARABLE><<<grid, 256, 0, stream>>>(fv->view, a, sh);
  ++fv->launches;
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

// interior box = bulk + up to six disjoint shell slabs (z slabs over the full x-y extent, then y, then x)
static void compute_split(dgb_fv* fv) {
  const dgb_view& v = fv->view;
  fv->split = false; fv->nShell = 0;
  if (v.dim != 3) return;
  const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
  const dgb_box& in = v.comp[0].box[DGB_BOX_INTERIOR];
  const int t = v.overlap;
  if (t <= 0) return;
  FvRegion cur;
  for (int i = 0; i < 3; ++i) { cur.origin[i] = in.origin[i]; cur.size[i] = in.size[i]; }
  bool any = false;
  for (int axis = 2; axis >= 0; --axis) {
    const bool lo = of.origin[axis] < in.origin[axis], hi = of.origin[axis]+of.size[axis] > in.origin[axis]+in.size[axis];
    if (cur.size[axis] < (lo ? t : 0) + (hi ? t : 0) + 1) { fv->nShell = 0; return; }        // no bulk left: keep one launch
    if (lo) { FvRegion r = cur; r.size[axis] = t; fv->shellPerm[fv->nShell] = axis == 0; fv->shell[fv->nShell++] = r;
              cur.origin[axis] += t; cur.size[axis] -= t; any = true; }
    if (hi) { FvRegion r = cur; r.origin[axis] = cur.origin[axis]+cur.size[axis]-t; r.size[axis] = t;
              fv->shellPerm[fv->nShell] = axis == 0; fv->shell[fv->nShell++] = r; cur.size[axis] -= t; any = true; }
  }
  fv->bulk = cur;
  fv->split = any;
}

static FvArgs base_args(dgb_fv* fv) {
  FvArgs a{};
  const dgb_box& in = fv->view.comp[0].box[DGB_BOX_INTERIOR];
  for (int i = 0; i < 3; ++i) { a.region.origin[i] = in.origin[i]; a.region.size[i] = i < fv->view.dim ? in.size[i] : 1; }
  a.problem = fv->problem; a.tab = fv->tab; a.consts = fv->dSlots + 5;
  static const unsigned skip = [] { const char* e = getenv("DGB_FV_SKIP_FACES"); return e ? unsigned(atoi(e)) : 0u; }();
  a.debugSkipFaces = skip;
  return a;
}

// CFL slots. The outflow sums depend on geometry and velocity only, and the velocity fields are time independent, so the
// value reduced while step n runs is used by step n+2: its all-reduce over the ranks then has a whole step to complete
// off the critical path. Step n reads slot n%4 (reduced), accumulates its own maximum into slot (n+2)%4 and clears
// slot (n+3)%4 for step n+1. [4] = dt of the last step, [5..6] = {0, boundary value}.
static void slot_args(dgb_fv* fv, long long n, FvArgs& a) {
  a.slotIn = fv->dSlots + (n % 4); a.slotOut = fv->dSlots + ((n+2) % 4); a.slotZero = fv->dSlots + ((n+3) % 4);
  a.dtOut = fv->dSlots + 4;
}

// geometry-only reduction that seeds the first two steps' dt
static int bootstrap(dgb_fv* fv, cudaStream_t stream) {
  DGB_CUDA(cudaMemsetAsync(fv->dSlots, 0, 5*sizeof(double), stream));
  FvArgs a = base_args(fv);
  a.slotOut = fv->dSlots + 0;
  if (int rc = launch_step<true>(fv, a, stream)) return rc;
  if (int rc = nccl_allreduce(fv->dSlots + 0, 1, fv->comm, stream)) return rc;
  DGB_CUDA(cudaMemcpyAsync(fv->dSlots + 1, fv->dSlots + 0, sizeof(double), cudaMemcpyDeviceToDevice, stream));   // steps 0 and 1
  fv->bootstrapped = true; fv->step = 0;
  return DGB_OK;
}

} // namespace dgb

extern "C" {

int dgb_fv_create(dgb_grid* g, int level, int problem, const double* params_host, int nparams, int mode, dgb_comm* comm, dgb_fv** out) {
  return guarded([&]() -> int {
    if (int rc = require_device()) return rc;
    if (problem < 0 || problem >= DGB_FV_PROBLEMS) throw ArgError("unknown problem id (the compiled-in FV problems are listed in dgb_functors.cuh)");
    if (mode != DGB_FV_EXACT && mode != DGB_FV_SEPARABLE) throw ArgError("unknown fv mode");
    if (nparams > 12) throw ArgError("at most 12 problem parameters");
    std::unique_ptr<dgb_fv> fv(new dgb_fv);
    fv->grid = g; fv->comm = comm; fv->level = level; fv->mode = mode;
    if (const char* e = getenv("DGB_FV_ZCHUNK")) { int z = atoi(e); if (z > 0) fv->zchunk = z; }
    if (const char* e = getenv("DGB_FV_VARIANT")) fv->variant = atoi(e);
    if (const char* e = getenv("DGB_FV_STAGING")) fv->staging = !strcmp(e, "8") ? 0 : !strcmp(e, "a16") ? 1 : !strcmp(e, "pair") ? 2 : !strcmp(e, "tma") ? 4 : 3;
    if (int rc = make_view(g, level, &fv->view)) return rc;
    if (g->me.nranks > 1 && !comm) throw ArgError("a dgb_comm is required on a multi-rank grid");
    const uint32_t block[4] = {1,0,0,0};
    dgb_mapper_init(&fv->view, block, &fv->mapper);
    fv->problem.id = problem;
    for (int i = 0; i < 12; ++i) fv->problem.p[i] = i < nparams ? params_host[i] : 0.0;
    DGB_CUDA(cudaMalloc(&fv->dSlots, 7*sizeof(double)));
    DGB_CUDA(cudaMemset(fv->dSlots, 0, 7*sizeof(double)));
    DGB_CUDA(cudaMemcpy(fv->dSlots + 6, &fv->problem.p[3], sizeof(double), cudaMemcpyHostToDevice));   // {0, b}: FvArgs::consts
    // per-axis tables over the stored cell range
    const dgb_view& v = fv->view;
    const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
    size_t total = 0;
    for (int i = 0; i < v.dim; ++i) total += 3*(size_t(of.size[i])+1);
    DGB_CUDA(cudaMalloc(&fv->dTables, total*sizeof(double)));
    double* p = fv->dTables;
    for (int i = 0; i < 3; ++i) { fv->tab.rw[i] = fv->tab.xc[i] = fv->tab.xv[i] = nullptr; fv->tab.first[i] = 0; }
    for (int i = 0; i < v.dim; ++i) {
      const int n = of.size[i];
      double* rw = p; double* xc = p + (n+1); double* xv = p + 2*(n+1);
      p += 3*(n+1);
      k_fv_tables<<<(n+1+127)/128, 128>>>(v, i, of.origin[i], n, rw, xc, xv);
      fv->tab.rw[i] = rw; fv->tab.xc[i] = xc; fv->tab.xv[i] = xv; fv->tab.first[i] = of.origin[i];
      ++fv->launches;
    }
    DGB_CUDA(cudaGetLastError());
    DGB_CUDA(cudaDeviceSynchronize());
    // the device-computed reciprocal z widths also travel in the kernel parameter space (k_fv_ring, FvZTable)
    fv->ztab.reset(new FvZTable);
    std::memset(fv->ztab.get(), 0, sizeof(FvZTable));
    fv->zparam = v.dim == 3 && of.size[2] <= DGB_RWZ_MAX;
    if (fv->zparam) DGB_CUDA(cudaMemcpy(fv->ztab->rw, fv->tab.rw[2], of.size[2]*sizeof(double), cudaMemcpyDeviceToHost));
    const int32_t items[1] = {1};
    // a plan of its own (not the grid's cache): the FV exchange may run while the user calls dgb_communicate with the same
    // layout on another stream, and the two must not share send / receive buffers
    fv->plan = build_plan(g, level, fv->mapper, DGB_InteriorBorder_All_Interface, DGB_ForwardCommunication, items, 1);
    compute_split(fv.get());
    if (getenv("DGB_FV_NO_SPLIT")) fv->split = false;
    if (const char* e = getenv("DGB_FV_TRACE")) fv->traceLeft = atoi(e);
    if (fv->split) {
      int prLow = 0, prHigh = 0;
      DGB_CUDA(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
      DGB_CUDA(cudaStreamCreateWithPriority(&fv->sComm, cudaStreamNonBlocking, prHigh));   // its small kernels must not queue behind the bulk's CTAs
      DGB_CUDA(cudaEventCreateWithFlags(&fv->evShell, cudaEventDisableTiming));
      DGB_CUDA(cudaEventCreateWithFlags(&fv->evBulk, cudaEventDisableTiming));
      DGB_CUDA(cudaEventCreateWithFlags(&fv->evDone, cudaEventDisableTiming));
      DGB_CUDA(cudaEventCreateWithFlags(&fv->evRed[0], cudaEventDisableTiming));
      DGB_CUDA(cudaEventCreateWithFlags(&fv->evRed[1], cudaEventDisableTiming));
    }
    *out = fv.release();
    return DGB_OK;
  });
}

static int fv_free(dgb_fv* fv, bool collective) {
  if (!fv) return DGB_OK;
  int rc = DGB_OK;
  // the fused exchange block is mapped by the other ranks: collective teardown (close, barrier, free) - or, when the caller
  // cannot promise that all ranks are here (dgb_fv_abandon), close the imports and leak the exported block
  if (collective) rc = fused_teardown(fv->fused, fv->comm, nullptr);
  else { cudaDeviceSynchronize(); fv->fused.release(!fv->fused.multi); }
  if (fv->dSlots) cudaFree(fv->dSlots);
  if (fv->dTables) cudaFree(fv->dTables);
  for (int i = 0; i < 2; ++i) if (fv->dStage[i]) cudaFree(fv->dStage[i]);
  if (fv->dHostVote) cudaFree(fv->dHostVote);
  for (cudaEvent_t e : fv->evIn) cudaEventDestroy(e);
  for (cudaEvent_t e : fv->evK) cudaEventDestroy(e);
  if (fv->evStart) cudaEventDestroy(fv->evStart);
  if (fv->evEnd) cudaEventDestroy(fv->evEnd);
  if (fv->sIn) cudaStreamDestroy(fv->sIn);
  if (fv->sComm) cudaStreamDestroy(fv->sComm);
  if (fv->evShell) cudaEventDestroy(fv->evShell);
  if (fv->evBulk) cudaEventDestroy(fv->evBulk);
  if (fv->evDone) cudaEventDestroy(fv->evDone);
  for (int k = 0; k < 2; ++k) if (fv->evRed[k]) cudaEventDestroy(fv->evRed[k]);
  if (fv->sOut) cudaStreamDestroy(fv->sOut);
  delete fv;
  return rc;
}

int dgb_fv_destroy(dgb_fv* fv) { return fv_free(fv, true); }
int dgb_fv_abandon(dgb_fv* fv) { return fv_free(fv, false); }

int dgb_fv_init_field(dgb_fv* fv, double* d_c, void* stream) {
  const dgb_view& v = fv->view;
  if (v.coord_mode == DGB_COORD_TENSORPRODUCT) for (int i = 0; i < v.dim; ++i) if (!v.tensor[i]) return fail(DGB_ECUDA, "tensor coordinates are not on the device");
  const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
  dim3 block(FBX, FBY, 1), grid((of.size[0]+FBX-1)/FBX, (of.size[1]+FBY-1)/FBY, v.dim > 2 ? of.size[2] : 1);
  if (v.dim == 3) k_fv_init<3><<<grid, block, 0, (cudaStream_t)stream>>>(v, fv->problem, d_c);
  else k_fv_init<2><<<grid, block, 0, (cudaStream_t)stream>>>(v, fv->problem, d_c);
  ++fv->launches;
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

// the send boxes of this step for the kernel's parameter space: origins / sizes from the plan, destinations by exchange form
static void fill_send_boxes(FvArgs& a, const CommPlan& plan, const std::vector<double*>& dst, const std::vector<int>* row, const std::vector<int>* plane) {
  a.send.n = plan.nSendBoxes;
  for (int b = 0; b < plan.nSendBoxes && b < 32; ++b) {
    const PlanBox& pb = plan.hostSend[b];
    FvSendBoxes::Box& o = a.send.box[b];
    for (int i = 0; i < 3; ++i) { o.origin[i] = pb.origin[i]; o.size[i] = pb.size[i]; }
    o.row = row ? (*row)[b] : pb.size[0]; o.plane = plane ? (*plane)[b] : pb.size[0]*pb.size[1];
    o.dst = dst[b];
    if (pb.face >= 0 && (a.debugSkipFaces >> pb.face & 1u)) o.size[0] = 0;          // developer knob: empty box
  }
}

// the step would run on the ring kernel (the only one that can read its x halo from the exchange buffer)
static bool fv_ring_eligible(const dgb_fv* fv) {
  return fv->view.dim == 3 && fv->mode == DGB_FV_SEPARABLE && fv->variant != 99 && fv_problem_column_invariant(fv->problem.id);
}
// on-demand unpack of the x-face receive boxes into the array whose overlap column the hybrid exchange left stale
static int fv_flush_xhalo(dgb_fv* fv, cudaStream_t stream) {
  if (fv->waitPending) {                                  // the deferred flag wait (+ CFL reduction) of the last registered step
    fv->waitPending = false;
    FusedState& fs = fv->fused;
    k_fv_wait<<<1, 64, 0, stream>>>(fs.ctl, fv->view.nranks, fv->waitEpoch, fv->waitParity, fv->dSlots + fv->waitSlot, fs.dTimeout);
    ++fv->launches;
    DGB_CUDA(cudaGetLastError());
  }
  if (!fv->xhPending) return DGB_OK;
  fv->xhPending = false;
  FusedState& fs = fv->fused;
  const long long before = fv->plan->launches;
  if (int rc = plan_unpack_boxes(*fv->plan, fs.hostXRecv, fs.dXRecv, fs.recv[fv->xhParity], fv->xhArray, stream)) return rc;
  fv->launches += fv->plan->launches - before;
  return DGB_OK;
}

// collective, once per object: try to set up the fused peer-memory exchange (see dgb_fv_step)
static int ensure_fused(dgb_fv* fv, cudaStream_t stream) {
  if (fv->fusedTried) return DGB_OK;
  fv->fusedTried = true;
  const bool exchanges = fv->plan->nSendBoxes || fv->plan->nRecvBoxes;
  const char* e = getenv("DGB_FV_FUSED");
  if (!(e && atoi(e) == 0) && fv->view.dim == 3 && (exchanges || fv->view.nranks > 1) && (fv->view.nranks == 1 || fv->comm)) {
    const dgb_box& in = fv->view.comp[0].box[DGB_BOX_INTERIOR];
    const bool nonEmpty = in.size[0] > 0 && in.size[1] > 0 && in.size[2] > 0;    // every rank must launch a step kernel
    if (int rc = fused_setup(*fv->plan, fv->view.rank, fv->view.nranks, fv->comm, stream, fv->fused, nonEmpty)) return rc;
  }
  return DGB_OK;
}

static const char* const kTimeoutMsg = "fused exchange: a rank did not deliver its step within 20 s (ranks out of step, or a peer has died); "
                                       "this dgb_fv object accepts no further steps";

int dgb_fv_step(dgb_fv* fv, const double* d_c_in, double* d_c_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (fv->fused.timed_out()) return fail(DGB_ENCCL, kTimeoutMsg);           // sticky: nothing is launched after a time-out
  if (fv->comm) if (int rc = nccl_check_async(fv->comm)) return rc;
  if (!fv->bootstrapped) if (int rc = bootstrap(fv, stream)) return rc;
  const long long n = fv->step;
  FvArgs a = base_args(fv);
  a.c = d_c_in; a.cnew = d_c_out;
  slot_args(fv, n, a);
  const bool exchanges = fv->plan->nSendBoxes || fv->plan->nRecvBoxes;
  // Fused form (3-D, default): ONE step kernel over the whole interior packs the send boxes straight into the receivers'
  // buffers (peer-mapped memory, NVLink) while it computes and raises a flag when the rank has delivered; a one-CTA kernel
  // waits for all ranks' flags (which also carry the CFL maxima: no all-reduce), then the overlap cells are scattered.
  // One stream, no NCCL call in the steady state. Collective set-up on the first step; DGB_FV_FUSED=0 or ranks without
  // peer mapping keep the NCCL forms below.
  if (int rc = ensure_fused(fv, stream)) return rc;
  if (fv->fused.on) {
    FusedState& fs = fv->fused;
    const int parity = int(fs.epoch & 1u);
    // registered output array (dgb_fv_register_fields): the kernel stores the halo cells straight into the receivers' arrays;
    // nothing to unpack, only the one-CTA wait for all ranks' flags (+ CFL reduction) remains
    const int slot = fs.direct_slot(d_c_out);
    // hybrid form (registered output, ring kernel, overlap 1): the x-face boxes - one cell per row of the receiver, every store a
    // DRAM read-modify-write of a 32-byte sector - keep going to the receiver's contiguous exchange buffer; the receiver's next
    // step reads its x neighbours from there and its overlap column is refreshed on demand only (dgb_fv_fence)
    static const bool hybridEnv = [] { const char* e = getenv("DGB_FV_HYBRID"); return !(e && atoi(e) == 0); }();
    const bool ringOk = fv_ring_eligible(fv);
    const bool hybrid = slot >= 0 && hybridEnv && fs.hybridOk && ringOk && (fs.xface[0].on || fs.xface[1].on);
    // input side: x halos of c_in still in the exchange buffer?
    if (fv->xhPending) {
      if (ringOk && d_c_in == fv->xhArray) {
        for (int sd = 0; sd < 2; ++sd) if (fs.xface[sd].on) {
          a.xh[sd] = fs.recv[fv->xhParity] + fs.xface[sd].buf_offset; a.xhY0[sd] = fs.xface[sd].y0; a.xhZ0[sd] = fs.xface[sd].z0; a.xhNy[sd] = fs.xface[sd].ny;
        }
        fv->wantPair = true;
      } else if (int rc = fv_flush_xhalo(fv, stream)) return rc;
    }
    a.pack = (slot >= 0 ? (hybrid ? fs.direct[slot].dHybrid : fs.direct[slot].dPack) : fs.dPack) + parity; a.packEpoch = fs.epoch + 1u; a.packSignal = 1;
    if (slot < 0) fill_send_boxes(a, *fv->plan, fs.hostDst[parity], nullptr, nullptr);
    else if (hybrid) fill_send_boxes(a, *fv->plan, fs.direct[slot].hDstH[parity], &fs.direct[slot].hRowH, &fs.direct[slot].hPlaneH);
    else fill_send_boxes(a, *fv->plan, fs.direct[slot].hDst, &fs.direct[slot].hRow, &fs.direct[slot].hPlane);
    static const bool debugNoWait = [] { const char* e = getenv("DGB_FV_DEBUG_NOWAIT"); return e && atoi(e) != 0; }();   // developer knob (timing only)
    if (fv->waitPending && debugNoWait) fv->waitPending = false;
    if (fv->waitPending) {                               // the previous registered step left its flag wait to this kernel's boundary CTAs
      a.waitCtl = fs.ctl; a.waitEpoch = fv->waitEpoch; a.waitParity = fv->waitParity; a.waitRanks = fv->view.nranks;
      a.waitSlot = fv->dSlots + fv->waitSlot; a.waitTimeout = fs.dTimeout;
      fv->waitPending = false;
    }
    const int rcl = launch_step<false>(fv, a, stream);
    const bool readBuffer = fv->wantPair;
    fv->wantPair = false;
    if (rcl) return rcl;
    if (readBuffer) fv->xhPending = false;               // consumed (the array itself was never refreshed; it is an input only)
    if (hybrid) { fv->xhPending = true; fv->xhArray = d_c_out; fv->xhParity = parity; }
    static const bool mergedEnv = [] { const char* e = getenv("DGB_FV_MERGED_WAIT"); return !(e && atoi(e) == 0); }();
    if (slot >= 0 && mergedEnv) {
      // registered output: ONE kernel per step. The wait for this step's flags is done by the boundary CTAs of the next step
      // kernel (interior CTAs start at once and absorb the skew between the ranks), or by dgb_fv_fence
      fv->waitPending = true; fv->waitEpoch = fs.epoch + 1u; fv->waitParity = parity; fv->waitSlot = int((n+2) % 4);
    } else if (slot >= 0 || fv->plan->nRecvBoxes == 0) {
      k_fv_wait<<<1, 64, 0, stream>>>(fs.ctl, fv->view.nranks, fs.epoch + 1u, parity, fv->dSlots + ((n+2) % 4), fs.dTimeout); ++fv->launches;
    }
    if (slot < 0) {
      const int32_t items[1] = {1};
      double* arrays[1] = { d_c_out };
      const long long before = fv->plan->launches;
      // the wait for all ranks' flags (+ CFL reduction) rides on the first unpack launch
      const HaloWait hw{fs.ctl, fv->view.nranks, fs.epoch + 1u, parity, fv->dSlots + ((n+2) % 4), fs.dTimeout};
      if (int rc = plan_unpack_from(*fv->plan, fs.recv[parity], DGB_OP_SET, arrays, items, 1, stream, &hw)) return rc;
      fv->launches += fv->plan->launches - before;
    }
    DGB_CUDA(cudaGetLastError());
    int64_t bytes = 0;
    for (auto& sg : fv->plan->sendSeg) if (sg.peer != fv->view.rank) bytes += sg.count*8;
    fv->grid->lastBytesSent = bytes;
    ++fs.epoch;
    fv->step = n+1;
    fv->exchangeMode = slot >= 0 ? (hybrid ? DGB_FV_EXCHANGE_HYBRID : DGB_FV_EXCHANGE_DIRECT) : DGB_FV_EXCHANGE_FUSED;
    return DGB_OK;
  }
  if (fv->split && exchanges) {
    // overlapped form. Caller's stream: bulk. Exchange stream (high priority; every NCCL call of the step): shell
    // slabs with in-kernel pack, NCCL send/recv, unpack; [after the bulk] all-reduce of the CFL slot, which
    // step n+2 reads - the caller's stream waits for it two steps later, and for the exchange at the end of this step.
    const bool trace = fv->traceLeft > 0;
    if (trace && !fv->tr[0]) { for (int k = 0; k < 4; ++k) cudaEventCreate(&fv->tr[k]); for (int k = 0; k < 3; ++k) cudaEventCreate(&fv->plan->trace[k]); }
    if (n >= 2) DGB_CUDA(cudaStreamWaitEvent(stream, fv->evRed[n % 2], 0));           // slot n%4 was reduced during step n-1
    if (trace) cudaEventRecord(fv->tr[0], stream);
    // the shell runs on the exchange stream, concurrently with the bulk: its strided, latency-bound accesses hide behind
    // the streaming bulk kernel, and the stream's priority schedules its CTAs ahead of the bulk's
    DGB_CUDA(cudaEventRecord(fv->evShell, stream));                                    // c_in is ready, slot n%4 is reduced
    DGB_CUDA(cudaStreamWaitEvent(fv->sComm, fv->evShell, 0));
    if (int rc = launch_shell(fv, a, true, fv->sComm)) return rc;
    if (trace) cudaEventRecord(fv->tr[1], fv->sComm);
    FvArgs b = a; b.region = fv->bulk;
    if (int rc = launch_step<false>(fv, b, stream)) return rc;
    DGB_CUDA(cudaEventRecord(fv->evBulk, stream));
    if (trace) cudaEventRecord(fv->tr[2], stream);
    {
      const int32_t items[1] = {1};
      double* arrays[1] = { d_c_out };
      const long long before = fv->plan->launches;
      if (int rc = run_plan(fv->grid, *fv->plan, fv->view.rank, DGB_OP_SET, arrays, items, 1, fv->comm, fv->sComm, true)) return rc;
      fv->launches += fv->plan->launches - before;
    }
    DGB_CUDA(cudaEventRecord(fv->evDone, fv->sComm));
    DGB_CUDA(cudaStreamWaitEvent(stream, fv->evDone, 0));                              // overlap cells of c_out are in place
    DGB_CUDA(cudaStreamWaitEvent(fv->sComm, fv->evBulk, 0));
    if (int rc = nccl_allreduce(fv->dSlots + ((n+2) % 4), 1, fv->comm, fv->sComm)) return rc;
    DGB_CUDA(cudaEventRecord(fv->evRed[n % 2], fv->sComm));
    fv->step = n+1;
    fv->exchangeMode = DGB_FV_EXCHANGE_OVERLAPPED;
    if (trace) {                                       // developer timeline (synchronises): ms since the start of the step
      cudaEventRecord(fv->tr[3], stream);
      cudaStreamSynchronize(stream);
      float t[6] = {0, 0, 0, 0, 0, 0};
      cudaEventElapsedTime(&t[0], fv->tr[0], fv->tr[1]); cudaEventElapsedTime(&t[1], fv->tr[0], fv->plan->trace[0]);
      cudaEventElapsedTime(&t[2], fv->tr[0], fv->plan->trace[1]); cudaEventElapsedTime(&t[3], fv->tr[0], fv->plan->trace[2]);
      cudaEventElapsedTime(&t[4], fv->tr[0], fv->tr[2]); cudaEventElapsedTime(&t[5], fv->tr[0], fv->tr[3]);
      fprintf(stderr, "[dgb_fv trace rank %d] shell_end %.3f pack_end %.3f xfer_end %.3f unpack_end %.3f bulk_end %.3f step_end %.3f ms\n",
              fv->view.rank, t[0], t[1], t[2], t[3], t[4], t[5]);
      if (--fv->traceLeft == 0) for (int k = 0; k < 3; ++k) { cudaEventDestroy(fv->plan->trace[k]); fv->plan->trace[k] = nullptr; }
    }
    return DGB_OK;
  }
  if (int rc = launch_step<false>(fv, a, stream)) return rc;
  if (int rc = nccl_allreduce(fv->dSlots + ((n+2) % 4), 1, fv->comm, stream)) return rc;
  // refresh the overlap cells of c_out: communicate(InteriorBorder_All_Interface, ForwardCommunication), set
  if (fv->plan->nSendBoxes || fv->plan->nRecvBoxes) {
    const int32_t items[1] = {1};
    double* arrays[1] = { d_c_out };
    const long long before = fv->plan->launches;
    if (int rc = run_plan(fv->grid, *fv->plan, fv->view.rank, DGB_OP_SET, arrays, items, 1, fv->comm, stream)) return rc;
    fv->launches += fv->plan->launches - before;
    fv->exchangeMode = DGB_FV_EXCHANGE_SERIAL;
  }
  fv->step = n+1;
  return DGB_OK;
}

int dgb_fv_register_fields(dgb_fv* fv, double* const* d_fields, int nfields, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (fv->fused.timed_out()) return fail(DGB_ENCCL, kTimeoutMsg);
  if (int rc = ensure_fused(fv, stream)) return rc;
  if (!fv->fused.on) return DGB_OK;                            // NCCL forms: nothing to register, the steps work as before
  DGB_CUDA(cudaStreamSynchronize(stream));
  return fused_register(*fv->plan, fv->view.rank, fv->view.nranks, fv->comm, stream, fv->fused, d_fields, nfields);
}
int dgb_fv_fence(dgb_fv* fv, void* stream_) {
  if (fv->fused.timed_out()) return fail(DGB_ENCCL, kTimeoutMsg);
  return fv_flush_xhalo(fv, (cudaStream_t)stream_);
}
int dgb_fv_registered_fields(const dgb_fv* fv, int* n) { *n = int(fv->fused.direct.size()); return DGB_OK; }

int dgb_fv_bootstrap_local(dgb_fv* fv, double** d_pending_max, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DGB_CUDA(cudaMemsetAsync(fv->dSlots, 0, 5*sizeof(double), stream));
  FvArgs a = base_args(fv);
  a.slotOut = fv->dSlots + 0;
  if (int rc = launch_step<true>(fv, a, stream)) return rc;
  fv->bootstrapped = true; fv->step = 0; fv->seedPending = true;      // slot 1 = slot 0 once the caller has reduced slot 0
  if (d_pending_max) *d_pending_max = fv->dSlots + 0;
  return DGB_OK;
}

int dgb_fv_step_local(dgb_fv* fv, const double* d_c_in, double* d_c_out, double** d_pending_max, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!fv->bootstrapped) return fail(DGB_EARG, "dgb_fv_step_local: call dgb_fv_bootstrap_local (and reduce its result over ranks) first");
  if (int rc = fv_flush_xhalo(fv, stream)) return rc;
  const long long n = fv->step;
  if (fv->seedPending) {                                      // the caller has reduced slot 0 by now
    DGB_CUDA(cudaMemcpyAsync(fv->dSlots + 1, fv->dSlots + 0, sizeof(double), cudaMemcpyDeviceToDevice, stream));
    fv->seedPending = false;
  }
  FvArgs a = base_args(fv);
  a.c = d_c_in; a.cnew = d_c_out;
  slot_args(fv, n, a);
  if (fv->split && getenv("DGB_FV_SPLIT_LOCAL")) {            // the shell / bulk decomposition of dgb_fv_step, without exchange
    if (int rc = launch_shell(fv, a, false, stream)) return rc;
    FvArgs b = a; b.region = fv->bulk;
    if (int rc = launch_step<false>(fv, b, stream)) return rc;
  } else if (int rc = launch_step<false>(fv, a, stream)) return rc;
  if (d_pending_max) *d_pending_max = fv->dSlots + ((n+2) % 4);
  fv->step = n+1;
  return DGB_OK;
}

int dgb_fv_update(dgb_fv* fv, const double* d_c_in, double* d_update, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = fv_flush_xhalo(fv, stream)) return rc;
  if (!fv->bootstrapped) if (int rc = bootstrap(fv, stream)) return rc;
  // same kernel, update array materialised, nothing advanced: the reduction goes to a scratch slot that is re-zeroed
  const long long n = fv->step;
  FvArgs a = base_args(fv);
  a.c = d_c_in; a.update = d_update;
  a.slotIn = fv->dSlots + (n % 4); a.slotOut = fv->dSlots + ((n+3) % 4); a.dtOut = fv->dSlots + 4;
  if (int rc = launch_step<false>(fv, a, stream)) return rc;
  DGB_CUDA(cudaMemsetAsync(fv->dSlots + ((n+3) % 4), 0, sizeof(double), stream));
  return DGB_OK;
}

int dgb_fv_last_dt(dgb_fv* fv, double* dt_host, void* stream) {
  if (int rc = fv_flush_xhalo(fv, (cudaStream_t)stream)) return rc;     // this call synchronises: a natural point to complete the halo
  DGB_CUDA(cudaMemcpyAsync(dt_host, fv->dSlots + 4, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  unsigned timedOut = 0;
  if (fv->fused.on) DGB_CUDA(cudaMemcpyAsync(&timedOut, &fv->fused.ctl->timeout, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  DGB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (timedOut || fv->fused.timed_out()) return fail(DGB_ENCCL, kTimeoutMsg);
  return DGB_OK;
}

int dgb_fv_step_host(dgb_fv* fv, const double* h_c_in, double* h_c_out, double* dt_host, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (fv->fused.timed_out()) return fail(DGB_ENCCL, kTimeoutMsg);
  if (fv->comm) if (int rc = nccl_check_async(fv->comm)) return rc;
  if (int rc = fv_flush_xhalo(fv, stream)) return rc;
  int64_t n = 0; dgb_view_size(&fv->view, 0, &n);
  const size_t bytes = size_t(n)*sizeof(double);
  if (fv->stageBytes < bytes) {
    for (int i = 0; i < 2; ++i) { if (fv->dStage[i]) cudaFree(fv->dStage[i]); fv->dStage[i] = nullptr; }
    DGB_CUDA(cudaMalloc(&fv->dStage[0], bytes));
    DGB_CUDA(cudaMalloc(&fv->dStage[1], bytes));
    fv->stageBytes = bytes;
  }
  const dgb_view& v = fv->view;
  const dgb_box& of = v.comp[0].box[DGB_BOX_OVERLAPFRONT];
  const dgb_box& in = v.comp[0].box[DGB_BOX_INTERIOR];
  // Pipelined form (3-D): the field crosses PCIe in z-slabs of the STORED box; the interior cells of slab i are updated as soon
  // as slab i+1 (whose first plane they read) has arrived, and slab i of the result leaves while later slabs are still coming
  // in - the two copy directions run concurrently on their own streams around the compute stream. Ranks that exchange
  // overlap cells (several ranks, or periodic directions) use the fused exchange: every slab launch stores its share of
  // the send boxes into the receivers' buffers (peer memory) as it goes, the LAST launch raises the flag, and after the
  // last download the unpack kernel waits for all ranks' flags and scatters the received cells STRAIGHT INTO THE HOST
  // RESULT (mapped pinned memory) - the stale overlap cells the slab downloads carried are overwritten in place.
  static const int nslab = [] { const char* e = getenv("DGB_FV_HOST_SLABS"); const int v = e ? atoi(e) : 32; return v >= 2 && v <= 256 ? v : 32; }();
  const bool exchanges = fv->plan->nSendBoxes || fv->plan->nRecvBoxes;
  bool pipelined = v.dim == 3 && of.size[2] >= 4*nslab && in.size[0] > 0 && in.size[1] > 0 && in.size[2] > 0 && !getenv("DGB_FV_HOST_SERIAL");
  double* d_host_out = nullptr;                        // device alias of h_c_out (exchanging ranks unpack into it)
  if (pipelined && (exchanges || v.nranks > 1)) {
    if (int rc = ensure_fused(fv, stream)) return rc;
    cudaPointerAttributes attr;
    const bool mapped = cudaPointerGetAttributes(&attr, h_c_out) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer;
    if (!mapped) cudaGetLastError();
    // every rank must take the same path (the fused exchange is collective): a rank with pageable buffers vetoes it for all
    double ok = (fv->fused.on && mapped) ? 1.0 : 0.0;
    if (v.nranks > 1) {
      if (!fv->dHostVote) DGB_CUDA(cudaMalloc(&fv->dHostVote, sizeof(double)));
      DGB_CUDA(cudaMemcpyAsync(fv->dHostVote, &ok, sizeof(double), cudaMemcpyHostToDevice, stream));
      if (int rc = nccl_allreduce(fv->dHostVote, 0, fv->comm, stream)) return rc;
      DGB_CUDA(cudaMemcpyAsync(&ok, fv->dHostVote, sizeof(double), cudaMemcpyDeviceToHost, stream));
      DGB_CUDA(cudaStreamSynchronize(stream));
    }
    pipelined = ok > 0.5;
    if (pipelined) d_host_out = static_cast<double*>(attr.devicePointer);
  }
  fv->hostPath = pipelined ? DGB_FV_HOST_PIPELINED : DGB_FV_HOST_SERIAL;
  if (!pipelined) {
    DGB_CUDA(cudaMemcpyAsync(fv->dStage[0], h_c_in, bytes, cudaMemcpyHostToDevice, stream));
    // every stored cell of the result is written: interior cells by the kernel, overlap cells by the exchange
    if (int rc = dgb_fv_step(fv, fv->dStage[0], fv->dStage[1], stream)) return rc;
    DGB_CUDA(cudaMemcpyAsync(h_c_out, fv->dStage[1], bytes, cudaMemcpyDeviceToHost, stream));
    if (dt_host) DGB_CUDA(cudaMemcpyAsync(dt_host, fv->dSlots + 4, sizeof(double), cudaMemcpyDeviceToHost, stream));
    DGB_CUDA(cudaStreamSynchronize(stream));
    return DGB_OK;
  }
  if (!fv->sIn) {
    DGB_CUDA(cudaStreamCreateWithFlags(&fv->sIn, cudaStreamNonBlocking));
    DGB_CUDA(cudaStreamCreateWithFlags(&fv->sOut, cudaStreamNonBlocking));
    fv->evIn.resize(nslab); fv->evK.resize(nslab);
    for (int i = 0; i < nslab; ++i) {
      DGB_CUDA(cudaEventCreateWithFlags(&fv->evIn[i], cudaEventDisableTiming));
      DGB_CUDA(cudaEventCreateWithFlags(&fv->evK[i], cudaEventDisableTiming));
    }
    DGB_CUDA(cudaEventCreateWithFlags(&fv->evStart, cudaEventDisableTiming));
    DGB_CUDA(cudaEventCreateWithFlags(&fv->evEnd, cudaEventDisableTiming));
  }
  if (!fv->bootstrapped) if (int rc = bootstrap(fv, stream)) return rc;
  DGB_CUDA(cudaEventRecord(fv->evStart, stream));
  DGB_CUDA(cudaStreamWaitEvent(fv->sIn, fv->evStart, 0));
  DGB_CUDA(cudaStreamWaitEvent(fv->sOut, fv->evStart, 0));
  const size_t plane = size_t(of.size[0])*of.size[1];
  const int nz = of.size[2];
  auto zbegin = [&](int i) { return int((long long)nz*i/nslab); };            // first stored plane of slab i
  for (int i = 0; i < nslab; ++i) {
    const size_t o = plane*zbegin(i), cnt = plane*(zbegin(i+1)-zbegin(i));
    DGB_CUDA(cudaMemcpyAsync(fv->dStage[0] + o, h_c_in + o, cnt*sizeof(double), cudaMemcpyHostToDevice, fv->sIn));
    DGB_CUDA(cudaEventRecord(fv->evIn[i], fv->sIn));
  }
  const long long sn = fv->step;
  const bool fused = fv->fused.on;
  FusedState& fs = fv->fused;
  const int parity = int(fs.epoch & 1u);
  // interior planes of slab i (stored plane p holds the cells with z = of.origin[2] + p)
  auto zlo = [&](int i) { return std::max(of.origin[2] + zbegin(i), in.origin[2]); };
  auto zhi = [&](int i) { return std::min(of.origin[2] + zbegin(i+1), in.origin[2] + in.size[2]); };
  int lastLaunch = -1;
  for (int i = 0; i < nslab; ++i) if (zhi(i) > zlo(i)) lastLaunch = i;
  for (int i = 0; i < nslab; ++i) {
    if (zhi(i) > zlo(i)) {
      DGB_CUDA(cudaStreamWaitEvent(stream, fv->evIn[std::min(i+1, nslab-1)], 0));
      FvArgs a = base_args(fv);
      a.region.origin[2] = zlo(i); a.region.size[2] = zhi(i)-zlo(i);
      a.c = fv->dStage[0]; a.cnew = fv->dStage[1];
      slot_args(fv, sn, a);
      if (fused) { a.pack = fs.dPack + parity; a.packEpoch = fs.epoch + 1u; a.packSignal = (i == lastLaunch) ? 1 : 0;
                   fill_send_boxes(a, *fv->plan, fs.hostDst[parity], nullptr, nullptr); }
      if (int rc = launch_step<false>(fv, a, stream)) return rc;
    }
    DGB_CUDA(cudaEventRecord(fv->evK[i], stream));
    DGB_CUDA(cudaStreamWaitEvent(fv->sOut, fv->evK[i], 0));
    const size_t o = plane*zbegin(i), cnt = plane*(zbegin(i+1)-zbegin(i));
    DGB_CUDA(cudaMemcpyAsync(h_c_out + o, fv->dStage[1] + o, cnt*sizeof(double), cudaMemcpyDeviceToHost, fv->sOut));
  }
  if (fused) {
    // after the last download: wait for every rank's flag (CTA 0 also reduces the CFL maxima into the slot of step n+2), then
    // scatter the received overlap cells into the host result
    const int32_t items[1] = {1};
    double* arrays[1] = { d_host_out };
    const long long before = fv->plan->launches;
    const HaloWait hw{fs.ctl, v.nranks, fs.epoch + 1u, parity, fv->dSlots + ((sn+2) % 4), fs.dTimeout};
    if (fv->plan->nRecvBoxes == 0) { k_fv_wait<<<1, 64, 0, fv->sOut>>>(fs.ctl, v.nranks, fs.epoch + 1u, parity, fv->dSlots + ((sn+2) % 4), fs.dTimeout); ++fv->launches; }
    if (int rc = plan_unpack_from(*fv->plan, fs.recv[parity], DGB_OP_SET, arrays, items, 1, fv->sOut, &hw)) return rc;
    fv->launches += fv->plan->launches - before;
    int64_t sent = 0;
    for (auto& sg : fv->plan->sendSeg) if (sg.peer != v.rank) sent += sg.count*8;
    fv->grid->lastBytesSent = sent;
    ++fs.epoch;
    fv->exchangeMode = DGB_FV_EXCHANGE_FUSED;
  } else if (int rc = nccl_allreduce(fv->dSlots + ((sn+2) % 4), 1, fv->comm, stream)) return rc;
  fv->step = sn+1;
  DGB_CUDA(cudaEventRecord(fv->evEnd, fv->sOut));
  DGB_CUDA(cudaStreamWaitEvent(stream, fv->evEnd, 0));
  if (dt_host) DGB_CUDA(cudaMemcpyAsync(dt_host, fv->dSlots + 4, sizeof(double), cudaMemcpyDeviceToHost, stream));
  DGB_CUDA(cudaStreamSynchronize(stream));
  if (fv->fused.timed_out()) return fail(DGB_ENCCL, kTimeoutMsg);
  return DGB_OK;
}

int dgb_fv_host_path(const dgb_fv* fv, int* path) { *path = fv->hostPath; return DGB_OK; }
int dgb_fv_kernel_info(const dgb_fv* fv, int32_t* info8) { for (int k = 0; k < 8; ++k) info8[k] = fv->kinfo[k]; return DGB_OK; }
int dgb_fv_chunk_rule(int nx, int ny, int nz, int packs, int tensor_map, int* zchunk) {
  if (nx <= 0 || ny <= 0 || nz <= 0 || !zchunk) return fail(DGB_EARG, "dgb_fv_chunk_rule: positive extents and a result pointer");
  const int size[3] = { nx, ny, nz };
  *zchunk = (tensor_map && !packs) ? tma_zchunk(size, 128, 4) : auto_zchunk(size, 128, 4, packs != 0);
  return DGB_OK;
}
int dgb_fv_launch_count(const dgb_fv* fv, int64_t* kernels) { *kernels = fv->launches; return DGB_OK; }
int dgb_fv_exchange_mode(const dgb_fv* fv, int* mode) { *mode = fv->exchangeMode; return DGB_OK; }

} // extern "C"
