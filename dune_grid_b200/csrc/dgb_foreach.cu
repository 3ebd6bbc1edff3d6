// dgb_foreach.cu — dgb_for_each_element / dgb_for_each_intersection: the generic functor sweeps of the bulk layer
// (SURVEY.md §8b "functor kernels"; north_star: "per-element and per-intersection functors run as batched device kernels
// behind a thin C-ABI"). The loop `for (e : elements(gv, partition)) [for (is : intersections(gv, e))] body(e, is)` of the
// reference (rangegenerators.hh:845,881) becomes ONE fused kernel: sweep + functor + reduction; nothing is materialised
// unless the caller asks for the per-entity values.
//
// Reduction order (deterministic, independent of the launch configuration's timing): a CTA owns a contiguous run of
// SWEEP_CHUNK*SWEEP_THREADS cells in iteration order; thread t adds the values of its cells (t, t+T, t+2T, ...; faces of a
// cell in reference order) in ascending order; the CTA sums its threads pairwise (tree over the thread index); the CTA
// partials are summed by a second one-CTA kernel, again thread-strided and then by the same tree. Same input -> same
// bits, on every run.
#include "dgb_internal.hh"
#include "dgb_device.cuh"
#include "dgb_functors.cuh"

namespace dgb {

static constexpr int SWEEP_THREADS = 256, SWEEP_CHUNK = 8;     // cells per CTA = 2048

struct FunctorArgs { int n; double a[16]; };

template<int T>
__device__ __forceinline__ double block_tree_sum(double v, double* sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
#pragma unroll
  for (int s = T/2; s > 0; s >>= 1) {
    if (t < s) sh[t] += sh[t+s];
    __syncthreads();
  }
  return sh[0];
}

// cell number `e` of the partition box (x fastest) -> context
template<int DIM>
__device__ __forceinline__ void make_element(const dgb_view& v, int kind, long long e, ElementCtx<DIM>& c) {
  const dgb_component& comp = v.comp[0];
  const dgb_box& b = comp.box[kind];
  const int s0 = b.size[0], s1 = b.size[1];
  const long long q = e / s0;
  const int i0 = int(e - q*s0);
  const int i2 = DIM > 2 ? int(q / s1) : 0;
  const int i1 = int(q - (long long)i2*s1);
  c.v = &v;
  c.coord[0] = b.origin[0]+i0; c.coord[1] = b.origin[1]+i1; c.coord[2] = DIM > 2 ? b.origin[2]+i2 : 0;
  c.idx = entity_index(v, comp, kind, c.coord);
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const double sh = periodic_shift(v, i, c.coord[i]);
    double lo = vertex_coord(v, i, c.coord[i]), hi = vertex_coord(v, i, c.coord[i]+1);
    if (sh != 0.0) { lo += sh; hi += sh; }
    c.ll[i] = lo; c.ur[i] = hi;
  }
}

// FACES = false: element loop; true: intersection loop (per_entity then holds 2*DIM values per cell, face-major: [k*n + cell])
template<class F, int DIM, bool FACES>
__global__ void __launch_bounds__(SWEEP_THREADS) k_for_each(const __grid_constant__ dgb_view v, int kind, long long n, const FunctorArgs fa,
                                                            double* __restrict__ per_entity, double* __restrict__ partials) {
  __shared__ double sh[SWEEP_THREADS];
  const long long first = (long long)blockIdx.x*SWEEP_THREADS*SWEEP_CHUNK;
  double acc = 0.0;
  const double* args = fa.n ? fa.a : nullptr;
  for (int j = 0; j < SWEEP_CHUNK; ++j) {
    const long long e = first + (long long)j*SWEEP_THREADS + threadIdx.x;
    if (e < n) {
      ElementCtx<DIM> c;
      make_element<DIM>(v, kind, e, c);
      if (!FACES) {
        const double val = F::template element<DIM>(c, args);
        if (per_entity) per_entity[e] = val;
        acc += val;
      } else {
#pragma unroll
        for (int k = 0; k < 2*DIM; ++k) {
          FaceCtx<DIM> f; f.e = &c; f.k = k;
          const double val = F::template face<DIM>(f, args);
          if (per_entity) per_entity[(long long)k*n + e] = val;
          acc += val;
        }
      }
    }
  }
  const double total = block_tree_sum<SWEEP_THREADS>(acc, sh);
  if (threadIdx.x == 0 && partials) partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SWEEP_THREADS) k_sum_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double sh[SWEEP_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += SWEEP_THREADS) acc += partials[i];
  const double total = block_tree_sum<SWEEP_THREADS>(acc, sh);
  if (threadIdx.x == 0) *out = total;
}

template<class F, bool FACES>
static int launch_for_each(const dgb_view& v, int kind, long long n, const FunctorArgs& fa, double* per_entity, double* partials, int nblocks,
                           cudaStream_t stream) {
  if (v.dim == 3) k_for_each<F, 3, FACES><<<nblocks, SWEEP_THREADS, 0, stream>>>(v, kind, n, fa, per_entity, partials);
  else k_for_each<F, 2, FACES><<<nblocks, SWEEP_THREADS, 0, stream>>>(v, kind, n, fa, per_entity, partials);
  return DGB_OK;
}

static int for_each(const dgb_view* v, int pitype, int functor_id, const double* args_host, int nargs, double* d_per_entity, double* d_sum,
                    bool faces, cudaStream_t stream) {
  const int kind = box_kind_of(pitype);
  if (kind == -2) return fail(DGB_EGRID, "YaspLevelIterator with this codim or partition type not implemented");
  if (nargs < 0 || nargs > 16) return fail(DGB_EARG, "at most 16 functor arguments");
  if (int rc = require_device()) return rc;
  if (v->coord_mode == DGB_COORD_TENSORPRODUCT) for (int i = 0; i < v->dim; ++i) if (!v->tensor[i]) return fail(DGB_ECUDA, "tensor coordinates are not on the device");
  FunctorArgs fa; fa.n = nargs;
  for (int i = 0; i < 16; ++i) fa.a[i] = (args_host && i < nargs) ? args_host[i] : 0.0;
  const long long n = kind == -1 ? 0 : partition_size(*v, 0, pitype);       // Ghost_Partition: empty range
  const int nblocks = int((n + SWEEP_THREADS*SWEEP_CHUNK - 1)/(SWEEP_THREADS*SWEEP_CHUNK));
  double* partials = nullptr;
  if (d_sum && nblocks > 0) DGB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&partials), size_t(nblocks)*sizeof(double), stream));
  bool found = false; int rc = DGB_OK;
  if (nblocks > 0) {
#define DGB_TRY(F_)                                                                                                  \
    if (!found && functor_id == F_::id) {                                                                            \
      found = true;                                                                                                  \
      if (faces ? !F_::faces : !F_::elements) rc = fail(DGB_EARG, std::string("functor '") + F_::name() + "' has no " + (faces ? "intersection" : "element") + " loop body"); \
      else rc = faces ? launch_for_each<F_, true>(*v, kind, n, fa, d_per_entity, partials, nblocks, stream)          \
                      : launch_for_each<F_, false>(*v, kind, n, fa, d_per_entity, partials, nblocks, stream);        \
    }
    DGB_FOR_EACH_SWEEP_FUNCTOR(DGB_TRY)
#undef DGB_TRY
    if (!found) rc = fail(DGB_EARG, "unknown functor id " + std::to_string(functor_id));
  } else {
#define DGB_KNOWN(F_) if (functor_id == F_::id) found = true;
    DGB_FOR_EACH_SWEEP_FUNCTOR(DGB_KNOWN)
#undef DGB_KNOWN
    if (!found) rc = fail(DGB_EARG, "unknown functor id " + std::to_string(functor_id));
  }
  if (rc == DGB_OK && d_sum) {
    if (nblocks > 0) k_sum_partials<<<1, SWEEP_THREADS, 0, stream>>>(partials, nblocks, d_sum);
    else DGB_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), stream));
  }
  if (partials) cudaFreeAsync(partials, stream);
  if (rc != DGB_OK) return rc;
  DGB_CUDA(cudaGetLastError());
  return DGB_OK;
}

} // namespace dgb

using namespace dgb;

extern "C" {

int dgb_functor_count(int* n) {
  int c = 0;
#define DGB_COUNT(F_) ++c;
  DGB_FOR_EACH_SWEEP_FUNCTOR(DGB_COUNT)
#undef DGB_COUNT
  *n = c;
  return DGB_OK;
}

int dgb_functor_info(int functor_id, const char** name, int* has_element_body, int* has_intersection_body) {
#define DGB_INFO(F_) if (functor_id == F_::id) { if (name) *name = F_::name(); if (has_element_body) *has_element_body = F_::elements; \
                                                 if (has_intersection_body) *has_intersection_body = F_::faces; return DGB_OK; }
  DGB_FOR_EACH_SWEEP_FUNCTOR(DGB_INFO)
#undef DGB_INFO
  return fail(DGB_EARG, "unknown functor id " + std::to_string(functor_id));
}

int dgb_for_each_element(const dgb_view* v, int pitype, int functor_id, const double* args_host, int nargs, double* d_per_element,
                         double* d_sum, void* stream) {
  return guarded([&] { return for_each(v, pitype, functor_id, args_host, nargs, d_per_element, d_sum, false, (cudaStream_t)stream); });
}

int dgb_for_each_intersection(const dgb_view* v, int pitype, int functor_id, const double* args_host, int nargs, double* d_per_face,
                              double* d_sum, void* stream) {
  return guarded([&] { return for_each(v, pitype, functor_id, args_host, nargs, d_per_face, d_sum, true, (cudaStream_t)stream); });
}

} // extern "C"
