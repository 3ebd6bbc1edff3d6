// dgb_functors.cuh — the functor boundary of the bulk layer (SURVEY.md §8b "functor kernels"): per-element and
// per-intersection functors are compiled-in DEVICE structs, selected by id through a registry; the library sweeps the
// grid view and calls them (no host callbacks, no CPU fallback). This header is what a maintainer edits to add one.
//
//   sweep functor   a struct with   static constexpr int id;  static const char* name();
//                                   static constexpr bool elements, faces;   (which loop bodies it has)
//                   and             template<int DIM> __device__ static double element(const ElementCtx<DIM>&, const double* args);
//                                   template<int DIM> __device__ static double face(const FaceCtx<DIM>&, const double* args);
//                   The returned value is the functor's contribution to the sweep's reduction (and, optionally, is stored per
//                   element / per face). ElementCtx / FaceCtx offer what the reference's loop body asks the entity,
//                   geometry and intersection facades for (dune/grid/common/entity.hh, geometry.hh, intersection.hh).
//   FV problem      a struct with   static constexpr int id;  static constexpr bool column_invariant;
//                                   __device__ static void velocity(const double* p, const double* x, double* u);   (3 components)
//                                   __device__ static double boundary(const double* p, const double* x);
//                                   __device__ static double initial(const double* p, int dim, const double* x);
//                   column_invariant = the velocity does not depend on z (and, as all problems, not on time): only then the
//                   production kernel k_fv_ring, which evaluates the x/y face factors once per column, may be used; other
//                   problems run on the general kernel k_fv_march (dgb_fv.cu).
//
// The three sweep functors shipped here are the loop bodies of the reference's doc/recipes/recipe-integration.cc:90-126.
#ifndef DGB_FUNCTORS_CUH
#define DGB_FUNCTORS_CUH
#include <cuda_runtime.h>
#include "dgb_device.cuh"

namespace dgb {

// ---- what a functor sees ---------------------------------------------------------------------------------------------
template<int DIM>
struct ElementCtx {
  const dgb_view* v;
  int coord[3];
  uint32_t idx;                  // indexSet.index(e)
  double ll[DIM], ur[DIM];       // e.geometry(): lower / upper corner (periodic wrap applied, yaspgridentity.hh:497-510)
  __device__ __forceinline__ uint32_t index() const { return idx; }
  __device__ __forceinline__ int partitionType() const { return partition_type(*v, v->comp[0], coord); }
  __device__ __forceinline__ double volume() const { double s = 1.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) s *= ur[i]-ll[i]; return s; }
  __device__ __forceinline__ double integrationElement(const double*) const { return volume(); }
  __device__ __forceinline__ void center(double* x) const {
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = 0.5*(ll[i]+ur[i]); }
  __device__ __forceinline__ void global(const double* xi, double* x) const {        // AxisAlignedCubeGeometry::global
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = ll[i] + xi[i]*(ur[i]-ll[i]); }
};

template<int DIM>
struct FaceCtx {
  const ElementCtx<DIM>* e;      // inside()
  int k;                         // indexInInside(): dir = k/2, side = k%2
  __device__ __forceinline__ int dir() const { return k >> 1; }
  __device__ __forceinline__ int side() const { return k & 1; }
  __device__ __forceinline__ bool neighbor() const { return face_neighbor(*e->v, e->coord, dir(), side()); }
  __device__ __forceinline__ bool boundary() const { return face_boundary(*e->v, e->coord, dir(), side()); }
  __device__ __forceinline__ int boundarySegmentIndex() const { return boundary_segment_index(*e->v, e->coord, dir(), side()); }
  __device__ __forceinline__ double centerUnitOuterNormal(int i) const { return i == dir() ? (side() ? 1.0 : -1.0) : 0.0; }
  __device__ __forceinline__ void center(double* x) const {                           // geometry().center() of the face
#pragma unroll
    for (int i = 0; i < DIM; ++i) { const double lo = (i == dir()) ? (side() ? e->ur[i] : e->ll[i]) : e->ll[i], hi = (i == dir()) ? lo : e->ur[i]; x[i] = 0.5*(lo+hi); } }
  __device__ __forceinline__ double volume() const { double s = 1.0;                  // geometry().volume() of the face
#pragma unroll
    for (int i = 0; i < DIM; ++i) if (i != dir()) s *= e->ur[i]-e->ll[i]; return s; }
};

template<int DIM> __device__ __forceinline__ double two_norm(const double* x) { double s = 0.0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) s += x[i]*x[i]; return sqrt(s); }

// ---- the recipe functors (doc/recipes/recipe-integration.cc) ---------------------------------------------------------
// :90-98  integral += u(e.geometry().center())*e.geometry().volume(),  u(x) = exp(|x|)
struct MidpointExpNorm {
  static constexpr int id = 0; static constexpr bool elements = true, faces = false;
  static const char* name() { return "recipe-integration: midpoint rule of u(x) = exp(|x|)"; }
  template<int DIM> __device__ static double element(const ElementCtx<DIM>& e, const double*) { double x[DIM]; e.center(x); return exp(two_norm<DIM>(x))*e.volume(); }
  template<int DIM> __device__ static double face(const FaceCtx<DIM>&, const double*) { return 0.0; }
};
// :100-111  quadrature rule of order 5 (Gauss-Legendre, 3 points per direction, first coordinate fastest):
//           integral2 += u(geo.global(qp.position()))*geo.integrationElement(qp.position())*qp.weight()
struct Gauss5ExpNorm {
  static constexpr int id = 1; static constexpr bool elements = true, faces = false;
  static const char* name() { return "recipe-integration: Gauss rule of order 5 of u(x) = exp(|x|)"; }
  template<int DIM> __device__ static double element(const ElementCtx<DIM>& e, const double*) {
    const double gx[3] = { 0.5 - 0.5*0.7745966692414834, 0.5, 0.5 + 0.5*0.7745966692414834 };      // (1 -+ sqrt(3/5))/2
    const double gw[3] = { 5.0/18.0, 8.0/18.0, 5.0/18.0 };
    constexpr int NQ = DIM == 3 ? 27 : 9;
    double s = 0.0;
    for (int q = 0; q < NQ; ++q) {
      double xi[DIM], x[DIM], w = 1.0; int r = q;
#pragma unroll
      for (int i = 0; i < DIM; ++i) { xi[i] = gx[r % 3]; w *= gw[r % 3]; r /= 3; }
      e.global(xi, x);
      s += exp(two_norm<DIM>(x))*e.integrationElement(xi)*w;
    }
    return s;
  }
  template<int DIM> __device__ static double face(const FaceCtx<DIM>&, const double*) { return 0.0; }
};
// :113-124  for intersections without neighbour: divergence += f(geoI.center())*I.centerUnitOuterNormal()*geoI.volume(), f(x) = x
struct BoundaryFluxIdentity {
  static constexpr int id = 2; static constexpr bool elements = false, faces = true;
  static const char* name() { return "recipe-integration: boundary flux of f(x) = x"; }
  template<int DIM> __device__ static double element(const ElementCtx<DIM>&, const double*) { return 0.0; }
  template<int DIM> __device__ static double face(const FaceCtx<DIM>& f, const double*) {
    if (f.neighbor()) return 0.0;
    double x[DIM], s = 0.0; f.center(x);
#pragma unroll
    for (int i = 0; i < DIM; ++i) s += x[i]*f.centerUnitOuterNormal(i);
    return s*f.volume();
  }
};
// element volume / face area, optionally scaled by args[0]: the simplest "user" functors (sum = |domain| / total face area)
struct Measure {
  static constexpr int id = 3; static constexpr bool elements = true, faces = true;
  static const char* name() { return "measure: element volume (element loop) / face volume (intersection loop)"; }
  template<int DIM> __device__ static double element(const ElementCtx<DIM>& e, const double* a) { return (a ? a[0] : 1.0)*e.volume(); }
  template<int DIM> __device__ static double face(const FaceCtx<DIM>& f, const double* a) { return (a ? a[0] : 1.0)*f.volume(); }
};

// the registry: DGB_FOR_EACH_SWEEP_FUNCTOR(X) expands X(Type) for every compiled-in sweep functor
#define DGB_FOR_EACH_SWEEP_FUNCTOR(X) X(MidpointExpNorm) X(Gauss5ExpNorm) X(BoundaryFluxIdentity) X(Measure)

// ---- FV transport problems (oracle/common/fv_problem.hh is the CPU statement of the same functions) ---------
// p = the params array of dgb_fv_create: p[3] inflow value, p[4..6] shell centre, p[7..8] shell radii
struct FvInitialShell {
  __device__ __forceinline__ static double eval(const double* p, int dim, const double* x) {
    double s = 0.0;
    for (int i = 0; i < dim; ++i) { const double d = x[i]-p[4+i]; s += d*d; }
    const double r = sqrt(s);
    return (r > p[7] && r < p[8]) ? 1.0 : 0.0;
  }
};
// problem 0: constant velocity p[0..2]
struct FvConstantVelocity {
  static constexpr int id = 0; static constexpr bool column_invariant = true;
  __device__ __forceinline__ static void velocity(const double* p, const double*, double* u) { u[0] = p[0]; u[1] = p[1]; u[2] = p[2]; }
  __device__ __forceinline__ static double boundary(const double* p, const double*) { return p[3]; }
  __device__ __forceinline__ static double initial(const double* p, int dim, const double* x) { return FvInitialShell::eval(p, dim, x); }
};
// problem 1: rigid rotation in the x-y plane about (0.5, 0.5) with angular speed p[0]
struct FvRotation {
  static constexpr int id = 1; static constexpr bool column_invariant = true;
  __device__ __forceinline__ static void velocity(const double* p, const double* x, double* u) { u[0] = -p[0]*(x[1]-0.5); u[1] = p[0]*(x[0]-0.5); u[2] = 0.0; }
  __device__ __forceinline__ static double boundary(const double* p, const double*) { return p[3]; }
  __device__ __forceinline__ static double initial(const double* p, int dim, const double* x) { return FvInitialShell::eval(p, dim, x); }
};
// problem 2: shear flow whose x velocity depends on z, u = (p[0]*(x_z - 0.5), p[1], p[2]*(x_x - 0.5)) (3-D; in 2-D the z
// coordinate is 0): NOT column invariant - the step runs on the general kernel
struct FvShear {
  static constexpr int id = 2; static constexpr bool column_invariant = false;
  __device__ __forceinline__ static void velocity(const double* p, const double* x, double* u) { u[0] = p[0]*(x[2]-0.5); u[1] = p[1]; u[2] = p[2]*(x[0]-0.5); }
  __device__ __forceinline__ static double boundary(const double* p, const double*) { return p[3]; }
  __device__ __forceinline__ static double initial(const double* p, int dim, const double* x) { return FvInitialShell::eval(p, dim, x); }
};
#define DGB_FOR_EACH_FV_PROBLEM(X) X(FvConstantVelocity) X(FvRotation) X(FvShear)
static constexpr int DGB_FV_PROBLEMS = 3;

} // namespace dgb
#endif
