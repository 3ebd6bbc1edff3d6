// dgb_geogrid.cuh — device evaluation of GeometryGrid geometries on cubes: corners are the analytic coordinate
// function applied to the host (Yasp) cell corners (geometrygrid/cornerstorage.hh:54-60, coordfunctioncaller.hh:40-43),
// the map is dune-geometry's MultiLinearGeometry (geometrygrid/geometry.hh:66,197-204). The interpolation follows
// MultiLinearGeometry's recursion order (weights multiplied from the highest axis down, corners accumulated in
// ascending order, last Jacobian row = top face minus bottom face) so results equal the CPU restatement
// (oracle/common/multilinear.hh) to the last bit for global/jt/detJ; J^-T uses the closed-form adjugate instead of the
// Cholesky-based right inverse of FieldMatrixHelper, which agrees to O(cond(J)^2 eps) (tolerance stated in the tests).
#ifndef DGB_GEOGRID_CUH
#define DGB_GEOGRID_CUH
#include <cuda_runtime.h>

namespace dgb {

struct Deformation { int id; double p[4]; };

// analytic coordinate functions (compiled-in device functors selected by id)
template<int DIM>
__device__ __forceinline__ void deform(const Deformation& f, const double* x, double* y) {
  if (f.id == 0) {
#pragma unroll
    for (int i = 0; i < DIM; ++i) y[i] = x[i];
    return;
  }
  const double a = f.p[0];
  const double twopi = 6.283185307179586476925286766559;
  if (DIM == 3) {
    const double s0 = sin(twopi*x[0]), s1 = sin(twopi*x[1]), s2 = sin(twopi*x[2]);
    y[0] = x[0] + a*(s1*s2);
    y[1] = x[1] + a*(s2*s0);
    y[2] = x[2] + a*(s0*s1);
  } else {
    const double s0 = sin(twopi*x[0]), s1 = sin(twopi*x[1]);
    y[0] = x[0] + a*s1;
    y[1] = x[1] + a*s0;
  }
}

// y (+)= rf * interpolation over the 2^D corners starting at corner FIRST, axes D-1..0
template<int DIM, int D, int FIRST, bool ADD>
struct MLGlobal {
  __device__ __forceinline__ static void run(const double (*c)[DIM], const double* x, double rf, double* y) {
    const double xn = x[D-1];
    const double cxn = 1.0 - xn;
    MLGlobal<DIM, D-1, FIRST, ADD>::run(c, x, rf*cxn, y);
    MLGlobal<DIM, D-1, FIRST + (1 << (D-1)), true>::run(c, x, rf*xn, y);
  }
};
template<int DIM, int FIRST, bool ADD>
struct MLGlobal<DIM, 0, FIRST, ADD> {
  __device__ __forceinline__ static void run(const double (*c)[DIM], const double*, double rf, double* y) {
#pragma unroll
    for (int i = 0; i < DIM; ++i) y[i] = ADD ? y[i] + rf*c[FIRST][i] : rf*c[FIRST][i];
  }
};
template<int DIM, int D, int FIRST, bool ADD>
struct MLJt {
  __device__ __forceinline__ static void run(const double (*c)[DIM], const double* x, double rf, double (*jt)[DIM]) {
    const double xn = x[D-1];
    const double cxn = 1.0 - xn;
    MLJt<DIM, D-1, FIRST, ADD>::run(c, x, rf*cxn, jt);
    MLJt<DIM, D-1, FIRST + (1 << (D-1)), true>::run(c, x, rf*xn, jt);
    MLGlobal<DIM, D-1, FIRST, ADD>::run(c, x, -rf, jt[D-1]);
    MLGlobal<DIM, D-1, FIRST + (1 << (D-1)), true>::run(c, x, rf, jt[D-1]);
  }
};
template<int DIM, int FIRST, bool ADD>
struct MLJt<DIM, 0, FIRST, ADD> {
  __device__ __forceinline__ static void run(const double (*)[DIM], const double*, double, double (*)[DIM]) {}
};

template<int DIM> __device__ __forceinline__ void ml_global(const double (*c)[DIM], const double* x, double* y) {
  MLGlobal<DIM, DIM, 0, false>::run(c, x, 1.0, y);
}
template<int DIM> __device__ __forceinline__ void ml_jt(const double (*c)[DIM], const double* x, double (*jt)[DIM]) {
  MLJt<DIM, DIM, 0, false>::run(c, x, 1.0, jt);
}

// signed determinant with the operation order of FieldMatrixHelper::sqrtDetAAT's explicit 2x2 / 3x3 formulas
template<int DIM> __device__ __forceinline__ double ml_det(const double (*A)[DIM]);
template<> __device__ __forceinline__ double ml_det<2>(const double (*A)[2]) {
  return A[0][0]*A[1][1] - A[1][0]*A[0][1];
}
template<> __device__ __forceinline__ double ml_det<3>(const double (*A)[3]) {
  const double v0 = A[0][1]*A[1][2] - A[1][1]*A[0][2];
  const double v1 = A[0][2]*A[1][0] - A[1][2]*A[0][0];
  const double v2 = A[0][0]*A[1][1] - A[1][0]*A[0][1];
  return v0*A[2][0] + v1*A[2][1] + v2*A[2][2];
}
// ret = A^-1 (right inverse of the transposed Jacobian = jacobianInverseTransposed)
template<int DIM> __device__ __forceinline__ void ml_inverse_transposed(const double (*A)[DIM], double det, double (*ret)[DIM]);
template<> __device__ __forceinline__ void ml_inverse_transposed<2>(const double (*A)[2], double det, double (*ret)[2]) {
  const double detInv = 1.0/det;
  ret[0][0] =  A[1][1]*detInv;
  ret[1][1] =  A[0][0]*detInv;
  ret[1][0] = -A[1][0]*detInv;
  ret[0][1] = -A[0][1]*detInv;
}
template<> __device__ __forceinline__ void ml_inverse_transposed<3>(const double (*A)[3], double det, double (*ret)[3]) {
  const double detInv = 1.0/det;
  ret[0][0] = (A[1][1]*A[2][2] - A[1][2]*A[2][1])*detInv;
  ret[0][1] = (A[0][2]*A[2][1] - A[0][1]*A[2][2])*detInv;
  ret[0][2] = (A[0][1]*A[1][2] - A[0][2]*A[1][1])*detInv;
  ret[1][0] = (A[1][2]*A[2][0] - A[1][0]*A[2][2])*detInv;
  ret[1][1] = (A[0][0]*A[2][2] - A[0][2]*A[2][0])*detInv;
  ret[1][2] = (A[0][2]*A[1][0] - A[0][0]*A[1][2])*detInv;
  ret[2][0] = (A[1][0]*A[2][1] - A[1][1]*A[2][0])*detInv;
  ret[2][1] = (A[0][1]*A[2][0] - A[0][0]*A[2][1])*detInv;
  ret[2][2] = (A[0][0]*A[1][1] - A[0][1]*A[1][0])*detInv;
}

} // namespace dgb
#endif
