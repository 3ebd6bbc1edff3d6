// multilinear.hh — CPU restatement of dune-geometry's MultiLinearGeometry on cubes and of
// Impl::FieldMatrixHelper (rightInvA / sqrtDetAAT), which GeometryGrid uses for every geometry
// (reference dune/grid/geometrygrid/geometry.hh:66,105,197-204).
// dune-geometry is an un-vendored dependency of the reference (dune.module:10, pinned only as
// ">= 2.12"), so this follows the published algorithm of multilineargeometry.hh / affinegeometry.hh
// as recalled in SURVEY.md Appendix C.2: GEOMETRY PARITY UNPINNED (no upstream binary or golden
// vector exists to check it against). Test infrastructure only.
#ifndef ORACLE_COMMON_MULTILINEAR_HH
#define ORACLE_COMMON_MULTILINEAR_HH
#include <cmath>
namespace orc {

  // y (+)= rf * multilinear interpolation of the 2^d corners starting at *cit over axes d-1..0;
  // corners are consumed in ascending order (recursion of MultiLinearGeometry::global).
  template<int cdim>
  inline void ml_global(int d, const double (*&cit)[cdim], const double* x, double rf, bool add, double* y) {
    if (d == 0) {
      const double* p = *cit; ++cit;
      for (int i = 0; i < cdim; ++i) y[i] = add ? y[i] + rf*p[i] : rf*p[i];
      return;
    }
    const double xn = 1.0*x[d-1];
    const double cxn = 1.0 - xn;
    ml_global<cdim>(d-1, cit, x, rf*cxn, add, y);
    ml_global<cdim>(d-1, cit, x, rf*xn, true, y);
  }

  // jt rows 0..d-1 (+)= rf * Jacobian transposed of the sub-cube starting at *cit
  template<int cdim>
  inline void ml_jt(int d, const double (*&cit)[cdim], const double* x, double rf, bool add, double (*jt)[cdim]) {
    if (d == 0) { ++cit; return; }
    const double xn = 1.0*x[d-1];
    const double cxn = 1.0 - xn;
    const double (*cit2)[cdim] = cit;
    ml_jt<cdim>(d-1, cit2, x, rf*cxn, add, jt);
    ml_jt<cdim>(d-1, cit2, x, rf*xn, true, jt);
    // last row: top-face interpolation minus bottom-face interpolation
    ml_global<cdim>(d-1, cit, x, -rf, add, jt[d-1]);
    ml_global<cdim>(d-1, cit, x, rf, true, jt[d-1]);
  }

  template<int dim>
  inline void multilinear_global(const double (*corners)[dim], const double* x, double* y) {
    const double (*cit)[dim] = corners;
    ml_global<dim>(dim, cit, x, 1.0, false, y);
  }
  template<int dim>
  inline void multilinear_jt(const double (*corners)[dim], const double* x, double (*jt)[dim]) {
    const double (*cit)[dim] = corners;
    ml_jt<dim>(dim, cit, x, 1.0, false, jt);
  }

  // FieldMatrixHelper::sqrtDetAAT for square 2x2 / 3x3 (explicit formulas)
  template<int dim> inline double sqrt_det_aat(const double (*A)[dim]);
  template<> inline double sqrt_det_aat<2>(const double (*A)[2]) {
    return std::abs(A[0][0]*A[1][1] - A[1][0]*A[0][1]);
  }
  template<> inline double sqrt_det_aat<3>(const double (*A)[3]) {
    const double v0 = A[0][1]*A[1][2] - A[1][1]*A[0][2];
    const double v1 = A[0][2]*A[1][0] - A[1][2]*A[0][0];
    const double v2 = A[0][0]*A[1][1] - A[1][0]*A[0][1];
    return std::abs(v0*A[2][0] + v1*A[2][1] + v2*A[2][2]);
  }

  // FieldMatrixHelper::sqrtDetAAT<m,n> for the Jacobians of sub-entities (m < n): explicit 2x3 formula; 1xn through the general
  // path (A A^T is 1x1, its Cholesky factor is the square root); 0xn (vertices): 1; square: as above
  template<int m, int n> inline double sqrt_det_aat_rect(const double (*A)[n]) {
    if constexpr (m == n) return sqrt_det_aat<n>(A);
    else if constexpr (m == 0) return 1.0;
    else if constexpr (m == 2 && n == 3) {
      const double v0 = A[0][0]*A[1][1] - A[0][1]*A[1][0];
      const double v1 = A[0][0]*A[1][2] - A[1][0]*A[0][2];
      const double v2 = A[0][1]*A[1][2] - A[0][2]*A[1][1];
      return std::sqrt(v0*v0 + v1*v1 + v2*v2);
    } else {
      static_assert(m == 1, "only faces of 2-D / 3-D cubes");
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += A[0][k]*A[0][k];
      return std::sqrt(s);
    }
  }

  // FieldMatrixHelper::rightInvA: ret (n x m) with A*ret = I; returns |det|.
  // 2x2: closed form. Otherwise: L L^T = A A^T (AAT_L + cholesky_L), (A A^T)^-1 = L^-T L^-1 (invL + LTL),
  // ret = A^T (A A^T)^-1.
  template<int dim> inline double right_inv(const double (*A)[dim], double (*ret)[dim]);
  template<> inline double right_inv<2>(const double (*A)[2], double (*ret)[2]) {
    const double det = (A[0][0]*A[1][1] - A[1][0]*A[0][1]);
    const double detInv = 1.0/det;
    ret[0][0] =  A[1][1]*detInv;
    ret[1][1] =  A[0][0]*detInv;
    ret[1][0] = -A[1][0]*detInv;
    ret[0][1] = -A[0][1]*detInv;
    return std::abs(det);
  }
  template<> inline double right_inv<3>(const double (*A)[3], double (*ret)[3]) {
    const int m = 3, n = 3;
    double aat[3][3], L[3][3];
    for (int i = 0; i < m; ++i)                       // AAT_L: lower triangle of A A^T
      for (int j = 0; j <= i; ++j) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += A[i][k]*A[j][k];
        aat[i][j] = s;
      }
    for (int i = 0; i < m; ++i) {                     // cholesky_L
      double xDiag = aat[i][i];
      for (int j = 0; j < i; ++j) xDiag -= L[i][j]*L[i][j];
      L[i][i] = std::sqrt(xDiag);
      const double invDiag = 1.0/L[i][i];
      for (int k = i+1; k < m; ++k) {
        double x = aat[k][i];
        for (int j = 0; j < i; ++j) x -= L[i][j]*L[k][j];
        L[k][i] = invDiag*x;
      }
    }
    double det = 1.0;                                  // invL (in place), det = prod of diagonal
    for (int i = 0; i < m; ++i) {
      double& pivot = L[i][i];
      det *= pivot;
      pivot = 1.0/pivot;
      for (int j = 0; j < i; ++j) {
        double x = L[i][j]*L[j][j];
        for (int k = j+1; k < i; ++k) x += L[i][k]*L[k][j];
        L[i][j] = -pivot*x;
      }
    }
    double inv[3][3];                                  // LTL: inv = L^T L (lower), mirrored
    for (int i = 0; i < m; ++i) {
      for (int j = 0; j < i; ++j) {
        double s = 0.0;
        for (int k = i; k < m; ++k) s += L[k][i]*L[k][j];
        inv[i][j] = s; inv[j][i] = s;
      }
      double s = 0.0;
      for (int k = i; k < m; ++k) s += L[k][i]*L[k][i];
      inv[i][i] = s;
    }
    for (int i = 0; i < n; ++i)                        // ATBT: ret = A^T * inv^T
      for (int j = 0; j < m; ++j) {
        double s = 0.0;
        for (int k = 0; k < m; ++k) s += A[k][i]*inv[j][k];
        ret[i][j] = s;
      }
    return det;
  }

  // analytic deformations F for GeometryGrid. id 0: identity; id 1:
  // F(x) = x + a (sin(2 pi x1) sin(2 pi x2), sin(2 pi x2) sin(2 pi x0), sin(2 pi x0) sin(2 pi x1)), a = params[0]
  // (2-D: F(x) = x + a (sin(2 pi x1), sin(2 pi x0)) ).
  inline void deformation(int id, const double* params, int dim, const double* x, double* y) {
    if (id == 0) { for (int i = 0; i < dim; ++i) y[i] = x[i]; return; }
    const double a = params[0];
    const double twopi = 6.283185307179586476925286766559;
    if (dim == 3) {
      const double s0 = std::sin(twopi*x[0]), s1 = std::sin(twopi*x[1]), s2 = std::sin(twopi*x[2]);
      y[0] = x[0] + a*(s1*s2);
      y[1] = x[1] + a*(s2*s0);
      y[2] = x[2] + a*(s0*s1);
    } else {
      const double s0 = std::sin(twopi*x[0]), s1 = std::sin(twopi*x[1]);
      y[0] = x[0] + a*s1;
      y[1] = x[1] + a*s0;
    }
  }
}
#endif
