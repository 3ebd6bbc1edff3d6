// fv_problem.hh — the synthetic finite-volume transport problems of SURVEY.md §8d, shared by both CPU
// oracles (test infrastructure only). The reference ships no FV example; these
// definitions ARE the specification of row a11 and the CUDA kernels restate them (csrc/dgb_fv.cuh).
#ifndef ORACLE_COMMON_FV_PROBLEM_HH
#define ORACLE_COMMON_FV_PROBLEM_HH
#include <cmath>
namespace orc {
  // velocity u(x): problem 0 constant (params[0..2]); problem 1 rigid rotation in the x-y plane about
  // (0.5,0.5) with angular speed params[0] (time independent, divergence free).
  // problem 2: shear flow u = (p0 (x_z - 0.5), p1, p2 (x_x - 0.5)) - the x velocity depends on z (x_z = 0 in 2-D)
  inline void fv_velocity(int problem, const double* params, int dim, const double* x, double* u) {
    if (problem == 2) {
      const double z = dim > 2 ? x[2] : 0.0;
      const double v[3] = { params[0]*(z-0.5), params[1], params[2]*(x[0]-0.5) };
      for (int i = 0; i < dim; ++i) u[i] = v[i];
    } else if (problem == 1) {
      u[0] = -params[0]*(x[1]-0.5);
      u[1] =  params[0]*(x[0]-0.5);
      for (int i = 2; i < dim; ++i) u[i] = 0.0;
    } else {
      for (int i = 0; i < dim; ++i) u[i] = params[i];
    }
  }
  // inflow boundary value b(x)
  inline double fv_boundary(int /*problem*/, const double* params, int /*dim*/, const double* /*x*/) { return params[3]; }
  // initial concentration: 1 inside the spherical shell r_in < |x - centre| < r_out, else 0
  inline double fv_initial(int /*problem*/, const double* params, int dim, const double* x) {
    double s = 0.0;
    for (int i = 0; i < dim; ++i) { double d = x[i]-params[4+i]; s += d*d; }
    double r = std::sqrt(s);
    return (r > params[7] && r < params[8]) ? 1.0 : 0.0;
  }
}
#endif
