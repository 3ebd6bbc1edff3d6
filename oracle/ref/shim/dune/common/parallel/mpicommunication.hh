// Stand-in for <dune/common/parallel/mpicommunication.hh> on top of the in-process mpi.h shim.
#ifndef SHIM_DUNE_COMMON_PARALLEL_MPICOMMUNICATION_HH
#define SHIM_DUNE_COMMON_PARALLEL_MPICOMMUNICATION_HH
#include <mpi.h>
#include "communication.hh"
namespace Dune {
  template<>
  class Communication<MPI_Comm> {
  public:
    Communication() : c_(nullptr) {}
    Communication(MPI_Comm c) : c_(c) {}
    int rank() const { return c_ ? c_->rank : 0; }
    int size() const { return c_ ? c_->world->size : 1; }
    operator MPI_Comm() const { return c_; }
    template<class T> T max(const T& x) const { T r = x; if (c_) vmpi_reduce<T>(c_, &x, &r, 1, [](T a, T b){ return a>b?a:b; }); return r; }
    template<class T> T min(const T& x) const { T r = x; if (c_) vmpi_reduce<T>(c_, &x, &r, 1, [](T a, T b){ return a<b?a:b; }); return r; }
    template<class T> T sum(const T& x) const { T r = x; if (c_) vmpi_reduce<T>(c_, &x, &r, 1, [](T a, T b){ return a+b; }); return r; }
    int barrier() const { return c_ ? MPI_Barrier(c_) : 0; }
    // reference dune/common/parallel/mpicommunication.hh allgather(sbuf, count, rbuf): rank-ordered concatenation
    template<class T> int allgather(const T* sbuf, int count, T* rbuf) const {
      if (!c_) { for (int i = 0; i < count; ++i) rbuf[i] = sbuf[i]; return 0; }
      std::vector<std::vector<char>> all;
      vmpi_allgather(c_, sbuf, int(sizeof(T))*count, all);
      for (std::size_t r = 0; r < all.size(); ++r) std::memcpy(rbuf + r*count, all[r].data(), sizeof(T)*count);
      return 0;
    }
    template<class Op, class T> int allreduce(const T* in, T* out, int n) const {
      if (!c_) { for (int i=0;i<n;++i) out[i]=in[i]; return 0; }
      vmpi_reduce<T>(c_, in, out, n, [](T a, T b){ return Op()(a,b); }); return 0;
    }
  private:
    MPI_Comm c_;
  };
}
#endif
