// Stand-in for <dune/common/parallel/communication.hh>: sequential Communication<No_Comm>.
#ifndef SHIM_DUNE_COMMON_PARALLEL_COMMUNICATION_HH
#define SHIM_DUNE_COMMON_PARALLEL_COMMUNICATION_HH
namespace Dune {
  struct No_Comm {};
  template<class C>
  class Communication {
  public:
    Communication() {}
    Communication(const C&) {}
    int rank() const { return 0; }
    int size() const { return 1; }
    template<class T> T sum(const T& x) const { return x; }
    template<class T> T max(const T& x) const { return x; }
    template<class T> T min(const T& x) const { return x; }
    int barrier() const { return 0; }
    template<class Op, class T> int allreduce(const T* in, T* out, int n) const { for (int i=0;i<n;++i) out[i]=in[i]; return 0; }
  };
}
#endif
