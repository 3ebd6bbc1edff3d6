// Stand-in for <dune/common/fmatrix.hh> and DiagonalMatrix (only what geometry needs).
#ifndef SHIM_DUNE_COMMON_FMATRIX_HH
#define SHIM_DUNE_COMMON_FMATRIX_HH
#include "fvector.hh"
namespace Dune {
  template<class K,int R,int C>
  class FieldMatrix {
  public:
    static constexpr int rows = R, cols = C;
    FieldMatrix() { for (int i=0;i<R;++i) r_[i]=K(0); }
    FieldMatrix(const K& v) { for (int i=0;i<R;++i) r_[i]=v; }
    FieldVector<K,C>& operator[](std::size_t i) { return r_[i]; }
    const FieldVector<K,C>& operator[](std::size_t i) const { return r_[i]; }
    template<class X,class Y> void mv(const X& x, Y& y) const {
      for (int i=0;i<R;++i) { y[i]=0; for (int j=0;j<C;++j) y[i]+=r_[i][j]*x[j]; } }
    template<class X,class Y> void mtv(const X& x, Y& y) const {
      for (int j=0;j<C;++j) { y[j]=0; for (int i=0;i<R;++i) y[j]+=r_[i][j]*x[i]; } }
  private:
    std::array<FieldVector<K,C>,R> r_;
  };
}
#endif
