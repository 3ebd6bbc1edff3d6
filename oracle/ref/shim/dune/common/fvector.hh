// Stand-in for <dune/common/fvector.hh>: dense fixed-size vector with the operations the
// YaspGrid headers use. Arithmetic is plain IEEE, component by component in ascending index.
#ifndef SHIM_DUNE_COMMON_FVECTOR_HH
#define SHIM_DUNE_COMMON_FVECTOR_HH
#include <array>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <ostream>
namespace Dune {
  template<class K, int SIZE>
  class FieldVector {
  public:
    typedef K value_type;
    typedef K field_type;
    typedef std::size_t size_type;
    static constexpr int dimension = SIZE;
    constexpr FieldVector() : d_{} {}
    FieldVector(const K& v) { for (int i=0;i<SIZE;++i) d_[i]=v; }
    FieldVector(std::initializer_list<K> l) : d_{} { int i=0; for (auto& v : l) if (i<SIZE) d_[i++]=v; }
    FieldVector(const std::array<K,SIZE>& a) : d_(a) {}
    K& operator[](size_type i) { return d_[i]; }
    const K& operator[](size_type i) const { return d_[i]; }
    static constexpr size_type size() { return SIZE; }
    K* data() { return d_.data(); }
    const K* data() const { return d_.data(); }
    auto begin() { return d_.begin(); }  auto end() { return d_.end(); }
    auto begin() const { return d_.begin(); }  auto end() const { return d_.end(); }
    FieldVector& operator=(const K& v) { for (int i=0;i<SIZE;++i) d_[i]=v; return *this; }
    FieldVector& operator+=(const FieldVector& o) { for (int i=0;i<SIZE;++i) d_[i]+=o.d_[i]; return *this; }
    FieldVector& operator-=(const FieldVector& o) { for (int i=0;i<SIZE;++i) d_[i]-=o.d_[i]; return *this; }
    FieldVector& operator*=(const K& k) { for (int i=0;i<SIZE;++i) d_[i]*=k; return *this; }
    FieldVector& operator/=(const K& k) { for (int i=0;i<SIZE;++i) d_[i]/=k; return *this; }
    FieldVector& axpy(const K& a, const FieldVector& x) { for (int i=0;i<SIZE;++i) d_[i]+=a*x.d_[i]; return *this; }
    K operator*(const FieldVector& o) const { K s=0; for (int i=0;i<SIZE;++i) s+=d_[i]*o.d_[i]; return s; }
    K dot(const FieldVector& o) const { return (*this)*o; }
    K two_norm2() const { K s=0; for (int i=0;i<SIZE;++i) s+=d_[i]*d_[i]; return s; }
    K two_norm() const { return std::sqrt(two_norm2()); }
    K infinity_norm() const { K m=0; for (int i=0;i<SIZE;++i) m=std::max(m,std::abs(d_[i])); return m; }
    bool operator==(const FieldVector& o) const { return d_==o.d_; }
    bool operator!=(const FieldVector& o) const { return !(d_==o.d_); }
  private:
    std::array<K,SIZE> d_;
  };
  template<class K,int N> FieldVector<K,N> operator+(FieldVector<K,N> a, const FieldVector<K,N>& b) { a+=b; return a; }
  template<class K,int N> FieldVector<K,N> operator-(FieldVector<K,N> a, const FieldVector<K,N>& b) { a-=b; return a; }
  template<class K,int N> FieldVector<K,N> operator*(const K& k, FieldVector<K,N> a) { a*=k; return a; }
  template<class K,int N> FieldVector<K,N> operator*(FieldVector<K,N> a, const K& k) { a*=k; return a; }
  template<class K,int N> std::ostream& operator<<(std::ostream& s, const FieldVector<K,N>& v) {
    for (int i=0;i<N;++i) s << (i?" ":"") << v[i];
    return s;
  }
}
#endif
