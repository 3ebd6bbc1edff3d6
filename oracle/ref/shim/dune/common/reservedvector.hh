// Stand-in for <dune/common/reservedvector.hh>: fixed-capacity vector whose elements never move
// (YaspGrid keeps iterators into its level array across resize).
#ifndef SHIM_DUNE_COMMON_RESERVEDVECTOR_HH
#define SHIM_DUNE_COMMON_RESERVEDVECTOR_HH
#include <array>
#include <cstddef>
namespace Dune {
  template<class T, int n>
  class ReservedVector {
  public:
    typedef T* iterator;
    typedef const T* const_iterator;
    typedef std::size_t size_type;
    ReservedVector() : sz_(0) {}
    void resize(size_type s) { sz_ = s; }
    void push_back(const T& t) { d_[sz_++] = t; }
    void pop_back() { --sz_; }
    void clear() { sz_ = 0; }
    size_type size() const { return sz_; }
    bool empty() const { return sz_==0; }
    T& operator[](size_type i) { return d_[i]; }
    const T& operator[](size_type i) const { return d_[i]; }
    T& back() { return d_[sz_-1]; }
    const T& back() const { return d_[sz_-1]; }
    T& front() { return d_[0]; }
    iterator begin() { return d_.data(); }
    iterator end() { return d_.data()+sz_; }
    const_iterator begin() const { return d_.data(); }
    const_iterator end() const { return d_.data()+sz_; }
  private:
    std::array<T,n> d_;
    size_type sz_;
  };
}
#endif
