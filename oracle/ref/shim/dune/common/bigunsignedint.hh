// Stand-in for <dune/common/bigunsignedint.hh>: k <= 128 bits suffices for YaspGrid's
// persistent ids up to dim 4 (4*24+5+4 = 105 bits). Only used by id-set code.
#ifndef SHIM_DUNE_COMMON_BIGUNSIGNEDINT_HH
#define SHIM_DUNE_COMMON_BIGUNSIGNEDINT_HH
#include <cstdint>
#include <functional>
#include <ostream>
namespace Dune {
  template<int k>
  class bigunsignedint {
    static_assert(k <= 128, "shim bigunsignedint supports at most 128 bits");
    typedef unsigned __int128 U;
    static constexpr U mask() { return k == 128 ? ~U(0) : ((U(1) << k) - 1); }
  public:
    bigunsignedint() : v_(0) {}
    template<class I, class = std::enable_if_t<std::is_integral<I>::value>>
    bigunsignedint(I x) : v_(U(static_cast<std::make_unsigned_t<I>>(x)) & mask()) {}
    bigunsignedint operator<<(int s) const { bigunsignedint r; r.v_ = (s>=128 ? U(0) : (v_ << s)) & mask(); return r; }
    bigunsignedint operator>>(int s) const { bigunsignedint r; r.v_ = (s>=128 ? U(0) : (v_ >> s)); return r; }
    bigunsignedint operator+(const bigunsignedint& o) const { bigunsignedint r; r.v_ = (v_+o.v_) & mask(); return r; }
    bigunsignedint operator|(const bigunsignedint& o) const { bigunsignedint r; r.v_ = v_|o.v_; return r; }
    bool operator==(const bigunsignedint& o) const { return v_==o.v_; }
    bool operator!=(const bigunsignedint& o) const { return v_!=o.v_; }
    bool operator<(const bigunsignedint& o) const { return v_<o.v_; }
    std::uint64_t lo() const { return std::uint64_t(v_); }
    std::uint64_t hi() const { return std::uint64_t(v_ >> 64); }
  private:
    U v_;
  };
  template<int k> std::ostream& operator<<(std::ostream& s, const bigunsignedint<k>& x) { return s << x.hi() << ":" << x.lo(); }
}
namespace std {
  template<int k> struct hash<Dune::bigunsignedint<k>> {
    size_t operator()(const Dune::bigunsignedint<k>& x) const { return hash<uint64_t>()(x.lo() ^ (x.hi()*0x9e3779b97f4a7c15ull)); }
  };
}
#endif
