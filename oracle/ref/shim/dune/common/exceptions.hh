// Stand-in for dune-common's <dune/common/exceptions.hh> (dune-common is not vendored in the
// reference tree). Test infrastructure only: lets the reference's YaspGrid headers compile
// unmodified for oracle/_ref. Written from the documented interface, not copied.
#ifndef SHIM_DUNE_COMMON_EXCEPTIONS_HH
#define SHIM_DUNE_COMMON_EXCEPTIONS_HH
#include <exception>
#include <sstream>
#include <string>
namespace Dune {
  class Exception : public std::exception {
  public:
    Exception() = default;
    void message(const std::string& m) { msg_ = m; }
    const char* what() const noexcept override { return msg_.c_str(); }
  private:
    std::string msg_;
  };
  class IOError : public Exception {};
  class MathError : public Exception {};
  class RangeError : public Exception {};
  class NotImplemented : public Exception {};
  class SystemError : public Exception {};
  class OutOfMemoryError : public SystemError {};
  class InvalidStateException : public Exception {};
  class ParallelError : public Exception {};
}
#define DUNE_THROW(E, m) do { E th__ex; std::ostringstream th__out; \
    th__out << #E << " [" << __func__ << ":" << __FILE__ << ":" << __LINE__ << "]: " << m; \
    th__ex.message(th__out.str()); throw th__ex; } while (0)
#endif
