// Stand-in for <dune/common/indent.hh> (dune-common is not vendored): the indentation helper the reference's VTK data array
// writers take (dune/grid/io/file/vtk/dataarraywriter.hh). RESTATED from the published class: an Indent prints its parent's
// indentation followed by `level` copies of the basic indentation string. Test infrastructure only.
#ifndef SHIM_DUNE_COMMON_INDENT_HH
#define SHIM_DUNE_COMMON_INDENT_HH
#include <ostream>
#include <string>
namespace Dune {
  class Indent {
    const Indent* parent;
    std::string basic_indent;
    unsigned level;
  public:
    inline Indent(const std::string& basic_indent_ = "  ", unsigned level_ = 0) : parent(0), basic_indent(basic_indent_), level(level_) {}
    inline Indent(unsigned level_) : parent(0), basic_indent("  "), level(level_) {}
    inline Indent(const Indent* parent_, const std::string& basic_indent_ = "  ", unsigned level_ = 1)
      : parent(parent_), basic_indent(basic_indent_), level(level_) {}
    inline Indent(const Indent* parent_, unsigned level_) : parent(parent_), basic_indent("  "), level(level_) {}
    inline Indent operator+(const std::string& newindent) const { return Indent(this, newindent); }
    inline Indent operator+(unsigned morelevel) const { return Indent(parent, basic_indent, level+morelevel); }
    inline Indent& operator++() { ++level; return *this; }
    inline Indent& operator--() { if (level > 0) --level; return *this; }
    friend inline std::ostream& operator<<(std::ostream& s, const Indent& indent);
  };
  inline std::ostream& operator<<(std::ostream& s, const Indent& indent) {
    if (indent.parent) s << *indent.parent;
    for (unsigned i = 0; i < indent.level; ++i) s << indent.basic_indent;
    return s;
  }
}
#endif
