// Stand-in for <dune/common/typetraits.hh> (nothing on the YaspGrid path needs its contents).
#ifndef SHIM_DUNE_COMMON_TYPETRAITS_HH
#define SHIM_DUNE_COMMON_TYPETRAITS_HH
#include <type_traits>
#endif
