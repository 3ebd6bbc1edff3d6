// Stand-in for <dune/common/hybridutilities.hh> (unused by the compiled path).
#ifndef SHIM_DUNE_COMMON_HYBRIDUTILITIES_HH
#define SHIM_DUNE_COMMON_HYBRIDUTILITIES_HH
#endif
