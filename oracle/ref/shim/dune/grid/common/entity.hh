// <dune/grid/common/entity.hh> is part of the facade stand-in (grid.hh); dune/grid/geometrygrid/hostcorners.hh includes it by name.
#ifndef SHIM_DUNE_GRID_COMMON_ENTITY_HH
#define SHIM_DUNE_GRID_COMMON_ENTITY_HH
#include <dune/grid/common/grid.hh>
#endif
