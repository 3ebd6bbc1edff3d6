// Stand-in for <dune/common/parallel/mpihelper.hh> (the oracle never initialises real MPI).
#ifndef SHIM_DUNE_COMMON_PARALLEL_MPIHELPER_HH
#define SHIM_DUNE_COMMON_PARALLEL_MPIHELPER_HH
#endif
