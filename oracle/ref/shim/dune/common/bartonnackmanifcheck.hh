// Stand-in for <dune/common/bartonnackmanifcheck.hh>: interface checks are plain calls here.
#ifndef CHECK_INTERFACE_IMPLEMENTATION
#define CHECK_INTERFACE_IMPLEMENTATION(dummy)
#endif
#ifndef CHECK_AND_CALL_INTERFACE_IMPLEMENTATION
#define CHECK_AND_CALL_INTERFACE_IMPLEMENTATION(__interface_method_to_call__) (__interface_method_to_call__)
#endif
