// Package load / unload hooks to append to each package's RcppExports.cpp (or a new init.cpp):
// the CUDA context is created lazily on the first .Call and released when the DLL is unloaded.
#include "b2f_r_context.h"
extern "C" void b2f_r_on_unload(void) { b2f_r_shutdown(); }   // call from R_unload_<pkg>(DllInfo*)
