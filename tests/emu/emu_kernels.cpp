// emu_kernels.cpp -- TEST INFRASTRUCTURE: the fold and frame kernels of libpacknet_b200 compiled FROM THEIR REAL SOURCE
// for the host (g++ -DPN_EMULATE, tests/emu/cuda_emu.h) behind the same C-ABI entry points, so that the CPU test tier can
// execute them.  Built by tests/test_kernels_emulated_cpu.py into tests/emu/_build/ (git-ignored).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char*) { return 0; }
void count_launch(int) {}
TraceScope::TraceScope(cudaStream_t stream, const char*, ...) : stream_(stream), index_(-1) {}
TraceScope::~TraceScope() {}
}  // namespace pn

extern "C" const char* pn_last_error_string(void) { return pn::g_err; }

#include "../../packnet_sfm_b200/csrc/fold_kernels.cu"
#include "../../packnet_sfm_b200/csrc/frame_kernels.cu"
