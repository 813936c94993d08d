// api.cu -- version, error text and launch accounting of the C-ABI (include/packnet_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace pn {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace pn

extern "C" int pn_version(void) { return 100; }
extern "C" const char* pn_last_error_string(void) { return pn::g_err; }
extern "C" uint64_t pn_launch_count(void) { return pn::g_launches.load(std::memory_order_relaxed); }
