// api.cu -- version, error text and launch accounting of the C-ABI (include/packnet_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <vector>

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

int sm_count() {
  static std::atomic<int> cache[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

// ---- per-call device timing (diagnostics: pn_trace_enable / pn_trace_dump) ----
struct TraceRec { char tag[112]; cudaEvent_t e0, e1; };
static std::vector<TraceRec> g_trace;
static std::mutex g_trace_mu;
static std::atomic<int> g_trace_on{0};

TraceScope::TraceScope(cudaStream_t stream, const char* fmt, ...) : stream_(stream), index_(-1) {
  if (!g_trace_on.load(std::memory_order_relaxed)) return;
  TraceRec r{};
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(r.tag, sizeof(r.tag), fmt, ap);
  va_end(ap);
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, stream);
  std::lock_guard<std::mutex> lk(g_trace_mu);
  g_trace.push_back(r);
  index_ = (int)g_trace.size() - 1;
}
TraceScope::~TraceScope() {
  if (index_ < 0) return;
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (index_ < (int)g_trace.size()) cudaEventRecord(g_trace[index_].e1, stream_);
}
}  // namespace pn

extern "C" void pn_trace_enable(int on) { pn::g_trace_on.store(on ? 1 : 0); }

// writes "tag<TAB>milliseconds" lines for every traced call since the last dump; returns the number of records
extern "C" int pn_trace_dump(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(pn::g_trace_mu);
  size_t off = 0;
  int n = 0;
  for (auto& r : pn::g_trace) {
    float ms = -1.0f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess) cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (buf && off < cap) off += (size_t)snprintf(buf + off, cap - off, "%s\t%.4f\n", r.tag, ms);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    ++n;
  }
  pn::g_trace.clear();
  return n;
}

extern "C" int pn_version(void) { return 100; }
extern "C" const char* pn_last_error_string(void) { return pn::g_err; }
extern "C" uint64_t pn_launch_count(void) { return pn::g_launches.load(std::memory_order_relaxed); }
