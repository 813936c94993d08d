// common.cuh -- shared host/device helpers of libpacknet_b200 (error reporting, launch accounting).
#pragma once
#ifdef PN_EMULATE
// host build of the plain SIMT kernels for the CPU test tier (tests/emu/cuda_emu.h); never part of the product library
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
// kernel launch, spelled as a macro so that the emulated host build can run the same dispatch code
#define PN_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
// the dynamic shared memory of a kernel (16-byte aligned)
#define PN_DYNAMIC_SHARED(type, name) extern __shared__ __align__(16) type name[]
#define PN_DYNAMIC_SHARED_PLAIN(type, name) extern __shared__ type name[]
#endif

#include <cstdint>
#include <cstdio>

#include "packnet_b200.h"

namespace pn {

// thread-local last-error text behind pn_last_error_string()
void set_error(const char* fmt, ...);
// cudaGetLastError() -> return code (0 ok, >0 cudaError_t) with the error text recorded
int check_launch(const char* what);
// launch accounting behind pn_launch_count()
void count_launch(int n = 1);
// SMs of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a B200): grid sizing / wave counting
int sm_count();
// device-time trace of one C-ABI call (no-op unless pn_trace_enable(1))
struct TraceScope {
  TraceScope(cudaStream_t stream, const char* fmt, ...);
  ~TraceScope();
  cudaStream_t stream_;
  int index_;
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#define PN_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      ::pn::set_error(__VA_ARGS__);      \
      return (code);                     \
    }                                    \
  } while (0)

#define PN_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::pn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                  \
    }                                                                                  \
  } while (0)

}  // namespace pn
