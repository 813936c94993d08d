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
int sm_count() { return 148; }
TraceScope::TraceScope(cudaStream_t stream, const char*, ...) : stream_(stream), index_(-1) {}
TraceScope::~TraceScope() {}
}  // namespace pn

extern "C" const char* pn_last_error_string(void) { return pn::g_err; }

#include "../../packnet_sfm_b200/csrc/fold_kernels.cu"
#include "../../packnet_sfm_b200/csrc/frame_kernels.cu"
#include "../../packnet_sfm_b200/csrc/layer_kernels.cu"
#include "../../packnet_sfm_b200/csrc/loss_kernels.cu"
#include "../../packnet_sfm_b200/csrc/pack_kernels.cu"
#include "../../packnet_sfm_b200/csrc/optim_kernels.cu"

// The tcgen05 / TMA convolution engine has no host emulation (tensor-core instructions): its entry points exist so that
// the ctypes declarations of packnet_sfm_b200/_lib_conv.py resolve, and refuse to run.
#define PN_EMU_UNSUPPORTED(name, ...)                                              \
  extern "C" int name(__VA_ARGS__) {                                               \
    pn::set_error(#name ": the tensor-core engine is not part of the host emulation"); \
    return PN_ERR_UNSUPPORTED;                                                     \
  }
PN_EMU_UNSUPPORTED(pn_conv2d_forward, const pn_conv_desc*, const void*, const void*, const void*, const void*, const float*, float*,
                   uint32_t*, pn_stream_t)
PN_EMU_UNSUPPORTED(pn_conv2d_dgrad, const pn_conv_desc*, const void*, const void*, const void*, const void*, float*, uint32_t*, pn_stream_t)
PN_EMU_UNSUPPORTED(pn_conv2d_packed_weight_elems, int, int, int, int, int, size_t*)
PN_EMU_UNSUPPORTED(pn_conv2d_pack_weight, const float*, void*, void*, int, int, int, int, int, pn_stream_t)
PN_EMU_UNSUPPORTED(pn_conv2d_wgrad, const pn_conv_desc*, const void*, const void*, const void*, const void*, float*, uint32_t*, pn_stream_t)
PN_EMU_UNSUPPORTED(pn_conv2d_wgrad_packed_elems, int, int, int, int, size_t*)
PN_EMU_UNSUPPORTED(pn_conv2d_unpack_weight_grad, const float*, float*, int, int, int, int, pn_stream_t)
PN_EMU_UNSUPPORTED(pn_tf32_residual, const float*, float*, size_t, pn_stream_t)
PN_EMU_UNSUPPORTED(pn_split_bf16, const float*, void*, void*, size_t, pn_stream_t)
extern "C" int pn_conv2d_rows_pad(int cout) {   // conv_engine.cu: rows_pad_of (N tile = min(128, ceil16(cout)), then a multiple of 64)
  if (cout <= 0) return 0;
  const int r16 = (cout + 15) / 16 * 16, bn = r16 < 128 ? r16 : 128;
  int r = (cout + bn - 1) / bn * bn;
  while (r % 64) r += bn;
  return r;
}
extern "C" int pn_version(void) { return 100; }
extern "C" uint64_t pn_launch_count(void) { return 0; }
extern "C" void pn_trace_enable(int) {}
extern "C" int pn_trace_dump(char*, size_t) { return 0; }
