// cuda_emu.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the SIMT kernels of
// fold_kernels.cu / frame_kernels.cu on the host, from their real source (compiled with g++ -DPN_EMULATE).
// One OS thread per CUDA thread of a block, blocks one after the other; __syncthreads = std::barrier over the block,
// __shfl_xor_sync = exchange through a per-warp buffer, atomicAdd = a mutex, __shared__ = static.
// It exists so that the CPU test tier can execute the kernels that otherwise only run on the GPU tier
// (tests/test_kernels_emulated_cpu.py).  It checks index arithmetic, staging, synchronisation placement and the host-side
// dispatch; it says nothing about performance, memory-model subtleties or tcgen05/TMA code (not used by these kernels).
#pragma once
#include <algorithm>
#include <barrier>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }

struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(8) float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct alignas(8) uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
using std::max;
using std::min;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) alignas(n)
// IEEE single operations with one rounding each: the host build is compiled with -ffp-contract=off
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __saturatef(float a) { return a != a ? 0.0f : (a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a)); }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline void __threadfence() {}

namespace cuda_emu {
struct BlockCtx {
  std::barrier<>* bar;
  double (*warp_buf)[32];
  std::barrier<>** warp_bar;
  void* dyn_smem;
};
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_gridDim, t_blockDim;
inline thread_local BlockCtx t_ctx;
inline thread_local unsigned t_shfl_count = 0;
inline std::mutex g_atomic_mu;
}  // namespace cuda_emu

#define threadIdx (cuda_emu::t_threadIdx)
#define blockIdx (cuda_emu::t_blockIdx)
#define gridDim (cuda_emu::t_gridDim)
#define blockDim (cuda_emu::t_blockDim)

inline void __syncthreads() { cuda_emu::t_ctx.bar->arrive_and_wait(); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> lk(cuda_emu::g_atomic_mu);
  const T old = *p;
  *p = old + v;
  return old;
}
// One barrier per shuffle: the exchange buffer is double-buffered by the parity of the thread's shuffle count.  A lane
// can only reach shuffle n+2 (which reuses buffer n % 2) after the barrier of shuffle n+1, i.e. after every lane has
// finished reading buffer n % 2.
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  const unsigned tid = cuda_emu::t_threadIdx.x, warp = tid / 32, lane = tid % 32;
  const unsigned par = (cuda_emu::t_shfl_count++) & 1u;
  double(*buf)[32] = cuda_emu::t_ctx.warp_buf + 2 * warp + par;
  (*buf)[lane] = (double)v;
  cuda_emu::t_ctx.warp_bar[warp]->arrive_and_wait();
  return (T)(*buf)[lane ^ (unsigned)lane_mask];
}

namespace cuda_emu {
// run `kernel(args...)` for every block of `grid` with `block.x` threads (1-D blocks only)
template <class K, class... Args>
void launch(K kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
  const unsigned nt = block.x, nwarps = (nt + 31) / 32;
  std::vector<float4> dyn((smem_bytes + 15) / 16 + 1);
  std::barrier<> bar((std::ptrdiff_t)nt);
  std::vector<std::unique_ptr<std::barrier<>>> wbars;
  std::vector<std::barrier<>*> wptr;
  for (unsigned w = 0; w < nwarps; ++w) {
    const unsigned lanes = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
    wbars.emplace_back(new std::barrier<>((std::ptrdiff_t)lanes));
    wptr.push_back(wbars.back().get());
  }
  std::vector<double[32]> wbuf(2 * nwarps);
  auto worker = [&](unsigned tid) {
    t_blockDim = block;
    t_gridDim = grid;
    t_ctx = BlockCtx{&bar, wbuf.data(), wptr.data(), dyn.data()};
    t_shfl_count = 0;
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_threadIdx = uint3{tid, 0, 0};
          t_blockIdx = uint3{bx, by, bz};
          kernel(args...);
          bar.arrive_and_wait();   // the next block reuses the static "shared memory"
        }
  };
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back(worker, t);
  for (auto& t : th) t.join();
}
}  // namespace cuda_emu

#define PN_LAUNCH(kernel, grid, block, smem, stream, ...) cuda_emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define PN_DYNAMIC_SHARED(type, name) type* name = reinterpret_cast<type*>(cuda_emu::t_ctx.dyn_smem)
#define PN_DYNAMIC_SHARED_PLAIN(type, name) type* name = reinterpret_cast<type*>(cuda_emu::t_ctx.dyn_smem)
