// cuda_emu.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the SIMT kernels of
// fold_kernels.cu / frame_kernels.cu on the host, from their real source (compiled with g++ -DPN_EMULATE).
// One OS thread per CUDA thread of a block, blocks one after the other; __syncthreads = std::barrier over the block,
// __shfl_xor_sync = exchange through a per-warp buffer, atomicAdd = a mutex, __shared__ = static.
// It exists so that the CPU test tier can execute the kernels that otherwise only run on the GPU tier
// (tests/test_kernels_emulated_cpu.py).  It checks index arithmetic, staging, synchronisation placement and the host-side
// dispatch; it says nothing about performance, memory-model subtleties or tcgen05/TMA code (not used by these kernels).
#pragma once
#include <barrier>
#include <cmath>
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
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static

namespace cuda_emu {
struct BlockCtx {
  std::barrier<>* bar;
  float (*warp_buf)[32];
  std::barrier<>** warp_bar;
};
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_gridDim, t_blockDim;
inline thread_local BlockCtx t_ctx;
inline std::mutex g_atomic_mu;
}  // namespace cuda_emu

#define threadIdx (cuda_emu::t_threadIdx)
#define blockIdx (cuda_emu::t_blockIdx)
#define gridDim (cuda_emu::t_gridDim)
#define blockDim (cuda_emu::t_blockDim)

inline void __syncthreads() { cuda_emu::t_ctx.bar->arrive_and_wait(); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float atomicAdd(float* p, float v) {
  std::lock_guard<std::mutex> lk(cuda_emu::g_atomic_mu);
  const float old = *p;
  *p = old + v;
  return old;
}
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const unsigned tid = cuda_emu::t_threadIdx.x, warp = tid / 32, lane = tid % 32;
  cuda_emu::t_ctx.warp_buf[warp][lane] = v;
  cuda_emu::t_ctx.warp_bar[warp]->arrive_and_wait();
  const float r = cuda_emu::t_ctx.warp_buf[warp][lane ^ (unsigned)lane_mask];
  cuda_emu::t_ctx.warp_bar[warp]->arrive_and_wait();
  return r;
}

namespace cuda_emu {
// run `kernel(args...)` for every block of `grid` with `block.x` threads (1-D blocks only)
template <class K, class... Args>
void launch(K kernel, dim3 grid, dim3 block, Args... args) {
  const unsigned nt = block.x, nwarps = (nt + 31) / 32;
  std::barrier<> bar((std::ptrdiff_t)nt);
  std::vector<std::unique_ptr<std::barrier<>>> wbars;
  std::vector<std::barrier<>*> wptr;
  for (unsigned w = 0; w < nwarps; ++w) {
    const unsigned lanes = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
    wbars.emplace_back(new std::barrier<>((std::ptrdiff_t)lanes));
    wptr.push_back(wbars.back().get());
  }
  std::vector<float[32]> wbuf(nwarps);
  auto worker = [&](unsigned tid) {
    t_blockDim = block;
    t_gridDim = grid;
    t_ctx = BlockCtx{&bar, wbuf.data(), wptr.data()};
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

#define PN_LAUNCH(kernel, grid, block, smem, stream, ...) cuda_emu::launch(kernel, dim3(grid), dim3(block), __VA_ARGS__)
