// Second tcgen05 micro-benchmark: how lean can the issuer loop get?  Compile-time MMA count, descriptors hoisted out of the
// loop, optional per-iteration commit / try_wait(on an already completed phase) / clock watchdog.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_tc2 tools/ubench_tc2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

struct Args { int iters, n1, n2; long long* out; long long* out_issue; };

// FLAGS: 1 = commit per iteration, 2 = try_wait (already complete) + fence per iteration, 4 = clock64 watchdog read per
// iteration, 8 = the whole body under `lane == 0` instead of elect, 16 = descriptors advance per iteration (stage walk)
template <int NMMA, int FLAGS>
__global__ void __launch_bounds__(192, 1) ubench(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 65536u, bars = base + 65536u + 65536u;
  const uint32_t done = bars + 128u, slot = bars + 136u, ready = bars + 144u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(bars + 8u * s, 1);
    mbar_init(done, 1);
    mbar_init(ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  auto idesc = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
  const uint32_t id1 = idesc(a.n1), id2 = idesc(a.n2);

  if (warp == 1) {
    const long long t0 = clock64();
    uint32_t la = ((sA >> 4) & 0x3FFFu) | (1u << 16), lb = ((sB >> 4) & 0x3FFFu) | (1u << 16);
    int s = 0;
    uint32_t ebar = bars;
    for (int i = 0; i < a.iters; ++i) {
      if (FLAGS & 2) {
        if (FLAGS & 4) {
          const long long w0 = clock64();
          while (!mbar_try_wait(ready, 1)) { if (clock64() - w0 > 4000000000LL) __trap(); }
        } else {
          mbar_wait(ready, 1);   // parity 1 of a fresh barrier: the "previous phase" is complete -> returns at once
        }
        tc_fence_after();
      }
      const bool me = (FLAGS & 8) ? ((threadIdx.x & 31) == 0) : elect_one();
      if (me) {
#pragma unroll
        for (int k = 0; k < NMMA; ++k) {
          const uint32_t off = 2u * ((k >> 1) & 3);
          if (k & 1) umma_bf16(tmem + 256, ((uint64_t)hi << 32) | (la + off + 512u), ((uint64_t)hi << 32) | (lb + off), id2, 1u);
          else umma_bf16(tmem, ((uint64_t)hi << 32) | (la + off), ((uint64_t)hi << 32) | (lb + off), id1, 1u);
        }
        if (FLAGS & 1) umma_commit(ebar);
      }
      if (FLAGS & 16) {
        la += 8u; lb += 256u; ebar += 8u;
        if (++s == 4) { s = 0; la -= 32u; lb -= 1024u; ebar -= 32u; }
      }
    }
    const long long t1 = clock64();
    if ((FLAGS & 8) ? ((threadIdx.x & 31) == 0) : elect_one()) umma_commit(done);
    mbar_wait(done, 0);
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { a.out[blockIdx.x] = t2 - t0; a.out_issue[blockIdx.x] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int NMMA, int FLAGS>
static void run(const char* tag, int n1, int n2) {
  const int grid = 148;
  Args a;
  a.iters = 2000; a.n1 = n1; a.n2 = n2;
  cudaMalloc(&a.out, sizeof(long long) * grid);
  cudaMalloc(&a.out_issue, sizeof(long long) * grid);
  const int smem = 1024 + 65536 + 65536 + 256;
  cudaFuncSetAttribute(ubench<NMMA, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) ubench<NMMA, FLAGS><<<grid, 192, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid), hi(grid);
  cudaMemcpy(h.data(), a.out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  cudaMemcpy(hi.data(), a.out_issue, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  cudaFree(a.out); cudaFree(a.out_issue);
  std::sort(h.begin(), h.end()); std::sort(hi.begin(), hi.end());
  printf("  %-46s nmma %2d N=%3d/%3d : total %7.1f  issue-side %7.1f clk/iter\n", tag, NMMA, n1, n2,
         (double)h[grid / 2] / a.iters, (double)hi[grid / 2] / a.iters);
}

int main() {
  printf("clk per iteration, median over 148 CTAs x 2000 iterations; M=128 bf16 K=16\n");
  run<1, 0>("bare loop", 64, 64);
  run<2, 0>("bare loop", 64, 64);
  run<4, 0>("bare loop", 64, 64);
  run<8, 0>("bare loop", 64, 64);
  run<16, 0>("bare loop", 64, 64);
  run<4, 0>("bare loop", 128, 128);
  run<8, 0>("bare loop", 128, 128);
  run<4, 0>("bare loop", 256, 256);
  run<8, 0>("bare loop", 256, 256);
  run<8, 0>("bare loop", 64, 128);
  run<8, 8>("lane==0 instead of elect", 64, 128);
  run<8, 1>("+commit", 64, 128);
  run<8, 2>("+try_wait+fence", 64, 128);
  run<8, 3>("+commit +try_wait", 64, 128);
  run<8, 7>("+commit +try_wait +watchdog clock", 64, 128);
  run<8, 16>("+stage walk", 64, 128);
  run<8, 19>("+commit +try_wait +stage walk", 64, 128);
  run<8, 23>("+commit +try_wait +watchdog +stage walk", 64, 128);
  run<4, 19>("+commit +try_wait +stage walk", 64, 64);
  run<4, 19>("+commit +try_wait +stage walk", 128, 128);
  run<16, 19>("+commit +try_wait +stage walk", 64, 128);
  run<1, 19>("+commit +try_wait +stage walk", 64, 64);
  return 0;
}
