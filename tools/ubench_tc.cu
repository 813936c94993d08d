// Micro-benchmark of the tcgen05 issue / commit / mbarrier hand-off costs that bound conv_igemm_kernel's pipeline
// (DESIGN.md 3.2).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tc tools/ubench_tc.cu
// Every mode runs ITER iterations in one CTA per SM and reports clocks per iteration (median over CTAs).
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
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void mbar_wait_test(uint32_t bar, uint32_t parity) { while (!mbar_test_wait(bar, parity)) {} }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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

struct Args {
  int mode, iters, nmma, n1, n2, stages, use_test_wait;
  long long* out;
};

// modes:
//  0  issuer alone: nmma MMAs (alternating N = n1 / n2) + one commit per iteration, never waits
//  1  issuer alone: nmma MMAs per iteration, ONE commit at the very end
//  2  two warps ping-pong over `stages` slots: producer waits empty / arrives full (plain mbarrier.arrive);
//     issuer waits full / nmma MMAs / tcgen05.commit -> empty
//  3  like 2, but the issuer releases the slot with a plain mbarrier.arrive (no commit; MMAs unsynchronised)
//  4  like 2, commit only every second iteration releases TWO slots (empty barrier per slot pair)
__global__ void __launch_bounds__(192, 1) ubench(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384u, bars = base + 49152u;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 64u + 8u * s; };
  const uint32_t done = bars + 128u, slot = bars + 136u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(done, 1);
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
  auto desc = [&](uint32_t addr) { return ((uint64_t)hi << 32) | (uint64_t)(((addr >> 4) & 0x3FFFu) | (1u << 16)); };
  auto idesc = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
  const uint32_t id1 = idesc(a.n1), id2 = idesc(a.n2);
  const int S = a.stages;

  if (warp == 0 && a.mode >= 2) {
    int s = 0, ph = 0;
    for (int i = 0; i < a.iters; ++i) {
      const int eb = (a.mode == 4) ? (s | 1) : s;   // slot pair shares the odd slot's empty barrier
      if (a.use_test_wait) mbar_wait_test(empty(eb), ph ^ 1); else mbar_wait(empty(eb), ph ^ 1);
      if (elect_one()) mbar_arrive(full(s));
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    long long t0 = clock64();
    int s = 0, ph = 0;
    for (int i = 0; i < a.iters; ++i) {
      if (a.mode >= 2) {
        if (a.use_test_wait) mbar_wait_test(full(s), ph); else mbar_wait(full(s), ph);
        tc_fence_after();
      }
      if (elect_one()) {
        for (int k = 0; k < a.nmma; ++k) {
          const uint32_t off = 2u * (k & 3);
          if (k & 1) umma_bf16(tmem + 256, desc(sA) + off, desc(sB) + off, id2, 1u);
          else umma_bf16(tmem, desc(sA) + off, desc(sB) + off, id1, 1u);
        }
        if (a.mode == 0 || a.mode == 2) umma_commit(empty(s));
        if (a.mode == 3) mbar_arrive(empty(s));
        if (a.mode == 4 && (s & 1)) umma_commit(empty(s));
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(done);
    mbar_wait(done, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) a.out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static double run(int mode, int nmma, int n1, int n2, int stages, int test_wait, int grid = 148) {
  Args a;
  a.mode = mode; a.iters = 2000; a.nmma = nmma; a.n1 = n1; a.n2 = n2; a.stages = stages; a.use_test_wait = test_wait;
  cudaMalloc(&a.out, sizeof(long long) * grid);
  const int smem = 1024 + 49152 + 256;
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) ubench<<<grid, 192, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), a.out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  cudaFree(a.out);
  std::sort(h.begin(), h.end());
  return (double)h[grid / 2] / a.iters;
}

int main() {
  printf("clk per iteration (median over 148 CTAs, 2000 iterations); M=128 bf16 K=16 MMAs, N alternating n1/n2\n");
  printf("-- issuer alone, commit every iteration (mode 0)\n");
  for (int nm : {0, 1, 2, 4, 8, 16}) printf("  nmma %2d N=64/64   : %7.1f\n", nm, run(0, nm, 64, 64, 4, 0));
  for (int nm : {4, 8, 16}) printf("  nmma %2d N=64/128  : %7.1f\n", nm, run(0, nm, 64, 128, 4, 0));
  for (int nm : {4, 8}) printf("  nmma %2d N=128/128 : %7.1f\n", nm, run(0, nm, 128, 128, 4, 0));
  for (int nm : {4, 8}) printf("  nmma %2d N=256/256 : %7.1f\n", nm, run(0, nm, 256, 256, 4, 0));
  printf("-- issuer alone, ONE commit at the end (mode 1)\n");
  for (int nm : {4, 8}) printf("  nmma %2d N=64/64   : %7.1f\n", nm, run(1, nm, 64, 64, 4, 0));
  for (int nm : {4, 8}) printf("  nmma %2d N=64/128  : %7.1f\n", nm, run(1, nm, 64, 128, 4, 0));
  for (int nm : {4}) printf("  nmma %2d N=128/128 : %7.1f\n", nm, run(1, nm, 128, 128, 4, 0));
  for (int nm : {4}) printf("  nmma %2d N=256/256 : %7.1f\n", nm, run(1, nm, 256, 256, 4, 0));
  printf("-- ping-pong producer<->issuer, commit releases the slot (mode 2)\n");
  for (int S : {1, 2, 4, 8})
    for (int nm : {0, 4, 8})
      printf("  stages %d nmma %2d N=64/128 try_wait : %7.1f   test_wait : %7.1f\n", S, nm, run(2, nm, 64, 128, S, 0), run(2, nm, 64, 128, S, 1));
  printf("-- ping-pong, slot released by a plain mbarrier.arrive (mode 3)\n");
  for (int S : {1, 4, 8})
    for (int nm : {0, 8})
      printf("  stages %d nmma %2d N=64/128 try_wait : %7.1f   test_wait : %7.1f\n", S, nm, run(3, nm, 64, 128, S, 0), run(3, nm, 64, 128, S, 1));
  printf("-- ping-pong, ONE commit per slot pair (mode 4)\n");
  for (int S : {4, 8})
    for (int nm : {0, 4, 8})
      printf("  stages %d nmma %2d N=64/128 try_wait : %7.1f\n", S, nm, run(4, nm, 64, 128, S, 0));
  printf("-- single CTA (grid 1), mode 2, stages 4\n");
  for (int nm : {0, 8}) printf("  nmma %2d : %7.1f\n", nm, run(2, nm, 64, 128, 4, 0, 1));
  return 0;
}
