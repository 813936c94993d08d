// Third tcgen05 micro-benchmark: TWO issuer warps alternating groups of MMAs into the SAME TMEM accumulator.
// Question 1: does the interleaved stream accumulate correctly (A = B = 1.0 -> every element == 16 * #MMAs)?
// Question 2: does the second warp hide the other's per-group scalar overhead (wait + descriptor prep + commit)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_tc3 tools/ubench_tc3.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Args { int iters, nissuers, work; long long* out; float* result; };

// warps: 0 = "producer" (keeps 4 full barriers cycling), 1 and 2 = issuers, 3 = reader of the accumulator at the end.
// `work` = extra dependent integer operations per group standing in for descriptor arithmetic.
template <int NMMA>
__global__ void __launch_bounds__(128, 1) ubench(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384u, bars = base + 32768u;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 64u + 8u * s; };
  const uint32_t done = bars + 128u, slot = bars + 136u, first = bars + 144u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A = B = 1.0 (bf16): the swizzle permutes equal values, so the layout does not matter
  for (uint32_t i = threadIdx.x; i < 32768u / 4; i += blockDim.x) {
    const uint32_t ones = 0x3F803F80u;
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 4u * i), "r"(ones));
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(done, a.nissuers);
    mbar_init(first, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
  const uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    int s = 0, ph = 0;
    for (int i = 0; i < a.iters; ++i) {
      mbar_wait(empty(s), ph ^ 1);
      if (elect_one()) mbar_arrive(full(s));
      if (++s == 4) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 || (warp == 2 && a.nissuers == 2)) {
    const int me = warp - 1;
    const long long t0 = clock64();
    int s = 0, ph = 0;
    uint32_t la = ((sA >> 4) & 0x3FFFu) | (1u << 16), lb = ((sB >> 4) & 0x3FFFu) | (1u << 16);
    uint32_t junk = threadIdx.x;
    for (int i = 0; i < a.iters; ++i) {
      const bool mine = (a.nissuers == 1) || ((i & 1) == me);
      if (mine) {
        mbar_wait(full(s), ph);
        tc_fence_after();
        for (int w = 0; w < a.work; ++w) junk = junk * 1664525u + 1013904223u;   // dependent chain ~ 4-6 clk each
        if (i == 1 && me == 1) mbar_wait(first, 0);   // the accumulate=0 MMA of group 0 must be in the pipe first
        if (elect_one()) {
          const uint32_t z = (junk == 0xFFFFFFFFu) ? 1u : 0u;   // keeps the chain alive, always 0
#pragma unroll
          for (int k = 0; k < NMMA; ++k) {
            const uint32_t off = 2u * (k & 3) + z;
            umma_bf16(tmem, ((uint64_t)hi << 32) | (la + off), ((uint64_t)hi << 32) | (lb + off), id, (i == 0 && k == 0) ? 0u : 1u);
          }
          umma_commit(empty(s));
          if (i == 0) mbar_arrive(first);
        }
      }
      if (++s == 4) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (lane == 0 && me == 0) a.out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 3 && blockIdx.x == 0) {   // TMEM lanes 96..127, first 16 columns
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)96 << 16), v);
    if (lane == 0) { a.result[0] = __uint_as_float(v[0]); a.result[1] = __uint_as_float(v[15]); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int NMMA>
static void run(int nissuers, int work) {
  const int grid = 148;
  Args a;
  a.iters = 2000; a.nissuers = nissuers; a.work = work;
  cudaMalloc(&a.out, sizeof(long long) * grid);
  cudaMalloc(&a.result, sizeof(float) * 2);
  const int smem = 1024 + 32768 + 256;
  cudaFuncSetAttribute(ubench<NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) ubench<NMMA><<<grid, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  float r[2];
  cudaMemcpy(h.data(), a.out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  cudaMemcpy(r, a.result, sizeof(r), cudaMemcpyDeviceToHost);
  cudaFree(a.out); cudaFree(a.result);
  std::sort(h.begin(), h.end());
  const double expect = 16.0 * NMMA * a.iters;
  printf("  issuers %d  nmma/group %2d  work %3d : %7.1f clk/group (MMA bound %4d)   acc = %.0f / %.0f expected %.0f  %s\n", nissuers, NMMA, work,
         (double)h[grid / 2] / a.iters, NMMA * 48, r[0], r[1], expect, (r[0] == expect && r[1] == expect) ? "OK" : "MISMATCH");
}

int main() {
  printf("two issuer warps alternating groups into one accumulator; M=128 N=64 bf16 K=16 (48 clk per MMA, smem bound)\n");
  for (int work : {0, 20, 40, 80}) { run<8>(1, work); run<8>(2, work); }
  for (int work : {0, 40, 80}) { run<16>(1, work); run<16>(2, work); }
  for (int work : {0, 40}) { run<4>(1, work); run<4>(2, work); }
  return 0;
}
