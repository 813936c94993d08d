// optim_kernels.cu -- the optimizer step of the training loop on ONE flat parameter buffer, fused with the per-step weight
// re-layout the convolution engine needs.
//
// Reference: Adam per parameter group (packnet_sfm/models/model_wrapper.py:128-166 -> torch.optim.Adam, lr 2e-4 for the 'Depth'
// and 'Pose' groups, weight_decay 0) applied after the gradient all-reduce (trainers/horovod_trainer.py:46-48,92-93).
//
// Round 1 spent 3.5 ms of a 38 ms step around the 128 M weights: pack_weight_kernel x93 (fp32 OIHW -> bf16 hi/lo tiles, forward
// and transposed, strided gathers, 1.9 ms), unpack_weight_grad x47 (0.6 ms), ATen's multi-tensor Adam (0.9 ms), copies.  Here:
//   * parameters, gradients and both moments live in four flat fp32 buffers (packnet_sfm_b200/optim.py); a convolution weight
//     that feeds the tensor-core engine is STORED as [Cout][tap][kpad] (kpad = Cin rounded up to 64) -- exactly what the
//     weight-gradient kernel accumulates into, so its gradient lands in the flat gradient buffer without a re-layout;
//   * ONE launch updates every element (float4 x2 per thread) and, for those weights, writes the bf16 hi / lo forward tiles
//     [chunk][tap][rows_pad][128 B, SWIZZLE_128B] of the NEW value from registers: one thread owns 8 consecutive reduction
//     channels = one 16-byte group of a tile row, so reads and writes are full 32-byte / 16-byte segments;
//   * the data gradient reads the same forward tiles as MN-major operands (conv_engine.cu, BT), so nothing else is packed.
// Algorithmic bytes per step: 16 B read + 12 B written per parameter + 4 B of tiles = 32 B x 129.9 M = 4.2 GB (0.63 ms at
// the measured 6.57 TB/s).
#include "common.cuh"

#ifdef PN_EMULATE
#include <cmath>
#include <cstring>
// bf16 round-to-nearest-even for the host emulation (bit pattern as unsigned short)
static inline unsigned short ok_bf16_rn(float f) {
  unsigned u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x0040u);
  const unsigned lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (unsigned short)(u >> 16);
}
static inline float ok_bf16_to_float(unsigned short h) {
  const unsigned u = (unsigned)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
#else
#include <cuda_bf16.h>
__device__ __forceinline__ unsigned short ok_bf16_rn(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
__device__ __forceinline__ float ok_bf16_to_float(unsigned short h) { return __bfloat162float(__ushort_as_bfloat16(h)); }
#endif

namespace pn {
namespace optim {

constexpr int THREADS = 256;
constexpr int PER_THREAD = 8;
constexpr int BLOCK_ELEMS = THREADS * PER_THREAD;   // 2048 = PN_ADAM_BLOCK

// hyper[] slots (device memory: a captured CUDA graph sees the values of the replay, not of the capture)
enum { H_STEP = 0, H_BETA1 = 1, H_BETA2 = 2, H_EPS = 3, H_BC1 = 4, H_BC2_SQRT = 5, H_GROUP0 = 8 };   // lr, wd per group from 8

__global__ void adam_prep_kernel(float* hyper) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float t = hyper[H_STEP] + 1.0f;
    hyper[H_STEP] = t;
    hyper[H_BC1] = (float)(1.0 - pow((double)hyper[H_BETA1], (double)t));
    hyper[H_BC2_SQRT] = (float)sqrt(1.0 - pow((double)hyper[H_BETA2], (double)t));
  }
}

__device__ __forceinline__ uint32_t pack2(unsigned short a, unsigned short b) { return (uint32_t)a | ((uint32_t)b << 16); }

// UPDATE = false: only (re)write the tiles of the stored weights (after construction / load_state_dict)
template <bool UPDATE>
__global__ void __launch_bounds__(THREADS) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                            float* __restrict__ v, const int32_t* __restrict__ block_info,
                                                            const pn_adam_conv_seg* __restrict__ segs, const float* __restrict__ hyper,
                                                            uint8_t* __restrict__ pk_hi, uint8_t* __restrict__ pk_lo) {
  const int info = __ldg(block_info + blockIdx.x);
  const int seg = (info & 0xFFFF) - 1, grp = (info >> 16) & 0xFF;
  if (!UPDATE && seg < 0) return;
  const size_t e0 = (size_t)blockIdx.x * BLOCK_ELEMS + (size_t)threadIdx.x * PER_THREAD;
  float w[PER_THREAD];
  {
    const float4 a = *reinterpret_cast<const float4*>(p + e0), b = *reinterpret_cast<const float4*>(p + e0 + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  }
  if (UPDATE) {
    const float beta1 = __ldg(hyper + H_BETA1), beta2 = __ldg(hyper + H_BETA2), eps = __ldg(hyper + H_EPS);
    const float lr = __ldg(hyper + H_GROUP0 + 2 * grp), wd = __ldg(hyper + H_GROUP0 + 2 * grp + 1);
    const float step_size = lr / __ldg(hyper + H_BC1), bc2s = __ldg(hyper + H_BC2_SQRT);
    float gr[PER_THREAD], mm[PER_THREAD], vv[PER_THREAD];
    {
      const float4 a = *reinterpret_cast<const float4*>(g + e0), b = *reinterpret_cast<const float4*>(g + e0 + 4);
      gr[0] = a.x; gr[1] = a.y; gr[2] = a.z; gr[3] = a.w; gr[4] = b.x; gr[5] = b.y; gr[6] = b.z; gr[7] = b.w;
      const float4 c = *reinterpret_cast<const float4*>(m + e0), d = *reinterpret_cast<const float4*>(m + e0 + 4);
      mm[0] = c.x; mm[1] = c.y; mm[2] = c.z; mm[3] = c.w; mm[4] = d.x; mm[5] = d.y; mm[6] = d.z; mm[7] = d.w;
      const float4 e = *reinterpret_cast<const float4*>(v + e0), f = *reinterpret_cast<const float4*>(v + e0 + 4);
      vv[0] = e.x; vv[1] = e.y; vv[2] = e.z; vv[3] = e.w; vv[4] = f.x; vv[5] = f.y; vv[6] = f.z; vv[7] = f.w;
    }
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      // torch.optim.Adam (no amsgrad; weight_decay = L2 added to the gradient), the op order of ATen's fused kernel
      const float gi = gr[i] + wd * w[i];
      mm[i] = mm[i] + (1.0f - beta1) * (gi - mm[i]);                       // lerp(exp_avg, grad, 1 - beta1)
      vv[i] = beta2 * vv[i] + (1.0f - beta2) * gi * gi;
      const float denom = sqrtf(vv[i]) / bc2s + eps;
      w[i] = w[i] - step_size * (mm[i] / denom);
    }
    *reinterpret_cast<float4*>(p + e0) = make_float4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<float4*>(p + e0 + 4) = make_float4(w[4], w[5], w[6], w[7]);
    *reinterpret_cast<float4*>(m + e0) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(m + e0 + 4) = make_float4(mm[4], mm[5], mm[6], mm[7]);
    *reinterpret_cast<float4*>(v + e0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(v + e0 + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
  }
  if (seg < 0) return;
  // forward tiles of a stored convolution weight [Cout][tap][kpad]: element (co, tap, ci) -> chunk ci/64, row co,
  // 16-byte group ((ci%64)/8) ^ (co&7)   (the image pack_weight_kernel documents in conv_engine.cu)
  const pn_adam_conv_seg S = segs[seg];
  const size_t e = e0 - (size_t)S.offset;
  const int ci = (int)(e % (size_t)S.kpad);
  const size_t rt = e / (size_t)S.kpad;
  const int tap = (int)(rt % (size_t)S.taps), co = (int)(rt / (size_t)S.taps);
  if (co >= S.cout) return;   // padding of the segment up to the block size
  const int cc = ci >> 6, grp16 = (ci & 63) >> 3;
  const size_t o = (size_t)S.packed_offset + (((size_t)cc * S.taps + tap) * S.rows_pad + co) * 128u + (size_t)((grp16 ^ (co & 7)) << 4);
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned short h0 = ok_bf16_rn(w[2 * i]), h1 = ok_bf16_rn(w[2 * i + 1]);
    h[i] = pack2(h0, h1);
    l[i] = pack2(ok_bf16_rn(w[2 * i] - ok_bf16_to_float(h0)), ok_bf16_rn(w[2 * i + 1] - ok_bf16_to_float(h1)));
  }
  *reinterpret_cast<uint4*>(pk_hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(pk_lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace optim
}  // namespace pn

using namespace pn;

extern "C" int pn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t numel,
                            const int32_t* block_info, const pn_adam_conv_seg* segs, float* hyper, void* packed_hi,
                            void* packed_lo, int update, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(params && block_info && hyper && numel > 0 && numel % optim::BLOCK_ELEMS == 0, PN_ERR_BAD_ARGUMENT,
             "pn_adam_step: numel (%lld) must be a positive multiple of %d", (long long)numel, optim::BLOCK_ELEMS);
  PN_REQUIRE(!update || (grads && exp_avg && exp_avg_sq), PN_ERR_BAD_ARGUMENT, "pn_adam_step: update needs grads and both moments");
  PN_REQUIRE(aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq) && aligned16(packed_hi) &&
                 aligned16(packed_lo), PN_ERR_ALIGNMENT, "pn_adam_step: buffers must be 16-byte aligned");
  const int64_t nblocks = numel / optim::BLOCK_ELEMS;
  PN_REQUIRE(nblocks < (1ll << 31), PN_ERR_UNSUPPORTED, "pn_adam_step: too many elements");
  TraceScope ts(stream, "adam_step n%lld update%d", (long long)numel, update);
  if (update) {
    PN_LAUNCH(optim::adam_prep_kernel, 1, 32, 0, stream, hyper);
    PN_LAUNCH(optim::adam_flat_kernel<true>, (unsigned)nblocks, optim::THREADS, 0, stream, params, grads, exp_avg, exp_avg_sq, block_info,
              segs, hyper, static_cast<uint8_t*>(packed_hi), static_cast<uint8_t*>(packed_lo));
    count_launch(2);
  } else {
    PN_LAUNCH(optim::adam_flat_kernel<false>, (unsigned)nblocks, optim::THREADS, 0, stream, params, grads, exp_avg, exp_avg_sq,
              block_info, segs, hyper, static_cast<uint8_t*>(packed_hi), static_cast<uint8_t*>(packed_lo));
    count_launch(1);
  }
  return check_launch("adam_flat_kernel");
}
