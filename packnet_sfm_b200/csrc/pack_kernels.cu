// pack_kernels.cu -- STAGED (functional.set_pack_tiled / PN_PACK_TILED; off by default until measured on a B200): the
// per-step weight packing of the bf16 precisions through shared memory.
//
// The packed layout is the one conv_engine.cu documents (pack_weight_kernel): [cchunk][tap][rows_pad][64 bf16 = 128 bytes],
// row = output channel (fprop) / input channel (dgrad), the eight 16-byte groups of a row XOR-swizzled with (row & 7):
//   fprop  (transposed = 0): value(row = co, k = ci, tap) = w[co][ci][tap]
//   dgrad  (transposed = 1): value(row = ci, k = co, tap) = w[co][ci][taps - 1 - tap]
// hi = rn_bf16(v), lo = rn_bf16(v - hi); rows >= the real row count and k >= the real reduction size are zero.
// The default kernel is one thread per OUTPUT element: neighbouring threads read `taps` floats apart (fprop) or
// Cin*taps floats apart (dgrad) -- 1.94 ms per step for 0.5 GB of weights read twice and 1 GB written (0.3 ms at the HBM
// rate).  Here a CTA owns (R rows, one 64-wide reduction chunk, all taps): contiguous runs in (64*taps floats per row for
// fprop, R*taps floats per output channel for dgrad), a small shared tile with an odd pitch, and for every (tap, row) one
// full 128-byte line of hi and one of lo out.  Same bits as the default kernel (tests compare them).
#include "common.cuh"

#ifdef PN_EMULATE
// bf16 round-to-nearest-even for the host emulation (bit pattern as unsigned short)
static inline unsigned short pk_bf16_rn(float f) {
  unsigned u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x0040u);   // NaN stays NaN (quiet)
  const unsigned lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (unsigned short)(u >> 16);
}
static inline float pk_bf16_to_float(unsigned short h) {
  const unsigned u = (unsigned)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
#else
#include <cuda_bf16.h>
__device__ __forceinline__ unsigned short pk_bf16_rn(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
__device__ __forceinline__ float pk_bf16_to_float(unsigned short h) { return __bfloat162float(__ushort_as_bfloat16(h)); }
#endif

namespace pn {
namespace packk {

constexpr int KC = 64;   // bf16 elements per 128-byte row
constexpr int RB = 8;    // rows per CTA (the swizzle period)

struct PackParams {
  const float* w;          // OIHW
  unsigned short* hi;      // packed bf16 bit patterns
  unsigned short* lo;      // or nullptr
  int Cout, Cin, taps, rows_pad, transposed;
};

// grid: (rows_pad / RB, cchunks); shared: [RB][KC*taps] (fprop) or [KC][RB*taps | 1] (dgrad) floats
__global__ void __launch_bounds__(256) pack_weight_tiled_kernel(const PackParams P) {
  PN_DYNAMIC_SHARED(float, sm);
  const int taps = P.taps;
  const int rows = P.transposed ? P.Cin : P.Cout, kred = P.transposed ? P.Cout : P.Cin;
  const int r0 = blockIdx.x * RB, cc = blockIdx.y, k0 = cc * KC;
  const int nk = min(KC, kred - k0);           // valid reduction entries of this chunk
  const int nr = max(0, min(RB, rows - r0));   // valid rows of this block
  const int pitch = P.transposed ? ((RB * taps) | 1) : (KC * taps);
  if (!P.transposed) {
    // row r: w[r][k0 .. k0+nk)[all taps] is one contiguous run of nk*taps floats; tile element (r, k, tap) at r*pitch + k*taps + tap
    for (int rl = 0; rl < nr; ++rl) {
      const float* src = P.w + ((size_t)(r0 + rl) * P.Cin + k0) * taps;
      for (int i = threadIdx.x; i < nk * taps; i += blockDim.x) sm[rl * pitch + i] = __ldg(src + i);
    }
  } else {
    // reduction entry k (an output channel): w[k0+k][r0 .. r0+nr)[all taps] is one contiguous run of nr*taps floats;
    // tile element (k, r, tap) at k*pitch + r*taps + tap
    const int run = nr * taps;
    for (int i = threadIdx.x; i < nk * run; i += blockDim.x) {
      const int kl = i / run, j = i - kl * run;
      sm[kl * pitch + j] = __ldg(P.w + ((size_t)(k0 + kl) * P.Cin + r0) * taps + j);
    }
  }
  __syncthreads();
  // one 128-byte line (64 bf16) per (tap, row): thread <-> position e inside the line
  for (int i = threadIdx.x; i < taps * RB * KC; i += blockDim.x) {
    const int e = i % KC, rl = (i / KC) % RB, tap = i / (KC * RB);
    const int r = r0 + rl;
    const int grp = e >> 3, within = e & 7;
    const int kl = ((grp ^ (r & 7)) << 3) + within;   // logical reduction index (inside the chunk) stored at position e
    float v = 0.0f;
    if (rl < nr && kl < nk) v = P.transposed ? sm[kl * pitch + rl * taps + (taps - 1 - tap)] : sm[rl * pitch + kl * taps + tap];
    const size_t o = (((size_t)cc * taps + tap) * P.rows_pad + r) * KC + e;
    const unsigned short h = pk_bf16_rn(v);
    P.hi[o] = h;
    if (P.lo) P.lo[o] = pk_bf16_rn(v - pk_bf16_to_float(h));
  }
}

}  // namespace packk
}  // namespace pn

using namespace pn;

extern "C" int pn_conv2d_pack_weight_tiled(const float* w_oihw, void* w_packed, void* w_packed_lo, int cout, int cin, int ksize,
                                           int transposed, int rows_pad, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(w_oihw && w_packed && cout > 0 && cin > 0 && ksize > 0, PN_ERR_BAD_ARGUMENT, "pn_conv2d_pack_weight_tiled: bad argument");
  const int rows = transposed ? cin : cout, kred = transposed ? cout : cin;
  PN_REQUIRE(rows_pad >= rows && rows_pad % packk::RB == 0, PN_ERR_BAD_ARGUMENT, "pn_conv2d_pack_weight_tiled: rows_pad %d for %d rows",
             rows_pad, rows);
  const int taps = ksize * ksize;
  const size_t smem = (transposed ? (size_t)packk::KC * ((packk::RB * taps) | 1) : (size_t)packk::RB * packk::KC * taps) * sizeof(float);
  PN_REQUIRE(smem <= 100 * 1024, PN_ERR_UNSUPPORTED, "pn_conv2d_pack_weight_tiled: kernel size %d", ksize);
  const int cchunks = (kred + packk::KC - 1) / packk::KC;
  PN_REQUIRE(cchunks <= 65535, PN_ERR_UNSUPPORTED, "pn_conv2d_pack_weight_tiled: %d reduction chunks", cchunks);
  packk::PackParams P{};
  P.w = w_oihw; P.hi = static_cast<unsigned short*>(w_packed); P.lo = static_cast<unsigned short*>(w_packed_lo);
  P.Cout = cout; P.Cin = cin; P.taps = taps; P.rows_pad = rows_pad; P.transposed = transposed;
  PN_CUDA(cudaFuncSetAttribute(packk::pack_weight_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PN_LAUNCH(packk::pack_weight_tiled_kernel, dim3(rows_pad / packk::RB, cchunks), 256, smem, stream, P);
  count_launch();
  return check_launch("pack_weight_tiled_kernel");
}
