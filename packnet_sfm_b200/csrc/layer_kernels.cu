// layer_kernels.cu -- the HBM-bound pieces of the pack / unpack blocks (NHWC fp32), sm_100a.
//
//   * feature stencil: the Conv3d(1->8, 3x3x3, pad 1) of PackLayerConv3d / UnpackLayerConv3d fused with
//     the surrounding data movement, so neither the space-to-depth tensor nor a separately shuffled copy
//     ever exists in HBM:
//        pack  : x[B,2H,2W,C] --(space-to-depth read)--> 27-tap stencil over (4C,H,W) --> Y[B,H,W,8*4C]
//                layers01.py:126-148 (packing), :236-237,:241-245 (conv3d + view)
//        unpack: u[B,H,W,Cu] --> stencil over (Cu,H,W) --(depth-to-space write)--> out[B,2H,2W,2Cu]
//                layers01.py:274-276,:281-285 (conv3d + view + PixelShuffle), written straight into the
//                decoder's concat buffer (channel stride/offset) -- torch.cat (PackNet01.py:138-175) by pointer
//     and their backward (data gradient, conv3d weight/bias gradient).
//   * GroupNorm(16) + ELU (layers01.py:31-32,37 / :61-62,72): statistics pass + fused apply pass, and backward.
//   * tf32 residual outputs (x - trunc_tf32(x)) are emitted by the producers for the tf32x3 GEMM path.
#include <atomic>
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#ifdef PN_EMULATE
#include <cmath>
#include <cstring>
#else
#include <cuda_bf16.h>
#include <cooperative_groups.h>
#endif

namespace pn {
namespace layers {

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---------------------------------------------------------------------------------------------------
// feature stencil
// ---------------------------------------------------------------------------------------------------
// "virtual" input volume V[d][h][w], d in [0,D): pack: D = 4C, V[4c+2i+j][h][w] = x[2h+i][2w+j][c];
//                                               unpack: D = Cu, V[d][h][w] = u[h][w][d]
// output feature (f, d, h, w), f in [0,8): pack: Y[h][w][f*D + d]; unpack: v = f*D+d -> out[2h+i][2w+j][v/4], i=(v%4)/2, j=v%2
struct StencilParams {
  int B, H, W, D;          // volume dims (H, W are the LOW resolution for both pack and unpack)
  int C;                   // pack: input channels (D = 4C); unpack: Cu (D = Cu)
  int tw;                  // output pixels per CTA along w
  int out_cstride;         // channels per output pixel in the destination buffer
  int out_coffset;         // first channel written
  const float* in;         // pack: x [B,2H,2W,C]; unpack: u [B,H,W,Cu]
  const float* w3;         // [8][27]
  const float* b3;         // [8]
  float* out;              // pack: [B,H,W,out_cstride]; unpack: [B,2H,2W,out_cstride]
  float* out_lo;           // optional tf32 residual of out (same layout) or nullptr
};

template <bool PACK>
__device__ __forceinline__ float load_volume(const StencilParams& P, int b, int d, int h, int w) {
  if (PACK) {
    const int c = d >> 2, i = (d >> 1) & 1, j = d & 1;
    return __ldg(P.in + (((size_t)b * 2 * P.H + (2 * h + i)) * 2 * P.W + (2 * w + j)) * P.C + c);
  } else {
    return __ldg(P.in + (((size_t)b * P.H + h) * P.W + w) * P.C + d);
  }
}

// Stage the 3 x (tw+2) x (D+2) zero-padded neighbourhood of the virtual volume into shared memory.  One warp takes
// one (row, column[, sub-pixel]) combination at a time -- the index arithmetic is warp-uniform -- and its lanes run
// over the contiguous channels of ONE source pixel, so every global load is coalesced.
template <bool PACK>
__device__ __forceinline__ void stage_volume(const float* __restrict__ in, int b, int H, int W, int C, int D, int h, int w0,
                                             int TWP, float* __restrict__ s_v) {
  const int DP = D + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int ncombo = 3 * TWP * (PACK ? 4 : 1);
  for (int cb = warp; cb < ncombo; cb += nwarps) {
    const int ij = PACK ? (cb & 3) : 0;
    const int rc = PACK ? (cb >> 2) : cb;
    const int cw = rc % TWP, r = rc / TWP;
    const int ww = w0 + cw - 1, hh = h + r - 1;
    float* row = s_v + (r * TWP + cw) * DP;
    const bool inside = (ww >= 0) && (ww < W) && (hh >= 0) && (hh < H);
    if (ij == 0) {
      if (lane == 0) row[0] = 0.0f;
      if (lane == 1) row[DP - 1] = 0.0f;
    }
    if (PACK) {
      const int i = ij >> 1, j = ij & 1;
      const float* src = in + (((size_t)b * 2 * H + (2 * hh + i)) * 2 * W + (2 * ww + j)) * C;
      for (int c = lane; c < C; c += 32) row[1 + 4 * c + ij] = inside ? __ldg(src + c) : 0.0f;
    } else {
      const float* src = in + (((size_t)b * H + hh) * W + ww) * C;
      for (int d = lane; d < D; d += 32) row[1 + d] = inside ? __ldg(src + d) : 0.0f;
    }
  }
}

// smem: s_v[3][tw+2][D+2] (zero padded in all three dims), s_w[8*27 + 8]
template <bool PACK>
__global__ void __launch_bounds__(256) stencil_fwd_kernel(const StencilParams P) {
  PN_DYNAMIC_SHARED_PLAIN(float, sm);
  const int D = P.D, DP = D + 2, TWP = P.tw + 2;
  float* s_v = sm;
  float* s_w = sm + 3 * TWP * DP;
  const int w0 = blockIdx.x * P.tw, h = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < 8 * 27 + 8; i += blockDim.x) s_w[i] = (i < 216) ? P.w3[i] : P.b3[i - 216];
  stage_volume<PACK>(P.in, b, P.H, P.W, P.C, D, h, w0, TWP, s_v);
  __syncthreads();
  // one (pixel, d) item per thread iteration -> 8 features from 27 shared-memory reads
  const int items = P.tw * D;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int d = it % D, pw = it / D;
    const int w = w0 + pw;
    if (w >= P.W) continue;
    float acc[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) acc[f] = s_w[216 + f];
#pragma unroll
    for (int dz = 0; dz < 3; ++dz)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float v = s_v[(dy * TWP + pw + dx) * DP + d + dz];
          const int t = (dz * 3 + dy) * 3 + dx;
#pragma unroll
          for (int f = 0; f < 8; ++f) acc[f] = fmaf(s_w[f * 27 + t], v, acc[f]);
        }
    if (PACK) {
      float* o = P.out + (((size_t)b * P.H + h) * P.W + w) * P.out_cstride + P.out_coffset + d;
      float* ol = P.out_lo ? P.out_lo + (((size_t)b * P.H + h) * P.W + w) * P.out_cstride + P.out_coffset + d : nullptr;
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        o[(size_t)f * D] = acc[f];
        if (ol) ol[(size_t)f * D] = acc[f] - tf32_trunc(acc[f]);
      }
    } else {
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        const int v = f * D + d, co = v >> 2, i = (v >> 1) & 1, j = v & 1;
        const size_t o = (((size_t)b * 2 * P.H + (2 * h + i)) * 2 * P.W + (2 * w + j)) * P.out_cstride + P.out_coffset + co;
        P.out[o] = acc[f];
        if (P.out_lo) P.out_lo[o] = acc[f] - tf32_trunc(acc[f]);
      }
    }
  }
}

// backward: g = dL/d(out features) in the forward's OUTPUT layout; produces dL/d(in) in the forward's INPUT
// layout, and accumulates dL/dw3 [8][27], dL/db3 [8] (atomicAdd into zeroed buffers).
struct StencilBwdParams {
  int B, H, W, D, C, tw;
  int g_cstride, g_coffset;  // layout of g (same meaning as out_cstride/out_coffset of the forward)
  const float* in;           // forward input (for the weight gradient)
  const float* g;            // gradient w.r.t. the forward output
  const float* w3;
  float* gin;                // gradient w.r.t. the forward input (same layout as `in`)
  float* gw3;                // [8*27] accumulated
  float* gb3;                // [8] accumulated
};

template <bool PACK>
__device__ __forceinline__ float load_gout(const StencilBwdParams& P, int b, int f, int d, int h, int w) {
  if (PACK) {
    return __ldg(P.g + (((size_t)b * P.H + h) * P.W + w) * P.g_cstride + P.g_coffset + (size_t)f * P.D + d);
  } else {
    const int v = f * P.D + d, co = v >> 2, i = (v >> 1) & 1, j = v & 1;
    return __ldg(P.g + (((size_t)b * 2 * P.H + (2 * h + i)) * 2 * P.W + (2 * w + j)) * P.g_cstride + P.g_coffset + co);
  }
}

// smem: s_v[3][tw+2][D+2] (forward input neighbourhood), s_g[3][tw+2][D+2] (one feature plane of g, with halo),
//       s_w[216], s_red[256]
template <bool PACK>
__global__ void __launch_bounds__(256) stencil_bwd_kernel(const StencilBwdParams P) {
  PN_DYNAMIC_SHARED_PLAIN(float, sm);
  const int D = P.D, DP = D + 2, TWP = P.tw + 2;
  float* s_v = sm;
  float* s_g = s_v + 3 * TWP * DP;
  float* s_w = s_g + 3 * TWP * DP;
  float* s_red = s_w + 216;
  float* s_gh = s_red + 8 * 28 + 32;  // unpack only: the high-res gradient neighbourhood [6][2*TWP][2*D], loaded once
  const int w0 = blockIdx.x * P.tw, h = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < 216; i += blockDim.x) s_w[i] = P.w3[i];
  const int total = 3 * TWP * DP;
  if (!PACK) {
    // g lives at high resolution with 2*D channels per pixel: stage rows 2(h-1)..2(h+1)+1, cols 2(w0-1)..2(w0+tw)+1
    // pixel-major (coalesced), then every feature plane f is re-indexed out of shared memory
    const int CG = 2 * D, GW = 2 * TWP;
    const int gtotal = 6 * GW * CG;
    for (int idx = threadIdx.x; idx < gtotal; idx += blockDim.x) {
      const int c = idx % CG, gx = (idx / CG) % GW, gy = idx / (CG * GW);
      const int yy = 2 * (h - 1) + gy, xx = 2 * (w0 - 1) + gx;
      float v = 0.0f;
      if (yy >= 0 && yy < 2 * P.H && xx >= 0 && xx < 2 * P.W)
        v = __ldg(P.g + (((size_t)b * 2 * P.H + yy) * 2 * P.W + xx) * P.g_cstride + P.g_coffset + c);
      s_gh[idx] = v;
    }
  }
  (void)total; (void)s_v;   // the weight gradient (which needs the forward input) lives in stencil_wgrad_kernel
  const int items = P.tw * D;
  constexpr int MAXI = 8;  // items per thread kept in registers (tw*D <= 256*MAXI is enforced by the host)
  float gin_acc[MAXI];
#pragma unroll
  for (int k = 0; k < MAXI; ++k) gin_acc[k] = 0.0f;
  for (int f = 0; f < 8; ++f) {
    __syncthreads();
    {  // feature plane f of g with its halo: one warp per (row, column), lanes over the depth
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
      for (int rc = warp; rc < 3 * TWP; rc += nwarps) {
        const int cw = rc % TWP, r = rc / TWP;
        const int ww = w0 + cw - 1, hh = h + r - 1;
        float* row = s_g + (r * TWP + cw) * DP;
        const bool inside = (ww >= 0) && (ww < P.W) && (hh >= 0) && (hh < P.H);
        if (lane == 0) row[0] = 0.0f;
        if (lane == 1) row[DP - 1] = 0.0f;
        if (PACK) {
          const float* src = P.g + (((size_t)b * P.H + hh) * P.W + ww) * P.g_cstride + P.g_coffset + (size_t)f * D;
          for (int d = lane; d < D; d += 32) row[1 + d] = inside ? __ldg(src + d) : 0.0f;
        } else {
          for (int d = lane; d < D; d += 32) {
            const int vv = f * D + d, co = vv >> 2, i = (vv >> 1) & 1, j = vv & 1;
            row[1 + d] = inside ? s_gh[((2 * r + i) * (2 * TWP) + (2 * cw + j)) * (2 * D) + co] : 0.0f;
          }
        }
      }
    }
    __syncthreads();
    int slot = 0;
    for (int it = threadIdx.x; it < items; it += blockDim.x, ++slot) {
      const int d = it % D, pw = it / D;
      if (w0 + pw >= P.W) continue;
      float a = 0.0f;
#pragma unroll
      for (int dz = 0; dz < 3; ++dz)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int t = (dz * 3 + dy) * 3 + dx;
            // data gradient: sum_t w[t] * g(centre - offset)  (transpose of the forward stencil)
            a = fmaf(s_w[f * 27 + t], s_g[((2 - dy) * TWP + pw + (2 - dx)) * DP + d + (2 - dz)], a);
          }
      if (slot < MAXI) gin_acc[slot] += a;
    }
  }
  // write the data gradient in the forward INPUT layout, through shared memory so the stores are pixel-major
  __syncthreads();
  float* s_out = s_g;  // tw * D floats (<= 3*(tw+2)*(D+2))
  {
    int slot = 0;
    for (int it = threadIdx.x; it < items; it += blockDim.x, ++slot)
      if (slot < MAXI) s_out[it] = gin_acc[slot];
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int ncombo = P.tw * (PACK ? 4 : 1);
    for (int cb = warp; cb < ncombo; cb += nwarps) {
      const int ij = PACK ? (cb & 3) : 0, pw = PACK ? (cb >> 2) : cb;
      const int w = w0 + pw;
      if (w >= P.W) continue;
      if (PACK) {
        const int i = ij >> 1, j = ij & 1;
        float* dst = P.gin + (((size_t)b * 2 * P.H + (2 * h + i)) * 2 * P.W + (2 * w + j)) * P.C;
        for (int c = lane; c < P.C; c += 32) dst[c] = s_out[pw * D + 4 * c + ij];
      } else {
        float* dst = P.gin + (((size_t)b * P.H + h) * P.W + w) * P.C;
        for (int d = lane; d < D; d += 32) dst[d] = s_out[pw * D + d];
      }
    }
  }
}

// Conv3d weight / bias gradient: dW[f][t] = sum over (b, d, h, w) of g[f][d][h][w] * V[d+dz-1][h+dy-1][w+dx-1], db[f] = sum g.
// One CTA walks a strided set of (sample, row) pairs and all pixel tiles of each row; warp f owns feature f and keeps its
// 27 + 1 partial sums in registers across the whole walk, so there is ONE warp reduction and 28 atomics per warp at the
// very end (the previous per-tile block reductions + 224 contended atomics per tile dominated the backward).
// smem: s_v[3][tw+2][D+2] (forward input with halo), s_gc[8][tw][D] (g at the tile's pixels, all 8 features)
template <bool PACK>
__global__ void __launch_bounds__(256) stencil_wgrad_kernel(const StencilBwdParams P) {
  PN_DYNAMIC_SHARED_PLAIN(float, sm);
  const int D = P.D, DP = D + 2, TWP = P.tw + 2;
  float* s_v = sm;
  float* s_gc = s_v + 3 * TWP * DP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = warp;  // 8 warps <-> 8 features
  float wacc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) wacc[t] = 0.0f;
  float bacc = 0.0f;
  const int rows = P.B * P.H;
  const int wtiles = (P.W + P.tw - 1) / P.tw;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / P.H, h = row % P.H;
    for (int wt = 0; wt < wtiles; ++wt) {
      const int w0 = wt * P.tw;
      __syncthreads();
      stage_volume<PACK>(P.in, b, P.H, P.W, P.C, D, h, w0, TWP, s_v);
      if (PACK) {
        // g[b][h][w][f*D + d]: one warp per (pixel, feature), lanes over the depth
        for (int cb = warp; cb < P.tw * 8; cb += 8) {
          const int ff = cb & 7, pw = cb >> 3;
          const int w = w0 + pw;
          const float* src = P.g + (((size_t)b * P.H + h) * P.W + w) * P.g_cstride + P.g_coffset + (size_t)ff * D;
          float* dst = s_gc + ((size_t)ff * P.tw + pw) * D;
          for (int d = lane; d < D; d += 32) dst[d] = (w < P.W) ? __ldg(src + d) : 0.0f;
        }
      } else {
        // g lives at high resolution: pixel (2h+i, 2w+j), channel co holds feature value v = 4*co + 2*i + j = f*D + d
        for (int cb = warp; cb < P.tw * 4; cb += 8) {
          const int ij = cb & 3, pw = cb >> 2;
          const int i = ij >> 1, j = ij & 1, w = w0 + pw;
          const float* src = P.g + (((size_t)b * 2 * P.H + (2 * h + i)) * 2 * P.W + (2 * w + j)) * P.g_cstride + P.g_coffset;
          for (int co = lane; co < 2 * D; co += 32) {
            const int v = 4 * co + ij, ff = v / D, d = v - ff * D;
            s_gc[((size_t)ff * P.tw + pw) * D + d] = (w < P.W) ? __ldg(src + co) : 0.0f;
          }
        }
      }
      __syncthreads();
      const float* gc_f = s_gc + (size_t)f * P.tw * D;
      for (int it = lane; it < P.tw * D; it += 32) {
        const int d = it % D, pw = it / D;
        const float gc = gc_f[it];
        bacc += gc;
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              wacc[(dz * 3 + dy) * 3 + dx] = fmaf(gc, s_v[(dy * TWP + pw + dx) * DP + d + dz], wacc[(dz * 3 + dy) * 3 + dx]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 28; ++t) {
    float v = (t < 27) ? wacc[t < 27 ? t : 0] : bacc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      if (t < 27) atomicAdd(P.gw3 + f * 27 + t, v);
      else atomicAdd(P.gb3 + f, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// feature stencil, register-tiled (D % 8 == 0): the production path
// ---------------------------------------------------------------------------------------------------
// The kernels above spend one or two shared-memory reads per FMA (weights and neighbours both come from shared
// memory), i.e. they run at the LDS rate, ~1/8 of the fp32 pipe (r01 step breakdown: 14.6 of 58 ms).  Here a thread
// owns a THREAD TILE of 8 consecutive depths of one pixel (x 8 features forward, x 2 rows backward, x 2 features
// for the weight gradient): a neighbour column is fetched as 10 floats (two LDS.128 + two LDS.32) and reused by
// 24..192 FMAs, weights sit in registers or arrive as broadcast LDS.128.  Shared bytes per FMA drop from ~8 to <= 1.1.
//   staged layout: cell (r, c) of a tile holds depths [0, D) at s[(r*NC + c)*PITCH + 4 + d], PITCH = D + 4; the 4 front
//   floats are zero (index 3 is depth -1) and depth D of a cell is the (zero) front pad of the next cell; lanes map
//   to consecutive cells, so with PITCH = 4 (mod 32) words the LDS.128 of a quarter-warp hit 8 different bank groups.
constexpr int SPAD = 4;

// rows [h0, h0+nr) x cols [w0, w0+nc) of a LOW-resolution H x W grid into the staged layout; pixels outside the grid
// are zero.  DIRECT: element d of pixel (hh, ww) is src[((b*H + hh)*W + ww)*pixstride + chan0 + d] (float4 loads).
// S2D (space-to-depth gather): element d comes from the 2x pixel (2hh + i, 2ww + j), ij = d & 3, channel chan0 + (d >> 2).
template <bool S2D>
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int b, int H, int W, int pixstride, int chan0, int D,
                                           int h0, int nr, int w0, int nc, float* __restrict__ s) {
  const int PITCH = D + SPAD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ncell = nr * nc;
  for (int cell = warp; cell < ncell; cell += nwarps) {
    const int c = cell % nc, r = cell / nc;
    const int hh = h0 + r, ww = w0 + c;
    const bool inside = (hh >= 0) && (hh < H) && (ww >= 0) && (ww < W);
    float* col = s + (size_t)cell * PITCH;
    if (lane == 0) *reinterpret_cast<float4*>(col) = z4;
    float4* dst = reinterpret_cast<float4*>(col + SPAD);
    if (S2D) {
      const size_t rowstride = (size_t)2 * W * pixstride;
      const float* p00 = src + (((size_t)b * 2 * H + 2 * hh) * 2 * W + 2 * ww) * pixstride + chan0;
      for (int q = lane; q < (D >> 2); q += 32) {
        float4 v = z4;
        if (inside) {
          v.x = __ldg(p00 + q); v.y = __ldg(p00 + pixstride + q);
          v.z = __ldg(p00 + rowstride + q); v.w = __ldg(p00 + rowstride + pixstride + q);
        }
        dst[q] = v;
      }
    } else {
      const float4* p = reinterpret_cast<const float4*>(src + (((size_t)b * H + hh) * W + ww) * pixstride + chan0);
      for (int q = lane; q < (D >> 2); q += 32) dst[q] = inside ? __ldg(p + q) : z4;
    }
  }
  if (threadIdx.x == 0) *reinterpret_cast<float4*>(s + (size_t)ncell * PITCH) = z4;   // depth D of the last cell
}

// STAGED variant of stage_tile for the head convolution and register-tiled feature-stencil kernels (pn_set_tuning(PN_TUNE_STAGE_FLAT, 1); off by default until
// measured on a B200; the default instantiations compile exactly as before): the same
// shared-memory image, but the (cell, float4) items are spread over ALL threads of the CTA, eight loads in flight per thread
// (explicit load phase / store phase: left to `#pragma unroll` nvcc keeps each store right behind its load), instead of one
// cell per warp iteration.  With D = 64 the original keeps 16 of 32 lanes busy and issues one dependent 256-byte load per
// warp iteration: the head convolution stages 340 cells = 43 serial round trips to DRAM per warp, which is what its
// 0.22 ms at 192x640 (126 MB: 17 us at the HBM rate) amounts to; the unpack stencils (D = 32) keep 8 lanes busy.
// PLANES: the nr "rows" of the tile are nr channel windows of ONE grid row h0 (row r reads channels chan0 + r * chan_step): the
// eight feature planes of an output gradient staged by one call = one load phase (round 1 staged them with eight calls, i.e.
// eight serial DRAM round trips per work item of the weight-gradient / data-gradient kernels).
template <bool S2D, bool PLANES = false>
__device__ __forceinline__ void stage_tile_flat(const float* __restrict__ src, int b, int H, int W, int pixstride, int chan0, int D,
                                                int h0, int nr, int w0, int nc, float* __restrict__ s, int chan_step = 0,
                                                int start = 0) {
  constexpr int U = 8;   // items in flight per thread: all loads of a batch are issued before the first store
  const int PITCH = D + SPAD, dq = D >> 2;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ncell = nr * nc, total = ncell * dq;
  const int step = (int)blockDim.x;
  // item `it` = (cell, q) = (it / dq, it % dq), cell = (r, c) = (cell / nc, cell % nc): the divisions are done once, the
  // thread then walks its items (it, it + step, ...) with carries
  const int step_q = step % dq, step_cell = step / dq;
  const int step_c = step_cell % nc, step_r = step_cell / nc;
  // `start` (a multiple of blockDim.x): the items before it were brought in by flat_prefetch / flat_store
  int q = ((int)threadIdx.x + start) % dq, cell0 = ((int)threadIdx.x + start) / dq;
  int c = cell0 % nc, r = cell0 / nc;
  for (int base = threadIdx.x + start; base < total; base += step * U) {
    float4 v[U];
    int off[U];   // shared-memory float offset of the item's float4 (low bit set: first float4 of its cell), -1: no item
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = z4;
      off[u] = -1;
      if (base + u * step < total) {
        const int hh = PLANES ? h0 : h0 + r, ww = w0 + c;
        const int ch = PLANES ? chan0 + r * chan_step : chan0;
        off[u] = ((r * nc + c) * PITCH + SPAD + 4 * q) | (q == 0 ? 1 : 0);
        if ((hh >= 0) && (hh < H) && (ww >= 0) && (ww < W)) {
          if (S2D) {
            const size_t rowstride = (size_t)2 * W * pixstride;
            const float* p00 = src + (((size_t)b * 2 * H + 2 * hh) * 2 * W + 2 * ww) * pixstride + ch;
            v[u].x = __ldg(p00 + q); v[u].y = __ldg(p00 + pixstride + q);
            v[u].z = __ldg(p00 + rowstride + q); v[u].w = __ldg(p00 + rowstride + pixstride + q);
          } else {
            v[u] = __ldg(reinterpret_cast<const float4*>(src + (((size_t)b * H + hh) * W + ww) * pixstride + ch) + q);
          }
        }
        // advance (q, c, r) by `step` items
        q += step_q;
        int carry = 0;
        if (q >= dq) { q -= dq; carry = 1; }
        c += step_c + carry;
        r += step_r;
        if (c >= nc) { c -= nc; r += 1; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (off[u] >= 0) {
        const int o = off[u] & ~1;   // offsets are multiples of 4 floats
        *reinterpret_cast<float4*>(s + o) = v[u];
        if (off[u] & 1) *reinterpret_cast<float4*>(s + o - SPAD) = z4;   // the cell's front pad
      }
    }
  }
  if (threadIdx.x == 0) *reinterpret_cast<float4*>(s + (size_t)ncell * PITCH) = z4;   // depth D of the last cell
}


// ---- asynchronous staging (cp.async): global -> shared copies that need no registers and no load/store pairing ------------------
// A work item's tiles are copied into the OTHER half of a double buffer while the FMAs of the current one run; rows outside the
// grid are zero-filled by the copy itself (src-size 0).  The emulated host build performs the copy at issue time.
__device__ __forceinline__ void pn_cp_async16(float* smem_dst, const float* gsrc, bool valid) {
#ifdef PN_EMULATE
  if (valid) std::memcpy(smem_dst, gsrc, 16); else std::memset(smem_dst, 0, 16);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
#endif
}
__device__ __forceinline__ void pn_cp_async4(float* smem_dst, const float* gsrc, bool valid) {
#ifdef PN_EMULATE
  if (valid) std::memcpy(smem_dst, gsrc, 4); else std::memset(smem_dst, 0, 4);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
#endif
}
__device__ __forceinline__ void pn_cp_async_commit() {
#ifndef PN_EMULATE
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void pn_cp_async_wait() {
#ifndef PN_EMULATE
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// the zero cells of a staged region that no copy ever touches: front pad of every cell and depth D of the last one
__device__ __forceinline__ void stage_pads_zero(float* __restrict__ s, int D, int ncell) {
  const int PITCH = D + SPAD;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int cell = threadIdx.x; cell <= ncell; cell += blockDim.x) *reinterpret_cast<float4*>(s + (size_t)cell * PITCH) = z4;
}

// stage_tile_flat's region (same item order, addresses and shared-memory image, pads excluded: stage_pads_zero) as
// asynchronous copies
template <bool S2D, bool PLANES>
__device__ __forceinline__ void stage_tile_async(const float* __restrict__ src, int b, int H, int W, int pixstride, int chan0, int D,
                                                 int h0, int nr, int w0, int nc, float* __restrict__ s, int chan_step = 0) {
  const int PITCH = D + SPAD, dq = D >> 2;
  const int total = nr * nc * dq, step = (int)blockDim.x;
  const int step_q = step % dq, step_cell = step / dq;
  const int step_c = step_cell % nc, step_r = step_cell / nc;
  int q = (int)threadIdx.x % dq, cell0 = (int)threadIdx.x / dq;
  int c = cell0 % nc, r = cell0 / nc;
  for (int it = threadIdx.x; it < total; it += step) {
    const int hh = PLANES ? h0 : h0 + r, ww = w0 + c;
    const int ch = PLANES ? chan0 + r * chan_step : chan0;
    float* dst = s + (size_t)(r * nc + c) * PITCH + SPAD + 4 * q;
    const bool inside = (hh >= 0) && (hh < H) && (ww >= 0) && (ww < W);
    if (S2D) {
      const size_t rowstride = (size_t)2 * W * pixstride;
      const float* p00 = inside ? src + (((size_t)b * 2 * H + 2 * hh) * 2 * W + 2 * ww) * pixstride + ch + q : src;
      pn_cp_async4(dst + 0, p00, inside);
      pn_cp_async4(dst + 1, inside ? p00 + pixstride : src, inside);
      pn_cp_async4(dst + 2, inside ? p00 + rowstride : src, inside);
      pn_cp_async4(dst + 3, inside ? p00 + rowstride + pixstride : src, inside);
    } else {
      pn_cp_async16(dst, inside ? src + (((size_t)b * H + hh) * W + ww) * pixstride + ch + 4 * q : src, inside);
    }
    q += step_q;
    int carry = 0;
    if (q >= dq) { q -= dq; carry = 1; }
    c += step_c + carry;
    r += step_r;
    if (c >= nc) { c -= nc; r += 1; }
  }
}

// Register prefetch of the FIRST U items per thread of a flat-staged region (same item order, addresses and shared-memory image
// as stage_tile_flat): flat_prefetch issues the loads -- typically for the NEXT work item / feature plane, right before the FMAs
// of the current one -- and flat_store writes them behind the barrier that ends those FMAs.  Regions with more than
// U * blockDim.x items finish with stage_tile_flat(..., start = U * blockDim.x).
template <int U>
struct FlatRegs { float4 v[U]; int off[U]; };

template <bool S2D, bool PLANES, int U>
__device__ __forceinline__ void flat_prefetch(FlatRegs<U>& R, const float* __restrict__ src, int b, int H, int W, int pixstride, int chan0,
                                              int D, int h0, int nr, int w0, int nc, int chan_step = 0) {
  const int PITCH = D + SPAD, dq = D >> 2, total = nr * nc * dq;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int it = (int)threadIdx.x + u * (int)blockDim.x;
    R.v[u] = z4;
    R.off[u] = -1;
    if (it < total) {
      const int q = it % dq, cell = it / dq, c = cell % nc, r = cell / nc;
      const int hh = PLANES ? h0 : h0 + r, ww = w0 + c;
      const int ch = PLANES ? chan0 + r * chan_step : chan0;
      R.off[u] = (cell * PITCH + SPAD + 4 * q) | (q == 0 ? 1 : 0);
      if ((hh >= 0) && (hh < H) && (ww >= 0) && (ww < W)) {
        if (S2D) {
          const size_t rowstride = (size_t)2 * W * pixstride;
          const float* p00 = src + (((size_t)b * 2 * H + 2 * hh) * 2 * W + 2 * ww) * pixstride + ch;
          R.v[u].x = __ldg(p00 + q); R.v[u].y = __ldg(p00 + pixstride + q);
          R.v[u].z = __ldg(p00 + rowstride + q); R.v[u].w = __ldg(p00 + rowstride + pixstride + q);
        } else {
          R.v[u] = __ldg(reinterpret_cast<const float4*>(src + (((size_t)b * H + hh) * W + ww) * pixstride + ch) + q);
        }
      }
    }
  }
}

template <int U>
__device__ __forceinline__ void flat_store(const FlatRegs<U>& R, float* __restrict__ s, int D, int ncell) {
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (R.off[u] >= 0) {
      const int o = R.off[u] & ~1;
      *reinterpret_cast<float4*>(s + o) = R.v[u];
      if (R.off[u] & 1) *reinterpret_cast<float4*>(s + o - SPAD) = z4;   // the cell's front pad
    }
  }
  if (threadIdx.x == 0) *reinterpret_cast<float4*>(s + (size_t)ncell * (D + SPAD)) = z4;   // depth D of the last cell
}

// the 10 staged floats around depths [d0, d0+8) of one cell: vv[j] = depth d0 - 1 + j
__device__ __forceinline__ void load_col10(const float* __restrict__ col, float vv[10]) {
  const float4 a = *reinterpret_cast<const float4*>(col);
  const float4 c = *reinterpret_cast<const float4*>(col + 4);
  vv[0] = col[-1];
  vv[1] = a.x; vv[2] = a.y; vv[3] = a.z; vv[4] = a.w;
  vv[5] = c.x; vv[6] = c.y; vv[7] = c.z; vv[8] = c.w;
  vv[9] = col[8];
}

// forward.  smem: s_v[3][tw+2][PITCH] + 4, s_w[27][8], s_b[8]
template <bool PACK, bool FLAT = false>
__global__ void __launch_bounds__(256, 2) stencil_fwd8_kernel(const StencilParams P) {
  PN_DYNAMIC_SHARED(float, sm);
  const int D = P.D, PITCH = D + SPAD, TWP = P.tw + 2;
  float* s_v = sm;
  float* s_w = sm + 3 * TWP * PITCH + 4;
  const int w0 = blockIdx.x * P.tw, h = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < 224; i += blockDim.x) {
    if (i < 216) { const int t = i >> 3, f = i & 7; s_w[i] = __ldg(P.w3 + f * 27 + t); }
    else s_w[i] = __ldg(P.b3 + (i - 216));
  }
  if (FLAT) stage_tile_flat<PACK>(P.in, b, P.H, P.W, P.C, 0, D, h - 1, 3, w0 - 1, TWP, s_v);
  else      stage_tile<PACK>(P.in, b, P.H, P.W, P.C, 0, D, h - 1, 3, w0 - 1, TWP, s_v);
  __syncthreads();
  const int ntiles = P.tw * (D >> 3);
  for (int it = threadIdx.x; it < ntiles; it += blockDim.x) {
    // pack: consecutive lanes = consecutive pixels (each thread stores whole 32-byte runs of its own pixel).  unpack: the
    // depth-to-space store sends a thread's 8 depths to 2 channels x 4 pixels, so consecutive lanes take consecutive depth
    // blocks of ONE pixel: D/8 lanes together write D/4 contiguous channels (32 bytes at D = 32: whole sectors; the
    // pixel-major order wrote 4-byte pieces 512 bytes apart -- r02n: unpack1 forward 189 us for 126 MB).  Shared-memory reads
    // stay conflict-free: a quarter-warp's float4 columns start at (lane / (D/8)) * (D + 4) + (lane % (D/8)) * 8 floats.
    const int nd = D >> 3;
    const int pw = PACK ? it % P.tw : it / nd, d0 = (PACK ? it / P.tw : it % nd) << 3;
    const int w = w0 + pw;
    if (w >= P.W) continue;
    float acc[8][8];
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      const float bf = s_w[216 + f];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[f][k] = bf;
    }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float vv[10];
        load_col10(s_v + (size_t)(dy * TWP + pw + dx) * PITCH + SPAD + d0, vv);
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
          const int t = (dz * 3 + dy) * 3 + dx;
          const float4 wa = *reinterpret_cast<const float4*>(s_w + t * 8);
          const float4 wb = *reinterpret_cast<const float4*>(s_w + t * 8 + 4);
          const float wf[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int f = 0; f < 8; ++f)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[f][k] = fmaf(wf[f], vv[k + dz], acc[f][k]);
        }
      }
    if (PACK) {
      const size_t o = (((size_t)b * P.H + h) * P.W + w) * P.out_cstride + P.out_coffset + d0;
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        float4* dst = reinterpret_cast<float4*>(P.out + o + (size_t)f * D);
        dst[0] = make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
        dst[1] = make_float4(acc[f][4], acc[f][5], acc[f][6], acc[f][7]);
        if (P.out_lo) {
          float4* dl = reinterpret_cast<float4*>(P.out_lo + o + (size_t)f * D);
          dl[0] = make_float4(acc[f][0] - tf32_trunc(acc[f][0]), acc[f][1] - tf32_trunc(acc[f][1]),
                              acc[f][2] - tf32_trunc(acc[f][2]), acc[f][3] - tf32_trunc(acc[f][3]));
          dl[1] = make_float4(acc[f][4] - tf32_trunc(acc[f][4]), acc[f][5] - tf32_trunc(acc[f][5]),
                              acc[f][6] - tf32_trunc(acc[f][6]), acc[f][7] - tf32_trunc(acc[f][7]));
        }
      }
    } else {
      // v = f*D + d0 + k -> channel v >> 2 at output pixel (2h + ((v >> 1) & 1), 2w + (v & 1)); d0 and D are multiples of 8, so
      // depths k and k + 4 of a thread are the adjacent channels co0, co0 + 1 of the same pixel: one 8-byte store
      const bool vec2 = ((P.out_cstride | P.out_coffset) & 1) == 0;
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        const int co0 = (f * D + d0) >> 2;
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) {
          const int i = ij >> 1, j = ij & 1;
          const size_t o = (((size_t)b * 2 * P.H + (2 * h + i)) * 2 * P.W + (2 * w + j)) * P.out_cstride + P.out_coffset + co0;
          const float v0 = acc[f][ij], v1 = acc[f][4 + ij];
          if (vec2) {
            *reinterpret_cast<float2*>(P.out + o) = make_float2(v0, v1);
            if (P.out_lo) *reinterpret_cast<float2*>(P.out_lo + o) = make_float2(v0 - tf32_trunc(v0), v1 - tf32_trunc(v1));
          } else {
            P.out[o] = v0; P.out[o + 1] = v1;
            if (P.out_lo) { P.out_lo[o] = v0 - tf32_trunc(v0); P.out_lo[o + 1] = v1 - tf32_trunc(v1); }
          }
        }
      }
    }
  }
}

// data gradient:  gin[d][h][w] = sum_f sum_{dz,dy,dx} W[f][dz][dy][dx] * g[f][d-dz+1][h-dy+1][w-dx+1]
// One CTA owns TH rows x tw columns; feature planes of g are staged one at a time WITH their halo ((TH+2) x (tw+2) cells)
// and a thread tile is 2 rows x 1 column x 8 depths (16 accumulators, up to MAXT tiles per thread), so the 4 x 3 staged
// cells around it feed 432 FMAs.  smem: s_g[(TH+2)][(tw+2)][PITCH] + 4 (reused as the output transpose buffer), s_w[216]
struct StencilBwd8Params {
  StencilBwdParams p;
  int th;   // rows per CTA (even)
};

template <bool PACK, int MAXT, bool FLAT = false, bool PRE = false, bool ASYNC = false>
__global__ void __launch_bounds__(256) stencil_bwd8_kernel(const StencilBwd8Params Q) {
  const StencilBwdParams& P = Q.p;
  PN_DYNAMIC_SHARED(float, sm);
  const int D = P.D, PITCH = D + SPAD, TWP = P.tw + 2, TH = Q.th;
  const size_t g_floats = (size_t)(TH + 2) * TWP * PITCH + 4;
  float* s_g = sm;
  float* s_w = sm + (ASYNC ? 2 : 1) * g_floats;     // ASYNC: two plane buffers (cp.async of plane f+1 during the FMAs of plane f)
  const int w0 = blockIdx.x * P.tw, h0 = blockIdx.y * TH, b = blockIdx.z;
  for (int i = threadIdx.x; i < 216; i += blockDim.x) s_w[i] = __ldg(P.w3 + i);
  const int ntiles = (TH >> 1) * P.tw * (D >> 3);
  float acc[MAXT][2][8];
#pragma unroll
  for (int m = 0; m < MAXT; ++m)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[m][rr][k] = 0.0f;
  // FLAT: the first UP items per thread of feature plane f+1 are loaded into registers before the FMAs of plane f (eight
  // serial load phases per CTA otherwise); a plane with more items finishes behind the barrier as before.
  constexpr int UP = 8;
  FlatRegs<UP> RP;
  const int plane_items = (TH + 2) * TWP * (D >> 2);
  auto prefetch = [&](int f) {
    if (PACK) flat_prefetch<false, false, UP>(RP, P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + f * D, D, h0 - 1, TH + 2, w0 - 1, TWP);
    else      flat_prefetch<true, false, UP>(RP, P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + ((f * D) >> 2), D, h0 - 1, TH + 2, w0 - 1, TWP);
  };
  if (PRE) prefetch(0);
  auto issue = [&](int f, int k) {
    if (PACK) stage_tile_async<false, false>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + f * D, D, h0 - 1, TH + 2, w0 - 1, TWP, sm + (size_t)k * g_floats);
    else      stage_tile_async<true, false>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + ((f * D) >> 2), D, h0 - 1, TH + 2, w0 - 1, TWP, sm + (size_t)k * g_floats);
    pn_cp_async_commit();
  };
  if (ASYNC) {
    stage_pads_zero(sm, D, (TH + 2) * TWP);
    stage_pads_zero(sm + g_floats, D, (TH + 2) * TWP);
    issue(0, 0);
  }
  for (int f = 0; f < 8; ++f) {
    if (ASYNC) {
      // buffer (f + 1) & 1 was read by plane f - 1, whose iteration ended with a barrier
      if (f + 1 < 8) { issue(f + 1, (f + 1) & 1); pn_cp_async_wait<1>(); }
      else pn_cp_async_wait<0>();
      s_g = sm + (size_t)(f & 1) * g_floats;
    } else {
      __syncthreads();   // the previous plane is consumed (and s_w is visible)
    }
    if (ASYNC) {
    } else if (FLAT && !PRE) {
      if (PACK) stage_tile_flat<false>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + f * D, D, h0 - 1, TH + 2, w0 - 1, TWP, s_g);
      else      stage_tile_flat<true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + ((f * D) >> 2), D, h0 - 1, TH + 2, w0 - 1, TWP, s_g);
    } else if (PRE) {
      flat_store<UP>(RP, s_g, D, (TH + 2) * TWP);
      if (plane_items > UP * (int)blockDim.x) {
        if (PACK) stage_tile_flat<false>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + f * D, D, h0 - 1, TH + 2, w0 - 1, TWP, s_g, 0, UP * (int)blockDim.x);
        else      stage_tile_flat<true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + ((f * D) >> 2), D, h0 - 1, TH + 2, w0 - 1, TWP, s_g, 0, UP * (int)blockDim.x);
      }
    } else {
      if (PACK) stage_tile<false>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + f * D, D, h0 - 1, TH + 2, w0 - 1, TWP, s_g);
      else      stage_tile<true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + ((f * D) >> 2), D, h0 - 1, TH + 2, w0 - 1, TWP, s_g);
    }
    __syncthreads();
    if (PRE && f + 1 < 8) prefetch(f + 1);     // in flight during the FMAs of this plane
    float wr[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) wr[t] = s_w[f * 27 + t];
#pragma unroll
    for (int m = 0; m < MAXT; ++m) {
      const int it = threadIdx.x + m * 256;
      if (it < ntiles) {
        const int pw = it % P.tw, rest = it / P.tw;
        const int r0 = (rest % (TH >> 1)) << 1, d0 = (rest / (TH >> 1)) << 3;
#pragma unroll
        for (int gr = 0; gr < 4; ++gr)
#pragma unroll
          for (int gc = 0; gc < 3; ++gc) {
            float gg[10];   // gg[j] = g depth d0 - 1 + j of staged cell (r0 + gr, pw + gc)
            load_col10(s_g + (size_t)((r0 + gr) * TWP + pw + gc) * PITCH + SPAD + d0, gg);
            const int dx = 2 - gc;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int dy = rr + 2 - gr;     // staged row r0 + gr is output row (r0 + rr) - dy + 1 (+1 halo)
              if (dy >= 0 && dy <= 2) {
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                  const float wv = wr[(dz * 3 + dy) * 3 + dx];
#pragma unroll
                  for (int k = 0; k < 8; ++k) acc[m][rr][k] = fmaf(wv, gg[k - dz + 2], acc[m][rr][k]);
                }
              }
            }
          }
      }
    }
    if (ASYNC) __syncthreads();   // plane f is consumed before the copies of plane f + 2 refill its buffer
  }
  // transpose through shared memory so that the stores run along the channels of one pixel
  __syncthreads();
  float* s_out = sm;   // [TH][tw][D] (the first plane buffer)
#pragma unroll
  for (int m = 0; m < MAXT; ++m) {
    const int it = threadIdx.x + m * 256;
    if (it < ntiles) {
      const int pw = it % P.tw, rest = it / P.tw;
      const int r0 = (rest % (TH >> 1)) << 1, d0 = (rest / (TH >> 1)) << 3;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float4* dst = reinterpret_cast<float4*>(s_out + ((size_t)(r0 + rr) * P.tw + pw) * D + d0);
        dst[0] = make_float4(acc[m][rr][0], acc[m][rr][1], acc[m][rr][2], acc[m][rr][3]);
        dst[1] = make_float4(acc[m][rr][4], acc[m][rr][5], acc[m][rr][6], acc[m][rr][7]);
      }
    }
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int ncombo = TH * P.tw * (PACK ? 4 : 1);
    for (int cb = warp; cb < ncombo; cb += nwarps) {
      const int ij = PACK ? (cb & 3) : 0, cell = PACK ? (cb >> 2) : cb;
      const int pw = cell % P.tw, r = cell / P.tw;
      const int w = w0 + pw, h = h0 + r;
      if (w >= P.W || h >= P.H) continue;
      const float* srcp = s_out + (size_t)cell * D;
      if (PACK) {
        const int i = ij >> 1, j = ij & 1;
        float* dst = P.gin + (((size_t)b * 2 * P.H + (2 * h + i)) * 2 * P.W + (2 * w + j)) * P.C;
        for (int c = lane; c < P.C; c += 32) dst[c] = srcp[4 * c + ij];
      } else {
        float* dst = P.gin + (((size_t)b * P.H + h) * P.W + w) * P.C;
        for (int d = lane; d < D; d += 32) dst[d] = srcp[d];
      }
    }
  }
}

// weight / bias gradient.  Warp (fp, half): feature pair {2fp, 2fp+1}, half of the thread tiles; 54 + 2 partial sums in
// registers over the whole persistent walk, one warp reduction + atomics at the end.
// smem: s_v[3][tw+2][PITCH] + 4, s_gc[8][tw][PITCH]
// ASYNC (default since round 2 where two buffers fit): the tiles of work item i+1 are copied with cp.async into the other half
// of a double buffer while the FMAs of work item i run -- no staging registers (the register-prefetch variant PRE took the
// kernel from 2 CTAs per SM to 1) and no exposed load phase (FLAT: nine, then two, serial DRAM round trips per work item).
template <bool PACK, bool FLAT = false, bool PRE = false, bool ASYNC = false>
__global__ void __launch_bounds__(256) stencil_wgrad8_kernel(const StencilBwdParams P) {
  PN_DYNAMIC_SHARED(float, sm);
  const int D = P.D, PITCH = D + SPAD, TWP = P.tw + 2;
  const size_t v_floats = (size_t)3 * TWP * PITCH + 4, buf_floats = v_floats + (size_t)8 * P.tw * PITCH + 4;
  float* s_v = sm;
  float* s_gc = sm + v_floats;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fp = warp & 3, half = warp >> 2;
  float wa[27], wb[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) { wa[t] = 0.0f; wb[t] = 0.0f; }
  float ba = 0.0f, bb = 0.0f;
  const int wtiles = (P.W + P.tw - 1) / P.tw;
  const int nwork = P.B * P.H * wtiles;
  const int ntiles = P.tw * (D >> 3);
  // FLAT: the staging loads of work item i+1 are issued before the FMAs of work item i (register prefetch: UA + UG float4 per
  // thread cover a whole work item at D = 32, tw = 16) -- round 1 paid nine serial DRAM round trips per work item (r02n:
  // unpack1 weight gradient 241 us for 142 MB).
  constexpr int UA = 4, UG = 8;
  FlatRegs<UA> RA;
  FlatRegs<UG> RG;
  auto prefetch = [&](int work) {
    const int wt = work % wtiles, row = work / wtiles;
    const int b = row / P.H, h = row % P.H, w0 = wt * P.tw;
    flat_prefetch<PACK, false, UA>(RA, P.in, b, P.H, P.W, P.C, 0, D, h - 1, 3, w0 - 1, TWP);
    if (PACK) flat_prefetch<false, true, UG>(RG, P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, D);
    else      flat_prefetch<true, true, UG>(RG, P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, D >> 2);
  };
  if (PRE && (int)blockIdx.x < nwork) prefetch(blockIdx.x);
  auto issue = [&](int work, int k) {      // ASYNC: all tiles of a work item -> buffer k, one commit group
    const int wt = work % wtiles, row = work / wtiles;
    const int b = row / P.H, h = row % P.H, w0 = wt * P.tw;
    float* bv = sm + (size_t)k * buf_floats;
    stage_tile_async<PACK, false>(P.in, b, P.H, P.W, P.C, 0, D, h - 1, 3, w0 - 1, TWP, bv);
    if (PACK) stage_tile_async<false, true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, bv + v_floats, D);
    else      stage_tile_async<true, true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, bv + v_floats, D >> 2);
    pn_cp_async_commit();
  };
  if (ASYNC) {
    for (int k = 0; k < 2; ++k) {
      stage_pads_zero(sm + (size_t)k * buf_floats, D, 3 * TWP);
      stage_pads_zero(sm + (size_t)k * buf_floats + v_floats, D, 8 * P.tw);
    }
    if ((int)blockIdx.x < nwork) issue(blockIdx.x, 0);
  }
  int iter = 0;
  for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++iter) {
    const int wt = work % wtiles, row = work / wtiles;
    const int b = row / P.H, h = row % P.H, w0 = wt * P.tw;
    if (ASYNC) {
      // buffer (iter + 1) & 1 was last read in iteration iter - 1, which ended with a barrier
      if (work + (int)gridDim.x < nwork) { issue(work + gridDim.x, (iter + 1) & 1); pn_cp_async_wait<1>(); }
      else pn_cp_async_wait<0>();
      s_v = sm + (size_t)(iter & 1) * buf_floats;
      s_gc = s_v + v_floats;
    } else {
      __syncthreads();
    }
    if (ASYNC) {
    } else if (FLAT && !PRE) {
      stage_tile_flat<PACK>(P.in, b, P.H, P.W, P.C, 0, D, h - 1, 3, w0 - 1, TWP, s_v);
      if (PACK) stage_tile_flat<false, true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, s_gc, D);       // the eight planes:
      else      stage_tile_flat<true, true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, s_gc, D >> 2);    // ONE load phase
    } else if (PRE) {
      flat_store<UA>(RA, s_v, D, 3 * TWP);
      flat_store<UG>(RG, s_gc, D, 8 * P.tw);
      const int step = (int)blockDim.x;
      if (3 * TWP * (D >> 2) > UA * step) stage_tile_flat<PACK>(P.in, b, P.H, P.W, P.C, 0, D, h - 1, 3, w0 - 1, TWP, s_v, 0, UA * step);
      if (8 * P.tw * (D >> 2) > UG * step) {
        if (PACK) stage_tile_flat<false, true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, s_gc, D, UG * step);
        else      stage_tile_flat<true, true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset, D, h, 8, w0, P.tw, s_gc, D >> 2, UG * step);
      }
    } else {
      stage_tile<PACK>(P.in, b, P.H, P.W, P.C, 0, D, h - 1, 3, w0 - 1, TWP, s_v);
      for (int f = 0; f < 8; ++f) {
        if (PACK) stage_tile<false>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + f * D, D, h, 1, w0, P.tw, s_gc + (size_t)f * P.tw * PITCH);
        else      stage_tile<true>(P.g, b, P.H, P.W, P.g_cstride, P.g_coffset + ((f * D) >> 2), D, h, 1, w0, P.tw, s_gc + (size_t)f * P.tw * PITCH);
      }
    }
    __syncthreads();
    if (PRE && work + (int)gridDim.x < nwork) prefetch(work + gridDim.x);     // in flight during the FMAs below
    const float* gA = s_gc + (size_t)(2 * fp) * P.tw * PITCH + SPAD;
    const float* gB = gA + (size_t)P.tw * PITCH;
    for (int it = half * 32 + lane; it < ntiles; it += 64) {
      const int pw = it % P.tw, d0 = (it / P.tw) << 3;
      const float4 a0 = *reinterpret_cast<const float4*>(gA + (size_t)pw * PITCH + d0);
      const float4 a1 = *reinterpret_cast<const float4*>(gA + (size_t)pw * PITCH + d0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(gB + (size_t)pw * PITCH + d0);
      const float4 b1 = *reinterpret_cast<const float4*>(gB + (size_t)pw * PITCH + d0 + 4);
      const float ga[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float gb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) { ba += ga[k]; bb += gb[k]; }
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          float vv[10];
          load_col10(s_v + (size_t)(dy * TWP + pw + dx) * PITCH + SPAD + d0, vv);
#pragma unroll
          for (int dz = 0; dz < 3; ++dz) {
            const int t = (dz * 3 + dy) * 3 + dx;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              wa[t] = fmaf(ga[k], vv[k + dz], wa[t]);
              wb[t] = fmaf(gb[k], vv[k + dz], wb[t]);
            }
          }
        }
    }
    if (ASYNC) __syncthreads();   // every thread is done with this buffer before the copies of iteration iter + 1 refill it
  }
#pragma unroll
  for (int t = 0; t < 28; ++t) {
    float va = (t < 27) ? wa[t < 27 ? t : 0] : ba;
    float vb = (t < 27) ? wb[t < 27 ? t : 0] : bb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      va += __shfl_xor_sync(0xffffffffu, va, o);
      vb += __shfl_xor_sync(0xffffffffu, vb, o);
    }
    if (lane == 0) {
      if (t < 27) { atomicAdd(P.gw3 + (2 * fp) * 27 + t, va); atomicAdd(P.gw3 + (2 * fp + 1) * 27 + t, vb); }
      else { atomicAdd(P.gb3 + 2 * fp, va); atomicAdd(P.gb3 + 2 * fp + 1, vb); }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm(16) + ELU   (NHWC; x may be the sum of two tensors for the residual block, layers01.py:72)
// ---------------------------------------------------------------------------------------------------
// stats[b][g] = (sum, sumsq) in double, accumulated atomically into a zeroed buffer
// TREE (staged, pn_set_tuning(PN_TUNE_GN_TREE, 1)): four independent loads per trip of the pixel loop, and the per-thread
// partial sums meet in the block through warp shuffles
// (lanes of the same float4 column, then the columns of a group) before ONE shared atomic per (warp, group), instead of
// 8 double-precision shared atomics per thread on 32 addresses -- shared fp64 atomics are compare-and-swap loops, and 16
// lanes of every warp collide on each address.
template <bool TREE>
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, const float* __restrict__ x2, int HW, int C,
                                                       int in_cstride, int pixels_per_cta, double* __restrict__ stats) {
  __shared__ double s_sum[16], s_sq[16];
  const int b = blockIdx.y;
  const int cg = C / 16;
  if (threadIdx.x < 16) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
  __syncthreads();
  const int p0 = blockIdx.x * pixels_per_cta;
  const int p1 = min(p0 + pixels_per_cta, HW);
  const int c4 = C / 4;  // float4 columns
  // thread -> (pixel lane, float4 column); a float4 never straddles a group because cg % 4 == 0 or cg in {1,2,4..}
  const int lanes = blockDim.x / c4 > 0 ? blockDim.x / c4 : 1;
  const int col = threadIdx.x % c4, pl = threadIdx.x / c4;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (pl < lanes) {
    int p = p0 + pl;
    if (TREE) {
      // four pixels per trip: the loads of a trip are independent (one 16-byte load in flight per thread and 4 CTAs per SM
      // cover less than half of the bytes the HBM latency needs in flight)
      for (; p + 3 * lanes < p1; p += 4 * lanes) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t o = ((size_t)b * HW + p + u * lanes) * in_cstride + col * 4;
          v[u] = *reinterpret_cast<const float4*>(x + o);
          if (x2) {
            const float4 w = *reinterpret_cast<const float4*>(x2 + o);
            v[u].x += w.x; v[u].y += w.y; v[u].z += w.z; v[u].w += w.w;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
          q[0] += v[u].x * v[u].x; q[1] += v[u].y * v[u].y; q[2] += v[u].z * v[u].z; q[3] += v[u].w * v[u].w;
        }
      }
    }
    for (; p < p1; p += lanes) {
      const size_t o = ((size_t)b * HW + p) * in_cstride + col * 4;
      float4 v = *reinterpret_cast<const float4*>(x + o);
      if (x2) {
        const float4 u = *reinterpret_cast<const float4*>(x2 + o);
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      }
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
    }
    const int gw_ = cg >> 2;   // float4 columns per group
    if (!(TREE && (cg & 3) == 0 && (gw_ & (gw_ - 1)) == 0 && gw_ <= 32 && (c4 & (c4 - 1)) == 0)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = (col * 4 + k) / cg;
        atomicAdd(&s_sum[g], (double)s[k]);
        atomicAdd(&s_sq[g], (double)q[k]);
      }
    }
  }
  if (TREE) {
    const int gw = cg >> 2;
    if ((cg & 3) == 0 && (gw & (gw - 1)) == 0 && gw <= 32 && (c4 & (c4 - 1)) == 0) {   // uniform over the block
      // a float4 column lies inside one group (cg % 4 == 0); idle pixel lanes contribute zeros
      double ds = (double)s[0] + (double)s[1] + (double)s[2] + (double)s[3];
      double dq = (double)q[0] + (double)q[1] + (double)q[2] + (double)q[3];
      for (int o = 16; o >= c4; o >>= 1) {          // the same column in the other pixel lanes of this warp (c4 < 32)
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
      }
      for (int o = 1; o < gw; o <<= 1) {            // the other columns of the group (consecutive lanes)
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
      }
      const int lane = threadIdx.x & 31;
      const bool leader = ((lane & (gw - 1)) == 0) && (c4 >= 32 || lane < c4) && (pl < lanes);
      if (leader) {
        const int g = (col * 4) / cg;
        atomicAdd(&s_sum[g], ds);
        atomicAdd(&s_sq[g], dq);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    atomicAdd(stats + ((size_t)b * 16 + threadIdx.x) * 2 + 0, s_sum[threadIdx.x]);
    atomicAdd(stats + ((size_t)b * 16 + threadIdx.x) * 2 + 1, s_sq[threadIdx.x]);
  }
}

// (sum, sumsq) doubles -> (mean, rstd) floats, once per (sample, group); stored behind the doubles in `stats`
__global__ void gn_finalize_stats_kernel(const double* __restrict__ stats, float* __restrict__ mr, int n, double cnt, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = stats[2 * i] / cnt;
  double var = stats[2 * i + 1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  mr[2 * i] = (float)mean;
  mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// bf16 pair of a value for the tensor-core engine: hi = rn_bf16(v), lo = rn_bf16(v - hi) (conv_engine.cu split_bf16_kernel)
__device__ __forceinline__ void split_pair(float v, unsigned short& hi, unsigned short& lo) {
#ifdef PN_EMULATE
  unsigned u;
  std::memcpy(&u, &v, 4);
  auto rn = [](float f) { unsigned w; std::memcpy(&w, &f, 4); if ((w & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((w >> 16) | 0x0040u);
                          w += 0x7fffu + ((w >> 16) & 1u); return (unsigned short)(w >> 16); };
  hi = rn(v);
  const unsigned hb = (unsigned)hi << 16;
  float hf;
  std::memcpy(&hf, &hb, 4);
  lo = rn(v - hf);
#else
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
#endif
}
__device__ __forceinline__ void store_split4(uint2* hi, uint2* lo, size_t i4, const float (&v)[4]) {
  unsigned short h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) split_pair(v[k], h[k], l[k]);
  hi[i4] = make_uint2((unsigned)h[0] | ((unsigned)h[1] << 16), (unsigned)h[2] | ((unsigned)h[3] << 16));
  lo[i4] = make_uint2((unsigned)l[0] | ((unsigned)l[1] << 16), (unsigned)l[2] | ((unsigned)l[3] << 16));
}

// y = ELU(gamma * (x - mean) * rstd + beta), written at (out_cstride, out_coffset); optional tf32 residual copy.
// Round 2: (mean, rstd) of every (sample, group) are derived from the fp64 sums in the kernel's prologue (the separate
// finalize launch is gone; CTA 0 leaves the floats in `mr` for the backward), and the kernel can also emit the bf16 hi / lo
// operand pair of y for the convolution that consumes it (y_hi / y_blo, contiguous [B,HW,C]): the split launch that re-read
// y is gone as well.
__global__ void __launch_bounds__(256) gn_elu_apply_kernel(const float* __restrict__ x, const float* __restrict__ x2, int HW, int C,
                                                           int in_cstride, float* __restrict__ mr, const double* __restrict__ sums,
                                                           double cnt, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float eps, float* __restrict__ y, float* __restrict__ y_lo, int out_cstride,
                                                           int out_coffset, int B, uint2* __restrict__ y_hi, uint2* __restrict__ y_blo) {
  PN_DYNAMIC_SHARED(float, s_mr);     // [B*16][2]
  for (int i = threadIdx.x; i < 16 * B; i += blockDim.x) {
    float mean, rstd;
    if (sums) {
      const double m = sums[2 * i] / cnt;
      double var = sums[2 * i + 1] / cnt - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) { mr[2 * i] = mean; mr[2 * i + 1] = rstd; }
    } else {
      mean = mr[2 * i]; rstd = mr[2 * i + 1];
    }
    s_mr[2 * i] = mean; s_mr[2 * i + 1] = rstd;
  }
  __syncthreads();
  const int c4 = C / 4, cg = C / 16;
  const size_t total = (size_t)B * HW * c4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(idx % c4);
    const size_t pix = idx / c4;
    const int b = (int)(pix / HW);
    const size_t o = pix * in_cstride + col * 4;
    float4 v = *reinterpret_cast<const float4*>(x + o);
    if (x2) {
      const float4 u = *reinterpret_cast<const float4*>(x2 + o);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    float in[4] = {v.x, v.y, v.z, v.w}, out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = col * 4 + k, g = c / cg;
      const float mean = s_mr[(b * 16 + g) * 2 + 0], rstd = s_mr[(b * 16 + g) * 2 + 1];
      const float z = (in[k] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      out[k] = z > 0.0f ? z : expm1f(z);   // nn.ELU(alpha=1)
    }
    const size_t oo = pix * out_cstride + out_coffset + col * 4;
    *reinterpret_cast<float4*>(y + oo) = make_float4(out[0], out[1], out[2], out[3]);
    if (y_lo)
      *reinterpret_cast<float4*>(y_lo + oo) = make_float4(out[0] - tf32_trunc(out[0]), out[1] - tf32_trunc(out[1]),
                                                          out[2] - tf32_trunc(out[2]), out[3] - tf32_trunc(out[3]));
    if (y_hi) store_split4(y_hi, y_blo, idx, out);
  }
}

// The same pass with the thread <-> (pixel lane, float4 column) mapping of the backward kernels: one CTA per (pixel range,
// sample), the four channels of a thread fixed -- gamma / beta / mean / rstd live in registers and the loop has no 64-bit
// division (r02n ncu: the flat-index kernel above moved 3.0 TB/s on the 192x640 maps where the backward apply pass moves 5.1).
__global__ void __launch_bounds__(256) gn_elu_apply4_kernel(const float* __restrict__ x, const float* __restrict__ x2, int HW, int C,
                                                            float* __restrict__ mr, const double* __restrict__ sums, double cnt,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                            int pixels_per_cta, float* __restrict__ y, int out_cstride, int out_coffset,
                                                            uint2* __restrict__ y_hi, uint2* __restrict__ y_blo) {
  __shared__ float s_mr[32];
  const int b = blockIdx.y, cg = C / 16, c4 = C / 4;
  if (threadIdx.x < 16) {
    const int i = b * 16 + threadIdx.x;
    const double m = sums[2 * i] / cnt;
    double var = sums[2 * i + 1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
    s_mr[2 * threadIdx.x] = mean; s_mr[2 * threadIdx.x + 1] = rstd;
    if (blockIdx.x == 0) { mr[2 * i] = mean; mr[2 * i + 1] = rstd; }
  }
  __syncthreads();
  const int p0 = blockIdx.x * pixels_per_cta, p1 = min(p0 + pixels_per_cta, HW);
  const int lanes = blockDim.x / c4 > 0 ? blockDim.x / c4 : 1;
  const int col = threadIdx.x % c4, pl = threadIdx.x / c4;
  if (pl >= lanes) return;
  float mean[4], rstd[4], gm[4], bt[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = col * 4 + k, g = c / cg;
    mean[k] = s_mr[2 * g]; rstd[k] = s_mr[2 * g + 1];
    gm[k] = __ldg(gamma + c); bt[k] = __ldg(beta + c);
  }
  auto load = [&](int p) {
    const size_t o = ((size_t)b * HW + p) * C + col * 4;
    float4 v = *reinterpret_cast<const float4*>(x + o);
    if (x2) {
      const float4 u = *reinterpret_cast<const float4*>(x2 + o);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    return v;
  };
  auto apply = [&](int p, const float4& v) {
    const float in[4] = {v.x, v.y, v.z, v.w};
    float out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float z = (in[k] - mean[k]) * rstd[k] * gm[k] + bt[k];
      out[k] = z > 0.0f ? z : expm1f(z);   // nn.ELU(alpha=1)
    }
    const size_t pix = (size_t)b * HW + p;
    *reinterpret_cast<float4*>(y + pix * out_cstride + out_coffset + col * 4) = make_float4(out[0], out[1], out[2], out[3]);
    if (y_hi) store_split4(y_hi, y_blo, (pix * C + col * 4) >> 2, out);
  };
  int p = p0 + pl;
  for (; p + 3 * lanes < p1; p += 4 * lanes) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = load(p + u * lanes);
#pragma unroll
    for (int u = 0; u < 4; ++u) apply(p + u * lanes, v[u]);
  }
  for (; p < p1; p += lanes) apply(p, load(p));
}

// backward, pass 1: per (b, c): sum dz, sum dz*xhat with dz = dy * ELU'(z) (ELU' = 1 for y>0 else y+1)
//   -> bc[b][c][2] (double, zeroed)
__global__ void __launch_bounds__(256) gn_elu_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                                const float* __restrict__ y, const float* __restrict__ dy, int HW, int C,
                                                                int in_cstride, int y_cstride, int y_coffset, int dy_cstride,
                                                                int dy_coffset, const float* __restrict__ mr, float eps,
                                                                int pixels_per_cta, double* __restrict__ bc) {
  const int b = blockIdx.y;
  const int cg = C / 16;
  const int p0 = blockIdx.x * pixels_per_cta, p1 = min(p0 + pixels_per_cta, HW);
  // thread -> channel (c = threadIdx.x % C), pixel lane
  const int lanes = blockDim.x / C > 0 ? blockDim.x / C : 1;
  for (int c = threadIdx.x % C + (threadIdx.x / C >= lanes ? C : 0); c < C; c += blockDim.x) {
    const int pl = (blockDim.x >= C) ? threadIdx.x / C : 0;
    const int g = c / cg;
    const float fmean = __ldg(mr + ((size_t)b * 16 + g) * 2 + 0), rstd = __ldg(mr + ((size_t)b * 16 + g) * 2 + 1);
    float s1 = 0.0f, s2 = 0.0f;
    for (int p = p0 + pl; p < p1; p += lanes) {
      const size_t pix = (size_t)b * HW + p;
      float xv = x[pix * in_cstride + c];
      if (x2) xv += x2[pix * in_cstride + c];
      const float yv = y[pix * y_cstride + y_coffset + c];
      const float dz = dy[pix * dy_cstride + dy_coffset + c] * (yv > 0.0f ? 1.0f : yv + 1.0f);
      s1 += dz;
      s2 += dz * (xv - fmean) * rstd;
    }
    atomicAdd(bc + ((size_t)b * C + c) * 2 + 0, (double)s1);
    atomicAdd(bc + ((size_t)b * C + c) * 2 + 1, (double)s2);
  }
}

// backward, pass 2: dx = rstd * (gamma*dz - mean_g(gamma*dz) - xhat * mean_g(gamma*dz*xhat)); also dgamma/dbeta
// (first CTA column only) from bc
__global__ void __launch_bounds__(256) gn_elu_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                               const float* __restrict__ y, const float* __restrict__ dy, int HW, int C,
                                                               int in_cstride, int y_cstride, int y_coffset, int dy_cstride,
                                                               int dy_coffset, const float* __restrict__ mr,
                                                               const float* __restrict__ gmeans, const float* __restrict__ gamma, float eps,
                                                               float* __restrict__ dx, float* __restrict__ dx_lo, int B) {
  const int cg = C / 16;
  const size_t total = (size_t)B * HW * C;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const size_t pix = idx / C;
    const int b = (int)(pix / HW), g = c / cg;
    const float fmean = __ldg(mr + ((size_t)b * 16 + g) * 2 + 0), rstd = __ldg(mr + ((size_t)b * 16 + g) * 2 + 1);
    const float m1 = __ldg(gmeans + ((size_t)b * 16 + g) * 2 + 0), m2 = __ldg(gmeans + ((size_t)b * 16 + g) * 2 + 1);
    float xv = x[pix * in_cstride + c];
    if (x2) xv += x2[pix * in_cstride + c];
    const float xhat = (xv - fmean) * rstd;
    const float yv = y[pix * y_cstride + y_coffset + c];
    const float dz = dy[pix * dy_cstride + dy_coffset + c] * (yv > 0.0f ? 1.0f : yv + 1.0f);
    const float r = rstd * (__ldg(gamma + c) * dz - m1 - xhat * m2);
    dx[idx] = r;
    if (dx_lo) dx_lo[idx] = r - tf32_trunc(r);
  }
}

// float4 versions of the two backward passes (all channel strides / offsets multiples of 4): thread <-> (pixel lane,
// float4 column), one CTA per (pixel range, sample) like gn_stats_kernel.  Pass 2 also accumulates sum_pixels dx per
// channel -- the bias gradient of the convolution that produced x (layers01.py:28) -- so that no separate reduction
// pass has to re-read dx.
__global__ void __launch_bounds__(256) gn_elu_bwd_reduce4_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                                 const float* __restrict__ y, const float* __restrict__ dy, int HW, int C,
                                                                 int in_cstride, int y_cstride, int y_coffset, int dy_cstride,
                                                                 int dy_coffset, const float* __restrict__ mr, int pixels_per_cta,
                                                                 double* __restrict__ bc) {
  __shared__ float s_red[2 * 1024];
  const int b = blockIdx.y, cg = C / 16, c4 = C / 4;
  const int p0 = blockIdx.x * pixels_per_cta, p1 = min(p0 + pixels_per_cta, HW);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_red[i] = 0.0f;
  __syncthreads();
  const int lanes = blockDim.x / c4 > 0 ? blockDim.x / c4 : 1;
  const int col = threadIdx.x % c4, pl = threadIdx.x / c4;
  if (pl < lanes) {
    float mean[4], rstd[4], s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = (col * 4 + k) / cg;
      mean[k] = __ldg(mr + ((size_t)b * 16 + g) * 2 + 0);
      rstd[k] = __ldg(mr + ((size_t)b * 16 + g) * 2 + 1);
    }
    for (int p = p0 + pl; p < p1; p += lanes) {
      const size_t pix = (size_t)b * HW + p;
      float4 xv = *reinterpret_cast<const float4*>(x + pix * in_cstride + col * 4);
      if (x2) {
        const float4 u = *reinterpret_cast<const float4*>(x2 + pix * in_cstride + col * 4);
        xv.x += u.x; xv.y += u.y; xv.z += u.z; xv.w += u.w;
      }
      const float4 yv = *reinterpret_cast<const float4*>(y + pix * y_cstride + y_coffset + col * 4);
      const float4 gv = *reinterpret_cast<const float4*>(dy + pix * dy_cstride + dy_coffset + col * 4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dz = gs[k] * (ys[k] > 0.0f ? 1.0f : ys[k] + 1.0f);
        s1[k] += dz;
        s2[k] += dz * (xs[k] - mean[k]) * rstd[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&s_red[(col * 4 + k) * 2 + 0], s1[k]);
      atomicAdd(&s_red[(col * 4 + k) * 2 + 1], s2[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(bc + (size_t)b * C * 2 + i, (double)s_red[i]);
}

// Round 2: the finalize launch between the two passes is folded in -- every CTA derives the two group means of ITS sample from
// the channel sums `bc` (16 threads x C/16 channels), the first CTA column also writes dgamma / dbeta -- and the kernel can
// emit the bf16 hi / lo pair of dx, the output-gradient operand of the convolution that produced x.
__global__ void __launch_bounds__(256) gn_elu_bwd_apply4_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                                const float* __restrict__ y, const float* __restrict__ dy, int HW, int C,
                                                                int in_cstride, int y_cstride, int y_coffset, int dy_cstride,
                                                                int dy_coffset, const float* __restrict__ mr,
                                                                const double* __restrict__ bc, double cnt, const float* __restrict__ gamma,
                                                                int pixels_per_cta, float* __restrict__ dx, float* __restrict__ dx_lo,
                                                                float* __restrict__ dsum, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, int B, uint2* __restrict__ dx_hi,
                                                                uint2* __restrict__ dx_blo) {
  __shared__ float s_red[1024];
  __shared__ float s_gm[32];
  const int b = blockIdx.y, cg = C / 16, c4 = C / 4;
  const int p0 = blockIdx.x * pixels_per_cta, p1 = min(p0 + pixels_per_cta, HW);
  if (threadIdx.x < 16) {
    // (mean_g(gamma*dz), mean_g(gamma*dz*xhat)) over the HW * C/16 elements of group g of sample b
    const int g = threadIdx.x;
    double m1 = 0.0, m2 = 0.0;
    for (int k = 0; k < cg; ++k) {
      const int cc = g * cg + k;
      const double gmv = (double)gamma[cc];
      m1 += gmv * bc[((size_t)b * C + cc) * 2 + 0];
      m2 += gmv * bc[((size_t)b * C + cc) * 2 + 1];
    }
    s_gm[2 * g + 0] = (float)(m1 / cnt);
    s_gm[2 * g + 1] = (float)(m2 / cnt);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    // dgamma[c] = sum_b bc[b][c][1], dbeta[c] = sum_b bc[b][c][0]
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double a = 0.0, d = 0.0;
      for (int bb = 0; bb < B; ++bb) { d += bc[((size_t)bb * C + c) * 2 + 0]; a += bc[((size_t)bb * C + c) * 2 + 1]; }
      dgamma[c] = (float)a;
      dbeta[c] = (float)d;
    }
  }
  if (dsum) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_red[i] = 0.0f;
  }
  __syncthreads();
  const int lanes = blockDim.x / c4 > 0 ? blockDim.x / c4 : 1;
  const int col = threadIdx.x % c4, pl = threadIdx.x / c4;
  if (pl < lanes) {
    float mean[4], rstd[4], m1[4], m2[4], gm[4], acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = col * 4 + k, g = c / cg;
      mean[k] = __ldg(mr + ((size_t)b * 16 + g) * 2 + 0);
      rstd[k] = __ldg(mr + ((size_t)b * 16 + g) * 2 + 1);
      m1[k] = s_gm[2 * g + 0];
      m2[k] = s_gm[2 * g + 1];
      gm[k] = __ldg(gamma + c);
    }
    for (int p = p0 + pl; p < p1; p += lanes) {
      const size_t pix = (size_t)b * HW + p;
      float4 xv = *reinterpret_cast<const float4*>(x + pix * in_cstride + col * 4);
      if (x2) {
        const float4 u = *reinterpret_cast<const float4*>(x2 + pix * in_cstride + col * 4);
        xv.x += u.x; xv.y += u.y; xv.z += u.z; xv.w += u.w;
      }
      const float4 yv = *reinterpret_cast<const float4*>(y + pix * y_cstride + y_coffset + col * 4);
      const float4 gv = *reinterpret_cast<const float4*>(dy + pix * dy_cstride + dy_coffset + col * 4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
      float r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xhat = (xs[k] - mean[k]) * rstd[k];
        const float dz = gs[k] * (ys[k] > 0.0f ? 1.0f : ys[k] + 1.0f);
        r[k] = rstd[k] * (gm[k] * dz - m1[k] - xhat * m2[k]);
        acc[k] += r[k];
      }
      const size_t o = pix * C + col * 4;
      *reinterpret_cast<float4*>(dx + o) = make_float4(r[0], r[1], r[2], r[3]);
      if (dx_lo)
        *reinterpret_cast<float4*>(dx_lo + o) = make_float4(r[0] - tf32_trunc(r[0]), r[1] - tf32_trunc(r[1]),
                                                            r[2] - tf32_trunc(r[2]), r[3] - tf32_trunc(r[3]));
      if (dx_hi) store_split4(dx_hi, dx_blo, o >> 2, r);
    }
    if (dsum) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&s_red[col * 4 + k], acc[k]);
    }
  }
  if (dsum) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dsum + i, s_red[i]);
  }
}

// dgamma[c] = sum_b bc[b][c][1], dbeta[c] = sum_b bc[b][c][0];
// gmeans[b][g] = (mean_g(gamma*dz), mean_g(gamma*dz*xhat)) over the (HW * C/16) elements of the group
__global__ void gn_bwd_finalize_kernel(const double* __restrict__ bc, const float* __restrict__ gamma, int B, int C, double cnt,
                                       float* __restrict__ gmeans, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int cg = C / 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * 16 + C; i += gridDim.x * blockDim.x) {
    if (i < B * 16) {
      const int b = i / 16, g = i % 16;
      double m1 = 0.0, m2 = 0.0;
      for (int k = 0; k < cg; ++k) {
        const int cc = g * cg + k;
        const double gm = (double)gamma[cc];
        m1 += gm * bc[((size_t)b * C + cc) * 2 + 0];
        m2 += gm * bc[((size_t)b * C + cc) * 2 + 1];
      }
      gmeans[i * 2 + 0] = (float)(m1 / cnt);
      gmeans[i * 2 + 1] = (float)(m2 / cnt);
    } else {
      const int c = i - B * 16;
      double a = 0.0, d = 0.0;
      for (int b = 0; b < B; ++b) { d += bc[((size_t)b * C + c) * 2 + 0]; a += bc[((size_t)b * C + c) * 2 + 1]; }
      dgamma[c] = (float)a;
      dbeta[c] = (float)d;
    }
  }
}

// bias gradient of an NHWC tensor: db[c] = sum over pixels (conv bias, layers01.py:28)
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ g, size_t pixels, int C, int pixels_per_cta,
                                                          float* __restrict__ out) {
  const size_t p0 = (size_t)blockIdx.x * pixels_per_cta;
  const size_t p1 = p0 + pixels_per_cta < pixels ? p0 + pixels_per_cta : pixels;
  __shared__ float s_acc[1024];
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_acc[c] = 0.0f;
  __syncthreads();
  const int lanes = blockDim.x / C > 0 ? blockDim.x / C : 1;
  for (int c = threadIdx.x % C + (threadIdx.x / C >= lanes ? C : 0); c < C; c += blockDim.x) {
    const int pl = (blockDim.x >= C) ? threadIdx.x / C : 0;
    float s = 0.0f;
    for (size_t p = p0 + pl; p < p1; p += lanes) s += g[p * C + c];
    atomicAdd(&s_acc[c], s);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(out + c, s_acc[c]);
}

// ---------------------------------------------------------------------------------------------------
// single-channel head convolution: Conv2d(C -> 1, 3x3, zero pad 1) of InvDepth (layers01.py:98-122)
//   One output channel is HBM-bound fp32 work (read x once), not a GEMM: x [B,H,W,C] NHWC, w [9][C], y [B,H,W].
// ---------------------------------------------------------------------------------------------------
struct HeadParams {
  int B, H, W, C, tw;      // tw x 8 output pixels per tile
  const float* x;
  const float* w;          // [9][C]  (tap-major)
  const float* bias;       // [1]
  const float* dy;         // [B,H,W]
  float* y;                // [B,H,W]
  float* dx;               // [B,H,W,C]
  float* dw;               // [9][C] accumulated
  float* db;               // [1] accumulated
};

// smem: s_x[(8+2)][(tw+2)][C+4], s_w[9][C]
template <bool FLAT>
__global__ void __launch_bounds__(256) head_fwd_kernel(const HeadParams P) {
  PN_DYNAMIC_SHARED(float, sm);
  const int C = P.C, PITCH = C + SPAD, TWP = P.tw + 2;
  float* s_x = sm;
  float* s_w = sm + 10 * TWP * PITCH + 4;
  const int w0 = blockIdx.x * P.tw, h0 = blockIdx.y * 8, b = blockIdx.z;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = __ldg(P.w + i);
  if (FLAT) stage_tile_flat<false>(P.x, b, P.H, P.W, C, 0, C, h0 - 1, 10, w0 - 1, TWP, s_x);
  else stage_tile<false>(P.x, b, P.H, P.W, C, 0, C, h0 - 1, 10, w0 - 1, TWP, s_x);
  __syncthreads();
  for (int it = threadIdx.x; it < 8 * P.tw; it += blockDim.x) {
    const int pw = it % P.tw, r = it / P.tw;
    const int w = w0 + pw, h = h0 + r;
    if (w >= P.W || h >= P.H) continue;
    float acc = __ldg(P.bias);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float4* xp = reinterpret_cast<const float4*>(s_x + (size_t)((r + dy) * TWP + pw + dx) * PITCH + SPAD);
        const float4* wp = reinterpret_cast<const float4*>(s_w + (dy * 3 + dx) * C);
        for (int q = 0; q < (C >> 2); ++q) {
          const float4 a = xp[q], ww = wp[q];
          acc = fmaf(a.x, ww.x, acc); acc = fmaf(a.y, ww.y, acc); acc = fmaf(a.z, ww.z, acc); acc = fmaf(a.w, ww.w, acc);
        }
      }
    P.y[((size_t)b * P.H + h) * P.W + w] = acc;
  }
}

// data gradient: dx[p][c] = sum_tap w[tap][c] * dy[p - tap + 1].  One thread per (pixel, channel quad).
__global__ void __launch_bounds__(256) head_dgrad_kernel(const HeadParams P) {
  PN_DYNAMIC_SHARED(float, sm);
  const int C = P.C, TWP = P.tw + 2;
  float* s_dy = sm;                 // [10][tw+2]
  float* s_w = sm + ((10 * TWP + 3) & ~3);
  const int w0 = blockIdx.x * P.tw, h0 = blockIdx.y * 8, b = blockIdx.z;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = __ldg(P.w + i);
  for (int i = threadIdx.x; i < 10 * TWP; i += blockDim.x) {
    const int c = i % TWP, r = i / TWP;
    const int hh = h0 - 1 + r, ww = w0 - 1 + c;
    s_dy[i] = (hh >= 0 && hh < P.H && ww >= 0 && ww < P.W) ? __ldg(P.dy + ((size_t)b * P.H + hh) * P.W + ww) : 0.0f;
  }
  __syncthreads();
  const int c4n = C >> 2;
  for (int it = threadIdx.x; it < 8 * P.tw * c4n; it += blockDim.x) {
    const int q = it % c4n, pix = it / c4n;
    const int pw = pix % P.tw, r = pix / P.tw;
    const int w = w0 + pw, h = h0 + r;
    if (w >= P.W || h >= P.H) continue;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float g = s_dy[(r + 2 - dy) * TWP + pw + 2 - dx];
        const float4 ww = *reinterpret_cast<const float4*>(s_w + (dy * 3 + dx) * C + 4 * q);
        a.x = fmaf(g, ww.x, a.x); a.y = fmaf(g, ww.y, a.y); a.z = fmaf(g, ww.z, a.z); a.w = fmaf(g, ww.w, a.w);
      }
    *reinterpret_cast<float4*>(P.dx + (((size_t)b * P.H + h) * P.W + w) * C + 4 * q) = a;
  }
}

// weight / bias gradient: dw[tap][c] = sum_p x[p + tap - 1][c] * dy[p].  Persistent CTAs; thread <-> (tap, channel quad)
// with its partial sums in registers over the whole walk (MAXQ combos per thread), atomics once at the end.
template <int MAXQ, bool FLAT>
__global__ void __launch_bounds__(256) head_wgrad_kernel(const HeadParams P) {
  PN_DYNAMIC_SHARED(float, sm);
  const int C = P.C, PITCH = C + SPAD, TWP = P.tw + 2;
  float* s_x = sm;
  float* s_dy = sm + 10 * TWP * PITCH + 4;   // [8][tw]
  const int c4n = C >> 2, ncombo = 9 * c4n;
  float4 acc[MAXQ];
#pragma unroll
  for (int m = 0; m < MAXQ; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  float bsum = 0.0f;
  const int tiles_x = (P.W + P.tw - 1) / P.tw, tiles_y = (P.H + 7) / 8;
  const int nwork = P.B * tiles_y * tiles_x;
  for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
    const int tx = work % tiles_x, ty = (work / tiles_x) % tiles_y, b = work / (tiles_x * tiles_y);
    const int w0 = tx * P.tw, h0 = ty * 8;
    __syncthreads();
    if (FLAT) stage_tile_flat<false>(P.x, b, P.H, P.W, C, 0, C, h0 - 1, 10, w0 - 1, TWP, s_x);
  else stage_tile<false>(P.x, b, P.H, P.W, C, 0, C, h0 - 1, 10, w0 - 1, TWP, s_x);
    for (int i = threadIdx.x; i < 8 * P.tw; i += blockDim.x) {
      const int pw = i % P.tw, r = i / P.tw;
      const int hh = h0 + r, ww = w0 + pw;
      const float g = (hh < P.H && ww < P.W) ? __ldg(P.dy + ((size_t)b * P.H + hh) * P.W + ww) : 0.0f;
      s_dy[i] = g;
      bsum += g;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MAXQ; ++m) {
      const int cb = threadIdx.x + m * 256;
      if (cb < ncombo) {
        const int q = cb % c4n, tap = cb / c4n;
        const int dy = tap / 3, dx = tap - 3 * dy;
        float4 a = acc[m];
        for (int r = 0; r < 8; ++r)
          for (int pw = 0; pw < P.tw; ++pw) {
            const float g = s_dy[r * P.tw + pw];
            const float4 xv = *reinterpret_cast<const float4*>(s_x + (size_t)((r + dy) * TWP + pw + dx) * PITCH + SPAD + 4 * q);
            a.x = fmaf(g, xv.x, a.x); a.y = fmaf(g, xv.y, a.y); a.z = fmaf(g, xv.z, a.z); a.w = fmaf(g, xv.w, a.w);
          }
        acc[m] = a;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MAXQ; ++m) {
    const int cb = threadIdx.x + m * 256;
    if (cb < ncombo) {
      const int q = cb % c4n, tap = cb / c4n;
      float* d = P.dw + tap * C + 4 * q;
      atomicAdd(d + 0, acc[m].x); atomicAdd(d + 1, acc[m].y); atomicAdd(d + 2, acc[m].z); atomicAdd(d + 3, acc[m].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(P.db, bsum);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// process-wide tuning switch behind pn_set_tuning(PN_TUNE_STAGE_FLAT, v): -1 = not set yet -> environment PN_STAGE_FLAT
static std::atomic<int> g_stage_flat{-1};
static int stage_flat() {
  int v = g_stage_flat.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("PN_STAGE_FLAT");
    v = (e && e[0] == '0') ? 0 : 1;   // ON by default since round 2 (B200: head fwd 0.239 -> 0.093 ms, step -1.5 ms); PN_STAGE_FLAT=0 = round-1 staging
    g_stage_flat.store(v, std::memory_order_relaxed);
  }
  return v;
}


// ---------------------------------------------------------------------------------------------------
// STAGED (functional.set_unpack_tiled / PN_UNPACK_TILED): weight-gradient re-layout through shared memory.
//   dw[co][ci][tap] = dwp[co][tap][ci]   (dwp rows are kpad floats apart: the weight-gradient kernel's native layout)
// The element-per-thread gather in conv_engine.cu (unpack_weight_grad_kernel) reads with a stride of kpad floats between
// neighbouring threads: 0.62 ms per step for 0.5 GB in and 0.5 GB out (0.14 ms at the HBM rate).  Here one CTA moves a
// (co, 128 input channels) block: coalesced rows in, transposed in shared memory (pitch = taps, odd or 1: conflict-free),
// one contiguous run of 128*taps floats out.
// ---------------------------------------------------------------------------------------------------
constexpr int UNPACK_CH = 128;
__global__ void __launch_bounds__(256) unpack_weight_grad_tiled_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cin,
                                                                       int taps, int kpad) {
  PN_DYNAMIC_SHARED(float, sm);   // [UNPACK_CH][taps]
  const int co = blockIdx.y, ci0 = blockIdx.x * UNPACK_CH;
  const int n = min(UNPACK_CH, Cin - ci0);
  const float* src = dwp + (size_t)co * taps * kpad + ci0;
  for (int i = threadIdx.x; i < taps * UNPACK_CH; i += blockDim.x) {
    const int tap = i / UNPACK_CH, c = i % UNPACK_CH;
    if (c < n) sm[c * taps + tap] = __ldg(src + (size_t)tap * kpad + c);
  }
  __syncthreads();
  float* out = dw + ((size_t)co * Cin + ci0) * taps;
  for (int i = threadIdx.x; i < n * taps; i += blockDim.x) out[i] = sm[i];
}

// register prefetch of the next work item / feature plane in the stencil backward kernels: OFF by default -- measured on the
// B200 (gpurun r02v, whole step 19.92 ms without / 20.15 ms with): the 48 prefetch registers take the kernels from 2 CTAs per
// SM to 1 (118 -> 226 and 110 -> 189 registers) and the data-gradient launches get slower (0.217 -> 0.259 ms); A/B: PN_STENCIL_PREFETCH=1
static int stencil_prefetch() {
  static std::atomic<int> v{-1};
  int x = v.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = std::getenv("PN_STENCIL_PREFETCH");
    x = (e && e[0] == '1') ? 1 : 0;
    v.store(x, std::memory_order_relaxed);
  }
  return x;
}

// cp.async double buffering of the stencil weight-gradient kernel (A/B: PN_STENCIL_ASYNC=0)
static int stencil_async() {
  static std::atomic<int> v{-1};
  int x = v.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = std::getenv("PN_STENCIL_ASYNC");
    x = (e && e[0] == '0') ? 0 : 1;
    v.store(x, std::memory_order_relaxed);
  }
  return x;
}

static std::atomic<int> g_gn_tree{-1};
static int gn_tree() {
  int v = g_gn_tree.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("PN_GN_TREE");
    v = (e && e[0] == '0') ? 0 : 1;   // ON by default since round 2 (B200: gn_stats 0.595 -> 0.363 ms per step); PN_GN_TREE=0 = fp64 shared atomics
    g_gn_tree.store(v, std::memory_order_relaxed);
  }
  return v;
}

static size_t stencil_smem_bytes(int D, int tw, bool bwd, bool unpack_bwd) {
  size_t f = (size_t)(bwd ? 2 : 1) * 3 * (tw + 2) * (D + 2) + 224;
  if (bwd) f += 216 + 8 * 28 + 32;
  if (unpack_bwd) f += (size_t)6 * 2 * (tw + 2) * 2 * D;
  return f * sizeof(float);
}

static int pick_tw(int D, int W, bool bwd, bool unpack_bwd = false) {
  // keep the working set under ~96 KB (200 KB for the unpack backward, which also stages the high-res gradient)
  // and tw*D <= 2048 items for the backward's per-thread accumulators
  const size_t budget = unpack_bwd ? 200 * 1024 : 96 * 1024;
  int tw = 8;
  while (tw > 1 && (stencil_smem_bytes(D, tw, bwd, unpack_bwd) > budget || (bwd && tw * D > 2048))) tw >>= 1;
  if (tw > W) tw = W;
  return tw;
}

}  // namespace layers
}  // namespace pn

using namespace pn;
using namespace pn::layers;

extern "C" int pn_feature_stencil_forward(int pack, const float* in, const float* w3, const float* b3, float* out, float* out_lo,
                                          int batch, int h_low, int w_low, int channels, int out_cstride, int out_coffset,
                                          pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(in && w3 && b3 && out && batch > 0 && h_low > 0 && w_low > 0 && channels > 0, PN_ERR_BAD_ARGUMENT,
             "pn_feature_stencil_forward: bad argument");
  TraceScope ts(stream, "stencil_fwd pack%d B%d H%d W%d C%d", pack, batch, h_low, w_low, channels);
  StencilParams P{};
  P.B = batch; P.H = h_low; P.W = w_low; P.C = channels; P.D = pack ? 4 * channels : channels;
  P.in = in; P.w3 = w3; P.b3 = b3; P.out = out; P.out_lo = out_lo;
  P.out_cstride = out_cstride; P.out_coffset = out_coffset;
  const bool vec_ok = aligned16(in) && aligned16(out) && (!out_lo || aligned16(out_lo)) && out_cstride % 4 == 0 && out_coffset % 4 == 0;
  if (P.D % 8 == 0 && vec_ok) {   // register-tiled production path
    int tw = 32;
    auto smem_of = [&](int t) { return ((size_t)3 * (t + 2) * (P.D + SPAD) + 4 + 224) * sizeof(float); };
    while (tw > 1 && (tw * (P.D >> 3) > 512 || smem_of(tw) > 200 * 1024)) tw >>= 1;
    while (tw > 1 && (tw >> 1) >= P.W) tw >>= 1;
    PN_REQUIRE(smem_of(tw) <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_feature_stencil_forward: depth %d too large", P.D);
    P.tw = tw;
    const size_t smem8 = smem_of(tw);
    dim3 grid8((P.W + tw - 1) / tw, P.H, P.B);
    auto launch8 = [&](auto kern) -> int {
      PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
      PN_LAUNCH(kern, grid8, 256, smem8, stream, P);
      return 0;
    };
    const bool flat = stage_flat() != 0;   // staged all-threads tile staging (pn_set_tuning)
    const int lrc8 = pack ? (flat ? launch8(stencil_fwd8_kernel<true, true>) : launch8(stencil_fwd8_kernel<true>))
                          : (flat ? launch8(stencil_fwd8_kernel<false, true>) : launch8(stencil_fwd8_kernel<false>));
    if (lrc8) return lrc8;
    count_launch();
    return check_launch("stencil_fwd8_kernel");
  }
  P.tw = pick_tw(P.D, P.W, false);
  PN_REQUIRE((size_t)3 * 3 * (P.D + 2) * 4 <= 200 * 1024, PN_ERR_UNSUPPORTED, "pn_feature_stencil_forward: depth %d too large", P.D);
  const size_t smem = ((size_t)3 * (P.tw + 2) * (P.D + 2) + 224) * sizeof(float);
  dim3 grid((P.W + P.tw - 1) / P.tw, P.H, P.B);
  if (pack) {
    PN_CUDA(cudaFuncSetAttribute(stencil_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PN_LAUNCH((stencil_fwd_kernel<true>), grid, 256, smem, stream, P);
  } else {
    PN_CUDA(cudaFuncSetAttribute(stencil_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PN_LAUNCH((stencil_fwd_kernel<false>), grid, 256, smem, stream, P);
  }
  count_launch();
  return check_launch("stencil_fwd_kernel");
}

// parts: 1 = data gradient (gin), 2 = weight / bias gradient (gw3, gb3), 3 = both.  The two halves are independent launches:
// the caller may put the weight half on another stream (the data half is the critical path of the backward).
static int stencil_backward_impl(int pack, const float* in, const float* g, const float* w3, float* gin, float* gw3,
                                 float* gb3, int batch, int h_low, int w_low, int channels, int g_cstride, int g_coffset,
                                 int parts, cudaStream_t stream) {
  PN_REQUIRE(in && g && w3 && batch > 0 && h_low > 0 && w_low > 0 && channels > 0 && (parts & 3) && (!(parts & 1) || gin) &&
                 (!(parts & 2) || (gw3 && gb3)), PN_ERR_BAD_ARGUMENT, "pn_feature_stencil_backward: bad argument");
  TraceScope ts(stream, "stencil_bwd pack%d B%d H%d W%d C%d", pack, batch, h_low, w_low, channels);
  StencilBwdParams P{};
  P.B = batch; P.H = h_low; P.W = w_low; P.C = channels; P.D = pack ? 4 * channels : channels;
  P.in = in; P.g = g; P.w3 = w3; P.gin = gin; P.gw3 = gw3; P.gb3 = gb3;
  P.g_cstride = g_cstride; P.g_coffset = g_coffset;
  if (parts & 2) {
    PN_CUDA(cudaMemsetAsync(gw3, 0, sizeof(float) * 216, stream));
    PN_CUDA(cudaMemsetAsync(gb3, 0, sizeof(float) * 8, stream));
  }
  const bool vec_ok = aligned16(in) && aligned16(g) && (!gin || aligned16(gin)) && g_cstride % 4 == 0 && g_coffset % 4 == 0;
  if (P.D % 8 == 0 && vec_ok) {   // register-tiled production path
    const int PITCH = P.D + SPAD;
    // data gradient
    if (parts & 1) {
      StencilBwd8Params Q8{};
      Q8.p = P;
      int th = (P.H >= 3) ? 4 : 2, tw = 32;
      auto tiles_of = [&](int t) { return (th >> 1) * t * (P.D >> 3); };
      auto smem_of = [&](int t) { return ((size_t)(th + 2) * (t + 2) * PITCH + 4 + 216) * sizeof(float); };
      while (tw > 1 && (tiles_of(tw) > 1024 || smem_of(tw) > 200 * 1024)) tw >>= 1;
      if (tiles_of(tw) > 1024 || smem_of(tw) > 200 * 1024) { th = 2; tw = 32; while (tw > 1 && (tiles_of(tw) > 1024 || smem_of(tw) > 200 * 1024)) tw >>= 1; }
      while (tw > 1 && (tw >> 1) >= P.W) tw >>= 1;
      PN_REQUIRE(tiles_of(tw) <= 1024 && smem_of(tw) <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_feature_stencil_backward: depth %d too large", P.D);
      Q8.p.tw = tw; Q8.th = th;
      const size_t plane_bytes = ((size_t)(th + 2) * (tw + 2) * PITCH + 4) * sizeof(float);
      const bool async8 = stage_flat() && !stencil_prefetch() && stencil_async() && smem_of(tw) + plane_bytes <= 200 * 1024;
      const size_t smem8 = smem_of(tw) + (async8 ? plane_bytes : 0);
      const int maxt = (tiles_of(tw) + 255) / 256;
      dim3 grid8((P.W + tw - 1) / tw, (P.H + th - 1) / th, P.B);
      auto launch = [&](auto kern) -> int {
        PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
        PN_LAUNCH(kern, grid8, 256, smem8, stream, Q8);
        return 0;
      };
      int lrc;
      if (async8) {
        if (pack) lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<true, 1, true, false, true>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<true, 2, true, false, true>) : launch(stencil_bwd8_kernel<true, 4, true, false, true>);
        else      lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<false, 1, true, false, true>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<false, 2, true, false, true>) : launch(stencil_bwd8_kernel<false, 4, true, false, true>);
      } else if (stage_flat() && stencil_prefetch()) {
        if (pack) lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<true, 1, true, true>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<true, 2, true, true>) : launch(stencil_bwd8_kernel<true, 4, true, true>);
        else      lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<false, 1, true, true>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<false, 2, true, true>) : launch(stencil_bwd8_kernel<false, 4, true, true>);
      } else if (stage_flat()) {
        if (pack) lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<true, 1, true>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<true, 2, true>) : launch(stencil_bwd8_kernel<true, 4, true>);
        else      lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<false, 1, true>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<false, 2, true>) : launch(stencil_bwd8_kernel<false, 4, true>);
      } else {
        if (pack) lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<true, 1>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<true, 2>) : launch(stencil_bwd8_kernel<true, 4>);
        else      lrc = (maxt <= 1) ? launch(stencil_bwd8_kernel<false, 1>) : (maxt <= 2) ? launch(stencil_bwd8_kernel<false, 2>) : launch(stencil_bwd8_kernel<false, 4>);
      }
      if (lrc) return lrc;
      count_launch();
      int rc8 = check_launch("stencil_bwd8_kernel");
      if (rc8) return rc8;
    }
    // weight / bias gradient
    if (!(parts & 2)) return PN_OK;
    {
      StencilBwdParams Q = P;
      int tw = 32;
      auto smem_of = [&](int t) { return ((size_t)3 * (t + 2) * PITCH + 4 + (size_t)8 * t * PITCH + 4) * sizeof(float); };
      while (tw > 1 && (tw * (P.D >> 3) > 2048 || smem_of(tw) > 200 * 1024)) tw >>= 1;
      while (tw > 1 && (tw >> 1) >= P.W) tw >>= 1;
      PN_REQUIRE(smem_of(tw) <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_feature_stencil_backward: depth %d too large for the weight gradient", P.D);
      Q.tw = tw;
      const bool flat = stage_flat() != 0, pre = flat && stencil_prefetch();
      const bool async = flat && !pre && stencil_async() && 2 * smem_of(tw) <= 200 * 1024;     // double buffer
      const size_t smem8 = async ? 2 * smem_of(tw) : smem_of(tw);
      const int nwork = Q.B * Q.H * ((Q.W + tw - 1) / tw);
      int ctas = (smem8 <= 110 * 1024) ? 148 * 2 : 148;
      if (ctas > nwork) ctas = nwork;
      auto launchw = [&](auto kern) -> int {
        PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
        PN_LAUNCH(kern, ctas, 256, smem8, stream, Q);
        return 0;
      };
      const int lrcw = pack ? (async ? launchw(stencil_wgrad8_kernel<true, true, false, true>) : pre ? launchw(stencil_wgrad8_kernel<true, true, true>) : flat ? launchw(stencil_wgrad8_kernel<true, true>) : launchw(stencil_wgrad8_kernel<true>))
                            : (async ? launchw(stencil_wgrad8_kernel<false, true, false, true>) : pre ? launchw(stencil_wgrad8_kernel<false, true, true>) : flat ? launchw(stencil_wgrad8_kernel<false, true>) : launchw(stencil_wgrad8_kernel<false>));
      if (lrcw) return lrcw;
      count_launch();
      return check_launch("stencil_wgrad8_kernel");
    }
  }
  P.tw = pick_tw(P.D, P.W, true, !pack);
  PN_REQUIRE(P.tw * P.D <= 2048, PN_ERR_UNSUPPORTED, "pn_feature_stencil_backward: depth %d too large", P.D);
  const size_t smem = stencil_smem_bytes(P.D, P.tw, true, !pack);
  PN_REQUIRE(smem <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_feature_stencil_backward: depth %d needs %zu bytes of shared memory", P.D, smem);
  dim3 grid((P.W + P.tw - 1) / P.tw, P.H, P.B);
  if (parts & 1) {
    if (pack) {
      PN_CUDA(cudaFuncSetAttribute(stencil_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PN_LAUNCH((stencil_bwd_kernel<true>), grid, 256, smem, stream, P);
    } else {
      PN_CUDA(cudaFuncSetAttribute(stencil_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PN_LAUNCH((stencil_bwd_kernel<false>), grid, 256, smem, stream, P);
    }
    count_launch();
    int rc = check_launch("stencil_bwd_kernel");
    if (rc) return rc;
  }
  if (!(parts & 2)) return PN_OK;
  // weight / bias gradient
  StencilBwdParams Q = P;
  Q.tw = 8;
  while (Q.tw > 1 && ((size_t)3 * (Q.tw + 2) * (Q.D + 2) + (size_t)8 * Q.tw * Q.D) * sizeof(float) > 200 * 1024) Q.tw >>= 1;
  if (Q.tw > Q.W) Q.tw = Q.W;
  const size_t smem_w = ((size_t)3 * (Q.tw + 2) * (Q.D + 2) + (size_t)8 * Q.tw * Q.D) * sizeof(float);
  PN_REQUIRE(smem_w <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_feature_stencil_backward: depth %d too large for the weight gradient", Q.D);
  int ctas = Q.B * Q.H;
  if (ctas > 148 * 2) ctas = 148 * 2;
  if (pack) {
    PN_CUDA(cudaFuncSetAttribute(stencil_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    PN_LAUNCH((stencil_wgrad_kernel<true>), ctas, 256, smem_w, stream, Q);
  } else {
    PN_CUDA(cudaFuncSetAttribute(stencil_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    PN_LAUNCH((stencil_wgrad_kernel<false>), ctas, 256, smem_w, stream, Q);
  }
  count_launch();
  return check_launch("stencil_wgrad_kernel");
}

extern "C" int pn_feature_stencil_backward(int pack, const float* in, const float* g, const float* w3, float* gin, float* gw3,
                                           float* gb3, int batch, int h_low, int w_low, int channels, int g_cstride, int g_coffset,
                                           pn_stream_t stream_) {
  return stencil_backward_impl(pack, in, g, w3, gin, gw3, gb3, batch, h_low, w_low, channels, g_cstride, g_coffset, 3,
                               reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int pn_feature_stencil_backward_parts(int pack, const float* in, const float* g, const float* w3, float* gin, float* gw3,
                                                 float* gb3, int batch, int h_low, int w_low, int channels, int g_cstride,
                                                 int g_coffset, int parts, pn_stream_t stream_) {
  return stencil_backward_impl(pack, in, g, w3, gin, gw3, gb3, batch, h_low, w_low, channels, g_cstride, g_coffset, parts,
                               reinterpret_cast<cudaStream_t>(stream_));
}

#ifndef PN_EMULATE
#include "gn_cluster_kernel.cuh"
#endif

static int groupnorm_elu_forward_impl(const float* x, const float* x2, const float* gamma, const float* beta, float eps, float* y,
                                      float* y_lo, double* stats, int batch, int hw, int channels, int out_cstride, int out_coffset,
                                      void* y_hi_bf16, void* y_lo_bf16, cudaStream_t stream) {
  PN_REQUIRE(x && gamma && beta && y && stats && batch > 0 && hw > 0, PN_ERR_BAD_ARGUMENT, "pn_groupnorm_elu_forward: bad argument");
  PN_REQUIRE(channels % 16 == 0 && channels % 4 == 0 && channels <= 1024, PN_ERR_UNSUPPORTED,
             "pn_groupnorm_elu_forward: channels %d (need a multiple of 16, <= 1024)", channels);
  PN_REQUIRE(out_cstride % 4 == 0 && out_coffset % 4 == 0, PN_ERR_ALIGNMENT, "pn_groupnorm_elu_forward: output channel window");
  PN_REQUIRE((!y_hi_bf16) == (!y_lo_bf16) && (!y_hi_bf16 || ((reinterpret_cast<uintptr_t>(y_hi_bf16) | reinterpret_cast<uintptr_t>(y_lo_bf16)) & 7) == 0),
             PN_ERR_BAD_ARGUMENT, "pn_groupnorm_elu_forward: the bf16 pair needs both 8-byte aligned pointers");
#ifndef PN_EMULATE
  {
    // small / medium maps (the tensor fits the L2): one cluster launch, statistics through distributed shared memory
    const GnClusterPlan plan = (y_lo || !aligned16(x) || (x2 && !aligned16(x2)) || !aligned16(y)) ? GnClusterPlan{0, 0, 0}
                                                                                                 : gn_cluster_plan(batch, hw, channels);
    if (plan.cl > 0) {
      GnClusterParams P{};
      P.x = x; P.x2 = x2; P.HW = hw; P.C = channels; P.cw = plan.cw; P.ppc = plan.ppc;
      P.gamma = gamma; P.beta = beta; P.eps = eps;
      P.mr = reinterpret_cast<float*>(stats + (size_t)2 * 16 * batch);
      P.y = y; P.out_cstride = out_cstride; P.out_coffset = out_coffset;
      P.hi = static_cast<uint2*>(y_hi_bf16); P.lo = static_cast<uint2*>(y_lo_bf16);
      const int rc = gn_cluster_launch(false, P, batch, plan.cl, stream);
      if (rc) return rc;
      return check_launch("gn_elu_cluster_fwd_kernel");
    }
  }
#endif
  PN_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * 16 * batch, stream));
  int ppc = (hw + 147) / 148;
  if (ppc < 32) ppc = 32;
  dim3 g1((hw + ppc - 1) / ppc, batch);
  if (gn_tree()) {
    PN_LAUNCH(gn_stats_kernel<true>, g1, 256, 0, stream, x, x2, hw, channels, channels, ppc, stats);
  } else {
    PN_LAUNCH(gn_stats_kernel<false>, g1, 256, 0, stream, x, x2, hw, channels, channels, ppc, stats);
  }
  count_launch();
  float* mr = reinterpret_cast<float*>(stats + (size_t)2 * 16 * batch);   // (mean, rstd) floats behind the doubles
  if (!y_lo && aligned16(x) && (!x2 || aligned16(x2)) && aligned16(y) && channels / 4 <= 256) {
    PN_LAUNCH(gn_elu_apply4_kernel, g1, 256, 0, stream, x, x2, hw, channels, mr, static_cast<const double*>(stats),
              (double)hw * (channels / 16), gamma, beta, eps, ppc, y, out_cstride, out_coffset, static_cast<uint2*>(y_hi_bf16),
              static_cast<uint2*>(y_lo_bf16));
    count_launch();
    return check_launch("gn_elu_apply4_kernel");
  }
  const size_t total = (size_t)batch * hw * (channels / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  PN_LAUNCH(gn_elu_apply_kernel, blocks, 256, sizeof(float) * 2 * 16 * batch, stream, x, x2, hw, channels, channels, mr,
            static_cast<const double*>(stats), (double)hw * (channels / 16), gamma, beta, eps, y, y_lo, out_cstride, out_coffset, batch,
            static_cast<uint2*>(y_hi_bf16), static_cast<uint2*>(y_lo_bf16));
  count_launch();
  return check_launch("gn_elu_apply_kernel");
}

extern "C" int pn_groupnorm_elu_forward(const float* x, const float* x2, const float* gamma, const float* beta, float eps, float* y,
                                        float* y_lo, double* stats, int batch, int hw, int channels, int out_cstride, int out_coffset,
                                        pn_stream_t stream_) {
  return groupnorm_elu_forward_impl(x, x2, gamma, beta, eps, y, y_lo, stats, batch, hw, channels, out_cstride, out_coffset, nullptr,
                                    nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int pn_groupnorm_elu_forward_split(const float* x, const float* x2, const float* gamma, const float* beta, float eps,
                                              float* y, void* y_hi_bf16, void* y_lo_bf16, double* stats, int batch, int hw,
                                              int channels, pn_stream_t stream_) {
  return groupnorm_elu_forward_impl(x, x2, gamma, beta, eps, y, nullptr, stats, batch, hw, channels, channels, 0, y_hi_bf16, y_lo_bf16,
                                    reinterpret_cast<cudaStream_t>(stream_));
}

static int groupnorm_elu_backward_impl(const float* x, const float* x2, const float* y, const float* dy, const float* gamma,
                                       float eps, const double* stats, double* bc, float* dx, float* dx_lo, float* dgamma,
                                       float* dbeta, float* dx_channel_sum, int batch, int hw, int channels, int y_cstride,
                                       int y_coffset, int dy_cstride, int dy_coffset, void* dx_hi_bf16, void* dx_lo_bf16,
                                       cudaStream_t stream) {
  PN_REQUIRE(x && y && dy && gamma && stats && bc && dx && dgamma && dbeta && batch > 0 && hw > 0, PN_ERR_BAD_ARGUMENT,
             "pn_groupnorm_elu_backward: bad argument");
  PN_REQUIRE(channels % 16 == 0 && channels <= 1024, PN_ERR_UNSUPPORTED, "pn_groupnorm_elu_backward: channels %d", channels);
  PN_REQUIRE((!dx_hi_bf16) == (!dx_lo_bf16), PN_ERR_BAD_ARGUMENT, "pn_groupnorm_elu_backward: the bf16 pair needs both pointers");
#ifndef PN_EMULATE
  {
    const bool vec0 = aligned16(x) && (!x2 || aligned16(x2)) && aligned16(y) && aligned16(dy) && aligned16(dx) && !dx_lo &&
                      y_cstride % 4 == 0 && y_coffset % 4 == 0 && dy_cstride % 4 == 0 && dy_coffset % 4 == 0;
    const GnClusterPlan plan = vec0 ? gn_cluster_plan(batch, hw, channels) : GnClusterPlan{0, 0, 0};
    if (plan.cl > 0) {
      // dgamma / dbeta / dsum meet over the samples with float atomics: one memset when the caller laid them out back to back
      if (dbeta == dgamma + channels && dx_channel_sum == dbeta + channels) {
        PN_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * 3 * channels, stream));
      } else {
        PN_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * channels, stream));
        PN_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * channels, stream));
        if (dx_channel_sum) PN_CUDA(cudaMemsetAsync(dx_channel_sum, 0, sizeof(float) * channels, stream));
      }
      GnClusterParams P{};
      P.x = x; P.x2 = x2; P.HW = hw; P.C = channels; P.cw = plan.cw; P.ppc = plan.ppc;
      P.gamma = gamma; P.eps = eps;
      P.mr = const_cast<float*>(reinterpret_cast<const float*>(stats + (size_t)2 * 16 * batch));
      P.yin = y; P.y_cstride = y_cstride; P.y_coffset = y_coffset;
      P.dy = dy; P.dy_cstride = dy_cstride; P.dy_coffset = dy_coffset;
      P.dx = dx; P.hi = static_cast<uint2*>(dx_hi_bf16); P.lo = static_cast<uint2*>(dx_lo_bf16);
      P.dgamma = dgamma; P.dbeta = dbeta; P.dsum = dx_channel_sum;
      const int rc = gn_cluster_launch(true, P, batch, plan.cl, stream);
      if (rc) return rc;
      return check_launch("gn_elu_cluster_bwd_kernel");
    }
  }
#endif
  PN_CUDA(cudaMemsetAsync(bc, 0, sizeof(double) * 2 * channels * batch, stream));
  int ppc = (hw + 147) / 148;
  if (ppc < 32) ppc = 32;
  dim3 g1((hw + ppc - 1) / ppc, batch);
  const float* mr = reinterpret_cast<const float*>(stats + (size_t)2 * 16 * batch);
  // bc holds 2*C*B doubles followed by 2*16*B floats of scratch for the group means
  float* gmeans = reinterpret_cast<float*>(bc + (size_t)2 * channels * batch);
  const bool vec = aligned16(x) && (!x2 || aligned16(x2)) && aligned16(y) && aligned16(dy) && aligned16(dx) && (!dx_lo || aligned16(dx_lo)) &&
                   y_cstride % 4 == 0 && y_coffset % 4 == 0 && dy_cstride % 4 == 0 && dy_coffset % 4 == 0;
  if (dx_channel_sum) PN_CUDA(cudaMemsetAsync(dx_channel_sum, 0, sizeof(float) * channels, stream));
  PN_REQUIRE((!dx_hi_bf16) == (!dx_lo_bf16) && (!dx_hi_bf16 || vec), PN_ERR_BAD_ARGUMENT,
             "pn_groupnorm_elu_backward: the bf16 pair needs both pointers and 16-byte aligned tensors");
  if (vec) {
    PN_LAUNCH(gn_elu_bwd_reduce4_kernel, g1, 256, 0, stream, x, x2, y, dy, hw, channels, channels, y_cstride, y_coffset, dy_cstride,
                                                      dy_coffset, mr, ppc, bc);
    count_launch();
    PN_LAUNCH(gn_elu_bwd_apply4_kernel, g1, 256, 0, stream, x, x2, y, dy, hw, channels, channels, y_cstride, y_coffset, dy_cstride, dy_coffset,
              mr, static_cast<const double*>(bc), (double)hw * (channels / 16), gamma, ppc, dx, dx_lo, dx_channel_sum, dgamma, dbeta, batch,
              static_cast<uint2*>(dx_hi_bf16), static_cast<uint2*>(dx_lo_bf16));
    count_launch();
    return check_launch("gn_elu_bwd kernels");
  }
  PN_LAUNCH(gn_elu_bwd_reduce_kernel, g1, 256, 0, stream, x, x2, y, dy, hw, channels, channels, y_cstride, y_coffset, dy_cstride, dy_coffset,
                                                   mr, eps, ppc, bc);
  count_launch();
  const size_t total = (size_t)batch * hw * channels;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  PN_LAUNCH(gn_bwd_finalize_kernel, (batch * 16 + channels + 127) / 128, 128, 0, stream, bc, gamma, batch, channels,
                                                                                 (double)hw * (channels / 16), gmeans, dgamma, dbeta);
  count_launch();
  PN_LAUNCH(gn_elu_bwd_apply_kernel, blocks, 256, 0, stream, x, x2, y, dy, hw, channels, channels, y_cstride, y_coffset, dy_cstride,
                                                      dy_coffset, mr, gmeans, gamma, eps, dx, dx_lo, batch);
  count_launch();
  if (dx_channel_sum) {
    int ppc2 = (int)(((size_t)batch * hw + 147 * 4) / (148 * 4));
    if (ppc2 < 64) ppc2 = 64;
    PN_LAUNCH(channel_sum_kernel, (int)(((size_t)batch * hw + ppc2 - 1) / ppc2), 256, 0, stream, dx, (size_t)batch * hw, channels, ppc2, dx_channel_sum);
    count_launch();
  }
  return check_launch("gn_elu_bwd kernels");
}

extern "C" int pn_groupnorm_elu_backward(const float* x, const float* x2, const float* y, const float* dy, const float* gamma,
                                         float eps, const double* stats, double* bc, float* dx, float* dx_lo, float* dgamma,
                                         float* dbeta, float* dx_channel_sum, int batch, int hw, int channels, int y_cstride,
                                         int y_coffset, int dy_cstride, int dy_coffset, pn_stream_t stream_) {
  return groupnorm_elu_backward_impl(x, x2, y, dy, gamma, eps, stats, bc, dx, dx_lo, dgamma, dbeta, dx_channel_sum, batch, hw, channels,
                                     y_cstride, y_coffset, dy_cstride, dy_coffset, nullptr, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int pn_groupnorm_elu_backward_split(const float* x, const float* x2, const float* y, const float* dy, const float* gamma,
                                               float eps, const double* stats, double* bc, float* dx, void* dx_hi_bf16, void* dx_lo_bf16,
                                               float* dgamma, float* dbeta, float* dx_channel_sum, int batch, int hw, int channels,
                                               pn_stream_t stream_) {
  return groupnorm_elu_backward_impl(x, x2, y, dy, gamma, eps, stats, bc, dx, nullptr, dgamma, dbeta, dx_channel_sum, batch, hw, channels,
                                     channels, 0, channels, 0, dx_hi_bf16, dx_lo_bf16, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int pn_conv2d_unpack_weight_grad_tiled(const float* dw_packed, float* dw_oihw, int cout, int cin, int ksize, int kpad,
                                                  pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(dw_packed && dw_oihw && cout > 0 && cin > 0 && ksize > 0 && kpad >= cin, PN_ERR_BAD_ARGUMENT,
             "pn_conv2d_unpack_weight_grad_tiled: bad argument");
  const int taps = ksize * ksize;
  const size_t smem = (size_t)UNPACK_CH * taps * sizeof(float);
  PN_REQUIRE(smem <= 48 * 1024 && cout <= 65535, PN_ERR_UNSUPPORTED, "pn_conv2d_unpack_weight_grad_tiled: kernel size %d / %d output channels",
             ksize, cout);
  dim3 grid((cin + UNPACK_CH - 1) / UNPACK_CH, cout);
  PN_LAUNCH(unpack_weight_grad_tiled_kernel, grid, 256, smem, stream, dw_packed, dw_oihw, cin, taps, kpad);
  count_launch();
  return check_launch("unpack_weight_grad_tiled_kernel");
}

extern "C" int pn_set_tuning(int key, int value) {
  PN_REQUIRE((key == PN_TUNE_STAGE_FLAT || key == PN_TUNE_GN_TREE) && (value == 0 || value == 1), PN_ERR_BAD_ARGUMENT,
             "pn_set_tuning: unknown key %d / value %d", key, value);
  (key == PN_TUNE_STAGE_FLAT ? g_stage_flat : g_gn_tree).store(value, std::memory_order_relaxed);
  return PN_OK;
}

extern "C" int pn_channel_sum(const float* g, float* out, size_t pixels, int channels, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(g && out && pixels > 0 && channels > 0 && channels <= 1024, PN_ERR_BAD_ARGUMENT, "pn_channel_sum: bad argument");
  PN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * channels, stream));
  int ppc = (int)((pixels + 147 * 4) / (148 * 4));
  if (ppc < 64) ppc = 64;
  PN_LAUNCH(channel_sum_kernel, (int)((pixels + ppc - 1) / ppc), 256, 0, stream, g, pixels, channels, ppc, out);
  count_launch();
  return check_launch("channel_sum_kernel");
}

static int head_tile_w(int C, int W) {
  int tw = 32;
  while (tw > 4 && ((size_t)10 * (tw + 2) * (C + SPAD) + 4 + (size_t)9 * C + 8 * tw) * sizeof(float) > 100 * 1024) tw >>= 1;
  while (tw > 4 && (tw >> 1) >= W) tw >>= 1;
  return tw;
}

// y[b,h,w] = bias + sum_{tap,c} x[b, h+dy-1, w+dx-1, c] * w[tap][c]      (Conv2d(C->1, 3x3, pad 1): layers01.py:110-116)
extern "C" int pn_head_conv_forward(const float* x, const float* w_tap_major, const float* bias, float* y, int batch, int height,
                                    int width, int channels, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(x && w_tap_major && bias && y && batch > 0 && height > 0 && width > 0 && channels > 0 && channels % 4 == 0,
             PN_ERR_BAD_ARGUMENT, "pn_head_conv_forward: bad argument (channels must be a multiple of 4)");
  PN_REQUIRE(aligned16(x) && aligned16(w_tap_major), PN_ERR_ALIGNMENT, "pn_head_conv_forward: pointers must be 16-byte aligned");
  HeadParams P{};
  P.B = batch; P.H = height; P.W = width; P.C = channels; P.x = x; P.w = w_tap_major; P.bias = bias; P.y = y;
  P.tw = head_tile_w(channels, width);
  const size_t smem = ((size_t)10 * (P.tw + 2) * (channels + SPAD) + 4 + (size_t)9 * channels) * sizeof(float);
  PN_REQUIRE(smem <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_head_conv_forward: %d channels need %zu bytes of shared memory", channels, smem);
  dim3 grid((width + P.tw - 1) / P.tw, (height + 7) / 8, batch);
  if (stage_flat()) {
    PN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PN_LAUNCH(head_fwd_kernel<true>, grid, 256, smem, stream, P);
  } else {
    PN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PN_LAUNCH(head_fwd_kernel<false>, grid, 256, smem, stream, P);
  }
  count_launch();
  return check_launch("head_fwd_kernel");
}

// dx [B,H,W,C], dw [9][C], db [1] from dy [B,H,W]
extern "C" int pn_head_conv_backward(const float* x, const float* dy, const float* w_tap_major, float* dx, float* dw_tap_major,
                                     float* dbias, int batch, int height, int width, int channels, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(x && dy && w_tap_major && dx && dw_tap_major && dbias && batch > 0 && height > 0 && width > 0 && channels > 0 &&
                 channels % 4 == 0, PN_ERR_BAD_ARGUMENT, "pn_head_conv_backward: bad argument");
  PN_REQUIRE(aligned16(x) && aligned16(w_tap_major) && aligned16(dx), PN_ERR_ALIGNMENT, "pn_head_conv_backward: pointers must be 16-byte aligned");
  PN_REQUIRE(9 * (channels / 4) <= 4 * 256, PN_ERR_UNSUPPORTED, "pn_head_conv_backward: %d channels (<= 452)", channels);
  HeadParams P{};
  P.B = batch; P.H = height; P.W = width; P.C = channels; P.x = x; P.w = w_tap_major; P.dy = dy; P.dx = dx; P.dw = dw_tap_major;
  P.db = dbias;
  P.tw = head_tile_w(channels, width);
  {
    const size_t smem = ((size_t)((10 * (P.tw + 2) + 3) & ~3) + (size_t)9 * channels) * sizeof(float);
    dim3 grid((width + P.tw - 1) / P.tw, (height + 7) / 8, batch);
    PN_CUDA(cudaFuncSetAttribute(head_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PN_LAUNCH(head_dgrad_kernel, grid, 256, smem, stream, P);
    count_launch();
    int rc = check_launch("head_dgrad_kernel");
    if (rc) return rc;
  }
  PN_CUDA(cudaMemsetAsync(dw_tap_major, 0, sizeof(float) * 9 * channels, stream));
  PN_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float), stream));
  const size_t smem = ((size_t)10 * (P.tw + 2) * (channels + SPAD) + 4 + (size_t)8 * P.tw) * sizeof(float);
  PN_REQUIRE(smem <= 227 * 1024, PN_ERR_UNSUPPORTED, "pn_head_conv_backward: %d channels need %zu bytes of shared memory", channels, smem);
  const int nwork = batch * ((height + 7) / 8) * ((width + P.tw - 1) / P.tw);
  int ctas = 148 * 2;
  if (ctas > nwork) ctas = nwork;
  const int maxq = (9 * (channels / 4) + 255) / 256;
  auto launch = [&](auto kern) -> int {
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PN_LAUNCH(kern, ctas, 256, smem, stream, P);
    return 0;
  };
  int lrc;
  if (stage_flat())
    lrc = (maxq <= 1) ? launch(head_wgrad_kernel<1, true>) : (maxq <= 2) ? launch(head_wgrad_kernel<2, true>) : launch(head_wgrad_kernel<4, true>);
  else
    lrc = (maxq <= 1) ? launch(head_wgrad_kernel<1, false>) : (maxq <= 2) ? launch(head_wgrad_kernel<2, false>) : launch(head_wgrad_kernel<4, false>);
  if (lrc) return lrc;
  count_launch();
  return check_launch("head_wgrad_kernel");
}
