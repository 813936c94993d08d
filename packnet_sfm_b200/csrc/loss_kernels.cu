// loss_kernels.cu -- fused photometric view-synthesis loss for sm_100a.
//
// One kernel does, per (scale, sample, 32x16 pixel tile):
//   inverse depth -> depth -> back-projection -> rigid transform -> projection -> bilinear gather
//   (zeros padding, align_corners) for every context frame, 3x3 reflection-padded SSIM + L1 against
//   the target for the warped AND the un-warped (auto-mask) candidates, per-pixel min / mean over the
//   candidates, edge-aware smoothness, and block -> global reduction of the loss.
// The backward kernel is the same tile program with a 2-pixel halo: it recomputes the forward values,
// then pushes dL/dloss through min -> SSIM/L1 -> bilinear sampling -> projection to d(inverse depth)
// (written per pixel) and d(pose) (reduced).  Nothing but the inputs and the per-pixel gradient
// touches HBM; the ~40 full-resolution intermediates per (scale, context) of the reference's op-by-op
// path live in shared memory and registers.
//
// Reference call sites replaced (paths relative to /root/reference/packnet_sfm):
//   losses/multiview_photometric_loss.py:14-53,127-344   geometry/camera.py:71-191
//   geometry/camera_utils.py:16-59   geometry/pose.py:80-86   utils/depth.py:103-198
//   utils/image.py:85-113,178-214
//
// Bit-exact warp indices: the coordinate chain below reproduces the reference's fp32 operation order
// with one IEEE rounding per op (__fmul_rn/__fadd_rn/__fdiv_rn are never contracted into FMAs).
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace pn {
namespace loss {

constexpr int TW = 32;   // tile width  (pixels)
constexpr int TH = 16;   // tile height (pixels)
constexpr int NT = 256;  // threads per CTA
constexpr int CAM_STRIDE_BASE = 18;  // Kinv[9], Kref[9]; then per context R[9], t[3]

struct ScaleParams {
  int h, w;
  int tiles_x, tiles_y, tile_base;
  const float* inv;                   // [B,1,h>>sh,w>>sh]: read through nearest up-sampling (index >> sh), model_utils.py:152-180
  int sh, iw;                         // log2 of that up-sampling factor; row pitch of the map (w >> sh)
  const float* img;                   // [B,3,h,w] target at this scale
  const float* ctx[PN_MAX_CONTEXT];   // [B,3,h,w] context frames at this scale
  float* ginv;                        // backward: [B,1,h>>sh,w>>sh]; sh > 0: pre-zeroed, the 2^sh x 2^sh block is summed (red.add)
  float photo_coef;                   // 1 / (B*h*w*n)            (x 1/#candidates in 'mean' mode)
  float sx_coef, sy_coef;             // smooth_w / (n*2^s) / (B*h*(w-1)),  .. / (B*(h-1)*w)
};

struct Params {
  int B, N, n;
  int total_tiles;
  int automask;
  float ssim_w, C1, C2;
  ScaleParams sc[PN_MAX_SCALES];
  const float* cams;        // [n][B][18 + 12N]
  const double* invsum;     // [n][B]  sum of inverse depth (prep kernel)
  double* photo_sum;        // [n]
  double* smooth_bs;        // [n][B]  smoothness loss share of (scale, sample), in loss units
  unsigned int* counter;    // last-block ticket
  float* out;               // [4]
  const float* grad_out;    // backward: device scalar
  float* gpose[PN_MAX_CONTEXT];  // backward: [B,4,4], pre-zeroed, atomically accumulated
};

// ---------------------------------------------------------------------------------------------------
// geometry: the reference's fp32 chain, one rounding per op
// ---------------------------------------------------------------------------------------------------
struct Projection {
  float ix, iy;      // unnormalised sampling coordinates (ATen grid_sampler_unnormalize, align_corners)
  float px, py, pz;  // K * (R X + t)
  float Z;           // clamp(pz, 1e-5)
  float X, Y, Zc;    // target-frame 3D point (camera.py:138)
};

__device__ __forceinline__ float dot3_rn(const float* __restrict__ m, float a, float b, float c) {
  // bmm row: (m0*a + m1*b) + m2*c, left to right, no FMA (camera.py:136,171,173; pose.py:84)
  return __fadd_rn(__fadd_rn(__fmul_rn(m[0], a), __fmul_rn(m[1], b)), __fmul_rn(m[2], c));
}

__device__ __forceinline__ void backproject(const float* __restrict__ Kinv, float u, float v, float depth,
                                            float& X, float& Y, float& Zc) {
  float rx = dot3_rn(Kinv + 0, u, v, 1.0f);
  float ry = dot3_rn(Kinv + 3, u, v, 1.0f);
  float rz = dot3_rn(Kinv + 6, u, v, 1.0f);
  X = __fmul_rn(rx, depth);   // camera.py:138  Xc = xnorm * depth ; Twc = identity (camera.py:145) is exact
  Y = __fmul_rn(ry, depth);
  Zc = __fmul_rn(rz, depth);
}

__device__ __forceinline__ Projection project_point(const float* __restrict__ Kref, const float* __restrict__ Rt,
                                                    float X, float Y, float Zc, float wm1, float hm1) {
  Projection p;
  p.X = X; p.Y = Y; p.Zc = Zc;
  float wx = __fadd_rn(dot3_rn(Rt + 0, X, Y, Zc), Rt[9]);    // pose.py:84-85
  float wy = __fadd_rn(dot3_rn(Rt + 3, X, Y, Zc), Rt[10]);
  float wz = __fadd_rn(dot3_rn(Rt + 6, X, Y, Zc), Rt[11]);
  p.px = dot3_rn(Kref + 0, wx, wy, wz);                      // camera.py:173
  p.py = dot3_rn(Kref + 3, wx, wy, wz);
  p.pz = dot3_rn(Kref + 6, wx, wy, wz);
  p.Z = fmaxf(p.pz, 1e-5f);                                   // camera.py:180
  float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(p.px, p.Z)), wm1), 1.0f);  // camera.py:181
  float yn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(p.py, p.Z)), hm1), 1.0f);  // camera.py:182
  p.ix = __fmul_rn(__fdiv_rn(__fadd_rn(xn, 1.0f), 2.0f), wm1);  // GridSampler.h grid_sampler_unnormalize
  p.iy = __fmul_rn(__fdiv_rn(__fadd_rn(yn, 1.0f), 2.0f), hm1);
  return p;
}

__device__ __forceinline__ float depth_from_inv(float inv) {
  return __fdiv_rn(1.0f, fmaxf(inv, 1e-6f));  // utils/depth.py:120
}

struct Taps {
  float x0f, y0f;
  bool nw, ne, sw, se;  // tap inside the image (padding_mode='zeros')
  int xi, yi;           // valid only where a tap is inside
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int w, int h) {
  Taps t;
  t.x0f = floorf(ix);
  t.y0f = floorf(iy);
  float x1f = t.x0f + 1.0f, y1f = t.y0f + 1.0f;
  bool x0in = (t.x0f >= 0.0f) && (t.x0f <= (float)(w - 1));
  bool x1in = (x1f >= 0.0f) && (x1f <= (float)(w - 1));
  bool y0in = (t.y0f >= 0.0f) && (t.y0f <= (float)(h - 1));
  bool y1in = (y1f >= 0.0f) && (y1f <= (float)(h - 1));
  t.nw = x0in && y0in; t.ne = x1in && y0in; t.sw = x0in && y1in; t.se = x1in && y1in;
  // every in-image tap has x0f in [-1, w-1]: the int cast is only used under those masks
  t.xi = (x0in || x1in) ? (int)t.x0f : 0;
  t.yi = (y0in || y1in) ? (int)t.y0f : 0;
  return t;
}

// bilinear sample of 3 channels (F.grid_sample bilinear / zeros / align_corners=True, camera_utils.py:58)
__device__ __forceinline__ void sample3(const float* __restrict__ src, size_t plane, int w, float ix, float iy,
                                        const Taps& t, float out[3]) {
  float x1f = t.x0f + 1.0f, y1f = t.y0f + 1.0f;
  float wnw = (x1f - ix) * (y1f - iy), wne = (ix - t.x0f) * (y1f - iy);
  float wsw = (x1f - ix) * (iy - t.y0f), wse = (ix - t.x0f) * (iy - t.y0f);
  const float* p = src + (ptrdiff_t)t.yi * w + t.xi;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pc = p + c * plane;
    float acc = 0.0f;
    if (t.nw) acc += __ldg(pc) * wnw;
    if (t.ne) acc += __ldg(pc + 1) * wne;
    if (t.sw) acc += __ldg(pc + w) * wsw;
    if (t.se) acc += __ldg(pc + w + 1) * wse;
    out[c] = acc;
  }
}

__device__ __forceinline__ int reflect_idx(int i, int n) {  // nn.ReflectionPad2d(1)
  return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

__device__ __forceinline__ float sgnf(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

__device__ __forceinline__ float block_sum(float v, float* red /* >= 8 floats */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < NT / 32) ? red[threadIdx.x] : 0.0f;
  if (warp == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

// ---------------------------------------------------------------------------------------------------
// prep: inverse-depth sums (for the per-sample mean, depth.py:160-162) and the per-(scale,sample)
// camera blocks (Camera.scaled / Kinv, camera.py:71-108; camera_utils.py:16-22)
// ---------------------------------------------------------------------------------------------------
struct PrepParams {
  int B, N, n, W;
  int h[PN_MAX_SCALES], w[PN_MAX_SCALES], sh[PN_MAX_SCALES];
  const float* inv[PN_MAX_SCALES];
  const float* K;
  const float* ref_K;
  const float* poses[PN_MAX_CONTEXT];
  float* cams;
  double* invsum;
};

__device__ void scale_K(const float* __restrict__ Kin, float sf, bool identity, float* Kout) {
#pragma unroll
  for (int i = 0; i < 9; ++i) Kout[i] = Kin[i];
  if (!identity) {  // camera_utils.py:16-22 on a clone (camera.py:104-108); scale 1 returns the camera itself
    Kout[0] = __fmul_rn(Kin[0], sf);
    Kout[4] = __fmul_rn(Kin[4], sf);
    Kout[2] = __fsub_rn(__fmul_rn(__fadd_rn(Kin[2], 0.5f), sf), 0.5f);
    Kout[5] = __fsub_rn(__fmul_rn(__fadd_rn(Kin[5], 0.5f), sf), 0.5f);
  }
}

__global__ void __launch_bounds__(256) loss_prep_kernel(PrepParams P) {
  __shared__ float red[8];
  const int sb = blockIdx.y;
  const int s = sb / P.B, b = sb % P.B;
  // the sum over the (virtually) up-sampled map = 4^sh x the sum over the stored one
  const int hw = (P.h[s] >> P.sh[s]) * (P.w[s] >> P.sh[s]);
  const float* inv = P.inv[s] + (size_t)b * hw;
  float acc = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) acc += __ldg(inv + i);
  float tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(P.invsum + sb, (double)tot * (double)(1 << (2 * P.sh[s])));
    if (blockIdx.x == 0) {
      const int stride = CAM_STRIDE_BASE + 12 * P.N;
      float* cam = P.cams + (size_t)sb * stride;
      // scale_factor = DW / float(W) (multiview_photometric_loss.py:156); torch multiplies in fp32
      const float sf = (float)((double)P.w[s] / (double)P.W);
      const bool identity = (P.w[s] == P.W);
      float Kt[9], Kr[9];
      scale_K(P.K + b * 9, sf, identity, Kt);
      scale_K(P.ref_K + b * 9, sf, identity, Kr);
      // Kinv: clone of K with four entries replaced (camera.py:75-80)
      float Ki[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) Ki[i] = Kt[i];
      Ki[0] = __fdiv_rn(1.0f, Kt[0]);
      Ki[4] = __fdiv_rn(1.0f, Kt[4]);
      Ki[2] = __fdiv_rn(__fmul_rn(-1.0f, Kt[2]), Kt[0]);
      Ki[5] = __fdiv_rn(__fmul_rn(-1.0f, Kt[5]), Kt[4]);
#pragma unroll
      for (int i = 0; i < 9; ++i) { cam[i] = Ki[i]; cam[9 + i] = Kr[i]; }
      for (int j = 0; j < P.N; ++j) {
        const float* T = P.poses[j] + b * 16;
        float* rt = cam + CAM_STRIDE_BASE + 12 * j;
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) rt[r * 3 + c] = T[r * 4 + c];
          rt[9 + r] = T[r * 4 + 3];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// the tile program
// ---------------------------------------------------------------------------------------------------
template <bool GRAD>
struct TileGeom {
  static constexpr int HALO = GRAD ? 2 : 1;
  static constexpr int RW = TW + 2 * HALO;
  static constexpr int RH = TH + 2 * HALO;
  static constexpr int RP = RW * RH;
  static constexpr int CW = TW + 2;  // coefficient region (tile + 1), backward only
  static constexpr int CH = TH + 2;
  static constexpr int CP = CW * CH;
};

template <int N, bool GRAD>
constexpr size_t tile_smem_floats() {
  using G = TileGeom<GRAD>;
  size_t f = (size_t)G::RP * (1 + 3 + 3 * N + 3 * N);  // inv, tgt, ref, warp
  if (GRAD) f += (size_t)G::CP * 10;                    // 9 coefficients + winner id
  return f + 16;
}

struct SsimTerms {
  float ssim, lossval;  // lossval = clamp((1-ssim)/2, 0, 1)
  float a, b, c;        // d ssim / d (mu_x, E[x^2], E[xy])
};

template <bool WITH_DERIV>
__device__ __forceinline__ SsimTerms ssim_from_sums(float sx, float sxx, float sxy, float sy, float syy, float C1,
                                                    float C2) {
  // multiview_photometric_loss.py:35-51 with avg_pool2d == window sum / 9
  const float mu_x = sx / 9.0f, mu_y = sy / 9.0f;
  const float mxy = mu_x * mu_y, mxx = mu_x * mu_x, myy = mu_y * mu_y;
  const float sig_x = sxx / 9.0f - mxx, sig_y = syy / 9.0f - myy, sig_xy = sxy / 9.0f - mxy;
  const float A1 = 2.0f * mxy + C1, A2 = 2.0f * sig_xy + C2;
  const float B1 = mxx + myy + C1, B2 = sig_x + sig_y + C2;
  const float Nn = A1 * A2, Dd = B1 * B2;
  SsimTerms t;
  t.ssim = Nn / Dd;
  t.lossval = fminf(fmaxf((1.0f - t.ssim) * 0.5f, 0.0f), 1.0f);
  if (WITH_DERIV) {
    const float invD = 1.0f / Dd;
    const float dN_dmu = 2.0f * mu_y * (A2 - A1);
    const float dD_dmu = 2.0f * mu_x * (B2 - B1);
    t.a = (dN_dmu - t.ssim * dD_dmu) * invD;
    t.b = -t.ssim * B1 * invD;
    t.c = 2.0f * A1 * invD;
  } else {
    t.a = t.b = t.c = 0.0f;
  }
  return t;
}

template <int N, bool MIN, bool GRAD>
__global__ void __launch_bounds__(NT) loss_tile_kernel(const Params P) {
  using G = TileGeom<GRAD>;
  constexpr int HALO = G::HALO, RW = G::RW, RP = G::RP;
  PN_DYNAMIC_SHARED(float, smem);
  float* s_inv = smem;
  float* s_tgt = s_inv + RP;            // [3][RP]
  float* s_ref = s_tgt + 3 * RP;        // [N][3][RP]
  float* s_warp = s_ref + 3 * N * RP;   // [N][3][RP]
  float* s_coef = s_warp + 3 * N * RP;  // [9][CP]      (GRAD)
  float* s_selj = s_coef + (GRAD ? 9 * G::CP : 0);  // [CP] winner context as float (-1: none) (GRAD)
  float* s_red = s_selj + (GRAD ? G::CP : 0);       // [16]

  // ---- which tile -------------------------------------------------------------------------------
  int s = 0;
#pragma unroll
  for (int i = 1; i < PN_MAX_SCALES; ++i)
    if (i < P.n && (int)blockIdx.x >= P.sc[i].tile_base) s = i;
  const ScaleParams& S = P.sc[s];
  const int b = blockIdx.y;
  const int tile = blockIdx.x - S.tile_base;
  const int tx0 = (tile % S.tiles_x) * TW, ty0 = (tile / S.tiles_x) * TH;
  const int h = S.h, w = S.w;
  const size_t plane = (size_t)h * w;
  const int rx0 = tx0 - HALO, ry0 = ty0 - HALO;
  const float wm1 = (float)(w - 1), hm1 = (float)(h - 1);

  const int cam_stride = CAM_STRIDE_BASE + 12 * P.N;
  const float* cam = P.cams + (size_t)(s * P.B + b) * cam_stride;
  __shared__ float s_cam[CAM_STRIDE_BASE + 12 * PN_MAX_CONTEXT];
  if (threadIdx.x < cam_stride) s_cam[threadIdx.x] = cam[threadIdx.x];
  __syncthreads();
  const float* Kinv = s_cam;
  const float* Kref = s_cam + 9;

  const float* inv_b = S.inv + (size_t)b * (plane >> (2 * S.sh));
  const float* img_b = S.img + (size_t)b * 3 * plane;

  // ---- phase 1: load the region, warp every context frame ------------------------------------------
  for (int idx = threadIdx.x; idx < RP; idx += NT) {
    const int i = idx % RW, j = idx / RW;
    const int x = rx0 + i, y = ry0 + j;
    const bool inside = (x >= 0) && (x < w) && (y >= 0) && (y < h);
    float inv = 0.0f, tg[3] = {0.0f, 0.0f, 0.0f};
    if (inside) {
      const size_t o = (size_t)y * w + x;
      inv = __ldg(inv_b + (size_t)(y >> S.sh) * S.iw + (x >> S.sh));
#pragma unroll
      for (int c = 0; c < 3; ++c) tg[c] = __ldg(img_b + c * plane + o);
    }
    s_inv[idx] = inv;
#pragma unroll
    for (int c = 0; c < 3; ++c) s_tgt[c * RP + idx] = tg[c];
    float X = 0.0f, Y = 0.0f, Zc = 0.0f;
    if (inside) backproject(Kinv, (float)x, (float)y, depth_from_inv(inv), X, Y, Zc);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float wv[3] = {0.0f, 0.0f, 0.0f}, rv[3] = {0.0f, 0.0f, 0.0f};
      if (inside) {
        const float* ctx_b = S.ctx[k] + (size_t)b * 3 * plane;
        const Projection pr = project_point(Kref, s_cam + CAM_STRIDE_BASE + 12 * k, X, Y, Zc, wm1, hm1);
        const Taps t = make_taps(pr.ix, pr.iy, w, h);
        sample3(ctx_b, plane, w, pr.ix, pr.iy, t, wv);
        if (P.automask) {
          const size_t o = (size_t)y * w + x;
#pragma unroll
          for (int c = 0; c < 3; ++c) rv[c] = __ldg(ctx_b + c * plane + o);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s_warp[(k * 3 + c) * RP + idx] = wv[c];
        s_ref[(k * 3 + c) * RP + idx] = rv[c];
      }
    }
  }
  __syncthreads();

  const float go = GRAD ? __ldg(P.grad_out) : 1.0f;
  const float wS = P.ssim_w / 3.0f, wL = (1.0f - P.ssim_w) / 3.0f;  // channel means, :214-216
  const int ncand = P.automask ? 2 * N : N;
  const float cand_w = MIN ? 1.0f : 1.0f / (float)ncand;               // 'mean': sum of means / len, :242

  // ---- phase 2: photometric candidates, min / mean, (GRAD) winner's SSIM derivative coefficients ----
  // region of pixels p whose photometric value is needed: the tile (forward) or tile+1 (backward)
  constexpr int PH = GRAD ? 1 : 0;
  constexpr int PW_ = TW + 2 * PH, PH_ = TH + 2 * PH;
  float photo_acc = 0.0f;
  // 'mean' mode treats the contexts one after the other (pass = context); 'min' needs them together
  const int passes = MIN ? 1 : N;
  float dR[GRAD ? N : 1][9], dT[GRAD ? N : 1][3];
  if (GRAD) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int i = 0; i < 9; ++i) dR[k][i] = 0.0f;
#pragma unroll
      for (int i = 0; i < 3; ++i) dT[k][i] = 0.0f;
    }
  }
  float ginv_acc[(TW * TH + NT - 1) / NT];  // per-thread gradient of the core pixels it owns (GRAD)
#pragma unroll
  for (int i = 0; i < (TW * TH + NT - 1) / NT; ++i) ginv_acc[i] = 0.0f;

  for (int pass = 0; pass < passes; ++pass) {
    for (int pidx = threadIdx.x; pidx < PW_ * PH_; pidx += NT) {
      const int pi = pidx % PW_, pj = pidx / PW_;
      const int x = tx0 - PH + pi, y = ty0 - PH + pj;
      const bool inside = (x >= 0) && (x < w) && (y >= 0) && (y < h);
      const bool core = inside && (x >= tx0) && (x < tx0 + TW) && (y >= ty0) && (y < ty0 + TH);
      float selj = -1.0f;
      float coef[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) coef[i] = 0.0f;
      if (inside) {
        int rc[3], rr[3];  // region columns / rows of the reflected 3x3 window
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          rc[d] = reflect_idx(x + d - 1, w) - rx0;
          rr[d] = (reflect_idx(y + d - 1, h) - ry0) * RW;
        }
        const int ctr = rr[1] + rc[1];
        float photo[2 * N];
#pragma unroll
        for (int q = 0; q < 2 * N; ++q) photo[q] = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float yv[9], sy = 0.0f, syy = 0.0f;
          const float* tg = s_tgt + c * RP;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float v = tg[rr[a] + rc[d]];
              yv[a * 3 + d] = v; sy += v; syy += v * v;
            }
#pragma unroll
          for (int q = 0; q < 2 * N; ++q) {
            const int k = q >> 1;
            const bool unwarped = (q & 1);
            if (unwarped && !P.automask) continue;
            if (!MIN && k != pass) continue;
            const float* src = (unwarped ? s_ref : s_warp) + (k * 3 + c) * RP;
            float sx = 0.0f, sxx = 0.0f, sxy = 0.0f;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                const float v = src[rr[a] + rc[d]];
                sx += v; sxx += v * v; sxy += v * yv[a * 3 + d];
              }
            const SsimTerms t = ssim_from_sums<false>(sx, sxx, sxy, sy, syy, P.C1, P.C2);
            photo[q] += wS * t.lossval + wL * fabsf(src[ctr] - yv[4]);
          }
        }
        // reduce over candidates: order [warp0, unwarp0, warp1, unwarp1, ...] (:326-334); first min wins
        int sel = -1;
        float val = 0.0f;
        if (MIN) {
          val = FLT_MAX;
#pragma unroll
          for (int q = 0; q < 2 * N; ++q) {
            if ((q & 1) && !P.automask) continue;
            if (photo[q] < val) { val = photo[q]; sel = q; }
          }
          if (sel < 0) { val = photo[0]; sel = 0; }  // all-NaN guard
        } else {
          sel = 2 * pass;
          val = photo[sel];
        }
        if (core) photo_acc += val * cand_w;
        if (GRAD && !(sel & 1)) {
          // winner is a warped frame: store d(loss)/d(ssim inputs) for the gather in phase 3
          const int k = sel >> 1;
          selj = (float)k;
          const float up = go * S.photo_coef * cand_w;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* tg = s_tgt + c * RP;
            const float* src = s_warp + (k * 3 + c) * RP;
            float sx = 0.0f, sxx = 0.0f, sxy = 0.0f, sy = 0.0f, syy = 0.0f;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                const float v = src[rr[a] + rc[d]], yv = tg[rr[a] + rc[d]];
                sx += v; sxx += v * v; sxy += v * yv; sy += yv; syy += yv * yv;
              }
            const SsimTerms t = ssim_from_sums<true>(sx, sxx, sxy, sy, syy, P.C1, P.C2);
            const float half = (1.0f - t.ssim) * 0.5f;
            const float g = (half >= 0.0f && half <= 1.0f) ? up * wS * (-0.5f) / 9.0f : 0.0f;
            coef[c * 3 + 0] = g * t.a;
            coef[c * 3 + 1] = g * 2.0f * t.b;
            coef[c * 3 + 2] = g * t.c;
          }
        }
      }
      if (GRAD) {
#pragma unroll
        for (int i = 0; i < 9; ++i) s_coef[i * G::CP + pidx] = coef[i];
        s_selj[pidx] = selj;
      }
    }
    if (GRAD) {
      __syncthreads();
      // ---- phase 3: gather the SSIM/L1 gradient onto the warped pixel, push through the sampler ----
      int slot = 0;
      for (int qidx = threadIdx.x; qidx < TW * TH; qidx += NT, ++slot) {
        const int qi = qidx % TW, qj = qidx / TW;
        const int x = tx0 + qi, y = ty0 + qj;
        if (x >= w || y >= h) continue;
        const int ridx = (qj + HALO) * RW + (qi + HALO);
        const float inv = s_inv[ridx];
        const float depth = depth_from_inv(inv);
        float X, Y, Zc;
        backproject(Kinv, (float)x, (float)y, depth, X, Y, Zc);
        float ddepth = 0.0f;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          if (!MIN && k != pass) continue;
          float Gc[3] = {0.0f, 0.0f, 0.0f};
          float xq[3], yq[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            xq[c] = s_warp[(k * 3 + c) * RP + ridx];
            yq[c] = s_tgt[c * RP + ridx];
          }
          bool any = false;
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int py = y + dy;
            if (py < 0 || py >= h) continue;
            const float my = ((py == 0 && y == 1 && dy == -1) || (py == h - 1 && y == h - 2 && dy == 1)) ? 2.0f : 1.0f;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int px = x + dx;
              if (px < 0 || px >= w) continue;
              const float mx = ((px == 0 && x == 1 && dx == -1) || (px == w - 1 && x == w - 2 && dx == 1)) ? 2.0f : 1.0f;
              const int cidx = (qj + 1 + dy) * G::CW + (qi + 1 + dx);
              if (s_selj[cidx] != (float)k) continue;
              any = true;
              const float m = mx * my;
#pragma unroll
              for (int c = 0; c < 3; ++c)
                Gc[c] += m * (s_coef[(c * 3 + 0) * G::CP + cidx] + s_coef[(c * 3 + 1) * G::CP + cidx] * xq[c] +
                              s_coef[(c * 3 + 2) * G::CP + cidx] * yq[c]);
            }
          }
          const int cself = (qj + 1) * G::CW + (qi + 1);
          if (s_selj[cself] == (float)k) {
            any = true;
            const float up = go * S.photo_coef * cand_w * wL;
#pragma unroll
            for (int c = 0; c < 3; ++c) Gc[c] += up * sgnf(xq[c] - yq[c]);
          }
          if (!any) continue;
          // d warped / d (ix, iy): grid_sampler_2d backward w.r.t. the grid, in-bounds taps only
          const float* Rt = s_cam + CAM_STRIDE_BASE + 12 * k;
          const Projection pr = project_point(Kref, Rt, X, Y, Zc, wm1, hm1);
          const Taps t = make_taps(pr.ix, pr.iy, w, h);
          const float x1f = t.x0f + 1.0f, y1f = t.y0f + 1.0f;
          const float* ctx_b = S.ctx[k] + (size_t)b * 3 * plane + (ptrdiff_t)t.yi * w + t.xi;
          float gix = 0.0f, giy = 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* pc = ctx_b + c * plane;
            const float g = Gc[c];
            if (t.nw) { const float v = __ldg(pc);         gix -= v * (y1f - pr.iy) * g; giy -= v * (x1f - pr.ix) * g; }
            if (t.ne) { const float v = __ldg(pc + 1);     gix += v * (y1f - pr.iy) * g; giy -= v * (pr.ix - t.x0f) * g; }
            if (t.sw) { const float v = __ldg(pc + w);     gix -= v * (pr.iy - t.y0f) * g; giy += v * (x1f - pr.ix) * g; }
            if (t.se) { const float v = __ldg(pc + w + 1); gix += v * (pr.iy - t.y0f) * g; giy += v * (pr.ix - t.x0f) * g; }
          }
          // ix == px/Z, iy == py/Z up to rounding (camera.py:181-182 then GridSampler unnormalise)
          const float iZ = 1.0f / pr.Z;
          const float dPx = gix * iZ, dPy = giy * iZ;
          const float dPz = (pr.pz >= 1e-5f) ? -(gix * pr.px + giy * pr.py) * iZ * iZ : 0.0f;
          float dXc[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) dXc[i] = Kref[0 + i] * dPx + Kref[3 + i] * dPy + Kref[6 + i] * dPz;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            dT[k][r] += dXc[r];
            dR[k][r * 3 + 0] += dXc[r] * X;
            dR[k][r * 3 + 1] += dXc[r] * Y;
            dR[k][r * 3 + 2] += dXc[r] * Zc;
          }
          const float dX = Rt[0] * dXc[0] + Rt[3] * dXc[1] + Rt[6] * dXc[2];
          const float dY = Rt[1] * dXc[0] + Rt[4] * dXc[1] + Rt[7] * dXc[2];
          const float dZ = Rt[2] * dXc[0] + Rt[5] * dXc[1] + Rt[8] * dXc[2];
          // X = ray * depth with ray = X / depth
          const float rx = dot3_rn(Kinv + 0, (float)x, (float)y, 1.0f), ry = dot3_rn(Kinv + 3, (float)x, (float)y, 1.0f),
                      rz = dot3_rn(Kinv + 6, (float)x, (float)y, 1.0f);
          ddepth += rx * dX + ry * dY + rz * dZ;
        }
        if (inv >= 1e-6f) ginv_acc[slot] += -ddepth * depth * depth;  // d(1/clamp(inv)) (depth.py:120)
      }
      __syncthreads();
    }
  }

  // ---- smoothness (utils/depth.py:146-198, multiview_photometric_loss.py:257-283) + gradient writes ----
  float smooth_acc = 0.0f;
  const bool do_smooth = (S.sx_coef != 0.0f) || (S.sy_coef != 0.0f);
  const double isum = P.invsum[s * P.B + b];
  const float mean = (float)(isum / (double)plane);
  const float mcl = fmaxf(mean, 1e-6f);
  const float inv_mcl = 1.0f / mcl;
  float bs_const = 0.0f;
  if (GRAD && do_smooth && mean >= 1e-6f)
    bs_const = -go * (float)(P.smooth_bs[s * P.B + b] / ((double)mcl * (double)plane));
  {
    int slot = 0;
    for (int qidx = threadIdx.x; qidx < TW * TH; qidx += NT, ++slot) {
      const int qi = qidx % TW, qj = qidx / TW;
      const int x = tx0 + qi, y = ty0 + qj;
      if (x >= w || y >= h) continue;
      const int ridx = (qj + HALO) * RW + (qi + HALO);
      float gq = 0.0f;
      if (do_smooth) {
        // explicit single roundings: an FMA-contracted (d0 - d1) would turn exact ties of the nearest-upsampled
        // scales into +-1 ulp noise and sign() of noise into a full-size gradient
        const float d0 = __fmul_rn(s_inv[ridx], inv_mcl);
        float i0[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) i0[c] = s_tgt[c * RP + ridx];
        if (x + 1 < w) {
          float ad = 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) ad += fabsf(i0[c] - s_tgt[c * RP + ridx + 1]);
          const float wx = expf(-ad / 3.0f);
          const float sx = __fsub_rn(d0, __fmul_rn(s_inv[ridx + 1], inv_mcl)) * wx;
          smooth_acc += fabsf(sx) * S.sx_coef;
          gq += sgnf(sx) * wx * S.sx_coef;
        }
        if (y + 1 < h) {
          float ad = 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) ad += fabsf(i0[c] - s_tgt[c * RP + ridx + RW]);
          const float wy = expf(-ad / 3.0f);
          const float sy = __fsub_rn(d0, __fmul_rn(s_inv[ridx + RW], inv_mcl)) * wy;
          smooth_acc += fabsf(sy) * S.sy_coef;
          gq += sgnf(sy) * wy * S.sy_coef;
        }
        if (GRAD) {
          if (x >= 1) {
            float ad = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) ad += fabsf(s_tgt[c * RP + ridx - 1] - i0[c]);
            const float wx = expf(-ad / 3.0f);
            const float sx = __fsub_rn(__fmul_rn(s_inv[ridx - 1], inv_mcl), d0) * wx;
            gq -= sgnf(sx) * wx * S.sx_coef;
          }
          if (y >= 1) {
            float ad = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) ad += fabsf(s_tgt[c * RP + ridx - RW] - i0[c]);
            const float wy = expf(-ad / 3.0f);
            const float sy = __fsub_rn(__fmul_rn(s_inv[ridx - RW], inv_mcl), d0) * wy;
            gq -= sgnf(sy) * wy * S.sy_coef;
          }
        }
      }
      if (GRAD) {
        const float gv = ginv_acc[slot] + go * gq * inv_mcl + bs_const;
        if (S.sh == 0) S.ginv[(size_t)b * plane + (size_t)y * w + x] = gv;
        else atomicAdd(S.ginv + (size_t)b * (plane >> (2 * S.sh)) + (size_t)(y >> S.sh) * S.iw + (x >> S.sh), gv);   // nearest backward
      }
    }
  }

  // ---- reductions ----------------------------------------------------------------------------------
  if (!GRAD) {
    const float ps = block_sum(photo_acc * S.photo_coef, s_red);
    const float ss = block_sum(smooth_acc, s_red);
    if (threadIdx.x == 0) {
      atomicAdd(P.photo_sum + s, (double)ps);
      if (do_smooth) atomicAdd(P.smooth_bs + s * P.B + b, (double)ss);
      __threadfence();
      const unsigned int ticket = atomicAdd(P.counter, 1u);
      if (ticket == (unsigned int)(P.total_tiles * P.B) - 1u) {
        __threadfence();
        double photo = 0.0, smooth = 0.0;
        for (int i = 0; i < P.n; ++i) photo += atomicAdd(P.photo_sum + i, 0.0);
        for (int i = 0; i < P.n * P.B; ++i) smooth += atomicAdd(P.smooth_bs + i, 0.0);
        const bool has_smooth = (P.sc[0].sx_coef != 0.0f) || (P.sc[0].sy_coef != 0.0f);
        const float loss = (float)(photo + smooth);
        P.out[0] = loss;
        // reference quirk: `loss += smoothness` is in place on the tensor aliased by the stored
        // 'photometric_loss' metric (multiview_photometric_loss.py:252,338; loss_base.py:72-74)
        P.out[1] = has_smooth ? loss : (float)photo;
        P.out[2] = (float)smooth;
        P.out[3] = 0.0f;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float* gp = P.gpose[k] + b * 16;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = block_sum(dR[k][r * 3 + c], s_red);
          if (threadIdx.x == 0 && v != 0.0f) atomicAdd(gp + r * 4 + c, v);
        }
        const float v = block_sum(dT[k][r], s_red);
        if (threadIdx.x == 0 && v != 0.0f) atomicAdd(gp + r * 4 + 3, v);
      }
    }
  }
}

}  // namespace loss
}  // namespace pn

#include "loss_group_kernel.cuh"   // the grouped-scale tile program (PN_LOSS_FLAG_GROUPED)

namespace pn {
namespace loss {

// ---------------------------------------------------------------------------------------------------
// inspection kernel: integer taps and float coordinates from the same device functions
// ---------------------------------------------------------------------------------------------------
// (GROUPED: through the grouped-scale program's copy of the coordinate chain, project_point_g)
template <bool GROUPED>
__global__ void warp_indices_kernel(const float* __restrict__ inv, const float* __restrict__ cams, int cam_stride,
                                    int B, int h, int w, int32_t* __restrict__ tap, float* __restrict__ coord) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = B * h * w;
  if (idx >= total) return;
  const int b = idx / (h * w), rem = idx % (h * w), y = rem / w, x = rem % w;
  const float* cam = cams + (size_t)b * cam_stride;
  float X, Y, Zc;
  backproject(cam, (float)x, (float)y, depth_from_inv(inv[idx]), X, Y, Zc);
  const Projection pr = GROUPED ? project_point_g(cam + 9, cam + CAM_STRIDE_BASE, X, Y, Zc, (float)(w - 1), (float)(h - 1))
                                : project_point(cam + 9, cam + CAM_STRIDE_BASE, X, Y, Zc, (float)(w - 1), (float)(h - 1));
  coord[2 * idx + 0] = pr.ix;
  coord[2 * idx + 1] = pr.iy;
  tap[2 * idx + 0] = (int32_t)floorf(fminf(fmaxf(pr.ix, -2.0e9f), 2.0e9f));
  tap[2 * idx + 1] = (int32_t)floorf(fminf(fmaxf(pr.iy, -2.0e9f), 2.0e9f));
}

// bilinear align_corners=True resize (F.interpolate, utils/image.py:145-146)
__global__ void resize_bilinear_ac_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes, int hi,
                                          int wi, int ho, int wo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)planes * ho * wo;
  if (idx >= total) return;
  const int x = (int)(idx % wo), y = (int)((idx / wo) % ho);
  const int p = (int)(idx / ((long long)wo * ho));
  // ATen area_pixel_compute_source_index(align_corners=True): src = scale * dst, scale = (in-1)/(out-1)
  const float sy_ = (ho > 1) ? (float)(hi - 1) / (float)(ho - 1) : 0.0f;
  const float sx_ = (wo > 1) ? (float)(wi - 1) / (float)(wo - 1) : 0.0f;
  const float fy = sy_ * (float)y, fx = sx_ * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + ((y0 < hi - 1) ? 1 : 0), x1 = x0 + ((x0 < wi - 1) ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float hy = 1.0f - ly, hx = 1.0f - lx;
  const float* s = src + (size_t)p * hi * wi;
  dst[idx] = hy * (hx * s[y0 * wi + x0] + lx * s[y0 * wi + x1]) + ly * (hx * s[y1 * wi + x0] + lx * s[y1 * wi + x1]);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct Workspace {
  // byte offsets into the caller's scratch buffer
  size_t invsum, photo_sum, smooth_bs, counter, cams, resized, total;
  size_t resized_img[PN_MAX_SCALES];                       // 0 when the scale uses the full-res tensors
  size_t resized_ctx[PN_MAX_SCALES][PN_MAX_CONTEXT];
};

static int validate(const pn_loss_desc* d) {
  PN_REQUIRE(d != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss: null descriptor");
  PN_REQUIRE(d->batch >= 1 && d->height >= 3 && d->width >= 3, PN_ERR_BAD_ARGUMENT, "pn_loss: bad image shape %dx%dx%d",
             d->batch, d->height, d->width);
  PN_REQUIRE(d->num_context >= 1 && d->num_context <= PN_MAX_CONTEXT, PN_ERR_BAD_ARGUMENT,
             "pn_loss: num_context %d outside 1..%d", d->num_context, PN_MAX_CONTEXT);
  PN_REQUIRE(d->num_scales >= 1 && d->num_scales <= PN_MAX_SCALES, PN_ERR_BAD_ARGUMENT,
             "pn_loss: num_scales %d outside 1..%d", d->num_scales, PN_MAX_SCALES);
  for (int i = 0; i < d->num_scales; ++i)
    PN_REQUIRE(d->scale_h[i] >= 3 && d->scale_w[i] >= 3 && d->scale_h[i] <= d->height && d->scale_w[i] <= d->width,
               PN_ERR_BAD_ARGUMENT, "pn_loss: scale %d has size %dx%d", i, d->scale_h[i], d->scale_w[i]);
  PN_REQUIRE(d->ssim_loss_weight > 0.0f && d->ssim_loss_weight <= 1.0f, PN_ERR_UNSUPPORTED,
             "pn_loss: ssim_loss_weight must be in (0,1] (the reference's SSIM-free branch keeps 3 channels, "
             "multiview_photometric_loss.py:216-217; not implemented)");
  PN_REQUIRE(!(d->automask && !d->reduce_min), PN_ERR_BAD_ARGUMENT,
             "pn_loss: automask requires photometric_reduce_op='min' (multiview_photometric_loss.py:112-114)");
  for (int i = 0; i < d->num_scales; ++i) {
    const int sh = d->inv_shift[i];
    PN_REQUIRE(sh >= 0 && sh <= 5 && (d->scale_h[i] & ((1 << sh) - 1)) == 0 && (d->scale_w[i] & ((1 << sh) - 1)) == 0,
               PN_ERR_BAD_ARGUMENT, "pn_loss: inv_shift[%d] = %d does not divide the %dx%d scale", i, sh, d->scale_h[i], d->scale_w[i]);
  }
  PN_REQUIRE(d->batch <= 65535, PN_ERR_UNSUPPORTED, "pn_loss: batch > 65535");
  PN_REQUIRE((d->flags & ~PN_LOSS_FLAG_GROUPED) == 0, PN_ERR_BAD_ARGUMENT, "pn_loss: unknown flags 0x%x", d->flags);
  return PN_OK;
}

static void layout(const pn_loss_desc* d, Workspace& ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  ws.invsum = take(sizeof(double) * PN_MAX_SCALES * d->batch);
  ws.photo_sum = take(sizeof(double) * PN_MAX_SCALES);
  ws.smooth_bs = take(sizeof(double) * PN_MAX_SCALES * d->batch);
  ws.counter = take(sizeof(unsigned int) * 4);
  ws.cams = take(sizeof(float) * PN_MAX_SCALES * d->batch * (CAM_STRIDE_BASE + 12 * d->num_context));
  ws.resized = off;
  for (int i = 0; i < PN_MAX_SCALES; ++i) {
    ws.resized_img[i] = 0;
    for (int j = 0; j < PN_MAX_CONTEXT; ++j) ws.resized_ctx[i][j] = 0;
  }
  for (int i = 0; i < d->num_scales; ++i) {
    if (d->scale_h[i] == d->height && d->scale_w[i] == d->width) continue;
    const size_t bytes = sizeof(float) * 3 * d->batch * d->scale_h[i] * d->scale_w[i];
    ws.resized_img[i] = take(bytes);
    for (int j = 0; j < d->num_context; ++j) ws.resized_ctx[i][j] = take(bytes);
  }
  ws.total = off;
}

template <int N, bool MIN, bool GRAD>
static int launch_tiles(const Params& P, dim3 grid, cudaStream_t stream) {
  const size_t smem = tile_smem_floats<N, GRAD>() * sizeof(float);
  auto kern = loss_tile_kernel<N, MIN, GRAD>;
  PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PN_LAUNCH(kern, grid, NT, smem, stream, P);
  count_launch();
  return check_launch("loss_tile_kernel");
}

template <bool GRAD>
static int dispatch_tiles(const pn_loss_desc* d, const Params& P, dim3 grid, cudaStream_t stream) {
#define PN_CASE(NN)                                                                       \
  case NN:                                                                                \
    return d->reduce_min ? launch_tiles<NN, true, GRAD>(P, grid, stream) : launch_tiles<NN, false, GRAD>(P, grid, stream);
  switch (d->num_context) {
    PN_CASE(1)
    PN_CASE(2)
    PN_CASE(3)
    PN_CASE(4)
  }
#undef PN_CASE
  set_error("pn_loss: unsupported num_context %d", d->num_context);
  return PN_ERR_UNSUPPORTED;
}

// ---- grouped-scale program: scales with the same size and images share a tile walk -----------------------------
template <bool GRAD>
static void make_groups(const Params& P, GParams& Q) {
  using GG = GroupGeom<GRAD>;
  Q.P = P;
  Q.ng = 0;
  int tile_base = 0;
  for (int i = 0; i < P.n; ++i) {
    const ScaleParams& S = P.sc[i];
    int gi = -1;
    for (int j = 0; j < Q.ng; ++j)
      if (Q.g[j].h == S.h && Q.g[j].w == S.w && P.sc[Q.g[j].scale[0]].img == S.img) gi = j;
    if (gi < 0) {
      gi = Q.ng++;
      GroupParams& G = Q.g[gi];
      G.h = S.h; G.w = S.w; G.ns = 0;
      G.tiles_x = (S.w + GG::CW - 1) / GG::CW;
      G.tiles_y = (S.h + GG::CH - 1) / GG::CH;
    }
    GroupParams& G = Q.g[gi];
    G.scale[G.ns++] = i;
  }
  for (int j = 0; j < Q.ng; ++j) {
    Q.g[j].tile_base = tile_base;
    tile_base += Q.g[j].tiles_x * Q.g[j].tiles_y;
  }
  Q.P.total_tiles = tile_base;
}

template <int N, bool MIN, bool GRAD, bool FUSED = false>
static int launch_groups(const GParams& Q, int batch, cudaStream_t stream) {
  const size_t smem = group_smem_floats<N, GRAD>() * sizeof(float);
#ifndef PN_EMULATE
  if (GRAD && FUSED && N == 2 && MIN) {      // A/B of the register target of the training launch (PN_LOSS_MINB=3: 3 CTAs per SM)
    static const int minb = [] { const char* e = std::getenv("PN_LOSS_MINB"); return (e && e[0] == '3') ? 3 : 2; }();
    if (minb == 3) {
      auto kern3 = loss_group_kernel<N, MIN, GRAD, FUSED, 3>;
      PN_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PN_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      PN_LAUNCH(kern3, dim3(Q.P.total_tiles, batch), GNT, smem, stream, Q);
      count_launch();
      return check_launch("loss_group_kernel");
    }
  }
#endif
  auto kern = loss_group_kernel<N, MIN, GRAD, FUSED>;
  PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PN_LAUNCH(kern, dim3(Q.P.total_tiles, batch), GNT, smem, stream, Q);
  count_launch();
  return check_launch("loss_group_kernel");
}

template <bool GRAD, bool FUSED = false>
static int dispatch_groups(const pn_loss_desc* d, const Params& P, cudaStream_t stream) {
  PN_REQUIRE(3LL * d->height * d->width < (1LL << 31), PN_ERR_UNSUPPORTED,
             "pn_loss (grouped): 3*H*W must fit 32-bit tap offsets");
  GParams Q{};
  make_groups<GRAD>(P, Q);
#define PN_CASE(NN)                                                                       \
  case NN:                                                                                \
    return d->reduce_min ? launch_groups<NN, true, GRAD, FUSED>(Q, d->batch, stream) : launch_groups<NN, false, GRAD, FUSED>(Q, d->batch, stream);
  switch (d->num_context) {
    PN_CASE(1)
    PN_CASE(2)
    PN_CASE(3)
    PN_CASE(4)
  }
#undef PN_CASE
  set_error("pn_loss: unsupported num_context %d", d->num_context);
  return PN_ERR_UNSUPPORTED;
}

// shared by forward and backward: pointers, per-scale images (resized when needed), prep kernel
static int setup(const pn_loss_desc* d, const float* image, const float* const* context, const float* const* inv_depths,
                 const float* K, const float* ref_K, const float* const* poses, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream, bool run_prep, Params& P, Workspace& ws) {
  int rc = validate(d);
  if (rc) return rc;
  PN_REQUIRE(image && context && inv_depths && K && ref_K && poses && workspace, PN_ERR_BAD_ARGUMENT,
             "pn_loss: null pointer argument");
  layout(d, ws);
  PN_REQUIRE(workspace_bytes >= ws.total, PN_ERR_WORKSPACE, "pn_loss: workspace %zu < required %zu", workspace_bytes,
             ws.total);
  PN_REQUIRE(aligned16(workspace) && aligned16(image) && aligned16(K) && aligned16(ref_K), PN_ERR_ALIGNMENT,
             "pn_loss: pointers must be 16-byte aligned");
  char* base = static_cast<char*>(workspace);
  P = Params{};
  P.B = d->batch; P.N = d->num_context; P.n = d->num_scales; P.automask = d->automask;
  P.ssim_w = d->ssim_loss_weight; P.C1 = d->C1; P.C2 = d->C2;
  P.cams = reinterpret_cast<float*>(base + ws.cams);
  P.invsum = reinterpret_cast<double*>(base + ws.invsum);
  P.photo_sum = reinterpret_cast<double*>(base + ws.photo_sum);
  P.smooth_bs = reinterpret_cast<double*>(base + ws.smooth_bs);
  P.counter = reinterpret_cast<unsigned int*>(base + ws.counter);
  const int ncand = d->automask ? 2 * d->num_context : d->num_context;
  (void)ncand;
  int tile_base = 0;
  for (int i = 0; i < d->num_scales; ++i) {
    ScaleParams& S = P.sc[i];
    S.h = d->scale_h[i]; S.w = d->scale_w[i];
    S.tiles_x = (S.w + TW - 1) / TW; S.tiles_y = (S.h + TH - 1) / TH; S.tile_base = tile_base;
    tile_base += S.tiles_x * S.tiles_y;
    PN_REQUIRE(inv_depths[i] != nullptr && aligned16(inv_depths[i]), PN_ERR_BAD_ARGUMENT, "pn_loss: inv_depths[%d]", i);
    S.inv = inv_depths[i];
    S.sh = d->inv_shift[i]; S.iw = S.w >> S.sh;
    const bool full = (S.h == d->height && S.w == d->width);
    S.img = full ? image : reinterpret_cast<const float*>(base + ws.resized_img[i]);
    for (int j = 0; j < d->num_context; ++j) {
      PN_REQUIRE(context[j] != nullptr && poses[j] != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss: context/poses[%d] null", j);
      S.ctx[j] = full ? context[j] : reinterpret_cast<const float*>(base + ws.resized_ctx[i][j]);
    }
    const double n = (double)d->num_scales, Bd = (double)d->batch;
    S.photo_coef = (float)(1.0 / (Bd * S.h * S.w * n));
    if (d->smooth_loss_weight > 0.0f) {
      const double sw = (double)d->smooth_loss_weight / (n * (double)(1 << i));
      S.sx_coef = (float)(sw / (Bd * S.h * (S.w - 1)));
      S.sy_coef = (float)(sw / (Bd * (S.h - 1) * S.w));
    } else {
      S.sx_coef = S.sy_coef = 0.0f;
    }
  }
  P.total_tiles = tile_base;
  if (!run_prep) return PN_OK;

  // zero the accumulators, resize images for reduced scales, camera blocks + inverse-depth sums
  PN_CUDA(cudaMemsetAsync(base, 0, ws.cams, stream));
  for (int i = 0; i < d->num_scales; ++i) {
    if (ws.resized_img[i] == 0) continue;
    const ScaleParams& S = P.sc[i];
    const long long total = 3LL * d->batch * S.h * S.w;
    const int blocks = (int)((total + 255) / 256);
    PN_LAUNCH(resize_bilinear_ac_kernel, blocks, 256, 0, stream, image, reinterpret_cast<float*>(base + ws.resized_img[i]),
                                                          3 * d->batch, d->height, d->width, S.h, S.w);
    count_launch();
    for (int j = 0; j < d->num_context; ++j) {
      PN_LAUNCH(resize_bilinear_ac_kernel, blocks, 256, 0, stream, context[j], reinterpret_cast<float*>(base + ws.resized_ctx[i][j]),
                                                            3 * d->batch, d->height, d->width, S.h, S.w);
      count_launch();
    }
  }
  PrepParams Q{};
  Q.B = d->batch; Q.N = d->num_context; Q.n = d->num_scales; Q.W = d->width;
  for (int i = 0; i < d->num_scales; ++i) { Q.h[i] = d->scale_h[i]; Q.w[i] = d->scale_w[i]; Q.sh[i] = d->inv_shift[i]; Q.inv[i] = inv_depths[i]; }
  Q.K = K; Q.ref_K = ref_K;
  for (int j = 0; j < d->num_context; ++j) Q.poses[j] = poses[j];
  Q.cams = reinterpret_cast<float*>(base + ws.cams);
  Q.invsum = reinterpret_cast<double*>(base + ws.invsum);
  int chunks = (d->height * d->width + 256 * 16 - 1) / (256 * 16);
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  PN_LAUNCH(loss_prep_kernel, dim3(chunks, d->num_scales * d->batch), 256, 0, stream, Q);
  count_launch();
  return check_launch("loss_prep_kernel");
}

}  // namespace loss
}  // namespace pn

using namespace pn;
using namespace pn::loss;

extern "C" int pn_loss_workspace_bytes(const pn_loss_desc* desc, size_t* bytes) {
  int rc = validate(desc);
  if (rc) return rc;
  PN_REQUIRE(bytes != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss_workspace_bytes: null out pointer");
  Workspace ws;
  layout(desc, ws);
  *bytes = ws.total;
  return PN_OK;
}

extern "C" int pn_loss_forward(const pn_loss_desc* desc, const float* image, const float* const* context,
                               const float* const* inv_depths, const float* K, const float* ref_K,
                               const float* const* poses, float* out, void* workspace, size_t workspace_bytes,
                               pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Params P;
  Workspace ws;
  int rc = setup(desc, image, context, inv_depths, K, ref_K, poses, workspace, workspace_bytes, stream, true, P, ws);
  if (rc) return rc;
  PN_REQUIRE(out != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss_forward: null out");
  P.out = out;
  if (desc->flags & PN_LOSS_FLAG_GROUPED) return dispatch_groups<false>(desc, P, stream);
  return dispatch_tiles<false>(desc, P, dim3(P.total_tiles, desc->batch), stream);
}

extern "C" int pn_loss_backward(const pn_loss_desc* desc, const float* image, const float* const* context,
                                const float* const* inv_depths, const float* K, const float* ref_K,
                                const float* const* poses, const float* grad_out, float* const* grad_inv_depths,
                                float* const* grad_poses, void* workspace, size_t workspace_bytes, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Params P;
  Workspace ws;
  // the workspace still holds the camera blocks, inverse-depth sums, resized images and the per-(scale,
  // sample) smoothness sums of the matching forward call
  int rc = setup(desc, image, context, inv_depths, K, ref_K, poses, workspace, workspace_bytes, stream, false, P, ws);
  if (rc) return rc;
  PN_REQUIRE(grad_out && grad_inv_depths && grad_poses, PN_ERR_BAD_ARGUMENT, "pn_loss_backward: null gradient pointer");
  P.grad_out = grad_out;
  for (int i = 0; i < desc->num_scales; ++i) {
    PN_REQUIRE(grad_inv_depths[i] != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss_backward: grad_inv_depths[%d] null", i);
    P.sc[i].ginv = grad_inv_depths[i];
    if (desc->inv_shift[i] > 0)   // the block sums of the nearest-upsample backward are accumulated
      PN_CUDA(cudaMemsetAsync(grad_inv_depths[i], 0, sizeof(float) * desc->batch * (desc->scale_h[i] >> desc->inv_shift[i]) *
                                                        (desc->scale_w[i] >> desc->inv_shift[i]), stream));
  }
  for (int j = 0; j < desc->num_context; ++j) {
    PN_REQUIRE(grad_poses[j] != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss_backward: grad_poses[%d] null", j);
    P.gpose[j] = grad_poses[j];
    PN_CUDA(cudaMemsetAsync(grad_poses[j], 0, sizeof(float) * 16 * desc->batch, stream));
  }
  if (desc->flags & PN_LOSS_FLAG_GROUPED) return dispatch_groups<true>(desc, P, stream);
  return dispatch_tiles<true>(desc, P, dim3(P.total_tiles, desc->batch), stream);
}

extern "C" int pn_loss_forward_backward(const pn_loss_desc* desc, const float* image, const float* const* context,
                                        const float* const* inv_depths, const float* K, const float* ref_K,
                                        const float* const* poses, float* out, float* const* unit_grad_inv_depths,
                                        float* const* unit_grad_poses, size_t grad_span_bytes, void* workspace,
                                        size_t workspace_bytes, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(desc && (desc->flags & PN_LOSS_FLAG_GROUPED), PN_ERR_UNSUPPORTED,
             "pn_loss_forward_backward: the one-launch training call exists for the grouped program (PN_LOSS_FLAG_GROUPED)");
  Params P;
  Workspace ws;
  int rc = setup(desc, image, context, inv_depths, K, ref_K, poses, workspace, workspace_bytes, stream, true, P, ws);
  if (rc) return rc;
  PN_REQUIRE(out && unit_grad_inv_depths && unit_grad_poses, PN_ERR_BAD_ARGUMENT, "pn_loss_forward_backward: null output pointer");
  P.out = out;
  for (int i = 0; i < desc->num_scales; ++i) {
    PN_REQUIRE(unit_grad_inv_depths[i] != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss_forward_backward: unit_grad_inv_depths[%d] null", i);
    P.sc[i].ginv = unit_grad_inv_depths[i];
  }
  for (int j = 0; j < desc->num_context; ++j) {
    PN_REQUIRE(unit_grad_poses[j] != nullptr, PN_ERR_BAD_ARGUMENT, "pn_loss_forward_backward: unit_grad_poses[%d] null", j);
    P.gpose[j] = unit_grad_poses[j];
  }
  if (grad_span_bytes) {
    // the caller laid every gradient output out inside one allocation starting at unit_grad_inv_depths[0]: one memset node
    PN_CUDA(cudaMemsetAsync(unit_grad_inv_depths[0], 0, grad_span_bytes, stream));
  } else {
    for (int i = 0; i < desc->num_scales; ++i)
      if (desc->inv_shift[i] > 0)
        PN_CUDA(cudaMemsetAsync(unit_grad_inv_depths[i], 0, sizeof(float) * desc->batch * (desc->scale_h[i] >> desc->inv_shift[i]) *
                                                                (desc->scale_w[i] >> desc->inv_shift[i]), stream));
    for (int j = 0; j < desc->num_context; ++j) PN_CUDA(cudaMemsetAsync(unit_grad_poses[j], 0, sizeof(float) * 16 * desc->batch, stream));
  }
  return dispatch_groups<true, true>(desc, P, stream);
}

extern "C" int pn_loss_backward_finish(const pn_loss_desc* desc, const float* grad_out, const float* const* unit_grad_inv_depths,
                                       const float* const* unit_grad_poses, float* const* grad_inv_depths, float* const* grad_poses,
                                       void* workspace, size_t workspace_bytes, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate(desc);
  if (rc) return rc;
  PN_REQUIRE(grad_out && unit_grad_inv_depths && unit_grad_poses && grad_inv_depths && grad_poses && workspace, PN_ERR_BAD_ARGUMENT,
             "pn_loss_backward_finish: null pointer argument");
  Workspace ws;
  layout(desc, ws);
  PN_REQUIRE(workspace_bytes >= ws.total, PN_ERR_WORKSPACE, "pn_loss_backward_finish: workspace %zu < required %zu", workspace_bytes, ws.total);
  char* base = static_cast<char*>(workspace);
  FinishParams F{};
  F.B = desc->batch; F.N = desc->num_context; F.n = desc->num_scales;
  int maxcount = 1;
  for (int i = 0; i < desc->num_scales; ++i) {
    PN_REQUIRE(unit_grad_inv_depths[i] && grad_inv_depths[i], PN_ERR_BAD_ARGUMENT, "pn_loss_backward_finish: gradient pointer %d null", i);
    const int sh = desc->inv_shift[i];
    F.raw[i] = unit_grad_inv_depths[i]; F.out[i] = grad_inv_depths[i];
    F.count[i] = (desc->scale_h[i] >> sh) * (desc->scale_w[i] >> sh);
    F.plane[i] = (double)desc->scale_h[i] * desc->scale_w[i];
    F.block[i] = (float)(1 << (2 * sh));
    F.smooth[i] = desc->smooth_loss_weight > 0.0f ? 1 : 0;
    if (F.count[i] > maxcount) maxcount = F.count[i];
  }
  for (int j = 0; j < desc->num_context; ++j) {
    PN_REQUIRE(unit_grad_poses[j] && grad_poses[j], PN_ERR_BAD_ARGUMENT, "pn_loss_backward_finish: pose gradient pointer %d null", j);
    F.raw_pose[j] = unit_grad_poses[j]; F.out_pose[j] = grad_poses[j];
  }
  F.invsum = reinterpret_cast<const double*>(base + ws.invsum);
  F.smooth_bs = reinterpret_cast<const double*>(base + ws.smooth_bs);
  F.grad_out = grad_out;
  int bx = (maxcount + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  PN_LAUNCH(loss_grad_finish_kernel, dim3(bx, desc->num_scales * desc->batch), 256, 0, stream, F);
  count_launch();
  return check_launch("loss_grad_finish_kernel");
}

extern "C" int pn_loss_warp_indices(const pn_loss_desc* desc, int scale, const float* inv_depth, const float* K,
                                    const float* ref_K, const float* pose, int32_t* tap_xy, float* coord_xy,
                                    void* workspace, size_t workspace_bytes, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate(desc);
  if (rc) return rc;
  PN_REQUIRE(scale >= 0 && scale < desc->num_scales, PN_ERR_BAD_ARGUMENT, "pn_loss_warp_indices: scale %d", scale);
  PN_REQUIRE(inv_depth && K && ref_K && pose && tap_xy && coord_xy && workspace, PN_ERR_BAD_ARGUMENT,
             "pn_loss_warp_indices: null pointer");
  // run the prep kernel for a single-scale, single-context view of the problem
  pn_loss_desc one = *desc;
  one.num_scales = 1; one.num_context = 1;
  one.scale_h[0] = desc->scale_h[scale]; one.scale_w[0] = desc->scale_w[scale];
  one.inv_shift[0] = 0;      // the hook takes the map at the scale's own resolution
  Workspace ws;
  layout(&one, ws);
  PN_REQUIRE(workspace_bytes >= ws.total, PN_ERR_WORKSPACE, "pn_loss_warp_indices: workspace %zu < %zu", workspace_bytes,
             ws.total);
  char* base = static_cast<char*>(workspace);
  PN_CUDA(cudaMemsetAsync(base, 0, ws.cams, stream));
  PrepParams Q{};
  Q.B = one.batch; Q.N = 1; Q.n = 1; Q.W = one.width;
  Q.h[0] = one.scale_h[0]; Q.w[0] = one.scale_w[0]; Q.sh[0] = 0; Q.inv[0] = inv_depth;
  Q.K = K; Q.ref_K = ref_K; Q.poses[0] = pose;
  Q.cams = reinterpret_cast<float*>(base + ws.cams);
  Q.invsum = reinterpret_cast<double*>(base + ws.invsum);
  PN_LAUNCH(loss_prep_kernel, dim3(1, one.batch), 256, 0, stream, Q);
  count_launch();
  const int total = one.batch * one.scale_h[0] * one.scale_w[0];
  if (desc->flags & PN_LOSS_FLAG_GROUPED) {
    PN_LAUNCH(warp_indices_kernel<true>, (total + 255) / 256, 256, 0, stream, inv_depth, Q.cams, CAM_STRIDE_BASE + 12, one.batch,
                                                                      one.scale_h[0], one.scale_w[0], tap_xy, coord_xy);
  } else {
    PN_LAUNCH(warp_indices_kernel<false>, (total + 255) / 256, 256, 0, stream, inv_depth, Q.cams, CAM_STRIDE_BASE + 12, one.batch,
                                                                       one.scale_h[0], one.scale_w[0], tap_xy, coord_xy);
  }
  count_launch();
  return check_launch("warp_indices_kernel");
}

extern "C" int pn_resize_bilinear_ac(const float* src, float* dst, int batch, int channels, int h_in, int w_in,
                                     int h_out, int w_out, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(src && dst && batch > 0 && channels > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0,
             PN_ERR_BAD_ARGUMENT, "pn_resize_bilinear_ac: bad argument");
  const long long total = (long long)batch * channels * h_out * w_out;
  PN_LAUNCH(resize_bilinear_ac_kernel, (int)((total + 255) / 256), 256, 0, stream, src, dst, batch * channels, h_in, w_in, h_out,
                                                                            w_out);
  count_launch();
  return check_launch("resize_bilinear_ac_kernel");
}
