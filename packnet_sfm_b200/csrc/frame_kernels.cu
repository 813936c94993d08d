// frame_kernels.cu -- the exact frame terms of the folded pack block (packnet_sfm_b200/folded.py), sm_100a.
//
// One (k+2)x(k+2) convolution of the space-to-depth tensor reproduces PackLayerConv3d (layers01.py:239-247) except on
// the frame of width m = k/2, where the reference's zero padding BETWEEN its Conv3d and Conv2d matters.  The repair is
// a sum of up to eight thin linear terms -- four ring rows / columns and four ring corners -- each of the form
//
//     z[b, row(a,l), col(a,l), co] += alpha * sum_{e < KE, nn < n}  w[co][a][e][nn] * line[b][l + e - pad][nn]
//
// where `line` is a border row / column of the space-to-depth tensor ([B][L][n], zero beyond both ends), `w` the folded
// weight of the term (pn_pack_fold_forward, channels-last) and (a, l) -> (row, col) an affine map into the frame; plus
// the Conv3d-bias correction dB[class(row)][class(col)][co], which depends on a frame pixel only through its border class.
// These are small fp32 GEMMs (M = B*L <= ~1300 pixels, K = KE*n, N = A*Co) in which every operand is contiguous along
// one GEMM index; all terms of a layer run in ONE launch (blockIdx.z = term), and the backward is one launch for the
// border-line gradients, one for the folded-weight gradients and one for the bias-class sums.
//
// fp32 FMA on CUDA cores on purpose: the terms are O(perimeter) -- 1.2 GMAC per forward for pack1..3 at B=4, 192x640,
// against 250 GMAC for the folded convolutions -- and must stay exact fp32 (they carry O(1) of the border pixels).
#include <algorithm>

#include "common.cuh"

namespace pn {
namespace frame {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct Term {
  const float* line;
  const float* w;
  float* dline;
  float* dw;
  long long line_bs, dline_bs;      // batch strides (floats)
  long long w_sco, w_sa, w_se;      // w element (co, a, e, nn) at co*w_sco + a*w_sa + e*w_se + nn
  int L, A, A2, KE, pad;
  int r0, ra1, ra2, rl;             // row = r0 + (a / A2) * ra1 + (a % A2) * ra2 + l * rl
  int c0, ca1, ca2, cl;             // col likewise
  float alpha;
  int bias_mode;                    // 0: none, 1: add dB on every pixel of the term, 2: only rows in [m, h - m)
};

struct Params {
  int B, h, w, Co, n, m, nterms;
  int ksplit;                       // CTAs per output tile along the middle reduction index (taps e forward / line backward, samples for the weights)
  Term t[8];
  const float* dB;                  // [2m+1][2m+1][Co]
  float* z;                         // forward: [B][h][w][Co], accumulated
  const float* gz;                  // backward
  float* gdB;                       // backward: [2m+1][2m+1][Co], atomically accumulated
};

__device__ __forceinline__ int border_class(int p, int len, int m) { return p < m ? p : (p >= len - m ? p - (len - m) + m + 1 : m); }


// (a / A2, a % A2) without a division for the two cases that occur (A2 = 1: sides, A2 = m = 2: corners of a 5x5 kernel)
__device__ __forceinline__ void split_a(int a, int A2, int* a1, int* a2) {
  if (A2 == 1) { *a1 = a; *a2 = 0; }
  else if (A2 == 2) { *a1 = a >> 1; *a2 = a & 1; }
  else { *a1 = a / A2; *a2 = a - *a1 * A2; }
}

__device__ __forceinline__ long long z_offset(const Params& P, const Term& T, int b, int a, int l, int* row, int* col) {
  int a1, a2;
  split_a(a, T.A2, &a1, &a2);
  const int r = T.r0 + a1 * T.ra1 + a2 * T.ra2 + l * T.rl;
  const int c = T.c0 + a1 * T.ca1 + a2 * T.ca2 + l * T.cl;
  *row = r;
  *col = c;
  return (((long long)b * P.h + r) * P.w + c) * P.Co;
}

// One 64x64 tile of C = A * B.  The reduction index is a triple (k2, k1, k0) with k0 the contiguous one, walked in steps
// of BK: the loaders never divide -- a thread decomposes ITS rows / columns once (ra / rb -> context) and receives the
// reduction triple ready-made.  A_KFAST / B_KFAST: whether consecutive threads of the tile load run along k0 (the operand
// is contiguous in the reduction index) or along the row / column index.
template <bool A_KFAST, bool B_KFAST, class RA, class LA, class RB, class LB>
__device__ __forceinline__ void gemm_tile(int K2, int K1b, int K1e, int K0, int m0, int n0, RA ra, LA la, RB rb, LB lb, float (&acc)[4][4]) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;
  constexpr int LD = (BM * BK) / NT;                     // 4 loads per thread and operand
  // fixed per-thread assignment of the tile loads
  int a_kk[LD], a_mm[LD], b_kk[LD], b_nn[LD];
#pragma unroll
  for (int i = 0; i < LD; ++i) {
    const int idx = t + i * NT;
    a_kk[i] = A_KFAST ? idx % BK : idx / BM;
    a_mm[i] = A_KFAST ? idx / BK : idx % BM;
    b_kk[i] = B_KFAST ? idx % BK : idx / BN;
    b_nn[i] = B_KFAST ? idx / BK : idx % BN;
  }
  auto ca0 = ra(m0 + a_mm[0]);
  decltype(ca0) ca[LD] = {ca0, ra(m0 + a_mm[1]), ra(m0 + a_mm[2]), ra(m0 + a_mm[3])};
  auto cb0 = rb(n0 + b_nn[0]);
  decltype(cb0) cb[LD] = {cb0, rb(n0 + b_nn[1]), rb(n0 + b_nn[2]), rb(n0 + b_nn[3])};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k2 = 0; k2 < K2; ++k2)
    for (int k1 = K1b; k1 < K1e; ++k1)
      for (int kb = 0; kb < K0; kb += BK) {
        // all eight global loads of the stage first, then the shared stores: left interleaved (load, store, load, ...) every
        // store waits for its own load and the stage costs eight serial round trips instead of one.
        // (Round 2 tried a register-prefetch pipeline over the stages + float4 operand reads: forward 0.45 -> 0.55 ms on the B200,
        // gpurun r02v -- the three launches are a few hundred CTAs of 2 waves each; reverted.)
        float av[LD], bv[LD];
#pragma unroll
        for (int i = 0; i < LD; ++i) {
          const int k0 = kb + a_kk[i];
          av[i] = (k0 < K0) ? la(ca[i], k2, k1, k0) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < LD; ++i) {
          const int k0 = kb + b_kk[i];
          bv[i] = (k0 < K0) ? lb(cb[i], k2, k1, k0) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < LD; ++i) As[a_kk[i]][a_mm[i]] = av[i];
#pragma unroll
        for (int i = 0; i < LD; ++i) Bs[b_kk[i]][b_nn[i]] = bv[i];
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
}

struct PixelCtx { int valid, b, l; };                    // a row of the border line: pixel l of sample b
struct WeightColCtx { int valid, a, co; long long off; };  // an output column (a, co) and its offset co*w_sco + a*w_sa
struct ChannelCtx { int valid, nn; };
struct WindowCtx { int valid, e, nn; };                  // a column (e, nn) of the line window

__device__ __forceinline__ PixelCtx pixel_ctx(const Term& T, int M, int p) {
  PixelCtx c;
  c.valid = p < M;
  c.b = c.valid ? p / T.L : 0;
  c.l = p - c.b * T.L;
  return c;
}
__device__ __forceinline__ WeightColCtx weight_col_ctx(const Params& P, const Term& T, int N, int c) {
  WeightColCtx w;
  w.valid = c < N;
  w.a = w.valid ? c / P.Co : 0;
  w.co = c - w.a * P.Co;
  w.off = (long long)w.co * T.w_sco + (long long)w.a * T.w_sa;
  return w;
}
__device__ __forceinline__ float line_at(const Params& P, const Term& T, int b, int pos, int nn) {
  if (pos < 0 || pos >= T.L) return 0.f;
  return __ldg(T.line + (long long)b * T.line_bs + (long long)pos * P.n + nn);
}

// forward: M = B*L pixels, N = A*Co columns (a, co), reduction (e, nn)
__global__ void __launch_bounds__(NT) frame_forward_kernel(const __grid_constant__ Params P) {
  // K split (round 2): the launch was ~300 CTAs walking 112 sixteen-wide stages each with one exposed L2 round trip per stage
  // (0.45 ms for 0.4 GMAC); one CTA per (tile, tap e) gives 7x the CTAs with 16 stages each, partial sums meet in the atomicAdd
  // the epilogue already used.
  const Term& T = P.t[blockIdx.z];
  const int M = P.B * T.L, N = T.A * P.Co;
  const int ks = blockIdx.x % P.ksplit;
  const int m0 = blockIdx.y * BM, n0 = (blockIdx.x / P.ksplit) * BN;
  if (m0 >= M || n0 >= N || ks >= T.KE) return;
  const int e_per = (T.KE + P.ksplit - 1) / P.ksplit, e0 = ks * e_per, e1 = min(T.KE, e0 + e_per);
  if (e0 >= e1) return;
  float acc[4][4];
  gemm_tile<true, true>(
      1, e0, e1, P.n, m0, n0,
      [&](int p) { return pixel_ctx(T, M, p); },
      [&](const PixelCtx& c, int, int e, int nn) -> float { return c.valid ? line_at(P, T, c.b, c.l + e - T.pad, nn) : 0.f; },
      [&](int c) { return weight_col_ctx(P, T, N, c); },
      [&](const WeightColCtx& c, int, int e, int nn) -> float { return c.valid ? __ldg(T.w + c.off + (long long)e * T.w_se + nn) : 0.f; },
      acc);
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  const int g = 2 * P.m + 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = m0 + ty * 4 + i;
    if (p >= M) continue;
    const int b = p / T.L, l = p - b * T.L;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= N) continue;
      const int a = c / P.Co, co = c - a * P.Co;
      int row, col;
      const long long off = z_offset(P, T, b, a, l, &row, &col);
      float v = T.alpha * acc[i][j];
      if (ks == 0 && (T.bias_mode == 1 || (T.bias_mode == 2 && row >= P.m && row < P.h - P.m)))
        v += __ldg(P.dB + ((long long)border_class(row, P.h, P.m) * g + border_class(col, P.w, P.m)) * P.Co + co);
      atomicAdd(P.z + off + co, v);
    }
  }
}

// backward, border lines: dline[b][j][nn] += alpha * sum_{a, e, co} gz[b, row(a,l), col(a,l), co] * w[co][a][e][nn], l = j - e + pad
// M = B*L pixels j, N = n, reduction (a, e, co)
__global__ void __launch_bounds__(NT) frame_backward_line_kernel(const __grid_constant__ Params P) {
  const Term& T = P.t[blockIdx.z];
  const int M = P.B * T.L, N = P.n;
  const int ks = blockIdx.x % P.ksplit;
  const int m0 = blockIdx.y * BM, n0 = (blockIdx.x / P.ksplit) * BN;
  if (m0 >= M || n0 >= N) return;
  const int e_per = (T.KE + P.ksplit - 1) / P.ksplit, e0 = ks * e_per, e1 = min(T.KE, e0 + e_per);
  if (e0 >= e1) return;
  float acc[4][4];
  gemm_tile<true, false>(
      T.A, e0, e1, P.Co, m0, n0,
      [&](int p) { return pixel_ctx(T, M, p); },
      [&](const PixelCtx& c, int a, int e, int co) -> float {
        const int l = c.l - e + T.pad;
        if (!c.valid || l < 0 || l >= T.L) return 0.f;
        int row, col;
        return __ldg(P.gz + z_offset(P, T, c.b, a, l, &row, &col) + co);
      },
      [&](int nn) { ChannelCtx c; c.valid = nn < N; c.nn = nn; return c; },
      [&](const ChannelCtx& c, int a, int e, int co) -> float {
        return c.valid ? __ldg(T.w + (long long)co * T.w_sco + (long long)a * T.w_sa + (long long)e * T.w_se + c.nn) : 0.f;
      },
      acc);
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = m0 + ty * 4 + i;
    if (p >= M) continue;
    const int b = p / T.L, j = p - b * T.L;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int nn = n0 + tx * 4 + jn;
      if (nn < N) atomicAdd(T.dline + (long long)b * T.dline_bs + (long long)j * P.n + nn, T.alpha * acc[i][jn]);
    }
  }
}

// backward, folded weights: dw[co][a][e][nn] = alpha * sum_{b, l} gz[b, row(a,l), col(a,l), co] * line[b][l + e - pad][nn]
// M = A*Co rows (a, co), N = KE*n columns (e, nn), reduction (b, l)
__global__ void __launch_bounds__(NT) frame_backward_weight_kernel(const __grid_constant__ Params P) {
  const Term& T = P.t[blockIdx.z];
  const int M = T.A * P.Co, N = T.KE * P.n;
  const int ks = blockIdx.x % P.ksplit;                 // split over the samples; dw is zeroed by the caller when ksplit > 1
  const int m0 = blockIdx.y * BM, n0 = (blockIdx.x / P.ksplit) * BN;
  if (m0 >= M || n0 >= N) return;
  const int b_per = (P.B + P.ksplit - 1) / P.ksplit, b0 = ks * b_per, b1 = min(P.B, b0 + b_per);
  if (b0 >= b1) return;
  float acc[4][4];
  gemm_tile<false, false>(
      1, b0, b1, T.L, m0, n0,
      [&](int r) { return weight_col_ctx(P, T, M, r); },
      [&](const WeightColCtx& c, int, int b, int l) -> float {
        if (!c.valid) return 0.f;
        int row, col;
        return __ldg(P.gz + z_offset(P, T, b, c.a, l, &row, &col) + c.co);
      },
      [&](int c) {
        WindowCtx w;
        w.valid = c < N;
        w.e = w.valid ? c / P.n : 0;
        w.nn = c - w.e * P.n;
        return w;
      },
      [&](const WindowCtx& c, int, int b, int l) -> float { return c.valid ? line_at(P, T, b, l + c.e - T.pad, c.nn) : 0.f; },
      acc);
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
    const int a = r / P.Co, co = r - a * P.Co;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= N) continue;
      const int e = c / P.n, nn = c - e * P.n;
      float* dst = T.dw + (long long)co * T.w_sco + (long long)a * T.w_sa + (long long)e * T.w_se + nn;
      if (P.ksplit > 1) atomicAdd(dst, T.alpha * acc[i][j]);
      else *dst = T.alpha * acc[i][j];
    }
  }
}

// backward, bias classes: gdB[class(row)][class(col)][co] += gz[b][row][col][co] over the frame pixels
__global__ void __launch_bounds__(NT) frame_backward_bias_kernel(const __grid_constant__ Params P) {
  const int m = P.m, h = P.h, w = P.w;
  const int F = h * w - (h - 2 * m) * (w - 2 * m);              // frame pixels per sample
  const long long total = (long long)P.B * F * P.Co;
  const int g = 2 * m + 1;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int co = (int)(i % P.Co);
    const long long bf = i / P.Co;
    const int f = (int)(bf % F), b = (int)(bf / F);
    int row, col;
    if (f < m * w) { row = f / w; col = f - row * w; }
    else if (f < 2 * m * w) { const int q = f - m * w; row = h - m + q / w; col = q % w; }
    else {
      const int q = f - 2 * m * w;                               // (h - 2m) rows x 2m columns
      row = m + q / (2 * m);
      const int cc = q % (2 * m);
      col = cc < m ? cc : w - 2 * m + cc;
    }
    const float v = __ldg(P.gz + (((long long)b * h + row) * w + col) * P.Co + co);
    atomicAdd(P.gdB + ((long long)border_class(row, h, m) * g + border_class(col, w, m)) * P.Co + co, v);
  }
}

static int fill(const pn_frame_desc* d, Params& P) {
  PN_REQUIRE(d, PN_ERR_BAD_ARGUMENT, "pn_pack_frame: null descriptor");
  PN_REQUIRE(d->batch > 0 && d->height > 0 && d->width > 0 && d->cout > 0 && d->n > 0, PN_ERR_BAD_ARGUMENT, "pn_pack_frame: bad shape");
  PN_REQUIRE(d->ksize == 3 || d->ksize == 5, PN_ERR_UNSUPPORTED, "pn_pack_frame: ksize %d (3 or 5)", d->ksize);
  const int m = d->ksize / 2;
  PN_REQUIRE(d->height >= 2 * m + 1 && d->width >= 2 * m + 1, PN_ERR_UNSUPPORTED, "pn_pack_frame: map %dx%d smaller than the frame",
             d->height, d->width);
  PN_REQUIRE(d->num_terms >= 1 && d->num_terms <= 8, PN_ERR_BAD_ARGUMENT, "pn_pack_frame: %d terms (1..8)", d->num_terms);
  P.B = d->batch; P.h = d->height; P.w = d->width; P.Co = d->cout; P.n = d->n; P.m = m; P.nterms = d->num_terms;
  for (int i = 0; i < d->num_terms; ++i) {
    const pn_frame_term& s = d->terms[i];
    Term& T = P.t[i];
    PN_REQUIRE(s.line && s.w && s.L > 0 && s.A > 0 && s.A2 > 0 && s.KE > 0, PN_ERR_BAD_ARGUMENT, "pn_pack_frame: term %d incomplete", i);
    T.line = s.line; T.w = s.w; T.dline = s.dline; T.dw = s.dw;
    T.line_bs = s.line_bstride; T.dline_bs = s.dline_bstride;
    T.w_sco = s.w_sco; T.w_sa = s.w_sa; T.w_se = s.w_se;
    T.L = s.L; T.A = s.A; T.A2 = s.A2; T.KE = s.KE; T.pad = s.pad;
    T.r0 = s.r0; T.ra1 = s.ra1; T.ra2 = s.ra2; T.rl = s.rl;
    T.c0 = s.c0; T.ca1 = s.ca1; T.ca2 = s.ca2; T.cl = s.cl;
    T.alpha = s.alpha; T.bias_mode = s.bias_mode;
    // the affine map must stay inside the map for every (a, l): check the four extreme combinations
    for (int ea = 0; ea < 2; ++ea)
      for (int el = 0; el < 2; ++el) {
        const int a = ea ? s.A - 1 : 0, l = el ? s.L - 1 : 0;
        const int a1 = a / s.A2, a2 = a % s.A2;
        const int r = s.r0 + a1 * s.ra1 + a2 * s.ra2 + l * s.rl, c = s.c0 + a1 * s.ca1 + a2 * s.ca2 + l * s.cl;
        PN_REQUIRE(r >= 0 && r < d->height && c >= 0 && c < d->width, PN_ERR_BAD_ARGUMENT,
                   "pn_pack_frame: term %d maps (a=%d, l=%d) to (%d, %d) outside the %dx%d map", i, a, l, r, c, d->height, d->width);
      }
  }
  return 0;
}

static unsigned cdiv(long long a, int b) { return (unsigned)((a + b - 1) / b); }

}  // namespace frame
}  // namespace pn

using namespace pn;

extern "C" int pn_pack_frame_forward(const pn_frame_desc* desc, const float* dB, float* z, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  frame::Params P{};
  if (int rc = frame::fill(desc, P)) return rc;
  PN_REQUIRE(dB && z, PN_ERR_BAD_ARGUMENT, "pn_pack_frame_forward: null argument");
  TraceScope ts(stream, "frame_fwd B%d H%d W%d Co%d n%d k%d", P.B, P.h, P.w, P.Co, P.n, desc->ksize);
  P.dB = dB; P.z = z;
  long long mmax = 0, nmax = 0;
  for (int i = 0; i < P.nterms; ++i) {
    mmax = std::max<long long>(mmax, (long long)P.B * P.t[i].L);
    nmax = std::max<long long>(nmax, (long long)P.t[i].A * P.Co);
  }
  int kemax = 1;
  for (int i = 0; i < P.nterms; ++i) kemax = std::max(kemax, P.t[i].KE);
  P.ksplit = (desc->flags & PN_FRAME_FLAG_NO_KSPLIT) ? 1 : kemax;
  PN_LAUNCH(frame::frame_forward_kernel, dim3(frame::cdiv(nmax, frame::BN) * P.ksplit, frame::cdiv(mmax, frame::BM), P.nterms), frame::NT, 0, stream, P);
  count_launch();
  return check_launch("frame_forward_kernel");
}

extern "C" int pn_pack_frame_backward(const pn_frame_desc* desc, const float* gz, float* gdB, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  frame::Params P{};
  if (int rc = frame::fill(desc, P)) return rc;
  PN_REQUIRE(gz && gdB, PN_ERR_BAD_ARGUMENT, "pn_pack_frame_backward: null argument");
  for (int i = 0; i < P.nterms; ++i)
    PN_REQUIRE(P.t[i].dline && P.t[i].dw, PN_ERR_BAD_ARGUMENT, "pn_pack_frame_backward: term %d has no gradient buffers", i);
  TraceScope ts(stream, "frame_bwd B%d H%d W%d Co%d n%d k%d", P.B, P.h, P.w, P.Co, P.n, desc->ksize);
  P.gz = gz; P.gdB = gdB;
  long long pmax = 0, amax = 0, kmax = 0;
  for (int i = 0; i < P.nterms; ++i) {
    pmax = std::max<long long>(pmax, (long long)P.B * P.t[i].L);
    amax = std::max<long long>(amax, (long long)P.t[i].A * P.Co);
    kmax = std::max<long long>(kmax, (long long)P.t[i].KE * P.n);
  }
  int kemax = 1;
  for (int i = 0; i < P.nterms; ++i) kemax = std::max(kemax, P.t[i].KE);
  const bool split = !(desc->flags & PN_FRAME_FLAG_NO_KSPLIT);
  P.ksplit = split ? kemax : 1;
  PN_LAUNCH(frame::frame_backward_line_kernel, dim3(frame::cdiv(P.n, frame::BN) * P.ksplit, frame::cdiv(pmax, frame::BM), P.nterms), frame::NT, 0, stream, P);
  // weights: split over the samples only when the caller zeroed dw (PN_FRAME_FLAG_DW_ZEROED): partial sums meet by atomicAdd
  P.ksplit = (split && (desc->flags & PN_FRAME_FLAG_DW_ZEROED)) ? P.B : 1;
  PN_LAUNCH(frame::frame_backward_weight_kernel, dim3(frame::cdiv(kmax, frame::BN) * P.ksplit, frame::cdiv(amax, frame::BM), P.nterms), frame::NT, 0, stream, P);
  const int m = P.m;
  const long long total = (long long)P.B * ((long long)P.h * P.w - (long long)(P.h - 2 * m) * (P.w - 2 * m)) * P.Co;
  unsigned blocks = frame::cdiv(total, frame::NT);
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  PN_LAUNCH(frame::frame_backward_bias_kernel, blocks, frame::NT, 0, stream, P);
  count_launch(3);
  return check_launch("frame_backward kernels");
}
