// fold_kernels.cu -- weight folding of the pack block (PackLayerConv3d, layers01.py:213-247), sm_100a.
//
// The Conv3d(1->8, 3x3x3, pad 1) and the Conv2d(8n -> Co, k x k) of a pack layer have nothing but zero padding between
// them, so their composition is one Conv2d(n -> Co, (k+2) x (k+2)) of the space-to-depth tensor whose weight is the FULL
// convolution of the two kernels (packnet_sfm_b200/folded.py derives it and the exact frame terms):
//
//     out[co][c''][ea][eb] = sum_{f, dc, dy, dx}  W2[co][f][c''-dc+1][ky0 + ea-(dy-dy0)][kx0 + eb-(dx-dx0)] * W3[f][dc][dy][dx]
//
// with the tap window [ky0,ky1) x [kx0,kx1) of W2 and the face [dy0,dy1) x [dx0,dx1) of W3 selected by the caller: the
// whole kernels give the interior weight W_eff; a border row / column of taps against the matching face of W3 gives the
// 1 x (k+2) / (k+2) x 1 / 1 x 1 weights of the frame terms.  The depth index c'' is clipped to [0, n) -- the Conv3d pads
// its depth with zeros.  Output channels are stored in the (i, j, c) order of our space-to-depth tensor
// (c'' = 4c + 2i + j  ->  (2i+j)*C + c), OIHW, ready for pn_conv2d_pack_weight.
//
// HBM-bound fp32 work (W2 of pack5 is 302 MB; 216 FMAs per folded weight, 27 per W2 gradient): one thread per
// (co, c'') keeps the whole (k+2)^2 patch in registers and reads every W2 value it needs exactly once through L1.
#include "common.cuh"

namespace pn {
namespace fold {

struct FoldParams {
  int cout, n, k, C;        // W2 [cout][8][n][k][k]; C = n / 4
  int ky0, kx0;             // first tap of the window (window size is a template parameter)
  int dy0, dx0;             // first row / column of the W3 face
  const float* w2;
  const float* w3;          // [8][3][3][3]
  float* out;               // forward: [cout][n (i,j,c)][EA][EB]
  const float* dout;        // backward: gradient of out
  const float* dS;          // backward, optional: [cout][8][k][k] added to every depth of dW2 (gradient of sum_c' W2)
  float* dw2;               // backward: [cout][8][n][k][k]; the window is overwritten or accumulated
  float* dw3;               // backward: [8][27], atomically accumulated (the caller zeroes it once per fold set)
  int accumulate;
  int ohwi;                 // layout of out / dout: 0 = OIHW [cout][n][EA][EB], 1 = OHWI [cout][EA][EB][n] (channels last)
};

__device__ __forceinline__ int perm_channel(int cpp, int C) { return (cpp & 3) * C + (cpp >> 2); }

constexpr int FOLD_THREADS = 128;

// grid (ceil(n / 128), cout); thread -> depth c'' of the folded weight
template <int KA, int KB, int DA, int DB>
__global__ void __launch_bounds__(FOLD_THREADS) fold_fwd_kernel(FoldParams P) {
  constexpr int EA = KA + DA - 1, EB = KB + DB - 1, E = EA * EB;
  __shared__ float w3s[8 * 3 * DA * DB];
  __shared__ float stage[FOLD_THREADS * E];
  const int tid = threadIdx.x;
  for (int i = tid; i < 8 * 3 * DA * DB; i += FOLD_THREADS) {
    const int dx = i % DB, dy = (i / DB) % DA, fdc = i / (DA * DB);
    w3s[i] = P.w3[fdc * 9 + (P.dy0 + dy) * 3 + (P.dx0 + dx)];
  }
  __syncthreads();
  const int co = blockIdx.y;
  const int c0 = blockIdx.x * FOLD_THREADS;
  const int cpp = c0 + tid;
  float acc[EA][EB];
#pragma unroll
  for (int a = 0; a < EA; ++a)
#pragma unroll
    for (int b = 0; b < EB; ++b) acc[a][b] = 0.f;
  if (cpp < P.n) {
    for (int f = 0; f < 8; ++f) {
#pragma unroll
      for (int dc = 0; dc < 3; ++dc) {
        const int cp = cpp - dc + 1;
        if (cp < 0 || cp >= P.n) continue;
        const float* src = P.w2 + ((((size_t)co * 8 + f) * P.n + cp) * P.k + P.ky0) * P.k + P.kx0;
        float v[KA][KB];
#pragma unroll
        for (int a = 0; a < KA; ++a)
#pragma unroll
          for (int b = 0; b < KB; ++b) v[a][b] = __ldg(src + a * P.k + b);
        const float* wf = w3s + (f * 3 + dc) * DA * DB;
#pragma unroll
        for (int dy = 0; dy < DA; ++dy)
#pragma unroll
          for (int dx = 0; dx < DB; ++dx) {
            const float w = wf[dy * DB + dx];
#pragma unroll
            for (int a = 0; a < KA; ++a)
#pragma unroll
              for (int b = 0; b < KB; ++b) acc[a + dy][b + dx] = fmaf(v[a][b], w, acc[a + dy][b + dx]);
          }
      }
    }
  }
  // stage so that the four (i,j) runs of this block leave as contiguous stores: run q holds the block's channels c
  const int q = tid & 3, cl = tid >> 2;
#pragma unroll
  for (int a = 0; a < EA; ++a)
#pragma unroll
    for (int b = 0; b < EB; ++b) stage[(q * 32 + cl) * E + a * EB + b] = acc[a][b];
  __syncthreads();
  const int cbase = c0 >> 2;                                   // first channel c of the block
  int nc = P.C - cbase;
  if (nc > 32) nc = 32;
  if (!P.ohwi) {
    for (int idx = tid; idx < 4 * 32 * E; idx += FOLD_THREADS) {
      const int qq = idx / (32 * E), r = idx - qq * 32 * E;
      if (r < nc * E) P.out[((size_t)co * P.n + (size_t)qq * P.C + cbase) * E + r] = stage[idx];
    }
  } else {
    // channels last: for every patch position the block owns four runs (one per (i,j)) of up to 32 consecutive channels
    for (int idx = tid; idx < 4 * 32 * E; idx += FOLD_THREADS) {
      const int c = idx & 31, qq = (idx >> 5) & 3, e = idx >> 7;
      if (c < nc) P.out[((size_t)co * E + e) * P.n + (size_t)qq * P.C + cbase + c] = stage[(qq * 32 + c) * E + e];
    }
  }
}

// grid (ceil(n / 128), 8, cout); thread -> depth c' of W2[co][f]
template <int KA, int KB, int DA, int DB>
__global__ void __launch_bounds__(FOLD_THREADS) fold_bwd_kernel(FoldParams P) {
  constexpr int EA = KA + DA - 1, EB = KB + DB - 1, E = EA * EB;
  __shared__ float w3s[3 * DA * DB];
  __shared__ float red[3 * DA * DB];
  const int tid = threadIdx.x;
  const int f = blockIdx.y, co = blockIdx.z;
  if (tid < 3 * DA * DB) {
    const int dx = tid % DB, dy = (tid / DB) % DA, dc = tid / (DA * DB);
    w3s[tid] = P.w3[(f * 3 + dc) * 9 + (P.dy0 + dy) * 3 + (P.dx0 + dx)];
    red[tid] = 0.f;
  }
  __syncthreads();
  const int cp = blockIdx.x * FOLD_THREADS + tid;
  float g2[KA][KB], g3[3][DA][DB], v[KA][KB];
#pragma unroll
  for (int a = 0; a < KA; ++a)
#pragma unroll
    for (int b = 0; b < KB; ++b) { g2[a][b] = 0.f; v[a][b] = 0.f; }
#pragma unroll
  for (int dc = 0; dc < 3; ++dc)
#pragma unroll
    for (int dy = 0; dy < DA; ++dy)
#pragma unroll
      for (int dx = 0; dx < DB; ++dx) g3[dc][dy][dx] = 0.f;
  const bool live = cp < P.n;
  float* dst = nullptr;
  if (live) {
    const size_t off = ((((size_t)co * 8 + f) * P.n + cp) * P.k + P.ky0) * P.k + P.kx0;
    dst = P.dw2 + off;
    const float* src = P.w2 + off;
#pragma unroll
    for (int a = 0; a < KA; ++a)
#pragma unroll
      for (int b = 0; b < KB; ++b) v[a][b] = __ldg(src + a * P.k + b);
#pragma unroll
    for (int dc = 0; dc < 3; ++dc) {
      const int cpp = cp + dc - 1;
      if (cpp < 0 || cpp >= P.n) continue;
      const int pc = perm_channel(cpp, P.C);
      const float* d = P.ohwi ? P.dout + (size_t)co * E * P.n + pc : P.dout + ((size_t)co * P.n + pc) * E;
      const size_t es = P.ohwi ? (size_t)P.n : 1;                // stride between patch positions
#pragma unroll
      for (int ea = 0; ea < EA; ++ea)
#pragma unroll
        for (int eb = 0; eb < EB; ++eb) {
          const float dv = __ldg(d + (size_t)(ea * EB + eb) * es);
#pragma unroll
          for (int dy = 0; dy < DA; ++dy)
#pragma unroll
            for (int dx = 0; dx < DB; ++dx) {
              const int a = ea - dy, b = eb - dx;            // compile-time after unrolling
              if (a >= 0 && a < KA && b >= 0 && b < KB) {
                g2[a][b] = fmaf(dv, w3s[(dc * DA + dy) * DB + dx], g2[a][b]);
                g3[dc][dy][dx] = fmaf(dv, v[a][b], g3[dc][dy][dx]);
              }
            }
        }
    }
    const float* ds = P.dS ? P.dS + (((size_t)co * 8 + f) * P.k + P.ky0) * P.k + P.kx0 : nullptr;
#pragma unroll
    for (int a = 0; a < KA; ++a)
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        float r = g2[a][b];
        if (ds) r += __ldg(ds + a * P.k + b);
        if (P.accumulate) r += dst[a * P.k + b];
        dst[a * P.k + b] = r;
      }
  }
  // dW3[f][dc][dy][dx]: warp tree, one shared atomic per warp, one global atomic per block
#pragma unroll
  for (int dc = 0; dc < 3; ++dc)
#pragma unroll
    for (int dy = 0; dy < DA; ++dy)
#pragma unroll
      for (int dx = 0; dx < DB; ++dx) {
        float s = g3[dc][dy][dx];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) atomicAdd(&red[(dc * DA + dy) * DB + dx], s);
      }
  __syncthreads();
  if (tid < 3 * DA * DB) {
    const int dx = tid % DB, dy = (tid / DB) % DA, dc = tid / (DA * DB);
    atomicAdd(P.dw3 + (f * 3 + dc) * 9 + (P.dy0 + dy) * 3 + (P.dx0 + dx), red[tid]);
  }
}

template <int KA, int KB, int DA, int DB>
static int launch_fold(const FoldParams& P, bool backward, cudaStream_t stream) {
  const int nblk = (P.n + FOLD_THREADS - 1) / FOLD_THREADS;
  if (backward) {
    PN_LAUNCH((fold_bwd_kernel<KA, KB, DA, DB>), dim3(nblk, 8, P.cout), FOLD_THREADS, 0, stream, P);
    count_launch();
    return check_launch("fold_bwd_kernel");
  }
  PN_LAUNCH((fold_fwd_kernel<KA, KB, DA, DB>), dim3(nblk, P.cout), FOLD_THREADS, 0, stream, P);
  count_launch();
  return check_launch("fold_fwd_kernel");
}

static int dispatch(const pn_fold_desc* d, FoldParams& P, bool backward, cudaStream_t stream) {
  PN_REQUIRE(d, PN_ERR_BAD_ARGUMENT, "pn_pack_fold: null descriptor");
  PN_REQUIRE(d->cout > 0 && d->cout <= 65535 && d->n > 0 && d->n % 4 == 0, PN_ERR_BAD_ARGUMENT,
             "pn_pack_fold: cout %d (1..65535), n %d (positive multiple of 4)", d->cout, d->n);
  PN_REQUIRE(d->ksize == 3 || d->ksize == 5, PN_ERR_UNSUPPORTED, "pn_pack_fold: ksize %d (3 or 5)", d->ksize);
  const int k = d->ksize, m = k / 2;
  const int ka = d->ky1 - d->ky0, kb = d->kx1 - d->kx0, da = d->dy1 - d->dy0, db = d->dx1 - d->dx0;
  PN_REQUIRE(d->ky0 >= 0 && d->kx0 >= 0 && d->ky1 <= k && d->kx1 <= k && d->dy0 >= 0 && d->dx0 >= 0 && d->dy1 <= 3 && d->dx1 <= 3,
             PN_ERR_BAD_ARGUMENT, "pn_pack_fold: tap window / face out of range");
  PN_REQUIRE((ka == k || ka == m) && (kb == k || kb == m) && (da == 3 || da == 1) && (db == 3 || db == 1) &&
                 ((ka == k) == (da == 3)) && ((kb == k) == (db == 3)),
             PN_ERR_UNSUPPORTED, "pn_pack_fold: window %dx%d with face %dx%d (whole kernel with whole face, or %d border taps with one face)",
             ka, kb, da, db, m);
  P.cout = d->cout; P.n = d->n; P.k = k; P.C = d->n / 4;
  P.ky0 = d->ky0; P.kx0 = d->kx0; P.dy0 = d->dy0; P.dx0 = d->dx0;
  P.ohwi = d->layout == PN_FOLD_LAYOUT_OHWI ? 1 : 0;
  PN_REQUIRE(d->layout == PN_FOLD_LAYOUT_OIHW || d->layout == PN_FOLD_LAYOUT_OHWI, PN_ERR_BAD_ARGUMENT, "pn_pack_fold: layout %d", d->layout);
  const bool fa = ka == k, fb = kb == k;
  if (k == 3) {
    if (fa && fb) return launch_fold<3, 3, 3, 3>(P, backward, stream);
    if (!fa && fb) return launch_fold<1, 3, 1, 3>(P, backward, stream);
    if (fa && !fb) return launch_fold<3, 1, 3, 1>(P, backward, stream);
    return launch_fold<1, 1, 1, 1>(P, backward, stream);
  }
  if (fa && fb) return launch_fold<5, 5, 3, 3>(P, backward, stream);
  if (!fa && fb) return launch_fold<2, 5, 1, 3>(P, backward, stream);
  if (fa && !fb) return launch_fold<5, 2, 3, 1>(P, backward, stream);
  return launch_fold<2, 2, 1, 1>(P, backward, stream);
}

}  // namespace fold
}  // namespace pn

using namespace pn;

extern "C" int pn_pack_fold_forward(const pn_fold_desc* desc, const float* w2, const float* w3, float* out, pn_stream_t stream) {
  PN_REQUIRE(w2 && w3 && out, PN_ERR_BAD_ARGUMENT, "pn_pack_fold_forward: null argument");
  TraceScope ts(reinterpret_cast<cudaStream_t>(stream), "fold_fwd Co%d n%d k%d win%dx%d", desc ? desc->cout : 0, desc ? desc->n : 0,
                desc ? desc->ksize : 0, desc ? desc->ky1 - desc->ky0 : 0, desc ? desc->kx1 - desc->kx0 : 0);
  fold::FoldParams P{};
  P.w2 = w2; P.w3 = w3; P.out = out;
  return fold::dispatch(desc, P, false, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int pn_pack_fold_backward(const pn_fold_desc* desc, const float* w2, const float* w3, const float* dout, const float* dS,
                                     float* dw2, float* dw3, int accumulate, pn_stream_t stream) {
  PN_REQUIRE(w2 && w3 && dout && dw2 && dw3, PN_ERR_BAD_ARGUMENT, "pn_pack_fold_backward: null argument");
  TraceScope ts(reinterpret_cast<cudaStream_t>(stream), "fold_bwd Co%d n%d k%d win%dx%d", desc ? desc->cout : 0, desc ? desc->n : 0,
                desc ? desc->ksize : 0, desc ? desc->ky1 - desc->ky0 : 0, desc ? desc->kx1 - desc->kx0 : 0);
  fold::FoldParams P{};
  P.w2 = w2; P.w3 = w3; P.dout = dout; P.dS = dS; P.dw2 = dw2; P.dw3 = dw3; P.accumulate = accumulate;
  return fold::dispatch(desc, P, true, reinterpret_cast<cudaStream_t>(stream));
}

// The nine folds of one pack layer in ONE call (interior, four sides, four corners -- the window table of
// packnet_sfm_b200/folded.py::fold_windows): saves the Python -> C round trips of the per-window entry points.
// outs / douts [9]: order main, top, bottom, left, right, tl, tr, bl, br; main OIHW, the others OHWI.
namespace pn {
namespace fold {
static void window_of(int which, int k, pn_fold_desc* d) {
  const int m = k / 2;
  struct R { int a0, a1; };
  const R lo{0, m}, hi{m + 1, k}, al{0, k}, f3{0, 3}, last{2, 3}, first{0, 1};
  const R ky[9] = {al, lo, hi, al, al, lo, lo, hi, hi};
  const R kx[9] = {al, al, al, lo, hi, lo, hi, lo, hi};
  const R dy[9] = {f3, last, first, f3, f3, last, last, first, first};
  const R dx[9] = {f3, f3, f3, last, first, last, first, last, first};
  d->ky0 = ky[which].a0; d->ky1 = ky[which].a1; d->kx0 = kx[which].a0; d->kx1 = kx[which].a1;
  d->dy0 = dy[which].a0; d->dy1 = dy[which].a1; d->dx0 = dx[which].a0; d->dx1 = dx[which].a1;
  d->layout = which == 0 ? PN_FOLD_LAYOUT_OIHW : PN_FOLD_LAYOUT_OHWI;
}
}  // namespace fold
}  // namespace pn

extern "C" int pn_pack_fold_set_forward(int cout, int n, int ksize, const float* w2, const float* w3, float* const* outs,
                                        pn_stream_t stream) {
  PN_REQUIRE(w2 && w3 && outs, PN_ERR_BAD_ARGUMENT, "pn_pack_fold_set_forward: null argument");
  for (int i = 0; i < 9; ++i) {
    pn_fold_desc d{};
    d.cout = cout; d.n = n; d.ksize = ksize;
    fold::window_of(i, ksize, &d);
    if (int rc = pn_pack_fold_forward(&d, w2, w3, outs[i], stream)) return rc;
  }
  return PN_OK;
}

extern "C" int pn_pack_fold_set_backward(int cout, int n, int ksize, const float* w2, const float* w3, const float* const* douts,
                                         const float* dS, float* dw2, float* dw3, pn_stream_t stream) {
  PN_REQUIRE(w2 && w3 && douts && douts[0] && dw2 && dw3, PN_ERR_BAD_ARGUMENT,
             "pn_pack_fold_set_backward: null argument (the interior gradient douts[0] is mandatory: it overwrites dw2)");
  for (int i = 0; i < 9; ++i) {
    if (!douts[i]) continue;
    pn_fold_desc d{};
    d.cout = cout; d.n = n; d.ksize = ksize;
    fold::window_of(i, ksize, &d);
    if (int rc = pn_pack_fold_backward(&d, w2, w3, douts[i], i == 0 ? dS : nullptr, dw2, dw3, i == 0 ? 0 : 1, stream)) return rc;
  }
  return PN_OK;
}
