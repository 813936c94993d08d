// gn_cluster_kernel.cuh -- GroupNorm(16)+ELU forward / backward in ONE launch per direction on thread-block clusters
// (included by layer_kernels.cu; device build only -- the CPU emulation tier keeps the two-pass kernels).
//
// Why: r02n ncu launch list of a step -- 35 of the 47 GroupNorm layers of PackNet01 work on maps of 48x160 pixels or less,
// where the statistics pass + apply pass (+ reduce pass + apply pass backward) are not bandwidth but LATENCY: ~12 us per
// launch for 4-16 MB (0.3-1.2 TB/s), two memsets per layer, fp64 atomics of every CTA on 128 addresses.  1.7 ms of the
// 3.7 ms the GroupNorm kernels take per step went to tensors that fit the L2 several times over.
//
// How: a cluster of CL CTAs owns (sample b, block of `cw` channels = whole groups); CTA r of the cluster takes the r-th
// pixel range.  Pass 1 reads the CTA's elements once from HBM (per-thread fp32 partial sums, fp64 from the first shuffle
// on), the CL partial results meet through DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank) between two cluster
// barriers, pass 2 re-reads the same elements -- now L2 hits -- and writes the result (and the bf16 hi / lo operand pair of
// the convolution that consumes it).  No global atomics on the statistics, no memset, no finalize launch.
// Used for tensors up to 8.5 MB (host: gn_cluster_plan -- 24x80x256 and smaller at B=4 192x640: 30 of the 47 layers).  Measured
// (tools/gn_bench.py, replayed graphs, us forward / backward, two-pass -> cluster): 12x40x512 24.6 / 34.8 -> 14.3 / 18.4,
// 6x20x512 24.6 / 34.8 -> 12.3 / 15.1, 24x80x256 20.4 / 30.8 -> 18.5 / 24.6; above that the <= 128 CTAs of a cluster grid
// do not keep enough bytes in flight (96x320x64: 41 / 63 -> 59 / 98) and the two-pass kernels stay (4.4-5.9 TB/s on
// 192x640).  Whole step: 20.75 -> 20.35 ms with the 8.5 MB bound (20.84 with every L2-sized tensor on clusters).
//   forward  = layers01.py:31-32,37 (Conv2D) / :61-62,72 (ResidualConv: x + shortcut)
#pragma once
#include <cooperative_groups.h>

namespace pn {
namespace layers {

namespace cgx = cooperative_groups;

constexpr int GNC_THREADS = 512;
constexpr int GNC_WARPS = GNC_THREADS / 32;
constexpr int GNC_MAXCW = 128;     // channels per CTA (32 for C >= 32; a whole group when groups are wider)

// Per-channel sums over the block of NQ quantities: thread (pixel lane, float4 column) holds v[q][4]; the pixel lanes of a
// warp meet by shuffles (ncol = cw / 4 divides 32, so column == lane % ncol), the warps through shared memory.
// Result: out[q][c], c < cw, valid for every thread after the call.
template <int NQ>
__device__ __forceinline__ void block_channel_sums(const float (&v)[NQ][4], int ncol, int cw, double* s_part /*[NQ][warps][cw]*/,
                                                   double* s_out /*[NQ][GNC_MAXCW]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double d[4] = {(double)v[q][0], (double)v[q][1], (double)v[q][2], (double)v[q][3]};
    for (int o = ncol; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
    }
    if (lane < ncol) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s_part[(q * GNC_WARPS + warp) * cw + lane * 4 + k] = d[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * cw; i += GNC_THREADS) {
    const int q = i / cw, c = i % cw;
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < GNC_WARPS; ++w) a += s_part[(q * GNC_WARPS + w) * cw + c];
    s_out[q * GNC_MAXCW + c] = a;
  }
  __syncthreads();
}

struct GnClusterParams {
  const float* x; const float* x2;     // [B,HW,C] (+ optional second addend)
  int HW, C, cw, ppc;                   // pixels, channels, channels per CTA, pixels per CTA (ceil(HW / cluster size))
  const float* gamma; const float* beta;
  float eps;
  float* mr;                            // [B*16][2] (mean, rstd): written by the forward, read by the backward
  // forward
  float* y; int out_cstride, out_coffset;
  uint2* hi; uint2* lo;                 // optional bf16 pair of the output, contiguous [B,HW,C]
  // backward
  const float* yin; int y_cstride, y_coffset;
  const float* dy; int dy_cstride, dy_coffset;
  float* dx;                            // [B,HW,C]
  float* dgamma; float* dbeta; float* dsum;   // [C] each, ZEROED by the caller (float atomics over samples / clusters)
};

template <int U>      // independent 16-byte loads in flight per thread and trip (4: small tensors; 8: keeps a 96x320 map streaming)
__global__ void __launch_bounds__(GNC_THREADS, 1) gn_elu_cluster_fwd_kernel(const GnClusterParams P) {
  cgx::cluster_group cluster = cgx::this_cluster();
  __shared__ double s_part[2 * GNC_WARPS * GNC_MAXCW];
  __shared__ double s_ch[2 * GNC_MAXCW];
  __shared__ double s_grp[2 * 8];        // this CTA's (sum, sumsq) of its groups (<= 8 groups per CTA): read by the cluster
  __shared__ float s_mr[2 * 8];
  const int C = P.C, cw = P.cw, cg = C / 16, ncol = cw / 4;
  const int ng = cw / cg;                // groups of this CTA (cw is a multiple of the group width)
  const int c0 = blockIdx.y * cw, b = blockIdx.z;
  const int col = threadIdx.x % ncol, pl = threadIdx.x / ncol, lanes = GNC_THREADS / ncol;
  const int p0 = (int)cluster.block_rank() * P.ppc, p1 = min(P.HW, p0 + P.ppc);
  const size_t base = (size_t)b * P.HW * C + c0 + col * 4;

  auto load = [&](int p) {
    float4 v = *reinterpret_cast<const float4*>(P.x + base + (size_t)p * C);
    if (P.x2) {
      const float4 u = *reinterpret_cast<const float4*>(P.x2 + base + (size_t)p * C);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    return v;
  };

  // ---- pass 1: sums ---------------------------------------------------------------------------------
  float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  int p = p0 + pl;
  for (; p + (U - 1) * lanes < p1; p += U * lanes) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = load(p + u * lanes);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc[0][0] += v[u].x; acc[0][1] += v[u].y; acc[0][2] += v[u].z; acc[0][3] += v[u].w;
      acc[1][0] += v[u].x * v[u].x; acc[1][1] += v[u].y * v[u].y; acc[1][2] += v[u].z * v[u].z; acc[1][3] += v[u].w * v[u].w;
    }
  }
  for (; p < p1; p += lanes) {
    const float4 v = load(p);
    acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
    acc[1][0] += v.x * v.x; acc[1][1] += v.y * v.y; acc[1][2] += v.z * v.z; acc[1][3] += v.w * v.w;
  }
  block_channel_sums<2>(acc, ncol, cw, s_part, s_ch);
  if (threadIdx.x < 2 * ng) {
    const int q = threadIdx.x / ng, g = threadIdx.x % ng;
    double a = 0.0;
    for (int k = 0; k < cg; ++k) a += s_ch[q * GNC_MAXCW + g * cg + k];
    s_grp[q * 8 + g] = a;
  }
  cluster.sync();
  if (threadIdx.x < ng) {
    const int g = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) {
      const double* rem = cluster.map_shared_rank(s_grp, r);
      s += rem[g]; q += rem[8 + g];
    }
    const double cnt = (double)P.HW * cg;
    const double m = s / cnt;
    double var = q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)P.eps));
    s_mr[2 * g] = mean; s_mr[2 * g + 1] = rstd;
    if (cluster.block_rank() == 0) {
      const int gg = c0 / cg + g;
      P.mr[((size_t)b * 16 + gg) * 2 + 0] = mean;
      P.mr[((size_t)b * 16 + gg) * 2 + 1] = rstd;
    }
  }
  cluster.sync();      // s_mr visible; no CTA leaves (or overwrites s_grp) while a neighbour still reads its shared memory

  // ---- pass 2: normalise + affine + ELU (re-read: L2 hits) ---------------------------------------------
  float mean[4], rstd[4], gm[4], bt[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cl = col * 4 + k, g = cl / cg;
    mean[k] = s_mr[2 * g]; rstd[k] = s_mr[2 * g + 1];
    gm[k] = __ldg(P.gamma + c0 + cl); bt[k] = __ldg(P.beta + c0 + cl);
  }
  auto apply = [&](int pp, const float4& v) {
    const float in[4] = {v.x, v.y, v.z, v.w};
    float out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float z = (in[k] - mean[k]) * rstd[k] * gm[k] + bt[k];
      out[k] = z > 0.0f ? z : expm1f(z);   // nn.ELU(alpha=1)
    }
    const size_t pix = (size_t)b * P.HW + pp;
    *reinterpret_cast<float4*>(P.y + pix * P.out_cstride + P.out_coffset + c0 + col * 4) = make_float4(out[0], out[1], out[2], out[3]);
    if (P.hi) store_split4(P.hi, P.lo, (pix * C + c0 + col * 4) >> 2, out);
  };
  p = p0 + pl;
  for (; p + (U - 1) * lanes < p1; p += U * lanes) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = load(p + u * lanes);
#pragma unroll
    for (int u = 0; u < U; ++u) apply(p + u * lanes, v[u]);
  }
  for (; p < p1; p += lanes) apply(p, load(p));
}

// backward: dz = dy * ELU'(z) (ELU' = 1 for y > 0 else y + 1); per (sample, channel) S1 = sum dz, S2 = sum dz * xhat;
// per group m1 = mean(gamma * dz), m2 = mean(gamma * dz * xhat); dx = rstd * (gamma * dz - m1 - xhat * m2);
// dgamma = sum_b S2, dbeta = sum_b S1, dsum = sum over samples and pixels of dx (bias gradient of the producing convolution)
__global__ void __launch_bounds__(GNC_THREADS, 1) gn_elu_cluster_bwd_kernel(const GnClusterParams P) {
  cgx::cluster_group cluster = cgx::this_cluster();
  __shared__ double s_part[2 * GNC_WARPS * GNC_MAXCW];
  __shared__ double s_ch[2 * GNC_MAXCW];      // this CTA's per-channel (S1, S2): read by the cluster
  __shared__ double s_tot[2 * GNC_MAXCW];     // cluster totals
  __shared__ float s_gm[2 * 8];
  const int C = P.C, cw = P.cw, cg = C / 16, ncol = cw / 4;
  const int ng = cw / cg;
  const int c0 = blockIdx.y * cw, b = blockIdx.z;
  const int col = threadIdx.x % ncol, pl = threadIdx.x / ncol, lanes = GNC_THREADS / ncol;
  const int p0 = (int)cluster.block_rank() * P.ppc, p1 = min(P.HW, p0 + P.ppc);
  const size_t xbase = (size_t)b * P.HW * C + c0 + col * 4;
  const size_t ybase = (size_t)b * P.HW * P.y_cstride + P.y_coffset + c0 + col * 4;
  const size_t gbase = (size_t)b * P.HW * P.dy_cstride + P.dy_coffset + c0 + col * 4;

  float mean[4], rstd[4], gm[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + col * 4 + k, g = c / cg;
    mean[k] = __ldg(P.mr + ((size_t)b * 16 + g) * 2 + 0);
    rstd[k] = __ldg(P.mr + ((size_t)b * 16 + g) * 2 + 1);
    gm[k] = __ldg(P.gamma + c);
  }
  struct Elem { float xh[4], dz[4]; };
  auto load = [&](int p) {
    float4 xv = *reinterpret_cast<const float4*>(P.x + xbase + (size_t)p * C);
    if (P.x2) {
      const float4 u = *reinterpret_cast<const float4*>(P.x2 + xbase + (size_t)p * C);
      xv.x += u.x; xv.y += u.y; xv.z += u.z; xv.w += u.w;
    }
    const float4 yv = *reinterpret_cast<const float4*>(P.yin + ybase + (size_t)p * P.y_cstride);
    const float4 gv = *reinterpret_cast<const float4*>(P.dy + gbase + (size_t)p * P.dy_cstride);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
    Elem e;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      e.xh[k] = (xs[k] - mean[k]) * rstd[k];
      e.dz[k] = gs[k] * (ys[k] > 0.0f ? 1.0f : ys[k] + 1.0f);
    }
    return e;
  };

  // ---- pass 1 -----------------------------------------------------------------------------------------
  float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  int p = p0 + pl;
  for (; p + lanes < p1; p += 2 * lanes) {
    const Elem e0 = load(p), e1 = load(p + lanes);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[0][k] += e0.dz[k] + e1.dz[k];
      acc[1][k] += e0.dz[k] * e0.xh[k] + e1.dz[k] * e1.xh[k];
    }
  }
  for (; p < p1; p += lanes) {
    const Elem e = load(p);
#pragma unroll
    for (int k = 0; k < 4; ++k) { acc[0][k] += e.dz[k]; acc[1][k] += e.dz[k] * e.xh[k]; }
  }
  block_channel_sums<2>(acc, ncol, cw, s_part, s_ch);
  cluster.sync();
  for (int i = threadIdx.x; i < 2 * cw; i += GNC_THREADS) {
    const int q = i / cw, c = i % cw;
    double a = 0.0;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) a += cluster.map_shared_rank(s_ch, r)[q * GNC_MAXCW + c];
    s_tot[q * GNC_MAXCW + c] = a;
    if (cluster.block_rank() == 0) atomicAdd((q == 0 ? P.dbeta : P.dgamma) + c0 + c, (float)a);
  }
  __syncthreads();
  if (threadIdx.x < ng) {
    const int g = threadIdx.x;
    double m1 = 0.0, m2 = 0.0;
    for (int k = 0; k < cg; ++k) {
      const double gv = (double)__ldg(P.gamma + c0 + g * cg + k);
      m1 += gv * s_tot[g * cg + k];
      m2 += gv * s_tot[GNC_MAXCW + g * cg + k];
    }
    const double cnt = (double)P.HW * cg;
    s_gm[2 * g] = (float)(m1 / cnt);
    s_gm[2 * g + 1] = (float)(m2 / cnt);
  }
  cluster.sync();      // s_gm visible; every remote read of s_ch is done

  // ---- pass 2 -----------------------------------------------------------------------------------------
  float m1[4], m2[4], dacc[1][4] = {{0, 0, 0, 0}};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int g = (col * 4 + k) / cg;
    m1[k] = s_gm[2 * g]; m2[k] = s_gm[2 * g + 1];
  }
  auto apply = [&](int pp, const Elem& e) {
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      r[k] = rstd[k] * (gm[k] * e.dz[k] - m1[k] - e.xh[k] * m2[k]);
      dacc[0][k] += r[k];
    }
    const size_t o = ((size_t)b * P.HW + pp) * C + c0 + col * 4;
    *reinterpret_cast<float4*>(P.dx + o) = make_float4(r[0], r[1], r[2], r[3]);
    if (P.hi) store_split4(P.hi, P.lo, o >> 2, r);
  };
  p = p0 + pl;
  for (; p + lanes < p1; p += 2 * lanes) {
    const Elem e0 = load(p), e1 = load(p + lanes);
    apply(p, e0);
    apply(p + lanes, e1);
  }
  for (; p < p1; p += lanes) apply(p, load(p));
  if (P.dsum) {
    block_channel_sums<1>(dacc, ncol, cw, s_part, s_tot);
    for (int c = threadIdx.x; c < cw; c += GNC_THREADS) atomicAdd(P.dsum + c0 + c, (float)s_tot[c]);
  }
}

// cluster size / channel block for a [B, HW, C] tensor; 0 = use the two-pass kernels
struct GnClusterPlan { int cl, cw, ppc; };
static GnClusterPlan gn_cluster_plan(int B, int HW, int C) {
  GnClusterPlan pl{0, 0, 0};
  const char* e = std::getenv("PN_GN_CLUSTER");
  if (e && e[0] == '0') return pl;
  const int cg = C / 16;
  int cw = C > 32 ? 32 : 16;          // 128 (64) contiguous bytes per pixel and CTA; C = 32 as two blocks: twice the CTAs
  if (cg > cw) cw = cg;
  if (cw > GNC_MAXCW || C % cw || cw % cg || cw / cg > 8 || (32 % (cw / 4))) return pl;
  // the second pass must find the tensor(s) in the L2 (126 MB): up to three input tensors of this size backward
  // (tuning knobs for tools/gn_bench.py: PN_GN_CLUSTER_MAX_MB, PN_GN_CLUSTER_CL = 8 | 16)
  const char* em = std::getenv("PN_GN_CLUSTER_MAX_MB");
  const double max_bytes = (em ? atof(em) : 8.5) * 1e6;
  if ((double)B * HW * C * 4.0 > max_bytes) return pl;
  const char* ec = std::getenv("PN_GN_CLUSTER_CL");
  const int cl_max = (ec && atoi(ec) == 16) ? 16 : 8;       // 16 = non-portable cluster size (opt-in attribute)
  const int nblk = C / cw;
  int cl = 1;
  while (cl < cl_max && (cl * nblk * B < 120 || (HW + cl - 1) / cl > 4096)) cl *= 2;
  while (cl > 1 && (HW + cl - 1) / cl < 16) cl /= 2;     // tiny maps: not less than a few pixels per CTA
  pl.cl = cl; pl.cw = cw; pl.ppc = (HW + cl - 1) / cl;
  return pl;
}

static int gn_cluster_launch(bool bwd, const GnClusterParams& P, int B, int cl, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cl, P.C / P.cw, B);
  cfg.blockDim = dim3(GNC_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const bool deep = P.ppc * (P.cw / 4) >= 8 * GNC_THREADS;     // at least one full trip of 8 loads per thread
  if (cl > 8) {
    if (bwd) PN_CUDA(cudaFuncSetAttribute(gn_elu_cluster_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    else if (deep) PN_CUDA(cudaFuncSetAttribute(gn_elu_cluster_fwd_kernel<8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    else PN_CUDA(cudaFuncSetAttribute(gn_elu_cluster_fwd_kernel<4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  if (bwd) PN_CUDA(cudaLaunchKernelEx(&cfg, gn_elu_cluster_bwd_kernel, P));
  else if (deep) PN_CUDA(cudaLaunchKernelEx(&cfg, gn_elu_cluster_fwd_kernel<8>, P));
  else     PN_CUDA(cudaLaunchKernelEx(&cfg, gn_elu_cluster_fwd_kernel<4>, P));
  count_launch();
  return 0;
}

}  // namespace layers
}  // namespace pn
