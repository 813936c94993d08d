// conv_engine.cu -- stride-1 "same" 2-D convolution as an implicit GEMM on the sm_100a tensor cores.
//
//   D[pixel, co] = sum_{tap, ci} X[pixel + tap, ci] * W[co, tap, ci]       (+ bias)
//
// * activations are NHWC fp32 (== torch.channels_last storage): an M-tile is an 8-wide x 16-tall (x nb
//   images) patch of 128 output pixels, a K-chunk is 32 input channels = one 128-byte row per pixel, so a
//   4-D TMA box IS the canonical K-major SWIZZLE_128B operand tile; the zero padding of the convolution
//   (layers01.py:30 ConstantPad2d) is the TMA out-of-bounds fill -- no im2col, no padded copy.
// * weights are pre-packed [Cout][tap][Cin] (K-major); one 3-D TMA box per (tap, K-chunk).
// * tcgen05.mma kind::tf32, M=128, N=BN, accumulators in TMEM, issued by one thread; a TMA producer
//   thread and four epilogue warps (tcgen05.ld -> +bias -> NHWC stores) run concurrently, linked by
//   mbarriers (tcgen05.commit frees the shared-memory stages).
// * HALO mode: the (16+k-1) x 16 pixel input patch of a tile is loaded ONCE per K-chunk and every tap
//   reads it through a shifted shared-memory descriptor (start address + base_offset), cutting the
//   activation traffic by ~k^2; PER-TAP mode reloads a shifted 8x16 box per tap (always 1024-B aligned).
// * precision: tf32x1 (one MMA per product; what cuDNN does for the reference on Ampere+ GPUs) or
//   tf32x3 (error-compensated: A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with lo = x - trunc_tf32(x); fp32-grade
//   results, needed for the 1e-3 depth parity bar -- DESIGN.md "precision").
//
// Reference call sites replaced: nn.Conv2d inside Conv2D / ResidualConv / PackLayerConv3d /
// UnpackLayerConv3d (packnet_sfm/networks/layers/packnet/layers01.py:28-29,36,58-72,234,272) and their
// autograd backward (dgrad = the same kernel on flipped/transposed weights).
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "common.cuh"

namespace pn {
namespace conv {

constexpr int TILE_W = 8;        // pixels per tile row (one 8-row swizzle atom)
constexpr int TILE_ROWS = 16;    // tile rows x images
constexpr int PATCH_PITCH = 16;  // pixels per row of the HALO patch (2048 B: keeps 8-row groups 1024-B aligned)
constexpr int NTHREADS_IGEMM = 224;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue, warp 6: second MMA issuer
constexpr int MAX_STAGES = 8;

struct KernelParams {
  int B, H, W, Cin, Cout, ks, pad;
  int th, nb;              // tile rows per image, images per tile (th * nb == 16)
  int tiles_x, tiles_y;    // tiles per image group
  int bn;                  // N tile (output channels per CTA)
  int nsplit;              // 1: single product, 3: error-compensated hi/lo split
  int kc;                  // K elements per 128-byte chunk: 32 (tf32 operands) or 64 (bf16 operands)
  int bf16;                // 1: operands are bf16 (kind::f16), 0: fp32 read as tf32 (kind::tf32)
  int halo;                // 1: patch reuse across taps
  int cchunks;             // ceil(Cin / 32)
  int ksplits;             // split of the channel-chunk loop over blockIdx.z (small maps: fill the 148 SMs)
  int stages;              // pipeline slots
  int nwork, nblocks;      // work items (pixel tile x N block) walked by the grid, N blocks per pixel tile
  int nsets, set_cols;     // accumulator sets in TMEM (2: epilogue of tile i overlaps the MMAs of tile i+1) and their column stride
  int group;               // (chunk, tap) items per slot
  int pitch;               // HALO patch: pixels per patch row (8 + k - 1)
  uint32_t patch_tx;       // bytes one patch load delivers (patch_bytes is rounded up to 1024)
  int force_base_offset0;  // debug knob (bring-up): 1 = put (addr>>7)&7 into base_offset (known to be WRONG)
  uint32_t a_stage_bytes, b_stage_bytes, patch_bytes;
  uint32_t tmem_cols;
  uint32_t idesc;          // N = bn
  uint32_t idesc2;         // N = 2*bn (tf32x3: A_hi x [B_hi ; B_lo])
  const uint8_t* wp;       // packed, pre-swizzled weight tiles [cchunk][tap][rows_pad][128 B] (see pack_weight kernels)
  const uint8_t* wp_lo;
  int rows_pad;            // rows of one (cchunk, tap) tile group: Cout rounded up to a multiple of bn
  int bt_nchunks;          // BT (data gradient from the forward packing): 64-channel chunks of the layer's Cin in the packing
  const float* bias;
  float* out;
  unsigned int* error_flag;
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must not hang the GPU box -- flag the error and trap instead
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, unsigned int* error_flag, int who) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      if (error_flag) atomicExch(error_flag, 0xDEAD0000u | (unsigned)who);
      __threadfence_system();
      __trap();
    }
  }
}
// Waits of the single-warp roles (TMA producer, MMA issuer).  Every lane probes, but the loop condition is a warp VOTE,
// i.e. provably warp-uniform: the compiler keeps the role's loop state (stage index, phases, descriptor words) in
// uniform registers across the wait instead of spilling it to vector registers + R2UR per use.  The fast path is one
// probe; the watchdog clock reads stay out of the steady state.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* error_flag, int who) {
  if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
  const long long t0 = clock64();
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      if (error_flag) atomicExch(error_flag, 0xDEAD0000u | (unsigned)who);
      __threadfence_system();
      __trap();
    }
  }
}
// long wait of a whole warp (the epilogue waiting for the accumulator): ONE lane polls, with a sleep between probes, so
// 128 spinning threads do not hammer the mbarrier unit the TMA producer and the MMA issuer depend on (ncu r01: 74 % of
// all stall samples of the convolution kernel sat in this loop and every pipeline stage took ~1500 cycles)
__device__ __forceinline__ void mbar_wait_warp_backoff(uint32_t bar, uint32_t parity, unsigned int* error_flag, int who) {
  if ((threadIdx.x & 31) == 0) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      __nanosleep(256);
      if (clock64() - t0 > 4000000000LL) {
        if (error_flag) atomicExch(error_flag, 0xDEAD0000u | (unsigned)who);
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1-D bulk copy global -> shared (UBLKCP): one request for a whole contiguous tile
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma(bool bf16, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (bf16) umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}
// one lane of a fully converged warp; the region it guards is what the compiler keeps on the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <bool BF16>
__device__ __forceinline__ void umma_k(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (BF16) umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major; 1) | [32,46) SBO>>4 | [46,48) version=1
//   [49,52) base_offset | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
// BT = data gradient read from the layer's FORWARD weight packing: the reduction runs over the layer's output channels,
// which are the ROWS of a forward tile [co][64 ci x 2 B, SWIZZLE_128B]; read as an MN-major B operand (K = rows, 64 N
// elements contiguous -- the layout conv_wgrad_kernel reads its dZ tiles in) the tile IS the transposed weight, so no
// second packing exists.  A K chunk = 64 rows = one contiguous 8 KB run per 64-wide N block; the taps are walked flipped.
template <bool BF16, bool SPLIT, bool HALO, bool BT = false>
__global__ void __launch_bounds__(NTHREADS_IGEMM, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                  const KernelParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-B alignment
  constexpr int nops = SPLIT ? 2 : 1;
  // An ITEM is one (channel chunk, tap): its weight tile(s) [B_hi][B_lo] and, in per-tap mode, its activation tile(s)
  // [A_hi][A_lo] in front of them.  A pipeline SLOT holds a GROUP of P.group consecutive items behind ONE full/empty
  // barrier pair: the wait + fence + commit hand-off costs ~100 issue cycles the tensor pipe cannot hide
  // (tools/ubench_tc2.cu), so it is paid once per group instead of once per tap.
  // carve: [HALO patches: 2 buffers x nops] [slots] [barriers]
  const uint32_t patch_region = HALO ? 2u * nops * P.patch_bytes : 0u;
  const uint32_t item_bytes = (HALO ? 0u : nops * P.a_stage_bytes) + nops * P.b_stage_bytes;
  const uint32_t slot_bytes = (uint32_t)P.group * item_bytes;
  const uint32_t slots_base = smem_base + patch_region;
  const uint32_t bars_base = slots_base + P.stages * slot_bytes;
  // barriers: full[2][stages], empty[stages], full_a[2], empty_a[2], tmem_full[2], tmem_empty[2], first[2]; TMEM address.
  // A slot has TWO full barriers used on alternate rounds (round = use count of the slot).  Each MMA issuer waits only
  // on its own groups, and with an odd slot count it meets a slot every OTHER round: on a single barrier it would skip a
  // phase, and mbarrier waits are by parity -- the phase two steps back reads as "complete" when copies finish out of
  // order (weights streaming from HBM; measured as a pipeline deadlock on the 16384-channel layers).  With two barriers
  // every waiter observes every phase of the barrier it waits on.
  auto full_bar = [&](int s, int round) { return bars_base + 8u * ((round & 1) * MAX_STAGES + s); };
  auto empty_bar = [&](int s) { return bars_base + 8u * (2 * MAX_STAGES + s); };
  auto fulla_bar = [&](int s) { return bars_base + 8u * (3 * MAX_STAGES + s); };
  auto emptya_bar = [&](int s) { return bars_base + 8u * (3 * MAX_STAGES + 2 + s); };
  auto tmemfull_bar = [&](int a) { return bars_base + 8u * (3 * MAX_STAGES + 4 + a); };
  auto tmemempty_bar = [&](int a) { return bars_base + 8u * (3 * MAX_STAGES + 6 + a); };
  auto first_bar = [&](int a) { return bars_base + 8u * (3 * MAX_STAGES + 8 + a); };
  const uint32_t tmem_slot = bars_base + 8u * (3 * MAX_STAGES + 10);
  auto patch_addr = [&](int buf, int op) { return smem_base + (uint32_t)(buf * nops + op) * P.patch_bytes; };
  const uint32_t b_off = HALO ? 0u : nops * P.a_stage_bytes;   // weight tiles inside an item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // WORK ITEMS are (pixel tile, N block) pairs; a CTA walks work = blockIdx.x, blockIdx.x + gridDim.x, ...  With one
  // work item per CTA this is the plain tiled kernel (used with the K split of small maps, blockIdx.z); with ~148 CTAs
  // it is PERSISTENT: barrier setup and the TMEM allocation are paid once, the producer prefetches the next tile while
  // the epilogue drains the current one, and with two accumulator sets in TMEM (P.nsets) the next tile's MMAs overlap
  // the epilogue as well.  Short-K layers (64 -> 64, 3x3) spent ~75 % of a tile's life outside the MMA loop.
  const int taps = P.ks * P.ks;
  const int cc_per = (P.cchunks + P.ksplits - 1) / P.ksplits;
  const int cc_begin = blockIdx.z * cc_per, cc_end = min(P.cchunks, cc_begin + cc_per);
  const bool has_work = cc_end > cc_begin;
  const int ncc = cc_end - cc_begin;
  // HALO: groups never straddle a channel chunk (they share its patch); per-tap: groups run over the flattened
  // (chunk, tap) item list.
  const int G = P.group;
  const int nitems = HALO ? taps : ncc * taps;
  const int ngroups = (nitems + G - 1) / G;
  struct Tile { int x0, y0, b0, n0, rot_c, rot_g; };
  auto tile_of = [&](int work) {
    Tile t;
    const int nblk = work % P.nblocks, m = work / P.nblocks;
    const int tx = m % P.tiles_x, ty = (m / P.tiles_x) % P.tiles_y, bg = m / (P.tiles_x * P.tiles_y);
    t.x0 = tx * TILE_W; t.y0 = ty * P.th; t.b0 = bg * P.nb; t.n0 = nblk * P.bn;
    // De-synchronise the CTAs: every CTA walks the same weight tiles, and when all 148 SMs ask the L2 for the SAME
    // lines at the same moment the few slices that hold them serialise the requests.  Each tile therefore starts its
    // reduction at its own rotation of the chunk and group order; the sum is order-independent up to fp32 rounding.
    const unsigned rot_seed = (unsigned)m * 2654435761u + (unsigned)nblk * 40503u;
    t.rot_c = (HALO && has_work) ? (int)((rot_seed >> 8) % (unsigned)ncc) : 0;
    t.rot_g = has_work ? (int)((rot_seed >> 4) % (unsigned)ngroups) : 0;
    return t;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s, 0), 1); mbar_init(full_bar(s, 1), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(fulla_bar(s), 1);
      mbar_init(emptya_bar(s), 2);      // both issuers release a patch
      mbar_init(tmemfull_bar(s), 2);    // ... and hand an accumulator set to the epilogue
      mbar_init(tmemempty_bar(s), 4);   // the four epilogue warps hand it back
      mbar_init(first_bar(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================== TMA producer ============================================
    // The whole warp walks the loop and ONE elected lane issues: in a converged warp the copy operands live in
    // uniform registers; under `if (lane == 0)` the compiler wraps every UTMALDG in a per-lane waterfall loop.
    if (has_work) {
      int s = 0, round = 0, pa = 0, pha = 0;
      const int nouter = HALO ? ncc : 1;
      for (int work = blockIdx.x; work < P.nwork; work += gridDim.x) {
        const Tile T = tile_of(work);
        for (int ci = 0; ci < nouter; ++ci) {
          int cc = cc_begin + (ci + T.rot_c) % ncc;
          if (HALO) {
            mbar_wait(emptya_bar(pa), pha ^ 1, P.error_flag, 1);
            if (elect_one()) {
              mbar_expect_tx(fulla_bar(pa), P.patch_tx * nops);
              tma_load_4d(patch_addr(pa, 0), &tmA, fulla_bar(pa), cc * P.kc, T.x0 - P.pad, T.y0 - P.pad, T.b0);
              if (nops == 2) tma_load_4d(patch_addr(pa, 1), &tmAlo, fulla_bar(pa), cc * P.kc, T.x0 - P.pad, T.y0 - P.pad, T.b0);
            }
            pa ^= 1;
            if (pa == 0) pha ^= 1;
          }
          int g = T.rot_g;
          for (int gi = 0; gi < ngroups; ++gi) {
            const int item0 = g * G, nt = min(G, nitems - item0);
            mbar_wait(empty_bar(s), (round & 1) ^ 1, P.error_flag, 2);
            const uint32_t fb = full_bar(s, round);
            if (elect_one()) {
              mbar_expect_tx(fb, (uint32_t)nt * item_bytes);
              uint32_t dst = slots_base + (uint32_t)s * slot_bytes;
              int tap = HALO ? item0 : item0 % taps;
              if (!HALO) cc = cc_begin + item0 / taps;
              for (int j = 0; j < nt; ++j) {
                if (!HALO) {
                  const int dy = tap / P.ks, dx = tap - dy * P.ks;
                  tma_load_4d(dst, &tmA, fb, cc * P.kc, T.x0 + dx - P.pad, T.y0 + dy - P.pad, T.b0);
                  if (nops == 2) tma_load_4d(dst + P.a_stage_bytes, &tmAlo, fb, cc * P.kc, T.x0 + dx - P.pad, T.y0 + dy - P.pad, T.b0);
                }
                if (BT) {
                  // rows [cc*64, cc*64+64) of the forward tiles (chunk = 64-wide N block, flipped tap): 8 KB runs
                  const int nb64 = P.bn >> 6, wtap = taps - 1 - tap;
                  for (int i = 0; i < nb64; ++i) {
                    const int nc = min((T.n0 >> 6) + i, P.bt_nchunks - 1);   // columns past the layer's Cin are never stored
                    const size_t woff = ((size_t)(nc * taps + wtap) * P.rows_pad + (size_t)cc * 64) * 128u;
                    bulk_load_1d(dst + b_off + (uint32_t)i * 8192u, P.wp + woff, 8192u, fb);
                    if (nops == 2) bulk_load_1d(dst + b_off + P.b_stage_bytes + (uint32_t)i * 8192u, P.wp_lo + woff, 8192u, fb);
                  }
                } else {
                // the weight tile of (channel chunk, tap, N block) is ONE contiguous, pre-swizzled run of bn*128 bytes
                const size_t woff = ((size_t)(cc * taps + tap) * P.rows_pad + T.n0) * 128u;
                bulk_load_1d(dst + b_off, P.wp + woff, P.b_stage_bytes, fb);
                if (nops == 2) bulk_load_1d(dst + b_off + P.b_stage_bytes, P.wp_lo + woff, P.b_stage_bytes, fb);
                }
                dst += item_bytes;
                if (++tap == taps) { tap = 0; if (!HALO) ++cc; }
              }
            }
            if (++g == ngroups) g = 0;
            if (++s == P.stages) { s = 0; ++round; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ===================================== MMA issuers =============================================
    // TWO issuer warps take alternate groups.  tcgen05.mma issue is nearly synchronous (the pipe queues ~2 instructions)
    // and the wait + descriptor arithmetic + commit of a group costs a single warp 200-400 cycles the tensor pipe spends
    // idle; with two warps one prepares its group while the other's MMAs run.  MMAs of both warps accumulate into the
    // same TMEM columns -- measured exact and MMA-bound in tools/ubench_tc3.cu; the only ordering that matters is that
    // the accumulate=0 MMAs of a tile's first group enter the pipe first (first_bar).
    // The warp stays converged and one ELECTED lane issues: descriptors and the instruction descriptor sit in uniform
    // registers and an MMA costs one issue slot.  (r01 SASS of the `if (lane == 0)` form: ~25 instructions incl. R2UR
    // round trips and a waterfall loop per UTCHMMA, ~130 clk each.)  Descriptors are (constant high word | start
    // address >> 4), advanced by adding 2 (= 32 bytes >> 4) per K step.
    // The x3 precisions issue TWO MMAs per K step instead of three: B_hi and B_lo of an item are adjacent in shared
    // memory, so A_hi x [B_hi ; B_lo] is one N = 2*bn instruction (TMEM columns [0, 2bn)) and A_lo x B_hi a second one
    // (columns [2bn, 3bn)); the epilogue adds the three column ranges.
    if (has_work) {
      const int me = (warp == 1) ? 0 : 1;
      int gcount = 0;   // groups issued by this CTA so far: issuer `me` owns those with (gcount & 1) == me
      int s = 0, round = 0, pa = 0, pha = 0;
      const uint32_t sbo_a = HALO ? (uint32_t)P.pitch * 128u : 1024u;
      const uint32_t hi_a = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t hi_b = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      auto lo_of = [](uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); };
      // BT: MN-major B -- LBO = 8 KB between 64-wide N blocks (B_lo's blocks follow B_hi's at the same stride, so the
      // N = 2*bn instruction still covers [B_hi ; B_lo]), SBO = 1024 B between 8-row K groups, 16 rows = 2048 B per K step
      constexpr uint32_t lbo_bt = ((8192u >> 4) - 1u) << 16;      // lo_of() already carries LBO field = 1
      constexpr uint32_t kstep_b = BT ? 128u : 2u;
      auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
      const uint32_t item_step = item_bytes >> 4, alo_step = (HALO ? P.patch_bytes : P.a_stage_bytes) >> 4;
      const int nouter = HALO ? ncc : 1;
      int iter = 0;
      for (int work = blockIdx.x; work < P.nwork; work += gridDim.x, ++iter) {
        const Tile T = tile_of(work);
        const int aset = iter % P.nsets, use = iter / P.nsets;     // accumulator set of this tile and its use count
        const uint32_t tm_d = tmem_base + (uint32_t)(aset * P.set_cols);
        const uint32_t tm_d2 = tm_d + 2u * (uint32_t)P.bn;
        const bool owner = (gcount & 1) == me;                      // this warp issues the tile's first (accumulate=0) group
        uint32_t acc = 1u;
        if (owner) {
          // the epilogue must have drained this accumulator set (tile iter - nsets) before it is overwritten
          mbar_wait(tmemempty_bar(aset), (use & 1) ^ 1, P.error_flag, 10);
          tc_fence_after();
          acc = 0u;
        }
        bool ordered = owner;   // non-owner: wait for the owner's first group before issuing into this set
        for (int ci = 0; ci < nouter; ++ci) {
          uint32_t pbase = 0;
          if (HALO) {
            mbar_wait(fulla_bar(pa), pha, P.error_flag, 3);
            pbase = lo_of(patch_addr(pa, 0));
          }
          int g = T.rot_g;
          for (int gi = 0; gi < ngroups; ++gi) {
            const int item0 = g * G, nt = min(G, nitems - item0);
            if ((gcount & 1) == me) {
              mbar_wait(full_bar(s, round), (round >> 1) & 1, P.error_flag, 4);
              if (!ordered) { mbar_wait(first_bar(aset), use & 1, P.error_flag, 9); ordered = true; }
              tc_fence_after();
              // descriptor arithmetic stays in converged code (uniform datapath); only the MMAs sit under the election
              uint32_t it = lo_of(slots_base + (uint32_t)s * slot_bytes);   // item 0 of the slot
              int dy = 0, dx = 0;
              if (HALO) { dy = item0 / P.ks; dx = item0 - dy * P.ks; }
              for (int j = 0; j < nt; ++j) {
                // HALO: the SWIZZLE_128B XOR is a pure function of the shared-memory ADDRESS bits (measured on B200,
                // profiles/r01_conv_probe.txt), exactly like the TMA write, so a 128-B-shifted start keeps base_offset 0
                // and the 8-row groups may sit at any multiple of 128 B (SBO = patch pitch).
                const uint32_t la = HALO ? pbase + (uint32_t)(dy * P.pitch + dx) * 8u : it;
                const uint32_t lb = it + (b_off >> 4) + (BT ? lbo_bt : 0u);
                if (elect_one()) {
                  if (nops == 2) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // one UMMA_K (8 tf32 / 16 bf16) = 32 bytes inside the 128-byte swizzle row
                      const uint32_t a1 = k ? 1u : acc;
                      umma_k<BF16>(tm_d2, desc(hi_a, la + alo_step + 2u * k), desc(hi_b, lb + kstep_b * k), P.idesc, a1);   // A_lo x B_hi
                      umma_k<BF16>(tm_d, desc(hi_a, la + 2u * k), desc(hi_b, lb + kstep_b * k), P.idesc2, a1);              // A_hi x [B_hi;B_lo]
                    }
                  } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_k<BF16>(tm_d, desc(hi_a, la + 2u * k), desc(hi_b, lb + kstep_b * k), P.idesc, k ? 1u : acc);
                  }
                }
                acc = 1u;
                it += item_step;
                if (HALO) { if (++dx == P.ks) { dx = 0; ++dy; } }
              }
              if (elect_one()) {
                umma_commit(empty_bar(s));  // frees this slot when the MMAs above have read it
                if (owner && ci == 0 && gi == 0)
                  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(first_bar(aset)) : "memory");
              }
            }
            ++gcount;
            if (++g == ngroups) g = 0;
            if (++s == P.stages) { s = 0; ++round; }
          }
          if (HALO) {
            if (elect_one()) umma_commit(emptya_bar(pa));
            pa ^= 1;
            if (pa == 0) pha ^= 1;
          }
        }
        // a non-owner without any group in this tile still observes the phase of first_bar (no skipped phases)
        if (!ordered) mbar_wait(first_bar(aset), use & 1, P.error_flag, 11);
        if (elect_one()) umma_commit(tmemfull_bar(aset));
      }
    }
  } else if (warp >= 2 && warp <= 5 && has_work) {
    // ===================================== epilogue ================================================
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;     // tile row == TMEM lane
    const int i = row & 7, gr = row >> 3;
    int iter = 0;
    for (int work = blockIdx.x; work < P.nwork; work += gridDim.x, ++iter) {
      const Tile T = tile_of(work);
      const int aset = iter % P.nsets, use = iter / P.nsets;
      mbar_wait_warp_backoff(tmemfull_bar(aset), use & 1, P.error_flag, 5);
      tc_fence_after();
      const int x = T.x0 + i, y = T.y0 + (gr % P.th), n = T.b0 + (gr / P.th);
      const bool valid = (x < P.W) && (y < P.H) && (n < P.B);
      float* orow = P.out + (((size_t)n * P.H + y) * P.W + x) * P.Cout + T.n0;
      const uint32_t trow = tmem_base + (uint32_t)(aset * P.set_cols) + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < P.bn; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        if (nops == 2) {  // x3: add A_hi*B_lo (columns bn..) and A_lo*B_hi (columns 2bn..)
          uint32_t u[16], t[16];
          tmem_ld16(trow + (uint32_t)(P.bn + c0), u);
          tmem_ld16(trow + (uint32_t)(2 * P.bn + c0), t);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = __float_as_uint((__uint_as_float(u[j]) + __uint_as_float(t[j])) + __uint_as_float(v[j]));
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int co = T.n0 + c0 + j;
            if (co < P.Cout) {
              float4 o;
              o.x = __uint_as_float(v[j + 0]); o.y = __uint_as_float(v[j + 1]);
              o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
              if (P.bias && blockIdx.z == 0) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(P.bias + co));
                o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
              }
              if (P.ksplits == 1) {
                *reinterpret_cast<float4*>(orow + c0 + j) = o;
              } else {  // K-split partial sums meet in the zero-initialised output
                atomicAdd(orow + c0 + j + 0, o.x); atomicAdd(orow + c0 + j + 1, o.y);
                atomicAdd(orow + c0 + j + 2, o.z); atomicAdd(orow + c0 + j + 3, o.w);
              }
            }
          }
        }
      }
      // hand the accumulator set back to the issuers (all tcgen05.ld of this warp have completed: wait::ld in tmem_ld16)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmemempty_bar(aset)) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------
// weight gradient:  dW[co, tap, ci] = sum_{b,y,x} dZ[b,y,x,co] * X[b, y+dy-p, x+dx-p, ci]
//   The reduction runs over pixels.  Both operands are read straight from the NHWC tensors as MN-MAJOR tiles:
//   a TMA box of (32 channels x 8 px x th rows) puts one pixel per 128-byte row, i.e. K (= pixels) along the
//   rows and 32 M/N elements (= channels) contiguous -- the canonical MN-major SWIZZLE_128B operand; a block of
//   128 input channels is four such boxes (LBO = bytes per box).  One MMA covers K = 8 pixels = one 1024-byte
//   swizzle atom.  The X patch of a pixel tile is loaded ONCE with its halo and every tap reads it through a
//   shifted descriptor (same trick as the forward HALO mode); the dZ tile is shared by all taps.  One CTA owns
//   (128-channel block, group of TG taps, BN block, pixel split) with TG accumulators side by side in TMEM;
//   partial sums of the pixel splits meet in the zero-initialised packed gradient via red.global.add.f32.
//   tf32x3 = three launches of this tf32x1 kernel on (X_lo,dZ), (X,dZ_lo), (X,dZ).
//   (A K-major formulation over NCHW copies was tried first: TMA rejects the 4-byte-misaligned innermost
//   coordinate a tap shift along x produces -- profiles/r01_conv_probe.txt.)
// ---------------------------------------------------------------------------------------------------
struct WgradParams {
  int B, H, W, Cin, Cout, ks, pad;
  int kpad;              // ceil32/ceil64(Cin): row pitch of the packed gradient
  int bf16;              // 1: bf16 operands (kind::f16, 64 channels per 128-byte block, plain SWIZZLE_128B)
  int blk_ch, mblks;     // channels per 128-byte block (32 tf32 / 64 bf16) and X blocks per CTA (128 / blk_ch)
  int bn, tg;            // N tile (multiple of blk_ch), taps per CTA
  int acc_stride;        // TMEM columns between the accumulators of two taps (power of two >= bn)
  int tap_groups;
  int th;                // pixel-tile rows (tile = 8 x th pixels = th MMA k-steps)
  int tiles_x, tiles_y, ntiles;   // pixel tiles per image / in total (B * tiles_y * tiles_x)
  int psplits;
  int stages;
  int pitch;             // pixels per patch row (8 + ks - 1)
  int pair;              // 1: Cin <= 64 (bf16): M block 1 is the X patch shifted to the NEXT tap -> two taps per accumulator
  uint32_t patch_bytes;  // one channel block of the X patch: pitch * (th + ks - 1) * 128 rounded up to 1024
  uint32_t patch_tx;     // bytes one patch box delivers
  uint32_t gblk_bytes;   // 8 * th * 128: one channel block of the dZ tile
  uint32_t tmem_cols, idesc;
  float* dwp;            // [Cout][taps][kpad]
  unsigned int* error_flag;
};

// MN-major descriptor for 32-bit (tf32) operands.  Transposing 4-byte elements needs the 32-byte-atom flavour of
// the 128-byte swizzle on BOTH sides: TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B and descriptor layout type 1
// (SWIZZLE_128B_BASE32B, cute Swizzle<2,5,2>, K atom = 4 rows); with the plain SWIZZLE_128B layout type the MMA
// returns zeros on B200 (measured, profiles/r01_conv_probe.txt).  LBO = byte distance between 32-element blocks
// along M/N, SBO = byte distance between the 4-row groups along K (UMMA_K = 8 rows = 2 groups).
// 16-bit (bf16) MN-major operands use the ordinary SWIZZLE_128B (layout type 2, K atom = 8 rows).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 1u) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS_IGEMM, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXlo,
                  const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmGlo, const WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int nops = SPLIT ? 2 : 1;
  const int nblk_n = P.bn / P.blk_ch;
  // stage: [X_hi blocks][X_lo blocks][dZ_hi blocks][dZ_lo blocks].  The x3 precisions run the three error-compensated
  // products X_lo*dZ_hi, X_hi*dZ_lo, X_hi*dZ_hi as three MMAs per K step into the SAME accumulator: one launch, one
  // epilogue and 4 tile loads where three x1 launches needed 6.
  const uint32_t x_bytes = (uint32_t)P.mblks * P.patch_bytes, g_bytes = (uint32_t)nblk_n * P.gblk_bytes;
  const uint32_t stage_bytes = nops * (x_bytes + g_bytes);
  const uint32_t bars_base = smem_base + P.stages * stage_bytes;
  auto full_bar = [&](int s) { return bars_base + 8u * s; };
  auto empty_bar = [&](int s) { return bars_base + 8u * (MAX_STAGES + s); };
  const uint32_t tmemfull_bar = bars_base + 8u * (2 * MAX_STAGES);
  const uint32_t tmem_slot = bars_base + 8u * (2 * MAX_STAGES + 1);
  auto stage_x = [&](int s, int op, int j) { return smem_base + s * stage_bytes + op * x_bytes + (uint32_t)j * P.patch_bytes; };
  auto stage_g = [&](int s, int op, int i) { return smem_base + s * stage_bytes + nops * x_bytes + op * g_bytes + (uint32_t)i * P.gblk_bytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = blockIdx.x * 128;
  const int tgi = blockIdx.y % P.tap_groups, nblk = blockIdx.y / P.tap_groups;
  const int n0 = nblk * P.bn;
  const int taps = P.ks * P.ks;
  const int tap0 = tgi * P.tg * (P.pair ? 2 : 1);   // pair mode: every accumulator holds two taps
  const int nacc = P.pair ? min(P.tg, (taps - tap0 + 1) >> 1) : min(P.tg, taps - tap0);
  const int per = (P.ntiles + P.psplits - 1) / P.psplits;
  const int t_begin = blockIdx.z * per, t_end = min(P.ntiles, t_begin + per);
  const bool has_work = t_end > t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 2); }   // two MMA issuers
    mbar_init(tmemfull_bar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (has_work) {   // converged warp, one elected lane issues (see conv_igemm_kernel)
      int s = 0, ph = 0;
      // CTAs that share a pixel split walk the same tiles: start each at its own rotation (L2 hot-spot avoidance)
      const int nt = t_end - t_begin;
      const int rot = (int)((blockIdx.x * 2654435761u + blockIdx.y * 40503u) >> 8) % nt;
      const uint32_t tx_bytes = nops * ((uint32_t)P.mblks * P.patch_tx + g_bytes);
      int t = t_begin + rot;
      for (int ti = 0; ti < nt; ++ti) {
        const int tx = t % P.tiles_x, ty = (t / P.tiles_x) % P.tiles_y, b = t / (P.tiles_x * P.tiles_y);
        const int x0 = tx * TILE_W, y0 = ty * P.th;
        mbar_wait(empty_bar(s), ph ^ 1, P.error_flag, 6);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), tx_bytes);
          for (int j = 0; j < P.mblks; ++j) {
            tma_load_4d(stage_x(s, 0, j), &tmX, full_bar(s), ci0 + P.blk_ch * j, x0 - P.pad, y0 - P.pad, b);
            if (SPLIT) tma_load_4d(stage_x(s, 1, j), &tmXlo, full_bar(s), ci0 + P.blk_ch * j, x0 - P.pad, y0 - P.pad, b);
          }
          for (int i = 0; i < nblk_n; ++i) {
            tma_load_4d(stage_g(s, 0, i), &tmG, full_bar(s), n0 + P.blk_ch * i, x0, y0, b);
            if (SPLIT) tma_load_4d(stage_g(s, 1, i), &tmGlo, full_bar(s), n0 + P.blk_ch * i, x0, y0, b);
          }
        }
        if (++t == t_end) t = t_begin;
        if (++s == P.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // Two issuer warps (see conv_igemm_kernel): warp 1 takes the even accumulators (taps / tap pairs) of every stage,
    // warp 6 the odd ones.  They never share an accumulator, both observe every phase of every full barrier, and a
    // stage is released when both have committed.
    if (has_work) {
      const int me = (warp == 1) ? 0 : 1;
      int s = 0, ph = 0;
      uint32_t acc = 0;
      // descriptor = constant high word | (start address >> 4); BF16: plain SWIZZLE_128B (layout 2), one MMA per PAIR
      // of tile rows (K = 16 pixels = two 8-row groups; patch rows pitch*128 B apart, dZ rows 1024 B).  TF32: layout 1
      // (SWIZZLE_128B_BASE32B), one MMA per tile row (K = 8 pixels = two 4-row groups 512 B apart).
      const uint32_t layout = BF16 ? 2u : 1u;
      const uint32_t sbo_a = BF16 ? (uint32_t)P.pitch * 128u : 512u, sbo_b = BF16 ? 1024u : 512u;
      const uint32_t hi_a = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      const uint32_t hi_b = ((sbo_b >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      const uint32_t lbo_a = P.pair ? 0u : ((P.patch_bytes >> 4) & 0x3FFFu) << 16, lbo_b = ((P.gblk_bytes >> 4) & 0x3FFFu) << 16;
      const int tstep = P.pair ? 2 : 1;
      auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
      constexpr int KSTEP = BF16 ? 2 : 1;   // tile rows per MMA
      const uint32_t xlo_off = x_bytes >> 4, glo_off = g_bytes >> 4;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(full_bar(s), ph, P.error_flag, 7);
        tc_fence_after();
        const uint32_t xa = (stage_x(s, 0, 0) >> 4) | lbo_a, ga = (stage_g(s, 0, 0) >> 4) | lbo_b;
        int dy = tap0 / P.ks, dx = tap0 % P.ks;
        uint32_t d_t = tmem_base;
        if (me) {
          d_t += (uint32_t)P.acc_stride;
          dx += tstep;
          if (dx >= P.ks) { dx -= P.ks; ++dy; }
        }
        for (int tt = me; tt < nacc; tt += 2) {
          uint32_t la0 = xa + (uint32_t)(dy * P.pitch + dx) * 8u;   // 128 B per patch pixel >> 4
          // pair mode: M rows 64..127 read the SAME patch one tap further: +1 pixel (128 B), or the start of the next
          // tap row ((pitch - ks + 1) = 8 pixels, 1024 B) -- the descriptor's M-block stride (LBO) expresses the shift
          if (P.pair) la0 |= ((dx + 1 < P.ks) ? 8u : 64u) << 16;
          if (elect_one()) {
            uint32_t la = la0, lb = ga, a1 = acc;
            for (int kk = 0; kk < P.th; kk += KSTEP) {
              if (SPLIT) {   // small terms first
                umma_k<BF16>(d_t, desc(hi_a, la + xlo_off), desc(hi_b, lb), P.idesc, a1);
                umma_k<BF16>(d_t, desc(hi_a, la), desc(hi_b, lb + glo_off), P.idesc, 1u);
                umma_k<BF16>(d_t, desc(hi_a, la), desc(hi_b, lb), P.idesc, 1u);
              } else {
                umma_k<BF16>(d_t, desc(hi_a, la), desc(hi_b, lb), P.idesc, a1);
              }
              a1 = 1u;
              la += (uint32_t)(KSTEP * P.pitch * 8);
              lb += (uint32_t)(KSTEP * 64);
            }
          }
          d_t += 2u * (uint32_t)P.acc_stride;
          dx += 2 * tstep;
          while (dx >= P.ks) { dx -= P.ks; ++dy; }
        }
        if (elect_one()) umma_commit(empty_bar(s));
        acc = 1u;
        if (++s == P.stages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tmemfull_bar);
    }
  } else if (warp >= 2 && warp <= 5 && has_work) {
    mbar_wait_warp_backoff(tmemfull_bar, 0, P.error_flag, 8);
    tc_fence_after();
    const int q = warp & 3;
    // pair mode: accumulator rows 0..63 are tap 2t, rows 64..127 tap 2t + 1, both over input channels 0..63
    const int ci = P.pair ? ((q & 1) * 32 + lane) : (ci0 + q * 32 + lane);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < nacc; ++t) {
      const int tap = P.pair ? (tap0 + 2 * t + (q >> 1)) : (tap0 + t);
      const bool valid = (ci < P.kpad) && (tap < taps);
      for (int c0 = 0; c0 < P.bn; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(trow + (uint32_t)(t * P.acc_stride + c0), v);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = n0 + c0 + j;
            if (co < P.Cout) atomicAdd(P.dwp + ((size_t)co * taps + tap) * P.kpad + ci, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

// packed gradient [Cout][tap][kpad] -> OIHW [Cout][Cin][k][k]
__global__ void unpack_weight_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin, int taps, int kpad) {
  const size_t total = (size_t)Cout * Cin * taps;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(idx % taps);
    const int ci = (int)((idx / taps) % Cin);
    const int co = (int)(idx / ((size_t)taps * Cin));
    dw[idx] = dwp[((size_t)co * taps + tap) * kpad + ci];
  }
}

// ---------------------------------------------------------------------------------------------------
// operand preparation kernels
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// lo = x - trunc_tf32(x): the part of x a tf32 tensor-core operand read drops
__global__ void tf32_residual_kernel(const float4* __restrict__ x, float4* __restrict__ lo, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 r;
    r.x = v.x - tf32_trunc(v.x); r.y = v.y - tf32_trunc(v.y); r.z = v.z - tf32_trunc(v.z); r.w = v.w - tf32_trunc(v.w);
    lo[i] = r;
  }
}

// Weight packing.  Output = the exact shared-memory image of every B stage tile, so the kernel fetches a tile with ONE
// contiguous bulk copy:   [cchunk][tap][rows_pad][128 bytes],  row = output channel (fprop) / input channel (dgrad),
// 128 bytes = kc consecutive reduction channels of that chunk, 16-byte groups XOR-swizzled with (row & 7) exactly as
// SWIZZLE_128B lays a K-major tile out (rows are 128 B apart, tiles start at multiples of 8 rows).
//   fprop  (transposed=0): value(row=co, k=ci, tap) = w[co][ci][tap]
//   dgrad  (transposed=1): value(row=ci, k=co, tap) = w[co][ci][taps-1-tap]     (correlation with the flipped kernel)
// Rows >= the real row count and k >= the real reduction size are zero.
template <typename T>
__device__ __forceinline__ void store_split(T* hi, T* lo, size_t idx, float v);
template <>
__device__ __forceinline__ void store_split<float>(float* hi, float* lo, size_t idx, float v) {
  hi[idx] = v;
  if (lo) lo[idx] = v - tf32_trunc(v);
}
template <>
__device__ __forceinline__ void store_split<__nv_bfloat16>(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  if (lo) lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wp, T* __restrict__ wp_lo, int Cout, int Cin,
                                   int ks, int kc, int rows_pad, int transposed) {
  const int taps = ks * ks;
  const int rows = transposed ? Cin : Cout, kred = transposed ? Cout : Cin;
  const int cchunks = (kred + kc - 1) / kc;
  const int epc = 16 / (int)sizeof(T);            // elements per 16-byte swizzle group
  const size_t total = (size_t)cchunks * taps * rows_pad * kc;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % kc);                 // position inside the (swizzled) 128-byte row
    const int r = (int)((idx / kc) % rows_pad);
    const int tap = (int)((idx / ((size_t)kc * rows_pad)) % taps);
    const int cc = (int)(idx / ((size_t)kc * rows_pad * taps));
    const int grp = e / epc, within = e % epc;
    const int k = cc * kc + ((grp ^ (r & 7)) * epc + within);   // logical reduction index stored at this position
    float v = 0.0f;
    if (r < rows && k < kred) {
      if (!transposed) v = w[((size_t)r * Cin + k) * taps + tap];
      else v = w[((size_t)k * Cin + r) * taps + (taps - 1 - tap)];
    }
    store_split<T>(wp, wp_lo, idx, v);
  }
}

__global__ void split_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const float in[4] = {v.x, v.y, v.z, v.w};
    unsigned short h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(in[k]);
      const __nv_bfloat16 lb = __float2bfloat16_rn(in[k] - __bfloat162float(hb));
      h[k] = __bfloat16_as_ushort(hb);
      l[k] = __bfloat16_as_ushort(lb);
    }
    hi[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    lo[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, bool bf16 = false) {
  EncodeTiledFn fn = encode_fn();
  PN_REQUIRE(fn != nullptr, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, PN_ERR_BAD_ARGUMENT, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PN_OK;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

static bool is_bf16(int precision) { return precision == PN_PRECISION_BF16X3 || precision == PN_PRECISION_BF16X1; }
static bool is_split(int precision) { return precision == PN_PRECISION_BF16X3 || precision == PN_PRECISION_TF32X3; }
int kc_of(int precision) { return is_bf16(precision) ? 64 : 32; }
int kpad_of(int c, int precision) { const int kc = kc_of(precision); return (c + kc - 1) / kc * kc; }
// N tile of the forward/dgrad kernel and the row padding of the packed weight (a multiple of the N tile)
int bn_of(int rows) { const int r16 = (rows + 15) / 16 * 16; return r16 < 128 ? r16 : 128; }
// ... and of 64: the data gradient reads the forward packing in 64-ROW reduction chunks (BT), whose tail must be zero rows
int rows_pad_of(int rows) {
  const int bn = bn_of(rows);
  int r = (rows + bn - 1) / bn * bn;
  while (r % 64) r += bn;
  return r;
}
size_t packed_weight_elems(int rows, int kred, int ksize, int precision) {
  const int kc = kc_of(precision);
  return (size_t)((kred + kc - 1) / kc) * ksize * ksize * rows_pad_of(rows) * kc;
}

// x [B,H,W,Cin] NHWC, wp [Cout][k*k][kpad(Cin)] -> y [B,H,W,Cout]
static int conv_forward(const pn_conv_desc* d, const void* x, const void* x_lo, const void* wp, const void* wp_lo,
                        const float* bias, float* y, unsigned int* error_flag, cudaStream_t stream, bool bt = false) {
  PN_REQUIRE(d && x && wp && y, PN_ERR_BAD_ARGUMENT, "pn_conv2d: null argument");
  PN_REQUIRE(d->batch > 0 && d->height > 0 && d->width > 0 && d->cin > 0 && d->cout > 0, PN_ERR_BAD_ARGUMENT,
             "pn_conv2d: bad shape");
  PN_REQUIRE(d->ksize >= 1 && d->ksize <= 7 && (d->ksize & 1), PN_ERR_UNSUPPORTED, "pn_conv2d: ksize %d (odd, <= 7)", d->ksize);
  PN_REQUIRE(d->precision == PN_PRECISION_TF32X1 || d->precision == PN_PRECISION_TF32X3 || is_bf16(d->precision),
             PN_ERR_BAD_ARGUMENT, "pn_conv2d: precision %d", d->precision);
  const bool bf16 = is_bf16(d->precision);
  const int esize = bf16 ? 2 : 4, kc = kc_of(d->precision);
  PN_REQUIRE(!bt || bf16, PN_ERR_UNSUPPORTED, "pn_conv2d_dgrad: the forward packing is read as an MN-major operand in the bf16 precisions only");
  PN_REQUIRE(d->cin % (16 / esize) == 0 && d->cout % 4 == 0, PN_ERR_UNSUPPORTED,
             "pn_conv2d: Cin (%d) must be a multiple of %d (16-byte TMA pitch) and Cout (%d) of 4", d->cin, 16 / esize, d->cout);
  PN_REQUIRE(!is_split(d->precision) || (x_lo && wp_lo), PN_ERR_BAD_ARGUMENT,
             "pn_conv2d: the x3 precisions need the residual operands x_lo and w_lo");
  PN_REQUIRE(aligned16(x) && aligned16(wp) && aligned16(y) && (!bias || aligned16(bias)), PN_ERR_ALIGNMENT,
             "pn_conv2d: pointers must be 16-byte aligned");

  KernelParams P{};
  P.B = d->batch; P.H = d->height; P.W = d->width; P.Cin = d->cin; P.Cout = d->cout; P.ks = d->ksize; P.pad = d->ksize / 2;
  P.nsplit = is_split(d->precision) ? 3 : 1;
  P.kc = kc; P.bf16 = bf16 ? 1 : 0;
  const int nops = (P.nsplit == 3) ? 2 : 1;
  // Tile rows (th) x folded images (nb = 16 / th).  The default is the tallest tile that the map fills; on the small maps
  // the choice is made by counting WAVES of the one-CTA-per-SM grid (r02l layer table: 512->512 3x3 at 12x40 ran 160 CTAs =
  // two waves, the second one 12 CTAs wide, at 90 TFLOP/s; 4 rows x 4 images gives 15 full tiles -> 120 CTAs, one wave of
  // half the length).  cost = (work items of the busiest CTA) x (its share of the reduction loop) + an epilogue term.
  const int nblocks_c = bt ? (d->cout + (d->cout <= 64 ? 64 : 128) - 1) / (d->cout <= 64 ? 64 : 128)
                           : (d->cout + bn_of(d->cout) - 1) / bn_of(d->cout);
  const int cchunks_c = (d->cin + kc - 1) / kc;
  const int sms = sm_count();
  auto plan_of = [&](int th, int* ks_out) -> double {
    const int nb = TILE_ROWS / th;
    const int tiles = ((d->width + TILE_W - 1) / TILE_W) * ((d->height + th - 1) / th) * ((d->batch + nb - 1) / nb);
    const int base = tiles * nblocks_c;
    const int taps = d->ksize * d->ksize;
    int ks = 1;
    if (base < sms) {
      ks = sms / base;
      if (ks > cchunks_c) ks = cchunks_c;
      const int per = (cchunks_c + ks - 1) / ks;
      ks = (cchunks_c + per - 1) / per;   // no empty splits
    }
    *ks_out = ks;
    const int per = (cchunks_c + ks - 1) / ks;
    const int rounds = (base * ks + sms - 1) / sms;            // persistent loop / waves of the plain grid
    const bool halo = (nb == 1 && d->ksize > 1);
    // per-tap staging re-fetches the activation tile for every tap (L2 traffic, one TMA per item): a few per cent
    const double item = halo ? 1.0 : 1.06;
    return rounds * (per * taps * item + (ks > 1 ? 4.0 : 2.0));
  };
  int th_auto = TILE_ROWS;
  while (th_auto > 1 && th_auto / 2 >= d->height) th_auto /= 2;
  P.th = th_auto;
  int ks_plan = 1;
  if (!(d->debug_flags & 8192) && d->mode == PN_CONV_MODE_AUTO) {
    double best = plan_of(th_auto, &ks_plan);
    for (int th = th_auto / 2; th >= 1 && TILE_ROWS / th <= d->batch; th /= 2) {
      int ks_c;
      const double c = plan_of(th, &ks_c);
      if (c < 0.93 * best) { best = c; P.th = th; ks_plan = ks_c; }
    }
  } else {
    plan_of(P.th, &ks_plan);
  }
  {
    // tuning knob (debug_flags bits 16..19): force the tile height to 1, 2, 4 or 8 rows
    const int forced_th = (d->debug_flags >> 16) & 15;
    if (forced_th == 1 || forced_th == 2 || forced_th == 4 || forced_th == 8) { P.th = forced_th; plan_of(P.th, &ks_plan); }
  }
  P.nb = TILE_ROWS / P.th;
  P.halo = (d->mode == PN_CONV_MODE_HALO) || (d->mode == PN_CONV_MODE_AUTO && P.nb == 1 && d->ksize > 1);
  if (d->mode == PN_CONV_MODE_PER_TAP) P.halo = 0;
  PN_REQUIRE(!(P.halo && P.nb != 1), PN_ERR_UNSUPPORTED, "pn_conv2d: halo mode needs maps at least 9 rows tall");
  P.force_base_offset0 = (d->debug_flags & 1) ? 1 : 0;
  P.tiles_x = (d->width + TILE_W - 1) / TILE_W;
  P.tiles_y = (d->height + P.th - 1) / P.th;
  const int bgroups = (d->batch + P.nb - 1) / P.nb;
  // N tile (<= 128, multiple of 16); the packed weight is padded to a multiple of it
  // (BT: whole 64-wide N blocks of the forward packing; its rows are the reduction dimension d->cin here)
  const int bn = bt ? (d->cout <= 64 ? 64 : 128) : bn_of(d->cout);
  P.bn = bn;
  P.rows_pad = bt ? rows_pad_of(d->cin) : rows_pad_of(d->cout);
  P.bt_nchunks = (d->cout + 63) / 64;
  P.wp = static_cast<const uint8_t*>(wp);
  P.wp_lo = static_cast<const uint8_t*>(wp_lo ? wp_lo : wp);
  P.cchunks = (d->cin + kc - 1) / kc;
  P.a_stage_bytes = TILE_W * TILE_ROWS * 128u;
  P.b_stage_bytes = (uint32_t)bn * 128u;
  // exact patch pitch: the MMA descriptor takes any 128-byte multiple as the 8-row-group stride (SBO)
  P.pitch = (d->debug_flags & 64) ? PATCH_PITCH : TILE_W + d->ksize - 1;
  P.patch_tx = (uint32_t)P.pitch * (P.th + d->ksize - 1) * 128u;
  P.patch_bytes = (P.patch_tx + 1023u) & ~1023u;
  P.set_cols = (int)pow2_cols(P.nsplit == 3 ? 3 * bn : bn);
  // instruction descriptor: fp32 accumulate, A/B format 2 = TF32 (kind::tf32) or 1 = BF16 (kind::f16), K-major, M = 128
  const uint32_t fmt = bf16 ? 1u : 2u;
  const uint32_t bmaj = bt ? (1u << 16) : 0u;     // B operand MN-major
  P.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | bmaj | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  P.idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | bmaj | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  P.bias = bias; P.out = y; P.error_flag = error_flag;

  const uint32_t patch_region = P.halo ? 2u * nops * P.patch_bytes : 0u;
  const uint32_t item_bytes = (P.halo ? 0u : nops * P.a_stage_bytes) + nops * P.b_stage_bytes;
  const uint32_t budget = 227u * 1024u - 1024u /*alignment slack*/ - 512u /*barriers*/;
  PN_REQUIRE(patch_region + 2 * item_bytes <= budget, PN_ERR_UNSUPPORTED, "pn_conv2d: tile does not fit in shared memory");
  // Items per slot: amortise the wait/commit hand-off over >= 16 MMAs (x3: 8 MMAs per item, x1: 4), but keep at least
  // three slots in flight so that a slot's refill (TMA latency) hides behind the other slots' MMAs.
  // (pack1, x3, 6 items fit: 2 x 3 slots 1.64 ms; 1 x 6 1.86 ms; 2 x 2 1.96 ms; 3 x 2 1.94 ms -- r01_conv_bench.)
  {
    const int total_items = (P.halo ? 1 : P.cchunks) * d->ksize * d->ksize;
    const int fit = (int)((budget - patch_region) / item_bytes);
    int group = (P.nsplit == 3) ? 2 : 4;
    const int forced = (d->debug_flags >> 8) & 15;   // tuning knob: force the group size
    if (forced) group = forced;
    while (group > 1 && (fit / group < (forced ? 2 : 3) || group > total_items)) --group;
    int stages = fit / group;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    P.group = group;
    P.stages = stages;
  }
  const size_t smem = 1024 + patch_region + (size_t)P.stages * P.group * item_bytes + 512;

  // tensor maps
  alignas(64) CUtensorMap tmA, tmAlo;
  {
    const uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->width, (uint64_t)d->height, (uint64_t)d->batch};
    const uint64_t strides[3] = {(uint64_t)d->cin * esize, (uint64_t)d->width * d->cin * esize,
                                 (uint64_t)d->height * d->width * d->cin * esize};
    uint32_t box[4];
    if (P.halo) { box[0] = kc; box[1] = (uint32_t)P.pitch; box[2] = P.th + d->ksize - 1; box[3] = 1; }
    else        { box[0] = kc; box[1] = TILE_W;      box[2] = P.th;                box[3] = P.nb; }
    int rc = make_map(&tmA, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, bf16);
    if (rc) return rc;
    rc = make_map(&tmAlo, x_lo ? x_lo : x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, bf16);
    if (rc) return rc;
  }
  // small maps: split the channel-chunk loop so that the grid covers the SMs without spilling into a second wave
  {
    int ks = ks_plan;
    if (d->debug_flags & 8192) {      // round-1 rule (A/B): round the split UP to at least 148 CTAs
      const int base = P.tiles_x * P.tiles_y * bgroups * ((d->cout + bn - 1) / bn);
      ks = 1;
      if (base < 120) ks = (148 + base - 1) / base;
      if (ks > P.cchunks) ks = P.cchunks;
      if (ks > 1) {
        const int per = (P.cchunks + ks - 1) / ks;
        ks = (P.cchunks + per - 1) / per;   // no empty splits
      }
    }
    P.ksplits = ks;
    if (ks > 1) PN_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)d->batch * d->height * d->width * d->cout, stream));
  }
  // work items and the persistent grid: more than one wave of tiles without a K split -> one CTA per SM walks them,
  // with two accumulator sets in TMEM when they fit (debug flag 4096 forces one tile per CTA)
  P.nblocks = (d->cout + bn - 1) / bn;
  P.nwork = P.tiles_x * P.tiles_y * bgroups * P.nblocks;
  const bool persistent = (P.ksplits == 1) && (P.nwork > sms) && !(d->debug_flags & 4096);
  P.nsets = (persistent && 2 * P.set_cols <= 512) ? 2 : 1;
  P.tmem_cols = pow2_cols(P.nsets * P.set_cols);
  dim3 grid(persistent ? sms : P.nwork, 1, P.ksplits);
  auto launch = [&](auto kern) -> int {
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NTHREADS_IGEMM, smem, stream>>>(tmA, tmAlo, P);
    return 0;
  };
  int lrc;
  const int variant = (bt ? 8 : 0) | (P.bf16 ? 4 : 0) | (P.nsplit == 3 ? 2 : 0) | (P.halo ? 1 : 0);
  switch (variant) {
    case 12: lrc = launch(conv_igemm_kernel<true, false, false, true>); break;
    case 13: lrc = launch(conv_igemm_kernel<true, false, true, true>); break;
    case 14: lrc = launch(conv_igemm_kernel<true, true, false, true>); break;
    case 15: lrc = launch(conv_igemm_kernel<true, true, true, true>); break;
    case 0: lrc = launch(conv_igemm_kernel<false, false, false>); break;
    case 1: lrc = launch(conv_igemm_kernel<false, false, true>); break;
    case 2: lrc = launch(conv_igemm_kernel<false, true, false>); break;
    case 3: lrc = launch(conv_igemm_kernel<false, true, true>); break;
    case 4: lrc = launch(conv_igemm_kernel<true, false, false>); break;
    case 5: lrc = launch(conv_igemm_kernel<true, false, true>); break;
    case 6: lrc = launch(conv_igemm_kernel<true, true, false>); break;
    default: lrc = launch(conv_igemm_kernel<true, true, true>); break;
  }
  if (lrc) return lrc;
  count_launch();
  return check_launch("conv_igemm_kernel");
}

// x [B,H,W,Cin], g [B,H,W,Cout] (NHWC; fp32 or bf16) -> dwp [Cout][k*k][kpad(Cin)] fp32, ACCUMULATED (caller zeroes it)
static int conv_wgrad(const pn_conv_desc* d, const void* x, const void* x_lo, const void* g, const void* g_lo, float* dwp,
                      unsigned int* error_flag, cudaStream_t stream) {
  PN_REQUIRE(d && x && g && dwp, PN_ERR_BAD_ARGUMENT, "pn_conv2d_wgrad: null argument");
  PN_REQUIRE(d->ksize >= 1 && d->ksize <= 7 && (d->ksize & 1), PN_ERR_UNSUPPORTED, "pn_conv2d_wgrad: ksize %d", d->ksize);
  const bool bf16 = is_bf16(d->precision), split = is_split(d->precision);
  const int esize = bf16 ? 2 : 4, blk_ch = bf16 ? 64 : 32, nops = split ? 2 : 1;
  PN_REQUIRE(d->cin % (16 / esize) == 0 && d->cout % (16 / esize) == 0, PN_ERR_UNSUPPORTED,
             "pn_conv2d_wgrad: Cin/Cout must be multiples of %d", 16 / esize);
  PN_REQUIRE(aligned16(x) && aligned16(g) && aligned16(dwp), PN_ERR_ALIGNMENT, "pn_conv2d_wgrad: alignment");
  WgradParams P{};
  P.B = d->batch; P.H = d->height; P.W = d->width; P.Cin = d->cin; P.Cout = d->cout; P.ks = d->ksize; P.pad = d->ksize / 2;
  P.kpad = kpad_of(d->cin, d->precision);
  P.bf16 = bf16 ? 1 : 0; P.blk_ch = blk_ch; P.mblks = 128 / blk_ch;
  const int taps = d->ksize * d->ksize;
  // few input channels: one 64-channel block would leave half of the M = 128 rows empty -> pair two taps per accumulator
  P.pair = (bf16 && P.kpad <= 64 && taps > 1 && !(d->debug_flags & 256)) ? 1 : 0;
  if (P.pair) P.mblks = 1;
  int bn = (d->cout + blk_ch - 1) / blk_ch * blk_ch;
  if (bn > 256) bn = 256;
  P.bn = bn;
  P.acc_stride = (int)pow2_cols(bn);
  int tg = 512 / P.acc_stride;
  if (tg > taps) tg = taps;
  if (d->debug_flags & 2) tg = 1;
  if (P.pair && tg > (taps + 1) / 2) tg = (taps + 1) / 2;
  {
    // Balanced tap groups: every CTA of a (channel block, pixel split) loads the same tiles whatever its tap count, so a
    // short last group (9 taps as 8 + 1: r02n ncu of 136 -> 64 3x3 at 192x640, SMs 60 % active) idles its SMs.  Among
    // gmin .. gmin + 2 groups take the one with the least padded work groups x ceil(units / groups), fewest groups on ties.
    const int units = P.pair ? (taps + 1) / 2 : taps;      // accumulators (taps or tap pairs) over all groups
    const int gmin = (units + tg - 1) / tg;
    int groups = gmin;
    if (!(d->debug_flags & 8192)) {
      int best = gmin * ((units + gmin - 1) / gmin);
      for (int g = gmin + 1; g <= gmin + 2 && g <= units; ++g) {
        const int padded = g * ((units + g - 1) / g);
        if (padded < best) { best = padded; groups = g; }
      }
      tg = (units + groups - 1) / groups;
    }
    P.tg = tg;
    P.tap_groups = (units + tg - 1) / tg;
  }
  P.pitch = TILE_W + d->ksize - 1;
  // pixel tile height: as tall as shared memory allows with >= 2 stages (X-patch blocks + dZ blocks per stage, hi and lo
  // for the x3 precisions); bf16 consumes tile rows in pairs (K = 16 pixels per MMA), so its height is even
  const uint32_t budget = 227u * 1024u - 1024u - 512u;
  int th = TILE_ROWS;
  while (th > 2 && th / 2 >= d->height) th /= 2;
  if (!bf16 && d->height < th) th = d->height;
  auto patch_of = [&](int t) { return ((uint32_t)P.pitch * (t + d->ksize - 1) * 128u + 1023u) & ~1023u; };
  auto stage_bytes_of = [&](int t) {
    return (uint32_t)nops * ((uint32_t)P.mblks * patch_of(t) + (uint32_t)(bn / blk_ch) * 8u * t * 128u);
  };
  while (th > (bf16 ? 2 : 1) && 2u * stage_bytes_of(th) > budget) th = bf16 ? th / 2 : (th + 1) / 2;
  PN_REQUIRE(stage_bytes_of(th) <= budget, PN_ERR_UNSUPPORTED, "pn_conv2d_wgrad: tile does not fit in shared memory");
  P.th = th;
  P.patch_tx = (uint32_t)P.pitch * (th + d->ksize - 1) * 128u;
  P.patch_bytes = patch_of(th);
  P.gblk_bytes = 8u * th * 128u;
  const uint32_t stage_bytes = stage_bytes_of(th);
  int stages = (int)(budget / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 1) stages = 1;
  P.stages = stages;
  P.tiles_x = (d->width + TILE_W - 1) / TILE_W;
  P.tiles_y = (d->height + th - 1) / th;
  P.ntiles = d->batch * P.tiles_x * P.tiles_y;
  const int mblocks = (P.kpad + 127) / 128, nblocks = (d->cout + bn - 1) / bn;
  const int base_ctas = mblocks * P.tap_groups * nblocks;
  // Pixel split: the kernel runs one CTA per SM, so the grid is counted in WAVES.  cost(ps) = waves x (pixel tiles per CTA +
  // an epilogue term for the red.add of the CTA's accumulators); fewest splits among the near-best (fewer atomics).
  // (round 1 aimed at 2 x 148 CTAs whatever the wave count: 512->512 3x3 at 12x40 ran 320 CTAs = 2.2 waves.)
  int psplits = 1;
  if (d->debug_flags & 8192) {
    psplits = (2 * 148 + base_ctas - 1) / base_ctas;
    if (psplits > P.ntiles) psplits = P.ntiles;
    if (psplits < 1) psplits = 1;
  } else {
    const int sms = sm_count();
    double best = 1e30;
    for (int ps = 1; ps <= P.ntiles; ++ps) {
      const int per = (P.ntiles + ps - 1) / ps;
      if (ps > 1 && (P.ntiles + per - 1) / per != ps) continue;        // same work per CTA with fewer CTAs exists
      const int waves = (base_ctas * ps + sms - 1) / sms;
      const double c = waves * (per + 1.5);
      if (c < 0.97 * best) { best = c; psplits = ps; }
    }
  }
  {
    const int per = (P.ntiles + psplits - 1) / psplits;
    psplits = (P.ntiles + per - 1) / per;
  }
  P.psplits = psplits;
  P.tmem_cols = pow2_cols(tg * P.acc_stride);
  // fp32 accumulate, M=128, N=bn, A and B MN-major (bits 15, 16); formats 2 = TF32 / 1 = BF16
  const uint32_t fmt = bf16 ? 1u : 2u;
  P.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  P.dwp = dwp; P.error_flag = error_flag;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 512;
  const CUtensorMapSwizzle sw = bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;

  alignas(64) CUtensorMap tmX, tmXlo, tmG, tmGlo;
  {
    const uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->width, (uint64_t)d->height, (uint64_t)d->batch};
    const uint64_t strides[3] = {(uint64_t)d->cin * esize, (uint64_t)d->width * d->cin * esize,
                                 (uint64_t)d->height * d->width * d->cin * esize};
    const uint32_t box[4] = {(uint32_t)blk_ch, (uint32_t)P.pitch, (uint32_t)(th + d->ksize - 1), 1};
    int rc = make_map(&tmX, x, 4, dims, strides, box, sw, bf16);
    if (rc) return rc;
    rc = make_map(&tmXlo, x_lo ? x_lo : x, 4, dims, strides, box, sw, bf16);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d->cout, (uint64_t)d->width, (uint64_t)d->height, (uint64_t)d->batch};
    const uint64_t strides[3] = {(uint64_t)d->cout * esize, (uint64_t)d->width * d->cout * esize,
                                 (uint64_t)d->height * d->width * d->cout * esize};
    const uint32_t box[4] = {(uint32_t)blk_ch, TILE_W, (uint32_t)th, 1};
    int rc = make_map(&tmG, g, 4, dims, strides, box, sw, bf16);
    if (rc) return rc;
    rc = make_map(&tmGlo, g_lo ? g_lo : g, 4, dims, strides, box, sw, bf16);
    if (rc) return rc;
  }
  dim3 grid(mblocks, P.tap_groups * nblocks, psplits);
  auto launch = [&](auto kern) -> int {
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NTHREADS_IGEMM, smem, stream>>>(tmX, tmXlo, tmG, tmGlo, P);
    return 0;
  };
  int lrc;
  if (bf16) lrc = split ? launch(conv_wgrad_kernel<true, true>) : launch(conv_wgrad_kernel<true, false>);
  else      lrc = split ? launch(conv_wgrad_kernel<false, true>) : launch(conv_wgrad_kernel<false, false>);
  if (lrc) return lrc;
  count_launch();
  return check_launch("conv_wgrad_kernel");
}

}  // namespace conv
}  // namespace pn

using namespace pn;

extern "C" int pn_conv2d_forward(const pn_conv_desc* desc, const void* x, const void* x_lo, const void* w_packed,
                                 const void* w_packed_lo, const float* bias, float* y, uint32_t* error_flag,
                                 pn_stream_t stream) {
  TraceScope ts(reinterpret_cast<cudaStream_t>(stream), "conv_igemm B%d H%d W%d Cin%d Cout%d k%d prec%d", desc ? desc->batch : 0,
                desc ? desc->height : 0, desc ? desc->width : 0, desc ? desc->cin : 0, desc ? desc->cout : 0, desc ? desc->ksize : 0,
                desc ? desc->precision : 0);
  return conv::conv_forward(desc, x, x_lo, w_packed, w_packed_lo, bias, y, error_flag, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int pn_conv2d_rows_pad(int cout) { return cout > 0 ? conv::rows_pad_of(cout) : 0; }

extern "C" int pn_conv2d_dgrad(const pn_conv_desc* desc, const void* g, const void* g_lo, const void* w_packed,
                               const void* w_packed_lo, float* gx, uint32_t* error_flag, pn_stream_t stream) {
  PN_REQUIRE(desc, PN_ERR_BAD_ARGUMENT, "pn_conv2d_dgrad: null descriptor");
  TraceScope ts(reinterpret_cast<cudaStream_t>(stream), "conv_dgrad B%d H%d W%d Cin%d Cout%d k%d prec%d", desc->batch, desc->height,
                desc->width, desc->cout, desc->cin, desc->ksize, desc->precision);
  pn_conv_desc dd = *desc;           // the convolution that runs: g [.., Cout] -> gx [.., Cin]
  dd.cin = desc->cout;
  dd.cout = desc->cin;
  return conv::conv_forward(&dd, g, g_lo, w_packed, w_packed_lo, nullptr, gx, error_flag, reinterpret_cast<cudaStream_t>(stream), true);
}

extern "C" int pn_conv2d_packed_weight_elems(int cout, int cin, int ksize, int transposed, int precision, size_t* elems) {
  PN_REQUIRE(elems && cout > 0 && cin > 0 && ksize > 0, PN_ERR_BAD_ARGUMENT, "pn_conv2d_packed_weight_elems: bad argument");
  const int rows = transposed ? cin : cout, k = transposed ? cout : cin;
  *elems = conv::packed_weight_elems(rows, k, ksize, precision);
  return PN_OK;
}

extern "C" int pn_conv2d_pack_weight(const float* w_oihw, void* w_packed, void* w_packed_lo, int cout, int cin, int ksize,
                                     int transposed, int precision, pn_stream_t stream) {
  PN_REQUIRE(w_oihw && w_packed && cout > 0 && cin > 0 && ksize > 0, PN_ERR_BAD_ARGUMENT, "pn_conv2d_pack_weight: bad argument");
  const int rows = transposed ? cin : cout, k = transposed ? cout : cin;
  const size_t total = conv::packed_weight_elems(rows, k, ksize, precision);
  const int kc = conv::kc_of(precision), rows_pad = conv::rows_pad_of(rows);
  // (a shared-memory tiled variant with contiguous source runs and 16-byte stores measured 2x SLOWER than this
  // element-per-thread gather -- 4.0 vs 1.9 ms per step over the 93 packings; profiles/r01_step_kernel_breakdown_final.txt)
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == PN_PRECISION_BF16X3 || precision == PN_PRECISION_BF16X1)
    conv::pack_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w_oihw, static_cast<__nv_bfloat16*>(w_packed),
                                                                    static_cast<__nv_bfloat16*>(w_packed_lo), cout, cin, ksize, kc,
                                                                    rows_pad, transposed);
  else
    conv::pack_weight_kernel<float><<<blocks, 256, 0, st>>>(w_oihw, static_cast<float*>(w_packed), static_cast<float*>(w_packed_lo),
                                                            cout, cin, ksize, kc, rows_pad, transposed);
  count_launch();
  return check_launch("pack_weight_kernel");
}

extern "C" int pn_tf32_residual(const float* x, float* lo, size_t n, pn_stream_t stream) {
  PN_REQUIRE(x && lo && (n % 4 == 0) && aligned16(x) && aligned16(lo), PN_ERR_BAD_ARGUMENT,
             "pn_tf32_residual: need 16-byte aligned pointers and n %% 4 == 0");
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  conv::tf32_residual_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(lo), n / 4);
  count_launch();
  return check_launch("tf32_residual_kernel");
}

extern "C" int pn_split_bf16(const float* x, void* hi, void* lo, size_t n, pn_stream_t stream) {
  PN_REQUIRE(x && hi && lo && (n % 4 == 0) && aligned16(x) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(lo) & 7) == 0,
             PN_ERR_BAD_ARGUMENT, "pn_split_bf16: need aligned pointers and n %% 4 == 0");
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  conv::split_bf16_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), static_cast<uint2*>(hi), static_cast<uint2*>(lo), n / 4);
  count_launch();
  return check_launch("split_bf16_kernel");
}

extern "C" int pn_conv2d_wgrad(const pn_conv_desc* desc, const void* x, const void* x_lo, const void* g, const void* g_lo,
                               float* dw_packed, uint32_t* error_flag, pn_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PN_REQUIRE(desc && dw_packed, PN_ERR_BAD_ARGUMENT, "pn_conv2d_wgrad: null argument");
  TraceScope ts(stream, "conv_wgrad B%d H%d W%d Cin%d Cout%d k%d prec%d", desc->batch, desc->height, desc->width, desc->cin,
                desc->cout, desc->ksize, desc->precision);
  const bool split = desc->precision == PN_PRECISION_TF32X3 || desc->precision == PN_PRECISION_BF16X3;
  PN_REQUIRE(!split || (x_lo && g_lo), PN_ERR_BAD_ARGUMENT, "pn_conv2d_wgrad: the x3 precisions need the residual operands");
  const size_t n = (size_t)desc->cout * desc->ksize * desc->ksize * conv::kpad_of(desc->cin, desc->precision);
  PN_CUDA(cudaMemsetAsync(dw_packed, 0, sizeof(float) * n, stream));
  return conv::conv_wgrad(desc, x, x_lo, g, g_lo, dw_packed, error_flag, stream);
}

extern "C" int pn_conv2d_wgrad_packed_elems(int cout, int cin, int ksize, int precision, size_t* elems) {
  PN_REQUIRE(elems && cout > 0 && cin > 0 && ksize > 0, PN_ERR_BAD_ARGUMENT, "pn_conv2d_wgrad_packed_elems: bad argument");
  *elems = (size_t)cout * ksize * ksize * conv::kpad_of(cin, precision);
  return PN_OK;
}

extern "C" int pn_conv2d_unpack_weight_grad(const float* dw_packed, float* dw_oihw, int cout, int cin, int ksize, int precision,
                                            pn_stream_t stream) {
  PN_REQUIRE(dw_packed && dw_oihw && cout > 0 && cin > 0 && ksize > 0, PN_ERR_BAD_ARGUMENT, "pn_conv2d_unpack_weight_grad: bad argument");
  const size_t total = (size_t)cout * cin * ksize * ksize;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  conv::unpack_weight_grad_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dw_packed, dw_oihw, cout, cin,
                                                                                            ksize * ksize, conv::kpad_of(cin, precision));
  count_launch();
  return check_launch("unpack_weight_grad_kernel");
}
