/*
 * packnet_b200.h -- C-ABI of libpacknet_b200.so: the B200-native (sm_100a) kernels behind the
 * PackNet-SfM self-supervised hot path.
 *
 * The reference (TRI-ML/packnet-sfm) has no FFI of its own: it is 100 % Python and reaches the GPU only
 * through PyTorch library ops.  Each entry point below therefore cites the
 * reference *call site(s)* whose library ops it replaces (paths relative to /root/reference).
 *
 * Contract:
 *   - plain pointers and sizes, no torch types; every pointer is DEVICE memory unless marked "host".
 *   - all tensors fp32, contiguous, 16-byte aligned.  Images are NCHW; feature maps of the conv path are
 *     NHWC (== torch.channels_last storage).
 *   - the library never allocates, frees or retains device memory; scratch comes from the caller
 *     (pn_*_workspace_bytes) and work is only enqueued on the caller's stream.
 *   - return 0 on success, <0 for an argument/shape error, >0 = cudaError_t.  pn_last_error_string()
 *     (thread-local) describes the last failure.  No exceptions cross the ABI, no CPU fallback exists.
 *   - re-entrant: forward runs on the Python main thread, backward on PyTorch's autograd thread.
 */
#ifndef PACKNET_B200_H
#define PACKNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pn_stream_t; /* == cudaStream_t */

#define PN_MAX_SCALES 4
#define PN_MAX_CONTEXT 4

enum {
  PN_OK = 0,
  PN_ERR_BAD_ARGUMENT = -1,
  PN_ERR_UNSUPPORTED = -2,
  PN_ERR_WORKSPACE = -3,
  PN_ERR_ALIGNMENT = -4
};

/* Library / ABI version (major*10000 + minor*100 + patch) and last error text of the calling thread. */
int pn_version(void);
const char* pn_last_error_string(void);
/* Number of kernel launches this library has enqueued so far (process-wide, monotonic). */
uint64_t pn_launch_count(void);
/* Process-wide tuning switches for STAGED kernel variants (same results; each is off until measured on a B200).
 *   PN_TUNE_STAGE_FLAT (0/1, initial value from the environment variable PN_STAGE_FLAT): the feature-stencil and head-convolution
 *   kernels stage their input tiles with all threads of the CTA (stage_tile_flat) instead of one cell per warp iteration.
 *   PN_TUNE_GN_TREE (0/1, environment PN_GN_TREE): the GroupNorm statistics kernel reduces its per-thread partial sums with warp
 *   shuffles before one shared atomic per (warp, group) instead of eight fp64 shared atomics per thread. */
#define PN_TUNE_STAGE_FLAT 1
#define PN_TUNE_GN_TREE 2
int pn_set_tuning(int key, int value);

/* ------------------------------------------------------------------------------------------------
 * Photometric view-synthesis loss
 *   replaces MultiViewPhotometricLoss.forward and everything it calls:
 *     packnet_sfm/losses/multiview_photometric_loss.py:127-165 (warp_ref_image), :14-53,:169-186 (SSIM),
 *     :188-223 (calc_photometric_loss), :225-253 (reduce), :257-283 (smoothness), :287-344 (forward)
 *     packnet_sfm/geometry/camera.py:71-80,84-108,112-148,150-191; camera_utils.py:16-22,27-59
 *     packnet_sfm/geometry/pose.py:80-86; utils/image.py:85-113,178-214; utils/depth.py:103-120,146-198
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t batch;                  /* B */
  int32_t height, width;          /* H, W of `image` / `context` */
  int32_t num_context;            /* N, 1..PN_MAX_CONTEXT */
  int32_t num_scales;             /* n, 1..PN_MAX_SCALES (after progressive scaling) */
  int32_t scale_h[PN_MAX_SCALES]; /* inverse-depth map sizes; == H,W for every scale when the */
  int32_t scale_w[PN_MAX_SCALES]; /*   caller upsampled them (SfmModel.py:87-88)              */
  float ssim_loss_weight;         /* 0 disables SSIM (pure L1, no channel mean: see :204-216)  */
  float smooth_loss_weight;       /* 0 disables the smoothness term */
  float C1, C2;                   /* SSIM constants */
  int32_t reduce_min;             /* 1: photometric_reduce_op='min', 0: 'mean' */
  int32_t automask;               /* 1: add the un-warped candidates (requires reduce_min) */
  int32_t flags;                  /* PN_LOSS_FLAG_*; 0 = the default tile program */
  int32_t inv_shift[PN_MAX_SCALES]; /* s > 0: inv_depths[i] is stored at (scale_h >> s) x (scale_w >> s) and read through nearest
                                     up-sampling (index >> s) -- SfmModel's upsample_output(mode='nearest'), models/model_utils.py:
                                     152-180 with SfmModel.py:87-88, without the full-resolution copies; grad_inv_depths[i] then has
                                     the stored shape and receives the sum over each 2^s x 2^s block.  0 = map at the scale's size. */
} pn_loss_desc;

/* flags: run the grouped-scale tile program (csrc/loss_group_kernel.cuh): one CTA carries a tile through every scale that
 * shares its image size, so that the target statistics, the un-warped candidates and the edge weights are computed once.
 * Same inputs, outputs, workspace and semantics; forward and backward may use different values of the flag. */
#define PN_LOSS_FLAG_GROUPED 1

/* Scratch size for one forward(+backward) pair with this descriptor. */
int pn_loss_workspace_bytes(const pn_loss_desc* desc, size_t* bytes);

/* Forward.
 *   image          [B,3,H,W]
 *   context        host array of N pointers, each [B,3,H,W]
 *   inv_depths     host array of n pointers, each [B,1,h_i,w_i]
 *   K, ref_K       [B,3,3]
 *   poses          host array of N pointers, each [B,4,4] (target -> context transform, Pose.mat)
 *   out            [4] floats: loss, metric photometric_loss, metric smoothness_loss, reserved
 *                  (photometric_loss carries the reference's in-place alias quirk: it includes the
 *                   smoothness term whenever smooth_loss_weight > 0; DESIGN.md "quirks")
 *   workspace      >= pn_loss_workspace_bytes; must be kept untouched until the matching backward ran
 */
int pn_loss_forward(const pn_loss_desc* desc, const float* image, const float* const* context,
                    const float* const* inv_depths, const float* K, const float* ref_K,
                    const float* const* poses, float* out, void* workspace, size_t workspace_bytes,
                    pn_stream_t stream);

/* Backward.  grad_out: device scalar dL/d(loss).  grad_inv_depths[i]: [B,1,h_i,w_i] (overwritten);
 * grad_poses[j]: [B,4,4] (overwritten; bottom row zero).  Same inputs and workspace as the forward. */
int pn_loss_backward(const pn_loss_desc* desc, const float* image, const float* const* context,
                     const float* const* inv_depths, const float* K, const float* ref_K,
                     const float* const* poses, const float* grad_out, float* const* grad_inv_depths,
                     float* const* grad_poses, void* workspace, size_t workspace_bytes, pn_stream_t stream);

/* Training call (grouped program only: desc->flags & PN_LOSS_FLAG_GROUPED): ONE tile launch produces the loss (`out`, as
 * pn_loss_forward) AND the gradients for dL/dloss = 1 -- MultiViewPhotometricLoss.forward followed by loss.backward()
 * (models/model_wrapper.py:193-199 -> losses/multiview_photometric_loss.py:287-344) without evaluating the tile program twice.
 *   unit_grad_inv_depths[i]  [B,1,h_i >> s_i,w_i >> s_i], unit_grad_poses[j] [B,4,4]: overwritten with the UNIT gradients
 *                            (minus the per-sample smoothness constant that pn_loss_backward_finish adds)
 *   grad_span_bytes          0, or the size of ONE allocation that starts at unit_grad_inv_depths[0] and holds every unit
 *                            gradient output: the library then zeroes it with a single memset
 * The workspace must be kept untouched until pn_loss_backward_finish ran. */
int pn_loss_forward_backward(const pn_loss_desc* desc, const float* image, const float* const* context,
                             const float* const* inv_depths, const float* K, const float* ref_K,
                             const float* const* poses, float* out, float* const* unit_grad_inv_depths,
                             float* const* unit_grad_poses, size_t grad_span_bytes, void* workspace,
                             size_t workspace_bytes, pn_stream_t stream);
/* Second half of the training call (autograd's backward of the loss node): grad = grad_out * (unit gradient + smoothness
 * constant of the sample), out of place (the unit gradients stay valid: the call may be repeated). */
int pn_loss_backward_finish(const pn_loss_desc* desc, const float* grad_out, const float* const* unit_grad_inv_depths,
                            const float* const* unit_grad_poses, float* const* grad_inv_depths, float* const* grad_poses,
                            void* workspace, size_t workspace_bytes, pn_stream_t stream);

/* Test/inspection hook for the "warp pixel indices bit-exact" bar: writes, for one (scale, context),
 * the integer bilinear tap origin floor(ix), floor(iy) as int32 [B,h,w,2] and the unnormalised float
 * coordinates ix, iy as fp32 [B,h,w,2] -- produced by the SAME device function the loss kernels use. */
int pn_loss_warp_indices(const pn_loss_desc* desc, int scale, const float* inv_depth, const float* K,
                         const float* ref_K, const float* pose, int32_t* tap_xy, float* coord_xy,
                         void* workspace, size_t workspace_bytes, pn_stream_t stream);

/* Bilinear resize with align_corners=True of an NCHW image (match_scales, utils/image.py:178-214 ->
 * F.interpolate).  src [B,C,h_in,w_in] -> dst [B,C,h_out,w_out]. */
int pn_resize_bilinear_ac(const float* src, float* dst, int batch, int channels, int h_in, int w_in,
                          int h_out, int w_out, pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stride-1 "same" 2-D convolution on the tcgen05 tensor cores (implicit GEMM, NHWC activations)
 *   replaces nn.Conv2d inside Conv2D / ResidualConv / PackLayerConv3d / UnpackLayerConv3d:
 *     packnet_sfm/networks/layers/packnet/layers01.py:28-30,36 (ConstantPad2d + Conv2d), :58-60,:69-71,
 *     :234 (pack conv on the 8x-inflated channel count), :272 (unpack conv) and their autograd backward.
 * ------------------------------------------------------------------------------------------------ */
/* Tensor-core precision of the GEMMs (fp32 accumulate in TMEM in every mode):
 *   TF32X1  one kind::tf32 MMA per product on the fp32 operands (what cuDNN gives the reference on Ampere+ with
 *           PyTorch's default cudnn.allow_tf32=True); depth maps off by ~1e-2 max-relative vs fp32
 *   TF32X3  error-compensated split, operands fp32 x and x - trunc_tf32(x)            (~22 mantissa bits)
 *   BF16X3  error-compensated split, operands bf16 hi = rn(x), lo = rn(x - hi)         (16 mantissa bits; kind::f16
 *           runs at twice the tf32 rate on half the operand bytes; PackNet01 depth maps stay within ~2.5e-4
 *           max-relative of fp32 -- DESIGN.md "precision") -- the default of the Python layer
 *   BF16X1  one bf16 MMA per product (8 mantissa bits; for experiments only) */
enum { PN_PRECISION_TF32X1 = 1, PN_PRECISION_BF16X1 = 2, PN_PRECISION_TF32X3 = 3, PN_PRECISION_BF16X3 = 4 };
enum { PN_CONV_MODE_AUTO = 0, PN_CONV_MODE_PER_TAP = 1, PN_CONV_MODE_HALO = 2 };

typedef struct {
  int32_t batch, height, width; /* input and output spatial size (stride 1, zero padding ksize/2) */
  int32_t cin, cout;            /* cin: multiple of 4 (fp32 operands) or 8 (bf16 operands); cout: multiple of 4 */
  int32_t ksize;                /* 1, 3, 5 or 7 */
  int32_t precision;            /* PN_PRECISION_* */
  int32_t mode;                 /* PN_CONV_MODE_*: how the activation operand is staged (AUTO picks) */
  int32_t debug_flags;          /* bring-up knobs, 0 in production */
} pn_conv_desc;

/* y[B,H,W,Cout] (fp32) = conv(x[B,H,W,Cin], w) + bias.  Operand element type follows desc->precision: fp32 for the
 * TF32 modes, bf16 for the BF16 modes.  w_packed comes from pn_conv2d_pack_weight; x_lo / w_packed_lo are the
 * residual operands (pn_tf32_residual / pn_split_bf16 / pack_weight), required by the X3 modes, ignored otherwise.
 * error_flag: optional (pinned host or device) word set to 0xDEADxxxx if a pipeline wait times out (the kernel
 * traps instead of hanging). */
int pn_conv2d_forward(const pn_conv_desc* desc, const void* x, const void* x_lo, const void* w_packed,
                      const void* w_packed_lo, const float* bias, float* y, uint32_t* error_flag,
                      pn_stream_t stream);

/* Data gradient (autograd backward of nn.Conv2d w.r.t. its input, layers01.py:28-30) from the layer's FORWARD packing:
 *   gx[B,H,W,Cin] = sum_{tap,co} g[pixel - tap + pad, co] * w[co, ci, tap]
 * desc describes the LAYER (cin = channels of gx, cout = channels of g); w_packed / w_packed_lo are what pn_conv2d_forward
 * took (pn_conv2d_pack_weight with transposed=0, or the tiles pn_adam_step writes).  bf16 precisions only: the [co][64 ci]
 * tiles are read as MN-major tensor-core operands, so the transposed packing (transposed=1) is never built. */
int pn_conv2d_dgrad(const pn_conv_desc* desc, const void* g, const void* g_lo, const void* w_packed, const void* w_packed_lo,
                    float* gx, uint32_t* error_flag, pn_stream_t stream);

/* Weight packing: OIHW fp32 [Cout,Cin,k,k] (nn.Conv2d.weight) -> the shared-memory image of every B-operand tile,
 * [channel chunk][tap][rows padded to the N tile][128 bytes, SWIZZLE_128B applied], so that the kernel fetches a
 * tile with one contiguous bulk copy.  transposed=0: rows = Cout (fprop); transposed=1: rows = Cin, taps flipped
 * (the operand of the data-gradient convolution).  A channel chunk is 32 (fp32) or 64 (bf16) reduction channels.
 * w_packed_lo (residual operand of the X3 modes) may be NULL. */
int pn_conv2d_packed_weight_elems(int cout, int cin, int ksize, int transposed, int precision, size_t* elems);
int pn_conv2d_pack_weight(const float* w_oihw, void* w_packed, void* w_packed_lo, int cout, int cin, int ksize,
                          int transposed, int precision, pn_stream_t stream);
/* STAGED alternative of pn_conv2d_pack_weight for the bf16 precisions (same bytes out): the re-layout through shared memory,
 * one CTA per (8 rows, 64-wide reduction chunk).  rows_pad = pn_conv2d_packed_weight_elems / (chunks * k*k * 64). */
int pn_conv2d_pack_weight_tiled(const float* w_oihw, void* w_packed, void* w_packed_lo, int cout, int cin, int ksize,
                                int transposed, int rows_pad, pn_stream_t stream);

/* Weight gradient (autograd backward of nn.Conv2d w.r.t. weight).  x [B,H,W,Cin] and g = dL/dy [B,H,W,Cout]
 * are the NHWC tensors themselves (read as MN-major operand tiles; the reduction runs over pixels); x_lo / g_lo
 * are their residuals (X3 modes).  dw_packed fp32 [Cout][k*k][kpad(Cin)] (kpad = Cin rounded up to 32 / 64;
 * pn_conv2d_wgrad_packed_elems) is zeroed and accumulated;
 * pn_conv2d_unpack_weight_grad converts it to OIHW. */
int pn_conv2d_wgrad(const pn_conv_desc* desc, const void* x, const void* x_lo, const void* g, const void* g_lo,
                    float* dw_packed, uint32_t* error_flag, pn_stream_t stream);
int pn_conv2d_wgrad_packed_elems(int cout, int cin, int ksize, int precision, size_t* elems);
int pn_conv2d_unpack_weight_grad(const float* dw_packed, float* dw_oihw, int cout, int cin, int ksize, int precision,
                                 pn_stream_t stream);
/* STAGED alternative of pn_conv2d_unpack_weight_grad (same result, bit for bit): the re-layout through shared memory, one CTA
 * per (output channel, 128 input channels).  kpad = row pitch of dw_packed in floats = pn_conv2d_wgrad_packed_elems / (cout*k*k). */
int pn_conv2d_unpack_weight_grad_tiled(const float* dw_packed, float* dw_oihw, int cout, int cin, int ksize, int kpad,
                                       pn_stream_t stream);

/* lo[i] = x[i] - trunc_tf32(x[i]) (the bits a tf32 tensor-core operand read drops); n % 4 == 0. */
int pn_tf32_residual(const float* x, float* lo, size_t n, pn_stream_t stream);
/* hi[i] = bf16_rn(x[i]), lo[i] = bf16_rn(x[i] - hi[i]); n % 4 == 0. */
int pn_split_bf16(const float* x, void* hi, void* lo, size_t n, pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step on flat buffers, fused with the weight re-layout of the convolution engine
 *   replaces torch.optim.Adam over the 'Depth' and 'Pose' parameter groups (packnet_sfm/models/model_wrapper.py:128-166)
 *   and every per-step pn_conv2d_pack_weight / pn_conv2d_unpack_weight_grad of a stored weight.
 * Layout contract (packnet_sfm_b200/optim.py builds it):
 *   params / grads / exp_avg / exp_avg_sq: fp32 [numel], numel a multiple of PN_ADAM_BLOCK; block b = elements
 *   [b*PN_ADAM_BLOCK, (b+1)*PN_ADAM_BLOCK) belongs to ONE parameter group and to at most one stored convolution weight:
 *   block_info[b] = (group << 16) | (segment index + 1, 0 = plain elements).
 *   A stored convolution weight occupies [offset, offset + cout*taps*kpad) as [cout][tap][kpad] (kpad = Cin rounded up to 64,
 *   the padding stays zero) -- the layout pn_conv2d_wgrad accumulates, so `grads + offset` IS its dw_packed -- and its
 *   forward tiles (the operand of pn_conv2d_forward AND pn_conv2d_dgrad) are (re)written at packed_hi/lo + packed_offset:
 *   [kpad/64][tap][rows_pad][128 bytes, SWIZZLE_128B], bf16 hi = rn(w), lo = rn(w - hi).  Tile rows >= cout are never
 *   written: the caller zeroes the tile buffers once.
 *   hyper (device, fp32[16]): [0] step count (incremented here), [1] beta1, [2] beta2, [3] eps, [4] 1-beta1^t, [5] sqrt(1-beta2^t)
 *   (both written here), [8+2g] lr and [9+2g] weight_decay (L2) of group g < 4.  Device-resident so that a captured CUDA graph
 *   replays with the current step count and learning rates.
 * update = 1: one Adam step (torch.optim.Adam semantics, no amsgrad) + tiles of the new values; update = 0: tiles only. */
#define PN_ADAM_BLOCK 2048
typedef struct {
  int64_t offset;          /* first element of the weight in the flat buffers (multiple of PN_ADAM_BLOCK) */
  int64_t packed_offset;   /* byte offset of its tiles in packed_hi / packed_lo (multiple of 1024) */
  int32_t cout, taps, kpad, rows_pad;
} pn_adam_conv_seg;
int pn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t numel,
                 const int32_t* block_info, const pn_adam_conv_seg* segs, float* hyper, void* packed_hi, void* packed_lo,
                 int update, pn_stream_t stream);
/* Rows of a forward tile group for `cout` output channels (cout rounded up to the N tile and to 64). */
int pn_conv2d_rows_pad(int cout);

/* ------------------------------------------------------------------------------------------------
 * Pack / unpack feature stencil: Conv3d(1->8, 3x3x3, pad 1) fused with space-to-depth (pack) or
 * depth-to-space (unpack).  NHWC.
 *   pack   (pack=1): in x[B,2h,2w,C]  -> out[B,h,w,8*4C] at channel f*4C + (4c+2i+j)
 *          replaces packing() + unsqueeze + conv3d + view, layers01.py:126-148,:241-245
 *   unpack (pack=0): in u[B,h,w,C]    -> out[B,2h,2w,2C], out[2y+i][2x+j][v/4] with v = f*C + c, i=(v%4)/2, j=v%2
 *          replaces unsqueeze + conv3d + view + PixelShuffle(2), layers01.py:281-285
 * The destination may be a wider buffer (out_cstride channels per pixel, first channel out_coffset): the
 * decoder's torch.cat (PackNet01.py:138-175) becomes a pointer offset.  out_lo (optional) receives the tf32
 * residual of every value written.  w3 = conv3d.weight [8,1,3,3,3] flattened, b3 = conv3d.bias [8].
 * ------------------------------------------------------------------------------------------------ */
int pn_feature_stencil_forward(int pack, const float* in, const float* w3, const float* b3, float* out,
                               float* out_lo, int batch, int h_low, int w_low, int channels, int out_cstride,
                               int out_coffset, pn_stream_t stream);
/* g: gradient in the forward OUTPUT layout (g_cstride/g_coffset as above); gin: gradient in the forward INPUT
 * layout (overwritten); gw3 [216], gb3 [8] (overwritten). */
int pn_feature_stencil_backward(int pack, const float* in, const float* g, const float* w3, float* gin,
                                float* gw3, float* gb3, int batch, int h_low, int w_low, int channels,
                                int g_cstride, int g_coffset, pn_stream_t stream);
/* The same, one half at a time: parts = 1 data gradient (gin), 2 weight / bias gradient (gw3, gb3), 3 both.  The halves are
 * independent launches; a caller may enqueue the weight half on another stream (autograd of nn.Conv3d w.r.t. its parameters
 * is off the critical path of the backward).  Pointers of a half that is not requested may be NULL. */
int pn_feature_stencil_backward_parts(int pack, const float* in, const float* g, const float* w3, float* gin,
                                      float* gw3, float* gb3, int batch, int h_low, int w_low, int channels,
                                      int g_cstride, int g_coffset, int parts, pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm(16, eps) + ELU on NHWC maps (layers01.py:31-32,37; with x2 != NULL the input is x + x2, the
 * residual sum of layers01.py:72).  stats: B*16*3 doubles of scratch -- [B,16,2] doubles (sum, sum of squares)
 * followed by [B,16,2] floats (mean, rstd) -- written by the forward and read by the backward.  The output may again be a channel window of a wider buffer.
 * backward scratch `bc`: 2*C*B doubles followed by 2*16*B floats.  dx_channel_sum (optional, [C]) receives
 * sum over pixels of dx = the bias gradient of the convolution that produced x (layers01.py:28), fused into pass 2.
 * ------------------------------------------------------------------------------------------------ */
int pn_groupnorm_elu_forward(const float* x, const float* x2, const float* gamma, const float* beta, float eps,
                             float* y, float* y_lo, double* stats, int batch, int hw, int channels,
                             int out_cstride, int out_coffset, pn_stream_t stream);
int pn_groupnorm_elu_backward(const float* x, const float* x2, const float* y, const float* dy,
                              const float* gamma, float eps, const double* stats, double* bc, float* dx,
                              float* dx_lo, float* dgamma, float* dbeta, float* dx_channel_sum, int batch, int hw,
                              int channels, int y_cstride, int y_coffset, int dy_cstride, int dy_coffset,
                              pn_stream_t stream);
/* The same two passes, additionally emitting the bf16 operand pair (hi = rn(v), lo = rn(v - hi), contiguous [B,HW,C]) of the
 * output for the tensor-core convolution that consumes it -- y in the forward, dx (the convolution's output gradient) in the
 * backward -- so that no separate pn_split_bf16 launch re-reads the tensor.  No channel windows. */
int pn_groupnorm_elu_forward_split(const float* x, const float* x2, const float* gamma, const float* beta, float eps, float* y,
                                   void* y_hi_bf16, void* y_lo_bf16, double* stats, int batch, int hw, int channels,
                                   pn_stream_t stream);
int pn_groupnorm_elu_backward_split(const float* x, const float* x2, const float* y, const float* dy, const float* gamma, float eps,
                                    const double* stats, double* bc, float* dx, void* dx_hi_bf16, void* dx_lo_bf16, float* dgamma,
                                    float* dbeta, float* dx_channel_sum, int batch, int hw, int channels, pn_stream_t stream);

/* out[c] = sum over pixels of g[pixel][c] (conv bias gradient), g dense [pixels, channels]. */
int pn_channel_sum(const float* g, float* out, size_t pixels, int channels, pn_stream_t stream);

/* Single-channel head convolution Conv2d(C -> 1, 3x3, zero pad 1) of InvDepth
 * (packnet_sfm/networks/layers/packnet/layers01.py:98-122; the sigmoid / min_depth scaling stays with the caller).
 * x [B,H,W,C] NHWC fp32 (C % 4 == 0), w_tap_major [9][C] (= weight[0, c, dy, dx] at [(dy*3+dx)*C + c]), y / dy [B,H,W]. */
int pn_head_conv_forward(const float* x, const float* w_tap_major, const float* bias, float* y, int batch, int height,
                         int width, int channels, pn_stream_t stream);
int pn_head_conv_backward(const float* x, const float* dy, const float* w_tap_major, float* dx, float* dw_tap_major,
                          float* dbias, int batch, int height, int width, int channels, pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pack-block weight folding (packnet_sfm_b200/folded.py): the Conv3d(1->8, 3x3x3, pad 1) and the Conv2d(8n -> Co, k x k)
 * of PackLayerConv3d (layers01.py:213-247: conv3d :236-237,:244-246, conv :234,:247) compose into one
 * Conv2d(n -> Co, (k+2) x (k+2)) of the space-to-depth tensor.  This entry point forms that weight -- and, with a
 * border window of taps against one face of the Conv3d kernel, the thin weights of the exact frame terms:
 *   out[co][c''][ea][eb] = sum_{f,dc,dy,dx} W2[co][f][c''-dc+1][ky0+ea-(dy-dy0)][kx0+eb-(dx-dx0)] * W3[f][dc][dy][dx]
 * over ky in [ky0,ky1), kx in [kx0,kx1), dy in [dy0,dy1), dx in [dx0,dx1); ea < (ky1-ky0)+(dy1-dy0)-1, eb likewise.
 *   w2  = conv.conv_base.weight [Co, 8n, k, k] read as [Co][8][n][k][k]; w3 = conv3d.weight [8,1,3,3,3]
 *   out = OIHW [Co][n][EA][EB] with the n channels in the (i, j, c) order of the space-to-depth tensor
 *         (reference order c'' = 4c + 2i + j  ->  (2i+j)*n/4 + c)
 * Supported: k = 3 or 5; each window dimension is the whole kernel (with the whole face) or the k/2 border taps
 * (with a single face).  backward: dw2 window (+)= (overwritten when accumulate == 0), dw3 [8][27] atomically
 * accumulated (caller zeroes), dS (optional, [Co][8][k][k]) is added to every depth of the dw2 window
 * (gradient of sum over depth of W2, the Conv3d-bias term).
 * ------------------------------------------------------------------------------------------------ */
enum { PN_FOLD_LAYOUT_OIHW = 0, PN_FOLD_LAYOUT_OHWI = 1 };
typedef struct {
  int32_t cout, n, ksize;
  int32_t ky0, ky1, kx0, kx1; /* tap window of W2 */
  int32_t dy0, dy1, dx0, dx1; /* face of W3 */
  int32_t layout;             /* of out / dout: OIHW [Co][n][EA][EB] (for pn_conv2d_pack_weight) or OHWI [Co][EA][EB][n]
                                 (channels last: the reduction index of the frame GEMMs is contiguous) */
} pn_fold_desc;
int pn_pack_fold_forward(const pn_fold_desc* desc, const float* w2, const float* w3, float* out, pn_stream_t stream);
int pn_pack_fold_backward(const pn_fold_desc* desc, const float* w2, const float* w3, const float* dout, const float* dS,
                          float* dw2, float* dw3, int accumulate, pn_stream_t stream);
/* All nine folds of one pack layer in one call.  outs / douts: host arrays of 9 device pointers in the order
 * main, top, bottom, left, right, tl, tr, bl, br (main OIHW [Co][n][k+2][k+2]; top/bottom [Co][m][k+2][n],
 * left/right [Co][k+2][m][n], corners [Co][m][m][n], m = k/2: OHWI).  backward: douts[0] mandatory (it overwrites dw2,
 * with dS added), NULL entries among the others are skipped; dw3 [8][27] is accumulated (caller zeroes). */
int pn_pack_fold_set_forward(int cout, int n, int ksize, const float* w2, const float* w3, float* const* outs,
                             pn_stream_t stream);
int pn_pack_fold_set_backward(int cout, int n, int ksize, const float* w2, const float* w3, const float* const* douts,
                              const float* dS, float* dw2, float* dw3, pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Frame terms of the folded pack block (packnet_sfm_b200/folded.py): the reference zero-pads BETWEEN the Conv3d and
 * the Conv2d of PackLayerConv3d (layers01.py:28-30,36 inside :247), which the single folded convolution cannot see.
 * Up to eight thin linear terms repair the frame of width m = k/2 exactly:
 *   z[b, row(a,l), col(a,l), co] += alpha * sum_{e<KE, nn<n} w[co][a][e][nn] * line[b][l+e-pad][nn]   (+ dB, see below)
 *   row = r0 + (a / A2) * ra1 + (a % A2) * ra2 + l * rl, col likewise;   a < A, l < L
 * line: a border row / column of the space-to-depth tensor, [B][L][n] with batch stride line_bstride, zero beyond both
 * ends; w: folded weight of the term (pn_pack_fold_forward, OHWI), element (co,a,e,nn) at co*w_sco + a*w_sa + e*w_se + nn.
 * bias_mode 1 adds dB[class(row)][class(col)][co] on every pixel of the term, 2 only on rows in [m, h-m), 0 never;
 * dB [2m+1][2m+1][Co] is the Conv3d-bias correction per border class (class(p) = p for p < m, m inside,
 * m+1+(p-(len-m)) for the last m).  All terms run in one launch and accumulate into z with atomics.
 * backward: dline (atomically accumulated; caller zeroes), dw (overwritten) per term; gdB accumulated (caller zeroes).
 * flags: PN_FRAME_FLAG_DW_ZEROED -- the caller zeroed every term's dw: the weight-gradient GEMMs are then split over the
 * samples and meet by atomic adds (4x the CTAs; the launches are latency bound); PN_FRAME_FLAG_NO_KSPLIT -- A/B: one CTA per
 * output tile walks the whole reduction (the round-1 schedule).
 * ------------------------------------------------------------------------------------------------ */
#define PN_FRAME_FLAG_DW_ZEROED 1
#define PN_FRAME_FLAG_NO_KSPLIT 2
typedef struct {
  const float* line;
  const float* w;
  float* dline;
  float* dw;
  int64_t line_bstride, dline_bstride;
  int64_t w_sco, w_sa, w_se;
  int32_t L, A, A2, KE, pad;
  int32_t r0, ra1, ra2, rl;
  int32_t c0, ca1, ca2, cl;
  float alpha;
  int32_t bias_mode;
} pn_frame_term;
typedef struct {
  int32_t batch, height, width; /* of z: the packed map */
  int32_t cout, n, ksize, num_terms, flags;
  pn_frame_term terms[8];
} pn_frame_desc;
int pn_pack_frame_forward(const pn_frame_desc* desc, const float* dB, float* z, pn_stream_t stream);
int pn_pack_frame_backward(const pn_frame_desc* desc, const float* gz, float* gdB, pn_stream_t stream);

/* Diagnostics: per-call device timing of the convolution / stencil entry points (CUDA events around each call).
 * pn_trace_dump writes "tag<TAB>milliseconds" lines for the calls traced since the last dump and returns their count. */
void pn_trace_enable(int on);
int pn_trace_dump(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* PACKNET_B200_H */
