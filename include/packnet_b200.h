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
} pn_loss_desc;

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

#ifdef __cplusplus
}
#endif
#endif /* PACKNET_B200_H */
