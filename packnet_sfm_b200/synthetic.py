"""Seeded synthetic KITTI-shaped inputs: low-pass-filtered frames, shifted+noised
context frames, KITTI-normalised pinhole intrinsics, smooth inverse-depth fields and small poses.

Used by tests, bench.py, the golden-vector generator and __graft_entry__.smoke().  CPU tensors; the
caller moves them to the device.  There is no dataset (and no network) in this environment."""
import math

import torch
import torch.nn.functional as F


def _lowpass(x, k=5):
    return F.avg_pool2d(F.pad(x, [k // 2] * 4, mode="replicate"), k, 1)


def make_frames(B, H, W, seed=1234, num_context=2):
    """-> dict(rgb [B,3,H,W], rgb_context list of [B,3,H,W], intrinsics [B,3,3])"""
    g = torch.Generator().manual_seed(seed)
    rgb = _lowpass(torch.rand(B, 3, H, W, generator=g))
    ctx = []
    for j in range(num_context):
        shift = 3 if j % 2 == 0 else -3
        noise = torch.rand(B, 3, H, W, generator=g)
        ctx.append((torch.roll(rgb, shift, dims=3) + 0.02 * noise).clamp(0.0, 1.0))
    K = torch.tensor([[0.58 * W, 0.0, 0.5 * W], [0.0, 1.92 * H, 0.5 * H], [0.0, 0.0, 1.0]])
    return {"rgb": rgb, "rgb_context": ctx, "intrinsics": K.unsqueeze(0).repeat(B, 1, 1).contiguous()}


def make_inv_depths(B, H, W, seed=4321, num_scales=4, full_res=True):
    """Smooth inverse-depth fields in [0.02, 1.0] (depth 1-50 m).  full_res=True mimics
    SfmModel's upsample_depth_maps=True (4 maps at HxW, scales 1..3 nearest-upsampled)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(num_scales):
        h, w = H >> i, W >> i
        base = torch.rand(B, 1, max(h // 8, 2), max(w // 8, 2), generator=g)
        d = F.interpolate(base, size=(h, w), mode="bilinear", align_corners=True)
        d = 0.02 + 0.98 * d
        if full_res and i > 0:
            d = F.interpolate(d, size=(H, W), mode="nearest")
        out.append(d.contiguous())
    return out


def make_pose_vecs(B, num_context=2, seed=99):
    """[B, num_context, 6]: translation U(-0.1,0.1) m, rotation U(-0.01,0.01) rad (PoseNet.py:82 scale)."""
    g = torch.Generator().manual_seed(seed)
    t = (torch.rand(B, num_context, 3, generator=g) * 2 - 1) * 0.1
    r = (torch.rand(B, num_context, 3, generator=g) * 2 - 1) * 0.01
    return torch.cat([t, r], dim=2)
