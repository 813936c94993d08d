"""Fused loss forward+backward at the bench shape, a few times (target of `ncu -k regex:loss_tile_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import synthetic  # noqa: E402
from packnet_sfm_b200.geometry import Pose  # noqa: E402
from packnet_sfm_b200.losses import MultiViewPhotometricLoss  # noqa: E402
from packnet_sfm_b200.models import YACS_LOSS_DEFAULTS  # noqa: E402

B, H, W = 4, 192, 640
dev = torch.device("cuda:0")
fr = synthetic.make_frames(B, H, W, seed=5)
inv = [d.to(dev).requires_grad_(True) for d in synthetic.make_inv_depths(B, H, W, seed=6)]
vec = synthetic.make_pose_vecs(B, seed=7).to(dev)
mats = [Pose.from_vec(vec[:, j], "euler").mat.requires_grad_(True) for j in range(2)]
loss_fn = MultiViewPhotometricLoss(**YACS_LOSS_DEFAULTS)
img, ctx, K = fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]], fr["intrinsics"].to(dev)
for _ in range(3):
    out = loss_fn(img, ctx, inv, K, K, [Pose(m) for m in mats])
    torch.autograd.grad(out["loss"], inv + mats)
torch.cuda.synchronize()
print("loss", float(out["loss"]))
