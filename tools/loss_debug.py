"""Gradient bring-up for the fused loss: per-scale errors vs the CPU oracle, split by term and location."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_num_threads(16)
from oracle import loss_oracle as LO  # noqa: E402
from packnet_sfm_b200 import synthetic  # noqa: E402
from packnet_sfm_b200.geometry import Pose  # noqa: E402
from packnet_sfm_b200.losses import MultiViewPhotometricLoss  # noqa: E402

dev = torch.device("cuda:0")


def run(tag, B, H, W, full_res, **kw):
    fr = synthetic.make_frames(B, H, W, seed=11)
    inv = synthetic.make_inv_depths(B, H, W, seed=12, full_res=full_res)
    vec = synthetic.make_pose_vecs(B, seed=13)
    mats = [LO.pose_from_vec(vec[:, j]) for j in range(2)]
    cfg = dict(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001, photometric_reduce_op="min", clip_loss=0.0,
               automask_loss=True)
    cfg.update(kw)
    K = fr["intrinsics"]
    inv_c = [d.clone().requires_grad_(True) for d in inv]
    mats_c = [m.clone().requires_grad_(True) for m in mats]
    ocfg = {k: v for k, v in cfg.items()}
    ref = LO.multiview_photometric_loss(fr["rgb"], fr["rgb_context"], inv_c, K, K, mats_c, **ocfg)
    ref["loss"].backward()
    inv_d = [d.to(dev).requires_grad_(True) for d in inv]
    mats_d = [m.to(dev).requires_grad_(True) for m in mats]
    out = MultiViewPhotometricLoss(**cfg)(fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]], inv_d, K.to(dev), K.to(dev),
                                         [Pose(m) for m in mats_d])
    out["loss"].backward()
    torch.cuda.synchronize()
    print("== %s: loss cuda %.8f oracle %.8f" % (tag, out["loss"].item(), ref["loss"].item()))
    for i, (a, b) in enumerate(zip(inv_d, inv_c)):
        ga, gb = a.grad.cpu(), b.grad
        err = (ga - gb)
        rel = float(err.norm() / gb.norm())
        idx = int(err.abs().argmax())
        hw = ga.shape[-1] * ga.shape[-2]
        bb, rem = idx // hw, idx % hw
        y, x = rem // ga.shape[-1], rem % ga.shape[-1]
        nbad = int((err.abs() > 1e-3 * gb.abs().max()).sum())
        print("   scale %d: rel_l2 %.3e  max|err| %.3e at (b=%d,y=%d,x=%d) got %.4e want %.4e  |g|max %.3e  bad px %d/%d  mean err %.3e" % (
            i, rel, float(err.abs().max()), bb, y, x, float(ga.flatten()[idx]), float(gb.flatten()[idx]), float(gb.abs().max()),
            nbad, err.numel(), float(err.mean())))
    for j, (a, b) in enumerate(zip(mats_d, mats_c)):
        print("   pose %d: rel_l2 %.3e" % (j, float((a.grad.cpu() - b.grad).norm() / b.grad.norm())))


run("default fullres", 1, 32, 64, True)
run("smooth=10 fullres", 1, 32, 64, True, smooth_loss_weight=10.0)
run("B3 17x33", 3, 17, 33, True)
run("B3 17x33 smooth=0", 3, 17, 33, True, smooth_loss_weight=0.0)
run("B3 17x33 smooth=10", 3, 17, 33, True, smooth_loss_weight=10.0)
run("B3 32x64 smooth=10", 3, 32, 64, True, smooth_loss_weight=10.0)
run("B1 17x33 smooth=10", 1, 17, 33, True, smooth_loss_weight=10.0)
run("bench shape", 4, 192, 640, True)
