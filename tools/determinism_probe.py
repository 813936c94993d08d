"""Run-to-run reproducibility of (1) the fused loss alone, (2) PackNet01 forward + backward with a fixed output gradient,
(3) the whole step's gradients -- same inputs, same parameters, twice each.  Split-K / weight-gradient atomics reorder fp32 sums
(expected: ~1e-6 relative); anything at the percent level is either the loss's per-pixel minimum amplifying that noise or a bug."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import packnet_oracle as PO  # noqa: E402  (seeded weights only)
from oracle.step_oracle import posenet_state_dict  # noqa: E402
from packnet_sfm_b200 import synthetic  # noqa: E402
from packnet_sfm_b200.geometry import Pose  # noqa: E402
from packnet_sfm_b200.losses import MultiViewPhotometricLoss  # noqa: E402
from packnet_sfm_b200.models import SelfSupModel, YACS_LOSS_DEFAULTS  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 64, 96)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


# (1) loss alone
fr = synthetic.make_frames(B, H, W, seed=91)
img, ctx, K = fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]], fr["intrinsics"].to(dev)
inv0 = [d.to(dev) for d in synthetic.make_inv_depths(B, H, W, seed=6, full_res=False)]
vec = synthetic.make_pose_vecs(B, seed=7).to(dev)
loss_fn = MultiViewPhotometricLoss(**YACS_LOSS_DEFAULTS)
runs = []
for _ in range(2):
    inv = [d.clone().requires_grad_(True) for d in inv0]
    mats = [Pose.from_vec(vec[:, j], "euler").mat.requires_grad_(True) for j in range(2)]
    out = loss_fn(img, ctx, inv, K, K, [Pose(m) for m in mats], nearest_upsample=True)
    out["loss"].backward()
    runs.append((float(out["loss"]), [d.grad.clone() for d in inv], [m.grad.clone() for m in mats]))
print("loss alone: loss %.9g / %.9g; ginv rel diff %s; gpose rel diff %s" % (
    runs[0][0], runs[1][0], ["%.1e" % rel(a, b) for a, b in zip(runs[0][1], runs[1][1])],
    ["%.1e" % rel(a, b) for a, b in zip(runs[0][2], runs[1][2])]))

# (2) network with a fixed output gradient, (3) whole step
batch = {"rgb": img, "rgb_context": ctx, "intrinsics": K, "rgb_original": img, "rgb_context_original": ctx}
g = torch.Generator().manual_seed(3)
res_net, res_step, invs = [], [], []
for r in range(2):
    model = SelfSupModel(flip_lr_prob=0.0).to(dev).train()
    model.depth_net.load_state_dict(PO.packnet01_state_dict(seed=42, randomize_affine=True))
    model.pose_net.load_state_dict(posenet_state_dict(43))
    outs = model.depth_net(img)["inv_depths"]
    if r == 0:
        gys = [(torch.rand(o.shape, generator=g) - 0.5).to(dev) for o in outs]
    torch.autograd.backward(outs, gys)
    res_net.append({k: p.grad.clone() for k, p in model.depth_net.named_parameters()})
    invs.append([o.detach().clone() for o in outs])
    for p in model.parameters():
        p.grad = None
    out = model(batch)
    out["loss"].backward()
    res_step.append((float(out["loss"]), {k: p.grad.clone() for k, p in model.named_parameters()}))
print("depth maps run to run:", ["%.1e" % rel(a, b) for a, b in zip(*invs)])
d = sorted(((rel(res_net[0][k], res_net[1][k]), k) for k in res_net[0]), reverse=True)
print("network alone (fixed gy): worst gradient rel diffs", [(("%.1e" % v), k) for v, k in d[:4]], "median %.1e" % d[len(d) // 2][0])
d = sorted(((rel(res_step[0][1][k], res_step[1][1][k]), k) for k in res_step[0][1] if float(res_step[0][1][k].norm()) > 1e-6), reverse=True)
print("whole step: loss %.9g / %.9g; worst gradient rel diffs" % (res_step[0][0], res_step[1][0]), [(("%.1e" % v), k) for v, k in d[:6]],
      "median %.1e" % d[len(d) // 2][0])
