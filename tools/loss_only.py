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
torch.cuda.set_stream(torch.cuda.Stream(dev))     # everything on one non-default stream (graph capture below; see bench.py)
fr = synthetic.make_frames(B, H, W, seed=5)
# the product's default: maps at H, H/2, H/4, H/8 read nearest-upsampled by the kernel (a8 fused); PN_LOSS_FULLRES=1 = pre-upsampled
FULL = os.environ.get("PN_LOSS_FULLRES") == "1"
inv = [d.to(dev).requires_grad_(True) for d in synthetic.make_inv_depths(B, H, W, seed=6, full_res=FULL)]
vec = synthetic.make_pose_vecs(B, seed=7).to(dev)
mats = [Pose.from_vec(vec[:, j], "euler").mat.requires_grad_(True) for j in range(2)]
loss_fn = MultiViewPhotometricLoss(**YACS_LOSS_DEFAULTS)
img, ctx, K = fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]], fr["intrinsics"].to(dev)
for _ in range(3):
    out = loss_fn(img, ctx, inv, K, K, [Pose(m) for m in mats], nearest_upsample=not FULL)
    torch.autograd.grad(out["loss"], inv + mats)
torch.cuda.synchronize()
print("loss", float(out["loss"]))

# CUDA-event timing of the forward and the backward call (L2 flushed before each), for the tile-vs-grouped comparison:
#   python tools/loss_only.py ; PN_LOSS_GROUPED=1 python tools/loss_only.py
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


state = {}


def fwd():
    state["o"] = loss_fn(img, ctx, inv, K, K, [Pose(m) for m in mats], nearest_upsample=not FULL)


def bwd():
    torch.autograd.grad(state["o"]["loss"], inv + mats, retain_graph=True)


fwd()
t_f, t_b = timed(fwd), timed(bwd)
P_s = B * H * W * 4
print("program %s, eager calls: fwd %.4f ms, bwd %.4f ms" % ("grouped" if os.environ.get("PN_LOSS_GROUPED", "1") == "1" else "tile", t_f, t_b))
# replayed graphs: device time without the host's enqueue gaps (how the training step runs it)
fwd(); bwd()
torch.cuda.synchronize()
gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
with torch.cuda.graph(gf):
    fwd()
with torch.cuda.graph(gb, pool=gf.pool()):
    bwd()
torch.cuda.synchronize()
t_f, t_b = timed(gf.replay), timed(gb.replay)
print("program %s, graph replay: fwd %.4f ms + bwd %.4f ms = %.4f ms -> %.0f GB/s algorithmic (%.1f MB)" % (
    "grouped" if os.environ.get("PN_LOSS_GROUPED", "1") == "1" else "tile", t_f, t_b, t_f + t_b, 92 * P_s / (t_f + t_b) / 1e6, 92 * P_s / 1e6))
