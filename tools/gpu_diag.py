"""Step-by-step GPU bring-up with timestamps (flushes after every step; dumps Python stacks if a step stalls)."""
import faulthandler
import os
import sys
import time

faulthandler.enable()
T0 = time.time()


def say(msg):
    print("[%7.1fs] %s" % (time.time() - T0, msg), flush=True)


faulthandler.dump_traceback_later(90, repeat=True)
say("start")
import torch  # noqa: E402
say("import torch done; cuda available=%s" % torch.cuda.is_available())
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
x = torch.zeros(4, device="cuda:0")
torch.cuda.synchronize()
say("cuda context up: %s" % torch.cuda.get_device_name(0))
from packnet_sfm_b200 import _lib, synthetic  # noqa: E402
lib = _lib.lib()
say("libpacknet_b200 loaded, version %d" % lib.pn_version())
from packnet_sfm_b200.losses import MultiViewPhotometricLoss, warp_tap_indices  # noqa: E402
from packnet_sfm_b200.geometry import Pose  # noqa: E402
from packnet_sfm_b200.models import YACS_LOSS_DEFAULTS  # noqa: E402
dev = torch.device("cuda:0")
B, H, W = 2, 32, 64
fr = synthetic.make_frames(B, H, W, seed=1)
inv = [d.to(dev).requires_grad_(True) for d in synthetic.make_inv_depths(B, H, W, seed=2)]
vec = synthetic.make_pose_vecs(B, seed=3).to(dev)
mats = [Pose.from_vec(vec[:, j], "euler").mat.requires_grad_(True) for j in range(2)]
K = fr["intrinsics"].to(dev)
say("inputs on device")
taps, coords = warp_tap_indices(inv[0].detach(), K, K, mats[0].detach())
torch.cuda.synchronize()
say("warp_tap_indices ok: %s" % (taps[0, 0, :2].tolist(),))
loss_fn = MultiViewPhotometricLoss(**YACS_LOSS_DEFAULTS)
out = loss_fn(fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]], inv, K, K, [Pose(m) for m in mats])
say("loss forward enqueued")
torch.cuda.synchronize()
say("loss forward done: %s" % out["loss"].item())
out["loss"].backward()
say("loss backward enqueued")
torch.cuda.synchronize()
say("loss backward done: ginv sum %s" % float(inv[0].grad.abs().sum()))
say("launch count %d" % _lib.launch_count())
