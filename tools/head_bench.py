"""InvDepth head convolution (C -> 1, 3x3) forward / backward at the four scales of the bench shape, default tile staging
against the staged all-threads staging (pn_set_tuning PN_TUNE_STAGE_FLAT), CUDA-event medians with the L2 flushed.
Round-1 profile: head_fwd 0.32 ms + head_wgrad 0.33 ms + head_dgrad 0.09 ms per step for ~170 MB of reads."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import _lib, functional as PF  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for B, H, W, C in ((4, 192, 640, 64), (4, 96, 320, 64), (4, 48, 160, 128), (4, 24, 80, 256)):
    x = (torch.rand(B, H, W, C, device=dev) - 0.5).requires_grad_(True)
    w = ((torch.rand(1, C, 3, 3, device=dev) - 0.5) * 0.2).requires_grad_(True)
    b = (torch.rand(1, device=dev) - 0.5).requires_grad_(True)
    gy = torch.rand(B, H, W, device=dev) - 0.5
    for flat in (0, 1):
        _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, flat)
        st = {}

        def fwd():
            st["y"] = PF.head_conv(x, w, b)

        def bwd():
            torch.autograd.grad(st["y"], (x, w, b), gy, retain_graph=True)

        fwd(); bwd()
        tf, tb = timed(fwd), timed(bwd)
        mb = B * H * W * C * 4 / 1e6
        print("B%d %dx%d C%d flat=%d: fwd %.4f ms (%.0f GB/s of x), bwd %.4f ms" % (B, H, W, C, flat, tf, mb / tf, tb))
_lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 0)
