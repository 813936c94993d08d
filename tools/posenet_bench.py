"""PoseNet forward + backward on the library convolutions under the precision / layout switches PyTorch offers
(PoseNet stays PyTorch host code, north_star; models.require_fp32_library_convolutions explains why fp32 is the default)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import synthetic  # noqa: E402
from packnet_sfm_b200.models import PoseNet  # noqa: E402

dev = torch.device("cuda:0")
fr = synthetic.make_frames(4, 192, 640, seed=1)
img, ctx = fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]]


def run(tag, tf32, bench, cl):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    net = PoseNet().to(dev).train()
    a, b = (img, ctx)
    if cl:
        net = net.to(memory_format=torch.channels_last)
        a, b = img.contiguous(memory_format=torch.channels_last), [c.contiguous(memory_format=torch.channels_last) for c in ctx]

    def step():
        for p in net.parameters():
            p.grad = None
        out = net(a, b)
        out.square().sum().backward()
        return out
    for _ in range(5):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %.3f ms fwd+bwd   pose[0,0] = %s" % (tag, e0.elapsed_time(e1) / 20, out[0, 0].detach().cpu().numpy().round(7)), flush=True)


run("cudnn tf32 (PyTorch default)", True, False, False)
run("cudnn fp32", False, False, False)
run("cudnn fp32 + benchmark", False, True, False)
run("cudnn fp32 + channels_last", False, False, True)
run("cudnn fp32 + channels_last + benchmark", False, True, True)
run("cudnn tf32 + channels_last + benchmark", True, True, True)
